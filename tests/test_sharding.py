"""CPU tests of the multi-GPU host logic: world_size-2 (and -3) process groups over gloo.

The data path has no collective, so what needs covering is the partitioning: every image and
every cube row is owned by exactly one rank, each rank's planned source window contains every
row its units read, shards computed independently reassemble to the single-rank result, and the
max-over-ranks timing reduction works.  The pixel work itself is done here by the CPU oracle
(test infrastructure) -- the product has no CPU path; the same partition drives the GPU ranks
in bench.py and test_multigpu_gpu.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
	s = socket.socket()
	s.bind(("127.0.0.1", 0))
	p = s.getsockname()[1]
	s.close()
	return p


def _worker(rank, world, port, out_dir):
	sys.path.insert(0, ROOT)
	import torch
	import torch.distributed as dist

	from cacaoengine_b200 import sharding
	from oracle import pyoracle as po

	dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
	try:
		# --- conversion batch sharded by image index (BASELINE config 4 in miniature)
		count, ppi = 11, 333
		lo, hi = sharding.image_range(rank, world, count)
		src = po.synth(count * ppi * 3, po.U8, 3, 0xC0FFEE + 4)
		mine = po.convert(src[lo * ppi * 3: hi * ppi * 3], 3, 4, po.U8, po.F32) if hi > lo else np.zeros(0, np.float32)
		parts = [None] * world
		dist.all_gather_object(parts, (lo, hi, mine))
		# --- cubemap sharded by (face, row band) with per-rank source windows (config 5 in miniature)
		W, H, N = 256, 128, 40
		pano = po.synth(W * H * 4, po.F32, 4, 0xC0FFEE + 5)
		units = sharding.units_for_rank(rank, world, N)
		w0, wn = sharding.source_window(units, W, H, N)
		window = pano.reshape(H, -1)[w0:w0 + wn]
		cube = np.zeros(6 * N * N * 4, np.float32)
		for u in units:
			# the rank only holds `window`: every tap row of the unit must lie inside it
			ys = np.array([po.project(u.face, i, j, N, W, H)[1] for j in range(u.row_begin, u.row_end) for i in (0, N // 2, N - 1)])
			assert w0 <= max(0, int(np.floor(ys.min()))) and min(H - 1, int(np.floor(ys.max())) + 1) < w0 + wn
			po.equirect_to_cube(pano, W, H, 4, po.F32, N, face_mask=1 << u.face, row_begin=u.row_begin, row_end=u.row_end, out=cube)
		assert window.shape[0] == wn
		cubes = [None] * world
		dist.all_gather_object(cubes, ([(u.face, u.row_begin, u.row_end) for u in units], cube))
		# --- timing reduction: max over ranks
		t = sharding.max_over_ranks(1.0 + rank)
		assert t == float(world)
		if rank == 0:
			# images: disjoint, ordered, complete
			spans = sorted((a, b) for a, b, _ in parts)
			assert spans[0][0] == 0 and spans[-1][1] == count and all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
			assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
			whole = np.concatenate([m for _, _, m in sorted(parts, key=lambda p: p[0])])
			assert (whole.view(np.uint32) == po.convert(src, 3, 4, po.U8, po.F32).view(np.uint32)).all()
			# cube rows: every (face, row) exactly once; reassembled cube == single-rank cube
			owner = np.zeros((6, N), np.int32)
			total = np.zeros_like(cube)
			for us, c in cubes:
				for f, r0, r1 in us:
					owner[f, r0:r1] += 1
					sl = slice((f * N + r0) * N * 4, (f * N + r1) * N * 4)
					total[sl] = c[sl]
			assert (owner == 1).all()
			assert (total.view(np.uint32) == po.equirect_to_cube(pano, W, H, 4, po.F32, N).view(np.uint32)).all()
			open(os.path.join(out_dir, "ok"), "w").write("ok")
		dist.barrier()
	finally:
		dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_partition_over_gloo(built_lib, oracle, tmp_path, world):
	import torch.multiprocessing as mp

	port = _free_port()
	mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
	assert (tmp_path / "ok").exists()


def test_image_range_properties():
	from cacaoengine_b200 import sharding

	for world in (1, 2, 3, 4, 8):
		for count in (0, 1, 7, 8, 256, 4096):
			spans = [sharding.image_range(r, world, count) for r in range(world)]
			assert spans[0][0] == 0 and spans[-1][1] == count
			assert all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
			sizes = [b - a for a, b in spans]
			assert max(sizes) - min(sizes) <= 1
	with pytest.raises(ValueError):
		sharding.image_range(2, 2, 10)


def test_cube_units_deal_evenly():
	"""Cost-balanced dealing: exact cover of the cube, and estimated per-rank cost within 5 % of the mean
	(the +-Y faces cost up to 2.2x per row at the pole, so equal row counts would NOT be balanced)."""
	from cacaoengine_b200 import sharding

	for world in (1, 2, 3, 4, 8):
		N = 4096
		per_rank = [sharding.units_for_rank(r, world, N) for r in range(world)]
		owner = np.zeros((6, N), np.int32)
		for us in per_rank:
			for u in us:
				owner[u.face, u.row_begin:u.row_end] += 1
		assert (owner == 1).all()
		costs = [sum(sharding.unit_cost(u, N) for u in us) for us in per_rank]
		assert max(costs) <= 1.03 * sum(costs) / world, (world, costs)
		# a band always travels with its mirror band (the launcher folds the two into one four-fold unit)
		for us in per_rank:
			have = {(u.face, u.row_begin, u.row_end) for u in us}
			assert all((u.face, N - u.row_end, N - u.row_begin) in have for u in us)
		assert all(len(us) <= 32 for us in per_rank) or world < 4  # fits one kernel launch per rank
	# determinism: every rank derives the same global plan
	assert sharding.units_for_rank(3, 8, 4096) == sharding.units_for_rank(3, 8, 4096)
	polar = sharding.CubeUnit(2, 1792, 2304)
	assert sharding.unit_cost(polar, 4096) > 2.0 * sharding.unit_cost(sharding.CubeUnit(0, 1792, 2304), 4096)


def test_band_edges_are_mirror_symmetric():
	from cacaoengine_b200 import sharding

	for N in (1, 2, 5, 40, 41, 70, 4096):
		for n in (1, 2, 3, 4, 5, 8, 16):
			e = sharding.band_edges(N, n)
			assert e[0] == 0 and e[-1] == N and all(a <= b for a, b in zip(e, e[1:]))
			# band k mirrors band n-1-k, up to the centre row of an odd face (kept by the upper band)
			for k in range(len(e) - 1):
				lo, hi = N - e[len(e) - 1 - k], N - e[len(e) - 2 - k]
				assert abs(lo - e[k]) <= (N & 1) and abs(hi - e[k + 1]) <= (N & 1)
		units = sharding.cube_units(N, 4, 2)
		owner = np.zeros((6, N), np.int32)
		for u in units:
			owner[u.face, u.row_begin:u.row_end] += 1
		assert (owner == 1).all()


def test_fold_units_keeps_the_cover(built_lib):
	"""cacao_img_equirect_fold_units (the launcher's planning step, host only): whatever the list --
	whole faces, mirror pairs, partial overlaps, odd faces, rows named twice -- every (face, row) is
	computed exactly as often as the caller listed it, mirror units stay inside the upper half, and
	slivers are not folded."""
	from cacaoengine_b200 import device, sharding

	M = device.UNIT_MIRROR

	def cover(N, units, folded=False):
		own = np.zeros((6, N), np.int32)
		for f, r0, r1 in units:
			if folded and f & M:
				assert r1 <= (N + 1) // 2 and r1 - r0 >= min(16, (N + 1) // 2)
				own[f & 7, r0:r1] += 1
				for j in range(r0, r1):
					if N - 1 - j != j:
						own[f & 7, N - 1 - j] += 1
			else:
				assert f <= 5
				own[f, r0:min(r1, N)] += 1
		return own

	for N in (1, 2, 9, 64, 70, 71):
		whole = [(f, 0, N) for f in range(6)]
		folded = device.equirect_fold_units(N, whole)
		assert folded == [(f | M, 0, (N + 1) // 2) for f in range(6)]
		assert (cover(N, folded, True) == 1).all()
	assert device.equirect_fold_units(70, [(0, 0, 70), (1, 0, 35), (2, 36, 70), (4, 3, 40)]) == [(M, 0, 35), (1, 0, 35), (2, 36, 70), (4, 3, 40)]
	assert device.equirect_fold_units(71, [(2, 30, 40), (3, 35, 36)]) == [(2, 30, 40), (3, 35, 36)]
	assert device.equirect_fold_units(64, [(7, 0, 3), (1, 5, 5), (9, 0, 64)]) == []
	# a sharded rank's list folds completely: half as many units, all mirrored
	for world in (2, 4, 8):
		for rank in range(world):
			us = sharding.units_for_rank(rank, world, 4096)
			folded = device.equirect_fold_units(4096, us)
			assert all(f & M for f, _, _ in folded) and sum(r1 - r0 for _, r0, r1 in folded) * 2 == sum(u.rows for u in us)
			assert (cover(4096, folded, True) == cover(4096, [(u.face, u.row_begin, u.row_end) for u in us])).all()
	rng = np.random.default_rng(7)
	for trial in range(300):
		N = int(rng.choice([1, 2, 3, 17, 40, 41, 100, 129]))
		units = []
		for _ in range(int(rng.integers(0, 9))):
			a, b = sorted(int(v) for v in rng.integers(0, N + 1, 2))
			units.append((int(rng.integers(0, 6)), a, b + int(rng.integers(0, 2)) * (b == N)))  # now and then past the face's end
		if rng.random() < 0.3 and units:  # mirror of an existing unit
			f, a, b = units[0]
			units.append((f, N - min(b, N), N - a))
		folded = device.equirect_fold_units(N, units)
		assert (cover(N, folded, True) == cover(N, [u for u in units if u[1] < min(u[2], N)])).all(), (N, units, folded)


def test_units_for_rank_random_sizes(built_lib):
	"""Any face size and rank count: the ranks' lists tile the cube exactly once, every band sits with its
	mirror band on the same rank (up to the centre row of an odd face), the launcher folds a rank's list
	into four-fold units without changing its cover, and the plan is the same on every rank."""
	from cacaoengine_b200 import device, sharding

	M = device.UNIT_MIRROR
	rng = np.random.default_rng(2026)
	for _ in range(40):
		N = int(rng.choice([1, 2, 3, 7, 16, 31, 64, 100, 257, 512, 1023, 2048]))
		world = int(rng.integers(1, 9))
		owner = np.zeros((6, N), np.int32)
		for rank in range(world):
			us = sharding.units_for_rank(rank, world, N)
			assert us == sharding.units_for_rank(rank, world, N)
			mine = np.zeros((6, N), np.int32)
			for u in us:
				assert 0 <= u.face < 6 and 0 <= u.row_begin < u.row_end <= N
				mine[u.face, u.row_begin:u.row_end] += 1
			owner += mine
			# mirror closure: row j is here iff row N-1-j is
			assert (mine == mine[:, ::-1]).all(), (N, world, rank)
			folded = device.equirect_fold_units(N, us)
			cover = np.zeros((6, N), np.int32)
			for f, r0, r1 in folded:
				cover[f & 7, r0:r1] += 1
				if f & M:
					for j in range(r0, r1):
						if N - 1 - j != j:
							cover[f & 7, N - 1 - j] += 1
			assert (cover == mine).all(), (N, world, rank, folded)
		assert (owner == 1).all(), (N, world)


def test_latitude_dealing_for_host_jobs(built_lib):
	"""units_for_rank_by_latitude (host-memory jobs: a rank pays for the source rows it uploads): the ranks'
	lists tile the cube exactly once for any face size and rank count, every rank gets about the same number of
	output rows, and at 8 ranks a rank's source window is a band of the panorama -- well under half of it --
	where the cost-balanced device dealing reads most of it."""
	from cacaoengine_b200 import sharding

	rng = np.random.default_rng(7)
	for _ in range(20):
		N = int(rng.choice([1, 3, 16, 31, 100, 257, 1024]))
		world = int(rng.integers(1, 9))
		W, H = 4 * max(N, 2), 2 * max(N, 2)
		owner = np.zeros((6, N), np.int32)
		rows = []
		for rank in range(world):
			us = sharding.units_for_rank_by_latitude(rank, world, N, W, H)
			assert us == sharding.units_for_rank_by_latitude(rank, world, N, W, H)
			for u in us:
				owner[u.face, u.row_begin:u.row_end] += 1
			rows.append(sum(u.rows for u in us))
		assert (owner == 1).all(), (N, world)
		if N >= 100:
			assert max(rows) - min(rows) <= 2 * (N // (2 * world) + 1), (N, world, rows)
	W, H, N = 16384, 8192, 4096
	by_lat = [sharding.source_window(sharding.units_for_rank_by_latitude(r, 8, N, W, H), W, H, N)[1] for r in range(8)]
	by_cost = [sharding.source_window(sharding.units_for_rank(r, 8, N), W, H, N)[1] for r in range(8)]
	assert max(by_lat) < 0.45 * H and sum(by_lat) < 0.5 * sum(by_cost), (by_lat, by_cost)
