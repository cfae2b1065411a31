"""Work partitioning across the GPUs of one box (one process per GPU).

The hot path has no exchange step: pixels, images and cube rows are independent, so ranks take
disjoint shards and never talk to each other on the data path.  The only
cross-rank traffic is the barrier / max-reduction that brackets a timed region and the final
gather of results to the host.

  * conversion batches shard by image index           -> image_range()
  * a cubemap shards by (face, row band) work units    -> cube_units() / units_for_rank()
    and each rank needs only the source rows its units read -> source_window()
"""
from __future__ import annotations

from dataclasses import dataclass

from . import device

FACE_NAMES = ("right", "left", "up", "down", "front", "back")  # +X,-X,+Y,-Y,+Z,-Z (tools/cubetool/commands.hpp:31)


def image_range(rank: int, world: int, count: int) -> tuple[int, int]:
	"""Contiguous, balanced range of image indices [lo, hi) for `rank` (sizes differ by at most 1)."""
	if not (0 <= rank < world):
		raise ValueError(f"rank {rank} outside world of {world}")
	base, extra = divmod(count, world)
	lo = rank * base + min(rank, extra)
	return lo, lo + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class CubeUnit:
	face: int
	row_begin: int
	row_end: int

	@property
	def rows(self) -> int:
		return self.row_end - self.row_begin


def band_edges(face_size: int, n: int) -> list[int]:
	"""Row boundaries of n bands, symmetric about the face's centre: band k and band n-1-k are mirror
	images of each other (row j <-> face_size-1-j), which is what lets the kernel resample the two with
	one projection per four pixels (csrc/equirect.cu, pair_mirror_units)."""
	n = max(1, min(n, face_size))
	edges = [0] * (n + 1)
	for k in range(n // 2 + 1):
		edges[k] = face_size * k // n
		edges[n - k] = face_size - edges[k]
	if n % 2 == 0:
		edges[n // 2] = (face_size + 1) // 2  # the centre row of an odd face goes to the upper band
	return edges


def cube_units(face_size: int, bands_per_face: int, polar_factor: int = 1) -> list[CubeUnit]:
	"""Work units of a cube: bands_per_face row bands on each side face and polar_factor times as
	many on the +-Y faces (their rows cost up to twice as much, so they are cut finer), band-major so
	that consecutive units are different faces.  Bands are mirror-symmetric (band_edges)."""
	units = []
	for f in range(6):
		edges = band_edges(face_size, bands_per_face * (polar_factor if f in (2, 3) else 1))
		for r0, r1 in zip(edges[:-1], edges[1:]):
			if r1 > r0:
				units.append(CubeUnit(f, r0, r1))
	return sorted(units, key=lambda u: (u.row_begin / face_size, u.face))


def mirror_pairs(units: list[CubeUnit], face_size: int) -> list[tuple[CubeUnit, ...]]:
	"""Group units with their mirror band on the same face: (upper, lower) tuples, 1-tuples for bands
	that are their own mirror (the middle band of an odd count) or whose partner is not in the list."""
	by_key = {(u.face, u.row_begin, u.row_end): u for u in units}
	groups, seen = [], set()
	for u in units:
		key = (u.face, u.row_begin, u.row_end)
		if key in seen:
			continue
		seen.add(key)
		lo, hi = face_size - u.row_end, face_size - u.row_begin
		# an odd face's centre row sits in the upper band only: its partner is one row shorter
		cands = [(u.face, lo, hi), (u.face, lo + 1, hi), (u.face, lo, hi - 1)]
		mate = next((by_key[c] for c in cands if c in by_key and c not in seen), None)
		if mate is None:
			groups.append((u,))
		else:
			seen.add((mate.face, mate.row_begin, mate.row_end))
			groups.append(tuple(sorted((u, mate), key=lambda q: q.row_begin)))
	return groups


def bands_for_world(world: int) -> int:
	"""Bands per side face (the +-Y faces get twice as many): world rounded up to a power of two -- about
	four mirror pairs per rank for the cost-aware dealing below to even out the polar faces with.  Finer
	cuts balance better on paper and run slower (dev/rank_costs.py on B200, 8 ranks: 0.0887 ms with 8
	bands, 0.0906 with 16, 0.0893 with 32)."""
	if world <= 1:
		return 1
	bands = 1
	while bands < world:
		bands *= 2
	return bands if bands == world else 2 * bands  # odd rank counts need finer change to come out even


# Cost of a row of the +-Y faces relative to a side-face row, against its distance from the face centre
# (0 = centre row, 1 = edge).  Measured on B200 with the four-fold kernel (dev/pair_costs.py, 16384x8192
# RGBA32F -> 4096^2, mirror pairs of 256-row bands: 28.5 us per 1000 side-face rows; 29.2 / 32.8 / 33.0 /
# 36.9 / 40.8 / 45.4 / 58.1 / 73.1 from the edge to the centre of a +-Y face, where neighbouring pixels
# fan out over all longitudes and taps stop coalescing).
_POLAR_COST = ((0.0, 2.8), (0.0625, 2.565), (0.1875, 2.039), (0.3125, 1.593), (0.4375, 1.432), (0.5625, 1.295), (0.6875, 1.158), (0.8125, 1.151), (0.9375, 1.025), (1.0, 1.0))


def _polar_row_cost(d: float) -> float:
	for (d0, c0), (d1, c1) in zip(_POLAR_COST, _POLAR_COST[1:]):
		if d <= d1:
			return c0 + (c1 - c0) * (d - d0) / (d1 - d0)
	return 1.0


def unit_cost(u: CubeUnit, face_size: int) -> float:
	"""Relative cost of a unit: a side-face row costs 1, a row of the +-Y faces 1.0 at the face edge rising
	to about 2.7 at the centre (table above)."""
	if u.face not in (2, 3):
		return float(u.rows)
	return sum(_polar_row_cost(abs((2 * r + 1) / face_size - 1.0)) for r in range(u.row_begin, u.row_end))


def units_for_rank(rank: int, world: int, face_size: int, bands_per_face: int | None = None) -> list[CubeUnit]:
	"""Cost-balanced dealing (longest-processing-time first) of MIRROR PAIRS of bands: a band and its
	mirror image always go to the same rank, whose launch then shares one projection between four
	pixels exactly like a whole-face launch does.  On the +-Y faces the pair also reads the same
	source rows.  Pairs are sorted by estimated cost and each is handed to the currently least-loaded
	rank.  Deterministic, so every rank computes the same plan without talking to the others."""
	if not (0 <= rank < world):
		raise ValueError(f"rank {rank} outside world of {world}")
	bands = bands_for_world(world) if bands_per_face is None else bands_per_face
	# the +-Y faces are cut twice as fine: their centre bands would otherwise be the largest units by
	# a factor of two and cap the balance
	units = cube_units(face_size, bands, polar_factor=2 if world > 1 else 1)
	groups = mirror_pairs(units, face_size)
	cost = [sum(unit_cost(u, face_size) for u in g) for g in groups]
	order = sorted(range(len(groups)), key=lambda k: (-cost[k], k))
	load = [0.0] * world
	mine = []
	for k in order:
		r = min(range(world), key=lambda q: (load[q], q))
		load[r] += cost[k]
		if r == rank:
			mine.extend(groups[k])
	return sorted(mine, key=lambda u: (u.face, u.row_begin))


def units_for_rank_by_latitude(rank: int, world: int, face_size: int, W: int, H: int) -> list[CubeUnit]:
	"""Dealing for HOST-memory jobs, where the upload of the source window is what a rank pays for, not its
	kernel: 2 x world row bands per face, ordered by the latitude they read (centre of their source-row
	window) and cut into `world` contiguous runs of equal output rows -- each rank then uploads one band of
	the panorama (about a quarter of it at 8 ranks) instead of nearly all of it.  The same plan
	libcacaoimage::EquirectToCubemap(src, N, ..., devices) makes (csrc/libcacaoimage.cpp); the device-resident
	dealing above balances kernel cost instead and lets windows overlap."""
	if not (0 <= rank < world):
		raise ValueError(f"rank {rank} outside world of {world}")
	if world == 1:
		return [CubeUnit(f, 0, face_size) for f in range(6)]
	bands = min(face_size, 2 * world)
	keyed = []
	for f in range(6):
		for b in range(bands):
			r0, r1 = face_size * b // bands, face_size * (b + 1) // bands
			if r1 <= r0:
				continue
			w0, wn = device.equirect_src_window(W, H, face_size, 1 << f, r0, r1)
			keyed.append((2 * w0 + wn, len(keyed), CubeUnit(f, r0, r1)))
	keyed.sort(key=lambda k: (k[0], k[1]))
	units = [k[2] for k in keyed]
	total = sum(u.rows for u in units)
	begin = [len(units)] * (world + 1)
	begin[0], acc, d = 0, 0, 0
	for i, u in enumerate(units):
		while d + 1 < world and acc * world >= (d + 1) * total:
			d += 1
			begin[d] = i
		acc += u.rows
	while d + 1 < world:
		d += 1
		begin[d] = len(units)
	return units[begin[rank]:begin[rank + 1]]


def source_window(units: list[CubeUnit], W: int, H: int, face_size: int) -> tuple[int, int]:
	"""(first_row, rows): the band of equirect rows that covers every unit in the list
	(host arithmetic identical to the kernel's, cacao_img_equirect_src_window)."""
	lo, hi = None, None
	for u in units:
		r0, n = device.equirect_src_window(W, H, face_size, 1 << u.face, u.row_begin, u.row_end)
		if n == 0:
			continue
		lo = r0 if lo is None else min(lo, r0)
		hi = r0 + n if hi is None else max(hi, r0 + n)
	return (0, 0) if lo is None else (lo, hi - lo)


def max_over_ranks(value: float) -> float:
	"""Multi-GPU timings are the max over ranks; single-process callers get the value back."""
	try:
		import torch
		import torch.distributed as dist
	except ImportError:
		return value
	if not (dist.is_available() and dist.is_initialized()):
		return value
	dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
	t = torch.tensor([value], dtype=torch.float64, device=dev)
	dist.all_reduce(t, op=dist.ReduceOp.MAX)
	return float(t.item())
