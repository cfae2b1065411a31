"""Multi-GPU parity: shards computed on different devices through the C ABI's `device` argument reassemble
bit for bit to the single-device result.  Partitioning comes from cacaoengine_b200.sharding, the same code
bench.py's ranks use.  The tests always run with at least three logical ranks; on a box with fewer devices
the ranks go round robin over the devices there are (two contexts' worth of work driven on device 0), so
the host-side orchestration -- dealing, source windows, band reassembly, one thread per rank in the C++
layer -- is exercised on a one-GPU box as well and nothing skips."""
import numpy as np
import pytest

from gpu_util import same_bits

pytestmark = pytest.mark.gpu


def _devices():
	from cacaoengine_b200 import device

	return device.device_count()


def _ranks():
	"""(logical ranks, physical device of each)"""
	have = _devices()
	n = max(3, have)
	return n, [r % have for r in range(n)]


def test_sharded_conversion_matches_single_device(built_lib, oracle):
	from cacaoengine_b200 import _capi as c, device, sharding

	n, phys = _ranks()
	count, ppi = 13, 40000
	src = oracle.synth(count * ppi * 3, c.FMT_U8, 3, 21)
	want = oracle.convert(src, 3, 4, c.FMT_U8, c.FMT_F32)
	got = np.empty_like(want)
	for r in range(n):
		lo, hi = sharding.image_range(r, n, count)
		if hi == lo:
			continue
		d = phys[r]
		device.init(d)
		a, b = device.malloc((hi - lo) * ppi * 3, d), device.malloc((hi - lo) * ppi * 16, d)
		device.upload(a, src[lo * ppi * 3: hi * ppi * 3], device=d)
		device.convert_batch(a, b, ppi, hi - lo, 3, 4, c.FMT_U8, c.FMT_F32, device=d)
		device.download(got[lo * ppi * 4: hi * ppi * 4], b, device=d)
		device.free(a, d), device.free(b, d)
	assert same_bits(got, want)


def test_sharded_cubemap_matches_single_device(built_lib, oracle):
	from cacaoengine_b200 import _capi as c, device, sharding

	n, phys = _ranks()
	W, H, N, ch, fmt = 1024, 512, 256, 4, c.FMT_F32
	pano = oracle.synth(W * H * ch, fmt, ch, 22)
	want = oracle.equirect_to_cube(pano, W, H, ch, fmt, N)
	got = np.zeros_like(want)
	row_bytes = W * ch * 4
	for r in range(n):
		d = phys[r]
		device.init(d)
		units = sharding.units_for_rank(r, n, N)
		w0, wn = sharding.source_window(units, W, H, N)
		a, b = device.malloc(wn * row_bytes, d), device.malloc(want.nbytes, d)
		device.upload(a, pano.reshape(H, -1)[w0:w0 + wn].copy(), device=d)
		for u in units:
			device.equirect_to_cube(a, W, H, ch, fmt, b, N, 1 << u.face, u.row_begin, u.row_end, src_row0=w0, src_rows=wn, device=d)
			sl = slice((u.face * N + u.row_begin) * N * ch, (u.face * N + u.row_end) * N * ch)
			device.download(got[sl], b + sl.start * 4, device=d)
		device.free(a, d), device.free(b, d)
	assert same_bits(got, want)


def test_host_api_on_second_device(built_lib, oracle):
	from cacaoengine_b200 import _capi as c

	last = _devices() - 1  # the last device of the box (device 0 on a one-GPU box), named explicitly
	src = oracle.synth(5000 * 4, c.FMT_U16, 4, 23)
	out = np.empty(5000 * 3, np.uint8)
	c.check(c.lib().cacao_img_convert(last, src.ctypes.data, out.ctypes.data, 5000, 4, 3, c.FMT_U16, c.FMT_U8, 0, c.MEM_HOST, None))
	assert (out == oracle.convert(src, 4, 3, c.FMT_U16, c.FMT_U8)).all()


def test_cubetool_project_on_two_gpus(built_lib, oracle, tmp_path):
	"""`ce-cubetool project --gpus 2`: the C++ host layer deals (face, band) units to one thread per
	rank through the C ABI's host-memory call (two threads on device 0 when the box has one GPU); output
	equals the single-device / oracle result."""
	import subprocess

	from cacaoengine_b200 import build as b

	W, H, N = 512, 256, 100
	pano = oracle.synth(W * H * 4, oracle.F32, 4, 41).reshape(H, W, 4)
	np.save(tmp_path / "p.npy", pano)
	r = subprocess.run([b.TOOL, "project", str(tmp_path / "p.npy"), "-o", str(tmp_path / "o"), "-s", str(N), "--gpus", "2"], capture_output=True, text=True)
	assert r.returncode == 0, r.stderr
	want = oracle.equirect_to_cube(pano.reshape(-1), W, H, 4, oracle.F32, N).reshape(6, N, N, 4)
	for k, nm in enumerate(["right", "left", "up", "down", "front", "back"]):
		got = np.load(tmp_path / "o" / f"p_{nm}.npy")
		assert (got.view(np.uint32) == want[k].view(np.uint32)).all(), nm


def test_table_kernel_on_every_device(built_lib, oracle):
	"""The 16 -> 8 bit kernel of convert_lanes.cuh needs 80 KiB of dynamic shared memory, an opt-in that is a
	per-device function attribute: one process driving several devices must set it on each.  A job large
	enough for that kernel (1.3 Mpixel RGBA16 -> RGB8, plain and mirrored) on every device of the box."""
	from cacaoengine_b200 import _capi as c, device

	w, h = 1280, 1024
	src = oracle.synth(w * h * 4, c.FMT_U16, 4, 24)
	want = oracle.convert(src, 4, 3, c.FMT_U16, c.FMT_U8)
	flipped = want.reshape(h, w * 3)[::-1].copy().reshape(-1)
	for d in reversed(range(_devices())):  # last device first: the attribute must not only exist where it was set first
		device.init(d)
		a, b = device.malloc(src.nbytes, d), device.malloc(want.nbytes, d)
		device.upload(a, src, device=d)
		got = np.empty_like(want)
		device.convert(a, b, w * h, 4, 3, c.FMT_U16, c.FMT_U8, device=d)
		device.download(got, b, device=d)
		assert same_bits(got, want), d
		device.convert_flip(a, b, w, h, 1, 4, 3, c.FMT_U16, c.FMT_U8, device=d)
		device.download(got, b, device=d)
		assert same_bits(got, flipped), d
		device.free(a, d), device.free(b, d)


def test_sharded_cubemap_assembled_on_one_device_by_the_launches(built_lib, oracle):
	"""SURVEY 8e, the device-side gather: every rank's resampling launch gets device 0's cube as its destination
	(peer access inside one process), so the bands cross NVLink as the kernel's own stores and the cube is complete
	on device 0 when the launches are -- no gather step, no per-rank output buffer.  On a one-GPU box all ranks are
	device 0 and the peer call is a no-op; on the 2- and 8-GPU boxes the stores really are remote."""
	from cacaoengine_b200 import _capi as c, device, sharding

	n, phys = _ranks()
	W, H, N, ch, fmt = 1024, 512, 256, 4, c.FMT_F32
	pano = oracle.synth(W * H * ch, fmt, ch, 25)
	want = oracle.equirect_to_cube(pano, W, H, ch, fmt, N)
	row_bytes = W * ch * 4
	device.init(0)
	cube = device.malloc(want.nbytes, 0)
	device.upload(cube, np.zeros(want.nbytes, np.uint8), device=0)
	device.sync(0)
	srcs = []
	for r in range(n):
		d = phys[r]
		device.init(d)
		device.enable_peer_access(d, 0)
		units = sharding.units_for_rank(r, n, N)
		w0, wn = sharding.source_window(units, W, H, N)
		a = device.malloc(wn * row_bytes, d)
		device.upload(a, pano.reshape(H, -1)[w0:w0 + wn].copy(), device=d)
		device.equirect_to_cube_units(a, W, H, ch, fmt, cube, N, units, src_row0=w0, src_rows=wn, device=d)
		srcs.append((a, d))
	for a, d in srcs:
		device.sync(d)
		device.free(a, d)
	got = np.empty_like(want)
	device.download(got, cube, device=0)
	device.free(cube, 0)
	assert same_bits(got, want)


_IPC_WRITER = r"""
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from cacaoengine_b200 import _capi as c, device, sharding
from oracle import pyoracle as po
handle, offset, dev, rank, world = bytes.fromhex(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
W, H, N, ch, fmt = 1024, 512, 256, 4, c.FMT_F32
device.init(dev)
cube = device.ipc_import(handle, offset, dev)
pano = po.synth(W * H * ch, fmt, ch, 26)
units = sharding.units_for_rank(rank, world, N)
w0, wn = sharding.source_window(units, W, H, N)
a = device.malloc(wn * W * ch * 4, dev)
device.upload(a, pano.reshape(H, -1)[w0:w0 + wn].copy(), device=dev)
device.equirect_to_cube_units(a, W, H, ch, fmt, cube, N, units, src_row0=w0, src_rows=wn, device=dev)
device.sync(dev)
device.ipc_close(cube, offset, dev)
device.free(a, dev)
print("done", len(units))
"""


def test_another_process_resamples_into_this_process_cube(built_lib, oracle):
	"""The same gather across PROCESSES (one process per GPU, as under torchrun): this process exports its cube
	buffer (cacao_img_ipc_export), each writer process opens it, resamples its rank's units straight into it from
	the last device of the box, and the cube read back here equals the oracle's."""
	import os
	import subprocess
	import sys

	from cacaoengine_b200 import _capi as c, device

	W, H, N, ch, fmt = 1024, 512, 256, 4, c.FMT_F32
	want = oracle.equirect_to_cube(oracle.synth(W * H * ch, fmt, ch, 26), W, H, ch, fmt, N)
	device.init(0)
	cube = device.malloc(want.nbytes, 0)
	device.upload(cube, np.zeros(want.nbytes, np.uint8), device=0)
	device.sync(0)
	handle, offset = device.ipc_export(cube, 0)
	root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
	world = 2
	for rank in range(world):
		dev = (_devices() - 1) if rank else 0
		r = subprocess.run([sys.executable, "-c", _IPC_WRITER, root, handle.hex(), str(offset), str(dev), str(rank), str(world)], capture_output=True, text=True, timeout=300)
		assert r.returncode == 0 and "done" in r.stdout, r.stderr[-2000:]
	got = np.empty_like(want)
	device.download(got, cube, device=0)
	device.free(cube, 0)
	assert same_bits(got, want)
