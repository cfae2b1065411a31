"""dev: device 1 resamples rank 1's units of the 2-way sharded config 5 straight into a cube buffer on device 0
(peer access, one process), then the same units into its own memory -- two launches each, for ncu's NVLink counters."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cacaoengine_b200 import FMT_F32, device, sharding
W, H, N = 16384, 8192, 4096
device.init(0); device.init(1)
device.enable_peer_access(1, 0)
units = sharding.units_for_rank(1, 2, N)
w0, wn = sharding.source_window(units, W, H, N)
src = device.malloc(wn * W * 16, 1)
device.synth(src, wn * W * 4, FMT_F32, 4, 3, first_sample=w0 * W * 4, device=1)
remote, local = device.malloc(6 * N * N * 16, 0), device.malloc(6 * N * N * 16, 1)
for dst in (remote, local):
	for _ in range(2):
		device.equirect_to_cube_units(src, W, H, 4, FMT_F32, dst, N, units, src_row0=w0, src_rows=wn, device=1)
	device.sync(1)
print("rows", sum(u.rows for u in units), "bytes", sum(u.rows for u in units) * N * 16)

if "--time" in sys.argv:
	import time
	for wide in ("0", "1"):
		os.environ["CACAO_EQUIRECT_WIDE_STORES"] = wide
		for name, dst in (("remote", remote), ("local", local)):
			for _ in range(5):
				device.equirect_to_cube_units(src, W, H, 4, FMT_F32, dst, N, units, src_row0=w0, src_rows=wn, device=1)
			device.sync(1)
			t0 = time.perf_counter()
			for _ in range(30):
				device.equirect_to_cube_units(src, W, H, 4, FMT_F32, dst, N, units, src_row0=w0, src_rows=wn, device=1)
			device.sync(1)
			dt = (time.perf_counter() - t0) / 30
			print(f"wide_stores={wide} {name}: {dt * 1e3:.4f} ms  ({sum(u.rows for u in units) * N * 16 / dt / 1e9:.0f} GB/s of output)")
	c0 = device.checksum(remote, 6 * N * N * 16, device=0)
	os.environ["CACAO_EQUIRECT_WIDE_STORES"] = "0"
	device.equirect_to_cube_units(src, W, H, 4, FMT_F32, remote, N, units, src_row0=w0, src_rows=wn, device=1)
	device.sync(1)
	print("identical:", c0 == device.checksum(remote, 6 * N * N * 16, device=0))
