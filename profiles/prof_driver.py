"""Small driver for ncu captures: launches each hot kernel of the BASELINE configs a few times on
device-resident synthetic data, through the C ABI only (no torch).  Usage:

    python profiles/prof_driver.py [cfg1] [cfg2] [cfg3] [cfg5] [rgb32f] [fused] [shard8] [cfg4a] [cfg4b] [lut] [engine] [narrow] [--reps N]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device  # noqa: E402


def main():
	args = [a for a in sys.argv[1:] if not a.startswith("--")]
	reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 2
	which = args or ["cfg2", "cfg3", "cfg4a", "cfg4b", "lut"]
	device.init(0)
	if "cfg2" in which:
		px = 1920 * 1080 * 256
		a, b = device.malloc(px * 4), device.malloc(px * 8)
		device.synth(a, px * 4, FMT_U8, 4, 1)
		for _ in range(reps):
			device.convert_batch(a, b, 1920 * 1080, 256, 4, 4, FMT_U8, FMT_F16)
		device.sync()
		device.free(a), device.free(b)
	if "cfg3" in which:
		W, H, N = 8192, 4096, 2048
		a, b = device.malloc(W * H * 16), device.malloc(6 * N * N * 16)
		device.synth(a, W * H * 4, FMT_F32, 4, 2)
		for _ in range(reps):
			device.equirect_to_cube(a, W, H, 4, FMT_F32, b, N)
		device.sync()
		device.free(a), device.free(b)
	for key, (W, H, N, ch, fmt, bps) in {"cfg1": (2048, 1024, 512, 4, FMT_U8, 1), "cfg5": (16384, 8192, 4096, 4, FMT_F32, 4), "rgb32f": (8192, 4096, 2048, 3, FMT_F32, 4)}.items():
		if key in which:
			a, b = device.malloc(W * H * ch * bps), device.malloc(6 * N * N * ch * bps)
			device.synth(a, W * H * ch, fmt, ch, 2)
			for _ in range(reps):
				device.equirect_to_cube(a, W, H, ch, fmt, b, N)
			device.sync()
			device.free(a), device.free(b)
	if "fused" in which:
		# config 3 with the cube written as RGBA16F in the same pass (Spec B.4)
		W, H, N = 8192, 4096, 2048
		a, b = device.malloc(W * H * 16), device.malloc(6 * N * N * 8)
		device.synth(a, W * H * 4, FMT_F32, 4, 2)
		for _ in range(reps):
			device.equirect_to_cube_cvt(a, W, H, 4, FMT_F32, b, N, 4, FMT_F16)
		device.sync()
		device.free(a), device.free(b)
	if "shard8" in which:
		# what rank 0 of an 8-GPU config-5 job launches: its mirror-pair unit list, folded into four-fold
		# units (first `reps` launches), then the same list computed two-fold (next `reps`)
		from cacaoengine_b200 import sharding
		W, H, N = 16384, 8192, 4096
		a, b = device.malloc(W * H * 16), device.malloc(6 * N * N * 16)
		device.synth(a, W * H * 4, FMT_F32, 4, 2)
		units = sharding.units_for_rank(0, 8, N)
		for env in ("0", "1"):
			os.environ["CACAO_EQUIRECT_NO_QUAD"] = env
			for _ in range(reps):
				device.equirect_to_cube_units(a, W, H, 4, FMT_F32, b, N, units)
		os.environ["CACAO_EQUIRECT_NO_QUAD"] = "0"
		device.sync()
		device.free(a), device.free(b)
	if "cfg4a" in which or "cfg4b" in which:
		px, ring = 4096 * 4096, 16
		a = device.malloc(ring * px * 3)
		device.synth(a, ring * px * 3, FMT_U8, 3, 3)
		if "cfg4a" in which:
			b = device.malloc(ring * px * 4)
			for _ in range(reps):
				device.convert_batch(a, b, px, ring, 3, 4, FMT_U8, FMT_U8)
			device.sync()
			device.free(b)
		if "cfg4b" in which:
			b = device.malloc(ring * px * 16)
			for _ in range(reps):
				device.convert_batch(a, b, px, ring, 3, 4, FMT_U8, FMT_F32)
			device.sync()
			device.free(b)
		device.free(a)
	if "lut" in which:
		# the engine's cubemap path: RGBA16 faces -> RGB8 in one kernel (Cubemap.cpp:28-31), 64 x 2048^2
		px = 2048 * 2048 * 64
		a, b = device.malloc(px * 8), device.malloc(px * 3)
		device.synth(a, px * 4, FMT_U16, 4, 4)
		for _ in range(reps):
			device.convert_batch(a, b, 2048 * 2048, 64, 4, 3, FMT_U16, FMT_U8)
		device.sync()
		device.free(a), device.free(b)
	if "engine" in which:
		# bench.py's extra.engine_cube_6x2048sq_rgba16_to_rgb8_flipped: six faces, 16 -> 8 bit, -> RGB, rows mirrored, one launch
		w = h = 2048
		a, b = device.malloc(6 * w * h * 8), device.malloc(6 * w * h * 3)
		device.synth(a, 6 * w * h * 4, FMT_U16, 4, 0xC0FFEE + 9)
		for _ in range(reps):
			device.convert_flip(a, b, w, h, 6, 4, 3, FMT_U16, FMT_U8)
		device.sync()
		device.free(a), device.free(b)
	if "narrow" in which:
		# bench.py's extra.cfg3shape_equirect_8192x4096_*_to_6x2048 legs: config 3's geometry on the narrow formats
		W, H, N = 8192, 4096, 2048
		for ch, fmt, bps in ((4, FMT_U8, 1), (4, FMT_U16, 2), (4, FMT_F16, 2), (1, FMT_U8, 1), (3, FMT_U8, 1)):
			a, b = device.malloc(W * H * ch * bps), device.malloc(6 * N * N * ch * bps)
			device.synth(a, W * H * ch, fmt, ch, 0xC0FFEE + 3)
			for _ in range(reps):
				device.equirect_to_cube(a, W, H, ch, fmt, b, N)
			device.sync()
			device.free(a), device.free(b)
	print("prof_driver done:", which, "launches", device.launch_count())


if __name__ == "__main__":
	main()
