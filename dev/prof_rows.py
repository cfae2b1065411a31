"""dev: one launch each of the gather kernel and of row-strip kernel variants on a few narrow shapes, for ncu.
Variants: "minb:steps" (steps 0 = no row loop); needs a CACAO_ROWS_SWEEP=1 build for anything but the default."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
B = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
which = [a for a in sys.argv[1:] if ":" not in a] or ["rgba8", "gray8", "rgba16f"]
variants = [a for a in sys.argv[1:] if ":" in a] or ["12:0"]
shapes = {"rgba8": (4, FMT_U8), "gray8": (1, FMT_U8), "rgba16f": (4, FMT_F16), "rgba16": (4, FMT_U16), "gray32f": (1, FMT_F32), "rgb8": (3, FMT_U8), "rgba32f": (4, FMT_F32)}
W, H, N = 8192, 4096, 2048
for name in which:
	ch, fmt = shapes[name]
	bpp = ch * B[fmt]
	a, b = device.malloc(W * H * bpp), device.malloc(6 * N * N * bpp)
	device.synth(a, W * H * ch, fmt, ch, 2)
	os.environ["CACAO_EQUIRECT_ROWS"] = "0"
	for _ in range(2):
		device.equirect_to_cube(a, W, H, ch, fmt, b, N)
	os.environ["CACAO_EQUIRECT_ROWS"] = "1"
	for v in variants:
		mb, st = v.split(":")
		os.environ["CACAO_EQUIRECT_MINB"] = mb
		if int(st): os.environ["CACAO_EQUIRECT_STEPS"] = st
		else: os.environ.pop("CACAO_EQUIRECT_STEPS", None)
		for _ in range(2):
			device.equirect_to_cube(a, W, H, ch, fmt, b, N)
	device.sync()
	device.free(a), device.free(b)
