"""dev: two launches each of the table shapes under the two-level and the bank-private table form, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cacaoengine_b200 import FMT_U8, FMT_U16, device
device.init(0)
faces = int(sys.argv[1]) if len(sys.argv) > 1 else 24
px = 2048 * 2048 * faces
a, b = device.malloc(px * 8), device.malloc(px * 4)
for sc, dc in ((4, 3), (4, 4)):
	device.synth(a, px * sc, FMT_U16, sc, 7)
	for mode in ("two", "lanes"):
		os.environ["CACAO_LUT_MODE"] = mode
		for _ in range(2):
			device.convert(a, b, px, sc, dc, FMT_U16, FMT_U8)
device.sync()
