"""dev: latency of the reference-shaped API on small images vs the compiled reference on the CPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cacaoengine_b200 as ce
from cacaoengine_b200 import device
from oracle import pyoracle as po
device.init(0)
rng = np.random.default_rng(0)
def bench(fn, reps):
	fn(); fn()
	t0 = time.perf_counter()
	for _ in range(reps): fn()
	return (time.perf_counter() - t0) / reps
print(f"{'size':>10s} {'op':>22s} {'gpu ms':>9s} {'ref cpu ms':>11s}")
for n in (16, 32, 64, 96, 128, 256, 1024, 2048, 4096):
	d8 = rng.integers(0, 256, n * n * 4).astype(np.uint8)
	d16 = rng.integers(0, 65536, n * n * 4).astype(np.uint16)
	img8 = ce.Image(n, n, ce.Layout.RGBA, 8, d8)
	img16 = ce.Image(n, n, ce.Layout.RGBA, 16, d16.view(np.uint8))
	reps = 200 if n <= 256 else (50 if n <= 1024 else 10)
	g1 = bench(lambda: ce.ChangeChannelLayout(img8, ce.Layout.RGB), reps)
	c1 = bench(lambda: po.ref_change_layout(d8, n, n, 4, 3), max(3, reps // 10)) if po.have_ref() else float("nan")
	g2 = bench(lambda: ce.Convert16To8BitColor(img16), reps)
	c2 = bench(lambda: po.ref_depth16to8(d16, n, n, 4), max(3, reps // 10)) if po.have_ref() else float("nan")
	print(f"{n:>6d}^2   {'RGBA8->RGB8':>22s} {g1*1e3:9.3f} {c1*1e3:11.3f}")
	print(f"{n:>6d}^2   {'RGBA16->RGBA8 (LUT)':>22s} {g2*1e3:9.3f} {c2*1e3:11.3f}")
