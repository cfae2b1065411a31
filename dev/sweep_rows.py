"""dev: the row-strip equirect kernel (equirect_rows.cu) against the gather kernel on every sample format /
layout at 8192x4096 -> 6 x 2048^2, bit equality included.  With a CACAO_ROWS_SWEEP=1 build it also sweeps
blocks per SM (registers per thread) and row steps per thread.

    python dev/sweep_rows.py [name-filter] [--sweep]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
ONLY = next((a for a in sys.argv[1:] if not a.startswith("--")), "")
SWEEP = "--sweep" in sys.argv
stream = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=20):
	for _ in range(3): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
B = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
NAMES = {FMT_U8: "8", FMT_U16: "16", FMT_F16: "16f", FMT_F32: "32f"}
shapes = [(f"{'gray' if ch == 1 else 'rgb' if ch == 3 else 'rgba'}{NAMES[f]}", 8192, 4096, 2048, ch, f) for f in (FMT_U8, FMT_U16, FMT_F16, FMT_F32) for ch in (1, 3, 4)]
shapes += [("cfg1-rgba8", 2048, 1024, 512, 4, FMT_U8), ("cfg5-rgba32f", 16384, 8192, 4096, 4, FMT_F32), ("odd-rgba8", 5000, 2500, 1111, 4, FMT_U8)]
for name, W, H, N, ch, fmt in shapes:
	if ONLY and ONLY not in name: continue
	bpp = ch * B[fmt]
	src = torch.empty(W * H * bpp, dtype=torch.uint8, device="cuda")
	d0 = torch.zeros(6 * N * N * bpp, dtype=torch.uint8, device="cuda")
	d1 = torch.zeros(6 * N * N * bpp, dtype=torch.uint8, device="cuda")
	device.synth(src.data_ptr(), W * H * ch, fmt, ch, 3, stream=stream)
	alg = W * H * bpp + 6 * N * N * bpp
	run0 = lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, d0.data_ptr(), N, stream=stream)
	run1 = lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, d1.data_ptr(), N, stream=stream)
	os.environ["CACAO_EQUIRECT_ROWS"] = "0"
	t0 = timeit(run0)
	os.environ["CACAO_EQUIRECT_ROWS"] = "1"
	os.environ.pop("CACAO_EQUIRECT_MINB", None); os.environ.pop("CACAO_EQUIRECT_STEPS", None)
	t1 = timeit(run1)
	same = bool(torch.equal(d0, d1))
	line = f"{name:14s} gather {t0:7.4f} ms ({alg/t0/1e6/6551.7:.3f})  rows {t1:7.4f} ms ({alg/t1/1e6/6551.7:.3f})  identical={same}"
	if SWEEP:
		best = None
		for mb in (8, 10, 12, 16):
			for st in (0, 2, 4, 8):  # 0: the form without a row loop
				os.environ["CACAO_EQUIRECT_MINB"] = str(mb)
				if st: os.environ["CACAO_EQUIRECT_STEPS"] = str(st)
				else: os.environ.pop("CACAO_EQUIRECT_STEPS", None)
				d1.zero_()
				t = timeit(run1, 10)
				ok = bool(torch.equal(d0, d1))
				line += f"\n      minb {mb:2d} steps {st}: {t:7.4f} ms ({alg/t/1e6/6551.7:.3f}) {'ok' if ok else 'DIFFERENT'}"
				if best is None or t < best[0]: best = (t, mb, st)
		line += f"\n      best: minb {best[1]} steps {best[2]} {best[0]:.4f} ms ({alg/best[0]/1e6/6551.7:.3f})"
	if "--wx" in sys.argv:  # gather kernel, warp footprint WX x 32/WX (development build)
		os.environ["CACAO_EQUIRECT_ROWS"] = "0"
		for wx in (8, 16, 32):
			os.environ["CACAO_EQUIRECT_WX"] = str(wx)
			d1.zero_()
			t = timeit(run1, 10)
			line += f"\n      gather, warp {wx}x{32 // wx}: {t:7.4f} ms ({alg/t/1e6/6551.7:.3f}) {'ok' if torch.equal(d0, d1) else 'DIFFERENT'}"
		os.environ.pop("CACAO_EQUIRECT_WX", None)
	print(line, flush=True)
	del src, d0, d1
