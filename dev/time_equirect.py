"""dev: time the equirect kernel on the BASELINE shapes and a few formats."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device

device.init(0)
stream = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=20):
	"""back-to-back launches between two events (what bench.py measures)"""
	for _ in range(3): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps

B = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
for name, W, H, N, ch, fmt in [("cfg3 rgba32f 8192x4096->2048", 8192, 4096, 2048, 4, FMT_F32), ("cfg5 rgba32f 16384x8192->4096", 16384, 8192, 4096, 4, FMT_F32),
		("rgba16f 8192x4096->2048", 8192, 4096, 2048, 4, FMT_F16), ("rgba8 8192x4096->2048", 8192, 4096, 2048, 4, FMT_U8), ("rgb8 8192x4096->2048", 8192, 4096, 2048, 3, FMT_U8),
		("rgb32f 8192x4096->2048", 8192, 4096, 2048, 3, FMT_F32), ("cfg1 rgba8 2048x1024->512", 2048, 1024, 512, 4, FMT_U8)]:
	bpp = ch * B[fmt]
	src = torch.empty(W * H * bpp, dtype=torch.uint8, device="cuda")
	dst = torch.empty(6 * N * N * bpp, dtype=torch.uint8, device="cuda")
	device.synth(src.data_ptr(), W * H * ch, fmt, ch, 3, stream=stream)
	ms = timeit(lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, dst.data_ptr(), N, stream=stream))
	alg = W * H * bpp + 6 * N * N * bpp
	print(f"{name:34s} {ms:8.4f} ms  {6*N*N/ms/1e6:8.1f} Gpx/s  {alg/ms/1e6:7.0f} GB/s  frac {alg/ms/1e6/6551.7:.3f}", flush=True)
	del src, dst
