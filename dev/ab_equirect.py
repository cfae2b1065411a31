"""dev: A/B an equirect kernel variant behind an environment switch (timing + bit equality).

    python dev/ab_equirect.py [ENV_KEY [VALUE [name-filter]]]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
KEY = sys.argv[1] if len(sys.argv) > 1 else "CACAO_EQUIRECT_TMA"
VAL = sys.argv[2] if len(sys.argv) > 2 else "1"
ONLY = sys.argv[3] if len(sys.argv) > 3 else ""
stream = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=20):
	for _ in range(3): fn()
	torch.cuda.synchronize()
	ts = []
	for _ in range(reps):
		a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		a.record(); fn(); b.record(); torch.cuda.synchronize()
		ts.append(a.elapsed_time(b))
	return sum(ts) / len(ts)
B = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
for name, W, H, N, ch, fmt in [("cfg3 rgba32f 8192x4096->2048", 8192, 4096, 2048, 4, FMT_F32), ("odd rgba32f 5000x2500->1111", 5000, 2500, 1111, 4, FMT_F32), ("cfg5 rgba32f 16384x8192->4096", 16384, 8192, 4096, 4, FMT_F32),
		("rgba16f 8192x4096->2048", 8192, 4096, 2048, 4, FMT_F16), ("rgba8 8192x4096->2048", 8192, 4096, 2048, 4, FMT_U8), ("gray32f 8192x4096->2048", 8192, 4096, 2048, 1, FMT_F32), ("gray8 8192x4096->2048", 8192, 4096, 2048, 1, FMT_U8), ("rgb8 8192x4096->2048", 8192, 4096, 2048, 3, FMT_U8), ("rgb32f 8192x4096->2048", 8192, 4096, 2048, 3, FMT_F32), ("rgba16 8192x4096->2048", 8192, 4096, 2048, 4, FMT_U16), ("cfg1 rgba8 2048x1024->512", 2048, 1024, 512, 4, FMT_U8), ("upsample rgba32f 2048x1024->2048", 2048, 1024, 2048, 4, FMT_F32), ("downsample rgba32f 8192x4096->512", 8192, 4096, 512, 4, FMT_F32)]:
	if ONLY and ONLY not in name: continue
	bpp = ch * B[fmt]
	src = torch.empty(W * H * bpp, dtype=torch.uint8, device="cuda")
	d0 = torch.zeros(6 * N * N * bpp, dtype=torch.uint8, device="cuda")
	d1 = torch.zeros(6 * N * N * bpp, dtype=torch.uint8, device="cuda")
	device.synth(src.data_ptr(), W * H * ch, fmt, ch, 3, stream=stream)
	os.environ[KEY] = "0"
	t0 = timeit(lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, d0.data_ptr(), N, stream=stream))
	os.environ[KEY] = VAL
	t1 = timeit(lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, d1.data_ptr(), N, stream=stream))
	same = bool(torch.equal(d0, d1))
	alg = W * H * bpp + 6 * N * N * bpp
	print(f"{name:36s} off {t0:8.4f} ms ({alg/t0/1e6/6551.7:.3f})   on {t1:8.4f} ms ({alg/t1/1e6/6551.7:.3f})   identical={same}", flush=True)
	del src, d0, d1
