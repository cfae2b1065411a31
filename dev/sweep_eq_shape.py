"""dev: sweep warp-footprint / block arrangements of the equirect kernel (CACAO_EQUIRECT_WX="WX,BWX")."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F32, FMT_U8, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=20):
	for _ in range(3): fn()
	torch.cuda.synchronize()
	ts = []
	for _ in range(reps):
		a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		a.record(); fn(); b.record(); torch.cuda.synchronize()
		ts.append(a.elapsed_time(b))
	ts.sort()
	return ts[len(ts) // 2]
B = {FMT_U8: 1, FMT_F32: 4}
variants = ["32,1"] + [f"{w},{b}" for w in (16, 8, 4) for b in (1, 2, 4)]
for name, W, H, N, ch, fmt in [("cfg3 rgba32f", 8192, 4096, 2048, 4, FMT_F32), ("cfg5 rgba32f", 16384, 8192, 4096, 4, FMT_F32), ("odd rgba32f 5000->1111", 5000, 2500, 1111, 4, FMT_F32), ("rgba8 8192->2048", 8192, 4096, 2048, 4, FMT_U8), ("upsample f32 2048->2048", 2048, 1024, 2048, 4, FMT_F32)]:
	bpp = ch * B[fmt]
	src = torch.empty(W * H * bpp, dtype=torch.uint8, device="cuda")
	d0 = torch.zeros(6 * N * N * bpp, dtype=torch.uint8, device="cuda")
	device.synth(src.data_ptr(), W * H * ch, fmt, ch, 3, stream=stream)
	alg = W * H * bpp + 6 * N * N * bpp
	out = []
	for v in variants:
		os.environ["CACAO_EQUIRECT_WX"] = v
		t = timeit(lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, d0.data_ptr(), N, stream=stream))
		out.append(f"{v}:{t*1000:.1f}us({alg/t/1e6/6551.7:.3f})")
	print(f"{name:24s} " + "  ".join(out), flush=True)
	del src, d0
