"""dev: the tensor-map TMA ring (CACAO_EQUIRECT_TMA=1) against the product kernels on the 4- and 8-byte texels,
ring depth 2 / 3 / 4 stages x CTAs per SM, bit equality on every line."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
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
for name, W, H, N, ch, fmt in (("rgba8", 8192, 4096, 2048, 4, FMT_U8), ("gray32f", 8192, 4096, 2048, 1, FMT_F32), ("rgba16f", 8192, 4096, 2048, 4, FMT_F16), ("rgba16", 8192, 4096, 2048, 4, FMT_U16), ("cfg1 rgba8", 2048, 1024, 512, 4, FMT_U8)):
	bpp = ch * B[fmt]
	src = torch.empty(W * H * bpp, dtype=torch.uint8, device="cuda")
	d0 = torch.zeros(6 * N * N * bpp, dtype=torch.uint8, device="cuda")
	d1 = torch.zeros(6 * N * N * bpp, dtype=torch.uint8, device="cuda")
	device.synth(src.data_ptr(), W * H * ch, fmt, ch, 3, stream=stream)
	alg = W * H * bpp + 6 * N * N * bpp
	for k in ("CACAO_EQUIRECT_TMA", "CACAO_EQUIRECT_TMA_STAGES", "CACAO_EQUIRECT_TMA_CTAS"): os.environ.pop(k, None)
	t0 = timeit(lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, d0.data_ptr(), N, stream=stream))
	line = f"{name:10s} {W}x{H}->{N}: product {t0:.4f} ms ({alg/t0/1e6/6551.7:.3f})"
	os.environ["CACAO_EQUIRECT_TMA"] = "1"
	for st in ("2", "3", "4"):
		for ctas in ("2", "4", "7"):
			os.environ["CACAO_EQUIRECT_TMA_STAGES"] = st; os.environ["CACAO_EQUIRECT_TMA_CTAS"] = ctas
			d1.zero_()
			t = timeit(lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, d1.data_ptr(), N, stream=stream), 10)
			line += f"\n      tma ring, {st} stages, <= {ctas} CTAs/SM: {t:.4f} ms ({alg/t/1e6/6551.7:.3f}) {'identical' if torch.equal(d0, d1) else 'DIFFERENT'}"
	print(line, flush=True)
	del src, d0, d1
