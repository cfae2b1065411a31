"""dev: the row-strip kernel with and without opposite-face sharing (CACAO_EQUIRECT_NO_OPPOSITE=1), outputs compared.
Record of a rejected experiment: the switch belongs to an experimental build that is not in the tree
(profiles/r2_equirect_rows_opposite_face_rejected.log); with the shipped library both columns are the same kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
B = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
def timeit(fn, reps=30):
	for _ in range(5): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps * 1000
shapes = [("gray8", 1, FMT_U8), ("rgba8", 4, FMT_U8), ("rgb8", 3, FMT_U8), ("gray16", 1, FMT_U16), ("gray32f", 1, FMT_F32), ("rgba16f", 4, FMT_F16), ("rgba16", 4, FMT_U16), ("rgba32f", 4, FMT_F32)]
for W, H, N in ((8192, 4096, 2048), (2048, 1024, 512), (3000, 1500, 777)):
	for name, ch, fmt in shapes:
		bpp = ch * B[fmt]
		src = torch.empty(W * H * bpp, dtype=torch.uint8, device="cuda"); out = [torch.zeros(6 * N * N * bpp, dtype=torch.uint8, device="cuda") for _ in range(2)]
		device.synth(src.data_ptr(), W * H * ch, fmt, ch, 3, stream=stream)
		os.environ["CACAO_EQUIRECT_ROWS"] = "1"
		res = []
		for k, off in enumerate(("1", "0")):
			os.environ["CACAO_EQUIRECT_NO_OPPOSITE"] = off
			res.append(timeit(lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, out[k].data_ptr(), N, stream=stream)))
		os.environ.pop("CACAO_EQUIRECT_ROWS"); os.environ.pop("CACAO_EQUIRECT_NO_OPPOSITE")
		os.environ["CACAO_EQUIRECT_ROWS"] = "0"
		g = timeit(lambda: device.equirect_to_cube(src.data_ptr(), W, H, ch, fmt, out[0].data_ptr(), N, stream=stream))
		os.environ.pop("CACAO_EQUIRECT_ROWS")
		same = torch.equal(out[0], out[1])
		print(f"{W}x{H}->{N} {name:8s} gather {g:8.2f} us  rows 4-fold {res[0]:8.2f} us  rows 8-fold {res[1]:8.2f} us  {100 * (res[0] / res[1] - 1):+5.1f} %  identical={same}", flush=True)
		del src, out
