"""dev: time the hot conversion shapes under different tiles-per-block settings (CACAO_TILE_ITERS)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device

device.init(0)
stream = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=10):
	for _ in range(3): fn()
	torch.cuda.synchronize()
	ts = []
	for _ in range(reps):
		a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		a.record(); fn(); b.record(); torch.cuda.synchronize()
		ts.append(a.elapsed_time(b))
	return sum(ts) / len(ts)

shapes = [
	("cfg2 rgba8->rgba16f", 1920 * 1080 * 256, 4, 4, FMT_U8, FMT_F16, 4, 8),
	("cfg4a rgb8->rgba8", 4096 * 4096 * 16, 3, 4, FMT_U8, FMT_U8, 3, 4),
	("cfg4b rgb8->rgba32f", 4096 * 4096 * 16, 3, 4, FMT_U8, FMT_F32, 3, 16),
	("lut rgba16->rgb8", 2048 * 2048 * 64, 4, 3, FMT_U16, FMT_U8, 8, 3),
	("lut rgba16->rgba8", 2048 * 2048 * 64, 4, 4, FMT_U16, FMT_U8, 8, 4),
	("rgba8->rgb8", 4096 * 4096 * 16, 4, 3, FMT_U8, FMT_U8, 4, 3),
	("rgba32f->rgba8", 4096 * 4096 * 8, 4, 4, FMT_F32, FMT_U8, 16, 4),
	("rgb8->gray8", 4096 * 4096 * 16, 3, 1, FMT_U8, FMT_U8, 3, 1),
]
for name, px, sc, dc, sf, df, sb, db in shapes:
	src = torch.empty(px * sb, dtype=torch.uint8, device="cuda")
	dst = torch.empty(px * db, dtype=torch.uint8, device="cuda")
	device.synth(src.data_ptr(), px * sb // (1 if sf == FMT_U8 else (4 if sf == FMT_F32 else 2)), sf, sc, 7, stream=stream)
	line = f"{name:22s}"
	for iters in sys.argv[1:] or ["1", "2", "4", "8", "16"]:
		os.environ["CACAO_TILE_ITERS"] = iters
		ms = timeit(lambda: device.convert(src.data_ptr(), dst.data_ptr(), px, sc, dc, sf, df, stream=stream))
		line += f"  it={iters}: {px * (sb + db) / ms / 1e6:7.0f} GB/s"
	print(line, flush=True)
	del src, dst
