"""dev: the U16->U8 table shapes against problem size and blocks' run length (CACAO_TILE_ITERS): how much of the
gap to the roofline at engine sizes (6 x 2048^2) is the per-block table staging / launch tail, how much the gather."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_U8, FMT_U16, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=30):
	for _ in range(5): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
smooth = "--smooth" in sys.argv
for sc, dc in ((4, 3), (4, 4), (1, 1), (3, 3)):
	for count in (1, 6, 24, 64):
		w = h = 2048
		px = w * h * count
		src = torch.empty(px * sc * 2, dtype=torch.uint8, device="cuda")
		dst = torch.empty(px * dc, dtype=torch.uint8, device="cuda")
		device.synth(src.data_ptr(), px * sc, FMT_U16, sc, 7, stream=stream)
		if smooth:
			t = (torch.arange(px * sc, device="cuda", dtype=torch.int64) // 64 % 65536).to(torch.int16)
			src.view(torch.int16).copy_(t)
		line = f"u16x{sc}->u8x{dc} {count:2d} x 2048^2:"
		for iters in ("auto", "1", "2", "4", "8", "16"):
			if iters == "auto": os.environ.pop("CACAO_TILE_ITERS", None)  # the launcher's wave-aligned choice
			else: os.environ["CACAO_TILE_ITERS"] = iters
			for flip in (False, True):
				fn = (lambda: device.convert_flip(src.data_ptr(), dst.data_ptr(), w, h, count, sc, dc, FMT_U16, FMT_U8, stream=stream)) if flip else (lambda: device.convert_batch(src.data_ptr(), dst.data_ptr(), w * h, count, sc, dc, FMT_U16, FMT_U8, stream=stream))
				ms = timeit(fn)
				line += f"  i{iters}{'f' if flip else ' '} {px * (2 * sc + dc) / ms / 1e6 / 6551.7:.3f}"
		print(line, flush=True)
		del src, dst
