"""dev: U16->U8 shapes under table mode (two-level 8 KiB vs direct 64 KiB) x tiles-per-block."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_U8, FMT_U16, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=8):
	for _ in range(3): fn()
	torch.cuda.synchronize()
	ts = []
	for _ in range(reps):
		a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		a.record(); fn(); b.record(); torch.cuda.synchronize()
		ts.append(a.elapsed_time(b))
	return sum(ts) / len(ts)
px = 2048 * 2048 * 64
src = torch.empty(px * 8, dtype=torch.uint8, device="cuda")
dst = torch.empty(px * 4, dtype=torch.uint8, device="cuda")
smooth = "--smooth" in sys.argv
for sc, dc in ((4, 4), (4, 3), (1, 1), (3, 3), (1, 3), (4, 1)):
	device.synth(src.data_ptr(), px * sc, FMT_U16, sc, 7, stream=stream)
	if smooth:  # neighbouring samples close in value, like a real image
		t = (torch.arange(px * sc, device="cuda", dtype=torch.int64) // 64 % 65536).to(torch.int16)
		src[: px * sc * 2].view(torch.int16).copy_(t)
	line = f"u16x{sc}->u8x{dc} "
	for wide in ("0", "1"):
		for direct in ("0", "1"):
			for iters in ("4", "16"):
				if wide == "0" and direct == "1": continue
				os.environ.pop("CACAO_NO_WIDE", None)
				if wide == "0": os.environ["CACAO_NO_WIDE"] = "1"
				os.environ["CACAO_LUT_DIRECT"] = direct
				os.environ["CACAO_TILE_ITERS"] = iters
				ms = timeit(lambda: device.convert(src.data_ptr(), dst.data_ptr(), px, sc, dc, FMT_U16, FMT_U8, stream=stream))
				line += f" w{wide}d{direct}i{iters}:{px * (2 * sc + dc) / ms / 1e6:6.0f}"
	print(line, flush=True)
