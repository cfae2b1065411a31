"""dev: where the bank-private table form starts to pay: U16->U8 shapes, two-level vs lanes, small to mid sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_U8, FMT_U16, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=50):
	for _ in range(10): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
pxmax = 4096 * 4096
src = torch.empty(pxmax * 8, dtype=torch.uint8, device="cuda")
dst = torch.empty(pxmax * 4, dtype=torch.uint8, device="cuda")
shapes = ((4, 3), (4, 4), (1, 1), (3, 3), (1, 3), (1, 4), (3, 4), (4, 1), (3, 1))
for sc, dc in shapes:
	device.synth(src.data_ptr(), pxmax * sc, FMT_U16, sc, 7, stream=stream)
	for side in (256, 512, 724, 1024, 1448, 2048, 4096):
		px = side * side // 32 * 32
		line = f"u16x{sc}->u8x{dc} {side:4d}^2 "
		for mode in ("two", "lanes", None):
			if mode: os.environ["CACAO_LUT_MODE"] = mode
			else: os.environ.pop("CACAO_LUT_MODE", None)
			ms = timeit(lambda: device.convert(src.data_ptr(), dst.data_ptr(), px, sc, dc, FMT_U16, FMT_U8, mode=1, stream=stream))
			line += f"  {mode or 'auto'}: {ms * 1000:6.1f} us"
		print(line, flush=True)
