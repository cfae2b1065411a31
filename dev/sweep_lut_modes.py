"""dev: U16->U8 shapes under the three table forms (two-level 8 KiB / direct 64 KiB / bank-private lanes 80.5 KiB)
at engine and batch sizes, uniform-random and smooth data; every lanes result compared with the two-level one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_U8, FMT_U16, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
PEAK = 6551.7
def timeit(fn, reps=20):
	for _ in range(5): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
iters_list = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1][0].isdigit() else [""]
for faces in (6, 24, 64):
	px = 2048 * 2048 * faces
	src = torch.empty(px * 8, dtype=torch.uint8, device="cuda")
	dst = torch.empty(px * 4, dtype=torch.uint8, device="cuda")
	ref = torch.empty(px * 4, dtype=torch.uint8, device="cuda")
	for smooth in (False, True):
		for sc, dc in ((4, 3), (4, 4), (1, 1), (3, 3), (1, 3), (1, 4), (3, 4), (4, 1), (3, 1)):
			device.synth(src.data_ptr(), px * sc, FMT_U16, sc, 7, stream=stream)
			if smooth:
				t = (torch.arange(px * sc, device="cuda", dtype=torch.int64) // 64 % 65536).to(torch.int16)
				src[: px * sc * 2].view(torch.int16).copy_(t)
			line = f"faces={faces:2d} {'smooth' if smooth else 'random'} u16x{sc}->u8x{dc} "
			gb = px * (2 * sc + dc) / 1e6
			for mode in ("two", "lanes"):
				for it in iters_list:
					os.environ["CACAO_LUT_MODE"] = mode
					os.environ.pop("CACAO_TILE_ITERS", None)
					if it: os.environ["CACAO_TILE_ITERS"] = it
					out = ref if mode == "two" else dst
					out.zero_()
					ms = timeit(lambda: device.convert(src.data_ptr(), out.data_ptr(), px, sc, dc, FMT_U16, FMT_U8, mode=1, stream=stream))
					same = "" if mode == "two" else (" ok" if torch.equal(dst[: px * dc], ref[: px * dc]) else " MISMATCH")
					line += f"  {mode}{('/' + it) if it else ''}: {ms * 1000:7.1f} us {gb / ms / PEAK:5.3f}{same}"
			os.environ.pop("CACAO_LUT_MODE", None); os.environ.pop("CACAO_TILE_ITERS", None)
			ms = timeit(lambda: device.convert(src.data_ptr(), dst.data_ptr(), px, sc, dc, FMT_U16, FMT_U8, mode=1, stream=stream))
			line += f"  auto: {ms * 1000:7.1f} us {gb / ms / PEAK:5.3f}"
			print(line, flush=True)
	del src, dst, ref
