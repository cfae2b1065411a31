"""dev: multi-channel -> one-channel conversions (tiny outputs): passes per block 1 / 2 / 4 (CACAO_TILE_ITERS), default kernel choice."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
B = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
NAME = {FMT_U8: "u8", FMT_U16: "u16", FMT_F16: "f16", FMT_F32: "f32"}
def timeit(fn, reps=20):
	for _ in range(5): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
which = sys.argv[1] if len(sys.argv) > 1 else "gray"
for px in (25 << 20, 64 << 20, 256 << 20):
	src = torch.empty(px * 16, dtype=torch.uint8, device="cuda"); dst = torch.empty(px * 16, dtype=torch.uint8, device="cuda")
	for sf, sc in itertools.product((FMT_U8, FMT_U16, FMT_F16, FMT_F32), (1, 3, 4)):
		device.synth(src.data_ptr(), px * sc, sf, sc, 5, stream=stream)
		for df, dc in itertools.product((FMT_U8, FMT_U16, FMT_F16, FMT_F32), (1, 3, 4)):
			if sf == df and sc == dc: continue
			if which == "gray" and not (dc == 1 and sc > 1): continue
			if sf == FMT_U16 and df == FMT_U8: continue  # table shapes: their own kernel and pass count
			line = f"{NAME[sf]}x{sc}->{NAME[df]}x{dc} {px >> 20:3d} Mpx:"
			best, base = None, None
			for it in ("", "1", "2", "4", "8"):
				if it: os.environ["CACAO_TILE_ITERS"] = it
				ms = timeit(lambda: device.convert(src.data_ptr(), dst.data_ptr(), px, sc, dc, sf, df, mode=1, stream=stream))
				os.environ.pop("CACAO_TILE_ITERS", None)
				fr = px * (sc * B[sf] + dc * B[df]) / ms / 1e6 / 6551.7
				line += f"  {'auto' if not it else 'x' + it} {fr:.3f}"
			print(line, flush=True)
	del src, dst
