import os, sys, csv
sys.path.insert(0, "/root/repo")
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
B = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
F = {"u8": FMT_U8, "u16": FMT_U16, "f16": FMT_F16, "f32": FMT_F32}
def timeit(fn, reps=20):
	for _ in range(5): fn()
	torch.cuda.synchronize()
	ts = []
	for _ in range(3):
		a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		a.record()
		for _ in range(reps): fn()
		b.record(); torch.cuda.synchronize()
		ts.append(a.elapsed_time(b) / reps)
	return sorted(ts)[1]
shapes = [("f16", 1, "u8", 3), ("f16", 1, "u8", 4), ("f32", 4, "u8", 3), ("f32", 3, "u8", 1), ("f32", 3, "u8", 3), ("u8", 1, "u8", 3), ("f32", 4, "u8", 4), ("f16", 1, "u8", 1), ("u8", 3, "u8", 4), ("u8", 1, "u16", 1)]
px = 64 << 20
src = torch.empty(px * 16, dtype=torch.uint8, device="cuda"); dst = torch.empty(px * 16, dtype=torch.uint8, device="cuda")
for sfn, sc, dfn, dc in shapes:
	sf, df = F[sfn], F[dfn]
	device.synth(src.data_ptr(), px * sc, sf, sc, 5, stream=stream)
	line = f"{sfn}x{sc}->{dfn}x{dc}:"
	for kern in ("", "tile", "wide"):
		for it in ("", "2", "4"):
			os.environ.pop("CACAO_NO_WIDE", None); os.environ.pop("CACAO_FORCE_WIDE", None); os.environ.pop("CACAO_TILE_ITERS", None)
			if kern == "tile": os.environ["CACAO_NO_WIDE"] = "1"
			if kern == "wide": os.environ["CACAO_FORCE_WIDE"] = "1"
			if it: os.environ["CACAO_TILE_ITERS"] = it
			ms = timeit(lambda: device.convert(src.data_ptr(), dst.data_ptr(), px, sc, dc, sf, df, mode=1, stream=stream))
			line += f"  {kern or 'auto'}{('/x' + it) if it else ''} {px * (sc * B[sf] + dc * B[df]) / ms / 1e6 / 6551.7:.3f}"
	print(line, flush=True)
