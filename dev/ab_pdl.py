"""dev: back-to-back launches with and without programmatic dependent launch (CACAO_NO_PDL=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=50):
	for _ in range(10): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps * 1000
def ab(name, fn, check=None):
	res = {}
	for rnd in range(3):
		for pdl in ("1", "0"):
			os.environ["CACAO_NO_PDL"] = "0" if pdl == "1" else "1"
			res.setdefault(pdl, []).append(timeit(fn))
	os.environ["CACAO_NO_PDL"] = "0"
	print(f"{name:58s} plain {min(res['0']):8.2f} us   pdl {min(res['1']):8.2f} us   {100 * (min(res['0']) / min(res['1']) - 1):+5.1f} %", flush=True)
w = h = 2048
for faces in (1, 6, 24):
	px = w * h * faces
	a = torch.empty(px * 8, dtype=torch.uint8, device="cuda"); b = torch.empty(px * 4, dtype=torch.uint8, device="cuda")
	device.synth(a.data_ptr(), px * 4, FMT_U16, 4, 9, stream=stream)
	ab(f"rgba16->rgb8 flipped, {faces} x 2048^2", lambda: device.convert_flip(a.data_ptr(), b.data_ptr(), w, h, faces, 4, 3, FMT_U16, FMT_U8, stream=stream))
	ab(f"rgba16->rgba8, {faces} x 2048^2", lambda: device.convert(a.data_ptr(), b.data_ptr(), px, 4, 4, FMT_U16, FMT_U8, stream=stream))
	ab(f"gray16->gray8, {faces} x 2048^2", lambda: device.convert(a.data_ptr(), b.data_ptr(), px, 1, 1, FMT_U16, FMT_U8, stream=stream))
	del a, b
if "eq" in sys.argv:
	from cacaoengine_b200 import sharding
	for name, W, H, N, ch, fmt, bps in (("cfg1 rgba8 2048x1024->512", 2048, 1024, 512, 4, FMT_U8, 1), ("cfg3 rgba32f 8192x4096->2048", 8192, 4096, 2048, 4, FMT_F32, 4), ("rgba8 8192x4096->2048", 8192, 4096, 2048, 4, FMT_U8, 1), ("gray8 8192x4096->2048", 8192, 4096, 2048, 1, FMT_U8, 1), ("rgba16f 8192x4096->2048", 8192, 4096, 2048, 4, FMT_F16, 2)):
		a = torch.empty(W * H * ch * bps, dtype=torch.uint8, device="cuda"); b = torch.empty(6 * N * N * ch * bps, dtype=torch.uint8, device="cuda")
		device.synth(a.data_ptr(), W * H * ch, fmt, ch, 3, stream=stream)
		ab(name, lambda: device.equirect_to_cube(a.data_ptr(), W, H, ch, fmt, b.data_ptr(), N, stream=stream))
		del a, b
	W, H, N = 16384, 8192, 4096
	a = torch.empty(W * H * 16, dtype=torch.uint8, device="cuda"); b = torch.empty(6 * N * N * 16, dtype=torch.uint8, device="cuda")
	device.synth(a.data_ptr(), W * H * 4, FMT_F32, 4, 3, stream=stream)
	for world in (2, 8):
		for rank in (0, world - 1):
			units = sharding.units_for_rank(rank, world, N)
			ab(f"cfg5 sharded, rank {rank} of {world}", lambda: device.equirect_to_cube_units(a.data_ptr(), W, H, 4, FMT_F32, b.data_ptr(), N, units, stream=stream))
