"""dev: achieved GB/s of every conversion shape (132) at ~64 Mpx, device-resident, one launch each.
Writes profiles-style CSV to stdout."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device

device.init(0)
stream = torch.cuda.current_stream().cuda_stream
B = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
NAME = {FMT_U8: "u8", FMT_U16: "u16", FMT_F16: "f16", FMT_F32: "f32"}
def timeit(fn, reps=5):
	for _ in range(2): fn()
	torch.cuda.synchronize()
	if "--back-to-back" in sys.argv:  # K launches between one pair of events, as bench.py times its steps
		ts = []
		for _ in range(3):  # median of three groups of 20
			a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			a.record()
			for _ in range(4 * reps): fn()
			b.record(); torch.cuda.synchronize()
			ts.append(a.elapsed_time(b) / (4 * reps))
		return sorted(ts)[1]
	ts = []
	for _ in range(reps):
		a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		a.record(); fn(); b.record(); torch.cuda.synchronize()
		ts.append(a.elapsed_time(b))
	return sum(ts) / len(ts)
px = 64 * 1024 * 1024
src = torch.empty(px * 16, dtype=torch.uint8, device="cuda")
dst = torch.empty(px * 16, dtype=torch.uint8, device="cuda")
print("src,dst,bytes_per_px,ms,GBps,frac_of_6551.7,Gpx_s")
for sf in (FMT_U8, FMT_U16, FMT_F16, FMT_F32):
	for sc in (1, 3, 4):
		device.synth(src.data_ptr(), px * sc, sf, sc, 5, stream=stream)
		for df, dc in itertools.product((FMT_U8, FMT_U16, FMT_F16, FMT_F32), (1, 3, 4)):
			if sf == df and sc == dc: continue
			ms = timeit(lambda: device.convert(src.data_ptr(), dst.data_ptr(), px, sc, dc, sf, df, mode=1, stream=stream))
			bpp = sc * B[sf] + dc * B[df]
			gbs = px * bpp / ms / 1e6
			print(f"{NAME[sf]}x{sc},{NAME[df]}x{dc},{bpp},{ms:.4f},{gbs:.0f},{gbs/6551.7:.3f},{px/ms/1e6:.1f}", flush=True)
