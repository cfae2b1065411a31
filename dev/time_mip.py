"""dev: throughput of cacao_img_downsample2x (one mip level) on device-resident data."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
s = torch.cuda.current_stream().cuda_stream
B = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
NAME = {FMT_U8: "u8", FMT_U16: "u16", FMT_F16: "f16", FMT_F32: "f32"}
n = 16384
for fmt in (FMT_U8, FMT_U16, FMT_F16, FMT_F32):
	for ch in (1, 3, 4):
		w = h = n if ch * B[fmt] <= 4 else n // 2
		bpp = ch * B[fmt]
		a = torch.empty(w * h * bpp, dtype=torch.uint8, device="cuda"); b = torch.empty(w * h * bpp // 4, dtype=torch.uint8, device="cuda")
		device.synth(a.data_ptr(), w * h * ch, fmt, ch, 1, stream=s)
		for _ in range(3): device.downsample2x(a.data_ptr(), b.data_ptr(), w, h, ch, fmt, stream=s)
		torch.cuda.synchronize()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		for _ in range(10): device.downsample2x(a.data_ptr(), b.data_ptr(), w, h, ch, fmt, stream=s)
		e1.record(); torch.cuda.synchronize()
		ms = e0.elapsed_time(e1) / 10
		byt = w * h * bpp * 1.25
		print(f"{NAME[fmt]}x{ch} {w}x{h}: {ms:.4f} ms  {byt/ms/1e6:.0f} GB/s ({byt/ms/1e6/6551.7:.2f})", flush=True)
		del a, b
