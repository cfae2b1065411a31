"""dev: throughput of cacao_img_unpack_texels (assimp texel swizzle / widen) on device-resident data."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_U8, device
device.init(0)
s = torch.cuda.current_stream().cuda_stream
n = 1 << 28
a = torch.empty(n * 4, dtype=torch.uint8, device="cuda"); b = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
device.synth(a.data_ptr(), n * 4, FMT_U8, 4, 1, stream=s)
for hint in ("bgra8888", "rgba8880", "rgba5650", "rgba4444", "rgba0800"):
	ch = sum(c != "0" for c in hint[4:])
	for _ in range(3): device.unpack_texels(a.data_ptr(), b.data_ptr(), n, hint, stream=s)
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(10): device.unpack_texels(a.data_ptr(), b.data_ptr(), n, hint, stream=s)
	e1.record(); torch.cuda.synchronize()
	ms = e0.elapsed_time(e1) / 10
	print(f"{hint}: {ms:.4f} ms  {n*(4+ch)/ms/1e6:.0f} GB/s ({n*(4+ch)/ms/1e6/6551.7:.2f} of the measured peak)  {n/ms/1e6:.0f} Gtexel/s", flush=True)
