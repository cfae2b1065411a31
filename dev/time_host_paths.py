"""dev: end-to-end host->host conversion throughput: pinned vs pageable buffers, several sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cacaoengine_b200 import FMT_F16, FMT_U8, FMT_U16, MEM_HOST, device
import cacaoengine_b200 as ce
device.init(0)
PPI = 1920 * 1080
def bench(name, fn, px, reps=3):
	fn()
	t0 = time.perf_counter()
	for _ in range(reps): fn()
	dt = (time.perf_counter() - t0) / reps
	print(f"{name:48s} {dt*1e3:9.2f} ms  {px/dt/1e9:7.3f} Gpx/s", flush=True)
for frames in (1, 16, 256):
	px = PPI * frames
	hp_in, h_in = device.host_alloc(px * 4); hp_out, h_out = device.host_alloc(px * 8)
	h_in[:] = 7
	bench(f"pinned   {frames:3d} x 1080p RGBA8->RGBA16F", lambda: device.convert_batch(hp_in, hp_out, PPI, frames, 4, 4, FMT_U8, FMT_F16, mem=MEM_HOST), px)
	a = np.full(px * 4, 7, np.uint8); b = np.empty(px * 8, np.uint8)
	bench(f"pageable {frames:3d} x 1080p RGBA8->RGBA16F", lambda: device.convert_batch(a.ctypes.data, b.ctypes.data, PPI, frames, 4, 4, FMT_U8, FMT_F16, mem=MEM_HOST), px)
	assert (b[:4096] == h_out[:4096]).all()
	device.host_free(hp_in); device.host_free(hp_out)
# the reference-shaped API on one 4096^2 RGBA16 face (what Cubemap.cpp does per face)
w = h = 4096
img = ce.Image(w, h, ce.Layout.RGBA, 16, np.random.default_rng(0).integers(0, 256, w * h * 8).astype(np.uint8))
bench("Image API: 4096^2 RGBA16 -> Convert16To8BitColor", lambda: ce.Convert16To8BitColor(img), w * h)
bench("Image API: 4096^2 RGBA16 -> ToRGB8 (fused)", lambda: ce.ToRGB8(img), w * h)
# flip throughput, device resident
import torch
for (w, h, bpp) in ((4096, 4096, 4), (8192, 8192, 16)):
	n = w * h * bpp
	a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
	s = torch.cuda.current_stream().cuda_stream
	for _ in range(3): device.flip_rows(a.data_ptr(), b.data_ptr(), w, h, bpp, stream=s)
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(10): device.flip_rows(a.data_ptr(), b.data_ptr(), w, h, bpp, stream=s)
	e1.record(); torch.cuda.synchronize()
	ms = e0.elapsed_time(e1) / 10
	print(f"flip {w}x{h}x{bpp}B: {ms:.4f} ms  {2*n/ms/1e6:.0f} GB/s", flush=True)
