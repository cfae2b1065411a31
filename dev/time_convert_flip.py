"""dev: conversion with the row mirror folded in (cacao_img_convert_flip) against convert + flip_rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F16, FMT_F32, FMT_U8, FMT_U16, device
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
BYTES = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}
def t(fn, reps=20):
	for _ in range(3): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
for name, w, h, count, sc, dc, sf, df in (("engine cube: 6 x 2048^2 RGBA16 -> RGB8", 2048, 2048, 6, 4, 3, FMT_U16, FMT_U8), ("64 x 2048^2 RGBA16 -> RGB8", 2048, 2048, 64, 4, 3, FMT_U16, FMT_U8),
		("64 x 2048^2 RGBA8 -> RGBA16F", 2048, 2048, 64, 4, 4, FMT_U8, FMT_F16), ("64 x 2048^2 RGB8 -> RGBA8", 2048, 2048, 64, 3, 4, FMT_U8, FMT_U8), ("16 x 4000x3000 RGB8 -> RGBA8 (tile kernel unit does not divide 4000)", 4000, 3000, 16, 3, 4, FMT_U8, FMT_U8),
		("16 x 4001x3000 RGB8 -> RGBA8 (byte-wise kernel)", 4001, 3000, 16, 3, 4, FMT_U8, FMT_U8)):
	px = w * h * count
	sb, db = sc * BYTES[sf], dc * BYTES[df]
	src = torch.empty(px * sb, dtype=torch.uint8, device="cuda"); mid = torch.empty(px * db, dtype=torch.uint8, device="cuda"); dst = torch.empty(px * db, dtype=torch.uint8, device="cuda")
	device.synth(src.data_ptr(), px * sc, sf, sc, 5, stream=stream)
	def two():
		device.convert_batch(src.data_ptr(), mid.data_ptr(), w * h, count, sc, dc, sf, df, stream=stream)
		for k in range(count): device.flip_rows(mid.data_ptr() + k * w * h * db, dst.data_ptr() + k * w * h * db, w, h, db, stream=stream)
	def one():
		device.convert_flip(src.data_ptr(), dst.data_ptr(), w, h, count, sc, dc, sf, df, stream=stream)
	def plain():
		device.convert_batch(src.data_ptr(), mid.data_ptr(), w * h, count, sc, dc, sf, df, stream=stream)
	t2, t1, t0 = t(two), t(one), t(plain)
	print(f"{name}: convert {t0:.4f} ms, convert+flip fused {t1:.4f} ms ({px * (sb + db) / t1 / 1e6:.0f} GB/s), convert then flip {t2:.4f} ms")
