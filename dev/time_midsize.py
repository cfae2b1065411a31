"""dev: where does the time go for mid-size host conversions?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cacaoengine_b200 import FMT_U8, FMT_U16, MEM_HOST, device
device.init(0)
def bench(fn, reps=20):
	fn(); fn()
	t0 = time.perf_counter()
	for _ in range(reps): fn()
	return (time.perf_counter() - t0) / reps * 1e3
n = 2048
px = n * n
for name, sc, dc, sf, df, sb, db in (("RGBA8->RGB8", 4, 3, FMT_U8, FMT_U8, 4, 3), ("RGBA16->RGBA8", 4, 4, FMT_U16, FMT_U8, 8, 4), ("RGBA16->RGB8", 4, 3, FMT_U16, FMT_U8, 8, 3), ("RGBA8->RGBA16", 4, 4, FMT_U8, FMT_U16, 4, 8)):
	a = np.full(px * sb, 7, np.uint8); b = np.empty(px * db, np.uint8); b[:] = 0
	pa, ha = device.host_alloc(px * sb); pb, hb = device.host_alloc(px * db); ha[:] = 7
	t_pg = bench(lambda: device.convert(a.ctypes.data, b.ctypes.data, px, sc, dc, sf, df, mem=MEM_HOST))
	t_pin = bench(lambda: device.convert(pa, pb, px, sc, dc, sf, df, mem=MEM_HOST))
	t_mix1 = bench(lambda: device.convert(a.ctypes.data, pb, px, sc, dc, sf, df, mem=MEM_HOST))
	t_mix2 = bench(lambda: device.convert(pa, b.ctypes.data, px, sc, dc, sf, df, mem=MEM_HOST))
	print(f"{name:16s} 2048^2: pageable {t_pg:.3f} ms  pinned {t_pin:.3f} ms  src-pageable {t_mix1:.3f}  dst-pageable {t_mix2:.3f}", flush=True)
	device.host_free(pa); device.host_free(pb)
