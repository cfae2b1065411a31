"""dev: host -> host times of the two engine load chains, pinned and pageable buffers."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
from cacaoengine_b200 import FMT_U16, _capi as c, device
device.init(0)
def best(fn, reps=12):
	for _ in range(3): fn()
	ts = []
	for _ in range(reps):
		t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
	return min(ts) * 1e3, sorted(ts)[len(ts) // 2] * 1e3
for w, h in ((4096, 4096), (2048, 2048), (4000, 3000)):
	flags = device.CUBE_FLIP_ROWS | device.CUBE_MIPS
	n_in, n_out = w * h * 8, device.engine_texture_bytes(w, h, 4, flags)
	sp, s_host = device.host_alloc(n_in); dp, d_host = device.host_alloc(n_out)
	s_host[:] = np.random.default_rng(1).integers(0, 256, n_in, dtype=np.uint8)
	a = best(lambda: device.engine_texture(sp, w, h, 4, FMT_U16, dp, flags, mem=c.MEM_HOST))
	src = s_host.copy(); dst = np.empty(n_out, np.uint8)
	b = best(lambda: device.engine_texture(src.ctypes.data, w, h, 4, FMT_U16, dst.ctypes.data, flags, mem=c.MEM_HOST))
	same = bool((dst == d_host).all())
	print(f"texture {w}x{h} rgba16 -> rgba8 flipped + mips: pinned {a[0]:.3f} ms (median {a[1]:.3f})  pageable {b[0]:.3f} ms (median {b[1]:.3f})  same={same}  floor(upload at 55 GB/s) {n_in / 55e6:.3f} ms", flush=True)
	device.host_free(sp); device.host_free(dp)
w = h = 2048
for flags in (device.CUBE_FLIP_ROWS, device.CUBE_FLIP_ROWS | device.CUBE_MIPS):
	n_face, n_out = w * h * 8, device.engine_cube_bytes(w, h, flags)
	sp, s_host = device.host_alloc(6 * n_face); dp, d_host = device.host_alloc(n_out)
	s_host[:] = np.random.default_rng(2).integers(0, 256, 6 * n_face, dtype=np.uint8)
	faces = (C.c_void_p * 6)(*[sp + f * n_face for f in range(6)])
	a = best(lambda: c.check(c.lib().cacao_img_engine_cube(-1, faces, w, h, 4, FMT_U16, dp, flags, c.MEM_HOST, None)))
	print(f"cube 6x{w}^2 rgba16 -> rgb8 flags={flags}: pinned {a[0]:.3f} ms (median {a[1]:.3f})  floor {6 * n_face / 55e6:.3f} ms", flush=True)
	device.host_free(sp); device.host_free(dp)
