"""dev: host->host equirect -> cubemap: the pipelined call against its three steps run one after the other."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cacaoengine_b200 import FMT_F32, FMT_U8, MEM_HOST, MEM_DEVICE, device
device.init(0)
def best(fn, reps=12):
	fn(); fn()
	ts = []
	for _ in range(reps):
		t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
	return min(ts)
for name, W, H, N, ch, fmt, b in (("cfg1 rgba8 2048x1024->512", 2048, 1024, 512, 4, FMT_U8, 1), ("cfg3 rgba32f 8192x4096->2048", 8192, 4096, 2048, 4, FMT_F32, 4), ("cfg5 rgba32f 16384x8192->4096", 16384, 8192, 4096, 4, FMT_F32, 4)):
	sb, db = W * H * ch * b, 6 * N * N * ch * b
	for kind in ("pinned", "pageable"):
		if kind == "pinned":
			sp, s = device.host_alloc(sb); dp, d = device.host_alloc(db)
		else:
			s = np.empty(sb, np.uint8); d = np.empty(db, np.uint8); sp, dp = s.ctypes.data, d.ctypes.data
		s[:] = 3
		da, dc = device.malloc(sb), device.malloc(db)
		t_up = best(lambda: (device.upload(da, s), device.sync()))
		t_k = best(lambda: (device.equirect_to_cube(da, W, H, ch, fmt, dc, N), device.sync()))
		t_dn = best(lambda: (device.download(d, dc), device.sync()))
		ref = d.copy()
		d[:] = 0
		t_pipe = best(lambda: device.equirect_to_cube(sp, W, H, ch, fmt, dp, N, mem=MEM_HOST))
		same = bool((d == ref).all())
		extra = ""
		if kind == "pinned" and "--sweep" in sys.argv:
			for cm in (256, 512, 1024, 2048, 4096, 8192):
				os.environ["CACAO_EQ_PIPE_MIN_KIB"] = "0"; os.environ["CACAO_EQ_CHUNK_KIB"] = str(cm)
				t = best(lambda: device.equirect_to_cube(sp, W, H, ch, fmt, dp, N, mem=MEM_HOST))
				extra += f"  chunk {cm} KiB: {t*1e3:.3f} ms{'' if (d == ref).all() else ' MISMATCH'}"
			os.environ.pop("CACAO_EQ_PIPE_MIN_KIB"); os.environ.pop("CACAO_EQ_CHUNK_KIB")
		print(f"{name:32s} {kind:8s}{extra} upload {t_up*1e3:8.2f}  kernel {t_k*1e3:7.3f}  readback {t_dn*1e3:8.2f}  sum {1e3*(t_up+t_k+t_dn):8.2f} ms   pipelined call {t_pipe*1e3:8.2f} ms  ({6*N*N/t_pipe/1e9:6.3f} Gpx/s)  identical={same}", flush=True)
		device.free(da); device.free(dc)
		if kind == "pinned":
			device.host_free(sp); device.host_free(dp)
