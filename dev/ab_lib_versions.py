"""dev: the same equirect launches through two builds of the library on the same box (A/B across commits)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
libs = {"now": C.CDLL(os.path.join(ROOT, "cacaoengine_b200/lib/libcacaoimage_b200.so"))}
for extra in sys.argv[1:]:  # other builds: name=path
	k, v = extra.split("=", 1)
	libs[k] = C.CDLL(os.path.join(ROOT, v))
u32, vp, i = C.c_uint32, C.c_void_p, C.c_int
for l in libs.values():
	l.cacao_img_equirect_to_cube.argtypes = [i, vp, u32, u32, u32, u32, i, i, vp, u32, u32, u32, u32, i, vp]
	l.cacao_img_synth.argtypes = [i, vp, C.c_uint64, C.c_uint64, i, i, u32, vp]
	assert l.cacao_img_init(0) == 0
def t(fn, reps):
	for _ in range(3): fn()
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): fn()
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
for name, W, H, N, ch, fmt, bps in (("cfg3", 8192, 4096, 2048, 4, 3, 4), ("cfg5", 16384, 8192, 4096, 4, 3, 4), ("cfg1", 2048, 1024, 512, 4, 0, 1)):
	src = torch.empty(W * H * ch * bps, dtype=torch.uint8, device="cuda"); dst = torch.empty(6 * N * N * ch * bps, dtype=torch.uint8, device="cuda")
	assert libs["now"].cacao_img_synth(0, src.data_ptr(), 0, W * H * ch, fmt, ch, 0xC0FFEE + 3, None) == 0
	torch.cuda.synchronize()
	for rnd in range(3):
		res = {}
		for key, l in libs.items():
			res[key] = t(lambda: l.cacao_img_equirect_to_cube(0, src.data_ptr(), W, H, 0, H, ch, fmt, dst.data_ptr(), N, 0x3F, 0, N, 1, None), 30 if name != "cfg5" else 10)
		print(f"{name} round {rnd}: " + "  ".join(f"{k} {v:.4f} ms" for k, v in res.items()), flush=True)
	del src, dst
