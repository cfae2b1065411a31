import time, numpy as np, torch
rt = torch.cuda.cudart()
torch.cuda.init()
for mb in (8, 64, 512, 2048):
	a = np.full(mb << 20, 1, np.uint8)
	t0 = time.perf_counter(); rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0); t1 = time.perf_counter()
	d = torch.empty(a.nbytes, dtype=torch.uint8, device="cuda")
	torch.cuda.synchronize(); t2 = time.perf_counter()
	rt.cudaMemcpy(d.data_ptr(), a.ctypes.data, a.nbytes, 1) if hasattr(rt, "cudaMemcpy") else d.copy_(torch.from_numpy(a), non_blocking=True)
	torch.cuda.synchronize(); t3 = time.perf_counter()
	rc2 = rt.cudaHostUnregister(a.ctypes.data); t4 = time.perf_counter()
	print(f"{mb:5d} MiB: register {1e3*(t1-t0):8.2f} ms ({a.nbytes/(t1-t0)/1e9:6.1f} GB/s) rc={rc}  h2d {1e3*(t3-t2):8.2f} ms ({a.nbytes/(t3-t2)/1e9:5.1f} GB/s)  unregister {1e3*(t4-t3):8.2f} ms")
# fresh (untouched) output buffer: np.empty pages are not faulted in yet
b = np.empty(2048 << 20, np.uint8)
t0 = time.perf_counter(); rt.cudaHostRegister(b.ctypes.data, b.nbytes, 0); t1 = time.perf_counter()
print(f"untouched 2048 MiB: register {1e3*(t1-t0):8.2f} ms ({b.nbytes/(t1-t0)/1e9:6.1f} GB/s)")
rt.cudaHostUnregister(b.ctypes.data)
# multi-thread memcpy speed
import threading
src = np.full(2048 << 20, 3, np.uint8); dst = np.empty_like(src); dst[:] = 0
for T in (1, 2, 4, 8, 16):
	n = src.nbytes // T
	def work(k): np.copyto(dst[k*n:(k+1)*n], src[k*n:(k+1)*n])
	ths = [threading.Thread(target=work, args=(k,)) for k in range(T)]
	t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; dt = time.perf_counter() - t0
	print(f"memcpy {T:2d} threads: {src.nbytes/dt/1e9:6.1f} GB/s")
