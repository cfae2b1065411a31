"""dev: how the cost model's weight of the +-Y rows moves the balance of the sharded config-5 job: every rank's
launch timed on one GPU with the polar cost table scaled by a factor (1.0 = as shipped)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cacaoengine_b200 import FMT_F32, device, sharding
device.init(0)
stream = torch.cuda.current_stream().cuda_stream
W, H, N = 16384, 8192, 4096
src = torch.empty(W * H * 16, dtype=torch.uint8, device="cuda"); dst = torch.empty(6 * N * N * 16, dtype=torch.uint8, device="cuda")
device.synth(src.data_ptr(), W * H * 4, FMT_F32, 4, 3, stream=stream)
def t(units, reps=30):
	for _ in range(5): device.equirect_to_cube_units(src.data_ptr(), W, H, 4, FMT_F32, dst.data_ptr(), N, units, stream=stream)
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): device.equirect_to_cube_units(src.data_ptr(), W, H, 4, FMT_F32, dst.data_ptr(), N, units, stream=stream)
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
whole = t([sharding.CubeUnit(f, 0, N) for f in range(6)])
print(f"whole cube, one launch: {whole:.4f} ms")
base = sharding._POLAR_COST
for world in (8, 4, 2):
	for scale in (1.0, 1.05, 1.1, 1.15, 1.2, 1.3):
		sharding._POLAR_COST = tuple((d, 1.0 + (c - 1.0) * scale + (scale - 1.0) * 0.0) for d, c in base) if "--excess" in sys.argv else tuple((d, c * scale) for d, c in base)
		lists = [sharding.units_for_rank(r, world, N) for r in range(world)]
		ms = [t(us) for us in lists]
		print(f"world={world} polar x{scale:.2f}: max {max(ms):.4f} mean {sum(ms)/world:.4f} ms  whole/world {whole/world:.4f}  eff {whole/world/max(ms):.3f}  per rank {[round(m, 4) for m in ms]}", flush=True)
sharding._POLAR_COST = base
