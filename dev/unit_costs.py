"""dev: cost of each (face, band) unit of config 5 on one GPU, and the per-rank sums for a dealing."""
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
def t(units, reps=10):
	for _ in range(2): device.equirect_to_cube_units(src.data_ptr(), W, H, 4, FMT_F32, dst.data_ptr(), N, units, stream=stream)
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): device.equirect_to_cube_units(src.data_ptr(), W, H, 4, FMT_F32, dst.data_ptr(), N, units, stream=stream)
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
for bands in (2, 4, 8):
	units = sharding.cube_units(N, bands, 2)
	cost = {(u.face, u.row_begin): t([u]) for u in units}
	print(f"bands={bands}: per-unit ms by face:")
	for f in range(6):
		print("  face", f, " ".join(f"{cost[(u.face,u.row_begin)]*1e3:6.1f}" for u in units if u.face == f), "us")
	for world in (2, 4, 8):
		sums = []
		for r in range(world):
			mine = sharding.units_for_rank(r, world, N, bands)
			sums.append((sum(cost[(u.face, u.row_begin)] for u in mine), t(mine)))
		print(f"  world={world}: per-rank sum-of-units ms {[round(s[0],4) for s in sums]}")
		print(f"  world={world}: per-rank one-launch   ms {[round(s[1],4) for s in sums]}  max {max(s[1] for s in sums):.4f}  ideal {sum(s[1] for s in sums)/world:.4f}")
print("whole cube one launch:", t([sharding.CubeUnit(f, 0, N) for f in range(6)]))
