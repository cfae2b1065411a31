"""dev: cost of every mirror pair of bands of config 5 (four-fold kernel), to fit sharding.unit_cost."""
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
	for _ in range(3): device.equirect_to_cube_units(src.data_ptr(), W, H, 4, FMT_F32, dst.data_ptr(), N, units, stream=stream)
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): device.equirect_to_cube_units(src.data_ptr(), W, H, 4, FMT_F32, dst.data_ptr(), N, units, stream=stream)
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
for bands in (8, 16):
	for f in (0, 2, 3, 4):
		edges = sharding.band_edges(N, bands)
		n = len(edges) - 1
		row = []
		for k in range(n // 2):
			pair = [sharding.CubeUnit(f, edges[k], edges[k + 1]), sharding.CubeUnit(f, edges[n - 1 - k], edges[n - k])]
			row.append(t(pair) * 1e3)
		print(f"bands={bands} face {f}: us per mirror pair, edge -> centre: " + " ".join(f"{v:6.1f}" for v in row) + f"   per 1000 rows: " + " ".join(f"{v / (2 * (edges[k+1]-edges[k])) * 1000:5.1f}" for k, v in enumerate(row)))
# several pairs in one launch: is the sum of parts a good predictor?
us = [sharding.CubeUnit(0, 0, 512), sharding.CubeUnit(0, 3584, 4096), sharding.CubeUnit(2, 1792, 2048), sharding.CubeUnit(2, 2048, 2304), sharding.CubeUnit(4, 1536, 2048), sharding.CubeUnit(4, 2048, 2560)]
print("three pairs one launch", t(us) * 1e3, "sum of parts", sum(t(us[k:k + 2]) for k in (0, 2, 4)) * 1e3)
