"""dev: config 5 (16384x8192 RGBA32F -> 6 x 4096^2) as a sharded job, every rank's launch timed on ONE
GPU: mirror-pair dealing with the four-fold kernel against the same lists computed two-fold
(CACAO_EQUIRECT_NO_QUAD=1), for 2 / 4 / 8 ranks."""
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
def t(units, reps=20):
	for _ in range(3): device.equirect_to_cube_units(src.data_ptr(), W, H, 4, FMT_F32, dst.data_ptr(), N, units, stream=stream)
	torch.cuda.synchronize()
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	for _ in range(reps): device.equirect_to_cube_units(src.data_ptr(), W, H, 4, FMT_F32, dst.data_ptr(), N, units, stream=stream)
	b.record(); torch.cuda.synchronize()
	return a.elapsed_time(b) / reps
whole = t([sharding.CubeUnit(f, 0, N) for f in range(6)])
print(f"whole cube, one launch: {whole:.4f} ms")
for world in (2, 4, 8):
	for bands in (None, sharding.bands_for_world(world) * 2):
		lists = [sharding.units_for_rank(r, world, N, bands) for r in range(world)]
		rows = [sharding.source_window(us, W, H, N)[1] for us in lists]
		for env in ("0", "1"):
			os.environ["CACAO_EQUIRECT_NO_QUAD"] = env
			ms = [t(us) for us in lists]
			print(f"world={world} bands={bands} {'two-fold ' if env == '1' else 'four-fold'}: max {max(ms):.4f} mean {sum(ms)/world:.4f} ms  whole/world {whole/world:.4f}  eff {whole/world/max(ms):.3f}  units/rank {len(lists[0])} src rows {max(rows)}  per rank {[round(m, 4) for m in ms]}")
		os.environ["CACAO_EQUIRECT_NO_QUAD"] = "0"
