"""dev: headline conversion host -> host from pinned memory against the pipeline chunk size (CACAO_CHUNK_MIB)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cacaoengine_b200 import FMT_F16, FMT_U8, MEM_HOST, device
device.init(0)
PPI = 1920 * 1080; frames = 256; px = PPI * frames
hp_in, h_in = device.host_alloc(px * 4); hp_out, h_out = device.host_alloc(px * 8)
h_in[:] = 7
fn = lambda: device.convert_batch(hp_in, hp_out, PPI, frames, 4, 4, FMT_U8, FMT_F16, mem=MEM_HOST)
fn()
ts = []
for _ in range(5):
	t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
print(os.environ.get("CACAO_CHUNK_MIB"), f"min {min(ts)*1e3:.2f} ms  median {sorted(ts)[2]*1e3:.2f} ms  {px/min(ts)/1e9:.3f} Gpx/s")
