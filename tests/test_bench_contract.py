"""bench.py's output contract: one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args, timeout=600):
	r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
	return r


def test_reference_arm_line(oracle):
	"""--impl reference: the CPU implementation of the same config on the host cores (runs anywhere)."""
	r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
	assert r.returncode == 0, r.stderr
	lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
	assert len(lines) == 1
	d = json.loads(lines[0])
	assert BASE_KEYS <= set(d)
	assert d["impl"] == "reference" and d["unit"] == "Gpixel/s" and d["higher_is_better"] is True and d["value"] > 0
	assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
	assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
	assert "workload" in d["config"] and "model" not in d["config"]
	assert d["gpu_launches"] == 0
	# the same config object as the GPU arm prints, and every step the whole 256-frame batch
	sys.path.insert(0, ROOT)
	import bench

	assert d["config"] == bench.headline_config(1) and d["config"]["frames_per_gpu"] == 256
	assert "all 256 frames" in d["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_stay_silent(oracle):
	env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
	r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=120)
	assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device(built_lib):
	from cacaoengine_b200 import device

	if device.device_count() > 0:
		pytest.skip("a CUDA device is present")
	r = run_bench("--steps", "1", "--warmup", "1", "--kernel-only")
	assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_gpu_arm_line(built_lib, oracle):
	r = run_bench("--steps", "3", "--warmup", "3", "--no-extra")
	assert r.returncode == 0, r.stderr[-2000:]
	lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
	assert len(lines) == 1, lines
	d = json.loads(lines[0])
	assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(d)
	assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
	assert d["gpu_launches"] == 3  # one kernel launch per step
	roof = d["roofline"]
	assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
	assert 0.5 < roof["frac"] < 1.3  # a sane fraction of the measured copy peak
	e2e = d["e2e"]
	assert e2e["h2d_bytes_per_step"] == 1920 * 1080 * 4 * 256 and e2e["d2h_bytes_per_step"] == 1920 * 1080 * 8 * 256
	assert 0 < e2e["value"] < d["value"] and e2e["readback_verified"] is True and "oracle" in e2e["verified_against"]
	pcie = e2e["roofline"]
	assert pcie["bound"] == "pcie" and 0.4 < pcie["frac"] <= 1.5 and abs(pcie["frac"] - pcie["achieved"] / pcie["peak"]) < 1e-6
	sys.path.insert(0, ROOT)
	import bench

	assert d["config"] == bench.headline_config(1)
	cpu = d["cpu_baseline"]
	assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] > 0
	assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
