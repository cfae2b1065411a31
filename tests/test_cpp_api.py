"""The C++ drop-in API (include/libcacaoimage.hpp) compiled and run as a reference user would."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path, built_lib):
	from cacaoengine_b200 import build as b

	exe = tmp_path / "test_cpp_api"
	subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests/cpp/test_cpp_api.cpp"), "-o", str(exe), "-L" + b.LIBDIR, "-lcacaoimage_b200", "-Wl,-rpath," + b.LIBDIR, "-lpthread"], check=True)
	return exe


def test_cpp_api_compiles_and_links(tmp_path, built_lib):
	"""CPU: a reference-style translation unit builds against the header and links the library."""
	exe = _compile(tmp_path, built_lib)
	from cacaoengine_b200 import device

	if device.device_count() == 0:
		# no device: the first pixel call must throw (no CPU fallback), not compute
		r = subprocess.run([str(exe)], capture_output=True, text=True)
		assert r.returncode != 0
		assert "cpp api ok" not in r.stdout


@pytest.mark.gpu
def test_cpp_api_known_answers(tmp_path, built_lib):
	exe = _compile(tmp_path, built_lib)
	from cacaoengine_b200 import device

	have = min(device.device_count(), 8)
	r = subprocess.run([str(exe), str(max(3, have)), str(have)], capture_output=True, text=True)
	assert r.returncode == 0, r.stdout + r.stderr
	assert "cpp api ok" in r.stdout
