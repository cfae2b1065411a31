"""In-tree build of libcacaoimage_b200.so (sm_100a only) with plain nvcc.

    python -m cacaoengine_b200.build [--force] [--verbose]

Objects go to cacaoengine_b200/_build/, the library to cacaoengine_b200/lib/.  Both are
git-ignored but travel with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libcacaoimage_b200.so")
TOOL = os.path.join(LIBDIR, "ce-cubetool")

CU_SOURCES = ["convert_u8.cu", "convert_u16.cu", "convert_f16.cu", "convert_f32.cu", "equirect.cu", "equirect_rows.cu", "equirect_tma.cu", "mip.cu", "misc_kernels.cu", "cabi.cu"]
CXX_SOURCES = ["libcacaoimage.cpp"]
TOOL_SOURCES = ["tools/cubetool/main.cpp", "tools/cubetool/project.cpp", "tools/cubetool/create.cpp", "tools/cubetool/extract.cpp", "tools/cubetool/imageio.cpp", "tools/cubetool/vp8l.cpp", "tools/cubetool/xjc.cpp"]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: the float spec fixes where FMAs happen (explicit fmaf only).  No --use_fast_math:
# IEEE division / sqrt and no flush-to-zero are part of the spec.
NVCC_FLAGS = ARCH + (["-DCACAO_ROWS_SWEEP"] if os.environ.get("CACAO_ROWS_SWEEP") else []) + ["-O3", "-std=c++20", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
CXX_FLAGS = ["-O2", "-std=c++20", "-fPIC", "-ffp-contract=off", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc() -> str:
	for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
		if cand and os.path.exists(cand):
			return cand
	raise RuntimeError("nvcc not found: libcacaoimage_b200 cannot be built (there is no CPU fallback)")


def _newer(target: str, deps: list[str]) -> bool:
	if not os.path.exists(target):
		return False
	t = os.path.getmtime(target)
	return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _headers() -> list[str]:
	hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp", ".inc"))]
	inc = os.path.join(ROOT, "include")
	hs += [os.path.join(inc, f) for f in os.listdir(inc)]
	hs.append(os.path.abspath(__file__))
	return hs


def _run(cmd: list[str], verbose: bool) -> None:
	if verbose:
		print(" ".join(cmd), flush=True)
	r = subprocess.run(cmd, capture_output=True, text=True)
	if r.returncode != 0:
		raise RuntimeError("build step failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
	if verbose and (r.stdout or r.stderr):
		print(r.stdout + r.stderr, flush=True)


def build(force: bool = False, verbose: bool = False) -> str:
	os.makedirs(OBJ, exist_ok=True)
	os.makedirs(LIBDIR, exist_ok=True)
	nvcc = _nvcc()
	hdrs = _headers()
	jobs = []
	objs = []
	for src in CU_SOURCES:
		o = os.path.join(OBJ, src.replace(".cu", ".o"))
		objs.append(o)
		if force or not _newer(o, [os.path.join(CSRC, src)] + hdrs):
			jobs.append([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", o])
	for src in CXX_SOURCES:
		o = os.path.join(OBJ, src.replace(".cpp", ".o"))
		objs.append(o)
		if force or not _newer(o, [os.path.join(CSRC, src)] + hdrs):
			jobs.append(["g++"] + CXX_FLAGS + ["-c", os.path.join(CSRC, src), "-o", o])
	if jobs:
		with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
			list(ex.map(lambda c: _run(c, verbose), jobs))
	if jobs or force or not os.path.exists(LIB):
		# static cudart: the library has no load-time dependency beyond libstdc++ / libc
		_run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"], verbose)
	build_tool(force, verbose)
	return LIB


def build_tool(force: bool = False, verbose: bool = False) -> str | None:
	srcs = [os.path.join(ROOT, s) for s in TOOL_SOURCES]
	if not all(os.path.exists(s) for s in srcs):
		return None
	tool_hdrs = [os.path.join(ROOT, "tools/cubetool", f) for f in os.listdir(os.path.join(ROOT, "tools/cubetool")) if f.endswith(".hpp")] + [os.path.join(ROOT, "tools/toolutil.hpp")]
	if force or not _newer(TOOL, srcs + tool_hdrs + [LIB]):
		_run(["g++", "-O2", "-std=c++20", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tools")] + srcs + ["-o", TOOL, "-L" + LIBDIR, "-lcacaoimage_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread", "-ldl"], verbose)
	return TOOL


if __name__ == "__main__":
	print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
