"""CPU tests of the boundary: the C-ABI library builds, loads, exports every symbol that
include/cacao_img.h declares, validates arguments before touching CUDA, and refuses to compute
without a device (no CPU fallback).  No kernel runs here."""
import ctypes as C
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
	text = open(os.path.join(ROOT, "include", "cacao_img.h")).read()
	text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
	return sorted(set(re.findall(r"\b(cacao_img_[a-z0-9_]+)\s*\(", text)))


def test_header_compiles_as_plain_c(tmp_path):
	src = tmp_path / "t.c"
	src.write_text('#include "cacao_img.h"\nint main(void){return CACAO_IMG_ABI_VERSION - 1;}\n')
	subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_library_exports_every_declared_symbol(built_lib):
	from cacaoengine_b200 import _capi

	names = declared_symbols()
	assert len(names) >= 20
	assert sorted(_capi.SIGNATURES) == names, "python binding and header disagree"
	for n in names:
		assert hasattr(built_lib, n), n
	assert built_lib.cacao_img_abi_version() == 1


def test_library_has_no_torch_or_oracle_dependency(built_lib):
	from cacaoengine_b200 import _capi

	out = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
	assert "torch" not in out and "oracle" not in out and "libcudart" not in out  # cudart is linked statically


def test_sass_is_sm100a_only(built_lib):
	from cacaoengine_b200 import _capi

	out = subprocess.run(["cuobjdump", "--list-elf", _capi.LIB_PATH], capture_output=True, text=True).stdout
	archs = set(re.findall(r"sm_(\d+a?)", out))
	assert archs == {"100a"}, archs


def test_sass_free_of_known_ptxas_miscompile(built_lib):
	"""ptxas 12.9 turns 'cvt.rn.f16.f32 into a .b16 register, then take its low byte' into F2I.U8.F16,
	a value conversion of the half (first GPU run: byte-wise edge kernel wrote 0x3400 for 0x34F5).
	cacao_device.cuh routes every f32->f16 through a .b32 register; make sure the pattern stays gone."""
	from cacaoengine_b200 import _capi

	sass = subprocess.run(["cuobjdump", "-sass", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
	assert not re.search(r"F2I\.[US](8|16)\.F16", sass)
	assert "STG.E.128" in sass and "LDG.E.128" in sass  # 16-byte vectorised global access is really there


def test_status_strings_are_the_reference_messages(built_lib):
	from cacaoengine_b200 import _capi as c

	s = lambda code: built_lib.cacao_img_status_string(code).decode()
	assert s(c.E_SAME_LAYOUT) == "Cannot change an image's channel layout to its current layout!"  # Layout.cpp:9
	assert s(c.E_NOT_16BIT) == "Source image for 16 to 8-bit color conversion does not have a 16-bit color depth!"  # Colordepth.cpp:8
	assert s(c.E_ZERO_DIM) == "Cannot flip an image with zeroed dimensions!"  # Flip.cpp:8
	assert s(c.E_BAD_DEPTH) == "Invalid color depth; only 8 and 16 are allowed."  # Flip.cpp:9
	assert s(c.E_EMPTY) == "Cannot flip an image with a zero-sized data buffer!"  # Flip.cpp:10


def test_lut_in_library_is_the_reference_table(built_lib):
	from cacaoengine_b200 import device

	lut = device.get_lut()
	assert hashlib.sha256(lut.tobytes()).hexdigest() == "f5e5d4fca73d029ea9f1e5d5a3d26766d984ba685ba5d948d92af2c3a0344306"


def test_argument_errors_come_before_cuda(built_lib):
	from cacaoengine_b200 import _capi as c

	buf = np.zeros(64, np.uint8)
	p = buf.ctypes.data
	assert built_lib.cacao_img_convert(-1, p, p, 4, 3, 3, c.FMT_U8, c.FMT_U8, 0, c.MEM_HOST, None) == c.E_SAME_LAYOUT
	assert b"current layout" in built_lib.cacao_img_last_error()
	assert built_lib.cacao_img_convert(-1, p, p, 4, 2, 3, c.FMT_U8, c.FMT_U8, 0, c.MEM_HOST, None) == c.E_BAD_ARG
	assert built_lib.cacao_img_convert(-1, p, p, 4, 3, 4, 9, c.FMT_U8, 0, c.MEM_HOST, None) == c.E_BAD_ARG
	assert built_lib.cacao_img_flip_rows(-1, p, p, 0, 4, 4, c.MEM_HOST, None) == c.E_ZERO_DIM
	assert built_lib.cacao_img_equirect_to_cube(-1, p, 0, 4, 0, 4, 4, 0, p, 4, 0x3F, 0, 4, c.MEM_HOST, None) == c.E_ZERO_DIM


def test_python_mirror_raises_reference_messages(built_lib):
	import cacaoengine_b200 as ce

	img = ce.Image(2, 2, ce.Layout.RGB, 8, np.zeros(12, np.uint8))
	with pytest.raises(RuntimeError, match="Cannot change an image's channel layout to its current layout!"):
		ce.ChangeChannelLayout(img, ce.Layout.RGB)
	with pytest.raises(RuntimeError, match="does not have a 16-bit color depth"):
		ce.Convert16To8BitColor(img)
	with pytest.raises(RuntimeError, match="zeroed dimensions"):
		ce.Flip(ce.Image(0, 2, ce.Layout.RGB, 8, np.zeros(0, np.uint8)))


def test_empty_and_ragged_images(built_lib):
	"""Edge cases that need no device: a 0x0 image converts to an empty image (the reference loops zero
	times, Layout.cpp:31); a buffer shorter than w*h*channels -- which the reference would read past --
	is rejected before anything is launched."""
	import cacaoengine_b200 as ce

	empty = ce.Image(0, 0, ce.Layout.RGBA, 8, np.zeros(0, np.uint8))
	out = ce.ChangeChannelLayout(empty, ce.Layout.RGB)
	assert out.layout == ce.Layout.RGB and out.data.size == 0 and (out.w, out.h) == (0, 0)
	out = ce.Convert16To8BitColor(ce.Image(0, 0, ce.Layout.RGB, 16, np.zeros(0, np.uint8)))
	assert out.bitsPerChannel == 8 and out.data.size == 0
	ragged = ce.Image(4, 4, ce.Layout.RGB, 8, np.zeros(4 * 4 * 3 - 1, np.uint8))
	with pytest.raises(RuntimeError, match="smaller than w \\* h"):
		ce.ChangeChannelLayout(ragged, ce.Layout.RGBA)
	with pytest.raises(RuntimeError, match="smaller than w \\* h"):
		ce.Flip(ragged)
	with pytest.raises(RuntimeError, match="smaller than w \\* h"):
		ce.EquirectToCubemap(ce.ImageEx(8, 4, ce.Layout.RGB, ce.FMT_F32, np.zeros(10, np.uint8)), 4)


def test_no_cpu_fallback_without_a_device(built_lib):
	"""On a box without CUDA the product must fail loudly, never compute on the CPU."""
	from cacaoengine_b200 import _capi as c, device

	if device.device_count() > 0:
		pytest.skip("a CUDA device is present")
	src = np.arange(12, dtype=np.uint8)
	dst = np.zeros(16, np.uint8)
	rc = built_lib.cacao_img_convert(-1, src.ctypes.data, dst.ctypes.data, 4, 3, 4, c.FMT_U8, c.FMT_U8, 0, c.MEM_HOST, None)
	assert rc == c.E_NO_DEVICE
	assert (dst == 0).all()


def test_src_window_planner_covers_all_taps(built_lib, oracle):
	"""Host-side planning (pure arithmetic, no CUDA): every row the oracle's taps touch lies inside
	the window the library plans for that work unit."""
	from cacaoengine_b200 import device

	for W, H, N, bands in ((256, 128, 48, ((0, 48), (0, 12), (12, 30), (40, 48), (23, 25), (24, 25))), (300, 140, 37, ((0, 37), (0, 9), (9, 19), (18, 19), (17, 30), (36, 37))), (64, 32, 1, ((0, 1),))):
		for face in range(6):
			for (r0, r1) in bands:
				w0, wn = device.equirect_src_window(W, H, N, 1 << face, r0, r1)
				ys = np.array([oracle.project(face, i, j, N, W, H)[1] for j in range(r0, r1) for i in range(N)])
				lo = max(0, int(np.floor(ys.min())))
				hi = min(H - 1, int(np.floor(ys.max())) + 1)
				assert w0 <= lo and w0 + wn - 1 >= hi, (face, r0, r1, w0, wn, lo, hi)
				assert wn <= H
				# and it is tight: the taps' rows plus the planner's slack (one row each side, one for floor + 1)
				assert w0 >= lo - 2 and w0 + wn - 1 <= hi + 2, (face, r0, r1, w0, wn, lo, hi)


def test_fused_equirect_conversion_rejects_undefined_pairs(built_lib):
	"""Spec B.4 is defined for float sources, same channels or RGB -> RGBA; everything else is refused
	before any device work (so this runs without a GPU)."""
	from cacaoengine_b200 import _capi as c

	lib = c.lib()
	call = lambda ch, fmt, dch, dfmt: lib.cacao_img_equirect_to_cube_units_cvt(-1, 16, 8, 4, 0, 4, ch, fmt, 16, 2, dch, dfmt, 0, None, 0, None)
	for ch, fmt, dch, dfmt in ((4, c.FMT_U8, 4, c.FMT_F16), (4, c.FMT_U16, 4, c.FMT_U8), (3, c.FMT_F32, 1, c.FMT_F32), (1, c.FMT_F32, 4, c.FMT_F16), (4, c.FMT_F32, 1, c.FMT_U8), (3, c.FMT_U8, 4, c.FMT_U8)):
		assert call(ch, fmt, dch, dfmt) == c.E_BAD_ARG, (ch, fmt, dch, dfmt)
		assert b"not defined" in lib.cacao_img_last_error()
	assert call(2, c.FMT_F32, 2, c.FMT_F16) == c.E_BAD_ARG and call(4, c.FMT_F32, 4, 7) == c.E_BAD_ARG
	assert lib.cacao_img_equirect_to_cube_units_cvt(-1, 16, 8, 4, 0, 4, 4, c.FMT_F32, 16, 2, 4, c.FMT_F16, 2, None, 0, None) == c.E_BAD_ARG  # unknown flag
