"""ctypes doorway to the CPU oracles.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (cacaoengine_b200) never does.

Two libraries:
  * port  -- oracle/_build/libcacao_oracle.so, our plain-C restatement (cacao_oracle.c)
  * ref   -- oracle/_ref/libcacaoimage_ref.so, the UNMODIFIED reference sources
             (/root/reference/libs/image/src/{Layout,Colordepth}.cpp + its LUTGen) behind
             oracle/ref_shim.cpp.  Exists only where oracle/Makefile could build it (or where the
             prebuilt file travelled with the repo snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libcacao_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcacaoimage_ref.so")
REFERENCE_ROOT = "/root/reference"

U8, U16, F16, F32 = 0, 1, 2, 3
MODE_REF_EXACT, MODE_FIXED = 0, 1
FMT_DTYPE = {U8: np.uint8, U16: np.uint16, F16: np.float16, F32: np.float32}
FMT_BYTES = {U8: 1, U16: 2, F16: 2, F32: 4}
FMT_NAME = {U8: "u8", U16: "u16", F16: "f16", F32: "f32"}

LUT_SHA256 = "f5e5d4fca73d029ea9f1e5d5a3d26766d984ba685ba5d948d92af2c3a0344306"


def build(ref: bool | None = None) -> None:
	"""Compile the checkers (idempotent).  ref=None: build _ref only if /root/reference exists."""
	subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
	if ref is None:
		ref = os.path.isdir(REFERENCE_ROOT)
	if ref:
		subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_port = None
_ref = None


def port() -> C.CDLL:
	global _port
	if _port is None:
		if not os.path.exists(PORT_SO):
			build(ref=False)
		lib = C.CDLL(PORT_SO)
		vp, u64, u32, i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
		lib.orc_build_lut.argtypes = [vp]
		lib.orc_lut.restype = vp
		lib.orc_change_layout.argtypes = [vp, vp, u64, i, i, i]
		lib.orc_depth16to8.argtypes = [vp, vp, u64]
		lib.orc_convert.argtypes = [vp, vp, u64, i, i, i, i, i]
		lib.orc_convert_mt.argtypes = [vp, vp, u64, i, i, i, i, i, i]
		lib.orc_flip_rows.argtypes = [vp, vp, u32, u32, u32]
		lib.orc_unpack_texels.argtypes = [vp, vp, C.c_uint64, C.c_char_p]
		lib.orc_downsample2x.argtypes = [vp, vp, u32, u32, C.c_int, C.c_int]
		lib.orc_equirect_to_cube.argtypes = [vp, u32, u32, i, i, vp, u32, u32, u32, u32, i]
		lib.orc_equirect_to_cube_f64.argtypes = [vp, u32, u32, i, i, vp, u32, u32, u32, u32]
		lib.orc_project.argtypes = [u32, u32, u32, u32, u32, u32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
		lib.orc_atan2.argtypes = [C.c_float, C.c_float]
		lib.orc_atan2.restype = C.c_float
		lib.orc_hash32.argtypes = [u32, u64]
		lib.orc_hash32.restype = u32
		lib.orc_synth.argtypes = [vp, u64, u64, i, i, u32]
		lib.orc_f32_to_f16.argtypes = [C.c_float]
		lib.orc_f32_to_f16.restype = C.c_uint16
		lib.orc_f16_to_f32.argtypes = [C.c_uint16]
		lib.orc_f16_to_f32.restype = C.c_float
		lib.orc_fnv1a64.argtypes = [vp, u64]
		lib.orc_fnv1a64.restype = u64
		lib.orc_max_threads.restype = i
		_port = lib
	return _port


def have_ref() -> bool:
	return os.path.exists(REF_SO)


def ref() -> C.CDLL:
	global _ref
	if _ref is None:
		lib = C.CDLL(REF_SO)
		vp, u32, i = C.c_void_p, C.c_uint32, C.c_int
		lib.ref_last_error.restype = C.c_char_p
		lib.ref_lut.restype = vp
		lib.ref_change_layout.argtypes = [vp, vp, u32, u32, i, i, i]
		lib.ref_depth16to8.argtypes = [vp, vp, u32, u32, i, i]
		lib.ref_batch_change_layout.argtypes = [vp, vp, u32, u32, i, i, i, i, i, C.POINTER(C.c_double)]
		lib.ref_batch_depth16to8.argtypes = [vp, vp, u32, u32, i, i, i, C.POINTER(C.c_double)]
		lib.ref_max_threads.restype = i
		if hasattr(lib, "ref_engine_cube"):
			lib.ref_engine_cube.argtypes = [vp, vp, u32, u32, i, i, i, i, C.POINTER(C.c_double)]
		_ref = lib
	return _ref


def _ptr(a: np.ndarray) -> int:
	assert a.flags["C_CONTIGUOUS"]
	return a.ctypes.data


# ---------------------------------------------------------------- port (restatement)
def lut() -> np.ndarray:
	out = np.empty(65536, np.uint8)
	port().orc_build_lut(_ptr(out))
	return out


def change_layout(src: np.ndarray, src_ch: int, dst_ch: int) -> np.ndarray:
	"""src: flat or (..., src_ch) array of uint8/uint16 samples."""
	bits = 8 if src.dtype == np.uint8 else 16
	s = np.ascontiguousarray(src).reshape(-1)
	px = s.size // src_ch
	dst = np.empty(px * dst_ch, s.dtype)
	rc = port().orc_change_layout(_ptr(s), _ptr(dst), px, src_ch, dst_ch, bits)
	if rc != 0:
		raise ValueError(f"orc_change_layout rc={rc}")
	return dst


def depth16to8(src: np.ndarray) -> np.ndarray:
	s = np.ascontiguousarray(src).view(np.uint8).reshape(-1)
	dst = np.empty(s.size // 2, np.uint8)
	port().orc_depth16to8(_ptr(s), _ptr(dst), s.size // 2)
	return dst


def convert(src: np.ndarray, src_ch: int, dst_ch: int, src_fmt: int, dst_fmt: int, mode: int = MODE_REF_EXACT,
			threads: int = 0) -> np.ndarray:
	s = np.ascontiguousarray(src).reshape(-1)
	assert s.dtype == FMT_DTYPE[src_fmt], (s.dtype, src_fmt)
	px = s.size // src_ch
	dst = np.empty(px * dst_ch, FMT_DTYPE[dst_fmt])
	if threads == 1:
		rc = port().orc_convert(_ptr(s), _ptr(dst), px, src_ch, dst_ch, src_fmt, dst_fmt, mode)
	else:
		rc = port().orc_convert_mt(_ptr(s), _ptr(dst), px, src_ch, dst_ch, src_fmt, dst_fmt, mode, threads)
	if rc != 0:
		raise ValueError(f"orc_convert rc={rc}")
	return dst


def flip_rows(src: np.ndarray, w: int, h: int, bpp: int) -> np.ndarray:
	s = np.ascontiguousarray(src).view(np.uint8).reshape(-1)
	dst = np.empty_like(s)
	rc = port().orc_flip_rows(_ptr(s), _ptr(dst), w, h, bpp)
	if rc != 0:
		raise ValueError(f"orc_flip_rows rc={rc}")
	return dst.view(src.dtype)


def downsample2x(src: np.ndarray, w: int, h: int, ch: int, fmt: int) -> np.ndarray:
	"""One mip level (Spec B.3): (max(1,h/2), max(1,w/2), ch) from (h, w, ch)."""
	s = np.ascontiguousarray(src).reshape(-1)
	assert s.dtype == FMT_DTYPE[fmt] and s.size == w * h * ch
	dst = np.empty(max(1, w // 2) * max(1, h // 2) * ch, FMT_DTYPE[fmt])
	rc = port().orc_downsample2x(_ptr(s), _ptr(dst), w, h, ch, fmt)
	if rc != 0:
		raise ValueError(f"orc_downsample2x rc={rc}")
	return dst


def unpack_texels(src: np.ndarray, hint: str) -> np.ndarray:
	"""(texels, out_ch) uint8 from (texels, 4) uint8 source bytes laid out as `hint` says (Model.cpp:245-316)."""
	s = np.ascontiguousarray(src, np.uint8).reshape(-1, 4)
	dst = np.zeros((s.shape[0], 4), np.uint8)
	ch = port().orc_unpack_texels(_ptr(s), _ptr(dst), s.shape[0], hint.encode())
	if ch < 0:
		raise ValueError(f"orc_unpack_texels rc={ch}")
	return dst.reshape(-1)[: s.shape[0] * ch].reshape(s.shape[0], ch).copy()


def equirect_to_cube(src: np.ndarray, W: int, H: int, ch: int, fmt: int, N: int, face_mask: int = 0x3F,
					 row_begin: int = 0, row_end: int | None = None, threads: int = 0,
					 out: np.ndarray | None = None) -> np.ndarray:
	s = np.ascontiguousarray(src).reshape(-1)
	assert s.dtype == FMT_DTYPE[fmt] and s.size == W * H * ch
	if out is None:
		out = np.zeros(6 * N * N * ch, FMT_DTYPE[fmt])
	rc = port().orc_equirect_to_cube(_ptr(s), W, H, ch, fmt, _ptr(out), N, face_mask, row_begin,
									 N if row_end is None else row_end, threads)
	if rc != 0:
		raise ValueError(f"orc_equirect_to_cube rc={rc}")
	return out


def equirect_to_cube_cvt(src: np.ndarray, W: int, H: int, ch: int, fmt: int, N: int, dst_ch: int, dst_fmt: int,
						 threads: int = 0) -> np.ndarray:
	"""Spec B.4 (resample + convert in one pass) stated as the composition it is defined to equal: the
	float source widened to fp32 (exact), Spec B.2 in fp32, then Spec B.1 from F32 to the destination
	format / layout.  Whole cube; nothing here is new arithmetic."""
	assert fmt in (F16, F32), "B.4 is defined for float sources"
	s = np.ascontiguousarray(src).reshape(-1)
	wide = s if fmt == F32 else convert(s, ch, ch, F16, F32, threads=threads)
	cube = equirect_to_cube(wide, W, H, ch, F32, N, threads=threads)
	if dst_ch == ch and dst_fmt == F32:
		return cube
	return convert(cube, ch, dst_ch, F32, dst_fmt, threads=threads)


def equirect_to_cube_f64(src: np.ndarray, W: int, H: int, ch: int, fmt: int, N: int, face_mask: int = 0x3F) -> np.ndarray:
	s = np.ascontiguousarray(src).reshape(-1)
	out = np.zeros(6 * N * N * ch, np.float64)
	rc = port().orc_equirect_to_cube_f64(_ptr(s), W, H, ch, fmt, _ptr(out), N, face_mask, 0, N)
	if rc != 0:
		raise ValueError(f"orc_equirect_to_cube_f64 rc={rc}")
	return out


def project(face: int, i: int, j: int, N: int, W: int, H: int) -> tuple[float, float]:
	xs, ys = C.c_float(), C.c_float()
	port().orc_project(face, i, j, N, W, H, C.byref(xs), C.byref(ys))
	return xs.value, ys.value


def synth(samples: int, fmt: int, ch: int, seed: int, first_sample: int = 0) -> np.ndarray:
	out = np.empty(samples, FMT_DTYPE[fmt])
	port().orc_synth(_ptr(out), first_sample, samples, fmt, ch, seed)
	return out


def fnv1a64(a: np.ndarray) -> int:
	b = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
	return int(port().orc_fnv1a64(_ptr(b), b.size))


# ---------------------------------------------------------------- ref (compiled reference)
def ref_lut() -> np.ndarray:
	p = ref().ref_lut()
	return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(65536,)).copy()


def ref_change_layout(src: np.ndarray, w: int, h: int, src_ch: int, dst_ch: int) -> np.ndarray:
	bits = 8 if src.dtype == np.uint8 else 16
	s = np.ascontiguousarray(src).reshape(-1)
	dst = np.empty(w * h * dst_ch, s.dtype)
	rc = ref().ref_change_layout(_ptr(s), _ptr(dst), w, h, src_ch, dst_ch, bits)
	if rc != 0:
		raise RuntimeError(ref().ref_last_error().decode())
	return dst


def ref_engine_cube(faces: np.ndarray, w: int, h: int, ch: int, flip: bool = False, threads: int = 6) -> tuple[np.ndarray, float]:
	"""Six faces back to back through the reference's own load chain (Cubemap.cpp:28-31: Convert16To8BitColor,
	ChangeChannelLayout -> RGB), a row mirror with Flip's intended semantics when `flip`, the six results
	contiguous.  Returns (cube bytes, seconds of the chain)."""
	bits = 8 if faces.dtype == np.uint8 else 16
	s = np.ascontiguousarray(faces).reshape(-1)
	assert s.size == 6 * w * h * ch
	dst = np.empty(6 * w * h * 3, np.uint8)
	secs = C.c_double()
	rc = ref().ref_engine_cube(_ptr(s), _ptr(dst), w, h, ch, bits, 1 if flip else 0, threads, C.byref(secs))
	if rc != 0:
		raise RuntimeError("ref_engine_cube failed")
	return dst, secs.value


def engine_cube(faces: np.ndarray, w: int, h: int, ch: int, flip: bool = False, mips: bool = False) -> np.ndarray:
	"""The same chain from the restatement (depth16to8, change_layout, row mirror, downsample2x per level):
	level 0 = six RGB8 faces back to back, then every mip level face-major."""
	f = np.ascontiguousarray(faces).reshape(6, -1)
	out = []
	for k in range(6):
		img = f[k]
		if img.dtype == np.uint16:
			img = depth16to8(img)
		if ch != 3:
			img = change_layout(img, ch, 3)
		img = img.reshape(h, w * 3)
		out.append(np.ascontiguousarray(img[::-1] if flip else img).reshape(-1))
	levels = [np.concatenate(out)]
	lw, lh = w, h
	while mips and (lw > 1 or lh > 1):
		out = [downsample2x(o, lw, lh, 3, U8) for o in out]
		lw, lh = max(1, lw // 2), max(1, lh // 2)
		levels.append(np.concatenate(out))
	return np.concatenate(levels)


def engine_texture(img: np.ndarray, w: int, h: int, ch: int, flip: bool = False, mips: bool = False) -> np.ndarray:
	"""The Tex2D load chain from the restatement: depth16to8 when 16-bit (layout kept, Tex2D.cpp:18), row mirror
	(OpenGLTex2D.cpp:16), downsample2x per level (the GL path's glGenerateMipmap): level 0, then every mip level."""
	s = np.ascontiguousarray(img).reshape(-1)
	if s.dtype == np.uint16:
		s = depth16to8(s)
	s = s.reshape(h, w * ch)
	level = np.ascontiguousarray(s[::-1] if flip else s).reshape(-1)
	out, lw, lh = [level], w, h
	while mips and (lw > 1 or lh > 1):
		level = downsample2x(level, lw, lh, ch, U8)
		lw, lh = max(1, lw // 2), max(1, lh // 2)
		out.append(level)
	return np.concatenate(out)


def ref_depth16to8(src: np.ndarray, w: int, h: int, ch: int) -> np.ndarray:
	bits = 8 if src.dtype == np.uint8 else 16
	s = np.ascontiguousarray(src).reshape(-1)
	dst = np.empty(w * h * ch, np.uint8)
	rc = ref().ref_depth16to8(_ptr(s), _ptr(dst), w, h, ch, bits)
	if rc != 0:
		raise RuntimeError(ref().ref_last_error().decode())
	return dst
