"""ctypes binding of include/cacao_img.h -- one Python function per C entry point, same names.

Loads cacaoengine_b200/lib/libcacaoimage_b200.so.  There is no fallback of any kind: a missing
library raises ImportError-like RuntimeError at first use, and compute calls on a box without a
CUDA device raise CacaoError(CACAO_E_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CACAO_IMG_LIB") or os.path.join(PKG, "lib", "libcacaoimage_b200.so")  # the override is for A/B runs of two builds (dev/)

# enum cacao_sample_fmt / cacao_mode / cacao_mem / cacao_status (include/cacao_img.h)
FMT_U8, FMT_U16, FMT_F16, FMT_F32 = 0, 1, 2, 3
MODE_REF_EXACT, MODE_FIXED = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
OK = 0
E_SAME_LAYOUT, E_NOT_16BIT, E_BAD_ARG, E_CUDA, E_NO_DEVICE, E_ZERO_DIM, E_BAD_DEPTH, E_EMPTY, E_NO_MEMORY = -1, -2, -3, -4, -5, -6, -7, -8, -9

FMT_BYTES = {FMT_U8: 1, FMT_U16: 2, FMT_F16: 2, FMT_F32: 4}

# every symbol include/cacao_img.h declares: name -> (restype, argtypes)
_vp, _u64, _u32, _i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
SIGNATURES = {
	"cacao_img_abi_version": (_i, []),
	"cacao_img_last_error": (C.c_char_p, []),
	"cacao_img_status_string": (C.c_char_p, [_i]),
	"cacao_img_device_count": (_i, []),
	"cacao_img_init": (_i, [_i]),
	"cacao_img_shutdown": (_i, []),
	"cacao_img_launch_count": (_u64, []),
	"cacao_img_malloc": (_i, [_i, C.POINTER(_vp), _u64]),
	"cacao_img_free": (_i, [_i, _vp]),
	"cacao_img_malloc_shareable": (_i, [_i, C.POINTER(_vp), _u64, C.POINTER(C.c_int), C.POINTER(_u64)]),
	"cacao_img_import_shareable": (_i, [_i, C.c_int, _u64, C.POINTER(_vp)]),
	"cacao_img_free_shareable": (_i, [_i, _vp]),
	"cacao_img_ipc_export": (_i, [_i, _vp, C.c_char_p, C.POINTER(_u64)]),
	"cacao_img_ipc_import": (_i, [_i, C.c_char_p, _u64, C.POINTER(_vp)]),
	"cacao_img_ipc_close": (_i, [_i, _vp, _u64]),
	"cacao_img_enable_peer_access": (_i, [_i, _i]),
	"cacao_img_host_alloc": (_i, [C.POINTER(_vp), _u64]),
	"cacao_img_host_free": (_i, [_vp]),
	"cacao_img_host_register": (_i, [_vp, _u64]),
	"cacao_img_host_unregister": (_i, [_vp]),
	"cacao_img_upload": (_i, [_i, _vp, _vp, _u64, _vp]),
	"cacao_img_download": (_i, [_i, _vp, _vp, _u64, _vp]),
	"cacao_img_sync": (_i, [_i, _vp]),
	"cacao_img_convert": (_i, [_i, _vp, _vp, _u64, _i, _i, _i, _i, _i, _i, _vp]),
	"cacao_img_convert_batch": (_i, [_i, _vp, _vp, _u64, _u64, _i, _i, _i, _i, _i, _i, _vp]),
	"cacao_img_convert_flip": (_i, [_i, _vp, _vp, _u32, _u32, _u64, _i, _i, _i, _i, _i, _i, _vp]),
	"cacao_img_engine_cube_bytes": (_u64, [_u32, _u32, _u32]),
	"cacao_img_engine_cube": (_i, [_i, C.POINTER(_vp), _u32, _u32, _i, _i, _vp, _u32, _i, _vp]),
	"cacao_img_engine_texture_bytes": (_u64, [_u32, _u32, _i, _u32]),
	"cacao_img_engine_texture": (_i, [_i, _vp, _u32, _u32, _i, _i, _vp, _u32, _i, _vp]),
	"cacao_img_shift_left_u16": (_i, [_i, _vp, _vp, _u64, _u32, _i, _vp]),
	"cacao_img_downsample2x": (_i, [_i, _vp, _vp, _u32, _u32, _i, _i, _i, _vp]),
	"cacao_img_unpack_texels": (_i, [_i, _vp, _vp, _u64, C.c_char_p, C.POINTER(C.c_int), _i, _vp]),
	"cacao_img_flip_rows": (_i, [_i, _vp, _vp, _u32, _u32, _u32, _i, _vp]),
	"cacao_img_equirect_to_cube": (_i, [_i, _vp, _u32, _u32, _u32, _u32, _i, _i, _vp, _u32, _u32, _u32, _u32, _i, _vp]),
	"cacao_img_equirect_to_cube_units": (_i, [_i, _vp, _u32, _u32, _u32, _u32, _i, _i, _vp, _u32, _vp, _u32, _vp]),
	"cacao_img_equirect_to_cube_units_host": (_i, [_i, _vp, _u32, _u32, _u32, _u32, _i, _i, C.POINTER(_vp), _u32, _vp, _u32]),
	"cacao_img_equirect_to_cube_units_cvt": (_i, [_i, _vp, _u32, _u32, _u32, _u32, _i, _i, _vp, _u32, _i, _i, _u32, _vp, _u32, _vp]),
	"cacao_img_equirect_to_cube_units_host_cvt": (_i, [_i, _vp, _u32, _u32, _u32, _u32, _i, _i, C.POINTER(_vp), _u32, _i, _i, _u32, _vp, _u32]),
	"cacao_img_equirect_src_window": (_i, [_u32, _u32, _u32, _u32, _u32, _u32, C.POINTER(_u32), C.POINTER(_u32)]),
	"cacao_img_equirect_fold_units": (_i, [_u32, _vp, _u32, _vp, _u32, C.POINTER(_u32)]),
	"cacao_img_synth": (_i, [_i, _vp, _u64, _u64, _i, _i, _u32, _vp]),
	"cacao_img_checksum": (_i, [_i, _vp, _u64, C.POINTER(_u64), _vp]),
	"cacao_img_get_lut": (_i, [_vp]),
}


class CacaoError(RuntimeError):
	"""Raised for any non-zero cacao_status; .status holds the code."""

	def __init__(self, status: int, message: str):
		super().__init__(message)
		self.status = status


_lib = None


def lib() -> C.CDLL:
	global _lib
	if _lib is None:
		if not os.path.exists(LIB_PATH):
			raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m cacaoengine_b200.build` "
							   "(libcacaoimage-b200 has no CPU or pure-Python fallback)")
		l = C.CDLL(LIB_PATH)
		for name, (res, args) in SIGNATURES.items():
			fn = getattr(l, name)
			fn.restype = res
			fn.argtypes = args
		_lib = l
	return _lib


def check(status: int) -> None:
	if status != OK:
		l = lib()
		msg = l.cacao_img_last_error().decode() or l.cacao_img_status_string(status).decode()
		raise CacaoError(status, msg)
