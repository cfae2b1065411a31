"""Device-buffer layer: thin wrappers over cacao_img_malloc / upload / download and the
device-pointer forms of the hot-path calls.  Pointers are plain ints (CUDA device addresses), so
torch tensors (`t.data_ptr()`) and cacao_img_malloc buffers both work; `stream` is a cudaStream_t
handle as an int (e.g. torch.cuda.current_stream().cuda_stream) or None for the default stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi


def device_count() -> int:
	return capi.lib().cacao_img_device_count()


def init(device: int = -1) -> None:
	capi.check(capi.lib().cacao_img_init(device))


def launch_count() -> int:
	return int(capi.lib().cacao_img_launch_count())


def malloc(nbytes: int, device: int = -1) -> int:
	p = C.c_void_p()
	capi.check(capi.lib().cacao_img_malloc(device, C.byref(p), nbytes))
	return int(p.value)


def free(ptr: int, device: int = -1) -> None:
	capi.check(capi.lib().cacao_img_free(device, ptr))


def malloc_shareable(nbytes: int, device: int = -1) -> tuple[int, int, int]:
	"""Device memory exported as a POSIX file descriptor: (device pointer, fd, mapped size).  See cacao_img.h."""
	p, fd, size = C.c_void_p(), C.c_int(-1), C.c_uint64(0)
	capi.check(capi.lib().cacao_img_malloc_shareable(device, C.byref(p), nbytes, C.byref(fd), C.byref(size)))
	return int(p.value), int(fd.value), int(size.value)


def import_shareable(fd: int, mapped_bytes: int, device: int = -1) -> int:
	p = C.c_void_p()
	capi.check(capi.lib().cacao_img_import_shareable(device, fd, mapped_bytes, C.byref(p)))
	return int(p.value)


def free_shareable(ptr: int, device: int = -1) -> None:
	capi.check(capi.lib().cacao_img_free_shareable(device, ptr))


def ipc_export(ptr: int, device: int = -1) -> tuple[bytes, int]:
	"""(64-byte handle, offset) of a device buffer for another process on the same box; see cacao_img.h."""
	buf, off = C.create_string_buffer(64), C.c_uint64(0)
	capi.check(capi.lib().cacao_img_ipc_export(device, ptr, buf, C.byref(off)))
	return buf.raw, int(off.value)


def ipc_import(handle: bytes, offset: int, device: int = -1) -> int:
	p = C.c_void_p()
	capi.check(capi.lib().cacao_img_ipc_import(device, C.create_string_buffer(handle, 64), offset, C.byref(p)))
	return int(p.value)


def ipc_close(ptr: int, offset: int, device: int = -1) -> None:
	capi.check(capi.lib().cacao_img_ipc_close(device, ptr, offset))


def enable_peer_access(device: int, peer: int) -> None:
	capi.check(capi.lib().cacao_img_enable_peer_access(device, peer))


def host_alloc(nbytes: int) -> tuple[int, np.ndarray]:
	"""Pinned host buffer: (address, uint8 numpy view).  Release with host_free(address)."""
	p = C.c_void_p()
	capi.check(capi.lib().cacao_img_host_alloc(C.byref(p), nbytes))
	arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))
	return int(p.value), arr


def host_free(ptr: int) -> None:
	capi.check(capi.lib().cacao_img_host_free(ptr))


def host_register(ptr: int, nbytes: int) -> None:
	"""Pin memory the caller owns (numpy array, shared-memory segment) for direct DMA; undo with host_unregister."""
	capi.check(capi.lib().cacao_img_host_register(ptr, nbytes))


def host_unregister(ptr: int) -> None:
	capi.check(capi.lib().cacao_img_host_unregister(ptr))


def upload(dst_dev: int, src: np.ndarray, device: int = -1, stream=None) -> None:
	s = np.ascontiguousarray(src)
	capi.check(capi.lib().cacao_img_upload(device, dst_dev, s.ctypes.data, s.nbytes, stream))
	if stream is None:
		sync(device)


def download(dst: np.ndarray, src_dev: int, device: int = -1, stream=None) -> None:
	assert dst.flags["C_CONTIGUOUS"]
	capi.check(capi.lib().cacao_img_download(device, dst.ctypes.data, src_dev, dst.nbytes, stream))
	sync(device, stream)


def sync(device: int = -1, stream=None) -> None:
	capi.check(capi.lib().cacao_img_sync(device, stream))


def convert(src: int, dst: int, pixels: int, src_ch: int, dst_ch: int, src_fmt: int, dst_fmt: int, mode: int = capi.MODE_REF_EXACT, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> None:
	capi.check(capi.lib().cacao_img_convert(device, src, dst, pixels, src_ch, dst_ch, src_fmt, dst_fmt, mode, mem, stream))


def convert_batch(src: int, dst: int, pixels_per_image: int, count: int, src_ch: int, dst_ch: int, src_fmt: int, dst_fmt: int, mode: int = capi.MODE_REF_EXACT, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> None:
	capi.check(capi.lib().cacao_img_convert_batch(device, src, dst, pixels_per_image, count, src_ch, dst_ch, src_fmt, dst_fmt, mode, mem, stream))


def shift_left_u16(src: int, dst: int, samples: int, shift: int, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> None:
	"""dst[k] = uint16(src[k] << shift): the JPEG decoder's 12 -> 16 bit widening (JPEG.cpp:79-82, shift 4)."""
	capi.check(capi.lib().cacao_img_shift_left_u16(device, src, dst, samples, shift, mem, stream))


def downsample2x(src: int, dst: int, w: int, h: int, ch: int, fmt: int, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> None:
	"""One mip level (2x2 box average): dst is max(1, w//2) x max(1, h//2)."""
	capi.check(capi.lib().cacao_img_downsample2x(device, src, dst, w, h, ch, fmt, mem, stream))


def unpack_texels(src: int, dst: int, texels: int, hint: str, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> int:
	"""Four-byte texels laid out as `hint` says -> 8-bit pixels in r,g,b,a order (Model.cpp:245-316); returns the channel count."""
	ch = C.c_int(0)
	capi.check(capi.lib().cacao_img_unpack_texels(device, src, dst, texels, hint.encode(), C.byref(ch), mem, stream))
	return int(ch.value)


def flip_rows(src: int, dst: int, w: int, h: int, bytes_per_pixel: int, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> None:
	capi.check(capi.lib().cacao_img_flip_rows(device, src, dst, w, h, bytes_per_pixel, mem, stream))


def convert_flip(src: int, dst: int, w: int, h: int, count: int, src_ch: int, dst_ch: int, src_fmt: int, dst_fmt: int, mode: int = capi.MODE_REF_EXACT, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> None:
	"""cacao_img_convert_flip: `count` images of w x h pixels converted and written bottom row first in one pass."""
	capi.check(capi.lib().cacao_img_convert_flip(device, src, dst, w, h, count, src_ch, dst_ch, src_fmt, dst_fmt, mode, mem, stream))


CUBE_FLIP_ROWS, CUBE_MIPS = 1, 2


def engine_cube_bytes(w: int, h: int, flags: int = 0) -> int:
	return int(capi.lib().cacao_img_engine_cube_bytes(w, h, flags))


def engine_cube(faces, w: int, h: int, ch: int, fmt: int, dst: int, flags: int = 0, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> None:
	"""cacao_img_engine_cube: six faces (addresses) -> one contiguous RGB8 cube (+ flipped rows, + mip chain)."""
	ptrs = (C.c_void_p * 6)(*[int(f) for f in faces])
	capi.check(capi.lib().cacao_img_engine_cube(device, ptrs, w, h, ch, fmt, dst, flags, mem, stream))


def engine_texture_bytes(w: int, h: int, ch: int, flags: int = 0) -> int:
	return int(capi.lib().cacao_img_engine_texture_bytes(w, h, ch, flags))


def engine_texture(src: int, w: int, h: int, ch: int, fmt: int, dst: int, flags: int = 0, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> None:
	"""cacao_img_engine_texture: the Tex2D load chain -- 16 -> 8 bit when needed (layout kept), row mirror, mip chain."""
	capi.check(capi.lib().cacao_img_engine_texture(device, src, w, h, ch, fmt, dst, flags, mem, stream))


def equirect_to_cube(src: int, W: int, H: int, ch: int, fmt: int, dst: int, N: int, face_mask: int = 0x3F, row_begin: int = 0, row_end: int | None = None, src_row0: int = 0, src_rows: int | None = None, mem: int = capi.MEM_DEVICE, device: int = -1, stream=None) -> None:
	capi.check(capi.lib().cacao_img_equirect_to_cube(device, src, W, H, src_row0, H if src_rows is None else src_rows, ch, fmt, dst, N, face_mask, row_begin, N if row_end is None else row_end, mem, stream))


def equirect_to_cube_units(src: int, W: int, H: int, ch: int, fmt: int, dst: int, N: int, units, src_row0: int = 0, src_rows: int | None = None, device: int = -1, stream=None) -> None:
	"""units: iterable of (face, row_begin, row_end) -- e.g. sharding.units_for_rank(...)."""
	flat = np.array([(u.face, u.row_begin, u.row_end) if hasattr(u, "face") else tuple(u) for u in units], np.uint32).reshape(-1, 3)
	capi.check(capi.lib().cacao_img_equirect_to_cube_units(device, src, W, H, src_row0, H if src_rows is None else src_rows, ch, fmt, dst, N, flat.ctypes.data, len(flat), stream))


def equirect_to_cube_units_host(src: int, W: int, H: int, ch: int, fmt: int, faces, N: int, units, src_row0: int = 0, src_rows: int | None = None, device: int = -1) -> None:
	"""Host-memory, pipelined form: src = host address of source rows [src_row0, src_row0 + src_rows),
	faces = six host addresses (0 for faces no unit names)."""
	flat = np.array([(u.face, u.row_begin, u.row_end) if hasattr(u, "face") else tuple(u) for u in units], np.uint32).reshape(-1, 3)
	ptrs = (C.c_void_p * 6)(*[int(f) if f else None for f in faces])
	capi.check(capi.lib().cacao_img_equirect_to_cube_units_host(device, src, W, H, src_row0, H if src_rows is None else src_rows, ch, fmt, ptrs, N, flat.ctypes.data, len(flat)))


def _flat_units(units):
	if units is None:
		return None, 0
	flat = np.array([(u.face, u.row_begin, u.row_end) if hasattr(u, "face") else tuple(u) for u in units], np.uint32).reshape(-1, 3)
	return flat, len(flat)


FLIP_ROWS = 1


def equirect_to_cube_cvt(src: int, W: int, H: int, ch: int, fmt: int, dst: int, N: int, dst_ch: int, dst_fmt: int, units=None, flags: int = 0, src_row0: int = 0, src_rows: int | None = None, device: int = -1, stream=None) -> None:
	"""Resampling + format / layout conversion in one pass (Spec B.4; float sources); flags=FLIP_ROWS
	stores the faces bottom row first (any format).  units=None: whole cube."""
	flat, n = _flat_units(units)
	capi.check(capi.lib().cacao_img_equirect_to_cube_units_cvt(device, src, W, H, src_row0, H if src_rows is None else src_rows, ch, fmt, dst, N, dst_ch, dst_fmt, flags, flat.ctypes.data if flat is not None else None, n, stream))


def equirect_to_cube_cvt_host(src: int, W: int, H: int, ch: int, fmt: int, faces, N: int, dst_ch: int, dst_fmt: int, units=None, flags: int = 0, src_row0: int = 0, src_rows: int | None = None, device: int = -1) -> None:
	"""Host-memory form: faces = six host addresses of N x N x dst_ch faces in dst_fmt."""
	flat, n = _flat_units(units)
	ptrs = (C.c_void_p * 6)(*[int(f) if f else None for f in faces])
	capi.check(capi.lib().cacao_img_equirect_to_cube_units_host_cvt(device, src, W, H, src_row0, H if src_rows is None else src_rows, ch, fmt, ptrs, N, dst_ch, dst_fmt, flags, flat.ctypes.data if flat is not None else None, n))


def equirect_src_window(W: int, H: int, N: int, face_mask: int, row_begin: int, row_end: int) -> tuple[int, int]:
	r0, n = C.c_uint32(), C.c_uint32()
	capi.check(capi.lib().cacao_img_equirect_src_window(W, H, N, face_mask, row_begin, row_end, C.byref(r0), C.byref(n)))
	return int(r0.value), int(n.value)


UNIT_MIRROR = 8


def equirect_fold_units(N: int, units) -> list[tuple[int, int, int]]:
	"""The launcher's folding of a unit list (host only): (face | UNIT_MIRROR, r0, r1) stands for rows
	[r0, r1) of the upper half and their mirror rows N-1-j, plain (face, r0, r1) for the rest."""
	flat = np.array([(u.face, u.row_begin, u.row_end) if hasattr(u, "face") else tuple(u) for u in units], np.uint32).reshape(-1, 3)
	n = C.c_uint32()
	capi.check(capi.lib().cacao_img_equirect_fold_units(N, flat.ctypes.data, len(flat), None, 0, C.byref(n)))
	out = np.zeros((max(1, n.value), 3), np.uint32)
	capi.check(capi.lib().cacao_img_equirect_fold_units(N, flat.ctypes.data, len(flat), out.ctypes.data, n.value, C.byref(n)))
	return [tuple(int(v) for v in row) for row in out[: n.value]]


def synth(dst: int, samples: int, fmt: int, ch: int, seed: int, first_sample: int = 0, device: int = -1, stream=None) -> None:
	capi.check(capi.lib().cacao_img_synth(device, dst, first_sample, samples, fmt, ch, seed, stream))


def checksum(src: int, nbytes: int, device: int = -1, stream=None) -> int:
	out = C.c_uint64()
	capi.check(capi.lib().cacao_img_checksum(device, src, nbytes, C.byref(out), stream))
	return int(out.value)


def get_lut() -> np.ndarray:
	out = np.empty(65536, np.uint8)
	capi.check(capi.lib().cacao_img_get_lut(out.ctypes.data))
	return out
