"""Host-side mirror of libcacaoimage's image API (same names, argument meaning and errors).

Reference: libs/image/include/libcacaoimage.hpp:10-34 (struct Image), :187 Convert16To8BitColor,
:196 Flip, :213 ChangeChannelLayout.  Buffers are numpy arrays; every pixel loop is a call into the
C ABI with CACAO_MEM_HOST (upload, sm_100a kernel, readback).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field, replace

import numpy as np

from . import _capi as capi


class Layout(enum.IntEnum):  # libcacaoimage.hpp:15-19
	Grayscale = 1
	RGB = 3
	RGBA = 4


_NP_OF_FMT = {capi.FMT_U8: np.uint8, capi.FMT_U16: np.uint16, capi.FMT_F16: np.float16, capi.FMT_F32: np.float32}


@dataclass
class Image:
	"""libcacaoimage::Image -- `data` is a flat uint8 buffer exactly like std::vector<unsigned char>."""
	w: int
	h: int
	layout: Layout
	bitsPerChannel: int
	data: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))
	format: int = 0  # Image::Format::PNG
	quality: int = 100
	lossy: bool = False


@dataclass
class ImageEx:
	"""Extension carrier with an explicit sample format (U8/U16/F16/F32)."""
	w: int
	h: int
	layout: Layout
	format: int
	data: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))


def _bytes(a: np.ndarray) -> np.ndarray:
	return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


SHORT_BUFFER = "Image data buffer is smaller than w * h * channels * bytes per channel!"


def _need(buf: np.ndarray, nbytes: int) -> None:
	# the reference walks w*h pixels whatever the buffer holds (Layout.cpp:20,31); a short buffer is an
	# error here instead of an out-of-bounds read
	if buf.size < nbytes:
		raise RuntimeError(SHORT_BUFFER)


def _raise_ref(status: int) -> None:
	# the reference's std::runtime_error texts (cacao_img_status_string carries them verbatim)
	raise RuntimeError(capi.lib().cacao_img_status_string(status).decode())


def Convert16To8BitColor(src: Image) -> Image:
	if src.bitsPerChannel != 16:  # Colordepth.cpp:8
		_raise_ref(capi.E_NOT_16BIT)
	s = _bytes(src.data)
	ch = int(src.layout)
	out = np.empty(s.size // 2, np.uint8)
	pixels = s.size // 2 // ch
	if pixels:
		capi.check(capi.lib().cacao_img_convert(-1, s.ctypes.data, out.ctypes.data, pixels, ch, ch, capi.FMT_U16, capi.FMT_U8, capi.MODE_REF_EXACT, capi.MEM_HOST, None))
	return replace(src, bitsPerChannel=8, data=out)


def ChangeChannelLayout(src: Image, layout: Layout) -> Image:
	if int(src.layout) == int(layout):  # Layout.cpp:9
		_raise_ref(capi.E_SAME_LAYOUT)
	fmt = capi.FMT_U16 if src.bitsPerChannel == 16 else capi.FMT_U8
	s = _bytes(src.data)
	pixels = src.w * src.h
	_need(s, pixels * int(src.layout) * capi.FMT_BYTES[fmt])
	out = np.empty(pixels * int(layout) * capi.FMT_BYTES[fmt], np.uint8)
	if pixels:
		capi.check(capi.lib().cacao_img_convert(-1, s.ctypes.data, out.ctypes.data, pixels, int(src.layout), int(layout), fmt, fmt, capi.MODE_REF_EXACT, capi.MEM_HOST, None))
	return replace(src, layout=Layout(layout), data=out)


def Flip(src):
	"""Vertical mirror with the semantics Flip.cpp:6-42 intends; accepts Image or ImageEx."""
	if not (src.w > 0 and src.h > 0):  # Flip.cpp:8
		_raise_ref(capi.E_ZERO_DIM)
	if isinstance(src, Image):
		if src.bitsPerChannel not in (8, 16):  # Flip.cpp:9
			_raise_ref(capi.E_BAD_DEPTH)
		bpp = int(src.layout) * (src.bitsPerChannel // 8)
	else:
		bpp = int(src.layout) * capi.FMT_BYTES[src.format]
	s = _bytes(src.data)
	if s.size == 0:  # Flip.cpp:10
		_raise_ref(capi.E_EMPTY)
	if src.h == 1:  # Flip.cpp:13
		return replace(src, data=s.copy())
	_need(s, src.w * src.h * bpp)
	out = np.empty_like(s)
	capi.check(capi.lib().cacao_img_flip_rows(-1, s.ctypes.data, out.ctypes.data, src.w, src.h, bpp, capi.MEM_HOST, None))
	return replace(src, data=out)


def Convert(src: ImageEx, layout: Layout, fmt: int, mode: int = capi.MODE_REF_EXACT) -> ImageEx:
	"""Fused layout + sample-format conversion (DESIGN.md Spec B.1)."""
	s = _bytes(src.data)
	pixels = src.w * src.h
	_need(s, pixels * int(src.layout) * capi.FMT_BYTES[src.format])
	out = np.empty(pixels * int(layout) * capi.FMT_BYTES[fmt], np.uint8)
	capi.check(capi.lib().cacao_img_convert(-1, s.ctypes.data, out.ctypes.data, pixels, int(src.layout), int(layout), src.format, fmt, mode, capi.MEM_HOST, None))
	return ImageEx(src.w, src.h, Layout(layout), fmt, out)


def ToRGB8(src: Image) -> Image:
	"""engine/src/core/Cubemap.cpp:28-31 (16->8 then ->RGB) as one kernel."""
	if src.bitsPerChannel == 8 and src.layout == Layout.RGB:
		return src
	fmt = capi.FMT_U16 if src.bitsPerChannel == 16 else capi.FMT_U8
	s = _bytes(src.data)
	pixels = src.w * src.h
	_need(s, pixels * int(src.layout) * capi.FMT_BYTES[fmt])
	out = np.empty(pixels * 3, np.uint8)
	if pixels:
		capi.check(capi.lib().cacao_img_convert(-1, s.ctypes.data, out.ctypes.data, pixels, int(src.layout), 3, fmt, capi.FMT_U8, capi.MODE_REF_EXACT, capi.MEM_HOST, None))
	return replace(src, layout=Layout.RGB, bitsPerChannel=8, data=out)


def EquirectToCubemap(src: ImageEx, face_size: int, face_mask: int = 0x3F, row_begin: int = 0, row_end: int | None = None, device: int = -1) -> np.ndarray:
	"""Equirect -> cube buffer (6*N*N pixels, faces +X,-X,+Y,-Y,+Z,-Z); only the selected rows are written."""
	s = _bytes(src.data)
	bpp = int(src.layout) * capi.FMT_BYTES[src.format]
	_need(s, src.w * src.h * bpp)
	out = np.zeros(6 * face_size * face_size * bpp, np.uint8)
	capi.check(capi.lib().cacao_img_equirect_to_cube(device, s.ctypes.data, src.w, src.h, 0, src.h, int(src.layout), src.format, out.ctypes.data, face_size, face_mask, row_begin, face_size if row_end is None else row_end, capi.MEM_HOST, None))
	return out.view(_NP_OF_FMT[src.format])
