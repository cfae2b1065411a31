"""Helpers shared by the -m gpu tests: every compute call goes through the C ABI."""
import numpy as np

from cacaoengine_b200 import _capi as capi
from cacaoengine_b200 import device

NP = {capi.FMT_U8: np.uint8, capi.FMT_U16: np.uint16, capi.FMT_F16: np.float16, capi.FMT_F32: np.float32}
FMTS = (capi.FMT_U8, capi.FMT_U16, capi.FMT_F16, capi.FMT_F32)
NAME = {capi.FMT_U8: "u8", capi.FMT_U16: "u16", capi.FMT_F16: "f16", capi.FMT_F32: "f32"}


GUARD = 256  # bytes of canary on both sides of every test buffer


class DevBuf:
	"""Device buffer with canaries: compute-sanitizer is closed on this GPU pool, so every test buffer
	carries 256 guard bytes before and after the payload and free() asserts no kernel touched them."""

	def __init__(self, nbytes, offset=0):
		self.nbytes = nbytes
		self.offset = offset
		self.total = GUARD + offset + nbytes + GUARD
		self.base = device.malloc(self.total)
		self.ptr = self.base + GUARD + offset
		self._lo = GUARD + offset
		device.upload(self.base, np.full(self._lo, 0xA5, np.uint8))
		device.upload(self.base + self._lo + nbytes, np.full(GUARD, 0xA5, np.uint8))

	def free(self):
		front, back = np.empty(self._lo, np.uint8), np.empty(GUARD, np.uint8)
		device.download(front, self.base)
		device.download(back, self.base + self._lo + self.nbytes)
		device.free(self.base)
		assert (front == 0xA5).all(), "kernel wrote BEFORE its buffer"
		assert (back == 0xA5).all(), "kernel wrote PAST its buffer"


def convert_host(src, sc, dc, sf, df, mode=capi.MODE_REF_EXACT, count=1):
	s = np.ascontiguousarray(src).reshape(-1)
	px = s.size // sc
	out = np.empty(px * dc, NP[df])
	capi.check(capi.lib().cacao_img_convert_batch(-1, s.ctypes.data, out.ctypes.data, px // count, count, sc, dc, sf, df, mode, capi.MEM_HOST, None))
	return out


def convert_device(src, sc, dc, sf, df, mode=capi.MODE_REF_EXACT, count=1, src_off=0, dst_off=0):
	s = np.ascontiguousarray(src).reshape(-1)
	px = s.size // sc
	out = np.empty(px * dc, NP[df])
	a, b = DevBuf(s.nbytes, src_off), DevBuf(out.nbytes, dst_off)
	try:
		device.upload(a.ptr, s)
		device.convert_batch(a.ptr, b.ptr, px // count, count, sc, dc, sf, df, mode)
		device.download(out, b.ptr)
	finally:
		a.free()
		b.free()
	return out


def random_samples(rng, n, fmt, specials=True):
	"""Random samples of a format; floats include out-of-range, tiny, huge and signed values (no NaN)."""
	if fmt == capi.FMT_U8:
		return rng.integers(0, 256, n).astype(np.uint8)
	if fmt == capi.FMT_U16:
		return rng.integers(0, 65536, n).astype(np.uint16)
	x = rng.uniform(-0.25, 1.25, n).astype(np.float32)
	if specials and n >= 16:
		x[:12] = np.float32([0.0, 1.0, -0.0, 0.5, 1e-7, 65504.0, -3.0, 2.5, 0.999999, 1e-3, np.inf, -np.inf])
	return x.astype(NP[fmt])


def same_bits(a, b):
	"""Bit-for-bit equality; for float arrays NaNs must sit at the same positions but their payload
	and sign are unspecified (x86 and the GPU pick different default NaNs for inf - inf)."""
	if a.dtype != b.dtype or a.shape != b.shape:
		return False
	if a.dtype.kind == "f":
		na, nb = np.isnan(a), np.isnan(b)
		if not (na == nb).all():
			return False
		ia, ib = a.view(f"u{a.itemsize}"), b.view(f"u{b.itemsize}")
		return bool((ia[~na] == ib[~nb]).all())
	return bool((a.view(np.uint8) == b.view(np.uint8)).all())
