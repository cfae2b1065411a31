"""GPU parity: the CUDA conversion path (through the C ABI) against the CPU oracle, bit for bit.

Integer ops are reference parity (the oracle is pinned to the compiled reference, test_oracle.py);
float-format ops are parity with our CPU restatement of DESIGN.md Spec B.1 -- the reference has no
float formats, so those are "parity unpinned vs reference".
"""
import itertools
import os

import numpy as np
import pytest

from gpu_util import FMTS, NAME, NP, DevBuf, convert_device, convert_host, random_samples, same_bits

pytestmark = pytest.mark.gpu

CH = (1, 3, 4)
SHAPES = [(sf, df, sc, dc) for sf, df, sc, dc in itertools.product(FMTS, FMTS, CH, CH) if not (sf == df and sc == dc)]
# 1 pixel; under one tile; tiles + tail; many tiles with a partial unroll group
SIZES = (1, 17, 1000, 40961)


@pytest.fixture(scope="module", autouse=True)
def _lib(built_lib):
	from cacaoengine_b200 import device

	device.init()
	yield


@pytest.mark.parametrize("sf,df,sc,dc", SHAPES, ids=[f"{NAME[a]}x{c}-{NAME[b]}x{d}" for a, b, c, d in SHAPES])
def test_every_shape_matches_oracle(oracle, sf, df, sc, dc):
	rng = np.random.default_rng(hash((sf, df, sc, dc)) & 0xFFFF)
	for mode in (0, 1):
		for px in SIZES:
			src = random_samples(rng, px * sc, sf)
			want = oracle.convert(src, sc, dc, sf, df, mode)
			got_d = convert_device(src, sc, dc, sf, df, mode)
			assert same_bits(got_d, want), (mode, px, "device")
			if px in (17, 40961):
				got_h = convert_host(src, sc, dc, sf, df, mode)
				assert same_bits(got_h, want), (mode, px, "host")


@pytest.mark.parametrize("sf,df,sc,dc", SHAPES, ids=[f"{NAME[a]}x{c}-{NAME[b]}x{d}" for a, b, c, d in SHAPES])
def test_every_shape_on_both_vector_kernels(oracle, monkeypatch, sf, df, sc, dc):
	"""The 128-bit tile kernel (convert_tiles.cuh) serves shapes the 256-bit sector kernel does not fit
	and buffers that are only 16-byte aligned; CACAO_NO_WIDE pins every shape onto it, and a
	16-byte-offset destination reaches it the natural way."""
	rng = np.random.default_rng(1 + (hash((sf, df, sc, dc)) & 0xFFFF))
	px = 40961
	src = random_samples(rng, px * sc, sf)
	want = oracle.convert(src, sc, dc, sf, df, 1)
	monkeypatch.setenv("CACAO_NO_WIDE", "1")
	assert same_bits(convert_device(src, sc, dc, sf, df, 1), want)
	monkeypatch.delenv("CACAO_NO_WIDE")
	# ... and CACAO_FORCE_WIDE pins every shape that fits onto the 256-bit sector kernel
	monkeypatch.setenv("CACAO_FORCE_WIDE", "1")
	assert same_bits(convert_device(src, sc, dc, sf, df, 1), want)
	assert same_bits(convert_device(src[: 1000 * sc], sc, dc, sf, df, 0), oracle.convert(src[: 1000 * sc], sc, dc, sf, df, 0))
	monkeypatch.delenv("CACAO_FORCE_WIDE")
	assert same_bits(convert_device(src, sc, dc, sf, df, 1, dst_off=16), want)
	assert same_bits(convert_device(src, sc, dc, sf, df, 1, src_off=16), want)


def test_golden_reference_vectors(golden):
	"""Outputs of the UNMODIFIED reference (tests/golden/ref_golden.npz) reproduced on the GPU."""
	from cacaoengine_b200 import _capi as c

	n = 0
	for key in golden.files:
		if not key.endswith("_in"):
			continue
		base = key[:-3]
		src, want = golden[key], golden[base + "_out"]
		if base.startswith("layout_"):
			_, bits, pair, _ = base.split("_")
			sc, dc = (int(x) for x in pair.split("to"))
			fmt = c.FMT_U8 if bits == "8" else c.FMT_U16
			got = convert_device(src, sc, dc, fmt, fmt)
			got_h = convert_host(src, sc, dc, fmt, fmt)
		else:
			ch = 4 if "rgba" in base else 1
			got = convert_device(src, ch, ch, c.FMT_U16, c.FMT_U8)
			got_h = convert_host(src, ch, ch, c.FMT_U16, c.FMT_U8)
		assert (got == want).all() and (got_h == want).all(), base
		n += 1
	assert n >= 26


def test_gray_weights_exhaustive(golden):
	"""All 65536 inputs through each BT.709 weight: the integer form + 1-in-625 double path must equal
	the reference's truncated double products (Layout.cpp:70-72)."""
	from cacaoengine_b200 import _capi as c

	allv = np.arange(65536, dtype=np.uint16)
	for k, name in enumerate("rgb"):
		px = np.zeros((65536, 3), np.uint16)
		px[:, k] = allv
		got = convert_device(px.reshape(-1), 3, 1, c.FMT_U16, c.FMT_U16)
		assert (got == golden[f"gray16_{name}_out"]).all(), name
	# 8-bit: all 2^24 RGB triples would be overkill; all values per channel + a random cloud
	for k in range(3):
		px = np.zeros((256, 4), np.uint8)
		px[:, k] = np.arange(256)
		want_k = (np.float64([0.2126, 0.7152, 0.0722][k]) * np.arange(256)).astype(np.uint8)
		assert (convert_device(px.reshape(-1), 4, 1, c.FMT_U8, c.FMT_U8) == want_k).all()


def test_unorm_to_float_exhaustive(oracle):
	from cacaoengine_b200 import _capi as c

	for fmt, n in ((c.FMT_U8, 256), (c.FMT_U16, 65536)):
		v = np.arange(n).astype(NP[fmt])
		f32 = convert_device(v, 1, 3, fmt, c.FMT_F32).reshape(-1, 3)
		assert same_bits(f32[:, 0].copy(), v.astype(np.float32) / np.float32(n - 1))
		f16 = convert_device(v, 1, 3, fmt, c.FMT_F16).reshape(-1, 3)
		assert same_bits(f16[:, 2].copy(), (v.astype(np.float32) / np.float32(n - 1)).astype(np.float16))
		assert same_bits(convert_device(v, 1, 3, fmt, c.FMT_F32), oracle.convert(v, 1, 3, fmt, c.FMT_F32))


def test_float_to_unorm_rounding_boundaries(oracle):
	from cacaoengine_b200 import _capi as c

	# values straddling every u8 rounding boundary (k + 0.5)/255 by a few ulps
	k = np.arange(255, dtype=np.float64)
	mid = ((k + 0.5) / 255.0).astype(np.float32)
	pts = np.concatenate([np.nextafter(mid, np.float32(0)), mid, np.nextafter(mid, np.float32(1)), np.float32([np.nan, -1e30, 1e30])])
	src = np.repeat(pts, 3)
	assert same_bits(convert_device(src, 3, 4, c.FMT_F32, c.FMT_U8), oracle.convert(src, 3, 4, c.FMT_F32, c.FMT_U8))
	assert same_bits(convert_device(src, 3, 4, c.FMT_F32, c.FMT_U16), oracle.convert(src, 3, 4, c.FMT_F32, c.FMT_U16))


def test_batch_pixel0_quirk_is_per_image(oracle):
	"""Gray->RGBA in REF_EXACT mode repeats pixel 0 of EACH image (Layout.cpp:50-51)."""
	from cacaoengine_b200 import _capi as c

	rng = np.random.default_rng(9)
	count, ppi = 5, 700
	src = rng.integers(0, 256, count * ppi).astype(np.uint8)
	want = np.concatenate([oracle.change_layout(src[k * ppi:(k + 1) * ppi], 1, 4) for k in range(count)])
	assert (convert_device(src, 1, 4, c.FMT_U8, c.FMT_U8, count=count) == want).all()
	assert (convert_host(src, 1, 4, c.FMT_U8, c.FMT_U8, count=count) == want).all()
	fixed = convert_device(src, 1, 4, c.FMT_U8, c.FMT_U8, mode=c.MODE_FIXED, count=count).reshape(-1, 4)
	assert (fixed[:, 0] == src).all() and (fixed[:, 3] == 255).all()


def test_unaligned_device_pointers_take_the_edge_path(oracle):
	from cacaoengine_b200 import _capi as c

	rng = np.random.default_rng(10)
	src = rng.integers(0, 256, 5000 * 3).astype(np.uint8)
	want = oracle.change_layout(src, 3, 4)
	assert (convert_device(src, 3, 4, c.FMT_U8, c.FMT_U8, src_off=1) == want).all()
	assert (convert_device(src, 3, 4, c.FMT_U8, c.FMT_U8, dst_off=4) == want).all()


def test_fused_engine_composition(oracle):
	"""RGBA16 -> RGB8 in one kernel == Convert16To8BitColor then ChangeChannelLayout (Cubemap.cpp:28-31)."""
	from cacaoengine_b200 import _capi as c

	rng = np.random.default_rng(12)
	src = rng.integers(0, 65536, 1024 * 64 * 4).astype(np.uint16)
	two_step = oracle.change_layout(oracle.depth16to8(src), 4, 3)
	assert (convert_device(src, 4, 3, c.FMT_U16, c.FMT_U8) == two_step).all()


def test_python_mirror_of_reference_api(oracle):
	import cacaoengine_b200 as ce

	rng = np.random.default_rng(13)
	w, h = 67, 29
	data = rng.integers(0, 65536, w * h * 4).astype(np.uint16)
	img = ce.Image(w, h, ce.Layout.RGBA, 16, data.view(np.uint8), format=2, quality=55, lossy=True)
	out8 = ce.Convert16To8BitColor(img)
	assert (out8.w, out8.h, out8.layout, out8.bitsPerChannel, out8.format, out8.quality, out8.lossy) == (w, h, ce.Layout.RGBA, 8, 2, 55, True)
	assert (out8.data == oracle.depth16to8(data)).all()
	rgb = ce.ChangeChannelLayout(out8, ce.Layout.RGB)
	assert rgb.layout == ce.Layout.RGB and (rgb.data == oracle.change_layout(out8.data, 4, 3)).all()
	assert (ce.ToRGB8(img).data == rgb.data).all()
	gray = ce.ChangeChannelLayout(img, ce.Layout.Grayscale)
	assert (gray.data.view(np.uint16) == oracle.change_layout(data, 4, 1)).all()
	fl = ce.Flip(rgb)
	assert (fl.data.reshape(h, -1) == rgb.data.reshape(h, -1)[::-1]).all() and (rgb.data == oracle.change_layout(out8.data, 4, 3)).all()
	ex = ce.Convert(ce.ImageEx(w, h, ce.Layout.RGBA, ce.FMT_U16, data.view(np.uint8)), ce.Layout.RGB, ce.FMT_F16)
	assert same_bits(ex.data.view(np.float16), oracle.convert(data, 4, 3, oracle.U16, oracle.F16))


@pytest.mark.parametrize("w,h,bpp", [(1, 1, 4), (5, 2, 3), (64, 33, 4), (67, 29, 6), (128, 128, 16), (3, 1000, 1)])
def test_flip_rows(oracle, w, h, bpp):
	from cacaoengine_b200 import _capi as c, device
	from gpu_util import DevBuf

	rng = np.random.default_rng(w * h)
	src = rng.integers(0, 256, w * h * bpp).astype(np.uint8)
	want = oracle.flip_rows(src, w, h, bpp)
	out = np.empty_like(src)
	c.check(c.lib().cacao_img_flip_rows(-1, src.ctypes.data, out.ctypes.data, w, h, bpp, c.MEM_HOST, None))
	assert (out == want).all()
	a, b = DevBuf(src.nbytes), DevBuf(src.nbytes)
	device.upload(a.ptr, src)
	device.flip_rows(a.ptr, b.ptr, w, h, bpp)
	out2 = np.empty_like(src)
	device.download(out2, b.ptr)
	a.free(), b.free()
	assert (out2 == want).all()
	# involution: flipping twice is the identity
	assert (oracle.flip_rows(out2, w, h, bpp) == src).all()


@pytest.mark.parametrize("sf,df,sc,dc", SHAPES, ids=[f"{NAME[a]}x{c}-{NAME[b]}x{d}" for a, b, c, d in SHAPES])
def test_convert_flip_every_shape(oracle, sf, df, sc, dc):
	"""cacao_img_convert_flip == convert, then mirror the rows of every image: widths the sector kernel's
	run divides (64, 96), widths only the tile kernel could take or nothing divides (1024 / 100 -> tile or
	byte-wise kernel), batches of images, one-row images, device and host memory."""
	from cacaoengine_b200 import _capi as c, device

	rng = np.random.default_rng(hash((sf, df, sc, dc, 7)) & 0xFFFF)
	for w, h, count, mode in ((64, 5, 3, 0), (96, 4, 1, 1), (100, 7, 2, 0), (1024, 3, 2, 0), (33, 1, 2, 0)):
		src = random_samples(rng, w * h * count * sc, sf)
		plain = oracle.convert(src, sc, dc, sf, df, mode) if not (sc == 1 and dc == 4 and mode == 0 and sf in (0, 1)) else np.concatenate([oracle.convert(src[k * w * h * sc:(k + 1) * w * h * sc], sc, dc, sf, df, mode) for k in range(count)])
		want = plain.reshape(count, h, w * dc)[:, ::-1, :].copy().reshape(-1)
		a, b = DevBuf(src.nbytes), DevBuf(want.nbytes)
		device.upload(a.ptr, src)
		device.convert_flip(a.ptr, b.ptr, w, h, count, sc, dc, sf, df, mode)
		got = np.empty_like(want)
		device.download(got, b.ptr)
		a.free(), b.free()
		assert same_bits(got, want), (w, h, count, mode, "device")
		if w in (64, 1024):
			got_h = np.empty_like(want)
			device.convert_flip(src.ctypes.data, got_h.ctypes.data, w, h, count, sc, dc, sf, df, mode, mem=c.MEM_HOST)
			assert same_bits(got_h, want), (w, h, count, mode, "host")


@pytest.mark.parametrize("form", ["two", "direct", "lanes"])
def test_every_form_of_the_depth_table(oracle, golden, monkeypatch, form):
	"""The 16 -> 8 bit table sits in shared memory in one of three forms (convert_pixel.cuh: two-level 8 KiB,
	direct 64 KiB, bank-private log-indexed 80.5 KiB); CACAO_LUT_MODE pins one.  Each must give the compiled
	reference's output for all 65,536 inputs (tests/golden) and the oracle's for every layout pair, plain and
	with the row mirror folded in."""
	from cacaoengine_b200 import _capi as c, device

	monkeypatch.setenv("CACAO_LUT_MODE", form)
	src_all, want_all = golden["depth_all_in"], golden["depth_all_out"]
	assert (convert_device(src_all, 1, 1, c.FMT_U16, c.FMT_U8) == want_all).all()
	rng = np.random.default_rng(11)
	for sc, dc in itertools.product(CH, CH):
		for mode in (0, 1):
			if sc == 1 and dc == 4 and mode == 0:
				continue  # pixel-0 quirk: byte-wise kernel, no table form involved
			px = 40961
			src = np.concatenate([np.repeat(np.arange(65536, dtype=np.uint16), sc)[: px * sc // 2], random_samples(rng, px * sc - px * sc // 2, c.FMT_U16)])
			want = oracle.convert(src, sc, dc, c.FMT_U16, c.FMT_U8, mode)
			assert same_bits(convert_device(src, sc, dc, c.FMT_U16, c.FMT_U8, mode), want), (sc, dc, mode)
		if sc == 1 and dc == 4:
			continue
		# 192: no warp of runs stays inside a row (lane-wise stores); 512: whole warps per row (staged stores)
		for w, h, count in ((192, 37, 3), (512, 9, 2)):
			src = random_samples(rng, w * h * count * sc, c.FMT_U16)
			want = oracle.convert(src, sc, dc, c.FMT_U16, c.FMT_U8, 1).reshape(count, h, w * dc)[:, ::-1, :].copy().reshape(-1)
			a, b = DevBuf(src.nbytes), DevBuf(want.nbytes)
			device.upload(a.ptr, src)
			device.convert_flip(a.ptr, b.ptr, w, h, count, sc, dc, c.FMT_U16, c.FMT_U8, 1)
			got = np.empty_like(want)
			device.download(got, b.ptr)
			a.free(), b.free()
			assert same_bits(got, want), (sc, dc, w, "flipped")


def test_large_table_jobs_take_the_bank_private_form(oracle):
	"""From about a megapixel on the dispatcher picks the bank-private form by itself (lanes_preferred,
	convert_lanes.cuh); RGBA16 -> RGB8 at six 1024^2 faces, plain and flipped, against the oracle."""
	from cacaoengine_b200 import _capi as c, device

	w, h, count = 1024, 1024, 6
	rng = np.random.default_rng(5)
	src = rng.integers(0, 65536, w * h * count * 4, dtype=np.uint16)
	want = oracle.convert(src, 4, 3, c.FMT_U16, c.FMT_U8)
	assert same_bits(convert_device(src, 4, 3, c.FMT_U16, c.FMT_U8), want)
	a, b = DevBuf(src.nbytes), DevBuf(want.nbytes)
	device.upload(a.ptr, src)
	device.convert_flip(a.ptr, b.ptr, w, h, count, 4, 3, c.FMT_U16, c.FMT_U8)
	got = np.empty_like(want)
	device.download(got, b.ptr)
	a.free(), b.free()
	assert same_bits(got, want.reshape(count, h, w * 3)[:, ::-1, :].copy().reshape(-1))


@pytest.mark.parametrize("w,h,count", [(2048, 1500, 1), (1000, 777, 3), (96, 4000, 2)])
def test_convert_flip_host_pipeline_in_row_bands(oracle, w, h, count):
	"""Host-memory cacao_img_convert_flip on images large enough for the chunked pipeline: chunks are bands
	of whole rows, each mirrored in itself and stored at the mirrored band.  Pageable (numpy) and
	pinned buffers; RGBA16 -> RGB8 (table shape) and RGB8 -> RGBA32F."""
	from cacaoengine_b200 import _capi as c, device

	rng = np.random.default_rng(w + h)
	for sc, dc, sf, df in ((4, 3, 1, 0), (3, 4, 0, 3)):
		src = random_samples(rng, w * h * count * sc, sf)
		want = oracle.convert(src, sc, dc, sf, df).reshape(count, h, w * dc)[:, ::-1, :].copy().reshape(-1)
		got = np.empty_like(want)
		device.convert_flip(src.ctypes.data, got.ctypes.data, w, h, count, sc, dc, sf, df, mem=c.MEM_HOST)
		assert same_bits(got, want), ("pageable", sc, dc)
		sp, s_host = device.host_alloc(src.nbytes)
		dp, d_host = device.host_alloc(want.nbytes)
		s_host[:] = src.view(np.uint8)
		d_host[:] = 0
		device.convert_flip(sp, dp, w, h, count, sc, dc, sf, df, mem=c.MEM_HOST)
		assert same_bits(d_host.view(want.dtype), want), ("pinned", sc, dc)
		device.host_free(sp), device.host_free(dp)


def test_convert_flip_engine_chain_full_size(oracle):
	"""The engine's three load-time passes -- 16 -> 8 bit, -> RGB (Cubemap.cpp:28-31), Flip (OpenGLCubemap.cpp:25)
	-- as one kernel on six 2048^2 RGBA16 faces: equals ToRGB8 followed by flip_rows (checksum), and a
	sampled face equals the oracle's composition."""
	from cacaoengine_b200 import _capi as c, device

	w = h = 2048
	px = w * h * 6
	a, mid, two, one = DevBuf(px * 8), DevBuf(px * 3), DevBuf(px * 3), DevBuf(px * 3)
	device.synth(a.ptr, px * 4, c.FMT_U16, 4, 0xC0FFEE + 9)
	before = device.launch_count()
	device.convert_flip(a.ptr, one.ptr, w, h, 6, 4, 3, c.FMT_U16, c.FMT_U8)
	assert device.launch_count() - before == 1
	device.convert_batch(a.ptr, mid.ptr, w * h, 6, 4, 3, c.FMT_U16, c.FMT_U8)
	for f in range(6):
		device.flip_rows(mid.ptr + f * w * h * 3, two.ptr + f * w * h * 3, w, h, 3)
	assert device.checksum(one.ptr, px * 3) == device.checksum(two.ptr, px * 3)
	rows = 64
	head = np.empty(rows * w * 4, np.uint16)
	device.download(head, a.ptr + 2 * w * h * 8)  # first rows of face 2 ...
	got = np.empty(rows * w * 3, np.uint8)
	device.download(got, one.ptr + 2 * w * h * 3 + (h - rows) * w * 3)  # ... are its last rows, mirrored
	want = oracle.convert(head, 4, 3, oracle.U16, oracle.U8).reshape(rows, w * 3)[::-1].reshape(-1)
	assert same_bits(got, want)
	for buf in (a, mid, two, one):
		buf.free()


def test_jpeg_12_to_16_bit_widening():
	"""libs/image/src/JPEG.cpp:79-82: `uint16_t(twelveBit[i]) << 4` over int16 samples -- restated in
	numpy (the reference loop needs libjpeg-turbo to build, so this row is parity-unpinned)."""
	from cacaoengine_b200 import _capi as c, device
	from gpu_util import DevBuf

	rng = np.random.default_rng(44)
	for n in (1, 7, 8, 1000, 40963):
		src = rng.integers(0, 4096, n).astype(np.int16)
		src[:1] = -1  # the cast to uint16_t wraps negatives before shifting
		want = (src.astype(np.uint16).astype(np.uint32) << 4).astype(np.uint16)
		out = np.empty(n, np.uint16)
		c.check(c.lib().cacao_img_shift_left_u16(-1, src.ctypes.data, out.ctypes.data, n, 4, c.MEM_HOST, None))
		assert (out == want).all()
		a = DevBuf(n * 2)
		device.upload(a.ptr, src)
		device.shift_left_u16(a.ptr, a.ptr, n, 4)  # in place
		got = np.empty(n, np.uint16)
		device.download(got, a.ptr)
		a.free()
		assert (got == want).all()
	assert c.lib().cacao_img_shift_left_u16(-1, src.ctypes.data, out.ctypes.data, 4, 16, c.MEM_HOST, None) == c.E_BAD_ARG


def test_device_synth_and_checksum_match_host(oracle):
	from cacaoengine_b200 import _capi as c, device
	from gpu_util import DevBuf

	for fmt, ch in ((c.FMT_U8, 3), (c.FMT_U16, 4), (c.FMT_F16, 4), (c.FMT_F32, 4), (c.FMT_F32, 3)):
		n = 100003 if fmt != c.FMT_U8 else 100004
		want = oracle.synth(n, fmt, ch, 0xC0FFEE, first_sample=12345)
		buf = DevBuf(want.nbytes)
		device.synth(buf.ptr, n, fmt, ch, 0xC0FFEE, first_sample=12345)
		got = np.empty_like(want)
		device.download(got, buf.ptr)
		assert same_bits(got, want), fmt
		nb = want.nbytes // 4 * 4
		words = want.view(np.uint8)[:nb].view(np.uint32).astype(np.uint64)
		idx = np.arange(words.size, dtype=np.uint64)
		h = (words + idx * np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)
		h ^= h >> np.uint64(16); h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
		h ^= h >> np.uint64(13); h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
		h ^= h >> np.uint64(16)
		assert device.checksum(buf.ptr, nb) == int(h.sum(dtype=np.uint64))
		buf.free()


def test_full_size_cfg2_batch_by_sampling(oracle):
	"""BASELINE config 2 at full size (256 x 1920x1080 RGBA8 -> RGBA16F), inputs generated on the
	device; sampled frames are read back and compared bit for bit, and a linearity-style property
	holds for the whole batch: converting the batch == converting each half (checksum of checksums)."""
	from cacaoengine_b200 import _capi as c, device
	from gpu_util import DevBuf

	ppi, count = 1920 * 1080, 256
	src, dst = DevBuf(ppi * count * 4), DevBuf(ppi * count * 8)
	device.synth(src.ptr, ppi * count * 4, c.FMT_U8, 4, 0xC0FFEE + 2)
	device.convert_batch(src.ptr, dst.ptr, ppi, count, 4, 4, c.FMT_U8, c.FMT_F16)
	whole = device.checksum(dst.ptr, ppi * count * 8)
	for k in (0, 101, 255):
		got = np.empty(ppi * 4, np.float16)
		device.download(got, dst.ptr + k * ppi * 8)
		want = oracle.convert(oracle.synth(ppi * 4, c.FMT_U8, 4, 0xC0FFEE + 2, first_sample=k * ppi * 4), 4, 4, c.FMT_U8, c.FMT_F16)
		assert same_bits(got, want), k
	# same batch converted as two independent halves into a fresh buffer
	dst2 = DevBuf(ppi * count * 8)
	half = count // 2
	device.convert_batch(src.ptr, dst2.ptr, ppi, half, 4, 4, c.FMT_U8, c.FMT_F16)
	device.convert_batch(src.ptr + half * ppi * 4, dst2.ptr + half * ppi * 8, ppi, count - half, 4, 4, c.FMT_U8, c.FMT_F16)
	assert device.checksum(dst2.ptr, ppi * count * 8) == whole
	for b in (src, dst, dst2):
		b.free()


def test_depth_table_kernel_beyond_4_gib(oracle, monkeypatch):
	"""The 16 -> 8 bit kernel of convert_lanes.cuh on a stream whose byte offsets exceed 2^32 (640 Mpixel RGBA16 =
	5.1 GB -> RGB8, flipped as 40 images of 4096^2): equal to the two-level form of the same conversion as a whole
	(checksum), equal to the oracle on sampled rows of the first, a middle and the last image, and equal to the
	conversion of each half on its own."""
	from cacaoengine_b200 import _capi as c, device
	from gpu_util import DevBuf

	w = h = 4096
	count = 40
	ppi = w * h
	src, dst, dst2 = DevBuf(ppi * count * 8), DevBuf(ppi * count * 3), DevBuf(ppi * count * 3)
	device.synth(src.ptr, ppi * count * 4, c.FMT_U16, 4, 0xC0FFEE + 11)
	device.convert_flip(src.ptr, dst.ptr, w, h, count, 4, 3, c.FMT_U16, c.FMT_U8)
	whole = device.checksum(dst.ptr, ppi * count * 3)
	rows = 8
	for k in (0, 19, count - 1):
		for y in (0, 1777, h - rows):  # source rows [y, y + rows) land, mirrored, at rows [h - y - rows, h - y)
			got = np.empty(rows * w * 3, np.uint8)
			device.download(got, dst.ptr + (k * ppi + (h - y - rows) * w) * 3)
			band = oracle.synth(rows * w * 4, c.FMT_U16, 4, 0xC0FFEE + 11, first_sample=(k * ppi + y * w) * 4)
			want = oracle.convert(band, 4, 3, c.FMT_U16, c.FMT_U8).reshape(rows, w * 3)[::-1].reshape(-1)
			assert same_bits(got, want), (k, y)
	monkeypatch.setenv("CACAO_LUT_MODE", "two")
	device.convert_flip(src.ptr, dst2.ptr, w, h, count, 4, 3, c.FMT_U16, c.FMT_U8)
	assert device.checksum(dst2.ptr, ppi * count * 3) == whole
	monkeypatch.delenv("CACAO_LUT_MODE")
	half = count // 2
	device.convert_flip(src.ptr, dst2.ptr, w, h, half, 4, 3, c.FMT_U16, c.FMT_U8)
	device.convert_flip(src.ptr + half * ppi * 8, dst2.ptr + half * ppi * 3, w, h, count - half, 4, 3, c.FMT_U16, c.FMT_U8)
	assert device.checksum(dst2.ptr, ppi * count * 3) == whole
	for b in (src, dst, dst2):
		b.free()


def test_round_trip_properties():
	"""Size-independent properties at a large size: RGB->RGBA->RGB is the identity, U8->U16->(>>8) too,
	and U8->F32->U8 returns every byte."""
	from cacaoengine_b200 import _capi as c, device
	from gpu_util import DevBuf

	px = 4096 * 4096
	a, b, d = DevBuf(px * 3), DevBuf(px * 16), DevBuf(px * 3)
	device.synth(a.ptr, px * 3, c.FMT_U8, 3, 77)
	ref = device.checksum(a.ptr, px * 3)
	device.convert(a.ptr, b.ptr, px, 3, 4, c.FMT_U8, c.FMT_U8)
	device.convert(b.ptr, d.ptr, px, 4, 3, c.FMT_U8, c.FMT_U8)
	assert device.checksum(d.ptr, px * 3) == ref
	device.convert(a.ptr, b.ptr, px, 3, 4, c.FMT_U8, c.FMT_F32)
	device.convert(b.ptr, d.ptr, px, 4, 3, c.FMT_F32, c.FMT_U8)
	assert device.checksum(d.ptr, px * 3) == ref
	device.convert(a.ptr, b.ptr, px, 3, 3, c.FMT_U8, c.FMT_F16)
	device.convert(b.ptr, d.ptr, px, 3, 3, c.FMT_F16, c.FMT_U8)
	assert device.checksum(d.ptr, px * 3) == ref
	for x in (a, b, d):
		x.free()


def test_buffers_beyond_4_gib():
	"""Offsets past 2^32 bytes: a 1.2 Gpixel RGBA8 -> RGBA32F conversion (4.8 GB in, 19 GB out) equals the
	same conversion done as three independent slices (checksum of the whole = position-aware checksums
	of the parts are not additive, so compare slice by slice)."""
	from cacaoengine_b200 import _capi as c, device
	from gpu_util import DevBuf

	px = 1200 * 1024 * 1024
	src, dst, ref = DevBuf(px * 4), DevBuf(px * 16), DevBuf((px // 3) * 16)
	device.synth(src.ptr, px * 4, c.FMT_U8, 4, 99)
	device.convert(src.ptr, dst.ptr, px, 4, 4, c.FMT_U8, c.FMT_F32)
	third = px // 3
	for k in range(3):
		device.convert(src.ptr + k * third * 4, ref.ptr, third, 4, 4, c.FMT_U8, c.FMT_F32)
		assert device.checksum(ref.ptr, third * 16) == device.checksum(dst.ptr + k * third * 16, third * 16), k
	# the very last pixels, read back and checked by hand: v / 255
	tail_in, tail_out = np.empty(64, np.uint8), np.empty(64, np.float32)
	device.download(tail_in, src.ptr + px * 4 - 64)
	device.download(tail_out, dst.ptr + px * 16 - 256)
	assert same_bits(tail_out, tail_in.astype(np.float32) / np.float32(255))
	for b in (src, dst, ref):
		b.free()


def test_randomised_sizes_and_batches(oracle):
	"""Seeded sweep: random shape, pixel count (tile edges, tails), batch split, mode and memory kind."""
	rng = np.random.default_rng(20261)
	for case in range(60):
		sf, df, sc, dc = SHAPES[int(rng.integers(0, len(SHAPES)))]
		count = int(rng.choice([1, 1, 2, 5]))
		ppi = int(rng.choice([1, 31, 32, 33, 511, 512, 513, 4096, 8191, 70001]))
		mode = int(rng.integers(0, 2))
		src = random_samples(rng, count * ppi * sc, sf)
		want = np.concatenate([oracle.convert(src[k * ppi * sc:(k + 1) * ppi * sc], sc, dc, sf, df, mode) for k in range(count)])
		got = convert_device(src, sc, dc, sf, df, mode, count=count) if case % 3 else convert_host(src, sc, dc, sf, df, mode, count=count)
		assert same_bits(got, want), (case, sf, df, sc, dc, count, ppi, mode)


def test_concurrent_callers(oracle):
	"""The reference functions are re-entrant and called from arbitrary engine threads (thread-pool tasks,
	the GL thread: OpenGLCubemap.cpp:13-25).  Eight threads hammer the host-memory and device-memory
	entry points of one device at once (ctypes drops the GIL); every result must still be exact."""
	from concurrent.futures import ThreadPoolExecutor

	from cacaoengine_b200 import _capi as c

	rng = np.random.default_rng(5)
	jobs = []
	for k in range(24):
		sf, df, sc, dc = SHAPES[int(rng.integers(0, len(SHAPES)))]
		px = int(rng.choice([100, 5000, 300000, 1500000]))
		src = random_samples(rng, px * sc, sf)
		jobs.append((k, src, sc, dc, sf, df, oracle.convert(src, sc, dc, sf, df, 1)))

	def run(job):
		k, src, sc, dc, sf, df, want = job
		got = convert_host(src, sc, dc, sf, df, 1) if k % 2 else convert_device(src, sc, dc, sf, df, 1)
		return same_bits(got, want)

	with ThreadPoolExecutor(max_workers=8) as ex:
		for rep in range(3):
			assert all(ex.map(run, jobs))


TEXEL_HINTS = ["rgba8888", "argb8888", "bgra8888", "abgr8888", "rgba8880", "bgra8880", "rgba5650", "bgra5650", "rgba4444", "argb1555", "rgba5551", "rgba8000", "rgba0800", "bgra0050", "rgba2222", "rgba1111", "rgba3320"]


@pytest.mark.parametrize("hint", TEXEL_HINTS)
def test_unpack_texels_matches_oracle(oracle, hint):
	"""cacao_img_unpack_texels (engine/src/core/Model.cpp:245-316: swizzle to r,g,b,a order, drop zero-count
	channels, widen sub-8-bit counts) against the oracle: vector body + scalar tail, device and host memory,
	aligned and unaligned buffers."""
	from cacaoengine_b200 import _capi as c, device

	for texels in (1, 15, 16, 17, 4099, 1 << 16):
		src = oracle.synth(texels * 4, c.FMT_U8, 4, 900 + texels).reshape(-1, 4)
		want = oracle.unpack_texels(src, hint)
		ch = want.shape[1]
		got = np.zeros(texels * ch, np.uint8)
		assert device.unpack_texels(src.ctypes.data, got.ctypes.data, texels, hint, mem=c.MEM_HOST) == ch
		assert np.array_equal(got.reshape(-1, ch), want), (hint, texels, "host")
		for shift in (0, 4):  # 16-byte aligned, and 4-byte aligned only (scalar kernel)
			a, b = DevBuf(texels * 4 + 16), DevBuf(texels * ch + 16)
			device.upload(a.ptr + shift, src)
			assert device.unpack_texels(a.ptr + shift, b.ptr + shift, texels, hint) == ch
			device.download(got, b.ptr + shift)
			a.free(), b.free()
			assert np.array_equal(got.reshape(-1, ch), want), (hint, texels, shift)


def test_unpack_texels_rejects_what_the_reference_rejects():
	from cacaoengine_b200 import _capi as c, device

	buf = np.zeros(64, np.uint8)
	for hint, text in (("rgba8800", "Invalid zeroed-channel layout for model embedded texture!"), ("rgba0888", "Only the alpha channel may be zero for a model embedded texture layout with only one zeroed channel!"),
					   ("rgba0000", "Impossible channel layout for model embedded texture!"), ("rgbx8888", "not four of r,g,b,a"), ("rgba888", "eight characters")):
		with pytest.raises(c.CacaoError) as e:
			device.unpack_texels(buf.ctypes.data, buf.ctypes.data, 4, hint, mem=c.MEM_HOST)
		assert e.value.status == c.E_BAD_ARG and text in str(e.value)


def test_unpack_texels_full_size_property(oracle):
	"""64 Mtexels on the device: the all-8-bit hints are pure byte permutations, so unpacking with
	"bgra8888" and then again with "bgra8888" (its own inverse) is the identity, and the RGB result of
	"rgba8880" equals the RGBA -> RGB layout change of the conversion kernel."""
	from cacaoengine_b200 import _capi as c, device

	n = 1 << 26
	a, b, d = DevBuf(n * 4), DevBuf(n * 4), DevBuf(n * 4)
	device.synth(a.ptr, n * 4, c.FMT_U8, 4, 4242)
	device.unpack_texels(a.ptr, b.ptr, n, "bgra8888")
	device.unpack_texels(b.ptr, d.ptr, n, "bgra8888")
	assert device.checksum(d.ptr, n * 4) == device.checksum(a.ptr, n * 4) != device.checksum(b.ptr, n * 4)
	device.unpack_texels(a.ptr, b.ptr, n, "rgba8880")
	device.convert(a.ptr, d.ptr, n, 4, 3, c.FMT_U8, c.FMT_U8)
	assert device.checksum(b.ptr, n * 3) == device.checksum(d.ptr, n * 3)
	for x in (a, b, d):
		x.free()


@pytest.mark.parametrize("fmt", FMTS, ids=[NAME[f] for f in FMTS])
@pytest.mark.parametrize("ch", (1, 3, 4))
def test_downsample2x_matches_oracle(oracle, fmt, ch):
	"""One mip level (Spec B.3) against the oracle: sizes that take the 256-bit vector kernel, odd sizes
	that take the per-pixel kernel, a vector body with a scalar remainder, one-wide / one-high images;
	device and host memory."""
	from cacaoengine_b200 import _capi as c, device

	for w, h in ((256, 64), (254, 63), (1, 1), (1, 7), (7, 1), (2, 2), (96, 5), (160, 40), (224, 3), (34, 34)):
		src = oracle.synth(w * h * ch, fmt, ch, 70 + w)
		want = oracle.downsample2x(src, w, h, ch, fmt)
		got = np.zeros_like(want)
		device.downsample2x(src.ctypes.data, got.ctypes.data, w, h, ch, fmt, mem=c.MEM_HOST)
		assert same_bits(got, want), (w, h, "host")
		a, b = DevBuf(src.nbytes), DevBuf(want.nbytes)
		device.upload(a.ptr, src)
		device.downsample2x(a.ptr, b.ptr, w, h, ch, fmt)
		got[:] = 0
		device.download(got, b.ptr)
		a.free(), b.free()
		assert same_bits(got, want), (w, h, "device")


def test_mip_chain_full_size(oracle):
	"""A 4096 x 4096 RGBA8 chain on the device: every level's checksum equals the oracle's for the
	levels small enough to recompute (<= 512^2, from the downloaded parent), and a constant image
	stays constant all the way down to 1 x 1."""
	from cacaoengine_b200 import _capi as c, device

	n, ch, fmt = 4096, 4, c.FMT_U8
	cur = DevBuf(n * n * ch)
	device.synth(cur.ptr, n * n * ch, fmt, ch, 808)
	w = n
	while w > 1:
		nxt = DevBuf((w // 2) * (w // 2) * ch)
		device.downsample2x(cur.ptr, nxt.ptr, w, w, ch, fmt)
		if w <= 1024:
			parent = np.empty(w * w * ch, np.uint8)
			device.download(parent, cur.ptr)
			child = np.empty((w // 2) * (w // 2) * ch, np.uint8)
			device.download(child, nxt.ptr)
			assert same_bits(child, oracle.downsample2x(parent, w, w, ch, fmt)), w
		cur.free()
		cur, w = nxt, w // 2
	cur.free()
	flat = DevBuf(1024 * 1024 * 4)
	device.upload(flat.ptr, np.full(1024 * 1024 * 4, 77, np.uint8))
	w = 1024
	while w > 1:
		nxt = DevBuf((w // 2) * (w // 2) * 4)
		device.downsample2x(flat.ptr, nxt.ptr, w, w, 4, fmt)
		flat.free()
		flat, w = nxt, w // 2
	last = np.empty(4, np.uint8)
	device.download(last, flat.ptr)
	flat.free()
	assert (last == 77).all()
