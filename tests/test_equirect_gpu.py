"""GPU parity for the equirect -> cubemap resampler against the CPU restatement of DESIGN.md Spec
B.2.  The reference has no such code (tools/cubetool/main.cpp:31-33), so this is parity with OUR
oracle -- "parity unpinned vs reference".  The tolerance north_star allows is 1 LSB (u8/u16) or
1e-5 relative (f16/f32); the design goal is stricter -- bit-exact -- and that is what is asserted,
with the tolerance check kept as the documented floor."""
import numpy as np
import pytest

from gpu_util import FMTS, NAME, NP, DevBuf, same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _lib(built_lib):
	from cacaoengine_b200 import device

	device.init()
	yield


def gpu_equirect_host(src, W, H, ch, fmt, N, mask=0x3F, r0=0, r1=None):
	from cacaoengine_b200 import _capi as c

	out = np.zeros(6 * N * N * ch, NP[fmt])
	c.check(c.lib().cacao_img_equirect_to_cube(-1, src.ctypes.data, W, H, 0, H, ch, fmt, out.ctypes.data, N, mask, r0, N if r1 is None else r1, c.MEM_HOST, None))
	return out


def gpu_equirect_device(src, W, H, ch, fmt, N, mask=0x3F, r0=0, r1=None, window=None):
	from cacaoengine_b200 import device

	out = np.zeros(6 * N * N * ch, NP[fmt])
	row_bytes = W * ch * src.itemsize
	w0, wn = window if window else (0, H)
	a, b = DevBuf(wn * row_bytes), DevBuf(out.nbytes)
	device.upload(a.ptr, src.reshape(H, -1)[w0:w0 + wn].copy())
	device.upload(b.ptr, out)
	device.equirect_to_cube(a.ptr, W, H, ch, fmt, b.ptr, N, mask, r0, r1, src_row0=w0, src_rows=wn)
	device.download(out, b.ptr)
	a.free(), b.free()
	return out


def within_tolerance(got, want, fmt):
	from cacaoengine_b200 import _capi as c

	if fmt in (c.FMT_U8, c.FMT_U16):
		return np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= 1
	g, w = got.astype(np.float64), want.astype(np.float64)
	return (np.abs(g - w) <= 1e-5 * np.abs(w) + 1e-30).all()


@pytest.mark.parametrize("fmt", FMTS, ids=[NAME[f] for f in FMTS])
@pytest.mark.parametrize("ch", (1, 3, 4))
def test_every_format_and_layout_matches_oracle(oracle, fmt, ch):
	W, H, N = 200, 100, 37  # odd face size: centre pixel has sc == 0 exactly
	src = oracle.synth(W * H * ch, fmt, ch, 0xC0FFEE + 1)
	want = oracle.equirect_to_cube(src, W, H, ch, fmt, N)
	for got in (gpu_equirect_device(src, W, H, ch, fmt, N), gpu_equirect_host(src, W, H, ch, fmt, N)):
		assert within_tolerance(got, want, fmt)
		assert same_bits(got, want)


@pytest.mark.parametrize("fmt", FMTS, ids=[NAME[f] for f in FMTS])
@pytest.mark.parametrize("ch", (1, 3, 4))
def test_row_mirror_sharing_is_bit_identical(oracle, monkeypatch, fmt, ch):
	"""Whole-face launches let one thread produce a pixel, its column mirror, its row mirror and the
	diagonal one from a single projection; row-band launches (and CACAO_EQUIRECT_NO_QUAD=1) only share
	the column mirror.  Both must equal the per-pixel oracle, for even, odd and degenerate face sizes."""
	for W, H, N in ((96, 48, 36), (130, 70, 37), (16, 8, 1), (16, 8, 2), (64, 32, 130)):
		src = oracle.synth(W * H * ch, fmt, ch, 77 + N)
		want = oracle.equirect_to_cube(src, W, H, ch, fmt, N)
		monkeypatch.delenv("CACAO_EQUIRECT_NO_QUAD", raising=False)
		assert same_bits(gpu_equirect_device(src, W, H, ch, fmt, N), want), ("quad", W, H, N)
		monkeypatch.setenv("CACAO_EQUIRECT_NO_QUAD", "1")
		assert same_bits(gpu_equirect_device(src, W, H, ch, fmt, N), want), ("pair", W, H, N)
		monkeypatch.delenv("CACAO_EQUIRECT_NO_QUAD", raising=False)
		# a band launch next to a whole-face launch of the other faces
		got = gpu_equirect_device(src, W, H, ch, fmt, N, mask=0x15, r0=0, r1=max(1, N // 2))
		ref = oracle.equirect_to_cube(src, W, H, ch, fmt, N, face_mask=0x15, row_begin=0, row_end=max(1, N // 2))
		assert same_bits(got, ref), ("band", W, H, N)


@pytest.mark.parametrize("fmt", FMTS, ids=[NAME[f] for f in FMTS])
def test_degenerate_panoramas(oracle, fmt):
	"""One-texel, one-row, one-column and few-texel panoramas (every tap wraps or clamps; the joint
	two-tap loads of three-channel texels must fall back near the end of the buffer), device and host."""
	for ch in (1, 3, 4):
		for W, H, N in ((1, 1, 3), (2, 1, 4), (1, 2, 5), (3, 2, 7), (5, 3, 16), (9, 1, 8), (8, 2, 33)):
			src = oracle.synth(W * H * ch, fmt, ch, 11 * W + H)
			want = oracle.equirect_to_cube(src, W, H, ch, fmt, N)
			assert same_bits(gpu_equirect_device(src, W, H, ch, fmt, N), want), (ch, W, H, N, "device")
			assert same_bits(gpu_equirect_host(src, W, H, ch, fmt, N), want), (ch, W, H, N, "host")


def test_cfg1_shape_rgba8(oracle):
	"""BASELINE config 1: 2048x1024 RGBA8 -> 6x512^2."""
	from cacaoengine_b200 import _capi as c

	W, H, N = 2048, 1024, 512
	src = oracle.synth(W * H * 4, c.FMT_U8, 4, 0xC0FFEE)
	want = oracle.equirect_to_cube(src, W, H, 4, c.FMT_U8, N)
	got = gpu_equirect_device(src, W, H, 4, c.FMT_U8, N)
	assert same_bits(got, want)


def test_smooth_panorama_f32_against_double_truth(oracle):
	"""End-to-end accuracy of the GPU result against the float64 libm resampler (not just against the
	fp32 restatement): within 1e-5 relative on a smooth HDR-like panorama."""
	from cacaoengine_b200 import _capi as c

	W, H, N = 1024, 512, 256
	y, x = np.mgrid[0:H, 0:W]
	phi, th = (x + 0.5) / W * 2 * np.pi, (y + 0.5) / H * np.pi
	img = np.stack([2.0 + np.sin(phi) * np.sin(th), 2.0 + np.cos(2 * phi) * np.sin(th) ** 2, 2.0 + np.cos(th), np.ones_like(phi)], -1).astype(np.float32).reshape(-1)
	got = gpu_equirect_device(img, W, H, 4, c.FMT_F32, N).astype(np.float64)
	truth = oracle.equirect_to_cube_f64(img, W, H, 4, c.FMT_F32, N)
	assert (np.abs(got - truth) / np.abs(truth)).max() < 1e-5


def test_face_masks_row_bands_and_source_windows(oracle):
	"""Work units (face, row band) fed only their planned source-row window reproduce the full result:
	the sharding scheme of DESIGN.md "Multi-GPU"."""
	from cacaoengine_b200 import _capi as c, device

	W, H, N, ch, fmt = 512, 256, 96, 4, c.FMT_F32
	src = oracle.synth(W * H * ch, fmt, ch, 5)
	full = oracle.equirect_to_cube(src, W, H, ch, fmt, N)
	acc = np.zeros_like(full)
	for f in range(6):
		for (r0, r1) in ((0, 24), (24, 48), (48, 96)):
			win = device.equirect_src_window(W, H, N, 1 << f, r0, r1)
			part = gpu_equirect_device(src, W, H, ch, fmt, N, 1 << f, r0, r1, window=win)
			sl = slice((f * N + r0) * N * ch, (f * N + r1) * N * ch)
			assert same_bits(part[sl], full[sl]), (f, r0, r1, win)
			assert (np.delete(part, np.arange(sl.start, sl.stop)) == 0).all()  # nothing else written
			acc[sl] = part[sl]
			if f in (0, 1, 4, 5):
				assert win[1] < H  # side-face bands never need the whole panorama
	assert same_bits(acc, full)


def test_unit_list_launch_matches_per_unit_calls(oracle):
	"""cacao_img_equirect_to_cube_units: a rank's whole unit list in one call (one launch when the units
	are equally tall) equals the oracle on exactly those rows and writes nothing else."""
	from cacaoengine_b200 import _capi as c, device, sharding

	W, H, N, ch, fmt = 300, 150, 70, 3, c.FMT_U16
	src = oracle.synth(W * H * ch, fmt, ch, 6)
	units = sharding.cube_units(N, 5)[::3]  # 10 units, mixed faces and bands
	assert len(units) == 10
	want = np.zeros(6 * N * N * ch, np.uint16)
	for u in units:
		oracle.equirect_to_cube(src, W, H, ch, fmt, N, face_mask=1 << u.face, row_begin=u.row_begin, row_end=u.row_end, out=want)
	a, b = DevBuf(src.nbytes), DevBuf(want.nbytes)
	device.upload(a.ptr, src)
	device.upload(b.ptr, np.zeros_like(want))
	before = device.launch_count()
	device.equirect_to_cube_units(a.ptr, W, H, ch, fmt, b.ptr, N, units)
	assert device.launch_count() - before == 1
	got = np.empty_like(want)
	device.download(got, b.ptr)
	assert same_bits(got, want)
	# a launch's grid is as tall as its tallest unit, so very different heights go to separate launches
	# (the whole face folds onto its upper 35 rows: 35, 35, 34 | 10 rows -> two); heights within a factor
	# of eight are cut to a common height instead
	mixed = [(0, 0, 70), (3, 10, 20), (1, 0, 35), (2, 36, 70)]
	want2 = np.zeros_like(want)
	for f, r0, r1 in mixed:
		oracle.equirect_to_cube(src, W, H, ch, fmt, N, face_mask=1 << f, row_begin=r0, row_end=r1, out=want2)
	device.upload(b.ptr, np.zeros_like(want))
	before = device.launch_count()
	device.equirect_to_cube_units(a.ptr, W, H, ch, fmt, b.ptr, N, mixed)
	assert device.launch_count() - before == 2
	device.download(got, b.ptr)
	assert same_bits(got, want2)
	similar = [(0, 0, 70), (1, 0, 35), (2, 36, 70), (4, 3, 40)]  # 70 -> 2 x 35: one launch of five pieces
	want3 = np.zeros_like(want)
	for f, r0, r1 in similar:
		oracle.equirect_to_cube(src, W, H, ch, fmt, N, face_mask=1 << f, row_begin=r0, row_end=r1, out=want3)
	device.upload(b.ptr, np.zeros_like(want))
	before = device.launch_count()
	device.equirect_to_cube_units(a.ptr, W, H, ch, fmt, b.ptr, N, similar)
	assert device.launch_count() - before == 1
	device.download(got, b.ptr)
	a.free(), b.free()
	assert same_bits(got, want3)
	with pytest.raises(c.CacaoError):
		device.equirect_to_cube_units(1, W, H, ch, fmt, 1, N, [(7, 0, 1)])


@pytest.mark.parametrize("fmt,ch,N", [(3, 4, 70), (0, 4, 71), (2, 3, 45), (1, 1, 96), (0, 3, 33)], ids=["f32x4-70", "u8x4-71", "f16x3-45", "u16x1-96", "u8x3-33"])
def test_mirror_folded_unit_lists_match_the_oracle(oracle, fmt, ch, N):
	"""Lists in which bands meet their mirror bands (folded by the launcher into four-fold units), meet
	them partly, or not at all -- odd and even faces, rows named twice, one-row bands at the centre:
	exactly the listed rows are written and they equal the oracle's."""
	from cacaoengine_b200 import device

	W, H = 4 * N + 3, 2 * N + 1
	src = oracle.synth(W * H * ch, fmt, ch, 40 + N)
	full = oracle.equirect_to_cube(src, W, H, ch, fmt, N)
	a, b = DevBuf(src.nbytes), DevBuf(full.nbytes)
	device.upload(a.ptr, src)
	c0, q = (N - 1) // 2, N // 4
	lists = [
		[(f, 0, N) for f in range(6)],
		[(0, 0, q), (0, N - q, N), (2, q, N - q), (3, 0, c0), (3, c0, c0 + 1), (3, c0 + 1, N)],
		[(1, 0, N // 2 + 5), (1, N // 2 - 7, N), (4, 3, N - 20), (5, c0, c0 + 1), (2, N - 18, N), (2, 0, 17)],
		[(4, 0, q), (4, N - q, N - 1), (5, 1, q + 20), (5, N - q - 20, N), (0, N // 2, N)],
	]
	rng = np.random.default_rng(N)
	for _ in range(4):
		units = []
		for _ in range(int(rng.integers(2, 9))):
			r0, r1 = sorted(int(v) for v in rng.integers(0, N + 1, 2))
			f = int(rng.integers(0, 6))
			units.append((f, r0, r1))
			if rng.random() < 0.6:
				units.append((f, N - r1, N - r0))
		lists.append(units)
	for units in lists:
		want = np.zeros_like(full)
		for f, r0, r1 in units:
			sl = slice((f * N + r0) * N * ch, (f * N + r1) * N * ch)
			want[sl] = full[sl]
		device.upload(b.ptr, np.zeros_like(full))
		device.equirect_to_cube_units(a.ptr, W, H, ch, fmt, b.ptr, N, units)
		got = np.empty_like(full)
		device.download(got, b.ptr)
		assert same_bits(got, want), (units, device.equirect_fold_units(N, units))
	a.free(), b.free()


FUSED = [(sf, ch, dc, df) for sf in (2, 3) for (ch, dc) in ((1, 1), (3, 3), (3, 4), (4, 4), (4, 3)) for df in (0, 1, 2, 3) if not (dc == ch and df == sf)]


def _float_pano(rng, W, H, ch, fmt, dst_fmt):
	"""Finite float samples: around [0, 1] with excursions on both sides for the quantising
	destinations, a wide dynamic range (f16 overflow included for F32 -> F16) for the float ones."""
	if dst_fmt in (0, 1):
		v = rng.uniform(-0.25, 1.25, W * H * ch)
	else:
		v = np.exp2(rng.uniform(-12, 17 if fmt == 3 else 14, W * H * ch)) * rng.choice([-1.0, 1.0], W * H * ch)
	return v.astype(NP[fmt])


@pytest.mark.parametrize("sf,ch,dc,df", FUSED, ids=[f"{NAME[sf]}x{ch}-{NAME[df]}x{dc}" for sf, ch, dc, df in FUSED])
def test_fused_resample_and_convert_matches_the_composition(oracle, sf, ch, dc, df):
	"""Spec B.4: cacao_img_equirect_to_cube_units_cvt == widen (exact) -> Spec B.2 in fp32 -> Spec B.1 to the
	destination, bit for bit, for every float source and every destination format / RGB -> RGBA."""
	from cacaoengine_b200 import device

	W, H, N = 93, 47, 29
	rng = np.random.default_rng(100 * sf + 10 * ch + df)
	src = _float_pano(rng, W, H, ch, sf, df)
	want = oracle.equirect_to_cube_cvt(src, W, H, ch, sf, N, dc, df)
	a, b = DevBuf(src.nbytes), DevBuf(want.nbytes)
	device.upload(a.ptr, src)
	device.upload(b.ptr, np.zeros_like(want))
	device.equirect_to_cube_cvt(a.ptr, W, H, ch, sf, b.ptr, N, dc, df)
	got = np.empty_like(want)
	device.download(got, b.ptr)
	assert same_bits(got, want)
	# a unit list (mirror pair + a lone band): only those rows are written, with the same values
	units = [(2, 3, 9), (2, N - 9, N - 3), (5, 11, 20)]
	device.upload(b.ptr, np.zeros_like(want))
	device.equirect_to_cube_cvt(a.ptr, W, H, ch, sf, b.ptr, N, dc, df, units=units)
	device.download(got, b.ptr)
	part = np.zeros_like(want)
	for f, r0, r1 in units:
		sl = slice((f * N + r0) * N * dc, (f * N + r1) * N * dc)
		part[sl] = want[sl]
	assert same_bits(got, part)
	a.free(), b.free()


@pytest.mark.parametrize("fmt,ch,dfmt,dch", [(0, 4, 0, 4), (1, 3, 1, 3), (2, 1, 2, 1), (3, 4, 3, 4), (3, 3, 2, 4), (2, 4, 0, 4)], ids=["u8x4", "u16x3", "f16x1", "f32x4", "f32x3-f16x4", "f16x4-u8x4"])
@pytest.mark.parametrize("N", [24, 31])
def test_flipped_faces_bottom_row_first(oracle, fmt, ch, dfmt, dch, N):
	"""CACAO_EQUIRECT_FLIP_ROWS: every face equals the unflipped face mirrored vertically -- whole cube
	(four-fold units), a unit list (row j lands at N-1-j, nothing else is written) and the host-memory
	pipeline with six separate face buffers, for plain and fused-conversion shapes."""
	from cacaoengine_b200 import device

	W, H = 4 * N + 1, 2 * N + 3
	rng = np.random.default_rng(N + fmt)
	if fmt >= 2:
		src = _float_pano(rng, W, H, ch, fmt, dfmt)
	else:
		src = oracle.synth(W * H * ch, fmt, ch, 77 + N)
	plain = (oracle.equirect_to_cube(src, W, H, ch, fmt, N) if (dfmt, dch) == (fmt, ch) else oracle.equirect_to_cube_cvt(src, W, H, ch, fmt, N, dch, dfmt)).reshape(6, N, N * dch)
	want = plain[:, ::-1, :].copy()
	a, b = DevBuf(src.nbytes), DevBuf(want.nbytes)
	device.upload(a.ptr, src)
	device.upload(b.ptr, np.zeros_like(want))
	device.equirect_to_cube_cvt(a.ptr, W, H, ch, fmt, b.ptr, N, dch, dfmt, flags=device.FLIP_ROWS)
	got = np.empty_like(want)
	device.download(got, b.ptr)
	assert same_bits(got, want)
	units = [(1, 2, 9), (1, N - 9, N - 2), (3, 0, 5), (4, N // 2 - 1, N // 2 + 2)]
	part = np.zeros_like(want)
	for f, r0, r1 in units:
		part[f, N - r1:N - r0] = want[f, N - r1:N - r0]
	device.upload(b.ptr, np.zeros_like(want))
	device.equirect_to_cube_cvt(a.ptr, W, H, ch, fmt, b.ptr, N, dch, dfmt, units=units, flags=device.FLIP_ROWS)
	device.download(got, b.ptr)
	assert same_bits(got, part)
	a.free(), b.free()
	faces = [np.zeros(N * N * dch, want.dtype) for _ in range(6)]
	device.equirect_to_cube_cvt_host(src.ctypes.data, W, H, ch, fmt, [f.ctypes.data for f in faces], N, dch, dfmt, units=units + [(0, 0, N)], flags=device.FLIP_ROWS)
	part[0] = want[0]
	for f in range(6):
		assert same_bits(faces[f], part[f].reshape(-1)), f


def test_fused_resample_host_path_and_full_size_property(oracle):
	"""Host-memory form with six separate face buffers (RGB32F panorama -> RGBA16F faces), and at
	BASELINE config 3's size: the fused RGBA32F -> RGBA16F cube equals resample-then-convert on the
	device (checksum), which the parity tests pin to the oracle."""
	from cacaoengine_b200 import _capi as c, device

	W, H, N, ch = 160, 80, 40, 3
	rng = np.random.default_rng(5)
	src = _float_pano(rng, W, H, ch, 3, 2)
	want = oracle.equirect_to_cube_cvt(src, W, H, ch, 3, N, 4, 2).reshape(6, -1)
	faces = [np.zeros(N * N * 4, np.float16) for _ in range(6)]
	device.equirect_to_cube_cvt_host(src.ctypes.data, W, H, ch, 3, [f.ctypes.data for f in faces], N, 4, 2)
	for f in range(6):
		assert same_bits(faces[f], want[f]), f
	W, H, N = 8192, 4096, 2048
	a, cube32, cube16, fused = DevBuf(W * H * 16), DevBuf(6 * N * N * 16), DevBuf(6 * N * N * 8), DevBuf(6 * N * N * 8)
	device.synth(a.ptr, W * H * 4, c.FMT_F32, 4, 0xC0FFEE + 3)
	device.equirect_to_cube(a.ptr, W, H, 4, c.FMT_F32, cube32.ptr, N)
	device.convert(cube32.ptr, cube16.ptr, 6 * N * N, 4, 4, c.FMT_F32, c.FMT_F16)
	device.equirect_to_cube_cvt(a.ptr, W, H, 4, c.FMT_F32, fused.ptr, N, 4, c.FMT_F16)
	assert device.checksum(fused.ptr, 6 * N * N * 8) == device.checksum(cube16.ptr, 6 * N * N * 8)
	for buf in (a, cube32, cube16, fused):
		buf.free()


@pytest.mark.parametrize("fmt", FMTS, ids=[NAME[f] for f in FMTS])
@pytest.mark.parametrize("ch", (1, 3, 4))
@pytest.mark.parametrize("rows", ("0", "1"), ids=["gather", "row-strip"])
def test_both_resampling_kernels_on_every_shape(oracle, monkeypatch, fmt, ch, rows):
	"""The gather kernel (equirect.cu) and the row-strip kernel (equirect_rows.cu: FP32-pipe floor and
	truncation, row pointers, the column part of the projection carried across row steps) are two
	implementations of Spec B.2; the launcher picks one per shape.  CACAO_EQUIRECT_ROWS forces either on
	every shape: whole cubes (even, odd, one-pixel and upsampling faces, panoramas one texel wide), unit
	lists with and without mirror partners, flipped output, row bands with source windows -- bit-exact
	against the oracle, with 1, 2 and 5 row steps per thread where the shape has a row loop."""
	from cacaoengine_b200 import device

	monkeypatch.setenv("CACAO_EQUIRECT_ROWS", rows)
	for steps in (None, "2", "5") if rows == "1" else (None,):
		if steps:
			monkeypatch.setenv("CACAO_EQUIRECT_STEPS", steps)
		for W, H, N in ((96, 48, 36), (130, 70, 37), (16, 8, 1), (1, 1, 5), (3, 2, 7), (64, 32, 130), (257, 129, 64)):
			src = oracle.synth(W * H * ch, fmt, ch, 5 + N)
			want = oracle.equirect_to_cube(src, W, H, ch, fmt, N)
			assert same_bits(gpu_equirect_device(src, W, H, ch, fmt, N), want), (W, H, N, steps)
	monkeypatch.delenv("CACAO_EQUIRECT_STEPS", raising=False)
	W, H, N = 301, 150, 75
	src = oracle.synth(W * H * ch, fmt, ch, 9)
	full = oracle.equirect_to_cube(src, W, H, ch, fmt, N)
	a, b = DevBuf(src.nbytes), DevBuf(full.nbytes)
	device.upload(a.ptr, src)
	units = [(0, 0, 20), (0, 55, 75), (2, 10, 65), (3, 0, 37), (3, 37, 38), (4, 3, 50), (5, 70, 75), (1, 30, 31)]
	for flags in (0, device.FLIP_ROWS):
		want = np.zeros_like(full).reshape(6, N, N * ch)
		ref = full.reshape(6, N, N * ch)
		for f, r0, r1 in units:
			if flags:
				want[f, N - r1:N - r0] = ref[f, r0:r1][::-1]
			else:
				want[f, r0:r1] = ref[f, r0:r1]
		device.upload(b.ptr, np.zeros_like(full))
		device.equirect_to_cube_cvt(a.ptr, W, H, ch, fmt, b.ptr, N, ch, fmt, units=units, flags=flags)
		got = np.empty_like(full)
		device.download(got, b.ptr)
		assert same_bits(got, want.reshape(-1)), flags
	a.free(), b.free()
	for f in (1, 2, 4):
		win = device.equirect_src_window(W, H, N, 1 << f, 20, 61)
		part = gpu_equirect_device(src, W, H, ch, fmt, N, 1 << f, 20, 61, window=win)
		sl = slice((f * N + 20) * N * ch, (f * N + 61) * N * ch)
		assert same_bits(part[sl], full[sl]), f


@pytest.mark.parametrize("fmt,ch", [(0, 4), (3, 1), (1, 4), (2, 4)], ids=["u8x4", "f32x1", "u16x4", "f16x4"])
def test_tensor_map_tma_ring_is_bit_identical(oracle, monkeypatch, fmt, ch):
	"""The tensor-map experiment (CACAO_EQUIRECT_TMA=1, equirect_tma.cu: persistent warp-specialised CTAs, four boxes
	per 32 x 8 tile staged by cp.async.bulk.tensor.2d through a four-stage mbarrier ring, LDS taps) must give exactly
	the product kernels' result: tiles whose boxes fit, tiles that fall back to gathers (polar caps, the seam on
	the back face, magnifications the box does not cover), even / odd faces, faces smaller than a tile, more
	tiles than CTAs (the ring wraps many times), and a CTA count forced to one per SM."""
	monkeypatch.setenv("CACAO_EQUIRECT_TMA", "1")
	for (W, H, N) in ((512, 256, 128), (1024, 512, 255), (256, 128, 400), (2048, 1024, 96), (64, 32, 7), (4096, 2048, 1024)):
		src = oracle.synth(W * H * ch, fmt, ch, 19)
		want = oracle.equirect_to_cube(src, W, H, ch, fmt, N)
		assert same_bits(gpu_equirect_device(src, W, H, ch, fmt, N), want), (W, H, N)
	monkeypatch.setenv("CACAO_EQUIRECT_TMA_CTAS", "1")
	W, H, N = 1024, 512, 256
	src = oracle.synth(W * H * ch, fmt, ch, 20)
	assert same_bits(gpu_equirect_device(src, W, H, ch, fmt, N), oracle.equirect_to_cube(src, W, H, ch, fmt, N))


def test_seam_and_poles(oracle):
	"""One-hot panoramas: a texel on the u=0/1 seam and texels in the pole rows land identically."""
	from cacaoengine_b200 import _capi as c

	W, H, N = 64, 32, 33
	for (yy, xx) in ((16, 0), (16, 63), (0, 10), (31, 40)):
		img = np.zeros((H, W, 1), np.float32)
		img[yy, xx, 0] = 1000.0
		src = img.reshape(-1)
		assert same_bits(gpu_equirect_device(src, W, H, 1, c.FMT_F32, N), oracle.equirect_to_cube(src, W, H, 1, c.FMT_F32, N))


def test_python_mirror_equirect(oracle):
	import cacaoengine_b200 as ce

	W, H, N = 128, 64, 20
	src = oracle.synth(W * H * 3, ce.FMT_U16, 3, 8)
	out = ce.EquirectToCubemap(ce.ImageEx(W, H, ce.Layout.RGB, ce.FMT_U16, src.view(np.uint8)), N)
	assert same_bits(out, oracle.equirect_to_cube(src, W, H, 3, ce.FMT_U16, N))


@pytest.mark.parametrize("W,H,N,seed", [(8192, 4096, 2048, 0xC0FFEE + 3), (16384, 8192, 4096, 0xC0FFEE + 5)], ids=["cfg3", "cfg5"])
def test_full_size_shapes_band_sample(oracle, W, H, N, seed):
	"""BASELINE configs 3 and 5 at full size (8192x4096 -> 6x2048^2, 16384x8192 -> 6x4096^2, RGBA32F) on
	the device; three row bands per face are compared with the oracle bit for bit, and the whole cube
	is checked through size-independent properties: producing it as 6x4 separate (face, band)
	launches, as one cost-dealt unit list per rank of an 8-rank job, and through the pipelined
	host-memory call gives the same checksum as one launch."""
	from cacaoengine_b200 import _capi as c, device, sharding

	ch, fmt = 4, c.FMT_F32
	src = DevBuf(W * H * 16)
	device.synth(src.ptr, W * H * 4, fmt, 4, seed)
	cube = DevBuf(6 * N * N * 16)
	device.equirect_to_cube(src.ptr, W, H, ch, fmt, cube.ptr, N)
	whole = device.checksum(cube.ptr, 6 * N * N * 16)
	host_src = oracle.synth(W * H * 4, fmt, 4, seed)
	for (r0, r1) in ((0, 8), (N // 2 - 4, N // 2 + 4), (N - 8, N)):
		want = oracle.equirect_to_cube(host_src, W, H, ch, fmt, N, row_begin=r0, row_end=r1)
		for f in range(6):
			got = np.empty((r1 - r0) * N * 4, np.float32)
			off = (f * N + r0) * N
			device.download(got, cube.ptr + off * 16)
			assert same_bits(got, want[off * 4: off * 4 + got.size]), (f, r0)
		del want
	cube2 = DevBuf(6 * N * N * 16)
	for f in range(6):
		for b in range(4):
			device.equirect_to_cube(src.ptr, W, H, ch, fmt, cube2.ptr, N, 1 << f, b * (N // 4), (b + 1) * (N // 4))
	assert device.checksum(cube2.ptr, 6 * N * N * 16) == whole
	# the eight unit lists of a sharded job (mixed side / half-height polar bands, one launch each)
	device.upload(cube2.ptr, np.zeros(1 << 20, np.uint8))
	for rank in range(8):
		device.equirect_to_cube_units(src.ptr, W, H, ch, fmt, cube2.ptr, N, sharding.units_for_rank(rank, 8, N))
	assert device.checksum(cube2.ptr, 6 * N * N * 16) == whole
	cube2.free()
	# host -> host through pinned buffers: the pipelined upload / resample / readback
	sp, s_host = device.host_alloc(W * H * 16)
	dp, d_host = device.host_alloc(6 * N * N * 16)
	s_host[:] = host_src.view(np.uint8)
	device.equirect_to_cube(sp, W, H, ch, fmt, dp, N, mem=c.MEM_HOST)
	device.upload(cube.ptr, d_host)
	assert device.checksum(cube.ptr, 6 * N * N * 16) == whole
	device.host_free(sp), device.host_free(dp)
	for x in (src, cube):
		x.free()


def test_source_beyond_4_gib_matches_windowed_run():
	"""A 32768x16384 RGBA32F panorama is 8 GiB: byte offsets exceed 2^32 while pixel indices still fit
	32 bits.  A band resampled from the full buffer must equal the same band resampled from just its
	planned source-row window (different base addresses, same pixels)."""
	from cacaoengine_b200 import _capi as c, device

	W, H, N, ch, fmt = 32768, 16384, 1024, 4, c.FMT_F32
	full = DevBuf(W * H * 16)
	device.synth(full.ptr, W * H * 4, fmt, 4, 77)
	cube_a, cube_b = DevBuf(6 * N * N * 16), DevBuf(6 * N * N * 16)
	for f, (r0, r1) in ((3, (0, 256)), (0, (700, 1024)), (5, (512, 640))):  # -Y reaches the last source rows
		w0, wn = device.equirect_src_window(W, H, N, 1 << f, r0, r1)
		device.equirect_to_cube(full.ptr, W, H, ch, fmt, cube_a.ptr, N, 1 << f, r0, r1)
		win = DevBuf(wn * W * 16)
		device.synth(win.ptr, wn * W * 4, fmt, 4, 77, first_sample=w0 * W * 4)
		device.equirect_to_cube(win.ptr, W, H, ch, fmt, cube_b.ptr, N, 1 << f, r0, r1, src_row0=w0, src_rows=wn)
		off, n = (f * N + r0) * N * 16, (r1 - r0) * N * 16
		assert device.checksum(cube_a.ptr + off, n) == device.checksum(cube_b.ptr + off, n), (f, r0, r1)
		win.free()
	for b in (full, cube_a, cube_b):
		b.free()


def test_randomised_shapes_against_oracle(oracle):
	"""Seeded sweep over odd panorama / face sizes, formats, layouts, face masks, row bands and source
	windows -- every case bit-exact against the oracle."""
	from cacaoengine_b200 import device

	rng = np.random.default_rng(20260)
	for case in range(40):
		fmt = int(rng.integers(0, 4))
		ch = int(rng.choice([1, 3, 4]))
		W, H = int(rng.integers(8, 300)), int(rng.integers(4, 150))
		N = int(rng.integers(1, 90))
		mask = int(rng.integers(1, 64))
		r0 = int(rng.integers(0, N))
		r1 = int(rng.integers(r0 + 1, N + 1))
		src = oracle.synth(W * H * ch, fmt, ch, 1000 + case)
		want = oracle.equirect_to_cube(src, W, H, ch, fmt, N, face_mask=mask, row_begin=r0, row_end=r1)
		window = device.equirect_src_window(W, H, N, mask, r0, r1) if case % 2 else None
		got = gpu_equirect_device(src, W, H, ch, fmt, N, mask, r0, r1, window=window)
		assert same_bits(got, want), (case, fmt, ch, W, H, N, mask, r0, r1, window)
