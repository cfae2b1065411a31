"""Packed / unpacked cubemap files: the VP8L codec, the .xjc container
and the `create` / `extract` subcommands.

The reference's own codec stack (libwebp 1.5.0, bzip2, yaml-cpp; un-vendored meson wraps) is absent
here, so file-format parity is pinned against independent implementations of the same formats:
Pillow's bundled libwebp for lossless WebP, Python's bz2 module for the payload stream, and the byte
layout of libs/formats/src/PackedContainer.cpp:108-158 / PackedEncode.cpp:18-48 typed out below."""
import bz2
import io
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ERR = "\x1b[0m\x1b[1;91mERROR: \x1b[0m"
AJC_KEYS = ["right", "left", "top", "bottom", "front", "back"]
FACE_NAMES = ["right", "left", "up", "down", "front", "back"]

PIL_Image = pytest.importorskip("PIL.Image")
from PIL import features  # noqa: E402

pytestmark = pytest.mark.skipif(not features.check("webp"), reason="Pillow without WebP support")


@pytest.fixture(scope="module")
def tool(built_lib):
	from cacaoengine_b200 import build as b

	return b.TOOL


@pytest.fixture(scope="module")
def codec(tmp_path_factory):
	exe = tmp_path_factory.mktemp("vp8l") / "vp8l_cli"
	subprocess.run(["g++", "-std=c++20", "-O2", "-Wall", "-I" + os.path.join(ROOT, "tools/cubetool"), os.path.join(ROOT, "tests/cpp/vp8l_cli.cpp"), os.path.join(ROOT, "tools/cubetool/vp8l.cpp"), "-o", str(exe)], check=True)
	return str(exe)


def run(*args):
	return subprocess.run(list(args), capture_output=True, text=True)


def picture(rng, h, w, c, kind):
	if kind == "noise":
		return rng.integers(0, 256, (h, w, c), dtype=np.uint8)
	if kind == "flat":
		return np.full((h, w, c), 77, np.uint8)
	if kind.startswith("pal"):
		n = int(kind[3:])
		pal = rng.integers(0, 256, (n, c), dtype=np.uint8)
		y, x = np.mgrid[0:h, 0:w]
		return pal[(y // 3 + x // 5 + rng.integers(0, 2, (h, w))) % n]
	y, x = np.mgrid[0:h, 0:w]
	a = np.stack([127 + 100 * np.sin(x / (17 + k)) * np.cos(y / (23 - k)) for k in range(c)], -1)
	return (a + rng.normal(0, 2, a.shape)).clip(0, 255).astype(np.uint8)


def libwebp_lossless(a, method):
	b = io.BytesIO()
	PIL_Image.fromarray(a, "RGBA" if a.shape[2] == 4 else "RGB").save(b, "WEBP", lossless=True, method=method, exact=True)
	return b.getvalue()


CASES = [(1, 1, 3, "noise"), (7, 5, 4, "noise"), (64, 64, 3, "smooth"), (200, 333, 4, "smooth"), (50, 50, 3, "pal5"), (31, 17, 4, "pal2"), (128, 128, 3, "flat"),
		 (257, 129, 3, "smooth"), (40, 40, 3, "pal13"), (513, 100, 3, "smooth"), (90, 90, 4, "pal200"), (300, 300, 4, "noise")]


@pytest.mark.parametrize("h,w,c,kind", CASES)
def test_decoder_reads_libwebp_streams(codec, tmp_path, h, w, c, kind):
	"""Files written by libwebp (every effort level: palettes, pixel bundling, predictor / cross-colour
	/ subtract-green transforms, colour cache, meta prefix codes, LZ77) decode to the same pixels."""
	a = picture(np.random.default_rng(h * 1000 + w), h, w, c, kind)
	for method in (0, 3, 6):
		f = tmp_path / "in.webp"
		f.write_bytes(libwebp_lossless(a, method))
		r = run(codec, "dec", str(f), str(tmp_path / "out.raw"))
		assert r.returncode == 0, r.stderr
		gw, gh, gc = map(int, r.stdout.split())
		got = np.fromfile(tmp_path / "out.raw", np.uint8).reshape(gh, gw, gc)
		assert (gh, gw) == (h, w)
		assert np.array_equal(got[..., :3], a[..., :3])
		if gc == 4:
			assert np.array_equal(got, a)
		else:  # alpha hint off: every pixel was opaque (or the source had no alpha)
			assert c == 3 or (a[..., 3] == 255).all()


@pytest.mark.parametrize("h,w,c,kind", CASES + [(2, 300, 3, "smooth"), (300, 2, 4, "noise"), (33, 65, 4, "smooth")])
def test_encoder_output_is_read_by_libwebp(codec, tmp_path, h, w, c, kind):
	a = picture(np.random.default_rng(h * 7 + w), h, w, c, kind)
	a.tofile(tmp_path / "in.raw")
	r = run(codec, "enc", str(tmp_path / "in.raw"), str(w), str(h), str(c), str(tmp_path / "e.webp"))
	assert r.returncode == 0, r.stderr
	im = PIL_Image.open(tmp_path / "e.webp")
	got = np.asarray(im.convert("RGBA" if c == 4 else "RGB"))
	assert np.array_equal(got, a)
	# and by this decoder
	r = run(codec, "dec", str(tmp_path / "e.webp"), str(tmp_path / "o.raw"))
	gw, gh, gc = map(int, r.stdout.split())
	assert np.array_equal(np.fromfile(tmp_path / "o.raw", np.uint8).reshape(gh, gw, gc)[..., :3], a[..., :3])
	if kind == "smooth":
		assert os.path.getsize(tmp_path / "e.webp") < 0.8 * a.nbytes  # it does compress


def test_decoder_rejects_garbage(codec, tmp_path):
	good = libwebp_lossless(picture(np.random.default_rng(1), 40, 40, 3, "smooth"), 4)
	for name, blob in (("short", good[: len(good) // 2]), ("notwebp", b"RIFF\x10\x00\x00\x00WAVEfmt 0123456789"), ("empty", b"")):
		(tmp_path / name).write_bytes(blob)
		r = run(codec, "dec", str(tmp_path / name), str(tmp_path / "o.raw"))
		assert r.returncode == 1 and "WebP lossless" in r.stderr
	lossy = io.BytesIO()
	PIL_Image.fromarray(picture(np.random.default_rng(2), 32, 32, 3, "smooth")).save(lossy, "WEBP", quality=50)
	(tmp_path / "lossy").write_bytes(lossy.getvalue())
	r = run(codec, "dec", str(tmp_path / "lossy"), str(tmp_path / "o.raw"))
	assert r.returncode == 1 and "lossy" in r.stderr


def test_decoder_survives_damaged_streams(codec, tmp_path):
	"""Bit flips, byte changes and truncation anywhere after the RIFF header: the decoder either
	decodes or reports an error -- it never crashes or hangs (the same sweep ran clean under
	ASan + UBSan, 1500 cases)."""
	rng = np.random.default_rng(11)
	seeds = [libwebp_lossless(picture(rng, 40, 40, 3, "smooth"), 0), libwebp_lossless(picture(rng, 33, 17, 4, "smooth"), 6), libwebp_lossless(picture(rng, 30, 30, 3, "pal5"), 6)]
	for it in range(150):
		s = bytearray(seeds[it % len(seeds)])
		for _ in range(int(rng.integers(1, 5))):
			p = int(rng.integers(12, len(s)))
			mode = int(rng.integers(0, 3))
			if mode == 0:
				s[p] ^= 1 << int(rng.integers(0, 8))
			elif mode == 1:
				s[p] = int(rng.integers(0, 256))
			elif len(s) > 40:
				s = s[:p]
		(tmp_path / "f.webp").write_bytes(bytes(s))
		r = subprocess.run([codec, "dec", str(tmp_path / "f.webp"), str(tmp_path / "o.raw")], capture_output=True, text=True, timeout=60)
		assert r.returncode == 0 or (r.returncode == 1 and "WebP lossless" in r.stderr), (it, r.returncode, r.stderr[-200:])


def make_faces(rng, n=96):
	faces = []
	for i in range(6):
		a = picture(rng, n, n, 4 if i % 2 else 3, "smooth")
		if a.shape[2] == 4:
			a[..., 3] = rng.integers(0, 256, (n, n))
		faces.append(a)
	return faces


def test_create_writes_the_reference_container_layout(tool, tmp_path):
	"""create: CA CA 00 C4, u16 version 1, u64 size, one bzip2 stream of 6 x u64 + six lossless WebP
	blobs (PackedContainer.cpp:108-158, PackedEncode.cpp:18-48) -- checked with bz2 + libwebp."""
	faces = make_faces(np.random.default_rng(3))
	for k, a in zip(AJC_KEYS, faces):
		PIL_Image.fromarray(a).save(tmp_path / f"f_{k}.png")
	# quoted and plain scalars, comments, document marker, faces relative to the .ajc
	(tmp_path / "sky.ajc").write_text("# sky\n---\n" + "".join(f'{k}: "f_{k}.png"\n' if i % 2 else f"{k}: f_{k}.png  # face\n" for i, k in enumerate(AJC_KEYS)))
	r = run(tool, "create", str(tmp_path / "sky.ajc"), "-o", str(tmp_path / "sky.xjc"))
	assert r.returncode == 0 and r.stdout == "" and r.stderr == ""
	f = (tmp_path / "sky.xjc").read_bytes()
	assert f[:4] == bytes([0xCA, 0xCA, 0x00, 0xC4])
	version, size = struct.unpack("<HQ", f[4:14])
	payload = bz2.decompress(f[14:])
	assert version == 1 and size == len(payload)
	sizes = struct.unpack("<6Q", payload[:48])
	assert 48 + sum(sizes) == len(payload)
	off = 48
	for i, s in enumerate(sizes):
		im = PIL_Image.open(io.BytesIO(payload[off:off + s]))
		off += s
		assert im.format == "WEBP" and np.array_equal(np.asarray(im), faces[i])
	# verbose mode narrates like the reference
	v = run(tool, "-V", "create", str(tmp_path / "sky.ajc"), "-o", str(tmp_path / "sky2.xjc"))
	assert "Loading input files:" in v.stdout and "Generating output... Done." in v.stdout


def test_extract_reads_reference_style_files(tool, tmp_path):
	"""extract on a container assembled the way the reference does (libwebp blobs, bzip2 -9)."""
	faces = make_faces(np.random.default_rng(4))
	blobs = [libwebp_lossless(a, 6) for a in faces]
	payload = struct.pack("<6Q", *map(len, blobs)) + b"".join(blobs)
	(tmp_path / "ref.xjc").write_bytes(bytes([0xCA, 0xCA, 0x00, 0xC4]) + struct.pack("<HQ", 1, len(payload)) + bz2.compress(payload, 9))
	r = run(tool, "extract", str(tmp_path / "ref.xjc"), "-A", "-o", str(tmp_path / "all"))
	assert r.returncode == 0, r.stderr
	assert sorted(os.listdir(tmp_path / "all")) == sorted(f"ref_{n}.png" for n in FACE_NAMES)
	for n, a in zip(FACE_NAMES, faces):
		assert np.array_equal(np.asarray(PIL_Image.open(tmp_path / "all" / f"ref_{n}.png")), a)
	r = run(tool, "extract", str(tmp_path / "ref.xjc"), "-f", "left", "up", "-o", str(tmp_path / "two"))
	assert r.returncode == 0 and sorted(os.listdir(tmp_path / "two")) == ["ref_left.png", "ref_up.png"]


def test_create_then_extract_round_trip_and_errors(tool, tmp_path):
	faces = make_faces(np.random.default_rng(5), 64)
	for k, a in zip(AJC_KEYS, faces):
		PIL_Image.fromarray(a).save(tmp_path / f"{k}.png")
	(tmp_path / "c.ajc").write_text("".join(f"{k}: {k}.png\n" for k in AJC_KEYS))
	assert run(tool, "create", str(tmp_path / "c.ajc"), "-o", str(tmp_path / "c.xjc")).returncode == 0
	assert run(tool, "extract", str(tmp_path / "c.xjc"), "--all-faces", "-o", str(tmp_path / "o")).returncode == 0
	for n, a in zip(FACE_NAMES, faces):
		assert np.array_equal(np.asarray(PIL_Image.open(tmp_path / "o" / f"c_{n}.png")), a)
	# the reference's error texts
	(tmp_path / "missing.ajc").write_text("".join(f"{k}: nowhere_{k}.png\n" for k in AJC_KEYS))
	r = run(tool, "create", str(tmp_path / "missing.ajc"), "-o", str(tmp_path / "m.xjc"))
	assert r.returncode == 1 and r.stderr == ERR + "Input file specifies a nonexistent file!\n"
	(tmp_path / "partial.ajc").write_text("".join(f"{k}: {k}.png\n" for k in AJC_KEYS[:4]))
	r = run(tool, "create", str(tmp_path / "partial.ajc"), "-o", str(tmp_path / "p.xjc"))
	assert r.returncode == 1 and "While parsing unpacked cubemap data, front (positive Z) face node is invalid: Node does not exist" in r.stderr
	(tmp_path / "list.ajc").write_text("".join(f"{k}: {k}.png\n" for k in AJC_KEYS[:5]) + "back:\n  - a\n  - b\n")
	r = run(tool, "create", str(tmp_path / "list.ajc"), "-o", str(tmp_path / "l.xjc"))
	assert r.returncode == 1 and "back (negative Z) face node is invalid: Node is of improper type!" in r.stderr
	# 16-bit and grayscale faces are rejected by the WebP stage exactly as in the reference (WebP.cpp:76,78)
	PIL_Image.fromarray(np.zeros((8, 8), np.uint8)).save(tmp_path / "gray.png")
	(tmp_path / "g.ajc").write_text("".join(f"{k}: gray.png\n" for k in AJC_KEYS))
	r = run(tool, "create", str(tmp_path / "g.ajc"), "-o", str(tmp_path / "g.xjc"))
	assert r.returncode == 1 and r.stderr == ERR + "Cannot encode a grayscale image to WebP format!\n"
	PIL_Image.fromarray(np.zeros((8, 8), np.uint16)).save(tmp_path / "deep.png")
	(tmp_path / "d.ajc").write_text("".join(f"{k}: deep.png\n" for k in AJC_KEYS))
	r = run(tool, "create", str(tmp_path / "d.ajc"), "-o", str(tmp_path / "d.xjc"))
	assert r.returncode == 1 and r.stderr == ERR + "Invalid color depth for WebP encoding; only 8 bits per channel are supported.\n"
	# damaged containers
	good = (tmp_path / "c.xjc").read_bytes()
	(tmp_path / "cut.xjc").write_bytes(good[: len(good) // 2])
	r = run(tool, "extract", str(tmp_path / "cut.xjc"), "-A", "-o", str(tmp_path / "x"))
	assert r.returncode == 1 and r.stderr == ERR + "Failed to decompress packed container payload!\n"
	(tmp_path / "shader.xjc").write_bytes(good[:3] + b"\x1b" + good[4:])
	r = run(tool, "extract", str(tmp_path / "shader.xjc"), "-A", "-o", str(tmp_path / "x"))
	assert r.returncode == 1 and r.stderr == ERR + "Packed container provided for cubemap decoding is not a cubemap!\n"


@pytest.mark.gpu
def test_project_create_extract_pipeline_on_gpu(tool, oracle, tmp_path):
	"""The whole tool chain: panorama -> `project` (GPU) -> `create --to-rgb8` (GPU: the engine's fused
	16 -> 8 bit + -> RGB conversion) -> `extract --flip` (GPU row mirror); every stage checked against
	the oracle."""
	from cacaoengine_b200 import _capi as c

	W, H, N = 256, 128, 48
	pano = oracle.synth(W * H * 4, c.FMT_U16, 4, 99).reshape(H, W, 4)
	PIL_Image  # 16-bit RGBA PNGs are beyond Pillow's writer: hand the panorama over as .npy
	np.save(tmp_path / "pano.npy", pano)
	r = run(tool, "project", str(tmp_path / "pano.npy"), "-o", str(tmp_path / "faces"), "-s", str(N))
	assert r.returncode == 0, r.stderr
	r = run(tool, "create", str(tmp_path / "faces" / "pano.ajc"), "-o", str(tmp_path / "pano.xjc"))
	assert r.returncode == 1 and "only 8 bits per channel" in r.stderr  # reference behaviour without the flag
	r = run(tool, "create", str(tmp_path / "faces" / "pano.ajc"), "-o", str(tmp_path / "pano.xjc"), "--to-rgb8")
	assert r.returncode == 0, r.stderr
	cube = oracle.equirect_to_cube(pano.reshape(-1), W, H, 4, c.FMT_U16, N).reshape(6, N, N, 4)
	want = [oracle.convert(cube[i].reshape(-1), 4, 3, c.FMT_U16, c.FMT_U8).reshape(N, N, 3) for i in range(6)]
	f = (tmp_path / "pano.xjc").read_bytes()
	payload = bz2.decompress(f[14:])
	sizes = struct.unpack("<6Q", payload[:48])
	off = 48
	for i, s in enumerate(sizes):
		assert np.array_equal(np.asarray(PIL_Image.open(io.BytesIO(payload[off:off + s]))), want[i])
		off += s
	assert run(tool, "extract", str(tmp_path / "pano.xjc"), "-A", "-o", str(tmp_path / "x"), "--flip").returncode == 0
	for i, n in enumerate(FACE_NAMES):
		assert np.array_equal(np.asarray(PIL_Image.open(tmp_path / "x" / f"pano_{n}.png")), want[i][::-1])
