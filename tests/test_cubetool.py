"""ce-cubetool: the reference tool's CLI contract (tools/cubetool/main.cpp:15-38, commands.hpp:9-11)
plus the new `project` subcommand."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ERR = "\x1b[0m\x1b[1;91mERROR: \x1b[0m"


@pytest.fixture(scope="module")
def tool(built_lib):
	from cacaoengine_b200 import build as b

	assert os.path.exists(b.TOOL)
	return b.TOOL


def run(tool, *args):
	return subprocess.run([tool, *args], capture_output=True, text=True)


def test_version_and_help(tool):
	r = run(tool, "--version")
	assert r.returncode == 0 and r.stdout.startswith("CubeTool v") and "For Cacao Engine v" in r.stdout
	assert run(tool, "-v").stdout == r.stdout
	h = run(tool, "--help")
	assert h.returncode == 0 and "SUBCOMMANDS:" in h.stdout and "create" in h.stdout and "extract" in h.stdout and "project" in h.stdout


def test_subcommand_is_required(tool):
	r = run(tool)
	assert r.returncode != 0 and "subcommand is required" in r.stderr.lower()


def test_reference_subcommands_validate_their_arguments(tool, tmp_path):
	"""create / extract keep the reference's option checks (create.cpp:14-26, extract.cpp:14-43):
	missing input file and bad -o are argument errors, bad content is a CUBE_ERROR with exit code 1."""
	for sub in ("create", "extract"):
		r = run(tool, sub, str(tmp_path / "missing"), "-o", str(tmp_path / "y"))
		assert r.returncode != 0 and "File does not exist" in r.stderr and r.stdout == ""
		assert run(tool, sub, "--help").returncode == 0
	junk = tmp_path / "junk.bin"
	junk.write_bytes(b"junk: [1, 2]\n")
	r = run(tool, "create", str(junk), "-o", str(tmp_path / "o.xjc"))
	assert r.returncode == 1 and r.stderr.startswith(ERR) and "right (positive X) face node is invalid: Node does not exist" in r.stderr
	r = run(tool, "extract", str(junk), "-A", "-o", str(tmp_path / "o"))
	assert r.returncode == 1 and r.stderr.startswith(ERR) and "Stream is not of a Cacao Engine packed object!" in r.stderr
	r = run(tool, "extract", str(junk), "-o", str(tmp_path / "o"))
	assert r.returncode == 1 and "No faces selected for extraction!" in r.stderr
	assert "Invalid face name" in run(tool, "extract", str(junk), "-f", "sideways").stderr
	assert run(tool, "extract", str(junk), "-A", "-f", "left").returncode != 0


def test_project_argument_validation(tool, tmp_path):
	assert run(tool, "project").returncode != 0
	assert "File does not exist" in run(tool, "project", str(tmp_path / "nope.png"), "-o", str(tmp_path)).stderr
	np.save(tmp_path / "p.npy", np.zeros((4, 8, 3), np.uint8))
	assert "-o is required" in run(tool, "project", str(tmp_path / "p.npy")).stderr
	assert "Invalid face name" in run(tool, "project", str(tmp_path / "p.npy"), "-o", str(tmp_path), "-f", "sideways").stderr
	bad = tmp_path / "bad.bin"
	bad.write_bytes(b"not an image")
	r = run(tool, "project", str(bad), "-o", str(tmp_path / "o"))
	assert r.returncode == 1 and r.stderr.startswith(ERR) and "Unrecognised image format" in r.stderr


def test_silent_by_default_verbose_on_request(tool, tmp_path):
	np.save(tmp_path / "p.npy", np.zeros((4, 8, 3), np.uint8))
	from cacaoengine_b200 import device

	if device.device_count() == 0:
		r = run(tool, "project", str(tmp_path / "p.npy"), "-o", str(tmp_path / "o"))
		assert r.returncode == 1 and r.stderr.startswith(ERR) and r.stdout == ""  # loud failure, no CPU fallback
		v = run(tool, "-V", "project", str(tmp_path / "p.npy"), "-o", str(tmp_path / "o"))
		assert "Opening input file" in v.stdout


@pytest.mark.gpu
def test_project_png_end_to_end(tool, tmp_path, oracle):
	import cv2

	W, H, N = 256, 128, 48
	pano = oracle.synth(W * H * 3, oracle.U8, 3, 31).reshape(H, W, 3)
	cv2.imwrite(str(tmp_path / "sky.png"), pano[:, :, ::-1])  # cv2 wants BGR
	r = run(tool, "project", str(tmp_path / "sky.png"), "-o", str(tmp_path / "out"), "--face-size", str(N))
	assert r.returncode == 0 and r.stdout == "", r.stderr
	want = oracle.equirect_to_cube(pano.reshape(-1), W, H, 3, oracle.U8, N).reshape(6, N, N, 3)
	names = ["right", "left", "up", "down", "front", "back"]
	for k, nm in enumerate(names):
		face = cv2.imread(str(tmp_path / "out" / f"sky_{nm}.png"), cv2.IMREAD_UNCHANGED)[:, :, ::-1]
		assert (face == want[k]).all(), nm
	ajc = (tmp_path / "out" / "sky.ajc").read_text().splitlines()
	assert ajc == [f"{key}: sky_{nm}.png" for key, nm in zip(["right", "left", "top", "bottom", "front", "back"], names)]


@pytest.mark.gpu
def test_project_float_npy_16bit_png_and_options(tool, tmp_path, oracle):
	import cv2

	W, H, N = 128, 64, 20
	pano = oracle.synth(W * H * 4, oracle.F32, 4, 32).reshape(H, W, 4)
	np.save(tmp_path / "hdr.npy", pano)
	r = run(tool, "-V", "project", str(tmp_path / "hdr.npy"), "-o", str(tmp_path / "o1"), "-s", str(N), "--bench")
	assert r.returncode == 0 and "Projecting onto 6 faces" in r.stdout and '"subcommand": "project"' in r.stdout
	want = oracle.equirect_to_cube(pano.reshape(-1), W, H, 4, oracle.F32, N).reshape(6, N, N, 4)
	got = np.load(tmp_path / "o1" / "hdr_front.npy")
	assert got.dtype == np.float32 and (got.view(np.uint32) == want[4].view(np.uint32)).all()
	# float panorama to half-float RGBA faces: one fused pass inside (Spec B.4), same answer as two
	np.save(tmp_path / "rgb.npy", pano[:, :, :3].copy())
	r = run(tool, "project", str(tmp_path / "rgb.npy"), "-o", str(tmp_path / "o3"), "-s", str(N), "--layout", "rgba", "--depth", "f16", "--format", "npy")
	assert r.returncode == 0, r.stderr
	want16 = oracle.equirect_to_cube_cvt(pano[:, :, :3].reshape(-1), W, H, 3, oracle.F32, N, 4, oracle.F16).reshape(6, N, N, 4)
	got16 = np.load(tmp_path / "o3" / "rgb_down.npy")
	assert got16.dtype == np.float16 and (got16.view(np.uint16) == want16[3].view(np.uint16)).all()
	# 16-bit PNG in, single face out, converted to 8-bit RGB and flipped
	p16 = oracle.synth(W * H * 4, oracle.U16, 4, 33).reshape(H, W, 4)
	cv2.imwrite(str(tmp_path / "deep.png"), p16[:, :, [2, 1, 0, 3]])
	r = run(tool, "project", str(tmp_path / "deep.png"), "-o", str(tmp_path / "o2"), "-s", str(N), "-f", "up", "--layout", "rgb", "--depth", "8", "--flip")
	assert r.returncode == 0, r.stderr
	cube16 = oracle.equirect_to_cube(p16.reshape(-1), W, H, 4, oracle.U16, N).reshape(6, N * N * 4)
	up8 = oracle.convert(cube16[2], 4, 3, oracle.U16, oracle.U8).reshape(N, N, 3)[::-1]
	face = cv2.imread(str(tmp_path / "o2" / "deep_up.png"), cv2.IMREAD_UNCHANGED)[:, :, ::-1]
	assert (face == up8).all()
	assert sorted(os.listdir(tmp_path / "o2")) == ["deep_up.png"]  # no .ajc for a partial set


@pytest.mark.gpu
def test_project_integer_panoramas_convert_without_the_reference_quirks(tool, tmp_path, oracle):
	"""`project --layout/--depth` on integer panoramas runs projection, then conversion.  That conversion is
	part of a NEW operation, so it uses the fixed rules, not the reference's bugs (Layout.cpp:50-51 repeats
	pixel 0 for Gray -> RGBA, :75 clamps 16-bit gray to 255): a gray panorama with --layout rgba keeps its
	picture, a 16-bit RGB panorama with --layout gray keeps its range.  One GPU (device-resident chain) and
	the host-memory path (--gpus semantics of EquirectToCubemap) must agree."""
	W, H, N = 128, 64, 24
	gray = oracle.synth(W * H, oracle.U8, 1, 51).reshape(H, W)
	np.save(tmp_path / "g.npy", gray)
	r = run(tool, "project", str(tmp_path / "g.npy"), "-o", str(tmp_path / "o1"), "-s", str(N), "--layout", "rgba", "--format", "npy")
	assert r.returncode == 0, r.stderr
	cube = oracle.equirect_to_cube(gray.reshape(-1), W, H, 1, oracle.U8, N).reshape(6, N * N)
	for k, nm in enumerate(["right", "left", "up", "down", "front", "back"]):
		got = np.load(tmp_path / "o1" / f"g_{nm}.npy")
		want = oracle.convert(cube[k], 1, 4, oracle.U8, oracle.U8, mode=1).reshape(N, N, 4)
		assert (got == want).all(), nm
		assert len(np.unique(got[:, :, 0])) > 16, "faces must not be one flat colour (the pixel-0 quirk)"
	rgb16 = oracle.synth(W * H * 3, oracle.U16, 3, 52).reshape(H, W, 3)
	np.save(tmp_path / "c.npy", rgb16)
	r = run(tool, "project", str(tmp_path / "c.npy"), "-o", str(tmp_path / "o2"), "-s", str(N), "--layout", "gray", "--format", "npy")
	assert r.returncode == 0, r.stderr
	cube = oracle.equirect_to_cube(rgb16.reshape(-1), W, H, 3, oracle.U16, N).reshape(6, N * N * 3)
	got = np.load(tmp_path / "o2" / "c_front.npy")
	want = oracle.convert(cube[4], 3, 1, oracle.U16, oracle.U16, mode=1).reshape(got.shape)
	assert got.dtype == np.uint16 and (got == want).all() and got.max() > 255
