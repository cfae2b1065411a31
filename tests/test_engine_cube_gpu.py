"""cacao_img_engine_cube: the engine's cubemap load chain as one call -- Convert16To8BitColor when 16-bit,
ChangeChannelLayout to RGB when not RGB (engine/src/core/Cubemap.cpp:28-31), the GL backend's row mirror
(engine/src/opengl/impls/OpenGLCubemap.cpp:25), the six faces contiguous (engine/src/vulkan/impls/VulkanCubemap.cpp:30-41),
optionally the mip chain.  Checked against the restatement and, where oracle/_ref travelled, against the
reference's OWN two functions run face by face."""
import numpy as np
import pytest

from gpu_util import DevBuf


def test_cube_size_is_a_host_side_calculation(built_lib):
	from cacaoengine_b200 import device

	assert device.engine_cube_bytes(4, 4) == 6 * 16 * 3
	assert device.engine_cube_bytes(4, 4, device.CUBE_MIPS) == 6 * 3 * (16 + 4 + 1)
	assert device.engine_cube_bytes(5, 3, device.CUBE_MIPS) == 6 * 3 * (15 + 2 + 1)
	assert device.engine_cube_bytes(1, 1, device.CUBE_MIPS) == 18
	assert device.engine_cube_bytes(2048, 2048, device.CUBE_FLIP_ROWS) == 6 * 2048 * 2048 * 3


def _faces(oracle, w, h, ch, fmt, seed):
	return oracle.synth(6 * w * h * ch, fmt, ch, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("ch", (1, 3, 4))
@pytest.mark.parametrize("bits", (8, 16))
def test_every_face_kind_every_flag_device_and_host(built_lib, oracle, ch, bits):
	from cacaoengine_b200 import _capi as c, device

	device.init()
	fmt = c.FMT_U8 if bits == 8 else c.FMT_U16
	for (w, h) in ((64, 64), (5, 3), (1, 1), (96, 40), (256, 256)):
		src = _faces(oracle, w, h, ch, fmt, 60 + w)
		face_bytes = w * h * ch * (bits // 8)
		for flags in (0, device.CUBE_FLIP_ROWS, device.CUBE_MIPS, device.CUBE_FLIP_ROWS | device.CUBE_MIPS):
			want = oracle.engine_cube(src, w, h, ch, flip=bool(flags & device.CUBE_FLIP_ROWS), mips=bool(flags & device.CUBE_MIPS))
			assert want.size == device.engine_cube_bytes(w, h, flags)
			if oracle.have_ref() and not (flags & device.CUBE_MIPS):
				ref, _ = oracle.ref_engine_cube(src, w, h, ch, flip=bool(flags & device.CUBE_FLIP_ROWS))
				assert (ref == want).all(), "restatement differs from the reference's own functions"
			# host memory (pageable numpy), faces as six separate arrays
			sep = [src.view(np.uint8)[f * face_bytes:(f + 1) * face_bytes].copy() for f in range(6)]
			got = np.full(want.size, 0xEE, np.uint8)
			device.engine_cube([s.ctypes.data for s in sep], w, h, ch, fmt, got.ctypes.data, flags, mem=c.MEM_HOST)
			assert (got == want).all(), ("host", w, h, flags)
			# device memory: contiguous faces (one launch) and scattered ones (six)
			a, b = DevBuf(src.nbytes), DevBuf(want.size)
			device.upload(a.ptr, src)
			device.engine_cube([a.ptr + f * face_bytes for f in range(6)], w, h, ch, fmt, b.ptr, flags)
			device.download(got, b.ptr)
			assert (got == want).all(), ("device", w, h, flags)
			a.free(), b.free()
			bufs = [DevBuf(face_bytes) for _ in range(6)]
			b = DevBuf(want.size)
			for f in range(6):
				device.upload(bufs[f].ptr, sep[f])
			device.engine_cube([x.ptr for x in bufs], w, h, ch, fmt, b.ptr, flags)
			device.download(got, b.ptr)
			assert (got == want).all(), ("device, scattered", w, h, flags)
			for x in bufs:
				x.free()
			b.free()


@pytest.mark.gpu
def test_pinned_pipeline_at_engine_size_against_the_compiled_reference(built_lib, oracle):
	"""Six 1024^2 RGBA16 faces (the engine's case) from pinned buffers, flipped: equal to
	Convert16To8BitColor + ChangeChannelLayout of the reference itself, face by face."""
	from cacaoengine_b200 import _capi as c, device

	device.init()
	w = h = 1024
	src = _faces(oracle, w, h, 4, c.FMT_U16, 99)
	nbytes = device.engine_cube_bytes(w, h, device.CUBE_FLIP_ROWS)
	hp_in, h_in = device.host_alloc(src.nbytes)
	hp_out, h_out = device.host_alloc(nbytes)
	h_in[:] = src.view(np.uint8)
	device.engine_cube([hp_in + f * w * h * 8 for f in range(6)], w, h, 4, c.FMT_U16, hp_out, device.CUBE_FLIP_ROWS, mem=c.MEM_HOST)
	want = oracle.ref_engine_cube(src, w, h, 4, flip=True)[0] if oracle.have_ref() else oracle.engine_cube(src, w, h, 4, flip=True)
	assert (h_out == want).all()
	device.host_free(hp_in), device.host_free(hp_out)


@pytest.mark.gpu
def test_argument_errors(built_lib, oracle):
	from cacaoengine_b200 import _capi as c, device

	device.init()
	buf = np.zeros(6 * 16 * 4, np.uint8)
	out = np.zeros(6 * 16 * 3, np.uint8)
	ptrs = [buf.ctypes.data + f * 64 for f in range(6)]
	for bad in (lambda: device.engine_cube(ptrs, 4, 4, 2, c.FMT_U8, out.ctypes.data, 0, mem=c.MEM_HOST),
				lambda: device.engine_cube(ptrs, 4, 4, 4, c.FMT_F16, out.ctypes.data, 0, mem=c.MEM_HOST),
				lambda: device.engine_cube(ptrs, 0, 4, 4, c.FMT_U8, out.ctypes.data, 0, mem=c.MEM_HOST),
				lambda: device.engine_cube(ptrs, 4, 4, 4, c.FMT_U8, out.ctypes.data, 8, mem=c.MEM_HOST),
				lambda: device.engine_cube(ptrs, 4, 4, 4, c.FMT_U8, buf.ctypes.data + 32, 0, mem=c.MEM_HOST)):
		with pytest.raises(c.CacaoError):
			bad()


@pytest.mark.gpu
@pytest.mark.parametrize("ch", (1, 3, 4))
@pytest.mark.parametrize("bits", (8, 16))
def test_texture_chain_every_kind_every_flag(built_lib, oracle, ch, bits):
	"""cacao_img_engine_texture: Tex2D's load chain (Convert16To8BitColor with the layout kept, Tex2D.cpp:18; Flip
	and the mip levels of OpenGLTex2D.cpp:16,51) -- device memory, pageable host memory in one band and in several
	(a 2.4 Mpixel image goes through as four bands, each mirrored in itself and placed at the mirrored band)."""
	from cacaoengine_b200 import _capi as c, device

	device.init()
	fmt = c.FMT_U8 if bits == 8 else c.FMT_U16
	for (w, h) in ((64, 64), (5, 3), (1, 1), (97, 41), (1600, 1501)):
		src = oracle.synth(w * h * ch, fmt, ch, 80 + w)
		for flags in (0, device.CUBE_FLIP_ROWS, device.CUBE_MIPS, device.CUBE_FLIP_ROWS | device.CUBE_MIPS):
			want = oracle.engine_texture(src, w, h, ch, flip=bool(flags & device.CUBE_FLIP_ROWS), mips=bool(flags & device.CUBE_MIPS))
			assert want.size == device.engine_texture_bytes(w, h, ch, flags)
			if oracle.have_ref() and bits == 16 and flags == 0:
				assert (oracle.ref_depth16to8(src, w, h, ch) == want).all()
			got = np.full(want.size, 0xEE, np.uint8)
			device.engine_texture(src.ctypes.data, w, h, ch, fmt, got.ctypes.data, flags, mem=c.MEM_HOST)
			assert (got == want).all(), ("host", w, h, flags)
			a, b = DevBuf(src.nbytes), DevBuf(want.size)
			device.upload(a.ptr, src)
			device.engine_texture(a.ptr, w, h, ch, fmt, b.ptr, flags)
			device.download(got, b.ptr)
			assert (got == want).all(), ("device", w, h, flags)
			a.free(), b.free()
	assert device.engine_texture_bytes(4, 4, 2) == 0
	with pytest.raises(c.CacaoError):
		device.engine_texture(src.ctypes.data, 4, 4, 4, c.FMT_F32, got.ctypes.data, 0, mem=c.MEM_HOST)


@pytest.mark.gpu
def test_host_register_pins_caller_memory(built_lib, oracle, capfd, monkeypatch):
	"""cacao_img_host_register: memory the caller owns (a numpy array here; a std::vector's storage or a
	shared-memory segment in bench.py's sharded e2e leg) becomes DMA-able in place -- the host-memory call takes
	the pinned pipeline instead of staging through the ring -- and the result is unchanged."""
	from cacaoengine_b200 import _capi as c, device

	device.init()
	px = 3 << 20
	src = oracle.synth(px * 4, c.FMT_U8, 4, 71)
	out = np.empty(px * 4, np.float16)
	want = oracle.convert(src, 4, 4, c.FMT_U8, c.FMT_F16)
	monkeypatch.setenv("CACAO_TRACE", "1")
	device.convert(src.ctypes.data, out.ctypes.data, px, 4, 4, c.FMT_U8, c.FMT_F16, mem=c.MEM_HOST)
	assert (out.view(np.uint16) == want.view(np.uint16)).all()
	device.host_register(src.ctypes.data, src.nbytes)
	device.host_register(out.ctypes.data, out.nbytes)
	try:
		out[:] = 0
		device.convert(src.ctypes.data, out.ctypes.data, px, 4, 4, c.FMT_U8, c.FMT_F16, mem=c.MEM_HOST)
		assert (out.view(np.uint16) == want.view(np.uint16)).all()
	finally:
		device.host_unregister(src.ctypes.data)
		device.host_unregister(out.ctypes.data)
	with pytest.raises(c.CacaoError):
		device.host_register(0, 16)


@pytest.mark.gpu
def test_load_chains_from_concurrent_threads(built_lib, oracle):
	"""The engine loads assets from thread-pool tasks: cube chains, texture chains and a resampling call issued
	from eight threads at once (ctypes drops the GIL) share one device's staging buffers and streams behind the
	pipeline mutex -- every result must still be exact."""
	from concurrent.futures import ThreadPoolExecutor

	from cacaoengine_b200 import _capi as c, device

	device.init()
	cube_src = oracle.synth(6 * 128 * 96 * 4, c.FMT_U16, 4, 91)
	cube_want = oracle.engine_cube(cube_src, 128, 96, 4, flip=True, mips=True)
	tex_src = oracle.synth(700 * 500 * 3, c.FMT_U16, 3, 92)
	tex_want = oracle.engine_texture(tex_src, 700, 500, 3, flip=True, mips=True)
	pano = oracle.synth(256 * 128 * 4, c.FMT_U8, 4, 93)
	pano_want = oracle.equirect_to_cube(pano, 256, 128, 4, c.FMT_U8, 48)

	def run(k):
		if k % 3 == 0:
			got = np.empty(cube_want.size, np.uint8)
			device.engine_cube([cube_src.ctypes.data + f * 128 * 96 * 8 for f in range(6)], 128, 96, 4, c.FMT_U16, got.ctypes.data, device.CUBE_FLIP_ROWS | device.CUBE_MIPS, mem=c.MEM_HOST)
			return bool((got == cube_want).all())
		if k % 3 == 1:
			got = np.empty(tex_want.size, np.uint8)
			device.engine_texture(tex_src.ctypes.data, 700, 500, 3, c.FMT_U16, got.ctypes.data, device.CUBE_FLIP_ROWS | device.CUBE_MIPS, mem=c.MEM_HOST)
			return bool((got == tex_want).all())
		got = np.empty(pano_want.size, np.uint8)
		c.check(c.lib().cacao_img_equirect_to_cube(-1, pano.ctypes.data, 256, 128, 0, 128, 4, c.FMT_U8, got.ctypes.data, 48, 0x3F, 0, 48, c.MEM_HOST, None))
		return bool((got == pano_want).all())

	with ThreadPoolExecutor(max_workers=8) as ex:
		assert all(ex.map(run, range(48)))


_CONSUMER = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, sys.argv[4])
from cacaoengine_b200 import device
fd, mapped, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
device.init(0)
p = device.import_shareable(fd, mapped)
out = np.empty(n, np.uint8)
device.download(out, p)
device.sync()
device.free_shareable(p)
print(hashlib.sha256(out.tobytes()).hexdigest())
"""


@pytest.mark.gpu
def test_cube_handed_over_through_a_file_descriptor(built_lib, oracle):
	"""SURVEY 8 f-4, the hand-off itself: the finished cube is written into device memory that was exported as a
	POSIX file descriptor (cacao_img_malloc_shareable -- the handle VK_KHR_external_memory_fd / GL_EXT_memory_object_fd
	import), and ANOTHER PROCESS with its own CUDA context maps that descriptor and finds the contiguous RGB8 cube
	there: no host copy, no staging buffer (engine/src/vulkan/impls/VulkanCubemap.cpp:30-41 builds one by hand).
	The consumer here is a CUDA context rather than a Vulkan device -- the image has no graphics headers -- but the
	exporting side is the same call either way."""
	import hashlib
	import os
	import subprocess
	import sys

	from cacaoengine_b200 import _capi as c, device

	device.init()
	w = h = 160
	src = _faces(oracle, w, h, 4, c.FMT_U16, 77)
	flags = device.CUBE_FLIP_ROWS | device.CUBE_MIPS
	want = oracle.engine_cube(src, w, h, 4, flip=True, mips=True)
	a = DevBuf(src.nbytes)
	device.upload(a.ptr, src)
	ptr, fd, mapped = device.malloc_shareable(want.size)
	assert fd >= 0 and mapped >= want.size
	face_bytes = w * h * 8
	device.engine_cube([a.ptr + f * face_bytes for f in range(6)], w, h, 4, c.FMT_U16, ptr, flags)
	device.sync()
	root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
	r = subprocess.run([sys.executable, "-c", _CONSUMER, str(fd), str(mapped), str(want.size), root], pass_fds=(fd,), capture_output=True, text=True, timeout=300)
	assert r.returncode == 0, r.stderr[-2000:]
	assert r.stdout.strip().splitlines()[-1] == hashlib.sha256(want.tobytes()).hexdigest()
	# the producer's own mapping still reads the same bytes, and a plain pointer is refused by free_shareable
	got = np.empty_like(want)
	device.download(got, ptr)
	assert (got == want).all()
	with pytest.raises(Exception):
		device.free_shareable(a.ptr)
	os.close(fd)
	device.free_shareable(ptr)
	a.free()
