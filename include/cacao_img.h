/*
 * cacao_img.h -- C ABI of the B200-native libcacaoimage / cubetool pixel hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI layer around this path: the
 * interface is the C++ header libs/image/include/libcacaoimage.hpp (free functions taking
 * `const Image&`, returning `Image`, throwing std::runtime_error).  include/libcacaoimage.hpp
 * in this repo keeps that header's API and forwards every pixel loop to the entry points
 * below; INTEGRATION.md shows the change a reference maintainer would make.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types; nothing throws across this boundary.
 *   - every function returns 0 (CACAO_OK) or a negative cacao_status; the text for the calling
 *     thread's last failure is cacao_img_last_error().
 *   - `device` is a CUDA ordinal, or -1 for the calling thread's current CUDA device.
 *   - `mem` says where src/dst live.  CACAO_MEM_DEVICE: both are device pointers on `device`;
 *     the work is enqueued on `stream` (a cudaStream_t, NULL = legacy default stream) and the call
 *     returns without synchronising.  CACAO_MEM_HOST: both are host pointers; the call uploads,
 *     converts and reads back through an internal chunked three-stream pipeline and returns when
 *     dst is complete (`stream` is ignored).
 *   - images are tightly packed, interleaved, top row first, pitch = w*channels*bytes(sample)
 *     (libs/image/src/Layout.cpp:18-19); 16-bit samples are native little-endian
 *     (libs/image/src/Colordepth.cpp:32).
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails
 *     with CACAO_E_NO_DEVICE.
 */
#ifndef CACAO_IMG_H
#define CACAO_IMG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CACAO_IMG_ABI_VERSION 1

typedef enum cacao_sample_fmt {
	CACAO_FMT_U8 = 0,  /* Image::bitsPerChannel == 8  (libcacaoimage.hpp:20) */
	CACAO_FMT_U16 = 1, /* Image::bitsPerChannel == 16 */
	CACAO_FMT_F16 = 2, /* extension: IEEE binary16 */
	CACAO_FMT_F32 = 3  /* extension: IEEE binary32 */
} cacao_sample_fmt;

typedef enum cacao_mode {
	CACAO_MODE_REF_EXACT = 0, /* bit-exact with the reference, quirks included ("Quirks") */
	CACAO_MODE_FIXED = 1      /* Gray->RGBA replicates per pixel; 16-bit gray clamps to 65535 */
} cacao_mode;

typedef enum cacao_mem { CACAO_MEM_HOST = 0, CACAO_MEM_DEVICE = 1 } cacao_mem;

typedef enum cacao_status {
	CACAO_OK = 0,
	CACAO_E_SAME_LAYOUT = -1, /* Layout.cpp:9   "Cannot change an image's channel layout to its current layout!" */
	CACAO_E_NOT_16BIT = -2,   /* Colordepth.cpp:8 "Source image for 16 to 8-bit color conversion does not have a 16-bit color depth!" */
	CACAO_E_BAD_ARG = -3,
	CACAO_E_CUDA = -4,
	CACAO_E_NO_DEVICE = -5,
	CACAO_E_ZERO_DIM = -6,  /* Flip.cpp:8  "Cannot flip an image with zeroed dimensions!" */
	CACAO_E_BAD_DEPTH = -7, /* Flip.cpp:9  "Invalid color depth; only 8 and 16 are allowed." */
	CACAO_E_EMPTY = -8,     /* Flip.cpp:10 "Cannot flip an image with a zero-sized data buffer!" */
	CACAO_E_NO_MEMORY = -9
} cacao_status;

/* ---- library / device ------------------------------------------------------------------ */
int cacao_img_abi_version(void);
const char* cacao_img_last_error(void);          /* thread-local, never NULL */
const char* cacao_img_status_string(int status); /* the reference's exception text where one exists */
int cacao_img_device_count(void);                /* >= 0, or a negative status */
int cacao_img_init(int device);                  /* idempotent; builds the device context (LUT upload, streams) */
int cacao_img_shutdown(void);                    /* releases every context */
uint64_t cacao_img_launch_count(void);           /* kernels launched by this library so far (all threads) */

/* ---- device-buffer / upload / readback layer (new; no reference counterpart) -------------- */
int cacao_img_malloc(int device, void** dev_ptr, uint64_t bytes);
int cacao_img_free(int device, void* dev_ptr);
/* Device memory another API or another process takes over WITHOUT a copy (the hand-off of a finished
 * cube / texture to the graphics backend, which today memcpy's six faces into a staging buffer --
 * engine/src/vulkan/impls/VulkanCubemap.cpp:30-41).  The allocation is a CUDA virtual-memory allocation exported as
 * a POSIX file descriptor: the handle type VK_KHR_external_memory_fd (VkImportMemoryFdInfoKHR, OPAQUE_FD) and
 * GL_EXT_memory_object_fd import, and what cacao_img_import_shareable maps in another process or context.  Use the
 * pointer as `dst` of any CACAO_MEM_DEVICE call (cacao_img_engine_cube writes the contiguous RGB8 cube straight
 * into it).  `*mapped_bytes` is the size rounded up to the allocation granularity -- the size the importer must
 * name.  The descriptor belongs to the caller (close() it once it has been handed over; the memory lives until
 * every mapping is freed).  The other direction needs no call here: memory a Vulkan / GL allocation exported and
 * the application imported (cudaImportExternalMemory) is an ordinary device pointer to this library. */
int cacao_img_malloc_shareable(int device, void** dev_ptr, uint64_t bytes, int* fd, uint64_t* mapped_bytes);
int cacao_img_import_shareable(int device, int fd, uint64_t mapped_bytes, void** dev_ptr);
int cacao_img_free_shareable(int device, void* dev_ptr); /* for pointers of either call; synchronises the device */
/* The device-side gather of a sharded job ("optional device-resident assembly on GPU 0 ... over NVLink"):
 * a rank exports a buffer (64 opaque bytes + the pointer's offset inside its allocation), another PROCESS imports it
 * and gets a pointer to the same memory, valid as `dst` (or `src`) of every CACAO_MEM_DEVICE call on the importer's
 * own GPU -- the kernel's stores cross NVLink and land in place, so the resampling launches of all ranks assemble
 * one cube on one GPU with no gather step.  cacao_img_ipc_close(ptr, offset) unmaps it (synchronises the device).
 * Inside one process that drives several devices, cacao_img_enable_peer_access(device, peer) does the same for
 * plain pointers.  The memory must come from cacao_img_malloc / cudaMalloc (not cacao_img_malloc_shareable). */
int cacao_img_ipc_export(int device, void* dev_ptr, unsigned char* handle64, uint64_t* offset);
int cacao_img_ipc_import(int device, const unsigned char* handle64, uint64_t offset, void** dev_ptr);
int cacao_img_ipc_close(int device, void* dev_ptr, uint64_t offset);
int cacao_img_enable_peer_access(int device, int peer);
int cacao_img_host_alloc(void** host_ptr, uint64_t bytes); /* pinned, portable across devices */
int cacao_img_host_free(void* host_ptr);
/* Pin memory the caller already owns (a std::vector's storage, a shared-memory segment several processes map,
 * an mmap'ed file) so the host-memory calls DMA it directly instead of staging it: cudaHostRegister, portable
 * across devices.  The range must stay mapped until cacao_img_host_unregister(host_ptr). */
int cacao_img_host_register(void* host_ptr, uint64_t bytes);
int cacao_img_host_unregister(void* host_ptr);
/* Copies are enqueued on `stream` and asynchronous for pinned host memory.  Pageable host memory
 * (>= 1 MiB) is staged through the library's pinned ring by its host copy threads and the call
 * returns when the bytes have landed. */
int cacao_img_upload(int device, void* dst_dev, const void* src_host, uint64_t bytes, void* stream);
int cacao_img_download(int device, void* dst_host, const void* src_dev, uint64_t bytes, void* stream);
int cacao_img_sync(int device, void* stream); /* stream == NULL: whole device */

/* ---- the hot path ------------------------------------------------------------------------ */

/*
 * Fused channel-layout + sample-format conversion of `pixels` pixels.
 *   replaces  ChangeChannelLayout   (libs/image/src/Layout.cpp:8-103)      src_fmt == dst_fmt in {U8,U16}
 *             Convert16To8BitColor  (libs/image/src/Colordepth.cpp:7-41)   U16 -> U8, src_ch == dst_ch
 *             both in one pass      (engine/src/core/Cubemap.cpp:28-31)    e.g. RGBA16 -> RGB8
 *                                   U16 -> U8 with a layout change applies the depth table FIRST, then the
 *                                   8-bit layout rules -- the engine's order, so the fused call equals
 *                                   ChangeChannelLayout(Convert16To8BitColor(img), layout) for every layout
 *                                   pair (an added alpha is 255; gray is computed from table outputs)
 *   extends   to F16/F32 per DESIGN.md "Spec B.1" (no reference counterpart).
 * Channel counts are Image::Layout values: 1 (Grayscale), 3 (RGB), 4 (RGBA).
 * src_ch == dst_ch && src_fmt == dst_fmt -> CACAO_E_SAME_LAYOUT (Layout.cpp:9).
 */
int cacao_img_convert(int device, const void* src, void* dst, uint64_t pixels, int src_ch, int dst_ch,
					  int src_fmt, int dst_fmt, int mode, int mem, void* stream);

/*
 * The same conversion over `count` equal-size images stored back to back (image k starts at
 * byte k*pixels_per_image*src_bytes_per_pixel).  One launch for the whole batch; each image is its
 * own domain for the Gray->RGBA pixel-0 quirk.
 */
int cacao_img_convert_batch(int device, const void* src, void* dst, uint64_t pixels_per_image, uint64_t count,
							int src_ch, int dst_ch, int src_fmt, int dst_fmt, int mode, int mem, void* stream);

/* The same conversion with the row mirror of cacao_img_flip_rows folded into the store addressing:
 * `count` images of w x h pixels stored back to back, each written bottom row first -- e.g. the
 * engine's load-time chain 16 -> 8 bit, -> RGB (engine/src/core/Cubemap.cpp:28-31) and the GL
 * backend's Flip before upload (engine/src/opengl/impls/OpenGLCubemap.cpp:25, OpenGLTex2D.cpp:16) as ONE
 * pass instead of three.  Widths that a vector kernel's unit divides (8 ... 64 pixels for the sector
 * kernel's shapes, 32 x 1 ... 16 pixels for the tile kernel) run at streaming speed; other widths take
 * the byte-wise kernel.  Same status codes as cacao_img_convert. */
int cacao_img_convert_flip(int device, const void* src, void* dst, uint32_t w, uint32_t h, uint64_t count,
						   int src_ch, int dst_ch, int src_fmt, int dst_fmt, int mode, int mem, void* stream);

/*
 * The engine's cubemap load chain as ONE call.  What Cacao::Cubemap::Cubemap does per face in two passes and two
 * allocations -- Convert16To8BitColor when the face is 16-bit, then ChangeChannelLayout to RGB when it is not
 * RGB already (engine/src/core/Cubemap.cpp:28-31) -- what the GL backend adds before upload (Flip,
 * engine/src/opengl/impls/OpenGLCubemap.cpp:25) and what the Vulkan backend adds after it (the six faces copied
 * into one contiguous buffer, engine/src/vulkan/impls/VulkanCubemap.cpp:30-41), plus, on request, the mip chain
 * the GL backend asks the driver for (glGenerateMipmap, OpenGLTex2D.cpp:51):
 *   faces[f]  the six faces +X,-X,+Y,-Y,+Z,-Z, each w x h pixels of `ch` channels (1, 3, 4) of `fmt`
 *             (CACAO_FMT_U8 or CACAO_FMT_U16); separate allocations are fine
 *   dst       ONE buffer: level 0 = six RGB8 faces back to back (6*w*h*3 bytes), then -- with
 *             CACAO_CUBE_MIPS -- level 1 (six faces of max(1,w/2) x max(1,h/2)), level 2, ... down to 1 x 1,
 *             every level face-major; cacao_img_engine_cube_bytes() says how large it is
 *   flags     CACAO_CUBE_FLIP_ROWS: every face bottom row first; CACAO_CUBE_MIPS: append the mip chain
 *             (2x2 box average per level, DESIGN.md "Spec B.3", of the faces as stored)
 * Values are the reference's, bit for bit (REF_EXACT rules, the Gray -> RGB replicate included).  Host memory:
 * one pipelined pass -- face k+1 goes up while face k is converted and face k-1 comes back; pinned buffers are
 * DMA'd directly, pageable ones (std::vector) are staged by the host copy pool.  Device memory: one launch when
 * the faces are contiguous, six otherwise, plus one per mip level.
 */
#define CACAO_CUBE_FLIP_ROWS 1u
#define CACAO_CUBE_MIPS 2u
uint64_t cacao_img_engine_cube_bytes(uint32_t w, uint32_t h, uint32_t flags);
int cacao_img_engine_cube(int device, const void* const* faces, uint32_t w, uint32_t h, int ch, int fmt, void* dst,
						  uint32_t flags, int mem, void* stream);

/*
 * The same for a 2-D texture.  Cacao::Tex2D::Tex2D converts a 16-bit image with Convert16To8BitColor and keeps its
 * layout (engine/src/core/Tex2D.cpp:18); the GL backend then Flips it and asks the driver for its mip levels
 * (engine/src/opengl/impls/OpenGLTex2D.cpp:16,51).  One call: src (w x h pixels, `ch` channels of U8 or U16) ->
 * dst = level 0 (w x h pixels of `ch` 8-bit channels; CACAO_CUBE_FLIP_ROWS: bottom row first) followed, with
 * CACAO_CUBE_MIPS, by every mip level down to 1 x 1 (cacao_img_engine_texture_bytes sizes dst; 0 for a bad
 * channel count).  Host memory: the image flows through in bands of rows, upload / conversion / readback
 * overlapped; the mip levels are computed on the device from level 0 and follow it.
 */
uint64_t cacao_img_engine_texture_bytes(uint32_t w, uint32_t h, int ch, uint32_t flags);
int cacao_img_engine_texture(int device, const void* src, uint32_t w, uint32_t h, int ch, int fmt, void* dst,
							 uint32_t flags, int mem, void* stream);

/*
 * dst[k] = (uint16_t)(src[k] << shift) over `samples` 16-bit samples, shift in [0,15].
 *   replaces the 12 -> 16 bit widening loop of the JPEG decoder, libs/image/src/JPEG.cpp:79-82
 *   (`imgDataSixteen[i] = uint16_t(twelveBit[i]) << 4`), one of the per-pixel loops next to the
 *   hot path.  In place (src == dst) is allowed; partly overlapping buffers are not.
 *   (Every other call needs disjoint src and dst and returns CACAO_E_BAD_ARG otherwise.)
 */
int cacao_img_shift_left_u16(int device, const void* src, void* dst, uint64_t samples, uint32_t shift, int mem,
							 void* stream);

/*
 * One mip level: dst (max(1,w/2) x max(1,h/2) pixels) = 2x2 box average of src (w x h), per channel on
 * stored values -- integers (a+b+c+d+2)>>2, floats ((a+b)+(c+d))*0.25f in fp32 (DESIGN.md "Spec B.3").
 * No reference counterpart: the engine leaves mip generation to the graphics API (glGenerateMipmap,
 * engine/src/opengl/impls/OpenGLTex2D.cpp:51); with the pixels already on the GPU the chain can be
 * produced where they are.  Call it repeatedly for a chain.
 */
int cacao_img_downsample2x(int device, const void* src, void* dst, uint32_t w, uint32_t h, int ch, int fmt,
						   int mem, void* stream);

/*
 * Texel unpack for textures embedded in model files: `texels` source texels of four bytes each, laid
 * out as `hint` says -- assimp's aiTexture::achFormatHint, four channel letters then four bit counts,
 * e.g. "rgba8888", "argb8888", "bgra5650" -- become tightly packed 8-bit pixels in r,g,b,a order;
 * channels with a zero count are dropped (none -> RGBA, alpha -> RGB, three -> Grayscale), counts
 * below 8 are widened by bit replication.
 *   replaces the swizzle / expand loop of engine/src/core/Model.cpp:245-316, with
 *   its validation: a hint with two or four zero counts, or one zero count that is not alpha, fails
 *   with CACAO_E_BAD_ARG and the reference's message in cacao_img_last_error().
 * *out_channels (may be NULL) receives 4, 3 or 1; dst needs texels * that many bytes.
 */
int cacao_img_unpack_texels(int device, const void* src, void* dst, uint64_t texels, const char* hint,
							int* out_channels, int mem, void* stream);

/*
 * Vertical mirror, out.row[y] = src.row[h-1-y], src untouched: the semantics
 * libs/image/src/Flip.cpp:6-42 intends (the reference itself overruns its buffer for h >= 2).
 */
int cacao_img_flip_rows(int device, const void* src, void* dst, uint32_t w, uint32_t h,
						uint32_t bytes_per_pixel, int mem, void* stream);

/*
 * Equirectangular -> cubemap, direction->UV projection + bilinear (DESIGN.md "Spec B.2"; new
 * `ce-cubetool project` subcommand -- the reference's cubetool has only create/extract,
 * tools/cubetool/main.cpp:31-33).
 *   src       rows [src_row0, src_row0+src_rows) of a W x H equirect image, `ch` channels of `fmt`
 *             (pass src_row0 = 0, src_rows = H for the whole image; a shard passes the window
 *             returned by cacao_img_equirect_src_window)
 *   dst       base of the FULL cube buffer: 6 faces x N x N pixels, face-major in the order
 *             +X,-X,+Y,-Y,+Z,-Z (libs/formats/include/libcacaoformats.hpp:213), rows top first
 *   face_mask bit f set = produce face f;  rows [row_begin,row_end) of each selected face are written
 * With CACAO_MEM_HOST the call is the pipelined pass described at
 * cacao_img_equirect_to_cube_units_host: only the source rows that are read go up, only the written
 * rows come back, and the two directions overlap.
 */
int cacao_img_equirect_to_cube(int device, const void* src, uint32_t W, uint32_t H, uint32_t src_row0,
							   uint32_t src_rows, int ch, int fmt, void* dst, uint32_t N, uint32_t face_mask,
							   uint32_t row_begin, uint32_t row_end, int mem, void* stream);

/* One (face, row band) work unit of a cubemap: rows [row_begin,row_end) of face `face` (0..5). */
typedef struct cacao_cube_unit {
	uint32_t face;
	uint32_t row_begin;
	uint32_t row_end;
} cacao_cube_unit;

/* The same resampling for an explicit list of work units -- what one GPU of a sharded job owns
 * (cacaoengine_b200/sharding.py deals units out) -- normally in ONE kernel launch (units are cut to a
 * common height; up to 32 pieces per launch) instead of one per unit.  src/dst/window arguments as above; device memory only. */
int cacao_img_equirect_to_cube_units(int device, const void* src, uint32_t W, uint32_t H, uint32_t src_row0,
									 uint32_t src_rows, int ch, int fmt, void* dst, uint32_t N,
									 const cacao_cube_unit* units, uint32_t n_units, void* stream);

/* The host-memory form of the unit list: src = host rows [src_row0, src_row0+src_rows) of the
 * equirect image, faces[f] = host base of face f's N x N pixels (may be NULL for faces no unit
 * names; the six faces need not be contiguous -- libcacaoimage::EquirectToCubemap hands over six
 * std::vectors).  One pipelined pass: the source goes up in row chunks, a unit is resampled as soon
 * as the last row of its window is on the device, and its rows travel back while later chunks still
 * go up, so the call costs about max(upload, readback) rather than their sum.  Pinned and pageable
 * buffers both work (pageable ones are staged through the library's pinned rings).  Returns when
 * every named row is in place. */
int cacao_img_equirect_to_cube_units_host(int device, const void* src, uint32_t W, uint32_t H, uint32_t src_row0,
											 uint32_t src_rows, int ch, int fmt, void* const* faces, uint32_t N,
											 const cacao_cube_unit* units, uint32_t n_units);

/* Resampling AND sample-format / layout conversion in one pass (DESIGN.md "Spec B.4"): the filtered
 * fp32 texel is converted to dst_fmt exactly as cacao_img_convert converts an F32 sample (Spec B.1) and
 * an RGB source may gain an opaque alpha (dst_ch = 4) on its way out -- e.g. an RGB32F .hdr panorama
 * straight to an RGBA16F cube, without the F32 cube ever touching memory.  Defined for FLOAT sources
 * (F16, F32), dst_ch == ch, 3 -> 4 or 4 -> 3 (alpha dropped); dst_ch == ch && dst_fmt == fmt is the plain resampling of any
 * format; everything else returns CACAO_E_BAD_ARG (resample, then convert).  dst holds 6 faces of
 * N x N x dst_ch samples of dst_fmt.  units == NULL && n_units == 0: the whole cube.
 * flags: CACAO_EQUIRECT_FLIP_ROWS stores every face bottom row first -- the vertical mirror the GL
 * backend applies before upload (engine/src/opengl/impls/OpenGLCubemap.cpp:25), folded into the store
 * addressing for every format; units keep naming rows of the unflipped face (row j lands at N-1-j). */
#define CACAO_EQUIRECT_FLIP_ROWS 1u
int cacao_img_equirect_to_cube_units_cvt(int device, const void* src, uint32_t W, uint32_t H, uint32_t src_row0,
										 uint32_t src_rows, int ch, int fmt, void* dst, uint32_t N, int dst_ch,
										 int dst_fmt, uint32_t flags, const cacao_cube_unit* units, uint32_t n_units,
										 void* stream);
/* Host-memory form (faces[f] = host base of face f in the DESTINATION format), pipelined like
 * cacao_img_equirect_to_cube_units_host. */
int cacao_img_equirect_to_cube_units_host_cvt(int device, const void* src, uint32_t W, uint32_t H,
											  uint32_t src_row0, uint32_t src_rows, int ch, int fmt,
											  void* const* faces, uint32_t N, int dst_ch, int dst_fmt, uint32_t flags,
											  const cacao_cube_unit* units, uint32_t n_units);

/* Host-side planning: the inclusive-exclusive window of source rows that the work unit
 * (face_mask, [row_begin,row_end)) reads, so a shard can upload just that band. */
int cacao_img_equirect_src_window(uint32_t W, uint32_t H, uint32_t N, uint32_t face_mask, uint32_t row_begin,
								  uint32_t row_end, uint32_t* src_row0, uint32_t* src_rows);

/* Host-side planning: how the launcher folds a unit list before it goes to the device.  A band that is
 * listed together with its mirror band on the same face (rows j and N-1-j: a whole face, the two
 * units of a mirror pair dealt by sharding.py, ...) becomes ONE unit of the upper half whose face has
 * CACAO_UNIT_MIRROR set -- the kernel then computes one projection per four pixels instead of per
 * two.  Rows without a listed mirror stay plain units.  Writes at most `cap` units to `out`
 * and the full count to *n_out (call with cap = 0 to size the buffer). */
#define CACAO_UNIT_MIRROR 8u
int cacao_img_equirect_fold_units(uint32_t N, const cacao_cube_unit* units, uint32_t n_units, cacao_cube_unit* out,
								  uint32_t cap, uint32_t* n_out);

/* ---- support for measurement and verification (device-side, deterministic) ---------------- */

/* Fill dst_dev with the counter-hash synthetic samples [first_sample, first_sample+samples)
 * (DESIGN.md "Synthetic inputs"; identical to oracle/cacao_oracle.c orc_synth on the host). */
int cacao_img_synth(int device, void* dst_dev, uint64_t first_sample, uint64_t samples, int fmt, int ch,
					uint32_t seed, void* stream);

/* Position-sensitive 64-bit checksum of a device buffer (sum over 32-bit words of
 * hash(word, index)); bytes must be a multiple of 4.  Synchronises `stream`. */
int cacao_img_checksum(int device, const void* src_dev, uint64_t bytes, uint64_t* out, void* stream);

/* The 65536-entry 16->8 table in use (libs/image/LUTGen.cpp:20-28), copied to out. */
int cacao_img_get_lut(uint8_t* out65536);

#ifdef __cplusplus
}
#endif
#endif /* CACAO_IMG_H */
