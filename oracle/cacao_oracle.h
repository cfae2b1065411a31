/*
 * cacao_oracle.h -- CPU oracle for the libcacaoimage / cubetool pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  The product (cacaoengine_b200/) never links it.
 *
 * Two tiers of parity:
 *   - integer ops (channel layout, 16->8 depth LUT): a plain-C restatement of the
 *     reference's own loops (libs/image/src/Layout.cpp, Colordepth.cpp, LUTGen.cpp),
 *     PINNED against the reference sources compiled by oracle/Makefile into
 *     oracle/_ref/ and against the golden vectors in tests/golden/.
 *   - float formats, flip and equirect->cubemap: the reference has no such code
 *     (libcacaoimage.hpp:20 "only 8 or 16"; tools/cubetool/main.cpp:31-33 has only
 *     create/extract; Flip.cpp:16 is UB for h>=2).  These functions restate the
 *     normative spec in DESIGN.md ("Spec B").  PARITY UNPINNED vs the reference.
 */
#ifndef CACAO_ORACLE_H
#define CACAO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sample formats -- numerically identical to include/cacao_img.h */
enum { ORC_U8 = 0, ORC_U16 = 1, ORC_F16 = 2, ORC_F32 = 3 };
/* modes */
enum { ORC_MODE_REF_EXACT = 0, ORC_MODE_FIXED = 1 };

/* error codes (0 = ok) */
enum { ORC_OK = 0, ORC_E_SAME_LAYOUT = -1, ORC_E_NOT_16BIT = -2, ORC_E_BAD_ARG = -3 };

/* LUTGen.cpp:15-28 -- 65536-entry sRGB OETF table, float32 math, truncation. */
void orc_build_lut(uint8_t lut[65536]);
/* the table built once on first use */
const uint8_t* orc_lut(void);

/* Layout.cpp:8-103 -- integer channel layout change, quirks included.
 * bits = 8 or 16; src_ch/dst_ch in {1,3,4}.  dst must hold pixels*dst_ch*bits/8 bytes. */
int orc_change_layout(const void* src, void* dst, uint64_t pixels, int src_ch, int dst_ch, int bits);

/* Colordepth.cpp:7-41 -- per-sample LUT lookup of little-endian u16 samples. */
int orc_depth16to8(const uint8_t* src, uint8_t* dst, uint64_t samples);

/* Spec B.1 -- generic fused convert: layout change in the source format's domain,
 * then per-sample format change.  Integer layout changes follow the reference
 * (mode REF_EXACT) or the quirk-free variant (mode FIXED). */
int orc_convert(const void* src, void* dst, uint64_t pixels, int src_ch, int dst_ch,
                int src_fmt, int dst_fmt, int mode);
/* same, OpenMP-parallel over disjoint pixel ranges (all host threads) */
int orc_convert_mt(const void* src, void* dst, uint64_t pixels, int src_ch, int dst_ch,
                   int src_fmt, int dst_fmt, int mode, int threads);

/* Spec f-1 -- vertical mirror with the semantics Flip.cpp:6-42 intends. */
int orc_flip_rows(const void* src, void* dst, uint32_t w, uint32_t h, uint32_t bytes_per_pixel);
/* f-4 follow-on (Spec B.3): one mip level, 2x2 box average; dst is max(1,w/2) x max(1,h/2) */
int orc_downsample2x(const void* src, void* dst, uint32_t w, uint32_t h, int ch, int fmt);
/* f-3: engine/src/core/Model.cpp:245-316 (texel unpack: swizzle to r,g,b,a order + widen sub-8-bit channels) */
int orc_unpack_texels(const uint8_t* src, uint8_t* dst, uint64_t texels, const char* hint);

/* Spec B.2 -- equirect (W x H, ch channels, fmt) -> rows [row_begin,row_end) of the faces in
 * face_mask (bit f = face f; order +X,-X,+Y,-Y,+Z,-Z).  dst is the FULL 6*N*N cube buffer
 * (face-major, row-major, top row first); only the selected rows are written. */
int orc_equirect_to_cube(const void* src, uint32_t W, uint32_t H, int ch, int fmt,
                         void* dst, uint32_t N, uint32_t face_mask,
                         uint32_t row_begin, uint32_t row_end, int threads);

/* Spec B.2 projection only: face pixel -> source sample position (xs, ys). */
void orc_project(uint32_t face, uint32_t i, uint32_t j, uint32_t N, uint32_t W, uint32_t H,
                 float* xs, float* ys);
/* the spec's polynomial atan2 (exposed so tests can bound it against libm) */
float orc_atan2(float y, float x);

/* Independent double-precision libm resampler (same geometry, no polynomial, no fp32
 * rounding) -- used only to bound the spec's error: f32 RGBA etc. -> double output. */
int orc_equirect_to_cube_f64(const void* src, uint32_t W, uint32_t H, int ch, int fmt,
                             double* dst, uint32_t N, uint32_t face_mask,
                             uint32_t row_begin, uint32_t row_end);

/* Spec 8(d) synthetic inputs: counter-hash generators, identical on host and device. */
uint32_t orc_hash32(uint32_t seed, uint64_t idx);
void orc_synth(void* dst, uint64_t first_sample, uint64_t samples, int fmt, int ch, uint32_t seed);

/* helpers */
uint16_t orc_f32_to_f16(float f);
float orc_f16_to_f32(uint16_t h);
uint64_t orc_fnv1a64(const void* p, uint64_t n);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
