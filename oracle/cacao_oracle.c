/*
 * cacao_oracle.c -- CPU oracle (plain C) for the libcacaoimage / cubetool pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see cacao_oracle.h.  Compile with
 *   gcc -O2 -std=c11 -ffp-contract=off -fopenmp -fPIC -shared
 * (-ffp-contract=off is REQUIRED: the float spec fixes where fused multiply-adds
 * happen; every fused op below is an explicit fmaf()).
 *
 * Citations are relative to /root/reference/.
 */
#include "cacao_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define ORC_X86 1
#endif

/* ------------------------------------------------------------------------------------------
 * a4: the 16->8 LUT.  Restates libs/image/LUTGen.cpp:8-28: everything in float32,
 * host powf, clamp to [0,255], then TRUNCATE (so lut[65535] == 254).
 * ---------------------------------------------------------------------------------------- */
static float orc_to_srgb(float lin) {
	/* LUTGen.cpp:15-17; note SRGB_EXPONENT is the float expression 1.0f / 2.4f */
	const float expo = 1.0f / 2.4f;
	return (lin <= 0.0031308f) ? lin * 12.92f : 1.055f * powf(lin, expo) - 0.055f;
}

void orc_build_lut(uint8_t lut[65536]) {
	for(uint32_t i = 0; i < 65536u; ++i) {
		float normalized = (float)i / 65535.0f;  /* LUTGen.cpp:23 */
		float s = orc_to_srgb(normalized) * 255.0f;
		if(s < 0.0f) s = 0.0f;                   /* std::clamp, LUTGen.cpp:25 */
		if(s > 255.0f) s = 255.0f;
		lut[i] = (uint8_t)s;                     /* truncation */
	}
}

static uint8_t g_lut[65536];
static int g_lut_ready = 0;

const uint8_t* orc_lut(void) {
	if(!g_lut_ready) {
#pragma omp critical(orc_lut_init)
		{
			if(!g_lut_ready) {
				orc_build_lut(g_lut);
				g_lut_ready = 1;
			}
		}
	}
	return g_lut;
}

/* ------------------------------------------------------------------------------------------
 * a2: ChangeChannelLayout, restated per case from libs/image/src/Layout.cpp:28-99.
 *
 * What the reference's do/while does, case by case (verified against the
 * compiled reference by tests/test_oracle_vs_ref.py):
 *   ->Gray      : trunc(0.2126*r) + trunc(0.7152*g) + trunc(0.0722*b), double products, each
 *                 truncated to the sample type first, sum clamped to 255 -- for 16-bit too
 *                 (Layout.cpp:33-41, :69-77).
 *   Gray->RGB   : (g,g,g)                                         (Layout.cpp:47-54,60)
 *   Gray->RGBA  : the source pointer never advances (step = clamp(oldCh-1,0,1) = 0 and the
 *                 ++srcp at :60 sits in the non-RGBA branch) => every pixel = (g0,g0,g0,MAX)
 *   RGB->RGBA   : (r,g,b,MAX)                                     (Layout.cpp:49-53,57-58)
 *   RGBA->RGB   : (r,g,b)                                         (Layout.cpp:49-52,60)
 * ---------------------------------------------------------------------------------------- */
static uint32_t gray_ref_u8(uint32_t r, uint32_t g, uint32_t b) {
	uint8_t tr = (uint8_t)(0.2126 * (double)r);
	uint8_t tg = (uint8_t)(0.7152 * (double)g);
	uint8_t tb = (uint8_t)(0.0722 * (double)b);
	uint32_t sum = (uint32_t)tr + tg + tb;
	return sum > 255u ? 255u : sum;
}

static uint32_t gray_ref_u16(uint32_t r, uint32_t g, uint32_t b, int mode) {
	uint16_t tr = (uint16_t)(0.2126 * (double)r);
	uint16_t tg = (uint16_t)(0.7152 * (double)g);
	uint16_t tb = (uint16_t)(0.0722 * (double)b);
	uint32_t sum = (uint32_t)tr + tg + tb;
	uint32_t cap = (mode == ORC_MODE_REF_EXACT) ? 255u : 65535u; /* Layout.cpp:75 clamps to 255 */
	return sum > cap ? cap : sum;
}

/* One integer pixel, already widened to u32 samples, through the layout rules. `first` is
 * pixel 0 of the image (needed for the Gray->RGBA quirk). */
static void layout_int_px(const uint32_t* in, const uint32_t* first, uint32_t* out, int src_ch,
						  int dst_ch, int bits, int mode) {
	const uint32_t maxv = (bits == 8) ? 255u : 65535u;
	if(dst_ch == 1) {
		out[0] = (bits == 8) ? gray_ref_u8(in[0], in[1], in[2]) : gray_ref_u16(in[0], in[1], in[2], mode);
		return;
	}
	if(src_ch == 1) {
		uint32_t g = (dst_ch == 4 && mode == ORC_MODE_REF_EXACT) ? first[0] : in[0];
		out[0] = out[1] = out[2] = g;
		if(dst_ch == 4) out[3] = maxv;
		return;
	}
	out[0] = in[0];
	out[1] = in[1];
	out[2] = in[2];
	if(dst_ch == 4) out[3] = (src_ch == 4) ? in[3] : maxv;
}

int orc_change_layout(const void* src, void* dst, uint64_t pixels, int src_ch, int dst_ch, int bits) {
	if(src_ch == dst_ch) return ORC_E_SAME_LAYOUT; /* Layout.cpp:9 */
	if((bits != 8 && bits != 16) || !(src_ch == 1 || src_ch == 3 || src_ch == 4) ||
	   !(dst_ch == 1 || dst_ch == 3 || dst_ch == 4))
		return ORC_E_BAD_ARG;
	uint32_t in[4], out[4], first[4] = {0, 0, 0, 0};
	const uint8_t* s8 = (const uint8_t*)src;
	const uint16_t* s16 = (const uint16_t*)src;
	uint8_t* d8 = (uint8_t*)dst;
	uint16_t* d16 = (uint16_t*)dst;
	for(uint64_t p = 0; p < pixels; ++p) {
		for(int c = 0; c < src_ch; ++c) in[c] = (bits == 8) ? s8[p * src_ch + c] : s16[p * src_ch + c];
		if(p == 0) memcpy(first, in, sizeof first);
		layout_int_px(in, first, out, src_ch, dst_ch, bits, ORC_MODE_REF_EXACT);
		for(int c = 0; c < dst_ch; ++c) {
			if(bits == 8)
				d8[p * dst_ch + c] = (uint8_t)out[c];
			else
				d16[p * dst_ch + c] = (uint16_t)out[c];
		}
	}
	return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * a3: Convert16To8BitColor, libs/image/src/Colordepth.cpp:26-37.
 * ---------------------------------------------------------------------------------------- */
int orc_depth16to8(const uint8_t* src, uint8_t* dst, uint64_t samples) {
	const uint8_t* lut = orc_lut();
	for(uint64_t k = 0; k < samples; ++k) {
		uint16_t value = (uint16_t)((uint16_t)src[2 * k] | ((uint16_t)src[2 * k + 1] << 8)); /* :32 */
		dst[k] = lut[value];                                                                 /* :35 */
	}
	return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * half <-> float, round-to-nearest-even, written out bit by bit (no compiler _Float16).
 * ---------------------------------------------------------------------------------------- */
uint16_t orc_f32_to_f16(float f) {
	uint32_t x;
	memcpy(&x, &f, 4);
	uint32_t sign = (x >> 16) & 0x8000u;
	uint32_t absx = x & 0x7FFFFFFFu;
	if(absx > 0x7F800000u) return (uint16_t)(sign | 0x7E00u | ((absx >> 13) & 0x3FFu)); /* NaN (quiet) */
	if(absx >= 0x477FF000u) {
		/* >= 65520 rounds to inf (65504 + half ulp) */
		return (uint16_t)(sign | 0x7C00u);
	}
	if(absx < 0x33000001u) return (uint16_t)sign; /* <= 2^-25 rounds to zero (tie -> even = 0) */
	int32_t e = (int32_t)(absx >> 23) - 127;
	uint32_t m = (absx & 0x7FFFFFu) | 0x800000u; /* 24-bit significand */
	uint32_t shift;
	uint32_t hexp;
	if(e < -14) { /* subnormal half */
		shift = (uint32_t)(13 + (-14 - e));
		hexp = 0;
	} else {
		shift = 13;
		hexp = (uint32_t)(e + 15);
	}
	uint32_t q = m >> shift;
	uint32_t rem = m & ((1u << shift) - 1u);
	uint32_t half = 1u << (shift - 1);
	if(rem > half || (rem == half && (q & 1u))) q += 1;
	uint32_t h;
	if(hexp == 0)
		h = q; /* may carry into exponent 1: correct */
	else
		h = ((hexp - 1) << 10) + q; /* q includes the hidden bit (0x400) */
	return (uint16_t)(sign | h);
}

float orc_f16_to_f32(uint16_t h) {
	uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
	uint32_t e = (h >> 10) & 0x1Fu;
	uint32_t m = h & 0x3FFu;
	uint32_t x;
	if(e == 0) {
		if(m == 0) {
			x = sign;
		} else {
			int sh = 0;
			while(!(m & 0x400u)) {
				m <<= 1;
				++sh;
			}
			m &= 0x3FFu;
			x = sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13);
		}
	} else if(e == 31) {
		x = sign | 0x7F800000u | (m << 13);
	} else {
		x = sign | ((e + 112u) << 23) | (m << 13);
	}
	float f;
	memcpy(&f, &x, 4);
	return f;
}

/* ------------------------------------------------------------------------------------------
 * a6 (NEW, Spec B.1): fused layout + sample-format conversion.
 * ---------------------------------------------------------------------------------------- */
static int fmt_bytes(int fmt) {
	return (fmt == ORC_U8) ? 1 : (fmt == ORC_F32) ? 4 : 2;
}

static float load_as_float(const uint8_t* p, int fmt) {
	switch(fmt) {
		case ORC_F32: {
			float f;
			memcpy(&f, p, 4);
			return f;
		}
		case ORC_F16: {
			uint16_t h;
			memcpy(&h, p, 2);
			return orc_f16_to_f32(h);
		}
		case ORC_U16: {
			uint16_t v;
			memcpy(&v, p, 2);
			return (float)v;
		}
		default: return (float)p[0];
	}
}

static uint32_t unorm_from_float(float x, uint32_t maxv) {
	/* Spec B.1: v = (x != x) ? 0 : uintN(min(max(x,0),1)*M + 0.5f) */
	if(x != x) return 0;
	float c = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
	return (uint32_t)(c * (float)maxv + 0.5f);
}

/* Pure sample-format change unorm -> float (same channel count) is a flat sample stream; these
 * typed loops compute exactly what the generic path below computes per sample --
 * (float)v / (float)M in IEEE fp32, then RNE to binary16 where asked -- without its per-sample
 * dispatch, so that the CPU baseline is a fair scalar loop in the style of the reference's own
 * (Layout.cpp:31-62) rather than an interpreter.  The binary16 rounding uses the CPU's F16C unit
 * when it has one (same IEEE round-to-nearest-even as orc_f32_to_f16; tests compare the two). */
#ifdef ORC_X86
__attribute__((target("f16c"))) static void u8_to_f16_hw(const uint8_t* s, uint16_t* d, uint64_t n) {
	for(uint64_t k = 0; k < n; ++k) d[k] = _cvtss_sh((float)s[k] / 255.0f, _MM_FROUND_TO_NEAREST_INT);
}
__attribute__((target("f16c"))) static void u16_to_f16_hw(const uint16_t* s, uint16_t* d, uint64_t n) {
	for(uint64_t k = 0; k < n; ++k) d[k] = _cvtss_sh((float)s[k] / 65535.0f, _MM_FROUND_TO_NEAREST_INT);
}
#endif
static int unorm_to_float_flat(const uint8_t* src, uint8_t* dst, uint64_t n, int src_fmt, int dst_fmt) {
	if(dst_fmt == ORC_F32) {
		float* d = (float*)dst;
		if(src_fmt == ORC_U8)
			for(uint64_t k = 0; k < n; ++k) d[k] = (float)src[k] / 255.0f;
		else
			for(uint64_t k = 0; k < n; ++k) d[k] = (float)((const uint16_t*)src)[k] / 65535.0f;
		return 1;
	}
	uint16_t* d = (uint16_t*)dst;
#ifdef ORC_X86
	if(__builtin_cpu_supports("f16c")) {
		if(src_fmt == ORC_U8)
			u8_to_f16_hw(src, d, n);
		else
			u16_to_f16_hw((const uint16_t*)src, d, n);
		return 1;
	}
#endif
	if(src_fmt == ORC_U8)
		for(uint64_t k = 0; k < n; ++k) d[k] = orc_f32_to_f16((float)src[k] / 255.0f);
	else
		for(uint64_t k = 0; k < n; ++k) d[k] = orc_f32_to_f16((float)((const uint16_t*)src)[k] / 65535.0f);
	return 1;
}

static void convert_range(const uint8_t* src, uint8_t* dst, uint64_t p0, uint64_t p1, const uint8_t* px0,
						  int src_ch, int dst_ch, int src_fmt, int dst_fmt, int mode) {
	const int sb = fmt_bytes(src_fmt), db = fmt_bytes(dst_fmt);
	const int src_int = (src_fmt == ORC_U8 || src_fmt == ORC_U16);
	const int bits = (src_fmt == ORC_U8) ? 8 : 16;
	const uint8_t* lut = orc_lut();
	if(src_int && src_ch == dst_ch && (dst_fmt == ORC_F16 || dst_fmt == ORC_F32)) {
		unorm_to_float_flat(src + p0 * (uint64_t)(src_ch * sb), dst + p0 * (uint64_t)(dst_ch * db), (p1 - p0) * (uint64_t)src_ch, src_fmt, dst_fmt);
		return;
	}
	uint32_t first[4] = {0, 0, 0, 0};
	if(src_int)
		for(int c = 0; c < src_ch; ++c) first[c] = (uint32_t)load_as_float(px0 + (size_t)c * sb, src_fmt);

	for(uint64_t p = p0; p < p1; ++p) {
		const uint8_t* sp = src + p * (uint64_t)(src_ch * sb);
		uint8_t* dp = dst + p * (uint64_t)(dst_ch * db);
		if(src_int) {
			/* step 1: layout change in the integer domain (reference rules) */
			uint32_t in[4] = {0, 0, 0, 0}, mid[4];
			for(int c = 0; c < src_ch; ++c) in[c] = (uint32_t)load_as_float(sp + (size_t)c * sb, src_fmt);
			if(src_fmt == ORC_U16 && dst_fmt == ORC_U8 && src_ch != dst_ch) {
				/* 16 -> 8 bit WITH a layout change: depth first, then the 8-bit layout rules -- the order of
				 * the engine's composition Convert16To8BitColor, ChangeChannelLayout
				 * (engine/src/core/Cubemap.cpp:28-31), so that the fused conversion equals the reference's two
				 * calls for every layout pair (added alpha = 255, Layout.cpp:57-58; gray from table outputs) */
				uint32_t in8[4] = {0, 0, 0, 0}, first8[4] = {0, 0, 0, 0};
				for(int c = 0; c < src_ch; ++c) {
					in8[c] = lut[in[c]];
					first8[c] = lut[first[c]];
				}
				layout_int_px(in8, first8, mid, src_ch, dst_ch, 8, mode);
				for(int c = 0; c < dst_ch; ++c) dp[c] = (uint8_t)mid[c];
				continue;
			}
			if(src_ch == dst_ch)
				memcpy(mid, in, sizeof mid);
			else
				layout_int_px(in, first, mid, src_ch, dst_ch, bits, mode);
			/* step 2: per-sample format change */
			const uint32_t M = (src_fmt == ORC_U8) ? 255u : 65535u;
			for(int c = 0; c < dst_ch; ++c) {
				uint32_t v = mid[c];
				uint8_t* o = dp + (size_t)c * db;
				if(dst_fmt == src_fmt) {
					if(db == 1) {
						o[0] = (uint8_t)v;
					} else {
						uint16_t t = (uint16_t)v;
						memcpy(o, &t, 2);
					}
				} else if(dst_fmt == ORC_U8) { /* U16 -> U8: the reference LUT (Colordepth.cpp:35) */
					o[0] = lut[v];
				} else if(dst_fmt == ORC_U16) { /* U8 -> U16: exact unorm widening */
					uint16_t t = (uint16_t)(v * 257u);
					memcpy(o, &t, 2);
				} else {
					float x = (float)v / (float)M; /* IEEE fp32 division */
					if(dst_fmt == ORC_F32) {
						memcpy(o, &x, 4);
					} else {
						uint16_t t = orc_f32_to_f16(x);
						memcpy(o, &t, 2);
					}
				}
			}
		} else {
			/* float source: layout change in fp32 (no quirks), then narrow/quantise */
			float in[4] = {0, 0, 0, 0}, mid[4];
			for(int c = 0; c < src_ch; ++c) in[c] = load_as_float(sp + (size_t)c * sb, src_fmt);
			if(src_ch == dst_ch) {
				memcpy(mid, in, sizeof mid);
			} else if(dst_ch == 1) {
				mid[0] = fmaf(0.0722f, in[2], fmaf(0.7152f, in[1], 0.2126f * in[0]));
			} else if(src_ch == 1) {
				mid[0] = mid[1] = mid[2] = in[0];
				mid[3] = 1.0f;
			} else {
				mid[0] = in[0];
				mid[1] = in[1];
				mid[2] = in[2];
				mid[3] = (src_ch == 4) ? in[3] : 1.0f;
			}
			for(int c = 0; c < dst_ch; ++c) {
				uint8_t* o = dp + (size_t)c * db;
				if(dst_fmt == ORC_F32) {
					memcpy(o, &mid[c], 4);
				} else if(dst_fmt == ORC_F16) {
					uint16_t t = orc_f32_to_f16(mid[c]); /* exact when the source was F16 */
					memcpy(o, &t, 2);
				} else if(dst_fmt == ORC_U16) {
					uint16_t t = (uint16_t)unorm_from_float(mid[c], 65535u);
					memcpy(o, &t, 2);
				} else {
					o[0] = (uint8_t)unorm_from_float(mid[c], 255u);
				}
			}
		}
	}
}

static int convert_args_ok(int src_ch, int dst_ch, int src_fmt, int dst_fmt) {
	if(!(src_ch == 1 || src_ch == 3 || src_ch == 4) || !(dst_ch == 1 || dst_ch == 3 || dst_ch == 4)) return 0;
	if(src_fmt < 0 || src_fmt > 3 || dst_fmt < 0 || dst_fmt > 3) return 0;
	return 1;
}

int orc_convert(const void* src, void* dst, uint64_t pixels, int src_ch, int dst_ch, int src_fmt, int dst_fmt,
				int mode) {
	if(!convert_args_ok(src_ch, dst_ch, src_fmt, dst_fmt)) return ORC_E_BAD_ARG;
	if(src_ch == dst_ch && src_fmt == dst_fmt) return ORC_E_SAME_LAYOUT;
	convert_range((const uint8_t*)src, (uint8_t*)dst, 0, pixels, (const uint8_t*)src, src_ch, dst_ch, src_fmt,
				  dst_fmt, mode);
	return ORC_OK;
}

int orc_max_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

int orc_convert_mt(const void* src, void* dst, uint64_t pixels, int src_ch, int dst_ch, int src_fmt, int dst_fmt,
				   int mode, int threads) {
	if(!convert_args_ok(src_ch, dst_ch, src_fmt, dst_fmt)) return ORC_E_BAD_ARG;
	if(src_ch == dst_ch && src_fmt == dst_fmt) return ORC_E_SAME_LAYOUT;
	if(threads < 1) threads = orc_max_threads();
	(void)orc_lut();
	const uint64_t chunk = 1u << 16;
	const int64_t nchunks = (int64_t)((pixels + chunk - 1) / chunk);
#pragma omp parallel for schedule(static) num_threads(threads)
	for(int64_t k = 0; k < nchunks; ++k) {
		uint64_t p0 = (uint64_t)k * chunk;
		uint64_t p1 = p0 + chunk > pixels ? pixels : p0 + chunk;
		convert_range((const uint8_t*)src, (uint8_t*)dst, p0, p1, (const uint8_t*)src, src_ch, dst_ch, src_fmt,
					  dst_fmt, mode);
	}
	return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * f-1 (Spec): Flip with the semantics libs/image/src/Flip.cpp:6-42 intends
 * (out.row[y] = src.row[h-1-y]; src untouched).  The reference itself is UB for h >= 2
 * (rowPitch multiplies by h, Flip.cpp:16) so there is nothing to pin against.
 * ---------------------------------------------------------------------------------------- */
int orc_flip_rows(const void* src, void* dst, uint32_t w, uint32_t h, uint32_t bytes_per_pixel) {
	if(w == 0 || h == 0 || bytes_per_pixel == 0) return ORC_E_BAD_ARG;
	const size_t pitch = (size_t)w * bytes_per_pixel;
	for(uint32_t y = 0; y < h; ++y)
		memcpy((uint8_t*)dst + (size_t)y * pitch, (const uint8_t*)src + (size_t)(h - 1 - y) * pitch, pitch);
	return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * f-4 follow-on (NEW, Spec B.3): one mip level, 2x2 box average.  No reference counterpart: the
 * engine leaves mips to the graphics API (glGenerateMipmap, engine/src/opengl/impls/OpenGLTex2D.cpp:51).
 *   out is max(1, w/2) x max(1, h/2); out(x, y) averages src(2x, 2y), src(min(2x+1, w-1), 2y),
 *   src(2x, min(2y+1, h-1)), src(min(2x+1, w-1), min(2y+1, h-1)) per channel, on stored values:
 *   integers (a + b + c + d + 2) >> 2; floats ((a + b) + (c + d)) * 0.25f in fp32 (F16 widened first,
 *   rounded to nearest even afterwards).
 * ---------------------------------------------------------------------------------------- */
int orc_downsample2x(const void* src, void* dst, uint32_t w, uint32_t h, int ch, int fmt) {
	if(w == 0 || h == 0 || !(ch == 1 || ch == 3 || ch == 4) || fmt < ORC_U8 || fmt > ORC_F32) return ORC_E_BAD_ARG;
	const uint32_t wo = w / 2 ? w / 2 : 1, ho = h / 2 ? h / 2 : 1;
	for(uint32_t y = 0; y < ho; ++y)
		for(uint32_t x = 0; x < wo; ++x) {
			const uint32_t x0 = 2 * x, x1 = (2 * x + 1 < w) ? 2 * x + 1 : w - 1, y0 = 2 * y, y1 = (2 * y + 1 < h) ? 2 * y + 1 : h - 1;
			for(int c = 0; c < ch; ++c) {
				const size_t ia = ((size_t)y0 * w + x0) * ch + c, ib = ((size_t)y0 * w + x1) * ch + c, ic = ((size_t)y1 * w + x0) * ch + c, id = ((size_t)y1 * w + x1) * ch + c;
				const size_t io = ((size_t)y * wo + x) * ch + c;
				if(fmt == ORC_U8) {
					const uint8_t* s = (const uint8_t*)src;
					((uint8_t*)dst)[io] = (uint8_t)(((uint32_t)s[ia] + s[ib] + s[ic] + s[id] + 2u) >> 2);
				} else if(fmt == ORC_U16) {
					const uint16_t* s = (const uint16_t*)src;
					((uint16_t*)dst)[io] = (uint16_t)(((uint32_t)s[ia] + s[ib] + s[ic] + s[id] + 2u) >> 2);
				} else if(fmt == ORC_F32) {
					const float* s = (const float*)src;
					const float top = s[ia] + s[ib], bot = s[ic] + s[id];
					((float*)dst)[io] = (top + bot) * 0.25f;
				} else {
					const uint16_t* s = (const uint16_t*)src;
					const float top = orc_f16_to_f32(s[ia]) + orc_f16_to_f32(s[ib]), bot = orc_f16_to_f32(s[ic]) + orc_f16_to_f32(s[id]);
					((uint16_t*)dst)[io] = orc_f32_to_f16((top + bot) * 0.25f);
				}
			}
		}
	return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * f-3: texel unpack for assimp embedded textures, engine/src/core/Model.cpp:245-316.
 *
 * hint = aiTexture::achFormatHint, e.g. "rgba8888", "argb8888", "bgra5650": four channel letters,
 * then four bit counts (Model.cpp:250-253).  Validation and layout choice follow :255-270
 * (zeroed channels: none -> RGBA, alpha only -> RGB, three -> Grayscale).  Each kept channel moves
 * to its place in r,g,b,a order (:280-293) and counts below 8 are widened as :305-310 do:
 *     v = (v & ((1 << bits) - 1)) << (8 - bits);  v |= v >> bits;
 * Reference-exact for hints without a zero count.  With zeroed channels the reference indexes the
 * 4-byte source texels with the output stride (base = texel * (4 - zeroed), :299,:304) and its
 * grayscale branch (:318-331) widens every source byte with a different, ineffective formula; those
 * cannot be the intent, so -- as with Flip -- the intended semantics are restated: 4 source bytes
 * per texel, out_ch output bytes per texel, the same widening everywhere.
 * Returns the output channel count (4, 3 or 1) or a negative error.
 * ---------------------------------------------------------------------------------------- */
int orc_unpack_texels(const uint8_t* src, uint8_t* dst, uint64_t texels, const char* hint) {
	static const char order[5] = "rgba";
	int src_index[4] = {-1, -1, -1, -1}, bits[4] = {0, 0, 0, 0}, zeroed = 0;
	if(!hint) return ORC_E_BAD_ARG;
	for(int i = 0; i < 8; ++i)
		if(hint[i] == 0) return ORC_E_BAD_ARG;
	for(int i = 0; i < 4; ++i) {
		const char* at = strchr(order, hint[i]);
		if(!at || hint[i + 4] < '0' || hint[i + 4] > '8') return ORC_E_BAD_ARG;
		const int c = (int)(at - order);
		if(src_index[c] >= 0) return ORC_E_BAD_ARG; /* letter used twice */
		src_index[c] = i;
		bits[c] = hint[i + 4] - '0';
	}
	for(int c = 0; c < 4; ++c)
		if(bits[c] == 0) ++zeroed;
	if(zeroed == 2 || zeroed == 4) return ORC_E_BAD_ARG;       /* Model.cpp:259,269 */
	if(zeroed == 1 && bits[3] != 0) return ORC_E_BAD_ARG;      /* :260 only alpha may be zero */
	const int out_ch = 4 - zeroed;
	for(uint64_t t = 0; t < texels; ++t) {
		int o = 0;
		for(int c = 0; c < 4; ++c) {
			if(bits[c] == 0) continue;
			uint8_t v = src[t * 4 + (uint64_t)src_index[c]];
			if(bits[c] < 8) {
				v = (uint8_t)((v & ((1 << bits[c]) - 1)) << (8 - bits[c]));
				v |= (uint8_t)(v >> bits[c]);
			}
			dst[t * (uint64_t)out_ch + (uint64_t)o++] = v;
		}
	}
	return out_ch;
}

/* ------------------------------------------------------------------------------------------
 * a7 (NEW, Spec B.2): equirectangular -> cubemap.
 *
 * Everything below is fp32 with a fixed operation order; fused multiply-adds appear ONLY as
 * explicit fmaf().  The CUDA kernel follows the same spec, so coordinates are bit-identical.
 * ---------------------------------------------------------------------------------------- */
#define ORC_PI 3.14159265358979323846f
#define ORC_PI_2 1.57079632679489661923f
#define ORC_PI_4 0.78539816339744830962f
#define ORC_INV_2PI 0.15915494309189533577f
#define ORC_INV_PI 0.31830988618379067154f
#define ORC_TAN_PI_8 0.41421356237309504880f

float orc_atan2(float y, float x) {
	const float ax = fabsf(x), ay = fabsf(y);
	const float a = ax < ay ? ax : ay; /* min */
	const float b = ax < ay ? ay : ax; /* max */
	if(b == 0.0f) return 0.0f;
	float num, den, off;
	if(a > ORC_TAN_PI_8 * b) { /* atan(t) = pi/4 + atan((t-1)/(t+1)) */
		num = a - b;
		den = a + b;
		off = ORC_PI_4;
	} else {
		num = a;
		den = b;
		off = 0.0f;
	}
	const float t = num / den;
	const float z = t * t;
	float p = fmaf(8.05374449538e-2f, z, -1.38776856032e-1f);
	p = fmaf(p, z, 1.99777106478e-1f);
	p = fmaf(p, z, -3.33329491539e-1f);
	float r = fmaf(p * z, t, t);
	r = r + off;
	if(ay > ax) r = ORC_PI_2 - r;
	if(x < 0.0f) r = ORC_PI - r;
	if(y < 0.0f) r = -r;
	return r;
}

static void face_dir(uint32_t face, float sc, float tc, float* x, float* y, float* z) {
	/* Khronos cube table, faces +X,-X,+Y,-Y,+Z,-Z
	 * (order: libs/formats/include/libcacaoformats.hpp:213, tools/cubetool/commands.hpp:31) */
	switch(face) {
		case 0: *x = 1.0f;  *y = -tc;  *z = -sc;  break;
		case 1: *x = -1.0f; *y = -tc;  *z = sc;   break;
		case 2: *x = sc;    *y = 1.0f; *z = tc;   break;
		case 3: *x = sc;    *y = -1.0f; *z = -tc; break;
		case 4: *x = sc;    *y = -tc;  *z = 1.0f; break;
		default: *x = -sc;  *y = -tc;  *z = -1.0f; break;
	}
}

void orc_project(uint32_t face, uint32_t i, uint32_t j, uint32_t N, uint32_t W, uint32_t H, float* xs,
				 float* ys) {
	const float invN = 1.0f / (float)N;
	/* (2k + 1 - N) / N as exact integer times 1/N: antisymmetric under k -> N-1-k */
	const float sc = (float)((int32_t)(2u * i + 1u) - (int32_t)N) * invN;
	const float tc = (float)((int32_t)(2u * j + 1u) - (int32_t)N) * invN;
	float x, y, z;
	face_dir(face, sc, tc, &x, &y, &z);
	const float phi = orc_atan2(x, z);
	const float rxz = sqrtf(fmaf(x, x, z * z));
	const float theta = orc_atan2(y, rxz);
	const float kx = (float)W * ORC_INV_2PI;
	const float ky = (float)H * ORC_INV_PI;
	const float cx = 0.5f * (float)W - 0.5f;
	const float cy = 0.5f * (float)H - 0.5f;
	*xs = fmaf(phi, kx, cx);
	*ys = fmaf(-theta, ky, cy);
}

static float clampf(float v, float lo, float hi) {
	return v < lo ? lo : (v > hi ? hi : v);
}

static void store_sample(uint8_t* o, float r, int fmt) {
	switch(fmt) {
		case ORC_F32: memcpy(o, &r, 4); break;
		case ORC_F16: {
			uint16_t t = orc_f32_to_f16(r);
			memcpy(o, &t, 2);
			break;
		}
		case ORC_U16: {
			/* uintN(min(max(r,0),M) + 0.5f); NaN cannot arise from integer sources */
			uint16_t t = (uint16_t)(clampf(r, 0.0f, 65535.0f) + 0.5f);
			memcpy(o, &t, 2);
			break;
		}
		default: o[0] = (uint8_t)(clampf(r, 0.0f, 255.0f) + 0.5f); break;
	}
}

int orc_equirect_to_cube(const void* src, uint32_t W, uint32_t H, int ch, int fmt, void* dst, uint32_t N,
						 uint32_t face_mask, uint32_t row_begin, uint32_t row_end, int threads) {
	if(!src || !dst || W == 0 || H == 0 || N == 0 || !(ch == 1 || ch == 3 || ch == 4) || fmt < 0 || fmt > 3)
		return ORC_E_BAD_ARG;
	if(row_end > N) row_end = N;
	if(threads < 1) threads = orc_max_threads();
	const int sb = fmt_bytes(fmt);
	const size_t bpp = (size_t)ch * sb;
	const uint8_t* s = (const uint8_t*)src;
	uint8_t* d = (uint8_t*)dst;
	for(uint32_t face = 0; face < 6; ++face) {
		if(!(face_mask & (1u << face))) continue;
#pragma omp parallel for schedule(static) num_threads(threads)
		for(int64_t jj = (int64_t)row_begin; jj < (int64_t)row_end; ++jj) {
			const uint32_t j = (uint32_t)jj;
			for(uint32_t i = 0; i < N; ++i) {
				float xs, ys;
				orc_project(face, i, j, N, W, H, &xs, &ys);
				const float fx0 = floorf(xs), fy0 = floorf(ys);
				const float fx = xs - fx0, fy = ys - fy0;
				int32_t x0 = (int32_t)fx0, y0 = (int32_t)fy0;
				int32_t x1 = x0 + 1, y1 = y0 + 1;
				/* columns wrap, rows clamp */
				if(x0 < 0) x0 += (int32_t)W;
				if(x0 >= (int32_t)W) x0 -= (int32_t)W;
				if(x1 < 0) x1 += (int32_t)W;
				if(x1 >= (int32_t)W) x1 -= (int32_t)W;
				if(y0 < 0) y0 = 0;
				if(y0 > (int32_t)H - 1) y0 = (int32_t)H - 1;
				if(y1 < 0) y1 = 0;
				if(y1 > (int32_t)H - 1) y1 = (int32_t)H - 1;
				const uint8_t* pa = s + ((size_t)y0 * W + (size_t)x0) * bpp;
				const uint8_t* pb = s + ((size_t)y0 * W + (size_t)x1) * bpp;
				const uint8_t* pc = s + ((size_t)y1 * W + (size_t)x0) * bpp;
				const uint8_t* pd = s + ((size_t)y1 * W + (size_t)x1) * bpp;
				uint8_t* o = d + (((size_t)face * N + j) * N + i) * bpp;
				for(int c = 0; c < ch; ++c) {
					const float a = load_as_float(pa + (size_t)c * sb, fmt);
					const float b = load_as_float(pb + (size_t)c * sb, fmt);
					const float cc = load_as_float(pc + (size_t)c * sb, fmt);
					const float dd = load_as_float(pd + (size_t)c * sb, fmt);
					const float top = fmaf(fx, b - a, a);
					const float bot = fmaf(fx, dd - cc, cc);
					const float r = fmaf(fy, bot - top, top);
					store_sample(o + (size_t)c * sb, r, fmt);
				}
			}
		}
	}
	return ORC_OK;
}

/* Independent check of the spec: same geometry in double precision with libm atan2. */
int orc_equirect_to_cube_f64(const void* src, uint32_t W, uint32_t H, int ch, int fmt, double* dst, uint32_t N,
							 uint32_t face_mask, uint32_t row_begin, uint32_t row_end) {
	if(!src || !dst || W == 0 || H == 0 || N == 0) return ORC_E_BAD_ARG;
	if(row_end > N) row_end = N;
	const int sb = fmt_bytes(fmt);
	const size_t bpp = (size_t)ch * sb;
	const uint8_t* s = (const uint8_t*)src;
	const double PI = 3.14159265358979323846;
	for(uint32_t face = 0; face < 6; ++face) {
		if(!(face_mask & (1u << face))) continue;
#pragma omp parallel for schedule(static)
		for(int64_t jj = (int64_t)row_begin; jj < (int64_t)row_end; ++jj) {
			const uint32_t j = (uint32_t)jj;
			for(uint32_t i = 0; i < N; ++i) {
				const double sc = (2.0 * i + 1.0) / N - 1.0, tc = (2.0 * j + 1.0) / N - 1.0;
				double x, y, z;
				switch(face) {
					case 0: x = 1; y = -tc; z = -sc; break;
					case 1: x = -1; y = -tc; z = sc; break;
					case 2: x = sc; y = 1; z = tc; break;
					case 3: x = sc; y = -1; z = -tc; break;
					case 4: x = sc; y = -tc; z = 1; break;
					default: x = -sc; y = -tc; z = -1; break;
				}
				const double phi = atan2(x, z), theta = atan2(y, sqrt(x * x + z * z));
				const double u = 0.5 + phi / (2.0 * PI), v = 0.5 - theta / PI;
				const double xs = u * W - 0.5, ys = v * H - 0.5;
				const double fx0 = floor(xs), fy0 = floor(ys);
				const double fx = xs - fx0, fy = ys - fy0;
				int64_t x0 = (int64_t)fx0, y0 = (int64_t)fy0;
				int64_t x1 = x0 + 1, y1 = y0 + 1;
				x0 = ((x0 % (int64_t)W) + W) % W;
				x1 = ((x1 % (int64_t)W) + W) % W;
				if(y0 < 0) y0 = 0;
				if(y0 > (int64_t)H - 1) y0 = H - 1;
				if(y1 < 0) y1 = 0;
				if(y1 > (int64_t)H - 1) y1 = H - 1;
				for(int c = 0; c < ch; ++c) {
					const double a = load_as_float(s + ((size_t)y0 * W + x0) * bpp + (size_t)c * sb, fmt);
					const double b = load_as_float(s + ((size_t)y0 * W + x1) * bpp + (size_t)c * sb, fmt);
					const double cc = load_as_float(s + ((size_t)y1 * W + x0) * bpp + (size_t)c * sb, fmt);
					const double dd = load_as_float(s + ((size_t)y1 * W + x1) * bpp + (size_t)c * sb, fmt);
					const double top = a + fx * (b - a), bot = cc + fx * (dd - cc);
					dst[(((size_t)face * N + j) * N + i) * ch + c] = top + fy * (bot - top);
				}
			}
		}
	}
	return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * Spec 8(d): synthetic inputs.  sample value = f(hash32(seed, flat sample index)).
 * ---------------------------------------------------------------------------------------- */
uint32_t orc_hash32(uint32_t seed, uint64_t idx) {
	uint32_t h = seed ^ ((uint32_t)idx * 0x9E3779B9u) ^ ((uint32_t)(idx >> 32) * 0x85EBCA6Bu);
	h ^= h >> 16; /* murmur3 fmix32 */
	h *= 0x85EBCA6Bu;
	h ^= h >> 13;
	h *= 0xC2B2AE35u;
	h ^= h >> 16;
	return h;
}

static float synth_float(uint32_t h) {
	/* (1 + m/65536) * 2^(e-8), e in [0,16): built from bits so host and device agree exactly */
	uint32_t e = (h >> 24) & 15u, m = (h >> 8) & 0xFFFFu;
	uint32_t bits = ((127u - 8u + e) << 23) | (m << 7);
	float f;
	memcpy(&f, &bits, 4);
	return f;
}

void orc_synth(void* dst, uint64_t first_sample, uint64_t samples, int fmt, int ch, uint32_t seed) {
	for(uint64_t k = 0; k < samples; ++k) {
		const uint64_t idx = first_sample + k;
		const uint32_t h = orc_hash32(seed, idx);
		const int is_alpha = (ch == 4) && ((idx & 3u) == 3u);
		switch(fmt) {
			case ORC_U8: ((uint8_t*)dst)[k] = (uint8_t)h; break;
			case ORC_U16: ((uint16_t*)dst)[k] = (uint16_t)h; break;
			case ORC_F16: ((uint16_t*)dst)[k] = is_alpha ? 0x3C00u : orc_f32_to_f16(synth_float(h)); break;
			default: ((float*)dst)[k] = is_alpha ? 1.0f : synth_float(h); break;
		}
	}
}

uint64_t orc_fnv1a64(const void* p, uint64_t n) {
	const uint8_t* b = (const uint8_t*)p;
	uint64_t h = 0xcbf29ce484222325ull;
	for(uint64_t i = 0; i < n; ++i) {
		h ^= b[i];
		h *= 0x100000001b3ull;
	}
	return h;
}
