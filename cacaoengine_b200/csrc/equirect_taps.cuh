// equirect_taps.cuh -- the per-texel pieces shared by the equirect -> cubemap kernels (equirect.cu,
// equirect_rows.cu): tap loads, integer <-> float sample conversion without the conversion unit, the
// packed-fp32 bilinear filter and the texel store.  DESIGN.md "Spec B.2" fixes every rounding here.
#pragma once

#include <cstddef>
#include <cstdint>

#include "cacao_device.cuh"
#include "convert_pixel.cuh"
#include "equirect.cuh"
#include "projection.cuh"

namespace cacao {

	namespace {

		// ---- one texel: raw words from global memory, then CH floats -------------------------------
		template<int F, int CH>
		__device__ __forceinline__ void load_words(const uint8_t* __restrict__ p, uint32_t (&w)[4]) {
			constexpr int sb = FmtTraits<F>::bytes;
			constexpr int bpp = sb * CH;
			if constexpr(bpp == 16) {
				const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
				w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
			} else if constexpr(bpp == 8) {
				const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
				w[0] = q.x; w[1] = q.y;
			} else if constexpr(bpp == 4) {
				w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
			} else if constexpr(bpp == 12) {
				const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
				w[0] = __ldg(q); w[1] = __ldg(q + 1); w[2] = __ldg(q + 2);
			} else if constexpr(bpp == 6) {
				const uint16_t* q = reinterpret_cast<const uint16_t*>(p);
				w[0] = __ldg(q) | (static_cast<uint32_t>(__ldg(q + 1)) << 16);
				w[1] = __ldg(q + 2);
			} else if constexpr(bpp == 3) {
				w[0] = __ldg(p) | (static_cast<uint32_t>(__ldg(p + 1)) << 8) | (static_cast<uint32_t>(__ldg(p + 2)) << 16);
			} else if constexpr(bpp == 2) {
				w[0] = __ldg(reinterpret_cast<const uint16_t*>(p));
			} else {
				w[0] = __ldg(p);
			}
		}

		// Integer samples become floats without an I2F: one PRMT drops the byte (or 16-bit half) into
		// the mantissa of 2^23, one exact subtraction removes the 2^23 again -- float(v) bit for bit,
		// on the full-rate pipes instead of the quarter-rate conversion unit.
		constexpr uint32_t kMagicBits = 0x4B000000u; // 8388608.0f
		constexpr float kMagic = 8388608.0f;

		// BIASED leaves the 2^23 in place for callers that remove it themselves (sample_and_store).
		template<int F, int CH, bool BIASED = false>
		__device__ __forceinline__ void words_to_floats(const uint32_t (&w)[4], float (&v)[CH]) {
			constexpr float kOff = BIASED ? 0.0f : kMagic;
#pragma unroll
			for(int c = 0; c < CH; ++c) {
				if constexpr(F == CACAO_FMT_U8)
					v[c] = __uint_as_float(__byte_perm(w[c >> 2], kMagicBits, 0x7540u | static_cast<uint32_t>(c & 3))) - kOff;
				else if constexpr(F == CACAO_FMT_U16)
					v[c] = __uint_as_float(__byte_perm(w[c >> 1], kMagicBits, (c & 1) ? 0x7632u : 0x7610u)) - kOff;
				else if constexpr(F == CACAO_FMT_F16)
					v[c] = f16_bits_to_f32(static_cast<uint16_t>((w[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu));
				else
					v[c] = __uint_as_float(w[c]);
			}
		}

		template<int F, int CH, bool BIASED = false>
		__device__ __forceinline__ void load_texel(const uint8_t* __restrict__ p, float (&v)[CH]) {
			uint32_t w[4] = {0, 0, 0, 0};
			load_words<F, CH>(p, w);
			words_to_floats<F, CH, BIASED>(w, v);
		}

		// Three-channel texels are 3 / 6 / 12 bytes: one texel is no power-of-two load, but the two
		// taps of a row (x0 and x0+1 when the pair does not straddle the u = 0/1 seam) are 6 / 12 / 24
		// CONTIGUOUS bytes.  load_pair3 fetches the 8-byte-aligned window around them with 64-bit loads
		// and realigns in registers (word select + funnel shift) -- 2-4 load instructions per row instead
		// of 6, and a third of the (request, line) pairs the L1 tag stage has to process, which is what
		// bounds this kernel.  `elem` = texel index of the x0 tap in the source window.
		template<int F, bool BIASED>
		__device__ __forceinline__ void load_pair3_at(const uint8_t* __restrict__ at, float (&a)[3], float (&b)[3]);
		template<int F, bool BIASED, typename IDX>
		__device__ __forceinline__ void load_pair3(const uint8_t* __restrict__ src, IDX elem, float (&a)[3], float (&b)[3]) {
			load_pair3_at<F, BIASED>(src + static_cast<size_t>(elem) * (3 * FmtTraits<F>::bytes), a, b);
		}
		// `at` = address of the x0 tap (the buffer is 16-byte aligned, so address and offset align alike)
		template<int F, bool BIASED>
		__device__ __forceinline__ void load_pair3_at(const uint8_t* __restrict__ at, float (&a)[3], float (&b)[3]) {
			constexpr int sb = FmtTraits<F>::bytes;
			const uintptr_t off = reinterpret_cast<uintptr_t>(at);
			const uint8_t* base = reinterpret_cast<const uint8_t*>(off & ~static_cast<uintptr_t>(7));
			const uint32_t o = static_cast<uint32_t>(off) & 7u;
			const bool e = (o & 4u) != 0;
			if constexpr(sb == 4) {
				// 24 bytes at a 4-byte-aligned offset: words q[o .. o+5] of the 16-byte-aligned window, o in 0..3
				const uint8_t* b16 = reinterpret_cast<const uint8_t*>(off & ~static_cast<uintptr_t>(15));
				const uint32_t o4 = (static_cast<uint32_t>(off) & 15u) >> 2;
				const uint4 lo = __ldg(reinterpret_cast<const uint4*>(b16));
				const uint4 hi = __ldg(reinterpret_cast<const uint4*>(b16 + 16));
				uint32_t q8 = 0;
				if(o4 == 3u) q8 = __ldg(reinterpret_cast<const uint32_t*>(b16 + 32));
				const uint32_t q[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, q8};
				const bool s1 = (o4 & 1u) != 0, s2 = (o4 & 2u) != 0;
				uint32_t u[8];
#pragma unroll
				for(int k = 0; k < 8; ++k) u[k] = s1 ? q[k + 1] : q[k];
				a[0] = __uint_as_float(s2 ? u[2] : u[0]);
				a[1] = __uint_as_float(s2 ? u[3] : u[1]);
				a[2] = __uint_as_float(s2 ? u[4] : u[2]);
				b[0] = __uint_as_float(s2 ? u[5] : u[3]);
				b[1] = __uint_as_float(s2 ? u[6] : u[4]);
				b[2] = __uint_as_float(s2 ? u[7] : u[5]);
			} else if constexpr(sb == 2) {
				// 12 bytes at a 2-byte-aligned offset: at most 18 bytes into the window
				const uint2 q01 = __ldg(reinterpret_cast<const uint2*>(base));
				const uint2 q23 = __ldg(reinterpret_cast<const uint2*>(base + 8));
				uint32_t q4 = 0;
				if(o == 6u) q4 = __ldg(reinterpret_cast<const uint32_t*>(base + 16));
				const uint32_t sh = (o & 3u) * 8u;
				const uint32_t w0 = e ? q01.y : q01.x, w1 = e ? q23.x : q01.y, w2 = e ? q23.y : q23.x, w3 = e ? q4 : q23.y;
				const uint32_t v[4] = {__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), 0u};
				float t[6];
#pragma unroll
				for(int k = 0; k < 6; ++k) {
					if constexpr(F == CACAO_FMT_U16)
						t[k] = __uint_as_float(__byte_perm(v[k >> 1], kMagicBits, (k & 1) ? 0x7632u : 0x7610u)) - (BIASED ? 0.0f : kMagic);
					else
						t[k] = f16_bits_to_f32(static_cast<uint16_t>((v[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu));
				}
				a[0] = t[0]; a[1] = t[1]; a[2] = t[2];
				b[0] = t[3]; b[1] = t[4]; b[2] = t[5];
			} else {
				// 6 bytes at any offset: at most 13 bytes into the window
				const uint2 q01 = __ldg(reinterpret_cast<const uint2*>(base));
				const uint2 q23 = __ldg(reinterpret_cast<const uint2*>(base + 8));
				const uint32_t sh = (o & 3u) * 8u;
				const uint32_t w0 = e ? q01.y : q01.x, w1 = e ? q23.x : q01.y, w2 = e ? q23.y : q23.x;
				const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
				constexpr float kOff = BIASED ? 0.0f : kMagic;
				a[0] = __uint_as_float(__byte_perm(v0, kMagicBits, 0x7540u)) - kOff;
				a[1] = __uint_as_float(__byte_perm(v0, kMagicBits, 0x7541u)) - kOff;
				a[2] = __uint_as_float(__byte_perm(v0, kMagicBits, 0x7542u)) - kOff;
				b[0] = __uint_as_float(__byte_perm(v0, kMagicBits, 0x7543u)) - kOff;
				b[1] = __uint_as_float(__byte_perm(v1, kMagicBits, 0x7540u)) - kOff;
				b[2] = __uint_as_float(__byte_perm(v1, kMagicBits, 0x7541u)) - kOff;
			}
		}

		// Integer stores: the spec's clamp to [0, M] cannot change anything -- every lerp is
		// fma(f, q - p, p) with f in [0, 1] and p, q exactly representable, so its exact value lies
		// between p and q and the (monotone) rounding keeps it there; the filtered value of integer
		// texels is therefore already inside [0, M] and never NaN.  The clamp is dropped here, not in
		// the spec (oracle/cacao_oracle.c keeps it), and the GPU tests compare the two bit for bit.
		// NORM (fused format conversion, Spec B.4): v is a normalised float sample and an integer store
		// quantises it as Spec B.1 does -- (x != x) ? 0 : uintN(min(max(x,0),1)*M + 0.5f).
		// FASTINT: the truncation of v + 0.5f runs on the FP32 pipe -- adding 2^23 with round-toward-zero
		// leaves trunc(v + 0.5f) in the low mantissa bits (v + 0.5f < 2^23), which the packing PRMTs pick
		// up directly -- instead of a quarter-rate F2I per sample (equirect_rows.cu).
		template<int F, int CH, bool NORM = false, bool FASTINT = false>
		__device__ __forceinline__ void store_texel(uint8_t* __restrict__ p, const float (&v)[CH]) {
			constexpr int sb = FmtTraits<F>::bytes;
			constexpr int bpp = sb * CH;
			uint32_t w[4] = {0, 0, 0, 0};
			if constexpr(F == CACAO_FMT_F32) {
#pragma unroll
				for(int c = 0; c < CH; ++c) w[c] = __float_as_uint(v[c]);
			} else if constexpr(F == CACAO_FMT_F16) {
#pragma unroll
				for(int c = 0; c < CH; c += 2) w[c >> 1] = (c + 1 < CH) ? f32x2_to_f16x2_bits(v[c], v[(c + 1 < CH) ? c + 1 : c]) : f32_to_f16_bits(v[c]);
			} else {
				uint32_t s[CH];
#pragma unroll
				for(int c = 0; c < CH; ++c) {
					if constexpr(NORM)
						s[c] = f32_to_unorm<(F == CACAO_FMT_U8) ? 255 : 65535>(v[c]);
					else if constexpr(FASTINT)
						s[c] = __float_as_uint(__fadd_rz(v[c] + 0.5f, kMagic)); // 0x4B00vvvv: the packing below takes the low bytes only
					else
						s[c] = __float2uint_rz(v[c] + 0.5f);
				}
				if constexpr(sb == 2) {
#pragma unroll
					for(int c = 0; c < CH; c += 2) w[c >> 1] = (c + 1 < CH) ? __byte_perm(s[c], s[(c + 1 < CH) ? c + 1 : c], 0x5410u) : s[c];
				} else if constexpr(CH == 1) {
					w[0] = s[0];
				} else {
					const uint32_t lo = __byte_perm(s[0], s[1], 0x0040u);
					const uint32_t hi = (CH == 4) ? __byte_perm(s[2], s[CH - 1], 0x0040u) : s[2];
					w[0] = __byte_perm(lo, hi, 0x5410u);
				}
			}
			if constexpr(bpp == 16) {
				*reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
			} else if constexpr(bpp == 8) {
				*reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
			} else if constexpr(bpp == 4) {
				*reinterpret_cast<uint32_t*>(p) = w[0];
			} else if constexpr(bpp == 12) {
				uint32_t* q = reinterpret_cast<uint32_t*>(p);
				q[0] = w[0]; q[1] = w[1]; q[2] = w[2];
			} else if constexpr(bpp == 6) {
				uint16_t* q = reinterpret_cast<uint16_t*>(p);
				q[0] = static_cast<uint16_t>(w[0]); q[1] = static_cast<uint16_t>(w[0] >> 16); q[2] = static_cast<uint16_t>(w[1]);
			} else if constexpr(bpp == 3) {
				p[0] = static_cast<uint8_t>(w[0]); p[1] = static_cast<uint8_t>(w[0] >> 8); p[2] = static_cast<uint8_t>(w[0] >> 16);
			} else if constexpr(bpp == 2) {
				*reinterpret_cast<uint16_t*>(p) = static_cast<uint16_t>(w[0]);
			} else {
				p[0] = static_cast<uint8_t>(w[0]);
			}
		}

		// ---- packed fp32 (sm_100 FFMA2): two channels per instruction --------------------------------
		// q - p is written fma(p, -1, q): p * (-1) is exact, so the single rounding is the subtraction's.
		__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
			uint64_t r;
			asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
			return r;
		}
		__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
			uint64_t r;
			asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
			return r;
		}
		// BIAS != 0: the four taps still carry the integer-conversion offset 2^23.  It cancels exactly
		// in q - p (both are integers below 2^24), and p alone loses it through one exact packed add.
		template<bool BIASED>
		__device__ __forceinline__ void lerp2(float fx, float fy, const float* a, const float* b, const float* c, const float* d, float* out) {
			const uint64_t m1 = pk2(-1.0f, -1.0f), vx = pk2(fx, fx), vy = pk2(fy, fy);
			uint64_t A = pk2(a[0], a[1]), C = pk2(c[0], c[1]);
			const uint64_t B = pk2(b[0], b[1]), D = pk2(d[0], d[1]);
			const uint64_t dx0 = fma2(A, m1, B), dx1 = fma2(C, m1, D);
			if constexpr(BIASED) {
				const uint64_t one = pk2(1.0f, 1.0f), off = pk2(-kMagic, -kMagic);
				A = fma2(A, one, off);
				C = fma2(C, one, off);
			}
			const uint64_t top = fma2(vx, dx0, A);
			const uint64_t bot = fma2(vx, dx1, C);
			const uint64_t o = fma2(vy, fma2(top, m1, bot), top);
			asm("mov.b64 {%0,%1}, %2;" : "=f"(out[0]), "=f"(out[1]) : "l"(o));
		}

		// The source rows of one output row: element offsets of the two tap rows and the row weight.
		template<typename IDX>
		struct RowTaps {
			IDX row0, row1;
			float fy;
		};
		template<typename IDX>
		__device__ __forceinline__ RowTaps<IDX> row_taps(float ys, const EquirectParams& prm) {
			const float fy0 = floorf(ys);
			int y0 = static_cast<int>(fy0);
			int y1 = y0 + 1;
			const int H = static_cast<int>(prm.H);
			y0 = min(max(y0, 0), H - 1); // rows clamp
			y1 = min(max(y1, 0), H - 1);
			// rows are stored from src_row0 on; the window was planned on the host with the same
			// arithmetic, the clamp below only guards memory if a caller passes a short window
			const int r0 = min(max(y0 - static_cast<int>(prm.src_row0), 0), static_cast<int>(prm.src_rows) - 1);
			const int r1 = min(max(y1 - static_cast<int>(prm.src_row0), 0), static_cast<int>(prm.src_rows) - 1);
			RowTaps<IDX> t;
			t.row0 = static_cast<IDX>(r0) * static_cast<IDX>(prm.W);
			t.row1 = static_cast<IDX>(r1) * static_cast<IDX>(prm.W);
			t.fy = ys - fy0;
			return t;
		}

		// Bilinear fetch + store of one output pixel whose source rows are already known.
		// IDX is uint32_t when the source window has < 2^32 pixels (cheaper address arithmetic).
		// DF / DC = F / CH: the filtered texel is stored in the source's format (Spec B.2).  Anything
		// else (float sources only): the fp32 result goes through Spec B.1's float -> DF conversion and gains
		// alpha 1 (RGB -> RGBA) or loses its alpha (RGBA -> RGB) on its way out -- resampling and conversion
		// in one pass (Spec B.4).
		template<int F, int CH, typename IDX, int DF = F, int DC = CH>
		__device__ __forceinline__ void sample_and_store(const uint8_t* __restrict__ src, uint8_t* __restrict__ out_px, float xs, int W, const RowTaps<IDX>& rt, IDX total_px) {
			constexpr int bpp = FmtTraits<F>::bytes * CH;
			const float fx0 = floorf(xs);
			const float fx = xs - fx0;
			int x0 = static_cast<int>(fx0);
			int x1 = x0 + 1;
			// columns wrap: xs lies in [-0.5, W-0.5], so one conditional step each way is enough
			if(x0 < 0) x0 += W;
			if(x0 >= W) x0 -= W;
			if(x1 < 0) x1 += W;
			if(x1 >= W) x1 -= W;
			constexpr bool kBiased = FmtTraits<F>::is_int; // integer taps keep their 2^23 until the lerp
			float a[CH], b[CH], c[CH], d[CH], out[CH];
			bool joint = false;
			if constexpr(CH == 3) {
				// the aligned windows of load_pair3 reach at most 24 bytes (8 texels) past the x0 tap
				const IDX last = (rt.row0 > rt.row1 ? rt.row0 : rt.row1) + static_cast<IDX>(x0);
				joint = (x1 == x0 + 1) && (last + static_cast<IDX>(8) <= total_px);
				if(joint) {
					load_pair3<F, kBiased, IDX>(src, rt.row0 + static_cast<IDX>(x0), a, b);
					load_pair3<F, kBiased, IDX>(src, rt.row1 + static_cast<IDX>(x0), c, d);
				}
			}
			if(!joint) {
				load_texel<F, CH, kBiased>(src + static_cast<size_t>(rt.row0 + static_cast<IDX>(x0)) * bpp, a);
				load_texel<F, CH, kBiased>(src + static_cast<size_t>(rt.row0 + static_cast<IDX>(x1)) * bpp, b);
				load_texel<F, CH, kBiased>(src + static_cast<size_t>(rt.row1 + static_cast<IDX>(x0)) * bpp, c);
				load_texel<F, CH, kBiased>(src + static_cast<size_t>(rt.row1 + static_cast<IDX>(x1)) * bpp, d);
			}
			constexpr int kPairs = CH / 2;
#pragma unroll
			for(int k = 0; k < kPairs; ++k) lerp2<kBiased>(fx, rt.fy, a + 2 * k, b + 2 * k, c + 2 * k, d + 2 * k, out + 2 * k);
#pragma unroll
			for(int k = 2 * kPairs; k < CH; ++k) {
				constexpr float kOff = kBiased ? kMagic : 0.0f;
				const float top = fmaf(fx, b[k] - a[k], a[k] - kOff);
				const float bot = fmaf(fx, d[k] - c[k], c[k] - kOff);
				out[k] = fmaf(rt.fy, bot - top, top);
			}
			if constexpr(DF == F && DC == CH) {
				store_texel<F, CH>(out_px, out);
			} else {
				static_assert(!FmtTraits<F>::is_int, "fused format conversion is defined for float sources");
				static_assert(DC == CH || (CH == 3 && DC == 4) || (CH == 4 && DC == 3), "fused layout change: RGB <-> RGBA only");
				float o[DC];
#pragma unroll
				for(int c = 0; c < DC; ++c) o[c] = c < CH ? out[c < CH ? c : 0] : 1.0f;
				store_texel<DF, DC, FmtTraits<DF>::is_int>(out_px, o);
			}
		}

	} // namespace

} // namespace cacao
