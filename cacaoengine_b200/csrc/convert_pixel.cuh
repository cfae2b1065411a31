// convert_pixel.cuh -- the per-pixel arithmetic of the fused layout + sample-format conversion.
//
// One place defines what a converted pixel is; the two vector kernels (convert_tiles.cuh,
// convert_wide.cuh) and the byte-wise edge kernel (convert_dispatch.cuh) all call it.
//
// Reference semantics reproduced here (paths relative to /root/reference/):
//   ->Gray (int)   trunc(0.2126*r)+trunc(0.7152*g)+trunc(0.0722*b), double products truncated
//                  one by one, sum clamped to 255 even at 16 bit      libs/image/src/Layout.cpp:33-41,69-77
//   Gray->RGB(A)   replicate; alpha = opaque max                         Layout.cpp:47-61,93-94
//   RGB<->RGBA     copy / drop / add opaque alpha                        Layout.cpp:49-62
//   U16->U8        table lookup                                          libs/image/src/Colordepth.cpp:32-35
// Everything touching F16/F32 is DESIGN.md "Spec B.1" (no reference counterpart).
#pragma once

#include "cacao_device.cuh"

namespace cacao {

	// trunc(c * (double)x) for the three BT.709 weights, x <= 65535, without touching the FP64 pipe
	// on the common path.  floor(x*num/10000) equals the reference's truncated double product for every
	// x in [0,65535] for 0.2126 and 0.0722; for 0.7152 = 447/625 the double product of a multiple of
	// 625 can land one ulp under the exact integer, so those inputs (1 in 625) take the reference's
	// own double multiply.  floor(x*num/10000) itself is one IMAD.HI: mulhi(x, ceil(num*2^32/10000))
	// is exact for x < 2^16 (the rounding-up of the constant adds < 2^-16, the nearest miss is 10^-4).
	// tests/test_convert_gpu.py checks all 3 x 65536 inputs on the device against the compiled
	// reference's outputs; tests/test_oracle.py checks the constants on the host.
	__device__ __forceinline__ uint32_t trunc_2126(uint32_t x) {
		return __umulhi(x, 913110048u); // ceil(2126 * 2^32 / 10000)
	}
	__device__ __forceinline__ uint32_t trunc_0722(uint32_t x) {
		return __umulhi(x, 310096639u); // ceil(722 * 2^32 / 10000)
	}
	template<bool Wide>
	__device__ __forceinline__ uint32_t trunc_7152(uint32_t x) {
		uint32_t t = __umulhi(x, 3071760611u); // ceil(7152 * 2^32 / 10000)
		if constexpr(Wide) {
			if(x % 625u == 0u) t = static_cast<uint32_t>(0.7152 * static_cast<double>(x));
		}
		return t;
	}

	// unorm -> float, bit-identical to IEEE (float)v / (float)M for every v in [0,M] (M = 255 or
	// 65535): one multiply by the rounded reciprocal plus one Newton correction.  Checked
	// exhaustively on the host (tests/test_oracle.py::test_division_forms) and on the device.
	template<int M>
	__device__ __forceinline__ float unorm_to_f32(uint32_t v) {
		constexpr float fm = static_cast<float>(M);
		constexpr float r = 1.0f / fm;
		const float fv = static_cast<float>(v);
		const float q = fv * r;
		const float e = __fmaf_rn(-q, fm, fv);
		return __fmaf_rn(e, r, q);
	}
	// RNE_f16((float)v / (float)M): the single multiply already rounds to the same binary16 for
	// every v (exhaustively checked), so the correction step is skipped.
	template<int M>
	__device__ __forceinline__ uint32_t unorm_to_f16(uint32_t v) {
		constexpr float r = 1.0f / static_cast<float>(M);
		return f32_to_f16_bits(static_cast<float>(v) * r);
	}

	// float -> unorm: (x != x) ? 0 : uintN(min(max(x,0),1)*M + 0.5f)
	template<int M>
	__device__ __forceinline__ uint32_t f32_to_unorm(float x) {
		if(x != x) return 0u;
		const float c = fminf(fmaxf(x, 0.0f), 1.0f);
		return __float2uint_rz(__fadd_rn(__fmul_rn(c, static_cast<float>(M)), 0.5f));
	}

	// Row mirror folded into a conversion's store addressing.  The stream is cut into
	// equal units (tiles, lane runs or single pixels) that never straddle a row; unit k of row y of
	// image i is stored where unit k of row rows-1-y of the same image would go.  per_row == 0: identity.
	struct FlipMap {
		uint32_t per_row = 0; // units per image row
		uint32_t rows = 0;    // rows per image
		__device__ __forceinline__ uint64_t operator()(uint64_t idx) const {
			if(per_row == 0) return idx;
			if(idx <= 0xFFFFFFFFull) {
				const uint32_t i32 = static_cast<uint32_t>(idx);
				const uint32_t row = i32 / per_row, k = i32 - row * per_row;
				const uint32_t img = row / rows, y = row - img * rows;
				return static_cast<uint64_t>(img * rows + (rows - 1u - y)) * per_row + k;
			}
			const uint64_t row = idx / per_row, k = idx - row * per_row;
			const uint64_t img = row / rows, y = row - img * rows;
			return (img * rows + (rows - 1u - y)) * per_row + k;
		}
	};

	struct PixelCtx {
		const uint8_t* lut;   // 65536-entry table in global memory (edge kernel); only read for U16 -> U8
		const uint16_t* lut2; // 4096-entry two-level form of the same table in shared memory (tile kernel)
		const uint8_t* lanes; // this lane's column of the bank-private form in shared memory (sector kernel)
		uint32_t gray16_cap;  // 255 in REF_EXACT mode (Layout.cpp:75), 65535 in FIXED mode
	};

	// Which form of the 16->8 table a kernel reads.
	enum : int { kLutDirect = 0, kLutTwoLevel = 1, kLutLanes = 2 };

	// The bank-private form (kLutLanes).  A warp's 32 independent 16-bit indices fall on the 32 shared-memory
	// banks like balls into bins -- 3.5 data-pipe passes per LDS whatever the table's size (ncu: 62.8 M
	// conflict wavefronts for 25.2 M gathers).  The only layout without conflicts gives every lane its own
	// bank: word (entry * 32 + lane).  That is 128 bytes per entry, so the table has to be SHORT, and it is
	// when it is indexed by the logarithm of the input: the table is an sRGB-style power curve, whose steps
	// thin out as the input grows (19 inputs apart at the dark end, 580 at the bright end).  float(v + 2048)
	// has its exponent and top 7 mantissa bits in its upper half-word: 128 segments per octave, 16 inputs wide
	// in the lowest octave and 512 in the highest, 644 segments in all for v in [0, 65535], and none of them
	// holds more than one step of the table (checked when the form is built, cabi.cu).  The low half-word of
	// the float is the position inside the segment, so with entry = (level << 16) + (0x10000 - step position)
	// - (segment's upper half-word << 16) the sum  bits + entry  carries into bit 16 exactly when v is at or
	// past the step, and bits 16..23 of the sum are the table's output.  Per sample: PRMT + FADD (v + 2048 as a
	// float, exact), shift + scaled add (address), LDS (one pass), IADD -- no compare, no select.
	constexpr int kLanesBias = 2048;
	constexpr int kLanesEntries = 644;
	constexpr uint32_t kLanesHiBase = (127u + 11u) << 7; // upper half-word of float(2048)
	constexpr int kLanesBytes = kLanesEntries * 128;
#ifndef CACAO_LANES_THREADS
#define CACAO_LANES_THREADS 512
#endif
	constexpr int kLanesThreads = CACAO_LANES_THREADS; // two blocks per SM

	// The reference's 16->8 table (Colordepth.cpp:35) is monotone, moves in unit steps, and its steps
	// are never closer than 19 inputs apart -- so a run of 16 consecutive inputs holds at most one
	// step.  Entry k of the two-level form describes inputs [16k, 16k+16): low byte = level at 16k,
	// high byte = offset of the step inside the run (16 = no step).  8 KiB instead of 64 KiB: it loads
	// 8x faster, leaves shared memory for 8 blocks per SM, and costs the same single LDS per sample.
	// cabi.cu builds it from the 65536-entry table and verifies it reproduces every entry.
	// 0x4B00vvvv is the float 2^23 + v.  The constant lives in constant memory so that it reaches PRMT as a
	// REGISTER operand and the byte selector as the immediate: with a literal, ptxas makes the literal the
	// immediate and spends two more instructions on masking the sample and materialising the selector.
	__constant__ uint32_t g_f32_2p23 = 0x4B000000u;
	// word: two packed 16-bit samples; half: which one
	__device__ __forceinline__ uint32_t lanes_magic(uint32_t word, int half) {
		return __byte_perm(word, g_f32_2p23, half ? 0x7632u : 0x7610u);
	}
	// bits 16..23 of the result are table[v]; magic = lanes_magic(...)
	__device__ __forceinline__ uint32_t lanes_sum(const PixelCtx& ctx, uint32_t magic) {
		const uint32_t bits = __float_as_uint(__fadd_rn(__uint_as_float(magic), static_cast<float>(kLanesBias) - 8388608.0f)); // float(v + bias), exact
		return bits + *reinterpret_cast<const uint32_t*>(ctx.lanes + (((bits >> 16) - kLanesHiBase) << 7)); // the constant folds into the LDS offset
	}

	template<int Mode>
	__device__ __forceinline__ uint32_t lut_16to8(const PixelCtx& ctx, uint32_t v) {
		if constexpr(Mode == kLutTwoLevel) {
			const uint32_t e = ctx.lut2[v >> 4];
			return (e & 0xFFu) + (((v & 15u) >= (e >> 8)) ? 1u : 0u);
		} else if constexpr(Mode == kLutLanes) {
			return lanes_sum(ctx, lanes_magic(v, 0)) >> 16;
		} else {
			return ctx.lut[v];
		}
	}

	// in[]: SC source samples; integer formats as their integer value, F16 as raw bits, F32 as
	// float bits.  out[]: DC destination samples in the same register convention.
	template<int SF, int DF, int SC, int DC, int LutMode = kLutDirect>
	__device__ __forceinline__ void convert_pixel(const uint32_t (&in)[SC], uint32_t (&out)[DC], const PixelCtx& ctx) {
		constexpr bool src_int = FmtTraits<SF>::is_int;
		if constexpr(SF == CACAO_FMT_U16 && DF == CACAO_FMT_U8 && LutMode == kLutLanes) {
			// ---- 16 -> 8 bit through the bank-private table form.  Register convention of THIS branch only:
			// in[] = lanes_magic() of the packed source word (the sample as the float 2^23 + v), out[] = the
			// result in bits 16..23, other bits unspecified -- the sector kernel builds and consumes both with
			// one PRMT each.  Same semantics as the two branches below (depth first, then the 8-bit layout rules).
			constexpr int used = (SC == DC) ? SC : (SC < 3 ? SC : 3);
			uint32_t t[used];
#pragma unroll
			for(int c = 0; c < used; ++c) t[c] = lanes_sum(ctx, in[c]);
			if constexpr(SC == DC) {
#pragma unroll
				for(int c = 0; c < DC; ++c) out[c] = t[c];
			} else if constexpr(DC == 1) {
				out[0] = min(trunc_2126(t[0] >> 16) + trunc_7152<false>(t[1] >> 16) + trunc_0722(t[2] >> 16), 255u) << 16;
			} else {
#pragma unroll
				for(int c = 0; c < 3; ++c) out[c] = t[SC == 1 ? 0 : c];
				if constexpr(DC == 4) out[3] = 255u << 16;
			}
		} else if constexpr(SF == CACAO_FMT_U16 && DF == CACAO_FMT_U8 && SC != DC) {
			// ---- 16 -> 8 bit WITH a layout change: the one pair where the depth conversion comes first --
			// the order of the engine's composition (Convert16To8BitColor, then ChangeChannelLayout:
			// engine/src/core/Cubemap.cpp:28-31), so the fused call equals the reference's two calls for every
			// layout pair: an added alpha is 255 (Layout.cpp:57-58), not table[65535] = 254, and gray is the
			// 8-bit rule on table outputs (Layout.cpp:33-41).  Only the channels that survive are looked up.
			constexpr int used = SC < 3 ? SC : 3;
			uint32_t t[used];
#pragma unroll
			for(int c = 0; c < used; ++c) t[c] = lut_16to8<LutMode>(ctx, in[c]);
			if constexpr(DC == 1) {
				out[0] = min(trunc_2126(t[0]) + trunc_7152<false>(t[1]) + trunc_0722(t[2]), 255u);
			} else {
#pragma unroll
				for(int c = 0; c < 3; ++c) out[c] = t[SC == 1 ? 0 : c];
				if constexpr(DC == 4) out[3] = 255u;
			}
		} else if constexpr(src_int) {
			// ---- step 1: layout change in the integer domain
			uint32_t mid[DC];
			if constexpr(SC == DC) {
#pragma unroll
				for(int c = 0; c < DC; ++c) mid[c] = in[c];
			} else if constexpr(DC == 1) {
				constexpr bool wide = (SF == CACAO_FMT_U16);
				const uint32_t sum = trunc_2126(in[0]) + trunc_7152<wide>(in[1]) + trunc_0722(in[2]);
				mid[0] = min(sum, wide ? ctx.gray16_cap : 255u);
			} else if constexpr(SC == 1) {
				mid[0] = mid[1] = mid[2] = in[0];
				if constexpr(DC == 4) mid[3] = (SF == CACAO_FMT_U8) ? 255u : 65535u;
			} else {
				mid[0] = in[0];
				mid[1] = in[1];
				mid[2] = in[2];
				if constexpr(DC == 4) mid[3] = (SF == CACAO_FMT_U8) ? 255u : 65535u;
			}
			// ---- step 2: sample-format change
			constexpr int M = (SF == CACAO_FMT_U8) ? 255 : 65535;
#pragma unroll
			for(int c = 0; c < DC; ++c) {
				if constexpr(DF == SF) {
					out[c] = mid[c];
				} else if constexpr(DF == CACAO_FMT_U8) {
					out[c] = lut_16to8<LutMode>(ctx, mid[c]);
				} else if constexpr(DF == CACAO_FMT_U16) {
					out[c] = mid[c] * 257u;
				} else if constexpr(DF == CACAO_FMT_F32) {
					out[c] = __float_as_uint(unorm_to_f32<M>(mid[c]));
				} else {
					out[c] = unorm_to_f16<M>(mid[c]);
				}
			}
		} else {
			// ---- float source: widen, layout change in fp32, narrow / quantise
			float f[SC];
#pragma unroll
			for(int c = 0; c < SC; ++c) f[c] = (SF == CACAO_FMT_F32) ? __uint_as_float(in[c]) : f16_bits_to_f32(static_cast<uint16_t>(in[c]));
			float mid[DC];
			if constexpr(SC == DC) {
#pragma unroll
				for(int c = 0; c < DC; ++c) mid[c] = f[c];
			} else if constexpr(DC == 1) {
				mid[0] = __fmaf_rn(0.0722f, f[2], __fmaf_rn(0.7152f, f[1], __fmul_rn(0.2126f, f[0])));
			} else if constexpr(SC == 1) {
				mid[0] = mid[1] = mid[2] = f[0];
				if constexpr(DC == 4) mid[3] = 1.0f;
			} else {
				mid[0] = f[0];
				mid[1] = f[1];
				mid[2] = f[2];
				if constexpr(DC == 4) mid[3] = 1.0f;
			}
#pragma unroll
			for(int c = 0; c < DC; ++c) {
				if constexpr(DF == CACAO_FMT_F32)
					out[c] = __float_as_uint(mid[c]);
				else if constexpr(DF == CACAO_FMT_F16)
					out[c] = f32_to_f16_bits(mid[c]);
				else if constexpr(DF == CACAO_FMT_U16)
					out[c] = f32_to_unorm<65535>(mid[c]);
				else
					out[c] = f32_to_unorm<255>(mid[c]);
			}
		}
	}

} // namespace cacao
