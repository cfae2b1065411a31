// equirect_rows.cu -- equirect -> cubemap, the row-strip kernel: the form the launcher picks for one-channel
// and 8-bit texels (equirect_rows_preferred), selectable on every plain resampling for tests and A/B timing.
//
// The gather kernel of equirect.cu is bound by the L1 data pipe on 16-byte texels and by INSTRUCTION ISSUE on
// narrow ones: RGBA8, Gray8 and Gray32F all take ~0.1 ms for 25 Mpixel whatever their bytes, 76-79 % of the issue
// slots busy at full occupancy (profiles/r2_equirect_rows_vs_gather_ncu.txt).  This kernel computes the same
// pixels -- Spec B.2 bit for bit, the same four-fold mirror sharing, unit lists, flipped output -- with fewer
// instructions per pixel (DESIGN.md 4.2a has the accounting and the measured gains):
//   * floor() of the sample position and the truncation of the integer store use the FP32 pipe
//     (FADD.RM / FADD.RZ with 1.5 * 2^23 / 2^23: the integer lands in the mantissa) instead of the
//     quarter-rate conversion unit (F2I, FRND); integer results are rounded and truncated as packed pairs;
//   * tap and store addresses are one IMAD.WIDE each off 64-bit row pointers that the mirror column shares;
//   * the cube-face table is a handful of selects on the warp-uniform face number, not a jump table;
//   * one-channel texels filter the pixel and its mirror-column pixel together in packed fp32 (FFMA2);
//   * LOOP form (one-channel texels of up to two bytes): a thread keeps its column (and the mirror column) for
//     several row steps.  On the four side faces the longitude and sqrt(x^2 + z^2) depend on the column only
//     (SURVEY.md B.2 "structure worth exploiting"): they are computed once per thread and only the latitude --
//     one atan2 -- per row step.  The +-Y faces recompute everything per step.
// No reference counterpart (tools/cubetool/main.cpp:31-33 has only create / extract).
#include <algorithm>
#include <cstdlib>

#include "equirect_taps.cuh"

namespace cacao {

	namespace {

		constexpr float kFloorMagic = 12582912.0f; // 1.5 * 2^23: v + kFloorMagic rounded down holds floor(v) in its mantissa for -2^22 <= v < 2^22

		// floor(v) as an integer, v - floor(v) as the return value; -1 <= v < 2^22
		__device__ __forceinline__ float floor_split(float v, int& whole) {
			const float t = __fadd_rd(v, kFloorMagic);
			whole = __float_as_int(t) - __float_as_int(kFloorMagic);
			return v - (t - kFloorMagic); // t - magic is floor(v) exactly
		}

		// base + index * scale as ONE IMAD.WIDE.  Written in PTX on purpose: left to itself the compiler
		// re-associates (row * pitch + x * bpp) + src into four instructions per tap.
		__device__ __forceinline__ const uint8_t* at(const uint8_t* base, uint32_t index, uint32_t scale) {
			uint64_t r;
			asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(index), "r"(scale), "l"(reinterpret_cast<uint64_t>(base)));
			return reinterpret_cast<const uint8_t*>(r);
		}
		__device__ __forceinline__ uint8_t* at(uint8_t* base, uint32_t index, uint32_t scale) {
			return const_cast<uint8_t*>(at(static_cast<const uint8_t*>(base), index, scale));
		}
		__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
			uint64_t r;
			asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
			return r;
		}
		__device__ __forceinline__ uint64_t fadd2_rz(uint64_t a, uint64_t b) {
			uint64_t r;
			asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
			return r;
		}

		// The byte sizes the addresses scale by, as kernel ARGUMENTS: with a literal power of two ptxas turns
		// mad.wide into shift + high-shift + add + add-with-carry (four instructions per tap); a scale it
		// cannot see stays one IMAD.WIDE with the scale in a uniform register.
		struct Scales {
			uint32_t sbpp, dbpp;      // source / destination texel
			uint32_t pitch, out_pitch; // source row / face row
		};

		// the two source rows of an output row as pointers, and the row weight
		struct RowPtrs {
			const uint8_t* p0;
			const uint8_t* p1;
			float fy;
			uint32_t tail; // texels from the start of the lower row to the end of the buffer (saturated): aligned tap windows must stay inside
		};
		__device__ __forceinline__ RowPtrs row_ptrs(const uint8_t* __restrict__ src, float ys, const EquirectParams& prm, const Scales& sc) {
			RowPtrs t;
			int y0;
			t.fy = floor_split(ys, y0);
			int y1 = y0 + 1;
			const int H1 = static_cast<int>(prm.H) - 1, R1 = static_cast<int>(prm.src_rows) - 1, s0 = static_cast<int>(prm.src_row0);
			y0 = min(max(y0, 0), H1); // rows clamp
			y1 = min(max(y1, 0), H1);
			// the window was planned on the host with the same arithmetic; this clamp only guards memory
			const int r0 = min(max(y0 - s0, 0), R1), r1 = min(max(y1 - s0, 0), R1);
			t.p0 = at(src, static_cast<uint32_t>(r0), sc.pitch);
			t.p1 = at(src, static_cast<uint32_t>(r1), sc.pitch);
			t.tail = static_cast<uint32_t>(min(R1 - max(r0, r1) + 1, 16)) * prm.W;
			return t;
		}

		// the four taps of one output pixel as floats (integers keep their 2^23 until the lerp); returns fx
		template<int F, int CH>
		__device__ __forceinline__ float fetch_taps(const RowPtrs& rp, float xs, int W, uint32_t sbpp, float (&a)[CH], float (&b)[CH], float (&c)[CH], float (&d)[CH]) {
			constexpr bool kBiased = FmtTraits<F>::is_int;
			int x0;
			const float fx = floor_split(xs, x0);
			// columns wrap: xs lies in [-0.5, W - 0.5] (W <= 2^20: the rounding of the projection cannot reach W)
			int x1 = x0 + 1;
			if(x0 < 0) x0 = W - 1;
			if(x1 >= W) x1 = 0;
			if constexpr(CH == 3) {
				// both taps of a row in one aligned window (load_pair3): not across the seam, and not where the
				// window (at most 8 texels past the x0 tap) would leave the buffer
				if(x1 != 0 && static_cast<uint32_t>(x0) + 8u <= rp.tail) {
					load_pair3_at<F, kBiased>(at(rp.p0, static_cast<uint32_t>(x0), sbpp), a, b);
					load_pair3_at<F, kBiased>(at(rp.p1, static_cast<uint32_t>(x0), sbpp), c, d);
					return fx;
				}
			}
			load_texel<F, CH, kBiased>(at(rp.p0, static_cast<uint32_t>(x0), sbpp), a);
			load_texel<F, CH, kBiased>(at(rp.p0, static_cast<uint32_t>(x1), sbpp), b);
			load_texel<F, CH, kBiased>(at(rp.p1, static_cast<uint32_t>(x0), sbpp), c);
			load_texel<F, CH, kBiased>(at(rp.p1, static_cast<uint32_t>(x1), sbpp), d);
			return fx;
		}

		// packed bilinear filter of two values at once, each with its own weights
		template<bool BIASED>
		__device__ __forceinline__ uint64_t lerp2p(uint64_t vx, uint64_t vy, uint64_t A, uint64_t B, uint64_t C, uint64_t D) {
			const uint64_t m1 = pk2(-1.0f, -1.0f);
			const uint64_t dx0 = fma2(A, m1, B), dx1 = fma2(C, m1, D);
			if constexpr(BIASED) {
				const uint64_t one = pk2(1.0f, 1.0f), off = pk2(-kMagic, -kMagic);
				A = fma2(A, one, off);
				C = fma2(C, one, off);
			}
			const uint64_t top = fma2(vx, dx0, A);
			const uint64_t bot = fma2(vx, dx1, C);
			return fma2(vy, fma2(top, m1, bot), top);
		}

		// Integer results that are still packed pairs: + 0.5 and the truncation (round-toward-zero add of 2^23,
		// see store_texel FASTINT) on both lanes at once.  Returns the two 0x4B00vvvv words.
		__device__ __forceinline__ void quantise2(uint64_t v, uint32_t& lo, uint32_t& hi) {
			const uint64_t t = fadd2_rz(fadd2(v, pk2(0.5f, 0.5f)), pk2(kMagic, kMagic));
			asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(t));
		}

		// one output pixel
		template<int F, int CH>
		__device__ __forceinline__ void sample_one(const RowPtrs& rp, float xs, int W, uint32_t sbpp, uint8_t* __restrict__ out_px) {
			constexpr bool kBiased = FmtTraits<F>::is_int;
			float a[CH], b[CH], c[CH], d[CH];
			const float fx = fetch_taps<F, CH>(rp, xs, W, sbpp, a, b, c, d);
			if constexpr(CH == 4 && kBiased) {
				// four integer channels: filter, round and truncate as two packed pairs
				const uint64_t vx = pk2(fx, fx), vy = pk2(rp.fy, rp.fy);
				const uint64_t o01 = lerp2p<true>(vx, vy, pk2(a[0], a[1]), pk2(b[0], b[1]), pk2(c[0], c[1]), pk2(d[0], d[1]));
				const uint64_t o23 = lerp2p<true>(vx, vy, pk2(a[2], a[3]), pk2(b[2], b[3]), pk2(c[2], c[3]), pk2(d[2], d[3]));
				uint32_t s0, s1, s2, s3;
				quantise2(o01, s0, s1);
				quantise2(o23, s2, s3);
				if constexpr(F == CACAO_FMT_U8) {
					*reinterpret_cast<uint32_t*>(out_px) = __byte_perm(__byte_perm(s0, s1, 0x0040u), __byte_perm(s2, s3, 0x0040u), 0x5410u);
				} else {
					*reinterpret_cast<uint2*>(out_px) = make_uint2(__byte_perm(s0, s1, 0x5410u), __byte_perm(s2, s3, 0x5410u));
				}
			} else {
				float out[CH];
				constexpr int kPairs = CH / 2;
#pragma unroll
				for(int k = 0; k < kPairs; ++k) lerp2<kBiased>(fx, rp.fy, a + 2 * k, b + 2 * k, c + 2 * k, d + 2 * k, out + 2 * k);
#pragma unroll
				for(int k = 2 * kPairs; k < CH; ++k) {
					constexpr float kOff = kBiased ? kMagic : 0.0f;
					const float top = fmaf(fx, b[k] - a[k], a[k] - kOff);
					const float bot = fmaf(fx, d[k] - c[k], c[k] - kOff);
					out[k] = fmaf(rp.fy, bot - top, top);
				}
				store_texel<F, CH, false, true>(out_px, out);
			}
		}

		// an output pixel and its mirror-column pixel (same rows, two longitudes)
		template<int F, int CH>
		__device__ __forceinline__ void sample_pair(const RowPtrs& rp, float xs_i, float xs_m, int W, uint32_t sbpp, uint8_t* __restrict__ out_i, uint8_t* __restrict__ out_m, bool has_m) {
			if constexpr(CH == 1) {
				// one channel: the two pixels share the packed-fp32 lanes
				constexpr bool kBiased = FmtTraits<F>::is_int;
				float a0[1], b0[1], c0[1], d0[1], a1[1], b1[1], c1[1], d1[1];
				const float fx0 = fetch_taps<F, 1>(rp, xs_i, W, sbpp, a0, b0, c0, d0);
				const float fx1 = fetch_taps<F, 1>(rp, xs_m, W, sbpp, a1, b1, c1, d1);
				const uint64_t o = lerp2p<kBiased>(pk2(fx0, fx1), pk2(rp.fy, rp.fy), pk2(a0[0], a1[0]), pk2(b0[0], b1[0]), pk2(c0[0], c1[0]), pk2(d0[0], d1[0]));
				if constexpr(kBiased) {
					uint32_t s0, s1;
					quantise2(o, s0, s1);
					if constexpr(F == CACAO_FMT_U8) {
						out_i[0] = static_cast<uint8_t>(s0);
						if(has_m) out_m[0] = static_cast<uint8_t>(s1);
					} else {
						*reinterpret_cast<uint16_t*>(out_i) = static_cast<uint16_t>(s0);
						if(has_m) *reinterpret_cast<uint16_t*>(out_m) = static_cast<uint16_t>(s1);
					}
				} else {
					float r0[1], r1[1];
					asm("mov.b64 {%0,%1}, %2;" : "=f"(r0[0]), "=f"(r1[0]) : "l"(o));
					store_texel<F, 1, false, true>(out_i, r0);
					if(has_m) store_texel<F, 1, false, true>(out_m, r1);
				}
			} else {
				sample_one<F, CH>(rp, xs_i, W, sbpp, out_i);
				if(has_m) sample_one<F, CH>(rp, xs_m, W, sbpp, out_m);
			}
		}

		// One thread = column i and its mirror column N-1-i of one unit, plus -- for units flagged kUnitMirror --
		// the mirror rows N-1-j: four pixels from one projection, exactly as equirect_kernel shares them.
		// LOOP: the thread walks `steps` rows that lie a block's height apart (the block sweeps a compact patch
		// of the face per step) and keeps what depends on the column only -- on the four side faces both
		// longitudes and sqrt(x^2 + z^2) -- across the steps; !LOOP: one row per thread, nothing carried.
		// Unit lists with and without the mirror flag and flipped output (row_add / row_mul) all go through this one
		// form; the fused conversions (Spec B.4) stay with the gather kernel.
		template<int F, int CH, int WX, bool LOOP, int MINB>
		__global__ void __launch_bounds__(128, MINB) equirect_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, EquirectParams prm, Scales scl, uint32_t steps) {
			constexpr uint32_t RY = 32 / WX, BH = 4 * RY; // a warp covers WX x RY pixels per step, a block WX x BH
			pdl_launch_dependents(); // the first projection overlaps the previous launch's tail (cacao_device.cuh)
			const CubeUnit unit = prm.units[blockIdx.z];
			const uint32_t face = unit.face & 7u;
			const bool mirror = (unit.face & kUnitMirror) != 0;
			const uint32_t N = prm.N;
			const uint32_t i = blockIdx.x * WX + (threadIdx.x % WX);
			uint32_t j = unit.row_begin + blockIdx.y * (LOOP ? BH * steps : BH) + threadIdx.y * RY + threadIdx.x / WX;
			if(i >= ((N + 1u) >> 1) || j >= unit.row_end) return;
			const uint32_t im = N - 1u - i;
			const bool has_m = im != i;
			const int W = static_cast<int>(prm.W);
			const float sc = face_coord(i, prm.pc);
			const bool polar = (face & 6u) == 2u; // +-Y: both angles depend on row and column
			const bool sc_is_z = face < 2u;       // which component the column mirror negates: z on the +-X faces, x elsewhere
			// face row j is stored at row row_add + row_mul * j: (0, 1) normally, (N-1, -1) for flipped output
			const uint32_t frow = face * N + prm.row_add;
			float xs_a = 0.0f, xs_b = 0.0f, xs_c = 0.0f, xs_d = 0.0f, rxz = 1.0f;
			pdl_wait(); // before the first touch of the caller's buffers (later, inside the loop, it costs the 32-register forms a spill)
			for(uint32_t step = 0; step < (LOOP ? steps : 1u) && j < unit.row_end; ++step, j += BH) {
				const float tc = face_coord(j, prm.pc);
				float xa, ya, za;
				face_direction_sel(face, sc, tc, xa, ya, za);
				if(!LOOP || polar || step == 0) {
					// first-quadrant longitude shared by all four mirror images; they differ in the quadrant only
					const float rphi = atan2_core(fabsf(za), fabsf(xa));
					rxz = sqrtf(fmaf(xa, xa, za * za));
					const float xb = sc_is_z ? xa : -xa, zb = sc_is_z ? -za : za;
					xs_a = fmaf(atan2_quadrant(rphi, xa, za), prm.pc.kx, prm.pc.cx);
					xs_b = fmaf(atan2_quadrant(rphi, xb, zb), prm.pc.kx, prm.pc.cx);
					// the row mirror negates y on the side faces (longitudes stay) and z on the +-Y faces
					xs_c = polar ? fmaf(atan2_quadrant(rphi, xa, -za), prm.pc.kx, prm.pc.cx) : xs_a;
					xs_d = polar ? fmaf(atan2_quadrant(rphi, xb, -zb), prm.pc.kx, prm.pc.cx) : xs_b;
				}
				const float theta = atan2_spec(ya, rxz);
				const RowPtrs rt = row_ptrs(src, fmaf(-theta, prm.pc.ky, prm.pc.cy), prm, scl);
				uint8_t* orow = at(dst, frow + j * prm.row_mul, scl.out_pitch);
				sample_pair<F, CH>(rt, xs_a, xs_b, W, scl.sbpp, at(orow, i, scl.dbpp), at(orow, im, scl.dbpp), has_m);
				const uint32_t jm = N - 1u - j;
				if(mirror && jm != j) { // uniform per block: units are indexed by blockIdx.z
					const RowPtrs rm = polar ? rt : row_ptrs(src, fmaf(theta, prm.pc.ky, prm.pc.cy), prm, scl);
					uint8_t* mrow = at(dst, frow + jm * prm.row_mul, scl.out_pitch);
					sample_pair<F, CH>(rm, xs_c, xs_d, W, scl.sbpp, at(mrow, i, scl.dbpp), at(mrow, im, scl.dbpp), has_m);
				}
			}
		}

		// Which form serves a shape, from the B200 sweep (dev/sweep_rows.py, profiles/r2_equirect_rows_sweep.log):
		// one-channel texels of up to two bytes walk rows (the projection is most of their work, and the loop
		// amortises the column part of it; 56-62 registers, 8-10 blocks per SM); Gray32F, RGBA8 and RGB8 take one
		// row per thread at full occupancy.  Everything else stays with the gather kernel of equirect.cu.
		template<int F, int CH>
		struct RowsPlan {
			static constexpr bool loop = CH == 1 && FmtTraits<F>::bytes <= 2;
			static constexpr int minb = loop ? (F == CACAO_FMT_U8 ? 8 : 10) : (CH == 3 ? 12 : 16);
			static constexpr uint32_t max_steps = F == CACAO_FMT_U8 ? 8 : 4;
		};

		template<int F, int CH>
		cudaError_t launch_rows(const void* src, void* dst, const EquirectParams& prm, int sm_count, cudaStream_t stream) {
			constexpr int sbpp = FmtTraits<F>::bytes * CH, dbpp = sbpp; // same format and layout out as in
			// a warp's stores should fill whole 32-byte sectors; wider texels prefer squarer footprints
			constexpr int WX = sbpp >= 4 ? 8 : (sbpp >= 2 ? 16 : 32);
			constexpr uint32_t BH = 4 * (32 / WX);
			using Plan = RowsPlan<F, CH>;
			const uint32_t half = (prm.N + 1u) >> 1;
			uint32_t max_rows = 0;
			for(uint32_t u = 0; u < prm.n_units; ++u) max_rows = std::max(max_rows, prm.units[u].row_end - prm.units[u].row_begin);
			// row steps per thread (loop form): as many as leave about three full waves of blocks
			const uint64_t quads = static_cast<uint64_t>(half) * max_rows * prm.n_units;
			const uint64_t wave = static_cast<uint64_t>(sm_count > 0 ? sm_count : 148) * Plan::minb * 128;
			uint32_t steps = static_cast<uint32_t>(std::min<uint64_t>(Plan::max_steps, std::max<uint64_t>(1, quads / (3 * wave))));
			const char* se = getenv("CACAO_EQUIRECT_STEPS");
			if(se) steps = static_cast<uint32_t>(std::max(1, atoi(se)));
			const uint8_t* s8 = static_cast<const uint8_t*>(src);
			uint8_t* d8 = static_cast<uint8_t*>(dst);
			cudaError_t pe = cudaSuccess;
			const dim3 block(32, 4, 1);
			const Scales scl = {static_cast<uint32_t>(sbpp), static_cast<uint32_t>(dbpp), prm.W * static_cast<uint32_t>(sbpp), prm.N * static_cast<uint32_t>(dbpp)};
			auto grid_for = [&](uint32_t st) { return dim3((half + WX - 1) / WX, (max_rows + BH * st - 1) / (BH * st), prm.n_units); };
#ifdef CACAO_ROWS_SWEEP // development build: loop form and blocks per SM (registers per thread) selectable at run time
			{
				const char* mb = getenv("CACAO_EQUIRECT_MINB");
				const int m = mb ? atoi(mb) : 0;
#define CACAO_ROWS_CASE(MB)                                                                                                     \
	if(m == MB) {                                                                                                                \
		if(se)                                                                                                                   \
			pe = launch_pdl(equirect_rows_kernel<F, CH, WX, true, MB>, grid_for(steps), block, 0, stream, s8, d8, prm, scl, steps);   \
		else                                                                                                                     \
			pe = launch_pdl(equirect_rows_kernel<F, CH, WX, false, MB>, grid_for(1), block, 0, stream, s8, d8, prm, scl, 1);          \
		return pe;                                                                                                               \
	}
				CACAO_ROWS_CASE(8)
				CACAO_ROWS_CASE(10)
				CACAO_ROWS_CASE(12)
				CACAO_ROWS_CASE(16)
#undef CACAO_ROWS_CASE
			}
#endif
			if(!Plan::loop) steps = 1;
			pe = launch_pdl(equirect_rows_kernel<F, CH, WX, Plan::loop, Plan::minb>, grid_for(steps), block, 0, stream, s8, d8, prm, scl, steps);
			return pe;
		}

		template<int F>
		cudaError_t rows_plain(const void* src, void* dst, int ch, const EquirectParams& prm, int sm_count, cudaStream_t stream) {
			switch(ch) {
				case 1: return launch_rows<F, 1>(src, dst, prm, sm_count, stream);
				case 3: return launch_rows<F, 3>(src, dst, prm, sm_count, stream);
				case 4: return launch_rows<F, 4>(src, dst, prm, sm_count, stream);
			}
			return cudaErrorInvalidValue;
		}

	} // namespace

	// The row-strip kernel serves every plain resampling (same format and layout out as in) of panoramas up to
	// 2^20 pixels a side (the floor trick's range); equirect_rows_preferred says where it measured ahead of the
	// gather kernel.  The fused conversions stay with the gather kernel.
	bool equirect_rows_supported(int ch, int fmt, int dst_ch, int dst_fmt, uint32_t W, uint32_t H) {
		return W <= (1u << 20) && H <= (1u << 20) && dst_ch == ch && dst_fmt == fmt && (ch == 1 || ch == 3 || ch == 4) && fmt >= CACAO_FMT_U8 && fmt <= CACAO_FMT_F32;
	}
	bool equirect_rows_preferred(int ch, int fmt) {
		return ch == 1 || (fmt == CACAO_FMT_U8 && (ch == 3 || ch == 4));
	}

	cudaError_t equirect_rows_launch(const void* src, void* dst, int ch, int fmt, const EquirectParams& prm, int sm_count, cudaStream_t stream) {
		switch(fmt) {
			case CACAO_FMT_U8: return rows_plain<CACAO_FMT_U8>(src, dst, ch, prm, sm_count, stream);
			case CACAO_FMT_U16: return rows_plain<CACAO_FMT_U16>(src, dst, ch, prm, sm_count, stream);
			case CACAO_FMT_F16: return rows_plain<CACAO_FMT_F16>(src, dst, ch, prm, sm_count, stream);
			case CACAO_FMT_F32: return rows_plain<CACAO_FMT_F32>(src, dst, ch, prm, sm_count, stream);
		}
		return cudaErrorInvalidValue;
	}

} // namespace cacao
