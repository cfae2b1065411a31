// convert_tiles.cuh -- the tiled, fully coalesced conversion kernel (one template, 132 shapes).
//
// A conversion is a flat 1-D stream: rows are tightly packed (libs/image/src/Layout.cpp:18-19), so
// an image -- and a contiguous batch of images -- is just `pixels` pixels in a row.
//
// Shape of the kernel (HBM-bound; no reuse, so no tiling for reuse -- smem is used only to
// re-shape the warp's data so BOTH global sides stay 16-byte vectorised and fully coalesced even
// for 3-, 6- and 12-byte pixels):
//
//   group  = G pixels, the smallest run whose source AND destination sizes are multiples of 16 B
//            (RGB8->RGBA8: G=16, 48 B -> 64 B;  RGBA8->RGBA16F: G=4, 16 B -> 32 B)
//   tile   = 32 groups = one warp's unit of work: NIN*512 B in, NOUT*512 B out
//   load   : NIN coalesced 16-B loads per lane (lane L, load k -> tile unit k*32+L)
//   shuffle: through the warp's private smem stage so lane L holds ITS group's NIN units
//   math   : G pixels per lane, entirely in registers, all indices compile-time
//   shuffle: back through the stage so lane L holds tile units k*32+L of the output
//   store  : NOUT coalesced 16-B stores per lane
//
// The stage pads each lane-row to an odd number of 16-B units, which makes the strided side of
// both transposes bank-conflict-free; the linear side then sees at most a 2-way conflict on one
// quarter-warp phase.  No __syncthreads anywhere: warps are independent (only __syncwarp).
#pragma once

#include <cstdlib>

#include "convert_pixel.cuh"

namespace cacao {

	template<int SF, int DF, int SC, int DC>
	struct TileShape {
		static constexpr int sb = FmtTraits<SF>::bytes, db = FmtTraits<DF>::bytes;
		static constexpr int sbpp = SC * sb, dbpp = DC * db;
		static constexpr int G = clcm(16 / cgcd(16, sbpp), 16 / cgcd(16, dbpp)); // pixels per group
		static constexpr int NIN = G * sbpp / 16;                                // 16-B units per group, source side
		static constexpr int NOUT = G * dbpp / 16;                               // ... destination side
		static constexpr int PIN = (NIN % 2 == 0) ? NIN + 1 : NIN;               // padded row pitches (odd)
		static constexpr int POUT = (NOUT % 2 == 0) ? NOUT + 1 : NOUT;
		static constexpr int STAGE_UNITS = 32 * ((NIN > 1 ? PIN : 0) > (NOUT > 1 ? POUT : 0) ? (NIN > 1 ? PIN : 0) : (NOUT > 1 ? POUT : 0));
		static constexpr bool needs_lut = (SF == CACAO_FMT_U16 && DF == CACAO_FMT_U8);
		// tiles in flight per warp iteration: keep >= 4 outstanding 16-B loads per lane
		static constexpr int U = NIN >= 4 ? 1 : (NIN >= 2 ? 2 : 4);
		static constexpr uint64_t tile_pixels = 32ull * G;
	};

	constexpr int kLut2Bytes = 4096 * 2;

	// sample k of a register-resident run of packed samples
	template<int Bytes>
	__device__ __forceinline__ uint32_t unpack_sample(const uint32_t* w, int k) {
		if constexpr(Bytes == 1)
			return (w[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
		else if constexpr(Bytes == 2)
			return (w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
		else
			return w[k];
	}
	template<int Bytes>
	__device__ __forceinline__ void pack_sample(uint32_t* w, int k, uint32_t v) {
		if constexpr(Bytes == 1)
			w[k >> 2] |= v << ((k & 3) * 8);
		else if constexpr(Bytes == 2)
			w[k >> 1] |= v << ((k & 1) * 16);
		else
			w[k] = v;
	}

	// FLIP: a separate instantiation, so that the plain stream kernel carries none of the index mapping
	template<int SF, int DF, int SC, int DC, bool FLIP = false>
	__global__ void __launch_bounds__(256) convert_tiles_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t n_tiles, uint32_t iters, const uint16_t* __restrict__ lut2_global, uint32_t gray16_cap, FlipMap flip) {
		using S = TileShape<SF, DF, SC, DC>;
		extern __shared__ __align__(16) uint8_t smem_raw[];

		const int lane = threadIdx.x & 31;
		const int warp = threadIdx.x >> 5;
		const int warps_per_block = blockDim.x >> 5;

		// shared-memory carve-up: [two-level 16->8 table, 8 KiB, only for U16->U8][per-warp stage]
		PixelCtx ctx;
		ctx.gray16_cap = gray16_cap;
		ctx.lut = nullptr;
		ctx.lut2 = nullptr;
		ctx.lanes = nullptr;
		uint8_t* stage_base = smem_raw;
		pdl_launch_dependents(); // cacao_device.cuh: the table (the library's own constant data) is staged before pdl_wait()
		if constexpr(S::needs_lut) {
			uint4* l4 = reinterpret_cast<uint4*>(smem_raw);
			const uint4* g4 = reinterpret_cast<const uint4*>(lut2_global);
			for(int i = threadIdx.x; i < kLut2Bytes / 16; i += blockDim.x) l4[i] = __ldg(g4 + i);
			__syncthreads();
			ctx.lut2 = reinterpret_cast<const uint16_t*>(smem_raw);
			stage_base = smem_raw + kLut2Bytes;
		}
		uint4* stage = reinterpret_cast<uint4*>(stage_base) + static_cast<size_t>(warp) * S::STAGE_UNITS;

		// A block owns a contiguous run of warps_per_block * U * iters tiles and walks it front to
		// back, so at any instant the whole grid touches one compact, advancing window of memory --
		// the access pattern of a plain one-shot elementwise kernel.  (Measured on B200: a classic
		// persistent grid-stride loop, where each warp's next tile is a whole grid away, loses 15%.)
		const uint64_t block_base = static_cast<uint64_t>(blockIdx.x) * warps_per_block * S::U * iters;
		pdl_wait(); // first touch of the caller's buffers
		for(uint32_t it = 0; it < iters; ++it) {
			const uint64_t t0 = block_base + (static_cast<uint64_t>(it) * warps_per_block + warp) * S::U;
			if(t0 >= n_tiles) break;
			// ---- coalesced loads for up to U tiles, all issued before any is consumed
			uint4 r[S::U][S::NIN];
#pragma unroll
			for(int u = 0; u < S::U; ++u) {
				if(t0 + u < n_tiles) {
					const uint4* tsrc = src + (t0 + u) * (32ull * S::NIN);
#pragma unroll
					for(int k = 0; k < S::NIN; ++k) r[u][k] = ldg_stream(tsrc + k * 32 + lane);
				}
			}
#pragma unroll
			for(int u = 0; u < S::U; ++u) {
				if(t0 + u >= n_tiles) break;
				// ---- linear -> per-lane groups
				if constexpr(S::NIN > 1) {
#pragma unroll
					for(int k = 0; k < S::NIN; ++k) {
						const int v = k * 32 + lane;
						stage[v + (S::PIN - S::NIN) * (v / S::NIN)] = r[u][k];
					}
					__syncwarp();
#pragma unroll
					for(int k = 0; k < S::NIN; ++k) r[u][k] = stage[lane * S::PIN + k];
					__syncwarp();
				}
				// ---- G pixels in registers
				uint32_t o[S::NOUT * 4];
#pragma unroll
				for(int i = 0; i < S::NOUT * 4; ++i) o[i] = 0u;
				uint32_t w[S::NIN * 4];
#pragma unroll
				for(int k = 0; k < S::NIN; ++k) {
					w[4 * k] = r[u][k].x;
					w[4 * k + 1] = r[u][k].y;
					w[4 * k + 2] = r[u][k].z;
					w[4 * k + 3] = r[u][k].w;
				}
#pragma unroll
				for(int p = 0; p < S::G; ++p) {
					uint32_t in[SC], out[DC];
#pragma unroll
					for(int c = 0; c < SC; ++c) in[c] = unpack_sample<S::sb>(w, p * SC + c);
					convert_pixel<SF, DF, SC, DC, kLutTwoLevel>(in, out, ctx);
#pragma unroll
					for(int c = 0; c < DC; ++c) pack_sample<S::db>(o, p * DC + c, out[c]);
				}
				uint4 ov[S::NOUT];
#pragma unroll
				for(int k = 0; k < S::NOUT; ++k) ov[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
				// ---- per-lane groups -> linear
				if constexpr(S::NOUT > 1) {
#pragma unroll
					for(int k = 0; k < S::NOUT; ++k) stage[lane * S::POUT + k] = ov[k];
					__syncwarp();
#pragma unroll
					for(int k = 0; k < S::NOUT; ++k) {
						const int v = k * 32 + lane;
						ov[k] = stage[v + (S::POUT - S::NOUT) * (v / S::NOUT)];
					}
					__syncwarp();
				}
				// ---- coalesced stores
				uint4* tdst = dst + (FLIP ? flip(t0 + u) : t0 + u) * (32ull * S::NOUT);
#pragma unroll
				for(int k = 0; k < S::NOUT; ++k) stg_stream(tdst + k * 32 + lane, ov[k]);
			}
		}
	}

	// Host-side launcher for one shape.  Returns the cudaError_t of the launch.
	struct TileLaunch {
		int sm_count;
		size_t max_smem_optin;
	};

	template<int SF, int DF, int SC, int DC>
	cudaError_t launch_convert_tiles(const void* src, void* dst, uint64_t n_tiles, const uint16_t* lut2_global, uint32_t gray16_cap, const TileLaunch& dev, cudaStream_t stream, FlipMap flip = FlipMap()) {
		using S = TileShape<SF, DF, SC, DC>;
		auto kern = flip.per_row ? convert_tiles_kernel<SF, DF, SC, DC, true> : convert_tiles_kernel<SF, DF, SC, DC, false>;
		const size_t stage_per_warp = static_cast<size_t>(S::STAGE_UNITS) * 16;
		const size_t lut_bytes = S::needs_lut ? kLut2Bytes : 0;
		const int warps = 8;
		const size_t smem = lut_bytes + stage_per_warp * warps;
		if(smem > 48 * 1024) {
			cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
			if(e != cudaSuccess) return e;
		}
		// tiles per block: one pass for plain shapes; LUT shapes amortise their 64 KiB table load over
		// several passes
		// Two float shapes whose blocks write little per pass measure ahead with four passes per block (RGBA32F -> Gray8
		// +8-12 %, RGB16F -> RGB8 +7-10 %: dev/sweep_to_gray.py all, profiles/r2_to_gray_passes_per_block.log)
		constexpr bool four_passes = DF == CACAO_FMT_U8 && ((SF == CACAO_FMT_F32 && SC == 4 && DC == 1) || (SF == CACAO_FMT_F16 && SC == 3 && DC == 3));
		uint32_t iters = (S::needs_lut || four_passes) ? 4u : 1u;
		if(const char* e = getenv("CACAO_TILE_ITERS")) iters = static_cast<uint32_t>(atoi(e) > 0 ? atoi(e) : 1);
		const uint64_t tiles_per_block = static_cast<uint64_t>(warps) * S::U * iters;
		const uint64_t blocks = (n_tiles + tiles_per_block - 1) / tiles_per_block;
		if(blocks == 0) return cudaSuccess;
		if(blocks > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
		(void)dev;
		return launch_pdl(kern, dim3(static_cast<unsigned>(blocks)), dim3(warps * 32), smem, stream, static_cast<const uint4*>(src), static_cast<uint4*>(dst), n_tiles, iters, lut2_global, gray16_cap, flip);
	}

} // namespace cacao
