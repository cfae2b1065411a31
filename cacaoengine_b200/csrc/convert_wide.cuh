// convert_wide.cuh -- conversion kernel built on sm_100's 256-bit global loads and stores.
//
// LDG.E.256 / STG.E.256 move one whole 32-byte DRAM sector per lane.  If a lane owns a run of G pixels
// whose source and destination sizes are both multiples of 32 bytes, it can load its run with KIN
// sector loads, convert in registers and store KOUT sectors -- every global access is a full,
// aligned sector, so nothing needs re-shaping through shared memory and no partial-sector write
// ever reaches L2.  Measured on B200 (dev/stream_bench.cu): lane-contiguous runs of 32..96 bytes
// stream at 6.9 TB/s, the same as a fully coalesced copy; 128 bytes costs 5 %, 256 bytes 25 %.
// So this kernel is used when both sides fit in <= 128 bytes per lane (RGBA8->RGBA16F: 32->64 B,
// RGB8->RGBA8: 96->128 B, RGBA16->RGBA8: 64->32 B, ...); everything else (e.g. RGB8->RGBA32F,
// 96->512 B) stays on the tile kernel in convert_tiles.cuh.  Where only the source side can be
// whole sectors within that budget (RGBA16->RGB8: 128 -> 48 B) the destination run is written with
// 16-byte stores instead: the minority of the traffic, and still contiguous per lane.
//
// Shared memory is used only by the U16->U8 shapes, for the 16->8 table.
#pragma once

#include <cstdlib>

#include "convert_pixel.cuh"
#include "convert_tiles.cuh"

namespace cacao {

	template<int SF, int DF, int SC, int DC>
	struct WideShape {
		static constexpr int sb = FmtTraits<SF>::bytes, db = FmtTraits<DF>::bytes;
		static constexpr int sbpp = SC * sb, dbpp = DC * db;
		// smallest run with whole sectors on the source side and whole 16-byte units on the destination
		// side; doubled when that makes the destination whole sectors too and still fits
		static constexpr int G1 = clcm(32 / cgcd(32, sbpp), 16 / cgcd(16, dbpp));
		static constexpr int K1 = G1 * sbpp / 32, N1 = G1 * dbpp / 16;
		static constexpr bool dbl = (N1 % 2 == 1) && (2 * K1 <= 4) && (2 * N1 <= 8);
		static constexpr int G = dbl ? 2 * G1 : G1;    // pixels per lane run
		static constexpr int KIN = G * sbpp / 32;      // 32-byte sectors per run, source
		static constexpr int NOUT16 = G * dbpp / 16;   // 16-byte units per run, destination
		static constexpr bool out_sectors = (NOUT16 % 2 == 0); // else: 16-byte stores (lane runs stay contiguous)
		static constexpr bool ok = KIN <= 4 && NOUT16 <= 8;
		// Where it pays (B200 sweep of all 132 shapes, profiles/r1_all_132_shapes*.csv): the table shapes
		// (shared memory is their bottleneck, and this kernel needs none for re-shaping: +15..30 %) and
		// multi-channel -> gray with narrow samples (large source runs, tiny outputs: +5..20 %).
		// Elsewhere the 128-bit tile kernel is equal or a few percent ahead and stays the default.
		// (+ Gray16F -> RGB8, the slowest shape of the tile kernel: 0.86 -> 0.95 here, dev/sweep_low_shapes.py)
		static constexpr bool preferred = ok && ((SF == CACAO_FMT_U16 && DF == CACAO_FMT_U8) || (DC == 1 && SC > 1 && sb <= 2) || (SF == CACAO_FMT_F16 && DF == CACAO_FMT_U8 && SC == 1 && DC == 3));
		static constexpr bool needs_lut = (SF == CACAO_FMT_U16 && DF == CACAO_FMT_U8);
		// runs per lane per pass: keep at least two sector loads in flight per lane
		static constexpr int U = KIN >= 2 ? 1 : 2;
	};

	// LUT_DIRECT: the 16->8 table sits in shared memory in full (64 KiB, one LDS.U8 per sample);
	// otherwise its 8 KiB two-level form (convert_pixel.cuh).
	template<int SF, int DF, int SC, int DC, bool LUT_DIRECT, bool FLIP = false>
	__global__ void __launch_bounds__(LUT_DIRECT ? 1024 : 256) convert_wide_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n_runs, uint32_t iters, const uint8_t* __restrict__ lut_global, const uint16_t* __restrict__ lut2_global, uint32_t gray16_cap, FlipMap flip) {
		using S = WideShape<SF, DF, SC, DC>;
		extern __shared__ __align__(16) uint8_t smem_raw[];
		PixelCtx ctx;
		ctx.gray16_cap = gray16_cap;
		ctx.lut = nullptr;
		ctx.lut2 = nullptr;
		ctx.lanes = nullptr;
		pdl_launch_dependents(); // cacao_device.cuh: the table (the library's own constant data) is staged before pdl_wait()
		if constexpr(S::needs_lut) {
			uint4* l4 = reinterpret_cast<uint4*>(smem_raw);
			const uint4* g4 = reinterpret_cast<const uint4*>(LUT_DIRECT ? static_cast<const void*>(lut_global) : static_cast<const void*>(lut2_global));
			constexpr int units = (LUT_DIRECT ? 65536 : kLut2Bytes) / 16;
			for(int i = threadIdx.x; i < units; i += blockDim.x) l4[i] = __ldg(g4 + i);
			__syncthreads();
			if constexpr(LUT_DIRECT)
				ctx.lut = smem_raw;
			else
				ctx.lut2 = reinterpret_cast<const uint16_t*>(smem_raw);
		}
		// a block owns a contiguous range of blockDim * U * iters runs and walks it front to back
		const uint64_t block_base = static_cast<uint64_t>(blockIdx.x) * blockDim.x * S::U * iters;
		pdl_wait(); // first touch of the caller's buffers
		for(uint32_t it = 0; it < iters; ++it) {
			const uint64_t r0 = block_base + (static_cast<uint64_t>(it) * blockDim.x + threadIdx.x) * S::U;
			if(r0 >= n_runs) break;
			uint32_t w[S::U][S::KIN * 8];
#pragma unroll
			for(int u = 0; u < S::U; ++u) {
				if(r0 + u < n_runs) {
					const uint8_t* p = src + (r0 + u) * (32ull * S::KIN);
#pragma unroll
					for(int k = 0; k < S::KIN; ++k) ldg256_stream(p + 32 * k, &w[u][8 * k]);
				}
			}
#pragma unroll
			for(int u = 0; u < S::U; ++u) {
				if(r0 + u >= n_runs) break;
				uint32_t o[S::NOUT16 * 4];
#pragma unroll
				for(int i = 0; i < S::NOUT16 * 4; ++i) o[i] = 0u;
#pragma unroll
				for(int p = 0; p < S::G; ++p) {
					uint32_t in[SC], out[DC];
#pragma unroll
					for(int c = 0; c < SC; ++c) in[c] = unpack_sample<S::sb>(w[u], p * SC + c);
					convert_pixel<SF, DF, SC, DC, LUT_DIRECT ? kLutDirect : kLutTwoLevel>(in, out, ctx);
#pragma unroll
					for(int c = 0; c < DC; ++c) pack_sample<S::db>(o, p * DC + c, out[c]);
				}
				uint8_t* q = dst + (FLIP ? flip(r0 + u) : r0 + u) * (16ull * S::NOUT16);
				if constexpr(S::out_sectors) {
#pragma unroll
					for(int k = 0; k < S::NOUT16 / 2; ++k) stg256_stream(q + 32 * k, &o[8 * k]);
				} else {
#pragma unroll
					for(int k = 0; k < S::NOUT16; ++k) stg_stream(q + 16 * k, make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
				}
			}
		}
	}

	// How many passes a block of a table shape makes over its contiguous range.  Every block stages the 8 KiB
	// table first, so short ranges pay for it again and again; but a grid that is not a whole number of waves
	// leaves SMs idle in the last one, and at the engine's sizes (six 2048^2 faces = 1.5 waves of 4-pass blocks)
	// that cost more than the table: 0.72 of the roofline against 0.88 for a 64-face batch of the same shape
	// (profiles/r2_lut_iters_by_size.log).  So: the fewest whole waves that keep a block at <= 8 passes, and the
	// passes that fill them exactly.  `resident` = blocks the device holds at once (occupancy x SMs).
	inline uint32_t wave_aligned_iters(uint64_t units, uint64_t units_per_pass, int resident, uint64_t max_passes = 8) {
		if(resident <= 0 || units_per_pass == 0) return 4;
		const uint64_t per_wave = units_per_pass * static_cast<uint64_t>(resident);
		const uint64_t waves = (units + per_wave * max_passes - 1) / (per_wave * max_passes);
		const uint64_t iters = (units + per_wave * waves - 1) / (per_wave * (waves ? waves : 1));
		return static_cast<uint32_t>(iters < 1 ? 1 : (iters > max_passes ? max_passes : iters));
	}

	template<class K>
	int resident_blocks(K kernel, int threads, size_t smem) {
		int dev = 0, sms = 0, per_sm = 0;
		if(cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
		if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess) return 0;
		return sms * per_sm;
	}

	template<int SF, int DF, int SC, int DC>
	cudaError_t launch_convert_wide(const void* src, void* dst, uint64_t n_runs, const uint8_t* lut_global, const uint16_t* lut2_global, uint32_t gray16_cap, cudaStream_t stream, FlipMap flip = FlipMap()) {
		using S = WideShape<SF, DF, SC, DC>;
		// table shapes: the 8 KiB two-level form is the default.  The full 64 KiB table (fewest
		// instructions per sample, but 1024-thread blocks and 3 blocks/SM) is kept selectable: on
		// uniform-random samples it is 2-15 % behind (RGBA16->RGBA8 5.84 vs 5.93 TB/s, Gray16->Gray8
		// 4.47 vs 5.23), on smooth images about level (dev/sweep_lut.py).
		bool direct = false;
		if(const char* e = getenv("CACAO_LUT_DIRECT")) direct = S::needs_lut && atoi(e) != 0 && !flip.per_row; // test switch; no flipped form
		if(const char* e = getenv("CACAO_LUT_MODE")) direct = S::needs_lut && e[0] == 'd' && !flip.per_row;   // two | direct | lanes (convert_lanes.cuh)
		const int threads = direct ? 1024 : 256;
		const size_t smem = S::needs_lut ? (direct ? 65536 : kLut2Bytes) : 0;
		// Four narrow channels -> one channel (RGBA8 -> Gray8, RGBA16 -> Gray16, ...): a block's pass writes so little that
		// two passes per block measure 3-15 % ahead of one (dev/sweep_to_gray.py, profiles/r2_to_gray_passes_per_block.log:
		// RGBA8 -> Gray8 0.87 -> 0.93 of the roofline at 64 Mpixel, RGBA16 -> Gray16 0.92 -> 0.98; RGB8 -> Gray8 +4 %); everything else
		// streams best with one.
		constexpr bool two_passes = !S::needs_lut && DC == 1 && ((SC == 4 && S::sb <= 2 && DF != CACAO_FMT_F32 && (FmtTraits<SF>::is_int || DF == SF)) || (SC == 3 && SF == CACAO_FMT_U8 && DF == CACAO_FMT_U8));
		uint32_t iters = S::needs_lut ? (direct ? 16u : 4u) : (two_passes ? 2u : 1u);
		if constexpr(S::needs_lut) {
			if(!direct) {
				static const int resident = resident_blocks(convert_wide_kernel<SF, DF, SC, DC, false>, 256, kLut2Bytes); // per shape; one device kind per process
				iters = wave_aligned_iters(n_runs, static_cast<uint64_t>(threads) * S::U, resident);
			}
		}
		if(const char* e = getenv("CACAO_TILE_ITERS")) iters = static_cast<uint32_t>(atoi(e) > 0 ? atoi(e) : 1);
		const uint64_t runs_per_block = static_cast<uint64_t>(threads) * S::U * iters;
		const uint64_t blocks = (n_runs + runs_per_block - 1) / runs_per_block;
		if(blocks == 0) return cudaSuccess;
		if(blocks > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
		if constexpr(S::needs_lut) {
			if(direct) {
				auto kern = convert_wide_kernel<SF, DF, SC, DC, true>;
				cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
				if(e != cudaSuccess) return e;
				return launch_pdl(kern, dim3(static_cast<unsigned>(blocks)), dim3(threads), smem, stream, static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), n_runs, iters, lut_global, lut2_global, gray16_cap, flip);
			}
		}
		if(flip.per_row)
			return launch_pdl(convert_wide_kernel<SF, DF, SC, DC, false, true>, dim3(static_cast<unsigned>(blocks)), dim3(threads), smem, stream, static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), n_runs, iters, lut_global, lut2_global, gray16_cap, flip);
		else
			return launch_pdl(convert_wide_kernel<SF, DF, SC, DC, false>, dim3(static_cast<unsigned>(blocks)), dim3(threads), smem, stream, static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), n_runs, iters, lut_global, lut2_global, gray16_cap, flip);
	}

} // namespace cacao
