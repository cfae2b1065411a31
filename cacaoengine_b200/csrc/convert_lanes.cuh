// convert_lanes.cuh -- the 16 -> 8 bit conversions (Colordepth.cpp:26-37, alone or fused with a layout change:
// engine/src/core/Cubemap.cpp:28-31) on the bank-private form of the depth table (convert_pixel.cuh, kLutLanes).
//
// The other two kernels keep the table in shared memory as an ordinary array, and a warp's 32 independent
// 16-bit indices cost 3.5 data-pipe passes per gather (ncu: 62.8 M conflict wavefronts for 25.2 M gathers) --
// the limiter of every table shape on uniform-random samples.  Here each lane reads its own bank, so a gather
// is one pass whatever the data, at 80.5 KiB of shared memory per block.  What the kernel is built around:
//
//   * two blocks per SM, each staging the 644-word table by itself (eight 16-byte stores per entry) -- before
//     pdl_wait(), i.e. under the tail of the previous launch on the stream when there is one (cacao_device.cuh);
//   * a lane owns a run of G pixels = KIN (2 or 3) whole 32-byte source sectors (LDG.256) and a whole number of
//     8-byte output units.  128-byte runs (RGBA16 at 16 pixels) stream 5 % slower on B200 (dev/stream_bench.cu)
//     and leave no registers to overlap passes, so RGBA16 runs are 8 pixels here, not 16 as in convert_wide.cuh;
//   * software pipelining: a sector's registers are re-loaded for the block's NEXT pass as soon as the last
//     pixel in them has been taken out, so a warp has loads in flight while it converts.  With ~30 instructions
//     per pixel and half the warps of a plain stream kernel, load -> wait -> convert -> store in turn leaves
//     too few bytes in flight (ncu on the unpipelined form: long-scoreboard 9-12 per issue, DRAM 63-80 %);
//   * outputs that are not whole sectors per lane (24 B for RGBA16 -> RGB8, 48 B for RGB16 -> RGB8) go through a
//     per-warp shared-memory stage and leave as coalesced 16-byte stores: lane-strided 8/16-byte stores would
//     reach L2 as partial sectors (ncu: 2x the sector writes of the bytes written).
#pragma once

#include <cstdlib>

#include "convert_pixel.cuh"
#include "convert_tiles.cuh"
#include "convert_wide.cuh"

namespace cacao {

	template<int SC, int DC>
	struct LanesShape {
		static constexpr int sbpp = SC * 2, dbpp = DC;
		static constexpr int G1 = clcm(32 / cgcd(32, sbpp), 8 / cgcd(8, dbpp));
		static constexpr int G = (G1 * sbpp / 32 == 1) ? 2 * G1 : G1; // pixels per lane run: at least two sectors in flight
		static constexpr int KIN = G * sbpp / 32;                     // 32-byte sectors per run
		static constexpr int OB = G * dbpp;                           // output bytes per run (a multiple of 8)
		static constexpr bool out_sectors = (OB % 32 == 0);
		static constexpr bool staged = !out_sectors && OB > 16;       // 24 or 48: transposed through shared memory
		static constexpr int threads = KIN == 2 ? 384 : 320;          // x 2 blocks per SM (B200 sweep, profiles/)
		static constexpr int stage_bytes = staged ? (threads / 32) * 32 * OB : 0;
		static constexpr int smem_bytes = kLanesBytes + stage_bytes;
		static_assert(KIN == 2 || KIN == 3, "runs are two or three sectors");
		static_assert(OB % 8 == 0 && OB <= 128, "runs end on 8-byte output units");
	};

	template<int SC, int DC, bool FLIP>
	__global__ void __launch_bounds__(LanesShape<SC, DC>::threads, 2)
		convert_lanes_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n_runs, uint32_t iters, const uint32_t* __restrict__ lanes_global, FlipMap flip) {
		using S = LanesShape<SC, DC>;
		constexpr int SF = CACAO_FMT_U16, DF = CACAO_FMT_U8;
		extern __shared__ __align__(16) uint8_t smem_raw[];
		const int lane = threadIdx.x & 31;
		// a block owns a contiguous range of blockDim * iters runs and walks it front to back
		const uint64_t pass_stride = blockDim.x;
		uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x * iters + threadIdx.x;
		uint32_t w[S::KIN * 8];
		auto issue = [&](int k, uint64_t run) {
			if(run < n_runs) ldg256_stream(src + run * (32ull * S::KIN) + 32 * k, &w[8 * k]);
		};
		pdl_launch_dependents();
		{
			// entry e becomes 32 equal words, one per bank: eight 16-byte stores.  The table is the library's own
			// constant data, so this happens BEFORE pdl_wait(): under the previous launch's tail when there is one.
			uint4* l4 = reinterpret_cast<uint4*>(smem_raw);
			for(int i = threadIdx.x; i < kLanesEntries * 8; i += blockDim.x) {
				const uint32_t e = __ldg(lanes_global + (i >> 3));
				l4[i] = make_uint4(e, e, e, e);
			}
		}
		pdl_wait();
#pragma unroll
		for(int k = 0; k < S::KIN; ++k) issue(k, r);
		__syncthreads();
		PixelCtx ctx;
		ctx.gray16_cap = 255u; // unused: 16 -> 8 bit takes the depth step first, gray is the 8-bit rule (convert_pixel.cuh)
		ctx.lut = nullptr;
		ctx.lut2 = nullptr;
		ctx.lanes = smem_raw + (lane << 2);
		uint8_t* stage = smem_raw + kLanesBytes + (threadIdx.x >> 5) * (32 * S::OB);
		// with the row mirror, a warp's 32 runs leave as one contiguous piece only if no warp straddles a row
		const bool warp_rows = !FLIP || (flip.per_row % 32u == 0u);

		for(uint32_t it = 0; it < iters; ++it, r += pass_stride) {
			const uint64_t warp_r0 = r - lane;
			if(warp_r0 >= n_runs) break;
			const bool mine = r < n_runs;
			const bool more = it + 1 < iters;
			uint32_t o[S::OB / 4];
			if(mine) {
#pragma unroll
				for(int i = 0; i < S::OB / 4; ++i) o[i] = 0u;
#pragma unroll
				for(int p = 0; p < S::G; ++p) {
					uint32_t in[SC], out[DC];
#pragma unroll
					for(int c = 0; c < SC; ++c) in[c] = lanes_magic(w[(p * SC + c) >> 1], (p * SC + c) & 1);
					convert_pixel<SF, DF, SC, DC, kLutLanes>(in, out, ctx);
#pragma unroll
					for(int c = 0; c < DC; ++c) {
						constexpr uint32_t put_byte2[4] = {0x3216u, 0x3260u, 0x3610u, 0x6210u};
						const int k = p * DC + c;
						o[k >> 2] = __byte_perm(o[k >> 2], out[c], put_byte2[k & 3]);
					}
					// sectors whose last pixel this was are free: fetch them for the next pass
#pragma unroll
					for(int k = 0; k < S::KIN; ++k)
						if((32 * (k + 1) - 1) / S::sbpp == p && more) issue(k, r + pass_stride);
				}
			}
			if constexpr(S::out_sectors) {
				if(mine) {
					uint8_t* q = dst + (FLIP ? flip(r) : r) * static_cast<uint64_t>(S::OB);
#pragma unroll
					for(int k = 0; k < S::OB / 32; ++k) stg256_stream(q + 32 * k, &o[8 * k]);
				}
			} else if constexpr(!S::staged) {
				// 8 or 16 bytes per lane: already contiguous across the warp
				if(mine) {
					uint8_t* q = dst + (FLIP ? flip(r) : r) * static_cast<uint64_t>(S::OB);
					if constexpr(S::OB == 16)
						stg_stream(q, make_uint4(o[0], o[1], o[2], o[3]));
					else
						asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(q), "r"(o[0]), "r"(o[1]) : "memory");
				}
			} else {
				if(warp_rows && warp_r0 + 32 <= n_runs) {
					// lane-major into the stage (odd strides in 8- / 16-byte units: conflict-free), unit-major out
					if constexpr(S::OB % 16 == 0) {
#pragma unroll
						for(int k = 0; k < S::OB / 16; ++k) *reinterpret_cast<uint4*>(stage + lane * S::OB + 16 * k) = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
					} else {
#pragma unroll
						for(int k = 0; k < S::OB / 8; ++k) *reinterpret_cast<uint2*>(stage + lane * S::OB + 8 * k) = make_uint2(o[2 * k], o[2 * k + 1]);
					}
					__syncwarp();
					uint8_t* q = dst + (FLIP ? flip(warp_r0) : warp_r0) * static_cast<uint64_t>(S::OB);
					constexpr int units = 32 * S::OB / 16;
#pragma unroll
					for(int k = 0; k < (units + 31) / 32; ++k) {
						const int u = k * 32 + lane;
						if(units % 32 == 0 || u < units) stg_stream(q + 16 * u, *reinterpret_cast<const uint4*>(stage + 16 * u));
					}
					__syncwarp();
				} else if(mine) {
					uint8_t* q = dst + (FLIP ? flip(r) : r) * static_cast<uint64_t>(S::OB);
#pragma unroll
					for(int k = 0; k < S::OB / 8; ++k) asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(q + 8 * k), "r"(o[2 * k]), "r"(o[2 * k + 1]) : "memory");
				}
			}
		}
	}

	// Which table shapes take this kernel by default (B200 sweep, profiles/r2_lut_forms_sweep.log); CACAO_LUT_MODE=
	// lanes|two|direct pins a form (test switch).  Returns false when the job stays on the other kernels.
	template<int SC, int DC>
	bool lanes_preferred(uint64_t n_runs) {
		if(const char* e = getenv("CACAO_LUT_MODE")) return e[0] == 'l';
		if(const char* e = getenv("CACAO_LUT_DIRECT"))
			if(atoi(e) != 0) return false;
		return n_runs >= 148ull * 2 * LanesShape<SC, DC>::threads;
	}

	template<int SC, int DC>
	cudaError_t launch_convert_lanes(const void* src, void* dst, uint64_t n_runs, const uint32_t* lanes_global, cudaStream_t stream, FlipMap flip = FlipMap()) {
		using S = LanesShape<SC, DC>;
		if(n_runs == 0) return cudaSuccess;
		auto kern = flip.per_row ? convert_lanes_kernel<SC, DC, true> : convert_lanes_kernel<SC, DC, false>;
		// a function attribute belongs to the device it was set on, and one process may drive several
		// (ConvertBatch(..., devices)): set it with every launch, as the other kernels above 48 KiB do
		if(cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::smem_bytes); e != cudaSuccess) return e;
		static const int resident = resident_blocks(kern, S::threads, S::smem_bytes); // per shape (the mirrored form has the same footprint); one device kind per process
		// passes per block: whole waves of long contiguous ranges; the shapes that write 96+ bytes per run measure
		// better with short ranges (profiles/r2_lut_forms_sweep.log)
		uint32_t iters = wave_aligned_iters(n_runs, S::threads, resident, S::OB >= 96 ? 8 : 32);
		if(const char* e = getenv("CACAO_TILE_ITERS")) iters = static_cast<uint32_t>(atoi(e) > 0 ? atoi(e) : 1);
		const uint64_t runs_per_block = static_cast<uint64_t>(S::threads) * iters;
		const uint64_t blocks = (n_runs + runs_per_block - 1) / runs_per_block;
		if(blocks > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
		return launch_pdl(kern, dim3(static_cast<unsigned>(blocks)), dim3(S::threads), S::smem_bytes, stream, static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), n_runs, iters, lanes_global, flip);
	}

} // namespace cacao
