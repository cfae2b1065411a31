// misc_kernels.cu -- flip, synthetic-input and checksum kernels.
#include "misc_kernels.cuh"

#include "cacao_device.cuh"

namespace cacao {

	namespace {

		// Vertical mirror.  A row is a contiguous run of `pitch` bytes; when pitch and both bases are
		// 16-byte aligned every row moves as coalesced 16-B vectors, otherwise byte by byte.
		__global__ void __launch_bounds__(256) flip_rows_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t units_per_row, uint32_t h) {
			const uint64_t total = units_per_row * h;
			const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
			for(uint64_t u = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < total; u += stride) {
				const uint64_t y = u / units_per_row, x = u - y * units_per_row;
				stg_stream(dst + u, ldg_stream(src + (h - 1 - y) * units_per_row + x));
			}
		}
		__global__ void __launch_bounds__(256) flip_rows_byte_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t pitch, uint32_t h) {
			const uint64_t total = pitch * h;
			const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
			for(uint64_t u = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < total; u += stride) {
				const uint64_t y = u / pitch, x = u - y * pitch;
				dst[u] = src[(h - 1 - y) * pitch + x];
			}
		}

		// 16-bit left shift on packed pairs: (w << s) with the bits that crossed the half-word border masked off
		__global__ void __launch_bounds__(256) shl16_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t units, uint32_t shift, uint32_t mask) {
			const uint64_t u = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
			if(u >= units) return;
			uint4 v = ldg_stream(src + u);
			v.x = (v.x << shift) & mask;
			v.y = (v.y << shift) & mask;
			v.z = (v.z << shift) & mask;
			v.w = (v.w << shift) & mask;
			stg_stream(dst + u, v);
		}
		__global__ void __launch_bounds__(256) shl16_scalar_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, uint64_t begin, uint64_t end, uint32_t shift) {
			const uint64_t k = begin + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
			if(k < end) dst[k] = static_cast<uint16_t>(static_cast<uint32_t>(src[k]) << shift);
		}

		__device__ __forceinline__ float synth_float(uint32_t hsh) {
			// (1 + m/65536) * 2^(e-8), e in [0,16): assembled from bits so host and device agree exactly
			const uint32_t e = (hsh >> 24) & 15u, m = (hsh >> 8) & 0xFFFFu;
			return __uint_as_float(((127u - 8u + e) << 23) | (m << 7));
		}

		template<int F>
		__global__ void __launch_bounds__(256) synth_kernel(typename FmtTraits<F>::type* __restrict__ dst, uint64_t first_sample, uint64_t samples, int ch, uint32_t seed) {
			const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
			for(uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < samples; k += stride) {
				const uint64_t idx = first_sample + k;
				const uint32_t hsh = hash32(seed, idx);
				const bool alpha = (ch == 4) && ((idx & 3u) == 3u);
				if constexpr(F == CACAO_FMT_U8)
					dst[k] = static_cast<uint8_t>(hsh);
				else if constexpr(F == CACAO_FMT_U16)
					dst[k] = static_cast<uint16_t>(hsh);
				else if constexpr(F == CACAO_FMT_F16)
					dst[k] = alpha ? static_cast<uint16_t>(0x3C00u) : f32_to_f16_bits(synth_float(hsh));
				else
					dst[k] = alpha ? 1.0f : synth_float(hsh);
			}
		}

		__global__ void __launch_bounds__(256) checksum_kernel(const uint32_t* __restrict__ src, uint64_t words, unsigned long long* __restrict__ accum) {
			unsigned long long local = 0;
			const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
			for(uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < words; k += stride) local += fmix32(src[k] + static_cast<uint32_t>(k) * 0x9E3779B9u);
#pragma unroll
			for(int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, o);
			if((threadIdx.x & 31) == 0) atomicAdd(accum, local);
		}

		// One-shot grids: a block per 256 items, consecutive blocks on consecutive memory (a capped,
		// grid-striding launch streams 15 % slower on B200, see convert_tiles.cuh).  The kernels keep
		// their stride loops, so the 2^31-block cap below is only a formality.
		unsigned grid_for(uint64_t items, int /*sm_count*/) {
			uint64_t blocks = (items + 255) / 256;
			if(blocks > 0x7FFFFFFFull) blocks = 0x7FFFFFFFull;
			if(blocks == 0) blocks = 1;
			return static_cast<unsigned>(blocks);
		}
	}

	cudaError_t flip_rows_launch(const void* src, void* dst, uint64_t pitch, uint32_t h, int sm_count, cudaStream_t stream) {
		const bool vec = (pitch % 16 == 0) && (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0);
		if(vec)
			flip_rows_vec_kernel<<<grid_for(pitch / 16 * h, sm_count), 256, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), pitch / 16, h);
		else
			flip_rows_byte_kernel<<<grid_for(pitch * h, sm_count), 256, 0, stream>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), pitch, h);
		return cudaGetLastError();
	}

	cudaError_t shift_left_u16_launch(const void* src, void* dst, uint64_t samples, uint32_t shift, cudaStream_t stream) {
		const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
		const uint64_t units = vec ? samples / 8 : 0;
		if(units) {
			const uint32_t half = (0xFFFFu << shift) & 0xFFFFu;
			shl16_vec_kernel<<<static_cast<unsigned>((units + 255) / 256), 256, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), units, shift, half | (half << 16));
			cudaError_t e = cudaGetLastError();
			if(e != cudaSuccess) return e;
		}
		if(units * 8 < samples) {
			const uint64_t rest = samples - units * 8;
			shl16_scalar_kernel<<<static_cast<unsigned>((rest + 255) / 256), 256, 0, stream>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), units * 8, samples, shift);
			return cudaGetLastError();
		}
		return cudaSuccess;
	}

	cudaError_t synth_launch(void* dst, uint64_t first_sample, uint64_t samples, int fmt, int ch, uint32_t seed, int sm_count, cudaStream_t stream) {
		const unsigned grid = grid_for(samples, sm_count);
		switch(fmt) {
			case CACAO_FMT_U8: synth_kernel<CACAO_FMT_U8><<<grid, 256, 0, stream>>>(static_cast<uint8_t*>(dst), first_sample, samples, ch, seed); break;
			case CACAO_FMT_U16: synth_kernel<CACAO_FMT_U16><<<grid, 256, 0, stream>>>(static_cast<uint16_t*>(dst), first_sample, samples, ch, seed); break;
			case CACAO_FMT_F16: synth_kernel<CACAO_FMT_F16><<<grid, 256, 0, stream>>>(static_cast<uint16_t*>(dst), first_sample, samples, ch, seed); break;
			case CACAO_FMT_F32: synth_kernel<CACAO_FMT_F32><<<grid, 256, 0, stream>>>(static_cast<float*>(dst), first_sample, samples, ch, seed); break;
			default: return cudaErrorInvalidValue;
		}
		return cudaGetLastError();
	}

	cudaError_t checksum_launch(const void* src, uint64_t words, unsigned long long* accum, int sm_count, cudaStream_t stream) {
		checksum_kernel<<<grid_for(words, sm_count), 256, 0, stream>>>(static_cast<const uint32_t*>(src), words, accum);
		return cudaGetLastError();
	}

} // namespace cacao
