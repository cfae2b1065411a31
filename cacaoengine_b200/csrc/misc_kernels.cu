// misc_kernels.cu -- flip, synthetic-input and checksum kernels.
#include "misc_kernels.cuh"

#include "cacao_device.cuh"

namespace cacao {

	namespace {

		// Vertical mirror.  A row is a contiguous run of `pitch` bytes; when pitch and both bases are
		// 16-byte aligned every row moves as coalesced 16-B vectors, otherwise byte by byte.
		__global__ void __launch_bounds__(256) flip_rows_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t units_per_row, uint32_t h) {
			const uint64_t total = units_per_row * h;
			pdl_launch_dependents();
			pdl_wait(); // cacao_device.cuh
			const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
			for(uint64_t u = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < total; u += stride) {
				const uint64_t y = u / units_per_row, x = u - y * units_per_row;
				stg_stream(dst + u, ldg_stream(src + (h - 1 - y) * units_per_row + x));
			}
		}
		__global__ void __launch_bounds__(256) flip_rows_byte_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t pitch, uint32_t h) {
			const uint64_t total = pitch * h;
			pdl_launch_dependents();
			pdl_wait();
			const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
			for(uint64_t u = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < total; u += stride) {
				const uint64_t y = u / pitch, x = u - y * pitch;
				dst[u] = src[(h - 1 - y) * pitch + x];
			}
		}

		// 16-bit left shift on packed pairs: (w << s) with the bits that crossed the half-word border masked off
		// In place (src == dst, which the header allows) nothing may promise the compiler disjoint, read-only
		// memory: that form has no __restrict__ and loads with plain ld.global instead of ld.global.nc.
		__global__ void __launch_bounds__(256) shl16_inplace_kernel(uint4* buf, uint64_t units, uint32_t shift, uint32_t mask) {
			const uint64_t u = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
			if(u >= units) return;
			uint4 v = buf[u];
			v.x = (v.x << shift) & mask;
			v.y = (v.y << shift) & mask;
			v.z = (v.z << shift) & mask;
			v.w = (v.w << shift) & mask;
			buf[u] = v;
		}
		__global__ void __launch_bounds__(256) shl16_scalar_inplace_kernel(uint16_t* buf, uint64_t begin, uint64_t end, uint32_t shift) {
			const uint64_t k = begin + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
			if(k < end) buf[k] = static_cast<uint16_t>(static_cast<uint32_t>(buf[k]) << shift);
		}
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

		// ---- texel unpack ----------------------------------------------------------------------------
		// One source word -> [r,g,b,a] bytes with the hint's sub-8-bit channels widened exactly as
		// Model.cpp:305-310 does: v = (v & ((1 << bits) - 1)) << (8 - bits); v |= v >> bits.
		template<bool ALL8>
		__device__ __forceinline__ uint32_t unpack_word(uint32_t w, const TexelPlan& pl) {
			const uint32_t t = __byte_perm(w, 0u, pl.sel);
			if constexpr(ALL8) return t;
			uint32_t r = 0;
#pragma unroll
			for(int c = 0; c < 4; ++c) {
				const uint32_t b = pl.bits[c];
				uint32_t v = ((t >> (8 * c)) & ((1u << b) - 1u)) << (8u - b);
				v |= v >> b;
				r |= (v & 0xFFu) << (8 * c);
			}
			return r;
		}

		// 16 texels (64 B = two whole 32-byte sectors) per thread in, 64 / 48 / 16 B out
		template<int OUT_CH, bool ALL8>
		__global__ void __launch_bounds__(256) unpack_texels_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t runs, TexelPlan pl) {
			const uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
			if(r >= runs) return;
			uint32_t t[16];
			ldg256_stream(src + r * 4, t);
			ldg256_stream(src + r * 4 + 2, t + 8);
#pragma unroll
			for(int k = 0; k < 16; ++k) t[k] = unpack_word<ALL8>(t[k], pl);
			if constexpr(OUT_CH == 4) {
				stg256_stream(dst + r * 4, t);
				stg256_stream(dst + r * 4 + 2, t + 8);
			} else if constexpr(OUT_CH == 3) {
#pragma unroll
				for(int k = 0; k < 4; ++k) { // four texels -> three words
					const uint32_t a = t[4 * k], b = t[4 * k + 1], c = t[4 * k + 2], d = t[4 * k + 3];
					t[3 * k + 0] = __byte_perm(a, b, 0x4210u);
					t[3 * k + 1] = __byte_perm(b, c, 0x5421u);
					t[3 * k + 2] = __byte_perm(c, d, 0x6542u);
				}
#pragma unroll
				for(int k = 0; k < 3; ++k) stg_stream(dst + r * 3 + k, make_uint4(t[4 * k], t[4 * k + 1], t[4 * k + 2], t[4 * k + 3]));
			} else {
				uint32_t o[4];
#pragma unroll
				for(int k = 0; k < 4; ++k) o[k] = __byte_perm(__byte_perm(t[4 * k], t[4 * k + 1], 0x0040u), __byte_perm(t[4 * k + 2], t[4 * k + 3], 0x0040u), 0x5410u);
				stg_stream(dst + r, make_uint4(o[0], o[1], o[2], o[3]));
			}
		}
		__global__ void __launch_bounds__(256) unpack_texels_tail_kernel(const uint32_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t begin, uint64_t end, TexelPlan pl) {
			const uint64_t k = begin + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
			if(k >= end) return;
			const uint32_t t = unpack_word<false>(src[k], pl);
			for(uint32_t c = 0; c < pl.out_ch; ++c) dst[k * pl.out_ch + c] = static_cast<uint8_t>(t >> (8 * c));
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
		if(vec) return launch_pdl(flip_rows_vec_kernel, dim3(grid_for(pitch / 16 * h, sm_count)), dim3(256), 0, stream, static_cast<const uint4*>(src), static_cast<uint4*>(dst), pitch / 16, h);
		return launch_pdl(flip_rows_byte_kernel, dim3(grid_for(pitch * h, sm_count)), dim3(256), 0, stream, static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), pitch, h);
	}

	cudaError_t shift_left_u16_launch(const void* src, void* dst, uint64_t samples, uint32_t shift, cudaStream_t stream) {
		const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
		const uint64_t units = vec ? samples / 8 : 0;
		const bool inplace = src == dst; // allowed by the header: no read-only / no-alias promises then
		if(units) {
			const uint32_t half = (0xFFFFu << shift) & 0xFFFFu;
			if(inplace)
				shl16_inplace_kernel<<<static_cast<unsigned>((units + 255) / 256), 256, 0, stream>>>(static_cast<uint4*>(dst), units, shift, half | (half << 16));
			else
				shl16_vec_kernel<<<static_cast<unsigned>((units + 255) / 256), 256, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), units, shift, half | (half << 16));
			cudaError_t e = cudaGetLastError();
			if(e != cudaSuccess) return e;
		}
		if(units * 8 < samples) {
			const uint64_t rest = samples - units * 8;
			if(inplace)
				shl16_scalar_inplace_kernel<<<static_cast<unsigned>((rest + 255) / 256), 256, 0, stream>>>(static_cast<uint16_t*>(dst), units * 8, samples, shift);
			else
				shl16_scalar_kernel<<<static_cast<unsigned>((rest + 255) / 256), 256, 0, stream>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), units * 8, samples, shift);
			return cudaGetLastError();
		}
		return cudaSuccess;
	}

	cudaError_t unpack_texels_launch(const void* src, void* dst, uint64_t texels, const TexelPlan& plan, cudaStream_t stream, unsigned* launches) {
		if(launches) *launches = 0;
		const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 31u) == 0; // 256-bit accesses
		const uint64_t runs = vec ? texels / 16 : 0;
		const bool all8 = plan.bits[0] == 8 && plan.bits[1] == 8 && plan.bits[2] == 8 && plan.bits[3] == 8;
		if(runs) {
			const unsigned grid = static_cast<unsigned>((runs + 255) / 256);
			const uint4* s4 = static_cast<const uint4*>(src);
			uint4* d4 = static_cast<uint4*>(dst);
#define CACAO_UNPACK(CHV)                                                                                   \
	if(all8)                                                                                                \
		unpack_texels_kernel<CHV, true><<<grid, 256, 0, stream>>>(s4, d4, runs, plan);                      \
	else                                                                                                    \
		unpack_texels_kernel<CHV, false><<<grid, 256, 0, stream>>>(s4, d4, runs, plan);
			if(plan.out_ch == 4) {
				CACAO_UNPACK(4)
			} else if(plan.out_ch == 3) {
				CACAO_UNPACK(3)
			} else {
				CACAO_UNPACK(1)
			}
#undef CACAO_UNPACK
			cudaError_t e = cudaGetLastError();
			if(e != cudaSuccess) return e;
			if(launches) *launches += 1;
		}
		if(runs * 16 < texels) {
			const uint64_t rest = texels - runs * 16;
			unpack_texels_tail_kernel<<<static_cast<unsigned>((rest + 255) / 256), 256, 0, stream>>>(static_cast<const uint32_t*>(src), static_cast<uint8_t*>(dst), runs * 16, texels, plan);
			if(launches) *launches += 1;
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
