// mip.cu -- one mip level: 2x2 box average (DESIGN.md "Spec B.3").
//
// No reference counterpart in libcacaoimage: the engine leaves mip generation to the graphics API
// (glGenerateMipmap, engine/src/opengl/impls/OpenGLTex2D.cpp:51).  It is the follow-on SURVEY.md
// names for row f-4: with the faces already resident and converted, the chain can be produced where
// the pixels are and handed over in one piece.
//
// Streaming kernel: a thread owns K output pixels of one output row, K chosen per texel size so that
// the two source rows it reads are whole 32-byte sectors (LDG.256) and its output whole 16-byte
// vectors; lanes are contiguous in both, so a warp reads 2 x 1-3 KiB and writes 0.5-1.5 KiB, fully
// coalesced.  Images whose rows do not keep that alignment (odd widths) take the per-pixel kernel.
#include "mip.cuh"

#include <algorithm>

#include "cacao_device.cuh"
#include "convert_tiles.cuh"

namespace cacao {

	namespace {
		// average of four samples held in the register convention of convert_pixel.cuh
		// (integers as values, F16 as raw bits, F32 as float bits)
		template<int F>
		__device__ __forceinline__ uint32_t box4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
			if constexpr(FmtTraits<F>::is_int) {
				return (a + b + c + d + 2u) >> 2; // round half up
			} else if constexpr(F == CACAO_FMT_F32) {
				return __float_as_uint(((__uint_as_float(a) + __uint_as_float(b)) + (__uint_as_float(c) + __uint_as_float(d))) * 0.25f);
			} else {
				const float s = (f16_bits_to_f32(static_cast<uint16_t>(a)) + f16_bits_to_f32(static_cast<uint16_t>(b))) + (f16_bits_to_f32(static_cast<uint16_t>(c)) + f16_bits_to_f32(static_cast<uint16_t>(d)));
				return f32_to_f16_bits(s * 0.25f);
			}
		}

		template<int F, int CH>
		struct MipShape {
			static constexpr int sb = FmtTraits<F>::bytes;
			static constexpr int bpp = sb * CH;
			static constexpr int K = clcm(2 * bpp, 32) / (2 * bpp); // output pixels per thread
			static constexpr int in_sectors = K * 2 * bpp / 32;     // per source row
			static constexpr int out_vec = K * bpp / 16;
			static_assert(K * bpp % 16 == 0, "a thread's output must be whole 16-byte vectors");
		};

		template<int F, int CH>
		__global__ void __launch_bounds__(256) downsample2x_vec_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t runs_per_row, uint32_t h, uint32_t ho, uint64_t src_pitch, uint64_t dst_pitch) {
			using S = MipShape<F, CH>;
			pdl_launch_dependents(); // a mip chain is a string of short dependent launches: each one's launch
			pdl_wait();              // latency and block scheduling overlap its predecessor (cacao_device.cuh)
			const uint32_t rx = blockIdx.x * blockDim.x + threadIdx.x;
			const uint32_t y = blockIdx.y;
			if(rx >= runs_per_row || y >= ho) return;
			const uint32_t y0 = 2u * y, y1 = min(2u * y + 1u, h - 1u);
			uint32_t r0[S::in_sectors * 8], r1[S::in_sectors * 8];
			const uint8_t* p0 = src + y0 * src_pitch + static_cast<uint64_t>(rx) * (S::in_sectors * 32);
			const uint8_t* p1 = src + y1 * src_pitch + static_cast<uint64_t>(rx) * (S::in_sectors * 32);
#pragma unroll
			for(int k = 0; k < S::in_sectors; ++k) {
				ldg256_stream(p0 + 32 * k, r0 + 8 * k);
				ldg256_stream(p1 + 32 * k, r1 + 8 * k);
			}
			uint32_t o[S::out_vec * 4];
#pragma unroll
			for(int i = 0; i < S::out_vec * 4; ++i) o[i] = 0u;
#pragma unroll
			for(int q = 0; q < S::K; ++q) {
#pragma unroll
				for(int c = 0; c < CH; ++c) {
					const uint32_t a = unpack_sample<S::sb>(r0, (2 * q) * CH + c), b = unpack_sample<S::sb>(r0, (2 * q + 1) * CH + c);
					const uint32_t cc = unpack_sample<S::sb>(r1, (2 * q) * CH + c), d = unpack_sample<S::sb>(r1, (2 * q + 1) * CH + c);
					pack_sample<S::sb>(o, q * CH + c, box4<F>(a, b, cc, d));
				}
			}
			uint8_t* q = dst + y * dst_pitch + static_cast<uint64_t>(rx) * (S::out_vec * 16);
#pragma unroll
			for(int k = 0; k < S::out_vec; ++k) stg_stream(q + 16 * k, make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
		}

		// one thread per output pixel, columns [x_begin, wo): any size, any alignment
		template<int F, int CH>
		__global__ void __launch_bounds__(256) downsample2x_px_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t w, uint32_t h, uint32_t wo, uint32_t ho, uint32_t x_begin) {
			using T = typename FmtTraits<F>::type;
			pdl_launch_dependents();
			pdl_wait();
			const uint32_t x = x_begin + blockIdx.x * blockDim.x + threadIdx.x;
			const uint32_t y = blockIdx.y;
			if(x >= wo || y >= ho) return;
			const uint32_t x0 = 2u * x, x1 = min(2u * x + 1u, w - 1u), y0 = 2u * y, y1 = min(2u * y + 1u, h - 1u);
			const T* s = reinterpret_cast<const T*>(src);
			T* d = reinterpret_cast<T*>(dst);
			auto at = [&](uint32_t xx, uint32_t yy, int c) -> uint32_t {
				const T v = s[(static_cast<uint64_t>(yy) * w + xx) * CH + c];
				if constexpr(F == CACAO_FMT_F32)
					return __float_as_uint(v);
				else
					return static_cast<uint32_t>(v);
			};
#pragma unroll
			for(int c = 0; c < CH; ++c) {
				const uint32_t r = box4<F>(at(x0, y0, c), at(x1, y0, c), at(x0, y1, c), at(x1, y1, c));
				T& out = d[(static_cast<uint64_t>(y) * wo + x) * CH + c];
				if constexpr(F == CACAO_FMT_F32)
					out = __uint_as_float(r);
				else
					out = static_cast<T>(r);
			}
		}

		template<int F, int CH>
		cudaError_t launch(const void* src, void* dst, uint32_t w, uint32_t h, cudaStream_t stream, unsigned* launches) {
			using S = MipShape<F, CH>;
			const uint32_t wo = std::max(1u, w / 2), ho = std::max(1u, h / 2);
			const uint64_t src_pitch = static_cast<uint64_t>(w) * S::bpp, dst_pitch = static_cast<uint64_t>(wo) * S::bpp;
			const bool vec = w >= 2 && src_pitch % 32 == 0 && dst_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 31u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
			const uint32_t runs = vec ? wo / S::K : 0;
			const uint8_t* s8 = static_cast<const uint8_t*>(src);
			uint8_t* d8 = static_cast<uint8_t*>(dst);
			// gridDim.y <= 65535: taller images go in slices of output rows
			for(uint32_t y0 = 0; y0 < ho; y0 += 65535u) {
				const uint32_t rows = std::min(65535u, ho - y0);
				const uint32_t h_left = h - 2u * y0;
				if(runs) {
					const dim3 grid((runs + 255) / 256, rows, 1);
					if(cudaError_t e = launch_pdl(downsample2x_vec_kernel<F, CH>, grid, dim3(256), 0, stream, s8 + 2ull * y0 * src_pitch, d8 + y0 * dst_pitch, runs, h_left, rows, src_pitch, dst_pitch); e != cudaSuccess) return e;
					if(launches) *launches += 1;
				}
				if(runs * S::K < wo) {
					const uint32_t rest = wo - runs * S::K;
					const dim3 grid((rest + 255) / 256, rows, 1);
					if(cudaError_t e = launch_pdl(downsample2x_px_kernel<F, CH>, grid, dim3(256), 0, stream, s8 + 2ull * y0 * src_pitch, d8 + y0 * dst_pitch, w, h_left, wo, rows, runs * S::K); e != cudaSuccess) return e;
					if(launches) *launches += 1;
				}
			}
			return cudaGetLastError();
		}

		template<int F>
		cudaError_t launch_ch(const void* src, void* dst, uint32_t w, uint32_t h, int ch, cudaStream_t stream, unsigned* launches) {
			switch(ch) {
				case 1: return launch<F, 1>(src, dst, w, h, stream, launches);
				case 3: return launch<F, 3>(src, dst, w, h, stream, launches);
				case 4: return launch<F, 4>(src, dst, w, h, stream, launches);
			}
			return cudaErrorInvalidValue;
		}
	}

	cudaError_t downsample2x_launch(const void* src, void* dst, uint32_t w, uint32_t h, int ch, int fmt, cudaStream_t stream, unsigned* launches) {
		if(launches) *launches = 0;
		if(w == 0 || h == 0) return cudaErrorInvalidValue;
		switch(fmt) {
			case CACAO_FMT_U8: return launch_ch<CACAO_FMT_U8>(src, dst, w, h, ch, stream, launches);
			case CACAO_FMT_U16: return launch_ch<CACAO_FMT_U16>(src, dst, w, h, ch, stream, launches);
			case CACAO_FMT_F16: return launch_ch<CACAO_FMT_F16>(src, dst, w, h, ch, stream, launches);
			case CACAO_FMT_F32: return launch_ch<CACAO_FMT_F32>(src, dst, w, h, ch, stream, launches);
		}
		return cudaErrorInvalidValue;
	}

} // namespace cacao
