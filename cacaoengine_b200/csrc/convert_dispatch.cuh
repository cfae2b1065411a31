// convert_dispatch.cuh -- runtime (format, layout) -> compiled shape, plus the scalar edge kernel.
//
// Split per SOURCE format into four translation units (convert_u8.cu ... convert_f32.cu) purely
// so nvcc can build them in parallel.
#pragma once

#include "convert_tiles.cuh"
#include "convert_wide.cuh"
#include "convert_lanes.cuh"

namespace cacao {

	struct ConvertJob {
		const void* src;
		void* dst;
		uint64_t pixels;           // all images together
		uint64_t pixels_per_image; // domain of the Gray->RGBA pixel-0 quirk
		int src_ch, dst_ch, src_fmt, dst_fmt;
		int mode;
		uint32_t flip_w = 0, flip_h = 0; // != 0: images are flip_w x flip_h pixels and come out bottom row first
		const uint8_t* lut_dev;   // 65536-byte table in device global memory
		const uint16_t* lut2_dev; // its 4096-entry two-level form (convert_pixel.cuh)
		const uint32_t* lanes_dev = nullptr; // its 644-word bank-private log-indexed form (convert_pixel.cuh)
		TileLaunch dev;
		cudaStream_t stream;
		unsigned launches; // out: kernels launched
	};

	// One thread per pixel, byte-wise loads and stores: used for the < one-tile tail, for buffers
	// that are not 16-byte aligned, and for the reference's Gray->RGBA quirk (Layout.cpp:50-51:
	// the source pointer never advances, so every output pixel repeats pixel 0 of its image).
	template<int SF, int DF, int SC, int DC, bool FLIP = false>
	__global__ void __launch_bounds__(256) convert_edge_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t p_begin, uint64_t p_end, uint64_t pixels_per_image, int pixel0_quirk, const uint8_t* __restrict__ lut_global, uint32_t gray16_cap, FlipMap flip) {
		constexpr int sb = FmtTraits<SF>::bytes, db = FmtTraits<DF>::bytes;
		PixelCtx ctx;
		ctx.lut = lut_global;
		ctx.lut2 = nullptr;
		ctx.lanes = nullptr;
		ctx.gray16_cap = gray16_cap;
		pdl_launch_dependents();
		pdl_wait(); // cacao_device.cuh: launched behind the vector kernel of the same call, whose tail it may overlap
		const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
		for(uint64_t p = p_begin + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < p_end; p += stride) {
			const uint64_t sp = pixel0_quirk ? (p / pixels_per_image) * pixels_per_image : p;
			const uint8_t* s = src + sp * (SC * sb);
			uint32_t in[SC], out[DC];
#pragma unroll
			for(int c = 0; c < SC; ++c) {
				uint32_t v = 0;
#pragma unroll
				for(int b = 0; b < sb; ++b) v |= static_cast<uint32_t>(s[c * sb + b]) << (8 * b);
				in[c] = v;
			}
			convert_pixel<SF, DF, SC, DC>(in, out, ctx);
			uint8_t* d = dst + (FLIP ? flip(p) : p) * (DC * db);
#pragma unroll
			for(int c = 0; c < DC; ++c) {
#pragma unroll
				for(int b = 0; b < db; ++b) d[c * db + b] = static_cast<uint8_t>(out[c] >> (8 * b));
			}
		}
	}

	template<int SF, int DF, int SC, int DC>
	cudaError_t run_convert_shape(ConvertJob& job) {
		using S = TileShape<SF, DF, SC, DC>;
		const bool quirk = FmtTraits<SF>::is_int && SC == 1 && DC == 4 && job.mode == CACAO_MODE_REF_EXACT;
		const bool aligned = ((reinterpret_cast<uintptr_t>(job.src) | reinterpret_cast<uintptr_t>(job.dst)) & 15u) == 0;
		const uint32_t gray16_cap = (job.mode == CACAO_MODE_REF_EXACT) ? 255u : 65535u;
		// preferred: the 256-bit sector kernel (convert_wide.cuh) when the shape fits it and both
		// buffers are 32-byte aligned; otherwise the 128-bit tile kernel (convert_tiles.cuh)
		uint64_t done = 0;
		bool wide = false;
		// Row mirror in the same pass: the kernel whose unit (lane run / tile) divides the image width
		// stores unit k of row y at row h-1-y; a width no unit divides goes through the byte-wise kernel.
		const bool flipped = job.flip_w != 0 && job.flip_h > 1;
		const bool tile_fits = !flipped || job.flip_w % S::tile_pixels == 0;
		bool lanes = false;
		if constexpr(SF == CACAO_FMT_U16 && DF == CACAO_FMT_U8) {
			// the bank-private table form (convert_lanes.cuh): large 16 -> 8 bit jobs on 32-byte aligned buffers
			using LS = LanesShape<SC, DC>;
			const bool aligned32 = ((reinterpret_cast<uintptr_t>(job.src) | reinterpret_cast<uintptr_t>(job.dst)) & 31u) == 0;
			const uint64_t n_runs = job.pixels / LS::G;
			lanes = !quirk && aligned32 && job.lanes_dev != nullptr && (!flipped || job.flip_w % LS::G == 0) && lanes_preferred<SC, DC>(n_runs);
			if(lanes) {
				if(n_runs > 0) {
					FlipMap fm;
					if(flipped) fm.per_row = job.flip_w / LS::G, fm.rows = job.flip_h;
					cudaError_t e = launch_convert_lanes<SC, DC>(job.src, job.dst, n_runs, job.lanes_dev, job.stream, fm);
					if(e != cudaSuccess) return e;
					job.launches += 1;
				}
				done = n_runs * LS::G;
			}
		}
		if constexpr(WideShape<SF, DF, SC, DC>::ok) {
			using WS = WideShape<SF, DF, SC, DC>;
			const bool aligned32 = ((reinterpret_cast<uintptr_t>(job.src) | reinterpret_cast<uintptr_t>(job.dst)) & 31u) == 0;
			const char* force = getenv("CACAO_FORCE_WIDE"); // dev/test switch: run every fitting shape on the sector kernel
			const bool wide_fits = !flipped || job.flip_w % WS::G == 0;
			wide = !lanes && !quirk && aligned32 && !getenv("CACAO_NO_WIDE") && wide_fits && (WS::preferred || (force && force[0] == '1') || (flipped && !(tile_fits && aligned)));
			if(wide) {
				const uint64_t n_runs = job.pixels / WS::G;
				if(n_runs > 0) {
					FlipMap fm;
					if(flipped) fm.per_row = job.flip_w / WS::G, fm.rows = job.flip_h;
					cudaError_t e = launch_convert_wide<SF, DF, SC, DC>(job.src, job.dst, n_runs, job.lut_dev, job.lut2_dev, gray16_cap, job.stream, fm);
					if(e != cudaSuccess) return e;
					job.launches += 1;
				}
				done = n_runs * WS::G;
			}
		}
		if(!wide && !lanes) {
			const uint64_t n_tiles = (quirk || !aligned || !tile_fits) ? 0 : job.pixels / S::tile_pixels;
			if(n_tiles > 0) {
				FlipMap fm;
				if(flipped) fm.per_row = static_cast<uint32_t>(job.flip_w / S::tile_pixels), fm.rows = job.flip_h;
				cudaError_t e = launch_convert_tiles<SF, DF, SC, DC>(job.src, job.dst, n_tiles, job.lut2_dev, gray16_cap, job.dev, job.stream, fm);
				if(e != cudaSuccess) return e;
				job.launches += 1;
			}
			done = n_tiles * S::tile_pixels;
		}
		if(done < job.pixels) {
			const uint64_t rest = job.pixels - done;
			uint64_t blocks = (rest + 255) / 256;
			const uint64_t cap = static_cast<uint64_t>(job.dev.sm_count) * 8;
			if(blocks > cap) blocks = cap;
			FlipMap fm; // in pixels
			if(flipped) fm.per_row = job.flip_w, fm.rows = job.flip_h;
			auto edge = flipped ? convert_edge_kernel<SF, DF, SC, DC, true> : convert_edge_kernel<SF, DF, SC, DC, false>;
			cudaError_t e = launch_pdl(edge, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, job.stream, static_cast<const uint8_t*>(job.src), static_cast<uint8_t*>(job.dst), done, job.pixels, job.pixels_per_image, quirk ? 1 : 0, job.lut_dev, gray16_cap, fm);
			if(e != cudaSuccess) return e;
			job.launches += 1;
		}
		return cudaSuccess;
	}

	template<int SF, int DF, int SC>
	cudaError_t dispatch_dst_ch(ConvertJob& job) {
		switch(job.dst_ch) {
			case 1:
				if constexpr(!(SC == 1 && SF == DF)) return run_convert_shape<SF, DF, SC, 1>(job);
				break;
			case 3:
				if constexpr(!(SC == 3 && SF == DF)) return run_convert_shape<SF, DF, SC, 3>(job);
				break;
			case 4:
				if constexpr(!(SC == 4 && SF == DF)) return run_convert_shape<SF, DF, SC, 4>(job);
				break;
		}
		return cudaErrorInvalidValue;
	}

	template<int SF, int DF>
	cudaError_t dispatch_src_ch(ConvertJob& job) {
		switch(job.src_ch) {
			case 1: return dispatch_dst_ch<SF, DF, 1>(job);
			case 3: return dispatch_dst_ch<SF, DF, 3>(job);
			case 4: return dispatch_dst_ch<SF, DF, 4>(job);
		}
		return cudaErrorInvalidValue;
	}

	template<int SF>
	cudaError_t dispatch_dst_fmt(ConvertJob& job) {
		switch(job.dst_fmt) {
			case CACAO_FMT_U8: return dispatch_src_ch<SF, CACAO_FMT_U8>(job);
			case CACAO_FMT_U16: return dispatch_src_ch<SF, CACAO_FMT_U16>(job);
			case CACAO_FMT_F16: return dispatch_src_ch<SF, CACAO_FMT_F16>(job);
			case CACAO_FMT_F32: return dispatch_src_ch<SF, CACAO_FMT_F32>(job);
		}
		return cudaErrorInvalidValue;
	}

	// defined in convert_u8.cu / convert_u16.cu / convert_f16.cu / convert_f32.cu
	cudaError_t convert_from_u8(ConvertJob& job);
	cudaError_t convert_from_u16(ConvertJob& job);
	cudaError_t convert_from_f16(ConvertJob& job);
	cudaError_t convert_from_f32(ConvertJob& job);

} // namespace cacao
