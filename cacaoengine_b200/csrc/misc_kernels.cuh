// misc_kernels.cuh -- flip, synthetic-input and checksum kernels (see misc_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace cacao {

	// out.row[y] = src.row[h-1-y]; pitch in bytes (the semantics libs/image/src/Flip.cpp:6-42 intends)
	cudaError_t flip_rows_launch(const void* src, void* dst, uint64_t pitch, uint32_t h, int sm_count, cudaStream_t stream);

	// dst[k] = uint16(src[k] << shift) (libs/image/src/JPEG.cpp:79-82 uses shift 4)
	cudaError_t shift_left_u16_launch(const void* src, void* dst, uint64_t samples, uint32_t shift, cudaStream_t stream);

	// Texel unpack for assimp embedded textures (engine/src/core/Model.cpp:245-316): four bytes per
	// source texel in the order the format hint names; out_ch output bytes per texel in r,g,b,a order.
	// sel: PRMT selector that brings a source word into [r,g,b,a] (nibble 4 = dropped channel);
	// bits[c]: the hint's bit count of output channel c (1..8; below 8 = expand by bit replication).
	struct TexelPlan {
		uint32_t sel;
		uint32_t out_ch; // 4, 3 or 1
		uint32_t bits[4];
	};
	cudaError_t unpack_texels_launch(const void* src, void* dst, uint64_t texels, const TexelPlan& plan, cudaStream_t stream, unsigned* launches);

	// counter-hash synthetic samples (DESIGN.md "Synthetic inputs")
	cudaError_t synth_launch(void* dst, uint64_t first_sample, uint64_t samples, int fmt, int ch, uint32_t seed, int sm_count, cudaStream_t stream);

	// sum over 32-bit words of fmix32(word + index * 0x9E3779B9) into *accum (device, pre-zeroed)
	cudaError_t checksum_launch(const void* src, uint64_t words, unsigned long long* accum, int sm_count, cudaStream_t stream);

} // namespace cacao
