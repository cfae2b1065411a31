// mip.cuh -- 2x2 box downsampling (one mip level), see mip.cu.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace cacao {

	// dst (max(1,w/2) x max(1,h/2)) = 2x2 box average of src (w x h), DESIGN.md "Spec B.3".
	// *launches gets the number of kernels enqueued (vector body and / or scalar remainder).
	cudaError_t downsample2x_launch(const void* src, void* dst, uint32_t w, uint32_t h, int ch, int fmt, cudaStream_t stream, unsigned* launches);

} // namespace cacao
