// equirect.cuh -- launcher interface of the equirect -> cubemap kernel (see equirect.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "projection.cuh"

namespace cacao {

	struct EquirectParams {
		uint32_t W, H;               // source size in pixels
		uint32_t N;                  // face edge in pixels
		uint32_t src_row0, src_rows; // window of source rows present in `src`
		uint32_t row_begin, row_end; // face rows to produce
		uint32_t faces;              // packed list: nibble k = k-th selected face
		ProjConst pc;
	};

	// Enqueue the resampling of rows [row_begin,row_end) of every face in face_mask.
	cudaError_t equirect_launch(const void* src, void* dst, int ch, int fmt, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, uint32_t N, uint32_t face_mask, uint32_t row_begin, uint32_t row_end, cudaStream_t stream);

	// Which source rows that work unit reads (host arithmetic identical to the kernel's).
	void equirect_src_window(uint32_t W, uint32_t H, uint32_t N, uint32_t face_mask, uint32_t row_begin, uint32_t row_end, uint32_t* src_row0, uint32_t* src_rows);

} // namespace cacao
