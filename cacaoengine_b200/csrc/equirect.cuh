// equirect.cuh -- launcher interface of the equirect -> cubemap kernel (see equirect.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "projection.cuh"

namespace cacao {

	constexpr int kMaxUnitsPerLaunch = 32;

	// one work unit = rows [row_begin,row_end) of one face
	struct CubeUnit {
		uint32_t face, row_begin, row_end;
	};
	// Internal to the launcher and the kernel (callers' units have face 0..5): the unit lies in the upper
	// half of its face and stands for its rows and their mirror rows N-1-j.
	constexpr uint32_t kUnitMirror = 8u;

	struct EquirectParams {
		uint32_t W, H;               // source size in pixels
		uint32_t N;                  // face edge in pixels
		uint32_t src_row0, src_rows; // window of source rows present in `src`
		uint32_t n_units;            // blockIdx.z selects the unit
		CubeUnit units[kMaxUnitsPerLaunch];
		ProjConst pc;
		uint32_t row_add, row_mul;   // face row j is stored at row row_add + row_mul * j: (0, 1), or (N-1, -1) = bottom row first
	};

	// Enqueue the resampling of a list of (face, row band) units in as few launches as their heights allow
	// (normally one), *launches gets the count.
	// dst_ch / dst_fmt other than ch / fmt: resampling and format conversion in one pass (Spec B.4; float
	// sources, same channels or RGB <-> RGBA -- equirect_cvt_supported); <= 0 / < 0 mean "as the source".
	// flags: CACAO_EQUIRECT_FLIP_ROWS stores every face bottom row first (units still name rows of the
	// unflipped face: row j lands at N-1-j).
	cudaError_t equirect_launch_units(const void* src, void* dst, int ch, int fmt, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, uint32_t N, const CubeUnit* units, uint32_t n_units, cudaStream_t stream, unsigned* launches, int dst_ch = 0, int dst_fmt = -1, uint32_t flags = 0);
	bool equirect_cvt_supported(int ch, int fmt, int dst_ch, int dst_fmt);

	// The row-strip kernel (equirect_rows.cu): same pixels, fewer instructions -- ahead on the one-channel and
	// 8-bit shapes, which are bound by instruction issue.
	// equirect_launch_units routes to it by shape; CACAO_EQUIRECT_ROWS=0 / 1 forces the gather kernel / the
	// row-strip kernel wherever it is defined (tests, A/B timing).
	bool equirect_rows_supported(int ch, int fmt, int dst_ch, int dst_fmt, uint32_t W, uint32_t H);
	bool equirect_rows_preferred(int ch, int fmt);
	cudaError_t equirect_rows_launch(const void* src, void* dst, int ch, int fmt, const EquirectParams& prm, int sm_count, cudaStream_t stream);

	// EXPERIMENT (equirect_tma.cu, CACAO_EQUIRECT_TMA=1): whole cubes of 4- / 8-byte texels through a persistent,
	// warp-specialised kernel whose source footprints are staged by 2-D tensor-map TMA copies in a four-stage ring.
	bool equirect_tma_supported(int ch, int fmt, uint32_t W, uint32_t H, const void* src);
	cudaError_t equirect_tma_launch(const void* src, void* dst, int ch, int fmt, const EquirectParams& prm, int sm_count, cudaStream_t stream);

	// The folding step of equirect_launch_units on its own (host only): mirror bands of the input become
	// single units flagged kUnitMirror; units with face > 5 or no rows are dropped.
	std::vector<CubeUnit> pair_mirror_units(const std::vector<CubeUnit>& in, uint32_t N);

	// Enqueue the resampling of rows [row_begin,row_end) of every face in face_mask (one launch).
	cudaError_t equirect_launch(const void* src, void* dst, int ch, int fmt, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, uint32_t N, uint32_t face_mask, uint32_t row_begin, uint32_t row_end, cudaStream_t stream);

	// Which source rows that work unit reads (host arithmetic identical to the kernel's).
	void equirect_src_window(uint32_t W, uint32_t H, uint32_t N, uint32_t face_mask, uint32_t row_begin, uint32_t row_end, uint32_t* src_row0, uint32_t* src_rows);

} // namespace cacao
