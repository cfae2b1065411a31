// convert_f32.cu -- instantiates every conversion shape whose SOURCE format is F32.
#include "convert_dispatch.cuh"

namespace cacao {
	cudaError_t convert_from_f32(ConvertJob& job) {
		return dispatch_dst_fmt<CACAO_FMT_F32>(job);
	}
}
