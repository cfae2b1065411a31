// convert_f16.cu -- instantiates every conversion shape whose SOURCE format is F16.
#include "convert_dispatch.cuh"

namespace cacao {
	cudaError_t convert_from_f16(ConvertJob& job) {
		return dispatch_dst_fmt<CACAO_FMT_F16>(job);
	}
}
