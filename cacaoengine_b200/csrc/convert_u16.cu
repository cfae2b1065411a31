// convert_u16.cu -- instantiates every conversion shape whose SOURCE format is U16.
#include "convert_dispatch.cuh"

namespace cacao {
	cudaError_t convert_from_u16(ConvertJob& job) {
		return dispatch_dst_fmt<CACAO_FMT_U16>(job);
	}
}
