// convert_u8.cu -- instantiates every conversion shape whose SOURCE format is U8.
#include "convert_dispatch.cuh"

namespace cacao {
	cudaError_t convert_from_u8(ConvertJob& job) {
		return dispatch_dst_fmt<CACAO_FMT_U8>(job);
	}
}
