// dev/l1_align_bench.cu -- does a warp-wide 16-byte-per-lane load cost more L1 data-pipe passes when its
// 8-lane groups straddle 128-byte lines?  (Hypothesis behind the equirect kernel's 8.5 passes per tap.)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while(0)

template<int STRIDE16>  // lane stride in 16-byte units (1 = contiguous texels)
__global__ void __launch_bounds__(256) probe(const uint4* __restrict__ buf, uint32_t mask, int off, int iters, uint32_t* out) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t acc = 0;
	uint32_t idx = (blockIdx.x * 8 + warp) * 64 + lane * STRIDE16 + off;
	for(int it = 0; it < iters; ++it) {
		const uint4 v = __ldg(buf + (idx & mask));
		acc ^= v.x ^ v.y ^ v.z ^ v.w;
		idx += 32 * STRIDE16;
	}
	if(acc == 0xDEADBEEF) out[0] = acc;
}

int main() {
	const uint32_t units = 1u << 12;  // 64 KiB: lives in L1 after the first pass
	uint4* buf; uint32_t* out;
	CK(cudaMalloc(&buf, units * 16)); CK(cudaMalloc(&out, 4));
	CK(cudaMemset(buf, 1, units * 16));
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	const int iters = 4096, blocks = 148 * 8;
	for(int off = 0; off < 8; ++off) {
		for(int rep = 0; rep < 2; ++rep) {
			CK(cudaEventRecord(a));
			probe<1><<<blocks, 256>>>(buf, units - 1, off, iters, out);
			CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
		}
		float ms; CK(cudaEventElapsedTime(&ms, a, b));
		const double reqs = (double)blocks * 8 * iters;
		printf("offset %d texels: %.3f ms  %.2f ns per warp request per SM-equivalent, %.1f GB/s L1->reg\n", off, ms, ms * 1e6 / (reqs / 148), reqs * 512 / ms / 1e6);
	}
	return 0;
}
