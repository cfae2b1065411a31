// dev/stream_bench.cu -- developer microbenchmark (not part of the product, not built by build.py).
//
// Explores how to stream "read R bytes, write W bytes" shapes on a B200: what a plain copy gets,
// what a 1:2 read:write mix gets, and which load/store widths, unrolls and grid shapes get there.
// The winners are folded back into cacaoengine_b200/csrc/convert_tiles.cuh.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false dev/stream_bench.cu -o dev/stream_bench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
	do {                                                                                        \
		cudaError_t e = (x);                                                                    \
		if(e != cudaSuccess) {                                                                  \
			fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); \
			exit(1);                                                                            \
		}                                                                                       \
	} while(0)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
	asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
	asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) {
	uint32_t r;
	asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
	return r;
}
// 4 u8 samples (one 32-bit word) -> 4 f16 (two words), RNE_f16(v * (1/255))
__device__ __forceinline__ void cvt_word(uint32_t w, uint32_t& o0, uint32_t& o1) {
	const float r = 1.0f / 255.0f;
	const float a = (float)(w & 0xFF) * r, b = (float)((w >> 8) & 0xFF) * r, c = (float)((w >> 16) & 0xFF) * r, d = (float)(w >> 24) * r;
	o0 = cvt2(a, b);
	o1 = cvt2(c, d);
}
__device__ __forceinline__ void cvt_unit(const uint4& in, uint4& o0, uint4& o1) {
	cvt_word(in.x, o0.x, o0.y);
	cvt_word(in.y, o0.z, o0.w);
	cvt_word(in.z, o1.x, o1.y);
	cvt_word(in.w, o1.z, o1.w);
}

// ---- A: plain copy, grid-stride, U 16-B loads in flight per thread
template<int U>
__global__ void __launch_bounds__(256) copy_kernel(const uint4* __restrict__ s, uint4* __restrict__ d, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
	for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i < n; i += stride) {
		uint4 v[U];
#pragma unroll
		for(int u = 0; u < U; ++u)
			if(i + u * 256 < n) v[u] = ldg_stream(s + i + u * 256);
#pragma unroll
		for(int u = 0; u < U; ++u)
			if(i + u * 256 < n) stg_stream(d + i + u * 256, v[u]);
	}
}

// ---- B: 1:2 expansion, lane-contiguous 32-B output per lane via two STG.128 (half sectors per instr)
// ---- C: same, one STG.256 per lane          MODE: 0 = two v4, 1 = v8, 2 = v8 + real convert math, 3 = two v4 + math
template<int U, int MODE>
__global__ void __launch_bounds__(256) expand_kernel(const uint4* __restrict__ s, uint4* __restrict__ d, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
	for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i < n; i += stride) {
		uint4 v[U];
#pragma unroll
		for(int u = 0; u < U; ++u)
			if(i + u * 256 < n) v[u] = ldg_stream(s + i + u * 256);
#pragma unroll
		for(int u = 0; u < U; ++u) {
			if(i + u * 256 < n) {
				uint4 a, b;
				if(MODE >= 2) {
					cvt_unit(v[u], a, b);
				} else {
					a = v[u];
					b = make_uint4(v[u].x ^ 1, v[u].y ^ 2, v[u].z ^ 3, v[u].w ^ 4);
				}
				uint4* o = d + 2 * (i + u * 256);
				if(MODE == 1 || MODE == 2) {
					stg256(o, a, b);
				} else {
					stg_stream(o, a);
					stg_stream(o + 1, b);
				}
			}
		}
	}
}

// ---- D: 32-B loads too (LDG.256: 8 px per lane -> two STG.256)
template<int U>
__global__ void __launch_bounds__(256) expand256_kernel(const uint4* __restrict__ s, uint4* __restrict__ d, uint64_t n /* 16-B units in */) {
	const uint64_t n2 = n / 2;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
	for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i < n2; i += stride) {
		uint4 v[U][2];
#pragma unroll
		for(int u = 0; u < U; ++u)
			if(i + u * 256 < n2) ldg256(s + 2 * (i + u * 256), v[u][0], v[u][1]);
#pragma unroll
		for(int u = 0; u < U; ++u) {
			if(i + u * 256 < n2) {
				uint4 a, b, c, e;
				cvt_unit(v[u][0], a, b);
				cvt_unit(v[u][1], c, e);
				uint4* o = d + 4 * (i + u * 256);
				stg256(o, a, b);
				stg256(o + 2, c, e);
			}
		}
	}
}

// ---- E: lane-contiguous chunks: each lane owns KIN*32 contiguous bytes in and KOUT*32 out (no smem, no
// warp-level coalescing beyond full 32-B sectors).  Does DRAM care?  KIN=KOUT: copy; 3->4: RGB8->RGBA8 math.
template<int KIN, int KOUT>
__global__ void __launch_bounds__(256) lane_chunk_kernel(const uint4* __restrict__ s, uint4* __restrict__ d, uint64_t n_lanes) {
	const uint64_t lane = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(lane >= n_lanes) return;
	uint4 v[KIN][2];
	const uint4* sp = s + lane * (2 * KIN);
#pragma unroll
	for(int k = 0; k < KIN; ++k) ldg256(sp + 2 * k, v[k][0], v[k][1]);
	uint4* dp = d + lane * (2 * KOUT);
	if constexpr(KIN == KOUT) {
#pragma unroll
		for(int k = 0; k < KOUT; ++k) stg256(dp + 2 * k, v[k][0], v[k][1]);
	} else {
		// 96 B = 32 RGB8 px -> 128 B RGBA8: word j of the output = bytes 3j..3j+2 of the input | 0xFF000000
		uint32_t w[KIN * 8 + 1];
#pragma unroll
		for(int k = 0; k < KIN; ++k) {
			w[8 * k + 0] = v[k][0].x; w[8 * k + 1] = v[k][0].y; w[8 * k + 2] = v[k][0].z; w[8 * k + 3] = v[k][0].w;
			w[8 * k + 4] = v[k][1].x; w[8 * k + 5] = v[k][1].y; w[8 * k + 6] = v[k][1].z; w[8 * k + 7] = v[k][1].w;
		}
		w[KIN * 8] = 0;
		uint32_t o[KOUT * 8];
#pragma unroll
		for(int j = 0; j < KOUT * 8; ++j) {
			const int byte = 3 * j, wi = byte >> 2, sh = (byte & 3) * 8;
			o[j] = (__funnelshift_r(w[wi], w[wi + 1], sh) & 0x00FFFFFFu) | 0xFF000000u;
		}
#pragma unroll
		for(int k = 0; k < KOUT; ++k) stg256(dp + 2 * k, make_uint4(o[8 * k], o[8 * k + 1], o[8 * k + 2], o[8 * k + 3]), make_uint4(o[8 * k + 4], o[8 * k + 5], o[8 * k + 6], o[8 * k + 7]));
	}
}

struct Timer {
	cudaEvent_t a, b;
	Timer() {
		CK(cudaEventCreate(&a));
		CK(cudaEventCreate(&b));
	}
};

template<class F>
float time_ms(F launch, int reps = 10) {
	Timer t;
	for(int i = 0; i < 3; ++i) launch();
	CK(cudaDeviceSynchronize());
	float best = 1e30f, sum = 0;
	for(int i = 0; i < reps; ++i) {
		CK(cudaEventRecord(t.a));
		launch();
		CK(cudaEventRecord(t.b));
		CK(cudaEventSynchronize(t.b));
		float ms;
		CK(cudaEventElapsedTime(&ms, t.a, t.b));
		best = ms < best ? ms : best;
		sum += ms;
	}
	CK(cudaGetLastError());
	printf("  best %.4f ms  mean %.4f ms", best, sum / reps);
	return sum / reps;
}

int main() {
	const uint64_t in_bytes = 1920ull * 1080 * 4 * 256;
	const uint64_t n = in_bytes / 16;
	uint4 *src, *dst;
	CK(cudaMalloc(&src, in_bytes));
	CK(cudaMalloc(&dst, in_bytes * 2));
	CK(cudaMemset(src, 0x5A, in_bytes));
	CK(cudaMemset(dst, 0, in_bytes * 2));
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	printf("SMs %d, in %.2f GB\n", sms, in_bytes / 1e9);

	auto report = [&](const char* name, double bytes, float ms) { printf("  -> %-44s %8.1f GB/s\n", name, bytes / ms / 1e6); };

	// A: copy in -> first half of dst (2 x in_bytes traffic)
	for(int bps : {4, 8, 16, 32, 0}) {
		unsigned grid = bps ? sms * bps : (unsigned)((n + 256 * 4 - 1) / (256 * 4));
		char nm[96];
		snprintf(nm, sizeof nm, "copy U=4 grid=%u", grid);
		report(nm, 2.0 * in_bytes, time_ms([&] { copy_kernel<4><<<grid, 256>>>(src, dst, n); }));
	}
	report("copy U=8 grid=sms*8", 2.0 * in_bytes, time_ms([&] { copy_kernel<8><<<sms * 8, 256>>>(src, dst, n); }));
	report("copy U=2 grid=sms*8", 2.0 * in_bytes, time_ms([&] { copy_kernel<2><<<sms * 8, 256>>>(src, dst, n); }));
	report("cudaMemcpyAsync D2D", 2.0 * in_bytes, time_ms([&] { CK(cudaMemcpyAsync(dst, src, in_bytes, cudaMemcpyDeviceToDevice)); }));
	report("cudaMemsetAsync (write only, 2x)", 2.0 * in_bytes, time_ms([&] { CK(cudaMemsetAsync(dst, 1, in_bytes * 2)); }));

	// B/C: 1:2 expansion
	const double b3 = 3.0 * in_bytes;
	for(int bps : {4, 8, 16, 0}) {
		unsigned grid = bps ? sms * bps : (unsigned)((n + 256 * 4 - 1) / (256 * 4));
		char nm[96];
		snprintf(nm, sizeof nm, "expand 2xSTG.128 U=4 grid=%u", grid);
		report(nm, b3, time_ms([&] { expand_kernel<4, 0><<<grid, 256>>>(src, dst, n); }));
		snprintf(nm, sizeof nm, "expand STG.256 U=4 grid=%u", grid);
		report(nm, b3, time_ms([&] { expand_kernel<4, 1><<<grid, 256>>>(src, dst, n); }));
		snprintf(nm, sizeof nm, "convert STG.256 U=4 grid=%u", grid);
		report(nm, b3, time_ms([&] { expand_kernel<4, 2><<<grid, 256>>>(src, dst, n); }));
		snprintf(nm, sizeof nm, "convert 2xSTG.128 U=4 grid=%u", grid);
		report(nm, b3, time_ms([&] { expand_kernel<4, 3><<<grid, 256>>>(src, dst, n); }));
	}
	for(int bps : {4, 8, 0}) {
		unsigned grid8 = bps ? sms * bps : (unsigned)((n + 256 * 8 - 1) / (256 * 8));
		unsigned grid2 = bps ? sms * bps : (unsigned)((n + 256 * 2 - 1) / (256 * 2));
		unsigned grid1 = bps ? sms * bps : (unsigned)((n + 256 - 1) / (256));
		char nm[96];
		snprintf(nm, sizeof nm, "convert STG.256 U=8 grid=%u", grid8);
		report(nm, b3, time_ms([&] { expand_kernel<8, 2><<<grid8, 256>>>(src, dst, n); }));
		snprintf(nm, sizeof nm, "convert STG.256 U=2 grid=%u", grid2);
		report(nm, b3, time_ms([&] { expand_kernel<2, 2><<<grid2, 256>>>(src, dst, n); }));
		snprintf(nm, sizeof nm, "convert STG.256 U=1 grid=%u", grid1);
		report(nm, b3, time_ms([&] { expand_kernel<1, 2><<<grid1, 256>>>(src, dst, n); }));
	}
	for(int bps : {4, 8, 0}) {
		unsigned g4 = bps ? sms * bps : (unsigned)((n / 2 + 256 * 4 - 1) / (256 * 4));
		unsigned g2 = bps ? sms * bps : (unsigned)((n / 2 + 256 * 2 - 1) / (256 * 2));
		unsigned g1 = bps ? sms * bps : (unsigned)((n / 2 + 256 - 1) / (256));
		char nm[96];
		snprintf(nm, sizeof nm, "convert LDG.256+2xSTG.256 U=4 grid=%u", g4);
		report(nm, b3, time_ms([&] { expand256_kernel<4><<<g4, 256>>>(src, dst, n); }));
		snprintf(nm, sizeof nm, "convert LDG.256+2xSTG.256 U=2 grid=%u", g2);
		report(nm, b3, time_ms([&] { expand256_kernel<2><<<g2, 256>>>(src, dst, n); }));
		snprintf(nm, sizeof nm, "convert LDG.256+2xSTG.256 U=1 grid=%u", g1);
		report(nm, b3, time_ms([&] { expand256_kernel<1><<<g1, 256>>>(src, dst, n); }));
	}
	// E: lane-contiguous chunk access
	{
		auto run = [&](auto kern, int kin, int kout, const char* nm) {
			const uint64_t lanes = in_bytes / (32ull * kin);
			const unsigned grid = (unsigned)((lanes + 255) / 256);
			report(nm, (double)lanes * 32.0 * (kin + kout), time_ms([&] { kern<<<grid, 256>>>(src, dst, lanes); }));
		};
		run(lane_chunk_kernel<1, 1>, 1, 1, "lane-chunk copy 32 B/lane");
		run(lane_chunk_kernel<2, 2>, 2, 2, "lane-chunk copy 64 B/lane");
		run(lane_chunk_kernel<3, 3>, 3, 3, "lane-chunk copy 96 B/lane");
		run(lane_chunk_kernel<4, 4>, 4, 4, "lane-chunk copy 128 B/lane");
		run(lane_chunk_kernel<8, 8>, 8, 8, "lane-chunk copy 256 B/lane");
		run(lane_chunk_kernel<3, 4>, 3, 4, "lane-chunk RGB8->RGBA8 96->128 B/lane");
	}
	CK(cudaFree(src));
	CK(cudaFree(dst));
	return 0;
}
