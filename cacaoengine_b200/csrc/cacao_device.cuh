// cacao_device.cuh -- shared device-side helpers for the sm_100a pixel kernels.
//
// Compiled with -fmad=false: the float spec (DESIGN.md "Spec B") fixes where fused
// multiply-adds happen, so every FMA in this library is an explicit fmaf()/__fmaf_rn().
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "cacao_img.h"

namespace cacao {

	constexpr int kWarp = 32;

	// ---- sample-format traits ------------------------------------------------------------------
	template<int Fmt>
	struct FmtTraits;
	template<>
	struct FmtTraits<CACAO_FMT_U8> {
		using type = uint8_t;
		static constexpr int bytes = 1;
		static constexpr bool is_int = true;
	};
	template<>
	struct FmtTraits<CACAO_FMT_U16> {
		using type = uint16_t;
		static constexpr int bytes = 2;
		static constexpr bool is_int = true;
	};
	template<>
	struct FmtTraits<CACAO_FMT_F16> {
		using type = uint16_t; // raw binary16 bits
		static constexpr int bytes = 2;
		static constexpr bool is_int = false;
	};
	template<>
	struct FmtTraits<CACAO_FMT_F32> {
		using type = float;
		static constexpr int bytes = 4;
		static constexpr bool is_int = false;
	};

	__host__ __device__ constexpr int fmt_bytes(int fmt) {
		return fmt == CACAO_FMT_U8 ? 1 : (fmt == CACAO_FMT_F32 ? 4 : 2);
	}

	constexpr int cgcd(int a, int b) {
		return b == 0 ? a : cgcd(b, a % b);
	}
	constexpr int clcm(int a, int b) {
		return a / cgcd(a, b) * b;
	}

	// ---- 16-byte global / shared accessors -----------------------------------------------------
	// Streaming data is touched exactly once: keep it out of L1 (no_allocate) so the LUT and the
	// resampler's reused texels own the cache.
	__device__ __forceinline__ uint4 ldg_stream(const void* p) {
		uint4 r;
		asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
		return r;
	}
	__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
		asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
	}

	// 256-bit forms (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256): one whole 32-byte DRAM sector per lane,
	// p 32-byte aligned
	__device__ __forceinline__ void ldg256_stream(const void* p, uint32_t* w) {
		asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
	}
	__device__ __forceinline__ void stg256_stream(void* p, const uint32_t* w) {
		asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
	}

	// ---- programmatic dependent launch ---------------------------------------------------------
	// Kernels of this library are usually launched back to back on one stream (faces of a cube, images of a
	// batch, a conversion after a resampling).  Launched with launch_pdl() below, a kernel's blocks may become
	// resident while the previous kernel's last blocks are still running: everything that does not touch the
	// caller's buffers -- index arithmetic, the projection of the block's first pixels, staging the depth table --
	// then overlaps the predecessor's tail instead of following it.  pdl_wait() returns once the predecessor
	// has completed and its writes are visible, so it MUST precede the first access to any caller buffer;
	// pdl_launch_dependents() lets the successor in.  After a non-participating predecessor (a copy, another
	// library's kernel) both are no-ops and the launch is an ordinary one.
	__device__ __forceinline__ void pdl_launch_dependents() {
		asm volatile("griddepcontrol.launch_dependents;");
	}
	__device__ __forceinline__ void pdl_wait() {
		asm volatile("griddepcontrol.wait;" ::: "memory");
	}

	// ---- binary16 <-> binary32, round-to-nearest-even ------------------------------------------
	// The bits come out of the packed two-lane convert into a 32-bit register on purpose: ptxas
	// 12.9 miscompiles "cvt.rn.f16.f32 into a .b16 register, then store/extract its low byte" into
	// F2I.U8.F16 (a VALUE conversion of the half) -- seen in the byte-wise edge kernel.  A .b32
	// result keeps that pattern from forming.  tests/test_abi.py greps the SASS for it.
	__device__ __forceinline__ uint32_t f32x2_to_f16x2_bits(float lo, float hi) {
		uint32_t r;
		asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
		return r;
	}
	__device__ __forceinline__ uint32_t f32_to_f16_bits(float f) {
		return f32x2_to_f16x2_bits(f, 0.0f) & 0xFFFFu;
	}
	__device__ __forceinline__ float f16_bits_to_f32(uint16_t h) {
		return __half2float(__ushort_as_half(h));
	}

	// ---- hashing (synthetic inputs, checksum) --------------------------------------------------
	__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
		h ^= h >> 16;
		h *= 0x85EBCA6Bu;
		h ^= h >> 13;
		h *= 0xC2B2AE35u;
		h ^= h >> 16;
		return h;
	}
	__host__ __device__ __forceinline__ uint32_t hash32(uint32_t seed, uint64_t idx) {
		return fmix32(seed ^ (static_cast<uint32_t>(idx) * 0x9E3779B9u) ^ (static_cast<uint32_t>(idx >> 32) * 0x85EBCA6Bu));
	}

	// Host side: <<<>>> with the programmatic-stream-serialization attribute.  CACAO_NO_PDL=1 launches plainly
	// (A/B switch).  Only for kernels that call pdl_wait() before their first global access.
	template<class... KArgs, class... Args>
	inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
		const char* e = getenv("CACAO_NO_PDL");
		const bool off = e && e[0] == '1';
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = grid;
		cfg.blockDim = block;
		cfg.dynamicSmemBytes = smem;
		cfg.stream = stream;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr;
		cfg.numAttrs = off ? 0 : 1;
		return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
	}

} // namespace cacao
