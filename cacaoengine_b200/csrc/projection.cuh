// projection.cuh -- cube-face pixel -> equirect sample position, DESIGN.md "Spec B.2".
//
// No reference counterpart (tools/cubetool/main.cpp:31-33 has only create/extract).  What the
// reference does constrain: face order +X,-X,+Y,-Y,+Z,-Z = right,left,up,down,front,back
// (libs/formats/include/libcacaoformats.hpp:213, tools/cubetool/commands.hpp:31) and top-row-first
// storage consumed by Vulkan without a flip (engine/src/vulkan/impls/VulkanCubemap.cpp:30-34).
//
// The whole chain is fp32 with a FIXED operation order: IEEE add/mul/div/sqrt plus fused
// multiply-adds only where written as fmaf().  The library is compiled with -fmad=false (device)
// so the compiler fuses nothing else; the host planner below runs the same code on the CPU.  The
// oracle restates this spec independently in C (oracle/cacao_oracle.c), and the two agree bit for
// bit in (xs, ys) -- so the filtered faces agree bit for bit too.
#pragma once

#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define CACAO_HD __host__ __device__ __forceinline__
#else
#define CACAO_HD inline
#endif

namespace cacao {

	constexpr float kPi = 3.14159265358979323846f;
	constexpr float kPi2 = 1.57079632679489661923f;
	constexpr float kPi4 = 0.78539816339744830962f;
	constexpr float kInv2Pi = 0.15915494309189533577f;
	constexpr float kInvPi = 0.31830988618379067154f;
	constexpr float kTanPi8 = 0.41421356237309504880f;

	// atan2 with one division, in two steps so that callers can share the expensive one.
	//
	// atan2_core(|x|, |y|): the angle of (|x|, |y|) in [0, pi/2].  Reduce to t in [-(sqrt2-1), sqrt2-1]
	// via atan(a/b) = pi/4 + atan((a-b)/(a+b)), then a 4-term odd polynomial (the classic
	// single-precision coefficient set); |error| <= ~1.1 ulp of pi against libm over the cube's directions.
	CACAO_HD float atan2_core(float ax, float ay) {
		const float a = ax < ay ? ax : ay;
		const float b = ax < ay ? ay : ax;
		if(b == 0.0f) return 0.0f;
		const bool hi = a > kTanPi8 * b;
		const float num = hi ? a - b : a;
		const float den = hi ? a + b : b;
		const float off = hi ? kPi4 : 0.0f;
		const float t = num / den;
		const float z = t * t;
		float p = fmaf(8.05374449538e-2f, z, -1.38776856032e-1f);
		p = fmaf(p, z, 1.99777106478e-1f);
		p = fmaf(p, z, -3.33329491539e-1f);
		float r = fmaf(p * z, t, t);
		r = r + off;
		if(ay > ax) r = kPi2 - r;
		return r;
	}
	// atan2_quadrant: place the first-quadrant angle by the signs of x and y.  Mirrored cube pixels
	// differ only here, which is what lets the kernel compute atan2_core once per mirrored pair.
	CACAO_HD float atan2_quadrant(float r, float y, float x) {
		if(x < 0.0f) r = kPi - r;
		if(y < 0.0f) r = -r;
		return r;
	}
	CACAO_HD float atan2_spec(float y, float x) {
		return atan2_quadrant(atan2_core(fabsf(x), fabsf(y)), y, x);
	}

	struct ProjConst {
		int32_t n;    // N
		float invN;   // 1.0f / N
		float kx, ky; // W/(2 pi), H/pi
		float cx, cy; // W/2 - 1/2, H/2 - 1/2
	};

	CACAO_HD ProjConst make_proj_const(uint32_t N, uint32_t W, uint32_t H) {
		ProjConst c;
		c.n = static_cast<int32_t>(N);
		c.invN = 1.0f / static_cast<float>(N);
		c.kx = static_cast<float>(W) * kInv2Pi;
		c.ky = static_cast<float>(H) * kInvPi;
		c.cx = 0.5f * static_cast<float>(W) - 0.5f;
		c.cy = 0.5f * static_cast<float>(H) - 0.5f;
		return c;
	}

	// Face-plane coordinate of pixel index k in [-1, 1]: (2k + 1 - N) / N, computed as an exact
	// integer times 1/N so that face_coord(N-1-k) == -face_coord(k) bit for bit (mirror symmetry).
	CACAO_HD float face_coord(uint32_t k, const ProjConst& c) {
		return static_cast<float>(static_cast<int32_t>(2u * k + 1u) - c.n) * c.invN;
	}

	// Khronos cube-map face table
	CACAO_HD void face_direction(uint32_t face, float sc, float tc, float& x, float& y, float& z) {
		switch(face) {
			case 0: x = 1.0f; y = -tc; z = -sc; break;   // +X right
			case 1: x = -1.0f; y = -tc; z = sc; break;   // -X left
			case 2: x = sc; y = 1.0f; z = tc; break;     // +Y up
			case 3: x = sc; y = -1.0f; z = -tc; break;   // -Y down
			case 4: x = sc; y = -tc; z = 1.0f; break;    // +Z front
			default: x = -sc; y = -tc; z = -1.0f; break; // -Z back
		}
	}

	// The same table without a jump: selects on the (warp-uniform) face number and negations only, so
	// every component has the bits the switch above produces, signed zeros included (equirect_rows.cu).
	CACAO_HD void face_direction_sel(uint32_t face, float sc, float tc, float& x, float& y, float& z) {
		const bool odd = (face & 1u) != 0;
		const uint32_t axis = face >> 1; // 0: +-X, 1: +-Y, 2: +-Z
		const float one = odd ? -1.0f : 1.0f;
		const float nsc = -sc, ntc = -tc;
		x = axis == 0u ? one : (face == 5u ? nsc : sc);
		y = axis == 1u ? one : ntc;
		z = axis == 0u ? (odd ? sc : nsc) : (axis == 1u ? (odd ? ntc : tc) : one);
	}

	// face pixel (column i, row j) -> continuous source position (texel centres at integer + 0.5
	// already removed: xs, ys index texel centres directly)
	CACAO_HD void project_pixel(uint32_t face, uint32_t i, uint32_t j, const ProjConst& c, float& xs, float& ys) {
		const float sc = face_coord(i, c);
		const float tc = face_coord(j, c);
		float x, y, z;
		face_direction(face, sc, tc, x, y, z);
		const float phi = atan2_spec(x, z);
		const float rxz = sqrtf(fmaf(x, x, z * z));
		const float theta = atan2_spec(y, rxz);
		xs = fmaf(phi, c.kx, c.cx);
		ys = fmaf(-theta, c.ky, c.cy);
	}

} // namespace cacao
