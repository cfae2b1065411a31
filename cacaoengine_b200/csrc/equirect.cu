// equirect.cu -- equirectangular -> cubemap resampling kernel (DESIGN.md "Spec B.2").
//
// New functionality behind `ce-cubetool project`; the reference's cubetool only packs/unpacks six
// ready-made faces (tools/cubetool/create.cpp:32-102, extract.cpp:50-101).
//
// One thread = one output pixel and its mirror images (column mirror always, row mirror too when the
// launch covers whole faces); a warp covers an 8 x 4 pixel patch, a block 8 x 16, so the four
// bilinear taps of neighbouring threads fall on the same or adjacent source lines (L1 absorbs the 4x
// tap overlap, L2 the reuse between neighbouring patches, and HBM sees each source texel about once
// -- ncu: 0.52 GB read for a 0.54 GB source).
#include "equirect.cuh"

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <utility>
#include <vector>

#include "equirect_taps.cuh"

namespace cacao {

	namespace {


		// One thread = the pixel (i, j) and its mirror images.
		//
		// Mirroring a pixel across the face's vertical axis (i -> N-1-i) negates exactly one of the
		// direction's x / z components (face_coord is bit-exactly antisymmetric), so the pair shares
		// |x|, |z| and y: the two divisions, two polynomials and the square root of the projection are
		// computed ONCE, the latitude (source rows, fy) is common, and only the longitude's quadrant
		// placement differs.  With QUAD (launches with units that come as a band of the upper half plus
		// its mirror band: whole faces, or the two halves of a mirror pair that a sharded rank was
		// dealt -- pair_mirror_units() below finds them) the thread also owns the
		// row mirror j -> N-1-j: on the side faces that negates y, hence theta exactly (atan2 is odd in
		// y by construction) while the longitudes stay; on the +-Y faces it negates z, so the latitude
		// stays and only the quadrant changes again.  Four pixels per projection; every value is the
		// one the per-pixel spec computes, bit for bit.
		// 32 registers (16 blocks/SM) everywhere except the three-channel shapes, whose window realignment
		// needs 40 (12 blocks/SM measured 7 % ahead of spilling at 32: rgb8 0.165 -> 0.154 ms).
		// MODE: kPair = two-fold (column mirror); kQuadUnits = four-fold per unit flag, any mix of units,
		// flipped or converted output; kQuadFaces = every unit a whole face, stored as is -- the case of a
		// one-GPU cube, kept as its own instantiation so that nothing the other cases need (per-unit row
		// bounds and flags, the store row map) sits in its 32 registers: the general form measured 1.5 %
		// behind it on BASELINE configs 3 and 5 (dev/ab_lib_versions.py).
		// kQuadUnitsMapped = kQuadUnits with the store row map (flipped output); the fused-conversion
		// instantiations always carry the map.
		constexpr int kPair = 0, kQuadFaces = 1, kQuadUnits = 2, kQuadUnitsMapped = 3;
		template<int F, int CH, typename IDX, int MODE, int WX, int DF = F, int DC = CH>
		__global__ void __launch_bounds__(128, CH == 3 ? 12 : 16) equirect_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, EquirectParams prm) {
			constexpr bool QUAD = MODE != kPair;
			constexpr bool FACES = MODE == kQuadFaces;
			constexpr bool MAPPED = MODE == kQuadUnitsMapped || DF != F || DC != CH;
			static_assert(!FACES || (DF == F && DC == CH), "the whole-face form stores in the source format");
			constexpr int bpp = FmtTraits<DF>::bytes * DC; // of the output; the taps know the source's
			constexpr int RY = 32 / WX; // a warp covers WX columns x RY rows, a block WX x 4*RY
			// cacao_device.cuh: the launch itself and the block scheduling overlap the previous launch's tail.  (Waiting
			// later, after the projection, would overlap that too, but costs the 32-register forms a spill: measured
			// 2 % slower on RGBA32F.)
			pdl_launch_dependents();
			pdl_wait();
			const CubeUnit unit = prm.units[blockIdx.z];
			const uint32_t i = blockIdx.x * WX + (threadIdx.x % WX);
			const uint32_t face = FACES ? unit.face : (unit.face & 7u);
			uint32_t by = blockIdx.y;
			if constexpr(!FACES) {
				// Folded +-Y units run from the face's edge to its centre row, and the rows by the centre (the pole:
				// neighbouring pixels fan out over all longitudes, every tap its own line) cost up to 2.5x the rows by
				// the edge.  Blocks are dispatched in index order, so those units are walked centre first: in a sharded
				// rank's launch -- its whole job, ~0.09 ms -- the expensive blocks then start early instead of being the
				// tail (longest-processing-time-first, by construction of the index).
				if((unit.face & kUnitMirror) && (face & 6u) == 2u) {
					const uint32_t nb = (unit.row_end - unit.row_begin + 4u * RY - 1u) / (4u * RY);
					if(by >= nb) return;
					by = nb - 1u - by;
				}
			}
			const uint32_t j = unit.row_begin + (by * 4u + threadIdx.y) * RY + threadIdx.x / WX;
			const uint32_t half = (prm.N + 1u) >> 1;
			// a unit flagged kUnitMirror lies in the upper half of the face (row_end <= half, the launcher
			// sees to that) and stands for its rows AND their mirror rows N-1-j
			if(i >= half || j >= (FACES ? half : unit.row_end)) return;
			const uint32_t im = prm.N - 1u - i;

			const float sc = face_coord(i, prm.pc);
			const float tc = face_coord(j, prm.pc);
			float xa, ya, za;
			face_direction(face, sc, tc, xa, ya, za);
			// which component the column mirror negates: z on the +-X faces, x elsewhere
			const bool sc_is_z = face < 2u;
			const float xb = sc_is_z ? xa : -xa, zb = sc_is_z ? -za : za;

			// shared: first-quadrant longitude, latitude
			const float rphi = atan2_core(fabsf(za), fabsf(xa));
			const float rxz = sqrtf(fmaf(xa, xa, za * za));
			const float theta = atan2_spec(ya, rxz);
			const int W = static_cast<int>(prm.W);
			const IDX total_px = static_cast<IDX>(prm.src_rows) * static_cast<IDX>(prm.W);
			const RowTaps<IDX> rt = row_taps<IDX>(fmaf(-theta, prm.pc.ky, prm.pc.cy), prm);

			// face row j is stored at row row_add + row_mul * j: (0, 1) normally, (N-1, -1) for flipped output
			// (bottom row first) -- the multiply-add the row address needed anyway, so the mirror is free
			const uint32_t frow = face * prm.N + (MAPPED ? prm.row_add : 0u);
			uint8_t* orow = MAPPED ? dst + static_cast<size_t>(frow + j * prm.row_mul) * prm.N * bpp : dst + (static_cast<size_t>(face) * prm.N + j) * prm.N * bpp;
			const float xs_a = fmaf(atan2_quadrant(rphi, xa, za), prm.pc.kx, prm.pc.cx);
			const float xs_b = fmaf(atan2_quadrant(rphi, xb, zb), prm.pc.kx, prm.pc.cx);
			sample_and_store<F, CH, IDX, DF, DC>(src, orow + static_cast<size_t>(i) * bpp, xs_a, W, rt, total_px);
			if(im != i) sample_and_store<F, CH, IDX, DF, DC>(src, orow + static_cast<size_t>(im) * bpp, xs_b, W, rt, total_px);
			if constexpr(QUAD) {
				const uint32_t jm = prm.N - 1u - j;
				if constexpr(FACES) {
					if(jm == j) return;
				} else {
					if(!(unit.face & kUnitMirror) || jm == j) return; // uniform per block: units are indexed by blockIdx.z
				}
				uint8_t* mrow = MAPPED ? dst + static_cast<size_t>(frow + jm * prm.row_mul) * prm.N * bpp : dst + (static_cast<size_t>(face) * prm.N + jm) * prm.N * bpp;
				const bool polar = (face & 6u) == 2u; // +-Y: the row mirror negates z, theta stays
				const RowTaps<IDX> rm = polar ? rt : row_taps<IDX>(fmaf(theta, prm.pc.ky, prm.pc.cy), prm);
				const float xs_c = polar ? fmaf(atan2_quadrant(rphi, xa, -za), prm.pc.kx, prm.pc.cx) : xs_a;
				const float xs_d = polar ? fmaf(atan2_quadrant(rphi, xb, -zb), prm.pc.kx, prm.pc.cx) : xs_b;
				sample_and_store<F, CH, IDX, DF, DC>(src, mrow + static_cast<size_t>(i) * bpp, xs_c, W, rm, total_px);
				if(im != i) sample_and_store<F, CH, IDX, DF, DC>(src, mrow + static_cast<size_t>(im) * bpp, xs_d, W, rm, total_px);
			}
		}

		template<int F, int CH, int DF = F, int DC = CH>
		cudaError_t launch(const void* src, void* dst, const EquirectParams& prm, cudaStream_t stream) {
			constexpr bool fused = DF != F || DC != CH;
			cudaError_t pe = cudaSuccess;
			const uint32_t half = (prm.N + 1u) >> 1;
			uint32_t max_rows = 0;
			for(uint32_t u = 0; u < prm.n_units; ++u) max_rows = std::max(max_rows, prm.units[u].row_end - prm.units[u].row_begin);
			constexpr int bpp = FmtTraits<F>::bytes * CH;
			const bool small = static_cast<uint64_t>(prm.src_rows) * prm.W < (1ull << 32);
			// gather kernel: 128-thread blocks (16 per SM at 32 registers), four warps stacked vertically.
			// A warp covers WX columns x 32/WX rows.  Square-ish footprints (8 x 4) touch fewer distinct
			// source lines per tap instruction than a 32 x 1 run, most of all on the +-Y faces where
			// neighbouring columns fan out over the longitudes (dev/sweep_eq_shape.py on B200: cfg3
			// 0.180 -> 0.169 ms, cfg5 0.682 -> 0.657 ms; 16 x 2 is level, 4 x 8 loses again).  WX stays
			// wide enough for a warp's store to fill whole 32-byte sectors.
			constexpr int WX = bpp >= 4 ? 8 : (bpp >= 2 ? 16 : 32);
			const dim3 block(32, 4, 1);
			// units that stand for a band and its mirror band (flagged by pair_mirror_units) need the
			// four-fold kernel; it also serves unflagged units of the same launch two-fold
			bool quad = false;
			for(uint32_t u = 0; u < prm.n_units; ++u) quad = quad || (prm.units[u].face & kUnitMirror) != 0;
			const uint32_t rows = max_rows;
			constexpr uint32_t bh = 4 * (32 / WX);
			const dim3 ggrid((half + WX - 1) / WX, (rows + bh - 1) / bh, prm.n_units);
			const uint8_t* s8 = static_cast<const uint8_t*>(src);
			uint8_t* d8 = static_cast<uint8_t*>(dst);
			// every unit the folded upper half of a whole face, stored unflipped: the dedicated form
			bool faces = quad && prm.row_mul == 1u;
			for(uint32_t u = 0; u < prm.n_units; ++u) faces = faces && (prm.units[u].face & kUnitMirror) && prm.units[u].row_begin == 0 && prm.units[u].row_end == half;
			if constexpr(fused) {
				// one instantiation per index width: the four-fold kernel serves unflagged units two-fold
				if(small)
					pe = launch_pdl(equirect_kernel<F, CH, uint32_t, kQuadUnits, WX, DF, DC>, ggrid, block, 0, stream, s8, d8, prm);
				else
					pe = launch_pdl(equirect_kernel<F, CH, uint64_t, kQuadUnits, WX, DF, DC>, ggrid, block, 0, stream, s8, d8, prm);
			} else if(faces) {
				EquirectParams whole = prm; // plain face numbers, as that form reads them
				for(uint32_t u = 0; u < whole.n_units; ++u) whole.units[u].face &= 7u;
#ifdef CACAO_ROWS_SWEEP // development build: warp footprint of whole-face launches selectable at run time (dev/sweep_rows.py --wx)
				if(const char* we = getenv("CACAO_EQUIRECT_WX")) {
					const int wx = atoi(we);
					if(small && (wx == 8 || wx == 16 || wx == 32) && wx != WX) {
						const uint32_t sbh = 4u * (32u / static_cast<uint32_t>(wx));
						const dim3 sg((half + wx - 1) / wx, (rows + sbh - 1) / sbh, prm.n_units);
						if(wx == 8) pe = launch_pdl(equirect_kernel<F, CH, uint32_t, kQuadFaces, 8>, sg, block, 0, stream, s8, d8, whole);
						if(wx == 16) pe = launch_pdl(equirect_kernel<F, CH, uint32_t, kQuadFaces, 16>, sg, block, 0, stream, s8, d8, whole);
						if(wx == 32) pe = launch_pdl(equirect_kernel<F, CH, uint32_t, kQuadFaces, 32>, sg, block, 0, stream, s8, d8, whole);
						return pe;
					}
				}
#endif
				if(small)
					pe = launch_pdl(equirect_kernel<F, CH, uint32_t, kQuadFaces, WX>, ggrid, block, 0, stream, s8, d8, whole);
				else
					pe = launch_pdl(equirect_kernel<F, CH, uint64_t, kQuadFaces, WX>, ggrid, block, 0, stream, s8, d8, whole);
			} else if(prm.row_mul != 1u) { // flipped output: the mapped form (it serves unflagged units two-fold)
				if(small)
					pe = launch_pdl(equirect_kernel<F, CH, uint32_t, kQuadUnitsMapped, WX>, ggrid, block, 0, stream, s8, d8, prm);
				else
					pe = launch_pdl(equirect_kernel<F, CH, uint64_t, kQuadUnitsMapped, WX>, ggrid, block, 0, stream, s8, d8, prm);
			} else if(small) {
				if(quad)
					pe = launch_pdl(equirect_kernel<F, CH, uint32_t, kQuadUnits, WX>, ggrid, block, 0, stream, s8, d8, prm);
				else
					pe = launch_pdl(equirect_kernel<F, CH, uint32_t, kPair, WX>, ggrid, block, 0, stream, s8, d8, prm);
			} else {
				if(quad)
					pe = launch_pdl(equirect_kernel<F, CH, uint64_t, kQuadUnits, WX>, ggrid, block, 0, stream, s8, d8, prm);
				else
					pe = launch_pdl(equirect_kernel<F, CH, uint64_t, kPair, WX>, ggrid, block, 0, stream, s8, d8, prm);
			}
			return pe;
		}

		// Fused resample + convert (Spec B.4): float sources, any destination format, same channels or
		// RGB <-> RGBA.
		template<int F, int CH, int DC>
		cudaError_t launch_cvt_fmt(const void* src, void* dst, int dst_fmt, const EquirectParams& prm, cudaStream_t stream) {
			switch(dst_fmt) {
				case CACAO_FMT_U8: return launch<F, CH, CACAO_FMT_U8, DC>(src, dst, prm, stream);
				case CACAO_FMT_U16: return launch<F, CH, CACAO_FMT_U16, DC>(src, dst, prm, stream);
				case CACAO_FMT_F16: return launch<F, CH, CACAO_FMT_F16, DC>(src, dst, prm, stream);
				case CACAO_FMT_F32: return launch<F, CH, CACAO_FMT_F32, DC>(src, dst, prm, stream);
			}
			return cudaErrorInvalidValue;
		}
		template<int F>
		cudaError_t launch_cvt(const void* src, void* dst, int ch, int dst_ch, int dst_fmt, const EquirectParams& prm, cudaStream_t stream) {
			if(ch == 1 && dst_ch == 1) return launch_cvt_fmt<F, 1, 1>(src, dst, dst_fmt, prm, stream);
			if(ch == 3 && dst_ch == 3) return launch_cvt_fmt<F, 3, 3>(src, dst, dst_fmt, prm, stream);
			if(ch == 3 && dst_ch == 4) return launch_cvt_fmt<F, 3, 4>(src, dst, dst_fmt, prm, stream);
			if(ch == 4 && dst_ch == 4) return launch_cvt_fmt<F, 4, 4>(src, dst, dst_fmt, prm, stream);
			if(ch == 4 && dst_ch == 3) return launch_cvt_fmt<F, 4, 3>(src, dst, dst_fmt, prm, stream);
			return cudaErrorInvalidValue;
		}

		template<int F>
		cudaError_t launch_ch(const void* src, void* dst, int ch, const EquirectParams& prm, cudaStream_t stream) {
			switch(ch) {
				case 1: return launch<F, 1>(src, dst, prm, stream);
				case 3: return launch<F, 3>(src, dst, prm, stream);
				case 4: return launch<F, 4>(src, dst, prm, stream);
			}
			return cudaErrorInvalidValue;
		}
	}

	// Rows j and N-1-j of a face share their whole projection (equirect_kernel), so a band that arrives
	// together with its mirror band -- a whole face, a band around the centre row, the two units of a
	// mirror pair in a sharded rank's list, the two polar bands with the same source window in the host
	// pipeline -- is folded into ONE unit of the upper half flagged kUnitMirror.  Rows whose mirror is
	// not in the list stay plain units, and so do overlaps of fewer than 16 rows (one block's height):
	// a sliver would only add a launch.  Every (face, row) of the input is covered exactly as often as
	// before; only who computes it changes, and the values are bit-identical either way.
	std::vector<CubeUnit> pair_mirror_units(const std::vector<CubeUnit>& in, uint32_t N) {
		using Iv = std::pair<uint32_t, uint32_t>;
		const uint32_t half = (N + 1u) >> 1; // upper half incl. the centre row of an odd face
		const uint32_t min_fold = std::min(16u, half);
		// sort by first row, then join intervals that touch (never ones that overlap: rows the caller
		// names twice are computed twice, as before)
		auto join = [](std::vector<Iv>& v) {
			std::sort(v.begin(), v.end());
			std::vector<Iv> r;
			for(const Iv& iv : v) {
				bool done = false;
				for(Iv& q : r)
					if(q.second == iv.first) {
						q.second = iv.second;
						done = true;
						break;
					}
				if(!done) r.push_back(iv);
			}
			v.swap(r);
		};
		std::vector<CubeUnit> out;
		for(uint32_t f = 0; f < 6; ++f) {
			// per face: intervals of upper rows present as themselves (A) and as mirror images of present
			// lower rows (B), both in upper-half coordinates
			std::vector<Iv> A, B;
			std::vector<uint32_t> cuts;
			for(const CubeUnit& u : in) {
				if(u.face != f || u.row_begin >= u.row_end) continue;
				if(u.row_begin < half) A.push_back({u.row_begin, std::min(u.row_end, half)});
				// the centre row of an odd face is its own mirror: N - half == half - 1 puts it in B as well
				const uint32_t lb = std::max(u.row_begin, N - half);
				if(u.row_end > lb) B.push_back({N - u.row_end, N - lb});
			}
			for(const Iv& iv : A) cuts.push_back(iv.first), cuts.push_back(iv.second);
			for(const Iv& iv : B) cuts.push_back(iv.first), cuts.push_back(iv.second);
			std::sort(cuts.begin(), cuts.end());
			cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
			auto count = [](const std::vector<Iv>& v, uint32_t s) {
				uint32_t n = 0;
				for(const Iv& iv : v) n += (iv.first <= s && s < iv.second) ? 1u : 0u;
				return n;
			};
			std::vector<Iv> both, upper, lower; // all in upper-half coordinates
			for(size_t k = 0; k + 1 < cuts.size(); ++k) {
				const Iv seg{cuts[k], cuts[k + 1]};
				uint32_t a = count(A, seg.first), b = count(B, seg.first);
				for(; a > 0 && b > 0; --a, --b) both.push_back(seg);
				for(; a > 0; --a) upper.push_back(seg);
				for(; b > 0; --b) lower.push_back(seg);
			}
			join(both);
			for(size_t k = 0; k < both.size();) {
				if(both[k].second - both[k].first < min_fold) {
					upper.push_back(both[k]);
					// an odd face's centre row was listed on both sides only as its own mirror
					const uint32_t e = ((N & 1u) && both[k].second == half) ? half - 1u : both[k].second;
					if(e > both[k].first) lower.push_back({both[k].first, e});
					both.erase(both.begin() + static_cast<std::ptrdiff_t>(k));
				} else
					++k;
			}
			// plain rows in face coordinates, joined across the centre line
			for(const Iv& iv : lower) upper.push_back({N - iv.second, N - iv.first});
			join(upper);
			for(const Iv& iv : both) out.push_back({f | kUnitMirror, iv.first, iv.second});
			for(const Iv& iv : upper) out.push_back({f, iv.first, iv.second});
		}
		return out;
	}

	bool equirect_cvt_supported(int ch, int fmt, int dst_ch, int dst_fmt) {
		if(dst_ch == ch && dst_fmt == fmt) return true;
		const bool float_src = fmt == CACAO_FMT_F16 || fmt == CACAO_FMT_F32;
		return float_src && (dst_ch == ch || (ch == 3 && dst_ch == 4) || (ch == 4 && dst_ch == 3));
	}

	cudaError_t equirect_launch_units(const void* src, void* dst, int ch, int fmt, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, uint32_t N, const CubeUnit* units, uint32_t n_units, cudaStream_t stream, unsigned* launches, int dst_ch, int dst_fmt, uint32_t flags) {
		if(launches) *launches = 0;
		if(dst_ch <= 0) dst_ch = ch;
		if(dst_fmt < 0) dst_fmt = fmt;
		if(!equirect_cvt_supported(ch, fmt, dst_ch, dst_fmt)) return cudaErrorInvalidValue;
		const bool fused = dst_ch != ch || dst_fmt != fmt;
		EquirectParams prm;
		prm.W = W;
		prm.H = H;
		prm.N = N;
		prm.src_row0 = src_row0;
		prm.src_rows = src_rows;
		const bool flip_rows = (flags & CACAO_EQUIRECT_FLIP_ROWS) != 0;
		prm.row_add = flip_rows ? N - 1u : 0u;
		prm.row_mul = flip_rows ? 0xFFFFFFFFu : 1u; // -1 in the kernel's modular row arithmetic
		prm.pc = make_proj_const(N, W, H);
		// The grid's y extent is that of the tallest unit in a launch, and blocks beyond a shorter unit's
		// rows exit at once but are not free (a sharded 16384x8192 job whose +-Y bands are half as tall as
		// its side-face bands lost 5 % to them on two ranks).  So the units of a call are first cut to a
		// common height -- the shortest unit's, when the heights are within a factor of eight -- then
		// sorted tallest first, and a launch takes up to kMaxUnitsPerLaunch units no shorter than 4/5
		// of its first one.  A rank's mixed list of side-face and half-height polar bands then still is
		// ONE launch without empty blocks.
		std::vector<CubeUnit> todo;
		for(uint32_t k = 0; k < n_units; ++k) {
			CubeUnit u = units[k];
			if(u.row_end > N) u.row_end = N;
			if(u.face > 5 || u.row_begin >= u.row_end) continue;
			todo.push_back(u);
		}
		// test / A-B switch: two-fold sharing only
		const char* qe = getenv("CACAO_EQUIRECT_NO_QUAD");
		if(!(qe && qe[0] == '1')) todo = pair_mirror_units(todo, N);
		// which kernel: the row-strip kernel where it measured ahead (one-channel and 8-bit texels: those shapes are
		// bound by instruction issue, which is what it saves), the gather kernel everywhere else
		bool use_rows = equirect_rows_supported(ch, fmt, dst_ch, dst_fmt, W, H) && equirect_rows_preferred(ch, fmt);
		if(const char* re = getenv("CACAO_EQUIRECT_ROWS")) use_rows = re[0] == '1' ? equirect_rows_supported(ch, fmt, dst_ch, dst_fmt, W, H) : (re[0] == '0' ? false : use_rows);
		int sm_count = 148;
		if(use_rows) {
			static int cached[64] = {};
			int dev = 0;
			if(cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
				if(cached[dev] == 0) {
					int n = 0;
					if(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached[dev] = n;
				}
				if(cached[dev] > 0) sm_count = cached[dev];
			}
		}
		// the tensor-map TMA experiment: a whole cube (six folded whole faces), stored as is
		if(const char* te = getenv("CACAO_EQUIRECT_TMA")) {
			bool whole = te[0] == '1' && !fused && !flip_rows && todo.size() == 6 && equirect_tma_supported(ch, fmt, W, H, src);
			uint32_t seen = 0;
			for(const CubeUnit& u : todo) {
				whole = whole && (u.face & kUnitMirror) && u.row_begin == 0 && u.row_end == ((N + 1u) >> 1);
				seen |= 1u << (u.face & 7u);
			}
			if(whole && seen == 0x3Fu) {
				int dev = 0, sms = 148;
				if(cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
				prm.n_units = 0;
				cudaError_t e = equirect_tma_launch(src, dst, ch, fmt, prm, sms, stream);
				if(e == cudaSuccess && launches) *launches = 1;
				if(e != cudaErrorNotSupported) return e;
			}
		}
		uint32_t hmin = 0xFFFFFFFFu, hmax = 0;
		for(const CubeUnit& u : todo) {
			hmin = std::min(hmin, u.row_end - u.row_begin);
			hmax = std::max(hmax, u.row_end - u.row_begin);
		}
		// gridDim.y is limited to 65535 blocks of at least 4 rows
		constexpr uint32_t kMaxRows = 65535u * 4u;
		uint32_t piece = hmax; // target height
		if(hmax > hmin && hmin >= 16 && hmax / hmin < 8) piece = hmin;
		if(piece > kMaxRows) piece = kMaxRows;
		if(piece < hmax) {
			std::vector<CubeUnit> cut;
			for(const CubeUnit& u : todo) {
				const uint32_t rows = u.row_end - u.row_begin;
				const uint32_t parts = std::max(1u, (rows + piece / 2) / piece);
				const uint32_t each = (rows + parts - 1) / parts;
				for(uint32_t r = u.row_begin; r < u.row_end; r += each) cut.push_back({u.face, r, std::min(u.row_end, r + each)});
			}
			todo.swap(cut);
		}
		// tallest first (launches group similar heights); among equals the +-Y units first: their blocks are the
		// expensive ones and blockIdx.z is the slowest dispatch index, so they start ahead of the side faces
		auto polar = [](const CubeUnit& u) { return ((u.face & 6u) == 2u) ? 1 : 0; };
		static const bool lpt = !(getenv("CACAO_EQUIRECT_NO_LPT") && getenv("CACAO_EQUIRECT_NO_LPT")[0] == '1');
		std::stable_sort(todo.begin(), todo.end(), [&](const CubeUnit& x, const CubeUnit& y) {
			const uint32_t hx = x.row_end - x.row_begin, hy = y.row_end - y.row_begin;
			if(hx != hy) return hx > hy;
			return lpt && polar(x) > polar(y);
		});
		size_t next = 0;
		while(next < todo.size()) {
			prm.n_units = 0;
			const uint64_t tallest = todo[next].row_end - todo[next].row_begin;
			while(next < todo.size() && prm.n_units < kMaxUnitsPerLaunch && 5ull * (todo[next].row_end - todo[next].row_begin) >= 4ull * tallest) prm.units[prm.n_units++] = todo[next++];
			cudaError_t e;
			if(use_rows) {
				e = equirect_rows_launch(src, dst, ch, fmt, prm, sm_count, stream);
			} else if(fused) {
				e = fmt == CACAO_FMT_F32 ? launch_cvt<CACAO_FMT_F32>(src, dst, ch, dst_ch, dst_fmt, prm, stream) : launch_cvt<CACAO_FMT_F16>(src, dst, ch, dst_ch, dst_fmt, prm, stream);
			} else {
				switch(fmt) {
					case CACAO_FMT_U8: e = launch_ch<CACAO_FMT_U8>(src, dst, ch, prm, stream); break;
					case CACAO_FMT_U16: e = launch_ch<CACAO_FMT_U16>(src, dst, ch, prm, stream); break;
					case CACAO_FMT_F16: e = launch_ch<CACAO_FMT_F16>(src, dst, ch, prm, stream); break;
					case CACAO_FMT_F32: e = launch_ch<CACAO_FMT_F32>(src, dst, ch, prm, stream); break;
					default: return cudaErrorInvalidValue;
				}
			}
			if(e != cudaSuccess) return e;
			if(launches) *launches += 1;
		}
		return cudaSuccess;
	}

	cudaError_t equirect_launch(const void* src, void* dst, int ch, int fmt, uint32_t W, uint32_t H, uint32_t src_row0, uint32_t src_rows, uint32_t N, uint32_t face_mask, uint32_t row_begin, uint32_t row_end, cudaStream_t stream) {
		CubeUnit units[6];
		uint32_t n = 0;
		for(uint32_t f = 0; f < 6; ++f)
			if(face_mask & (1u << f)) units[n++] = {f, row_begin, row_end};
		return equirect_launch_units(src, dst, ch, fmt, W, H, src_row0, src_rows, N, units, n, stream, nullptr, ch, fmt);
	}

	// Host-side planning with the same arithmetic as the kernel, in O(1) evaluations per face.
	// On a side face theta = atan2(-tc, sqrt(1 + sc^2)) is monotone in tc for any sc and, for a given
	// tc, monotone in |sc|: over a full-width row band its extremes lie on the band's first / last row
	// at the centre or edge columns.  On the +-Y faces theta is a monotone function of sc^2 + tc^2: the
	// extremes are the pixel nearest the face centre (centre columns; the centre row if the band
	// contains it, else the nearer border row) and a corner of the band.  fp32 rounding moves ys by
	// ~1e-3 rows at most; the window carries a whole row of slack on each side.
	void equirect_src_window(uint32_t W, uint32_t H, uint32_t N, uint32_t face_mask, uint32_t row_begin, uint32_t row_end, uint32_t* src_row0, uint32_t* src_rows) {
		const ProjConst pc = make_proj_const(N, W, H);
		float lo = 3.0e38f, hi = -3.0e38f;
		auto visit = [&](uint32_t f, uint32_t i, uint32_t j) {
			float xs, ys;
			project_pixel(f, i, j, pc, xs, ys);
			lo = ys < lo ? ys : lo;
			hi = ys > hi ? ys : hi;
		};
		if(row_end > N) row_end = N;
		for(uint32_t f = 0; f < 6; ++f) {
			if(!(face_mask & (1u << f)) || row_begin >= row_end) continue;
			const uint32_t cols[4] = {0, (N - 1) / 2, N / 2, N - 1};
			uint32_t rows[4] = {row_begin, row_end - 1, row_begin, row_end - 1};
			if((N - 1) / 2 >= row_begin && (N - 1) / 2 < row_end) rows[2] = (N - 1) / 2;
			if(N / 2 >= row_begin && N / 2 < row_end) rows[3] = N / 2;
			for(uint32_t j : rows)
				for(uint32_t i : cols) visit(f, i, j);
		}
		if(hi < lo) {
			*src_row0 = 0;
			*src_rows = 0;
			return;
		}
		// floor(lo) .. floor(hi)+1 are the rows the taps touch; one more row of slack each side
		long long first = static_cast<long long>(floorf(lo)) - 1;
		long long last = static_cast<long long>(floorf(hi)) + 2;
		if(first < 0) first = 0;
		if(last > static_cast<long long>(H) - 1) last = static_cast<long long>(H) - 1;
		if(last < first) last = first;
		*src_row0 = static_cast<uint32_t>(first);
		*src_rows = static_cast<uint32_t>(last - first + 1);
	}

} // namespace cacao
