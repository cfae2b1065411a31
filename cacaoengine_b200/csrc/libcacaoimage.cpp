// libcacaoimage.cpp -- the reference's C++ image API (include/libcacaoimage.hpp) and its extensions
// (include/libcacaoimage_ext.hpp) implemented over the C ABI in cacao_img.h.
//
// Mirrors the reference's conventions: inputs by const&, results by value owning a fresh
// std::vector, metadata copied through (libs/image/src/Layout.cpp:12-13, Colordepth.cpp:11-18),
// failures as std::runtime_error carrying the reference's message texts
// (libs/common/include/libcacaocommon.hpp:20-25).
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>

#include "cacao_img.h"
#include "libcacaoimage.hpp"
#include "libcacaoimage_ext.hpp"

namespace {
	// status -> exception.  Argument errors that the reference reports get the reference's text;
	// everything else (CUDA, no device) carries the C layer's detailed message.
	void check(int status) {
		if(status == CACAO_OK) return;
		switch(status) {
			case CACAO_E_SAME_LAYOUT:
			case CACAO_E_NOT_16BIT:
			case CACAO_E_ZERO_DIM:
			case CACAO_E_BAD_DEPTH:
			case CACAO_E_EMPTY: throw std::runtime_error(cacao_img_status_string(status));
			default: throw std::runtime_error(std::string("libcacaoimage-b200: ") + cacao_img_last_error());
		}
	}

	int fmt_of_bits(uint8_t bits) {
		return bits == 16 ? CACAO_FMT_U16 : CACAO_FMT_U8;
	}
	std::size_t bytes_of(int fmt) {
		return fmt == CACAO_FMT_U8 ? 1 : (fmt == CACAO_FMT_F32 ? 4 : 2);
	}
}

namespace libcacaoimage {

	Image Convert16To8BitColor(const Image& src) {
		if(src.bitsPerChannel != 16) check(CACAO_E_NOT_16BIT); // Colordepth.cpp:8
		Image out = {};
		out.w = src.w;
		out.h = src.h;
		out.format = src.format;
		out.layout = src.layout;
		out.bitsPerChannel = 8;
		out.lossy = src.lossy;
		out.quality = src.quality;
		// the reference sizes the result from the buffer, not from w*h (Colordepth.cpp:21-23)
		const int ch = static_cast<int>(src.layout);
		out.data.resize(src.data.size() / 2);
		const uint64_t pixels = src.data.size() / 2 / ch;
		if(pixels) check(cacao_img_convert(-1, src.data.data(), out.data.data(), pixels, ch, ch, CACAO_FMT_U16, CACAO_FMT_U8, CACAO_MODE_REF_EXACT, CACAO_MEM_HOST, nullptr));
		return out;
	}

	Image ChangeChannelLayout(const Image& src, Image::Layout layout) {
		if(src.layout == layout) check(CACAO_E_SAME_LAYOUT); // Layout.cpp:9
		Image out = {};
		out.w = src.w;
		out.h = src.h;
		out.layout = layout;
		out.bitsPerChannel = src.bitsPerChannel;
		out.format = src.format;
		out.quality = src.quality;
		out.lossy = src.lossy;
		const int fmt = fmt_of_bits(src.bitsPerChannel);
		const uint64_t pixels = static_cast<uint64_t>(src.w) * src.h;
		// the reference walks w*h pixels whatever the buffer holds (Layout.cpp:20,31) and would read past
		// a short one; that is the one place this wrapper is stricter
		if(src.data.size() < pixels * static_cast<int>(src.layout) * bytes_of(fmt)) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
		out.data.resize(pixels * static_cast<int>(layout) * bytes_of(fmt));
		if(pixels) check(cacao_img_convert(-1, src.data.data(), out.data.data(), pixels, static_cast<int>(src.layout), static_cast<int>(layout), fmt, fmt, CACAO_MODE_REF_EXACT, CACAO_MEM_HOST, nullptr));
		return out;
	}

	Image Flip(const Image& src) {
		// validation order and texts: Flip.cpp:8-10
		if(!(src.w > 0 && src.h > 0)) check(CACAO_E_ZERO_DIM);
		if(!(src.bitsPerChannel == 8 || src.bitsPerChannel == 16)) check(CACAO_E_BAD_DEPTH);
		if(src.data.empty()) check(CACAO_E_EMPTY);
		if(src.h == 1) return src; // Flip.cpp:13
		Image out = {};
		out.w = src.w;
		out.h = src.h;
		out.layout = src.layout;
		out.bitsPerChannel = src.bitsPerChannel;
		out.format = src.format;
		out.quality = src.quality;
		out.lossy = src.lossy;
		out.data.resize(src.data.size());
		const uint32_t bpp = static_cast<uint32_t>(src.layout) * (src.bitsPerChannel / 8);
		if(src.data.size() < static_cast<std::size_t>(src.w) * src.h * bpp) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
		check(cacao_img_flip_rows(-1, src.data.data(), out.data.data(), src.w, src.h, bpp, CACAO_MEM_HOST, nullptr));
		return out;
	}

	// ---- extensions ----------------------------------------------------------------------------

	std::size_t BytesPerPixel(Image::Layout layout, SampleFormat format) {
		return static_cast<std::size_t>(layout) * bytes_of(static_cast<int>(format));
	}

	ImageEx FromImage(const Image& src) {
		ImageEx out;
		out.w = src.w;
		out.h = src.h;
		out.layout = src.layout;
		out.format = src.bitsPerChannel == 16 ? SampleFormat::U16 : SampleFormat::U8;
		out.data = src.data;
		return out;
	}

	Image ToImage(const ImageEx& src) {
		if(src.format != SampleFormat::U8 && src.format != SampleFormat::U16) throw std::runtime_error("libcacaoimage::Image holds only 8- or 16-bit integer samples");
		Image out = {};
		out.w = src.w;
		out.h = src.h;
		out.layout = src.layout;
		out.bitsPerChannel = src.format == SampleFormat::U16 ? 16 : 8;
		out.data = src.data;
		out.format = Image::Format::PNG;
		out.quality = 100;
		out.lossy = false;
		return out;
	}

	ImageEx Convert(const ImageEx& src, Image::Layout layout, SampleFormat format, ConvertMode mode) {
		return Convert(src, layout, format, mode, false);
	}

	ImageEx Convert(const ImageEx& src, Image::Layout layout, SampleFormat format, ConvertMode mode, bool flipRows) {
		ImageEx out;
		out.w = src.w;
		out.h = src.h;
		out.layout = layout;
		out.format = format;
		const uint64_t pixels = static_cast<uint64_t>(src.w) * src.h;
		if(src.layout == layout && src.format == format) check(CACAO_E_SAME_LAYOUT);
		if(src.data.size() < pixels * BytesPerPixel(src.layout, src.format)) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
		out.data.resize(pixels * BytesPerPixel(layout, format));
		if(pixels && flipRows && src.h > 1)
			check(cacao_img_convert_flip(-1, src.data.data(), out.data.data(), src.w, src.h, 1, static_cast<int>(src.layout), static_cast<int>(layout), static_cast<int>(src.format), static_cast<int>(format), static_cast<int>(mode), CACAO_MEM_HOST, nullptr));
		else if(pixels)
			check(cacao_img_convert(-1, src.data.data(), out.data.data(), pixels, static_cast<int>(src.layout), static_cast<int>(layout), static_cast<int>(src.format), static_cast<int>(format), static_cast<int>(mode), CACAO_MEM_HOST, nullptr));
		return out;
	}

	std::vector<ImageEx> ConvertBatch(const std::vector<ImageEx>& srcs, Image::Layout layout, SampleFormat format, ConvertMode mode, const std::vector<int>& devices) {
		std::vector<ImageEx> out(srcs.size());
		for(std::size_t i = 0; i < srcs.size(); ++i) {
			if(srcs[i].layout == layout && srcs[i].format == format) check(CACAO_E_SAME_LAYOUT);
			if(srcs[i].data.size() < static_cast<std::size_t>(srcs[i].w) * srcs[i].h * BytesPerPixel(srcs[i].layout, srcs[i].format)) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
			out[i].w = srcs[i].w;
			out[i].h = srcs[i].h;
			out[i].layout = layout;
			out[i].format = format;
			out[i].data.resize(static_cast<std::size_t>(srcs[i].w) * srcs[i].h * BytesPerPixel(layout, format));
		}
		const std::vector<int> devs = devices.empty() ? std::vector<int>{-1} : devices;
		std::vector<std::string> errors(devs.size());
		auto worker = [&](std::size_t k) {
			// contiguous, balanced index range of device k
			const std::size_t base = srcs.size() / devs.size(), extra = srcs.size() % devs.size();
			const std::size_t lo = k * base + std::min(k, extra), hi = lo + base + (k < extra ? 1 : 0);
			for(std::size_t i = lo; i < hi; ++i) {
				const uint64_t pixels = static_cast<uint64_t>(srcs[i].w) * srcs[i].h;
				if(pixels == 0) continue;
				const int rc = cacao_img_convert(devs[k], srcs[i].data.data(), out[i].data.data(), pixels, static_cast<int>(srcs[i].layout), static_cast<int>(layout), static_cast<int>(srcs[i].format), static_cast<int>(format), static_cast<int>(mode), CACAO_MEM_HOST, nullptr);
				if(rc != CACAO_OK) {
					errors[k] = cacao_img_last_error();
					return;
				}
			}
		};
		if(devs.size() == 1) {
			worker(0);
		} else {
			std::vector<std::thread> pool;
			for(std::size_t k = 0; k < devs.size(); ++k) pool.emplace_back(worker, k);
			for(auto& t : pool) t.join();
		}
		for(const auto& e : errors)
			if(!e.empty()) throw std::runtime_error("libcacaoimage-b200: " + e);
		return out;
	}

	Image ToRGB8(const Image& src) {
		return ToRGB8(src, false);
	}

	Image ToRGB8(const Image& src, bool flipRows) {
		if(src.bitsPerChannel == 8 && src.layout == Image::Layout::RGB) return flipRows ? Flip(src) : src;
		Image out = {};
		out.w = src.w;
		out.h = src.h;
		out.layout = Image::Layout::RGB;
		out.bitsPerChannel = 8;
		out.format = src.format;
		out.quality = src.quality;
		out.lossy = src.lossy;
		const uint64_t pixels = static_cast<uint64_t>(src.w) * src.h;
		if(src.data.size() < pixels * static_cast<int>(src.layout) * (src.bitsPerChannel / 8)) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
		out.data.resize(pixels * 3);
		if(pixels && flipRows && src.h > 1)
			check(cacao_img_convert_flip(-1, src.data.data(), out.data.data(), src.w, src.h, 1, static_cast<int>(src.layout), 3, fmt_of_bits(src.bitsPerChannel), CACAO_FMT_U8, CACAO_MODE_REF_EXACT, CACAO_MEM_HOST, nullptr));
		else if(pixels)
			check(cacao_img_convert(-1, src.data.data(), out.data.data(), pixels, static_cast<int>(src.layout), 3, fmt_of_bits(src.bitsPerChannel), CACAO_FMT_U8, CACAO_MODE_REF_EXACT, CACAO_MEM_HOST, nullptr));
		return out;
	}

	std::size_t CubeRGB8::FaceBytes(unsigned int level) const {
		unsigned int lw = w, lh = h;
		for(unsigned int l = 0; l < level; ++l) {
			lw = lw / 2 ? lw / 2 : 1;
			lh = lh / 2 ? lh / 2 : 1;
		}
		return static_cast<std::size_t>(lw) * lh * 3;
	}

	CubeRGB8 EngineCube(const std::array<Image, 6>& faces, bool flipRows, bool mips, int device) {
		CubeRGB8 cube;
		cube.w = faces[0].w;
		cube.h = faces[0].h;
		bool uniform = true;
		for(const Image& f : faces) {
			if(f.w != cube.w || f.h != cube.h) throw std::runtime_error("Not all cubemap faces are the same size!"); // Cubemap.cpp:25
			if(!(f.bitsPerChannel == 8 || f.bitsPerChannel == 16)) check(CACAO_E_BAD_DEPTH);
			if(f.data.size() < static_cast<std::size_t>(f.w) * f.h * static_cast<int>(f.layout) * (f.bitsPerChannel / 8)) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
			uniform = uniform && f.layout == faces[0].layout && f.bitsPerChannel == faces[0].bitsPerChannel;
		}
		if(cube.w == 0 || cube.h == 0) check(CACAO_E_ZERO_DIM);
		const uint32_t flags = (flipRows ? CACAO_CUBE_FLIP_ROWS : 0u) | (mips ? CACAO_CUBE_MIPS : 0u);
		cube.data.resize(cacao_img_engine_cube_bytes(cube.w, cube.h, flags));
		std::size_t off = 0;
		for(unsigned int lw = cube.w, lh = cube.h;; lw = lw / 2 ? lw / 2 : 1, lh = lh / 2 ? lh / 2 : 1) {
			cube.levelOffset.push_back(off);
			off += 6 * static_cast<std::size_t>(lw) * lh * 3;
			if(!mips || (lw == 1 && lh == 1)) break;
		}
		const void* ptrs[6];
		std::array<Image, 6> converted; // only for mixed sets: each face brought to RGB8 on its own first
		if(uniform) {
			for(int f = 0; f < 6; ++f) ptrs[f] = faces[f].data.data();
			check(cacao_img_engine_cube(device, ptrs, cube.w, cube.h, static_cast<int>(faces[0].layout), fmt_of_bits(faces[0].bitsPerChannel), cube.data.data(), flags, CACAO_MEM_HOST, nullptr));
		} else {
			for(int f = 0; f < 6; ++f) {
				converted[f] = ToRGB8(faces[f]);
				ptrs[f] = converted[f].data.data();
			}
			check(cacao_img_engine_cube(device, ptrs, cube.w, cube.h, 3, CACAO_FMT_U8, cube.data.data(), flags, CACAO_MEM_HOST, nullptr));
		}
		return cube;
	}

	TextureU8 EngineTexture(const Image& image, bool flipRows, bool mips, int device) {
		if(!(image.w > 0 && image.h > 0)) check(CACAO_E_ZERO_DIM);
		if(!(image.bitsPerChannel == 8 || image.bitsPerChannel == 16)) check(CACAO_E_BAD_DEPTH);
		const int ch = static_cast<int>(image.layout);
		if(image.data.size() < static_cast<std::size_t>(image.w) * image.h * ch * (image.bitsPerChannel / 8)) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
		TextureU8 tex;
		tex.w = image.w;
		tex.h = image.h;
		tex.layout = image.layout;
		const uint32_t flags = (flipRows ? CACAO_CUBE_FLIP_ROWS : 0u) | (mips ? CACAO_CUBE_MIPS : 0u);
		tex.data.resize(cacao_img_engine_texture_bytes(image.w, image.h, ch, flags));
		std::size_t off = 0;
		for(unsigned int lw = image.w, lh = image.h;; lw = lw / 2 ? lw / 2 : 1, lh = lh / 2 ? lh / 2 : 1) {
			tex.levelOffset.push_back(off);
			off += static_cast<std::size_t>(lw) * lh * ch;
			if(!mips || (lw == 1 && lh == 1)) break;
		}
		check(cacao_img_engine_texture(device, image.data.data(), image.w, image.h, ch, fmt_of_bits(image.bitsPerChannel), tex.data.data(), flags, CACAO_MEM_HOST, nullptr));
		return tex;
	}

	ImageEx Flip(const ImageEx& src) {
		if(!(src.w > 0 && src.h > 0)) check(CACAO_E_ZERO_DIM);
		if(src.data.empty()) check(CACAO_E_EMPTY);
		ImageEx out = src;
		if(src.h == 1) return out;
		check(cacao_img_flip_rows(-1, src.data.data(), out.data.data(), src.w, src.h, static_cast<uint32_t>(BytesPerPixel(src.layout, src.format)), CACAO_MEM_HOST, nullptr));
		return out;
	}

	bool FusedProjectionDefined(Image::Layout srcLayout, SampleFormat srcFormat, Image::Layout layout, SampleFormat format) {
		if(srcLayout == layout && srcFormat == format) return true;
		const bool float_src = srcFormat == SampleFormat::F16 || srcFormat == SampleFormat::F32;
		return float_src && (layout == srcLayout || (srcLayout == Image::Layout::RGB && layout == Image::Layout::RGBA) || (srcLayout == Image::Layout::RGBA && layout == Image::Layout::RGB));
	}

	std::array<ImageEx, 6> EquirectToCubemap(const ImageEx& src, unsigned int faceSize, const std::vector<int>& devices) {
		return EquirectToCubemap(src, faceSize, src.layout, src.format, devices, false);
	}

	std::array<ImageEx, 6> EquirectToCubemap(const ImageEx& src, unsigned int faceSize, Image::Layout layout, SampleFormat format, const std::vector<int>& devices, bool flipRows) {
		if(!FusedProjectionDefined(src.layout, src.format, layout, format)) {
			// no one-pass form (integer source, or a layout change other than RGB <-> RGBA): two passes
			// (the row mirror is folded into the first one: conversion does not care about rows)
			// ConvertMode::Fixed: this is a new operation, not a replacement of a reference call -- the
			// reference's Gray -> RGBA pixel-0 repeat and its clamp of 16-bit gray to 255 are bugs to reproduce
			// where its own functions are replaced, not to spread to faces that never went through them
			std::array<ImageEx, 6> faces = EquirectToCubemap(src, faceSize, src.layout, src.format, devices, flipRows);
			for(auto& f : faces) f = Convert(f, layout, format, ConvertMode::Fixed);
			return faces;
		}
		if(src.w == 0 || src.h == 0 || faceSize == 0) check(CACAO_E_ZERO_DIM);
		const std::size_t bpp = BytesPerPixel(src.layout, src.format);
		if(src.data.size() < static_cast<std::size_t>(src.w) * src.h * bpp) throw std::runtime_error("equirect source buffer is smaller than w*h*bytes-per-pixel");
		const std::size_t face_bytes = static_cast<std::size_t>(faceSize) * faceSize * BytesPerPixel(layout, format);
		// the six results are allocated up front and every device reads its rows back straight into them
		std::array<ImageEx, 6> faces;
		for(unsigned f = 0; f < 6; ++f) {
			faces[f].w = faces[f].h = faceSize;
			faces[f].layout = layout;
			faces[f].format = format;
			faces[f].data.resize(face_bytes);
		}

		// One device: the whole cube as six units; the library pipelines upload, resampling and
		// readback by itself.  Several devices: (face, row band) units ordered by the latitude they
		// read and dealt in contiguous runs of equal output size, so every device uploads only a band
		// of the source (about 1/4 of it with 8 devices) instead of nearly all of it -- the host side of
		// this call is PCIe- and memcpy-bound, the kernels are noise.  Each device thread then makes ONE
		// pipelined call for its run and the rows land straight in the six result vectors.
		const std::vector<int> devs = devices.empty() ? std::vector<int>{-1} : devices;
		std::vector<cacao_cube_unit> units;
		if(devs.size() == 1) {
			for(unsigned f = 0; f < 6; ++f) units.push_back({f, 0, faceSize});
		} else {
			const unsigned bands = static_cast<unsigned>(std::min<std::size_t>(faceSize, 2 * devs.size()));
			struct Keyed {
				cacao_cube_unit u;
				uint64_t mid; // twice the centre of its source-row window
			};
			std::vector<Keyed> keyed;
			for(unsigned f = 0; f < 6; ++f)
				for(unsigned b = 0; b < bands; ++b) {
					const unsigned r0 = static_cast<unsigned>(static_cast<uint64_t>(faceSize) * b / bands);
					const unsigned r1 = static_cast<unsigned>(static_cast<uint64_t>(faceSize) * (b + 1) / bands);
					if(r1 <= r0) continue;
					uint32_t w0 = 0, wn = 0;
					check(cacao_img_equirect_src_window(src.w, src.h, faceSize, 1u << f, r0, r1, &w0, &wn));
					keyed.push_back({{f, r0, r1}, 2ull * w0 + wn});
				}
			std::stable_sort(keyed.begin(), keyed.end(), [](const Keyed& x, const Keyed& y) { return x.mid < y.mid; });
			for(const Keyed& k : keyed) units.push_back(k.u);
		}
		// contiguous runs with (nearly) equal numbers of output rows
		uint64_t total_rows = 0;
		for(const auto& un : units) total_rows += un.row_end - un.row_begin;
		std::vector<std::size_t> run_begin(devs.size() + 1, units.size());
		{
			uint64_t acc = 0;
			std::size_t d = 0;
			run_begin[0] = 0;
			for(std::size_t u = 0; u < units.size(); ++u) {
				while(d + 1 < devs.size() && acc * devs.size() >= (d + 1) * total_rows) run_begin[++d] = u;
				acc += units[u].row_end - units[u].row_begin;
			}
			while(d + 1 < devs.size()) run_begin[++d] = units.size();
		}
		std::vector<std::string> errors(devs.size());
		auto worker = [&](std::size_t k) {
			const std::size_t u0 = run_begin[k], u1 = run_begin[k + 1];
			if(u1 <= u0) return;
			uint32_t lo = UINT32_MAX, hi = 0;
			for(std::size_t u = u0; u < u1; ++u) {
				uint32_t w0 = 0, wn = 0;
				if(cacao_img_equirect_src_window(src.w, src.h, faceSize, 1u << units[u].face, units[u].row_begin, units[u].row_end, &w0, &wn) != CACAO_OK) {
					errors[k] = cacao_img_last_error();
					return;
				}
				lo = std::min(lo, w0);
				hi = std::max(hi, w0 + wn);
			}
			const std::size_t pitch = static_cast<std::size_t>(src.w) * bpp;
			void* dst[6];
			for(unsigned f = 0; f < 6; ++f) dst[f] = faces[f].data.data();
			if(cacao_img_equirect_to_cube_units_host_cvt(devs[k], src.data.data() + static_cast<std::size_t>(lo) * pitch, src.w, src.h, lo, hi - lo, static_cast<int>(src.layout), static_cast<int>(src.format), dst, faceSize, static_cast<int>(layout), static_cast<int>(format), flipRows ? CACAO_EQUIRECT_FLIP_ROWS : 0u, units.data() + u0, static_cast<uint32_t>(u1 - u0)) != CACAO_OK)
				errors[k] = cacao_img_last_error();
		};
		if(devs.size() == 1) {
			worker(0);
		} else {
			std::vector<std::thread> pool;
			for(std::size_t k = 0; k < devs.size(); ++k) pool.emplace_back(worker, k);
			for(auto& t : pool) t.join();
		}
		for(const auto& e : errors)
			if(!e.empty()) throw std::runtime_error("libcacaoimage-b200: " + e);
		return faces;
	}
	ImageEx Downsample2x(const ImageEx& src) {
		if(src.w == 0 || src.h == 0) check(CACAO_E_ZERO_DIM);
		if(src.data.size() < static_cast<std::size_t>(src.w) * src.h * BytesPerPixel(src.layout, src.format)) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
		ImageEx out;
		out.w = src.w / 2 ? src.w / 2 : 1;
		out.h = src.h / 2 ? src.h / 2 : 1;
		out.layout = src.layout;
		out.format = src.format;
		out.data.resize(static_cast<std::size_t>(out.w) * out.h * BytesPerPixel(src.layout, src.format));
		check(cacao_img_downsample2x(-1, src.data.data(), out.data.data(), src.w, src.h, static_cast<int>(src.layout), static_cast<int>(src.format), CACAO_MEM_HOST, nullptr));
		return out;
	}

	std::vector<ImageEx> GenerateMipChain(const ImageEx& base) {
		std::vector<ImageEx> levels;
		DeviceImage cur = DeviceImage::Upload(base);
		while(cur.Width() > 1 || cur.Height() > 1) {
			cur = cur.Downsample2x();
			levels.push_back(cur.Download());
		}
		return levels;
	}

	Image UnpackTexels(const unsigned char* texels, unsigned int w, unsigned int h, const std::string& hint) {
		Image out = {};
		out.w = w;
		out.h = h;
		out.bitsPerChannel = 8; // Model.cpp:248
		out.format = Image::Format::PNG;
		out.quality = 0;
		out.lossy = false;
		const uint64_t n = static_cast<uint64_t>(w) * h;
		std::vector<unsigned char> tmp(n ? n * 4 : 4);
		int ch = 0;
		const int rc = cacao_img_unpack_texels(-1, n ? texels : tmp.data(), tmp.data(), n, hint.c_str(), &ch, CACAO_MEM_HOST, nullptr);
		if(rc == CACAO_E_BAD_ARG) throw std::runtime_error(cacao_img_last_error()); // the reference's texts for rejected hints
		check(rc);
		out.layout = static_cast<Image::Layout>(ch);
		tmp.resize(n * static_cast<uint64_t>(ch));
		out.data = std::move(tmp);
		return out;
	}

	// ---- DeviceImage ---------------------------------------------------------------------------

	namespace {
		std::shared_ptr<void> device_alloc(int device, std::size_t bytes) {
			void* p = nullptr;
			check(cacao_img_malloc(device, &p, bytes));
			return std::shared_ptr<void>(p, [device](void* q) { cacao_img_free(device, q); });
		}
	}

	DeviceImage::DeviceImage(unsigned int w, unsigned int h, Image::Layout layout, SampleFormat format, int device) : w_(w), h_(h), layout_(layout), format_(format), device_(device) {
		if(w == 0 || h == 0) check(CACAO_E_ZERO_DIM);
		owner_ = device_alloc(device, Bytes());
		ptr_ = owner_.get();
	}

	DeviceImage DeviceImage::Upload(const ImageEx& src, int device) {
		DeviceImage d(src.w, src.h, src.layout, src.format, device);
		if(src.data.size() < d.Bytes()) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
		check(cacao_img_upload(device, d.ptr_, src.data.data(), d.Bytes(), nullptr));
		check(cacao_img_sync(device, nullptr)); // src may go away as soon as this returns
		return d;
	}

	DeviceImage DeviceImage::Upload(const Image& src, int device) {
		if(!(src.bitsPerChannel == 8 || src.bitsPerChannel == 16)) check(CACAO_E_BAD_DEPTH);
		DeviceImage d(src.w, src.h, src.layout, src.bitsPerChannel == 16 ? SampleFormat::U16 : SampleFormat::U8, device);
		if(src.data.size() < d.Bytes()) throw std::runtime_error("Image data buffer is smaller than w * h * channels * bytes per channel!");
		check(cacao_img_upload(device, d.ptr_, src.data.data(), d.Bytes(), nullptr));
		check(cacao_img_sync(device, nullptr));
		return d;
	}

	ImageEx DeviceImage::Download() const {
		if(Empty()) check(CACAO_E_EMPTY);
		ImageEx out;
		out.w = w_;
		out.h = h_;
		out.layout = layout_;
		out.format = format_;
		out.data.resize(Bytes());
		check(cacao_img_download(device_, out.data.data(), ptr_, Bytes(), nullptr));
		check(cacao_img_sync(device_, nullptr));
		return out;
	}

	DeviceImage DeviceImage::Convert(Image::Layout layout, SampleFormat format, ConvertMode mode, bool flipRows) const {
		if(Empty()) check(CACAO_E_EMPTY);
		if(layout == layout_ && format == format_) check(CACAO_E_SAME_LAYOUT);
		DeviceImage out(w_, h_, layout, format, device_);
		if(flipRows && h_ > 1)
			check(cacao_img_convert_flip(device_, ptr_, out.ptr_, w_, h_, 1, static_cast<int>(layout_), static_cast<int>(layout), static_cast<int>(format_), static_cast<int>(format), static_cast<int>(mode), CACAO_MEM_DEVICE, nullptr));
		else
			check(cacao_img_convert(device_, ptr_, out.ptr_, static_cast<uint64_t>(w_) * h_, static_cast<int>(layout_), static_cast<int>(layout), static_cast<int>(format_), static_cast<int>(format), static_cast<int>(mode), CACAO_MEM_DEVICE, nullptr));
		return out;
	}

	DeviceImage DeviceImage::ToRGB8(bool flipRows) const {
		if(layout_ == Image::Layout::RGB && format_ == SampleFormat::U8) return flipRows ? Flip() : *this;
		return Convert(Image::Layout::RGB, SampleFormat::U8, ConvertMode::ReferenceExact, flipRows);
	}

	DeviceImage DeviceImage::Flip() const {
		if(Empty()) check(CACAO_E_EMPTY);
		if(h_ == 1) return *this;
		DeviceImage out(w_, h_, layout_, format_, device_);
		check(cacao_img_flip_rows(device_, ptr_, out.ptr_, w_, h_, static_cast<uint32_t>(BytesPerPixel(layout_, format_)), CACAO_MEM_DEVICE, nullptr));
		return out;
	}

	DeviceImage DeviceImage::Downsample2x() const {
		if(Empty()) check(CACAO_E_EMPTY);
		DeviceImage out(w_ / 2 ? w_ / 2 : 1, h_ / 2 ? h_ / 2 : 1, layout_, format_, device_);
		check(cacao_img_downsample2x(device_, ptr_, out.ptr_, w_, h_, static_cast<int>(layout_), static_cast<int>(format_), CACAO_MEM_DEVICE, nullptr));
		return out;
	}

	std::array<DeviceImage, 6> DeviceImage::EquirectToCubemap(unsigned int faceSize) const {
		return EquirectToCubemap(faceSize, layout_, format_, false);
	}

	std::array<DeviceImage, 6> DeviceImage::EquirectToCubemap(unsigned int faceSize, Image::Layout layout, SampleFormat format, bool flipRows) const {
		if(Empty()) check(CACAO_E_EMPTY);
		if(faceSize == 0) check(CACAO_E_ZERO_DIM);
		if(!FusedProjectionDefined(layout_, format_, layout, format)) {
			std::array<DeviceImage, 6> faces = EquirectToCubemap(faceSize, layout_, format_, flipRows);
			for(auto& f : faces) f = f.Convert(layout, format, ConvertMode::Fixed); // see the host form above
			return faces;
		}
		const std::size_t face_bytes = static_cast<std::size_t>(faceSize) * faceSize * BytesPerPixel(layout, format);
		std::shared_ptr<void> cube = device_alloc(device_, 6 * face_bytes);
		check(cacao_img_equirect_to_cube_units_cvt(device_, ptr_, w_, h_, 0, h_, static_cast<int>(layout_), static_cast<int>(format_), cube.get(), faceSize, static_cast<int>(layout), static_cast<int>(format), flipRows ? CACAO_EQUIRECT_FLIP_ROWS : 0u, nullptr, 0, nullptr));
		std::array<DeviceImage, 6> faces;
		for(unsigned f = 0; f < 6; ++f) {
			faces[f].owner_ = cube;
			faces[f].ptr_ = static_cast<unsigned char*>(cube.get()) + f * face_bytes;
			faces[f].w_ = faces[f].h_ = faceSize;
			faces[f].layout_ = layout;
			faces[f].format_ = format;
			faces[f].device_ = device_;
		}
		return faces;
	}
}
