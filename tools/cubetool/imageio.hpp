// imageio.hpp -- minimal uncompressed / zlib-only image file I/O for ce-cubetool's `project`.
//
// The reference decodes through libpng / libjpeg-turbo / libwebp / libtiff / aseprite-tga
// (libs/image/src/*.cpp); those are un-vendored meson wraps and absent here, and entropy codecs are
// outside the GPU hot path anyway.  This file covers what the projection subcommand needs with
// nothing but zlib: PNG (8/16-bit gray, RGB, RGBA, non-interlaced), PFM, NPY and binary PNM, plus
// lossless WebP through this tool's own VP8L codec (vp8l.hpp).
#pragma once

#include <filesystem>
#include <string>
#include <vector>

#include "libcacaoimage_ext.hpp"

namespace cubeio {
	// Reads by magic bytes, not by extension (as decode::DecodeGeneric does, Forwarding.cpp:8-67).
	// Throws std::runtime_error with a human-readable reason.
	libcacaoimage::ImageEx ReadImage(const std::filesystem::path& path);
	libcacaoimage::ImageEx DecodeImage(const std::vector<unsigned char>& file);

	// format: "png" (U8/U16 only), "webp" (lossless, U8 RGB/RGBA only), "pfm" (F32 gray/RGB only), "npy" (anything)
	void WriteImage(const libcacaoimage::ImageEx& img, const std::filesystem::path& path, const std::string& format);

	// "png" for integer samples, "npy" for float samples
	std::string DefaultFormat(const libcacaoimage::ImageEx& img);
}
