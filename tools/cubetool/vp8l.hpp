// vp8l.hpp -- WebP lossless (VP8L) bitstream, the face encoding inside a packed cubemap (.xjc).
//
// The reference stores every face of a packed cubemap as a lossless WebP blob made by libwebp
// (libs/formats/src/PackedEncode.cpp:39-44 -> libs/image/src/WebP.cpp:88-96,
// WebPEncodeLosslessRGB / WebPEncodeLosslessRGBA) and reads them back with WebPDecode*
// (libs/image/src/WebP.cpp:10-72).  libwebp is an un-vendored meson wrap (subprojects/libwebp.wrap,
// v1.5.0) and absent from this image, so this file implements the published container + VP8L
// bitstream format directly: a complete decoder (all four transforms, colour cache, meta prefix
// codes, LZ77 with the 2-D distance map) so that files written by the reference tool can be read,
// and an encoder that emits valid, reasonably compact streams (predictor + subtract-green
// transforms, length-limited canonical prefix codes) that libwebp reads.  Host-side, sequential
// entropy coding: NOT part of the GPU hot path.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace vp8l {
	// True when `data` is a RIFF/WEBP file (any flavour).
	bool IsWebP(const uint8_t* data, std::size_t size);

	// Decodes a lossless WebP file into interleaved 8-bit samples, top row first: RGBA when the
	// stream's alpha flag is set, RGB otherwise (what WebPDecodeRGBA / WebPDecodeRGB hand to
	// libs/image/src/WebP.cpp:38-66).  Throws std::runtime_error for lossy (VP8) or malformed files.
	struct Decoded {
		unsigned w = 0, h = 0;
		int channels = 0; // 3 or 4
		std::vector<uint8_t> pixels;
	};
	Decoded Decode(const uint8_t* data, std::size_t size);

	// Encodes interleaved RGB (channels == 3) or RGBA (channels == 4) 8-bit pixels as a complete
	// RIFF/WEBP/VP8L file.  1 <= w, h <= 16384.
	std::vector<uint8_t> Encode(const uint8_t* pixels, unsigned w, unsigned h, int channels);
}
