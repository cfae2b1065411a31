// libcacaoimage.hpp -- source-compatible stand-in for the reference's public image header
// (libs/image/include/libcacaoimage.hpp).  Same namespace, same `Image` members in the same order
// (libstdc++ x86-64: sizeof 56; w@0 h@4 layout@8 bitsPerChannel@9 data@16 format@40 quality@44
// lossy@48), same free functions, same exceptions with the same texts.  The three pixel
// transforms run on a B200 through the C ABI in cacao_img.h; the codec entry points are declared
// for source compatibility and are not part of this library (they wrap third-party entropy
// codecs and are out of scope).
#pragma once

#include <cstddef>
#include <cstdint>
#include <istream>
#include <ostream>
#include <vector>

namespace libcacaoimage {

	// A decoded image: interleaved channels, rows packed top first, 16-bit samples little-endian.
	struct Image {
		unsigned int w;
		unsigned int h;
		enum class Layout : uint8_t { Grayscale = 1, RGB = 3, RGBA = 4 } layout; // channel count
		uint8_t bitsPerChannel;                                                  // 8 or 16
		std::vector<unsigned char> data;
		enum class Format { PNG, JPEG, WebP, TGA, TIFF } format; // container the pixels came from
		int quality;                                             // 0..100, encoder hint
		bool lossy;                                              // encoder hint
	};

	// Codec surface of the reference (libcacaoimage.hpp:37-176): declared, not provided here.
	namespace decode {
		Image DecodeGeneric(std::istream& input);
		Image DecodePNG(std::istream& input);
		Image DecodeJPEG(std::istream& input);
		Image DecodeWebP(std::istream& input);
		Image DecodeTGA(std::istream& input);
		Image DecodeTIFF(std::istream& input);
	}
	namespace encode {
		std::size_t Reencode(const Image& src, std::ostream& out);
		std::size_t EncodePNG(const Image& src, std::ostream& out);
		std::size_t EncodeJPEG(const Image& src, std::ostream& out);
		std::size_t EncodeWebP(const Image& src, std::ostream& out);
		std::size_t EncodeTGA(const Image& src, std::ostream& out);
		std::size_t EncodeTIFF(const Image& src, std::ostream& out);
	}

	// 16-bit -> 8-bit through the engine's sRGB table (reference: libcacaoimage.hpp:187,
	// src/Colordepth.cpp:7-41).  Throws std::runtime_error unless src.bitsPerChannel == 16.
	Image Convert16To8BitColor(const Image& src);

	// Vertical mirror for bottom-row-first APIs (reference: libcacaoimage.hpp:196).  Implements what
	// src/Flip.cpp:6-42 intends -- out.row[y] = src.row[h-1-y], src left untouched -- rather than its
	// out-of-bounds behaviour; validation and messages follow Flip.cpp:8-10.
	Image Flip(const Image& src);

	// Gray / RGB / RGBA re-layout, bit-exact with the reference including its quirks (reference:
	// libcacaoimage.hpp:213, src/Layout.cpp:8-103).  Throws std::runtime_error when `layout` is the
	// image's current layout.
	Image ChangeChannelLayout(const Image& src, Image::Layout layout);
}
