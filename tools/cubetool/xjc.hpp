// xjc.hpp -- the packed (.xjc) and unpacked (.ajc) cubemap files of Cacao Engine, host side.
//
// Restates, for cubemaps only, what libcacaoformats does:
//   * PackedContainer (libs/formats/src/PackedContainer.cpp:41-158): bytes CA CA 00, a format code
//     (0xC4 = cubemap), u16 version, u64 uncompressed size, then ONE bzip2 stream (level 9, work
//     factor 30) holding the payload;
//   * the cubemap payload (libs/formats/src/PackedEncode.cpp:18-48, PackedDecode.cpp:15-62): six u64
//     blob sizes in the order +X,-X,+Y,-Y,+Z,-Z, then the six blobs -- lossless WebP on the way in,
//     anything decode::DecodeGeneric understands on the way out;
//   * the unpacked definition (libs/formats/src/UnpackedDecode.cpp:17-67): a YAML map with the scalar
//     keys right,left,top,bottom,front,back naming the face files relative to the .ajc.
// bzip2 itself is the system's libbz2 (the reference links the same library as a meson wrap,
// subprojects/bzip2.wrap), bound at run time so the tool still starts without it.
// Sequential host code, outside the GPU hot path.
#pragma once

#include <array>
#include <cstdint>
#include <filesystem>
#include <string>
#include <vector>

#include "libcacaoimage.hpp"

namespace xjc {
	constexpr uint8_t kCubemapCode = 0xC4;

	struct Container {
		uint8_t code = 0;
		uint16_t version = 0;
		std::vector<unsigned char> payload;
	};

	// Throw std::runtime_error with the reference's messages where it has one.
	Container ReadContainer(const std::vector<unsigned char>& file);
	std::vector<unsigned char> WriteContainer(const Container& c);

	// payload <-> six encoded face blobs
	std::array<std::vector<unsigned char>, 6> SplitCubemapPayload(const Container& c);
	Container MakeCubemapContainer(const std::array<std::vector<unsigned char>, 6>& blobs);

	// The six face paths of an .ajc, in slot order, as written in the file (relative to it).
	std::array<std::string, 6> ReadAjc(const std::string& text);

	// Face blob -> Image, by magic bytes like decode::DecodeGeneric (libs/image/src/Forwarding.cpp:8-67).
	// This build decodes PNG, lossless WebP and binary PNM; other known types raise a clear error.
	libcacaoimage::Image DecodeFace(const std::vector<unsigned char>& blob);

	// encode::EncodeWebP with lossy = false (libs/image/src/WebP.cpp:73-96), same validation messages.
	std::vector<unsigned char> EncodeFaceWebP(const libcacaoimage::Image& img);
}
