// libcacaoimage_ext.hpp -- additions beside the reference API: float sample formats, fused and
// batched conversion, device-resident images and the equirect -> cubemap projection.  None of this
// exists in the reference (libcacaoimage.hpp:20 allows only 8/16-bit integers; cubetool has no
// projection, tools/cubetool/main.cpp:31-33); semantics are DESIGN.md "Spec B".
#pragma once

#include <array>
#include <cstdint>
#include <vector>

#include "cacao_img.h"
#include "libcacaoimage.hpp"

namespace libcacaoimage {

	enum class SampleFormat : uint8_t { U8 = CACAO_FMT_U8, U16 = CACAO_FMT_U16, F16 = CACAO_FMT_F16, F32 = CACAO_FMT_F32 };
	enum class ConvertMode : uint8_t { ReferenceExact = CACAO_MODE_REF_EXACT, Fixed = CACAO_MODE_FIXED };

	// Host image with an explicit sample format (Image carries only 8/16-bit integers).
	struct ImageEx {
		unsigned int w = 0, h = 0;
		Image::Layout layout = Image::Layout::RGBA;
		SampleFormat format = SampleFormat::U8;
		std::vector<unsigned char> data;
	};

	ImageEx FromImage(const Image& src); // U8 / U16 according to bitsPerChannel
	Image ToImage(const ImageEx& src);   // throws unless U8 / U16

	std::size_t BytesPerPixel(Image::Layout layout, SampleFormat format);

	// Fused layout + format conversion in one pass over the pixels (one upload, one kernel, one readback).
	ImageEx Convert(const ImageEx& src, Image::Layout layout, SampleFormat format, ConvertMode mode = ConvertMode::ReferenceExact);

	// The same conversion for a batch of images, sharded by contiguous image-index range over `devices`
	// (CUDA ordinals; empty = current device), one host thread per device, no cross-device traffic.
	// Images may differ in size and source layout/format.
	std::vector<ImageEx> ConvertBatch(const std::vector<ImageEx>& srcs, Image::Layout layout, SampleFormat format, ConvertMode mode = ConvertMode::ReferenceExact, const std::vector<int>& devices = {});

	// What engine/src/core/Cubemap.cpp:28-31 does per face in two passes and two allocations
	// (Convert16To8BitColor, then ChangeChannelLayout to RGB), as one kernel.
	Image ToRGB8(const Image& src);

	// Equirectangular panorama -> six N x N faces in the order +X,-X,+Y,-Y,+Z,-Z
	// (right,left,up,down,front,back), same layout and sample format as the source.
	// devices: CUDA ordinals to spread face/row-band work units over (empty = current device).
	std::array<ImageEx, 6> EquirectToCubemap(const ImageEx& src, unsigned int faceSize, const std::vector<int>& devices = {});

	// Flip for any sample format.
	ImageEx Flip(const ImageEx& src);
}
