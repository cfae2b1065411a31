// libcacaoimage_ext.hpp -- additions beside the reference API: float sample formats, fused and
// batched conversion, device-resident images and the equirect -> cubemap projection.  None of this
// exists in the reference (libcacaoimage.hpp:20 allows only 8/16-bit integers; cubetool has no
// projection, tools/cubetool/main.cpp:31-33); semantics are DESIGN.md "Spec B".
#pragma once

#include <array>
#include <cstdint>
#include <memory>
#include <string>
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
	// flipRows: the result comes out bottom row first -- Convert and Flip (the GL backend's step before
	// upload, engine/src/opengl/impls/OpenGLTex2D.cpp:16) as one pass over the pixels.
	ImageEx Convert(const ImageEx& src, Image::Layout layout, SampleFormat format, ConvertMode mode, bool flipRows);

	// The same conversion for a batch of images, sharded by contiguous image-index range over `devices`
	// (CUDA ordinals; empty = current device), one host thread per device, no cross-device traffic.
	// Images may differ in size and source layout/format.
	std::vector<ImageEx> ConvertBatch(const std::vector<ImageEx>& srcs, Image::Layout layout, SampleFormat format, ConvertMode mode = ConvertMode::ReferenceExact, const std::vector<int>& devices = {});

	// What engine/src/core/Cubemap.cpp:28-31 does per face in two passes and two allocations
	// (Convert16To8BitColor, then ChangeChannelLayout to RGB), as one kernel.
	Image ToRGB8(const Image& src);
	// ... and OpenGLCubemap.cpp:25's Flip in the same kernel: three reference passes, one here
	Image ToRGB8(const Image& src, bool flipRows);

	// The engine's whole cubemap load chain in one call: what Cacao::Cubemap::Cubemap does per face
	// (Convert16To8BitColor when 16-bit, ChangeChannelLayout to RGB when not RGB, engine/src/core/Cubemap.cpp:20-32),
	// the GL backend's Flip before upload (engine/src/opengl/impls/OpenGLCubemap.cpp:25, flipRows) and the Vulkan
	// backend's copy of the six faces into one contiguous buffer (engine/src/vulkan/impls/VulkanCubemap.cpp:30-41),
	// optionally with the mip chain behind it (2x2 box per level; the GL path leaves that to glGenerateMipmap).
	// One upload / kernel / readback pipeline instead of up to eighteen passes over the pixels; values are the
	// reference's bit for bit.  Throws the engine's message when the faces differ in size.
	struct CubeRGB8 {
		unsigned int w = 0, h = 0;            // of a level-0 face
		std::vector<unsigned char> data;      // level 0: six RGB8 faces back to back (+X,-X,+Y,-Y,+Z,-Z); then the mip levels, each face-major
		std::vector<std::size_t> levelOffset; // byte offset of every level in data (one entry without mips)
		std::size_t FaceBytes(unsigned int level = 0) const;
		const unsigned char* Face(unsigned int face, unsigned int level = 0) const { return data.data() + levelOffset[level] + face * FaceBytes(level); }
	};
	CubeRGB8 EngineCube(const std::array<Image, 6>& faces, bool flipRows = false, bool mips = false, int device = -1);

	// The same for a 2-D texture: Cacao::Tex2D::Tex2D's Convert16To8BitColor (engine/src/core/Tex2D.cpp:18, layout
	// kept), the GL backend's Flip (engine/src/opengl/impls/OpenGLTex2D.cpp:16) and, on request, the mip levels it
	// asks the driver for (:51) -- one pipelined pass.  data = level 0, then every mip level down to 1 x 1.
	struct TextureU8 {
		unsigned int w = 0, h = 0;
		Image::Layout layout = Image::Layout::RGBA;
		std::vector<unsigned char> data;
		std::vector<std::size_t> levelOffset; // byte offset of every level in data (one entry without mips)
	};
	TextureU8 EngineTexture(const Image& image, bool flipRows = false, bool mips = false, int device = -1);

	// Equirectangular panorama -> six N x N faces in the order +X,-X,+Y,-Y,+Z,-Z
	// (right,left,up,down,front,back), same layout and sample format as the source.
	// devices: CUDA ordinals to spread face/row-band work units over (empty = current device).
	std::array<ImageEx, 6> EquirectToCubemap(const ImageEx& src, unsigned int faceSize, const std::vector<int>& devices = {});
	// The same with the faces delivered in another layout / sample format.  Float panoramas (F16, F32)
	// going to any sample format, with the same channels or RGB <-> RGBA, are resampled AND converted in
	// one pass over the pixels (DESIGN.md "Spec B.4": the fp32 filter result is converted as Convert()
	// converts an F32 sample) -- an RGB32F .hdr panorama becomes an RGBA16F cube without the F32 cube ever
	// existing.  FusedProjectionDefined() tells; every other pair runs the two passes for you.
	// flipRows: every face comes out bottom row first -- the vertical mirror the GL backend applies before
	// upload (engine/src/opengl/impls/OpenGLCubemap.cpp:25) -- folded into the store addressing (any format).
	std::array<ImageEx, 6> EquirectToCubemap(const ImageEx& src, unsigned int faceSize, Image::Layout layout, SampleFormat format, const std::vector<int>& devices = {}, bool flipRows = false);
	bool FusedProjectionDefined(Image::Layout srcLayout, SampleFormat srcFormat, Image::Layout layout, SampleFormat format);

	// Flip for any sample format.
	ImageEx Flip(const ImageEx& src);

	// Mip levels (2x2 box average per level, DESIGN.md "Spec B.3"; the engine leaves this to the graphics
	// API, engine/src/opengl/impls/OpenGLTex2D.cpp:51).  Downsample2x gives the next level
	// (max(1,w/2) x max(1,h/2)); GenerateMipChain gives every level below `base` down to 1x1 with one
	// upload, the whole chain on the GPU, and one readback per level.
	ImageEx Downsample2x(const ImageEx& src);
	std::vector<ImageEx> GenerateMipChain(const ImageEx& base);

	// The uncompressed-texture branch of Cacao::Model::GetTexture (engine/src/core/Model.cpp:245-316):
	// w*h four-byte texels laid out as assimp's format hint says ("rgba8888", "argb8888", "bgra5650",
	// ...) become an 8-bit RGBA / RGB / Grayscale Image -- channels in r,g,b,a order, zero-count
	// channels dropped, counts below 8 widened by bit replication.  Throws the reference's messages
	// for the hints it rejects (Model.cpp:259-269).
	Image UnpackTexels(const unsigned char* texels, unsigned int w, unsigned int h, const std::string& hint);

	// An image that lives in the memory of one GPU.  The host API above pays one PCIe round trip per
	// call; a chain such as "project, then convert, then flip" (what `ce-cubetool project --depth 8
	// --layout rgb --flip` does, or the engine's 16 -> 8 bit, -> RGB, Flip sequence of
	// engine/src/core/Cubemap.cpp:28-31 + engine/src/opengl/impls/OpenGLCubemap.cpp:25) uploads once,
	// runs its kernels back to back on the device and downloads once.  Copies share the pixels
	// (reference-counted device allocation); every operation returns a new image and leaves its input
	// untouched, like the host functions.  Work is enqueued in order on the device's default stream;
	// Download() returns when the pixels are in host memory.  Same exceptions as the host API.
	class DeviceImage {
	  public:
		DeviceImage() = default;
		// uninitialised pixels
		DeviceImage(unsigned int w, unsigned int h, Image::Layout layout, SampleFormat format, int device = -1);

		static DeviceImage Upload(const ImageEx& src, int device = -1);
		static DeviceImage Upload(const Image& src, int device = -1);
		ImageEx Download() const;

		DeviceImage Convert(Image::Layout layout, SampleFormat format, ConvertMode mode = ConvertMode::ReferenceExact, bool flipRows = false) const;
		DeviceImage ToRGB8(bool flipRows = false) const; // U16 -> U8 through the engine's table, then -> RGB (and the row mirror), one kernel
		DeviceImage Flip() const;
		DeviceImage Downsample2x() const; // the next mip level
		// six faces that are views of ONE contiguous face-major allocation (+X,-X,+Y,-Y,+Z,-Z) -- the
		// layout the Vulkan backend concatenates by hand (engine/src/vulkan/impls/VulkanCubemap.cpp:30-41)
		std::array<DeviceImage, 6> EquirectToCubemap(unsigned int faceSize) const;
		// faces in another layout / format: one fused pass where Spec B.4 defines it, else project + Convert
		std::array<DeviceImage, 6> EquirectToCubemap(unsigned int faceSize, Image::Layout layout, SampleFormat format, bool flipRows = false) const;

		unsigned int Width() const { return w_; }
		unsigned int Height() const { return h_; }
		Image::Layout Layout() const { return layout_; }
		SampleFormat Format() const { return format_; }
		int Device() const { return device_; }
		void* Data() const { return ptr_; } // device pointer, tightly packed rows, top row first
		std::size_t Bytes() const { return static_cast<std::size_t>(w_) * h_ * BytesPerPixel(layout_, format_); }
		bool Empty() const { return ptr_ == nullptr; }

	  private:
		std::shared_ptr<void> owner_; // the allocation (possibly shared with sibling views)
		void* ptr_ = nullptr;
		unsigned int w_ = 0, h_ = 0;
		Image::Layout layout_ = Image::Layout::RGBA;
		SampleFormat format_ = SampleFormat::U8;
		int device_ = -1;
	};
}
