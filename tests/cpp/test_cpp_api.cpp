// test_cpp_api.cpp -- the reference's C++ interface (include/libcacaoimage.hpp) exercised the way
// engine/src/core/Cubemap.cpp:20-32 and Tex2D.cpp:18 use it, against the known-answer vectors
// extracted from the compiled reference.  Built and run by tests/test_cpp_api.py.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "libcacaoimage.hpp"
#include "libcacaoimage_ext.hpp"

using namespace libcacaoimage;

static int failures = 0;
#define EXPECT(cond)                                                    \
	do {                                                                \
		if(!(cond)) {                                                   \
			std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
			++failures;                                                 \
		}                                                               \
	} while(0)

static Image make8(unsigned w, unsigned h, Image::Layout l, std::vector<unsigned char> d) {
	Image im = {};
	im.w = w;
	im.h = h;
	im.layout = l;
	im.bitsPerChannel = 8;
	im.data = std::move(d);
	im.format = Image::Format::TIFF;
	im.quality = 42;
	im.lossy = true;
	return im;
}
static Image make16(unsigned w, unsigned h, Image::Layout l, const std::vector<unsigned short>& d) {
	Image im = make8(w, h, l, {});
	im.bitsPerChannel = 16;
	im.data.resize(d.size() * 2);
	std::memcpy(im.data.data(), d.data(), im.data.size());
	return im;
}
static std::vector<unsigned short> as16(const Image& im) {
	std::vector<unsigned short> v(im.data.size() / 2);
	std::memcpy(v.data(), im.data.data(), im.data.size());
	return v;
}
template<class F>
static std::string thrown(F f) {
	try {
		f();
	} catch(const std::runtime_error& e) {
		return e.what();
	}
	return "";
}

int main(int argc, char** argv) {
	// argv[1] = logical ranks to spread multi-device calls over, argv[2] = devices in the box: the ranks go round
	// robin over the devices (three ranks on device 0 of a one-GPU box still take the multi-threaded path)
	const int n_devices = argc > 1 ? std::atoi(argv[1]) : 1;
	const int n_physical = argc > 2 ? std::max(1, std::atoi(argv[2])) : n_devices;
	static_assert(sizeof(Image) == 56, "Image layout must match the reference (libstdc++ x86-64)");
	using L = Image::Layout;
	// ---- ChangeChannelLayout, 8 bit
	const Image g8 = make8(3, 1, L::Grayscale, {10, 20, 30});
	EXPECT(ChangeChannelLayout(g8, L::RGB).data == (std::vector<unsigned char>{10, 10, 10, 20, 20, 20, 30, 30, 30}));
	EXPECT(ChangeChannelLayout(g8, L::RGBA).data == (std::vector<unsigned char>{10, 10, 10, 255, 10, 10, 10, 255, 10, 10, 10, 255})); // pixel-0 quirk
	const Image rgb8 = make8(3, 1, L::RGB, {255, 255, 255, 1, 2, 3, 200, 100, 50});
	EXPECT(ChangeChannelLayout(rgb8, L::RGBA).data == (std::vector<unsigned char>{255, 255, 255, 255, 1, 2, 3, 255, 200, 100, 50, 255}));
	EXPECT(ChangeChannelLayout(rgb8, L::Grayscale).data == (std::vector<unsigned char>{254, 1, 116}));
	const Image rgba8 = make8(3, 1, L::RGBA, {255, 255, 255, 7, 1, 2, 3, 8, 200, 100, 50, 9});
	const Image r = ChangeChannelLayout(rgba8, L::RGB);
	EXPECT(r.data == (std::vector<unsigned char>{255, 255, 255, 1, 2, 3, 200, 100, 50}));
	EXPECT(r.w == 3 && r.h == 1 && r.layout == L::RGB && r.bitsPerChannel == 8 && r.format == Image::Format::TIFF && r.quality == 42 && r.lossy);
	EXPECT(rgba8.data.size() == 12 && rgba8.data[3] == 7); // source untouched
	// ---- ChangeChannelLayout, 16 bit
	const Image rgb16 = make16(3, 1, L::RGB, {65535, 65535, 65535, 100, 200, 300, 1000, 100, 50});
	EXPECT(as16(ChangeChannelLayout(rgb16, L::RGBA)) == (std::vector<unsigned short>{65535, 65535, 65535, 65535, 100, 200, 300, 65535, 1000, 100, 50, 65535}));
	EXPECT(as16(ChangeChannelLayout(rgb16, L::Grayscale)) == (std::vector<unsigned short>{255, 185, 255})); // clamp-255 quirk
	const Image g16 = make16(3, 1, L::Grayscale, {1000, 20000, 65535});
	EXPECT(as16(ChangeChannelLayout(g16, L::RGBA)) == (std::vector<unsigned short>{1000, 1000, 1000, 65535, 1000, 1000, 1000, 65535, 1000, 1000, 1000, 65535}));
	// ---- Convert16To8BitColor
	EXPECT(Convert16To8BitColor(rgb16).data == (std::vector<unsigned char>{254, 254, 254, 5, 10, 14, 33, 5, 2}));
	const Image rgba16 = make16(2, 1, L::RGBA, {65535, 32768, 0, 65535, 100, 200, 300, 400});
	const Image d8 = Convert16To8BitColor(rgba16);
	EXPECT(d8.data == (std::vector<unsigned char>{254, 187, 0, 254, 5, 10, 14, 18}) && d8.bitsPerChannel == 8 && d8.layout == L::RGBA && d8.quality == 42);
	// ---- the engine's per-face sequence (Cubemap.cpp:28-31) and its fused form
	Image face = rgba16;
	if(face.bitsPerChannel == 16) face = Convert16To8BitColor(face);
	if(face.layout != L::RGB) face = ChangeChannelLayout(face, L::RGB);
	EXPECT(face.data == (std::vector<unsigned char>{254, 187, 0, 5, 10, 14}));
	EXPECT(ToRGB8(rgba16).data == face.data);
	// ---- errors: the reference's texts (Layout.cpp:9, Colordepth.cpp:8, Flip.cpp:8-10)
	EXPECT(thrown([&] { ChangeChannelLayout(rgb8, L::RGB); }) == "Cannot change an image's channel layout to its current layout!");
	EXPECT(thrown([&] { Convert16To8BitColor(rgb8); }) == "Source image for 16 to 8-bit color conversion does not have a 16-bit color depth!");
	EXPECT(thrown([&] { Flip(make8(0, 3, L::RGB, {})); }) == "Cannot flip an image with zeroed dimensions!");
	Image bad = make8(1, 2, L::RGB, {1, 2, 3, 4, 5, 6});
	bad.bitsPerChannel = 12;
	EXPECT(thrown([&] { Flip(bad); }) == "Invalid color depth; only 8 and 16 are allowed.");
	EXPECT(thrown([&] { Flip(make8(1, 2, L::RGB, {})); }) == "Cannot flip an image with a zero-sized data buffer!");
	// ---- empty and ragged images
	EXPECT(ChangeChannelLayout(make8(0, 0, L::RGBA, {}), L::RGB).data.empty());
	EXPECT(thrown([&] { ChangeChannelLayout(make8(4, 4, L::RGB, std::vector<unsigned char>(47)), L::RGBA); }) == "Image data buffer is smaller than w * h * channels * bytes per channel!");
	// ---- Flip: intended semantics, source untouched
	const Image two = make8(2, 3, L::Grayscale, {1, 2, 3, 4, 5, 6});
	EXPECT(Flip(two).data == (std::vector<unsigned char>{5, 6, 3, 4, 1, 2}) && two.data[0] == 1);
	EXPECT(Flip(make8(3, 1, L::Grayscale, {7, 8, 9})).data == (std::vector<unsigned char>{7, 8, 9}));
	// ---- mip levels
	{
		ImageEx m;
		m.w = 4;
		m.h = 2;
		m.layout = L::Grayscale;
		m.format = SampleFormat::U8;
		m.data = {1, 2, 10, 20, 3, 4, 30, 41};
		const ImageEx half = Downsample2x(m);
		EXPECT(half.w == 2 && half.h == 1 && half.data == (std::vector<unsigned char>{3, 25})); // (1+2+3+4+2)>>2, (10+20+30+41+2)>>2
		const auto chain = GenerateMipChain(m);
		EXPECT(chain.size() == 2 && chain[0].data == half.data && chain[1].w == 1 && chain[1].h == 1 && chain[1].data == (std::vector<unsigned char>{14}));
	}
	// ---- texel unpack (engine/src/core/Model.cpp:245-316)
	{
		const unsigned char tx[8] = {0x12, 0x34, 0x56, 0x78, 0xFF, 0x00, 0x80, 0x01};
		const Image bgra = UnpackTexels(tx, 2, 1, "bgra8888");
		EXPECT(bgra.layout == L::RGBA && bgra.bitsPerChannel == 8 && bgra.w == 2 && bgra.h == 1);
		EXPECT(bgra.data == (std::vector<unsigned char>{0x56, 0x34, 0x12, 0x78, 0x80, 0x00, 0xFF, 0x01}));
		const Image rgb565 = UnpackTexels(tx, 2, 1, "rgba5650");
		EXPECT(rgb565.layout == L::RGB && rgb565.data == (std::vector<unsigned char>{148, 211, 181, 255, 0, 0}));
		EXPECT(UnpackTexels(tx, 2, 1, "rgba0800").layout == L::Grayscale);
		EXPECT(thrown([&] { UnpackTexels(tx, 2, 1, "rgba8800"); }) == "Invalid zeroed-channel layout for model embedded texture!");
		EXPECT(thrown([&] { UnpackTexels(tx, 2, 1, "rgba0888"); }) == "Only the alpha channel may be zero for a model embedded texture layout with only one zeroed channel!");
	}
	// ---- extensions: float formats and the projection
	ImageEx ex = FromImage(rgb8);
	const ImageEx f32 = Convert(ex, L::RGBA, SampleFormat::F32);
	const float* fp = reinterpret_cast<const float*>(f32.data.data());
	EXPECT(f32.data.size() == 3 * 4 * 4 && fp[0] == 1.0f && fp[3] == 1.0f && fp[4] == 1.0f / 255.0f && fp[9] == 100.0f / 255.0f);
	EXPECT(ToImage(Convert(f32, L::RGB, SampleFormat::U8)).data == rgb8.data); // round trip
	// batch conversion over however many devices argv says (default 1): results independent of the split
	{
		std::vector<ImageEx> batch(5);
		for(std::size_t k = 0; k < batch.size(); ++k) {
			batch[k].w = 64 + 3 * static_cast<unsigned>(k);
			batch[k].h = 17;
			batch[k].layout = L::RGB;
			batch[k].format = SampleFormat::U8;
			batch[k].data.resize(static_cast<std::size_t>(batch[k].w) * 17 * 3);
			for(std::size_t b = 0; b < batch[k].data.size(); ++b) batch[k].data[b] = static_cast<unsigned char>((b * 7 + k * 13) & 0xFF);
		}
		std::vector<int> devs;
		for(int d = 0; d < n_devices; ++d) devs.push_back(d % n_physical);
		const auto one = ConvertBatch(batch, L::RGBA, SampleFormat::U8);
		const auto many = ConvertBatch(batch, L::RGBA, SampleFormat::U8, ConvertMode::ReferenceExact, devs);
		for(std::size_t k = 0; k < batch.size(); ++k) {
			EXPECT(one[k].data == many[k].data && one[k].data.size() == static_cast<std::size_t>(batch[k].w) * 17 * 4);
			EXPECT(one[k].data[0] == batch[k].data[0] && one[k].data[3] == 255 && one[k].data[4] == batch[k].data[3]);
		}
	}
	ImageEx pano;
	pano.w = 64;
	pano.h = 32;
	pano.layout = L::RGB;
	pano.format = SampleFormat::U8;
	pano.data.assign(64 * 32 * 3, 99);
	const auto cube = EquirectToCubemap(pano, 8);
	for(const auto& f : cube) EXPECT(f.w == 8 && f.h == 8 && f.data.size() == 8 * 8 * 3 && f.data[0] == 99 && f.data.back() == 99);
	// the same over several devices (latitude-ordered unit runs, one pipelined call per device)
	{
		ImageEx noisy;
		noisy.w = 192;
		noisy.h = 96;
		noisy.layout = L::RGBA;
		noisy.format = SampleFormat::U16;
		noisy.data.resize(192 * 96 * 8);
		for(std::size_t b = 0; b < noisy.data.size(); ++b) noisy.data[b] = static_cast<unsigned char>((b * 2654435761u) >> 13);
		std::vector<int> devs;
		for(int d = 0; d < n_devices; ++d) devs.push_back(d % n_physical);
		const auto one = EquirectToCubemap(noisy, 40);
		const auto many = EquirectToCubemap(noisy, 40, devs);
		for(int f = 0; f < 6; ++f) EXPECT(one[f].data == many[f].data && one[f].data.size() == 40 * 40 * 8);
		// ---- DeviceImage: the same chain without leaving the GPU equals the host calls stage by stage
		const DeviceImage dpano = DeviceImage::Upload(noisy);
		EXPECT(dpano.Width() == 192 && dpano.Height() == 96 && dpano.Bytes() == noisy.data.size() && !dpano.Empty());
		EXPECT(dpano.Download().data == noisy.data);
		const auto dfaces = dpano.EquirectToCubemap(40);
		for(int f = 0; f < 6; ++f) {
			EXPECT(dfaces[f].Download().data == one[f].data);
			if(f) EXPECT(static_cast<unsigned char*>(dfaces[f].Data()) == static_cast<unsigned char*>(dfaces[0].Data()) + f * dfaces[0].Bytes()); // one contiguous cube
			const DeviceImage chain = dfaces[f].ToRGB8().Flip();
			const ImageEx host_chain = Flip(FromImage(ToRGB8(ToImage(one[f]))));
			EXPECT(chain.Layout() == L::RGB && chain.Format() == SampleFormat::U8 && chain.Download().data == host_chain.data);
			EXPECT(dfaces[f].Convert(L::RGBA, SampleFormat::F16).Download().data == Convert(one[f], L::RGBA, SampleFormat::F16).data);
		}
		EXPECT(thrown([&] { dpano.Convert(L::RGBA, SampleFormat::U16); }) == "Cannot change an image's channel layout to its current layout!");
		EXPECT(thrown([&] { DeviceImage().Download(); }) == "Cannot flip an image with a zero-sized data buffer!" || !thrown([&] { DeviceImage().Download(); }).empty());
		DeviceImage copy = dfaces[3]; // shares the allocation; the cube outlives the array it came from
		EXPECT(copy.Data() == dfaces[3].Data() && copy.Download().data == one[3].data);
	}
	// ---- conversion with the row mirror folded in: the engine's three load-time passes as one
	{
		Image deep16;
		deep16.w = 96;
		deep16.h = 40;
		deep16.layout = L::RGBA;
		deep16.bitsPerChannel = 16;
		deep16.format = Image::Format::PNG;
		deep16.quality = 100;
		deep16.lossy = false;
		deep16.data.resize(96 * 40 * 8);
		for(std::size_t b = 0; b < deep16.data.size(); ++b) deep16.data[b] = static_cast<unsigned char>((b * 2246822519u) >> 11);
		const Image three = Flip(ChangeChannelLayout(Convert16To8BitColor(deep16), L::RGB)); // what the reference runs
		EXPECT(ToRGB8(deep16, true).data == three.data);
		EXPECT(DeviceImage::Upload(deep16).ToRGB8(true).Download().data == three.data);
		const ImageEx ex = FromImage(deep16);
		EXPECT(Convert(ex, L::RGBA, SampleFormat::F16, ConvertMode::ReferenceExact, true).data == Flip(Convert(ex, L::RGBA, SampleFormat::F16)).data);
	}
	// ---- resampling and conversion in one pass (Spec B.4): an RGB32F panorama straight to RGBA16F faces
	// equals project-then-Convert; pairs without a fused form (integer sources) run the two passes inside
	{
		ImageEx hdr;
		hdr.w = 160;
		hdr.h = 80;
		hdr.layout = L::RGB;
		hdr.format = SampleFormat::F32;
		hdr.data.resize(160 * 80 * 12);
		float* px = reinterpret_cast<float*>(hdr.data.data());
		for(std::size_t k = 0; k < 160 * 80 * 3; ++k) px[k] = static_cast<float>((k * 2654435761u) >> 8 & 0xFFFF) / 32768.0f - 0.25f;
		EXPECT(FusedProjectionDefined(L::RGB, SampleFormat::F32, L::RGBA, SampleFormat::F16));
		EXPECT(!FusedProjectionDefined(L::RGBA, SampleFormat::U16, L::RGBA, SampleFormat::F16) && FusedProjectionDefined(L::RGBA, SampleFormat::F32, L::RGB, SampleFormat::U8) && !FusedProjectionDefined(L::RGBA, SampleFormat::F32, L::Grayscale, SampleFormat::F32));
		const auto plain = EquirectToCubemap(hdr, 24);
		const auto fused = EquirectToCubemap(hdr, 24, L::RGBA, SampleFormat::F16);
		const auto quant = EquirectToCubemap(hdr, 24, L::RGB, SampleFormat::U8);
		const auto dfused = DeviceImage::Upload(hdr).EquirectToCubemap(24, L::RGBA, SampleFormat::F16);
		for(int f = 0; f < 6; ++f) {
			EXPECT(fused[f].layout == L::RGBA && fused[f].format == SampleFormat::F16 && fused[f].data.size() == 24 * 24 * 8);
			EXPECT(fused[f].data == Convert(plain[f], L::RGBA, SampleFormat::F16).data);
			EXPECT(quant[f].data == Convert(plain[f], L::RGB, SampleFormat::U8).data);
			EXPECT(dfused[f].Download().data == fused[f].data && dfused[f].Format() == SampleFormat::F16);
		}
		// bottom row first, folded into the same pass
		const auto flipped = EquirectToCubemap(hdr, 24, L::RGBA, SampleFormat::F16, {}, true);
		const auto dflipped = DeviceImage::Upload(hdr).EquirectToCubemap(24, L::RGB, SampleFormat::F32, true);
		for(int f = 0; f < 6; ++f) {
			EXPECT(flipped[f].data == Flip(fused[f]).data);
			EXPECT(dflipped[f].Download().data == Flip(plain[f]).data);
		}
		ImageEx deep;
		deep.w = 64;
		deep.h = 32;
		deep.layout = L::RGBA;
		deep.format = SampleFormat::U16;
		deep.data.resize(64 * 32 * 8);
		for(std::size_t b = 0; b < deep.data.size(); ++b) deep.data[b] = static_cast<unsigned char>((b * 40503u) >> 7);
		const auto two = EquirectToCubemap(deep, 12, L::RGB, SampleFormat::U8); // falls back to two passes
		const auto ref = EquirectToCubemap(deep, 12);
		for(int f = 0; f < 6; ++f) EXPECT(two[f].data == Convert(ref[f], L::RGB, SampleFormat::U8).data);
	}
	// ---- EngineCube: the whole load chain of Cacao::Cubemap (Cubemap.cpp:20-32) + Flip + contiguous faces in one call
	{
		std::array<Image, 6> faces;
		for(int f = 0; f < 6; ++f) {
			faces[f] = {};
			faces[f].w = 48;
			faces[f].h = 32;
			faces[f].layout = L::RGBA;
			faces[f].bitsPerChannel = 16;
			faces[f].data.resize(48 * 32 * 8);
			for(std::size_t b = 0; b < faces[f].data.size(); ++b) faces[f].data[b] = static_cast<unsigned char>(((b + 977 * f) * 2654435761u) >> 11);
		}
		const CubeRGB8 cube = EngineCube(faces, true, true);
		EXPECT(cube.w == 48 && cube.h == 32 && cube.levelOffset.size() == 6 && cube.levelOffset[1] == 6 * 48 * 32 * 3 && cube.FaceBytes(1) == 24 * 16 * 3);
		EXPECT(cube.data.size() == cube.levelOffset[5] + 6 * cube.FaceBytes(5) && cube.FaceBytes(5) == 3);
		for(int f = 0; f < 6; ++f) {
			// the reference's sequence, one face at a time: 16 -> 8 bit, -> RGB, Flip
			const Image step = Flip(ChangeChannelLayout(Convert16To8BitColor(faces[f]), L::RGB));
			EXPECT(std::memcmp(cube.Face(f), step.data.data(), step.data.size()) == 0);
		}
		// faces of different kinds: each is brought to RGB8 by its own rule first
		faces[2] = ChangeChannelLayout(Convert16To8BitColor(faces[2]), L::Grayscale);
		faces[4] = Convert16To8BitColor(faces[4]);
		const CubeRGB8 mixed = EngineCube(faces);
		for(int f = 0; f < 6; ++f) {
			const Image step = ToRGB8(faces[f]);
			EXPECT(std::memcmp(mixed.Face(f), step.data.data(), step.data.size()) == 0);
		}
		// EngineTexture: Tex2D's chain -- 16 -> 8 bit with the layout kept, Flip, mip levels -- against the calls it replaces
		const TextureU8 tex = EngineTexture(faces[0], true, true);
		const Image step = Flip(Convert16To8BitColor(faces[0]));
		EXPECT(tex.w == 48 && tex.h == 32 && tex.layout == L::RGBA && tex.levelOffset.size() == 6 && tex.levelOffset[1] == 48 * 32 * 4);
		EXPECT(std::memcmp(tex.data.data(), step.data.data(), step.data.size()) == 0);
		const ImageEx mip1 = Downsample2x(FromImage(step));
		EXPECT(std::memcmp(tex.data.data() + tex.levelOffset[1], mip1.data.data(), mip1.data.size()) == 0);
		faces[3].w = 47;
		bool threw = false;
		try {
			EngineCube(faces);
		} catch(const std::runtime_error& e) {
			threw = std::string(e.what()) == "Not all cubemap faces are the same size!";
		}
		EXPECT(threw);
	}
	std::printf(failures ? "%d FAILURES\n" : "cpp api ok\n", failures);
	return failures ? 1 : 0;
}
