// xjc.cpp -- see xjc.hpp.
#include "xjc.hpp"

#include <dlfcn.h>

#include <cctype>
#include <cstring>
#include <map>
#include <sstream>
#include <stdexcept>

#include "imageio.hpp"
#include "vp8l.hpp"

namespace xjc {
	namespace {
		void check(bool cond, const std::string& msg) {
			if(!cond) throw std::runtime_error(msg);
		}

		// ---- libbz2, bound at run time ---------------------------------------------------------
		using CompressFn = int (*)(char*, unsigned int*, char*, unsigned int, int, int, int);
		using DecompressFn = int (*)(char*, unsigned int*, char*, unsigned int, int, int);
		struct Bz2 {
			CompressFn compress = nullptr;
			DecompressFn decompress = nullptr;
			Bz2() {
				void* h = nullptr;
				for(const char* name : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}) {
					h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
					if(h) break;
				}
				if(!h) return;
				compress = reinterpret_cast<CompressFn>(dlsym(h, "BZ2_bzBuffToBuffCompress"));
				decompress = reinterpret_cast<DecompressFn>(dlsym(h, "BZ2_bzBuffToBuffDecompress"));
			}
		};
		const Bz2& bz2() {
			static const Bz2 lib;
			check(lib.compress && lib.decompress, "libbz2 (bzip2 runtime library) was not found on this system; packed containers cannot be read or written!");
			return lib;
		}

		std::string trim(const std::string& s) {
			size_t a = 0, b = s.size();
			while(a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
			while(b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
			return s.substr(a, b - a);
		}
	}

	Container ReadContainer(const std::vector<unsigned char>& f) {
		// PackedContainer.cpp:47-49 tests the magic with masks; a real file has exactly CA CA 00
		check(f.size() >= 3 && (f[0] & 0xCA) == 0xCA && (f[1] & 0xCA) == 0xCA && f[2] == 0x00, "Stream is not of a Cacao Engine packed object!");
		check(f.size() >= 4 && (f[3] == 0xC4 || f[3] == 0x1B || f[3] == 0x3E || f[3] == 0xAA || f[3] == 0x7A), "Stream is not a valid Cacao Engine format!");
		check(f.size() >= 14, "Failed to decompress packed container payload!");
		Container c;
		c.code = f[3];
		std::memcpy(&c.version, f.data() + 4, 2);
		uint64_t size = 0;
		std::memcpy(&size, f.data() + 6, 8);
		check(size <= 0xFFFFFFFFull, "Failed to decompress packed container payload!"); // the bzip2 buffer API is 32-bit
		c.payload.resize(size);
		unsigned int outSize = static_cast<unsigned int>(size);
		std::vector<unsigned char> in(f.begin() + 14, f.end());
		const int status = bz2().decompress(reinterpret_cast<char*>(c.payload.data()), &outSize, reinterpret_cast<char*>(in.data()), static_cast<unsigned int>(in.size()), 0, 0);
		check(status == 0 && outSize == size, "Failed to decompress packed container payload!");
		check(!c.payload.empty(), "Cannot make empty PackedContainer!");
		return c;
	}

	std::vector<unsigned char> WriteContainer(const Container& c) {
		check(!c.payload.empty(), "Cannot make empty PackedContainer!");
		check(c.payload.size() <= 0xFFFFFFFFull, "Failed to compress payload for packed container export!");
		std::vector<unsigned char> out = {0xCA, 0xCA, 0x00, c.code};
		out.resize(14);
		std::memcpy(out.data() + 4, &c.version, 2);
		const uint64_t size = c.payload.size();
		std::memcpy(out.data() + 6, &size, 8);
		// level 9, work factor 30, destination sized as the bzip2 manual says (PackedContainer.cpp:143-146)
		std::vector<unsigned char> comp(static_cast<size_t>(size * 1.01) + 600);
		unsigned int destSize = static_cast<unsigned int>(comp.size());
		std::vector<unsigned char> src(c.payload);
		const int status = bz2().compress(reinterpret_cast<char*>(comp.data()), &destSize, reinterpret_cast<char*>(src.data()), static_cast<unsigned int>(size), 9, 0, 30);
		check(status == 0, "Failed to compress payload for packed container export!");
		out.insert(out.end(), comp.begin(), comp.begin() + destSize);
		return out;
	}

	std::array<std::vector<unsigned char>, 6> SplitCubemapPayload(const Container& c) {
		check(c.code == kCubemapCode, "Packed container provided for cubemap decoding is not a cubemap!");
		check(c.payload.size() > 48, "Cubemap packed container is too small to contain face size data!");
		uint64_t sizes[6], sum = 48;
		std::memcpy(sizes, c.payload.data(), 48);
		for(uint64_t s : sizes) {
			check(s <= c.payload.size(), "Cubemap packed container is too small to contain face data of given sizes!");
			sum += s;
		}
		check(c.payload.size() >= sum, "Cubemap packed container is too small to contain face data of given sizes!");
		std::array<std::vector<unsigned char>, 6> blobs;
		size_t off = 48;
		for(int i = 0; i < 6; ++i) {
			blobs[i].assign(c.payload.begin() + off, c.payload.begin() + off + sizes[i]);
			off += sizes[i];
		}
		return blobs;
	}

	Container MakeCubemapContainer(const std::array<std::vector<unsigned char>, 6>& blobs) {
		Container c;
		c.code = kCubemapCode;
		c.version = 1; // PackedEncode.cpp:47
		c.payload.resize(48);
		for(int i = 0; i < 6; ++i) {
			const uint64_t s = blobs[i].size();
			std::memcpy(c.payload.data() + 8 * i, &s, 8);
		}
		for(const auto& b : blobs) c.payload.insert(c.payload.end(), b.begin(), b.end());
		return c;
	}

	// The subset of YAML an .ajc needs: one block mapping of scalar keys to scalar values, optional
	// document markers, comments, and single- or double-quoted scalars.
	std::array<std::string, 6> ReadAjc(const std::string& text) {
		std::map<std::string, std::string> map;
		std::map<std::string, bool> nonScalar;
		std::istringstream in(text);
		std::string line, lastKey;
		while(std::getline(in, line)) {
			if(!line.empty() && line.back() == '\r') line.pop_back();
			const std::string t = trim(line);
			if(t.empty() || t[0] == '#' || t == "---" || t == "...") continue;
			const bool indented = std::isspace(static_cast<unsigned char>(line[0])) != 0;
			if(indented || t[0] == '-') {
				// nested content under the previous key: that key is not a scalar
				if(!lastKey.empty()) nonScalar[lastKey] = true;
				continue;
			}
			const size_t colon = t.find(':');
			check(colon != std::string::npos && (colon + 1 == t.size() || std::isspace(static_cast<unsigned char>(t[colon + 1]))), "Failed to parse unpacked cubemap data stream!");
			std::string key = trim(t.substr(0, colon)), val = trim(t.substr(colon + 1));
			if(key.size() >= 2 && (key.front() == '"' || key.front() == '\'') && key.back() == key.front()) key = key.substr(1, key.size() - 2);
			if(val.size() >= 2 && (val.front() == '"' || val.front() == '\'') && val.back() == val.front()) {
				val = val.substr(1, val.size() - 2);
			} else {
				const size_t hash = val.find(" #");
				if(hash != std::string::npos) val = trim(val.substr(0, hash));
			}
			if(!val.empty() && (val[0] == '[' || val[0] == '{')) nonScalar[key] = true;
			map[key] = val;
			lastKey = key;
		}
		static const char* keys[6] = {"right", "left", "top", "bottom", "front", "back"};
		static const char* what[6] = {"right (positive X) face", "left (negative X) face", "top (positive Y) face", "bottom (negative Y) face", "front (positive Z) face", "back (negative Z) face"};
		std::array<std::string, 6> out;
		for(int i = 0; i < 6; ++i) {
			// message format of libs/formats/src/YAMLValidate.hpp:20-37
			const std::string pre = std::string("While parsing unpacked cubemap data, ") + what[i] + " node is invalid: ";
			const auto it = map.find(keys[i]);
			check(it != map.end(), pre + "Node does not exist");
			check(!nonScalar[keys[i]] && !it->second.empty(), pre + "Node is of improper type!");
			out[i] = it->second;
		}
		return out;
	}

	libcacaoimage::Image DecodeFace(const std::vector<unsigned char>& b) {
		using libcacaoimage::Image;
		Image img;
		if(vp8l::IsWebP(b.data(), b.size())) {
			vp8l::Decoded d;
			try {
				d = vp8l::Decode(b.data(), b.size());
			} catch(const std::exception& e) {
				throw std::runtime_error(std::string("Failed to decode WebP data! (") + e.what() + ")");
			}
			// metadata as libs/image/src/WebP.cpp:47-54 sets it
			img.w = d.w;
			img.h = d.h;
			img.format = Image::Format::WebP;
			img.bitsPerChannel = 8;
			img.layout = d.channels == 4 ? Image::Layout::RGBA : Image::Layout::RGB;
			img.quality = 80;
			img.lossy = true;
			img.data = std::move(d.pixels);
			return img;
		}
		if(b.size() >= 3 && b[0] == 0xFF && b[1] == 0xD8 && b[2] == 0xFF) throw std::runtime_error("JPEG decoding is not part of this build (libjpeg-turbo is absent); convert the face to PNG first.");
		if(b.size() >= 4 && b[0] == 'I' && b[1] == 'I' && b[2] == '*' && b[3] == 0) throw std::runtime_error("TIFF decoding is not part of this build (libtiff is absent); convert the face to PNG first.");
		libcacaoimage::ImageEx ex;
		try {
			ex = cubeio::DecodeImage(b);
		} catch(const std::exception&) { // malformed input of any kind (length_error / bad_alloc from a hostile header included)
			throw std::runtime_error("Unknown file type!"); // Forwarding.cpp:66
		}
		check(ex.format == libcacaoimage::SampleFormat::U8 || ex.format == libcacaoimage::SampleFormat::U16, "Cubemap faces must have 8 or 16-bit integer samples!");
		img = libcacaoimage::ToImage(ex);
		img.format = Image::Format::PNG;
		return img;
	}

	std::vector<unsigned char> EncodeFaceWebP(const libcacaoimage::Image& src) {
		using libcacaoimage::Image;
		check(src.w > 0 && src.h > 0, "Cannot encode an image with zeroed dimensions!");
		check(src.bitsPerChannel == 8, "Invalid color depth for WebP encoding; only 8 bits per channel are supported.");
		check(!src.data.empty(), "Cannot encode an image with a zero-sized data buffer!");
		check(src.layout != Image::Layout::Grayscale, "Cannot encode a grayscale image to WebP format!");
		check(src.w <= 16384 && src.h <= 16384, "Failed to encode image to WebP!");
		check(src.data.size() >= static_cast<size_t>(src.w) * src.h * static_cast<size_t>(src.layout), "Failed to encode image to WebP!");
		return vp8l::Encode(src.data.data(), src.w, src.h, static_cast<int>(src.layout));
	}
}
