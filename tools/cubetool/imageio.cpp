// imageio.cpp -- see imageio.hpp.
#include "imageio.hpp"

#include "vp8l.hpp"

#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

using libcacaoimage::Image;
using libcacaoimage::ImageEx;
using libcacaoimage::SampleFormat;

namespace {
	std::vector<unsigned char> slurp(const std::filesystem::path& p) {
		std::ifstream in(p, std::ios::binary);
		if(!in.is_open()) throw std::runtime_error("Failed to open input file stream for reading!");
		return std::vector<unsigned char>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
	}

	uint32_t be32(const unsigned char* p) {
		return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3];
	}
	void put_be32(std::vector<unsigned char>& v, uint32_t x) {
		v.push_back(x >> 24);
		v.push_back(x >> 16);
		v.push_back(x >> 8);
		v.push_back(x);
	}

	Image::Layout layout_of(int ch) {
		if(ch == 1) return Image::Layout::Grayscale;
		if(ch == 3) return Image::Layout::RGB;
		if(ch == 4) return Image::Layout::RGBA;
		throw std::runtime_error("Unsupported channel count (only 1, 3 or 4 channels are allowed)");
	}

	// ---------------------------------------------------------------- PNG
	int paeth(int a, int b, int c) {
		const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
		return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
	}

	ImageEx read_png(const std::vector<unsigned char>& f) {
		size_t pos = 8;
		uint32_t w = 0, h = 0;
		int depth = 0, ctype = -1, interlace = 0;
		std::vector<unsigned char> idat;
		while(pos + 12 <= f.size()) {
			const uint32_t len = be32(&f[pos]);
			const std::string type(reinterpret_cast<const char*>(&f[pos + 4]), 4);
			if(pos + 12 + len > f.size()) throw std::runtime_error("PNG chunk runs past the end of the file");
			const unsigned char* d = &f[pos + 8];
			if(type == "IHDR") {
				if(len != 13) throw std::runtime_error("PNG IHDR chunk has the wrong size");
				w = be32(d);
				h = be32(d + 4);
				depth = d[8];
				ctype = d[9];
				interlace = d[12];
			} else if(type == "IDAT") {
				idat.insert(idat.end(), d, d + len);
			} else if(type == "IEND") {
				break;
			}
			pos += 12 + len;
		}
		if(w == 0 || h == 0) throw std::runtime_error("PNG has no IHDR");
		if(w > 65536 || h > 65536) throw std::runtime_error("PNG dimensions beyond 65536 pixels are not supported");
		if(interlace) throw std::runtime_error("Interlaced PNG files are not supported");
		if(depth != 8 && depth != 16) throw std::runtime_error("Only 8- and 16-bit PNG files are supported");
		int ch;
		switch(ctype) {
			case 0: ch = 1; break;
			case 2: ch = 3; break;
			case 6: ch = 4; break;
			default: throw std::runtime_error("Only grayscale, RGB and RGBA PNG files are supported (no palette, no gray+alpha)");
		}
		const size_t bpp = size_t(ch) * depth / 8, pitch = bpp * w;
		std::vector<unsigned char> raw((pitch + 1) * h);
		uLongf rawLen = raw.size();
		if(uncompress(raw.data(), &rawLen, idat.data(), idat.size()) != Z_OK || rawLen != raw.size()) throw std::runtime_error("PNG pixel data failed to inflate");
		ImageEx out;
		out.w = w;
		out.h = h;
		out.layout = layout_of(ch);
		out.format = depth == 8 ? SampleFormat::U8 : SampleFormat::U16;
		out.data.resize(pitch * h);
		std::vector<unsigned char> zero(pitch, 0);
		for(uint32_t y = 0; y < h; ++y) {
			const unsigned char* in = &raw[(pitch + 1) * y];
			unsigned char* cur = &out.data[pitch * y];
			const unsigned char* up = y ? &out.data[pitch * (y - 1)] : zero.data();
			const int filter = in[0];
			for(size_t x = 0; x < pitch; ++x) {
				const int a = x >= bpp ? cur[x - bpp] : 0, b = up[x], c = x >= bpp ? up[x - bpp] : 0;
				int v = in[1 + x];
				switch(filter) {
					case 0: break;
					case 1: v += a; break;
					case 2: v += b; break;
					case 3: v += (a + b) / 2; break;
					case 4: v += paeth(a, b, c); break;
					default: throw std::runtime_error("PNG uses an unknown row filter");
				}
				cur[x] = static_cast<unsigned char>(v);
			}
		}
		if(depth == 16) // PNG is big-endian; Image buffers are native little-endian (Colordepth.cpp:32)
			for(size_t i = 0; i + 1 < out.data.size(); i += 2) std::swap(out.data[i], out.data[i + 1]);
		return out;
	}

	void write_png(const ImageEx& img, const std::filesystem::path& path) {
		if(img.format != SampleFormat::U8 && img.format != SampleFormat::U16) throw std::runtime_error("PNG output needs 8- or 16-bit integer samples");
		const int ch = static_cast<int>(img.layout), depth = img.format == SampleFormat::U8 ? 8 : 16;
		const size_t bpp = size_t(ch) * depth / 8, pitch = bpp * img.w;
		std::vector<unsigned char> raw((pitch + 1) * img.h);
		for(uint32_t y = 0; y < img.h; ++y) {
			unsigned char* o = &raw[(pitch + 1) * y];
			o[0] = 0;
			std::memcpy(o + 1, &img.data[pitch * y], pitch);
			if(depth == 16)
				for(size_t x = 0; x + 1 < pitch; x += 2) std::swap(o[1 + x], o[2 + x]);
		}
		uLongf zlen = compressBound(raw.size());
		std::vector<unsigned char> z(zlen);
		if(compress2(z.data(), &zlen, raw.data(), raw.size(), 1) != Z_OK) throw std::runtime_error("PNG deflate failed");
		z.resize(zlen);
		std::vector<unsigned char> f = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
		auto chunk = [&f](const char* type, const std::vector<unsigned char>& body) {
			put_be32(f, static_cast<uint32_t>(body.size()));
			const size_t start = f.size();
			f.insert(f.end(), type, type + 4);
			f.insert(f.end(), body.begin(), body.end());
			put_be32(f, static_cast<uint32_t>(crc32(0, &f[start], static_cast<uInt>(f.size() - start))));
		};
		std::vector<unsigned char> ihdr;
		put_be32(ihdr, img.w);
		put_be32(ihdr, img.h);
		ihdr.push_back(static_cast<unsigned char>(depth));
		ihdr.push_back(ch == 1 ? 0 : (ch == 3 ? 2 : 6));
		ihdr.push_back(0);
		ihdr.push_back(0);
		ihdr.push_back(0);
		chunk("IHDR", ihdr);
		chunk("IDAT", z);
		chunk("IEND", {});
		std::ofstream out(path, std::ios::binary);
		if(!out.is_open()) throw std::runtime_error("Failed to open output file stream for writing!");
		out.write(reinterpret_cast<const char*>(f.data()), static_cast<std::streamsize>(f.size()));
	}

	// ---------------------------------------------------------------- PFM / PNM (text header, binary body)
	std::string next_token(const std::vector<unsigned char>& f, size_t& pos) {
		while(pos < f.size()) {
			if(f[pos] == '#') {
				while(pos < f.size() && f[pos] != '\n') ++pos;
			} else if(std::isspace(f[pos])) {
				++pos;
			} else {
				break;
			}
		}
		std::string t;
		while(pos < f.size() && !std::isspace(f[pos])) t.push_back(static_cast<char>(f[pos++]));
		return t;
	}

	ImageEx read_pfm(const std::vector<unsigned char>& f) {
		size_t pos = 0;
		const std::string magic = next_token(f, pos);
		const int ch = magic == "PF" ? 3 : 1;
		const unsigned w = std::stoul(next_token(f, pos)), h = std::stoul(next_token(f, pos));
		const float scale = std::stof(next_token(f, pos));
		++pos; // single whitespace after the header
		if(scale > 0) throw std::runtime_error("Big-endian PFM files are not supported");
		const size_t pitch = size_t(w) * ch * 4;
		if(pos + pitch * h > f.size()) throw std::runtime_error("PFM file is truncated");
		ImageEx out;
		out.w = w;
		out.h = h;
		out.layout = layout_of(ch);
		out.format = SampleFormat::F32;
		out.data.resize(pitch * h);
		for(unsigned y = 0; y < h; ++y) // PFM stores the bottom row first
			std::memcpy(&out.data[pitch * y], &f[pos + pitch * (h - 1 - y)], pitch);
		return out;
	}

	void write_pfm(const ImageEx& img, const std::filesystem::path& path) {
		const int ch = static_cast<int>(img.layout);
		if(img.format != SampleFormat::F32 || ch == 4) throw std::runtime_error("PFM output needs 32-bit float grayscale or RGB samples");
		std::ofstream out(path, std::ios::binary);
		if(!out.is_open()) throw std::runtime_error("Failed to open output file stream for writing!");
		out << (ch == 3 ? "PF" : "Pf") << "\n" << img.w << " " << img.h << "\n-1.0\n";
		const size_t pitch = size_t(img.w) * ch * 4;
		for(unsigned y = 0; y < img.h; ++y) out.write(reinterpret_cast<const char*>(&img.data[pitch * (img.h - 1 - y)]), static_cast<std::streamsize>(pitch));
	}

	ImageEx read_pnm(const std::vector<unsigned char>& f) {
		size_t pos = 0;
		const std::string magic = next_token(f, pos);
		const int ch = magic == "P6" ? 3 : 1;
		const unsigned w = std::stoul(next_token(f, pos)), h = std::stoul(next_token(f, pos)), maxv = std::stoul(next_token(f, pos));
		++pos;
		const int bytes = maxv > 255 ? 2 : 1;
		const size_t n = size_t(w) * h * ch * bytes;
		if(pos + n > f.size()) throw std::runtime_error("PNM file is truncated");
		ImageEx out;
		out.w = w;
		out.h = h;
		out.layout = layout_of(ch);
		out.format = bytes == 1 ? SampleFormat::U8 : SampleFormat::U16;
		out.data.assign(f.begin() + pos, f.begin() + pos + n);
		if(bytes == 2)
			for(size_t i = 0; i + 1 < n; i += 2) std::swap(out.data[i], out.data[i + 1]);
		return out;
	}

	// ---------------------------------------------------------------- NPY (version 1.0 / 2.0, C order)
	ImageEx read_npy(const std::vector<unsigned char>& f) {
		if(f.size() < 12) throw std::runtime_error("NPY file is truncated");
		const int major = f[6];
		size_t hlen, hstart;
		if(major == 1) {
			hlen = f[8] | (f[9] << 8);
			hstart = 10;
		} else {
			hlen = f[8] | (f[9] << 8) | (f[10] << 16) | (size_t(f[11]) << 24);
			hstart = 12;
		}
		const std::string hdr(reinterpret_cast<const char*>(&f[hstart]), hlen);
		auto field = [&hdr](const std::string& key) {
			const size_t k = hdr.find("'" + key + "'");
			if(k == std::string::npos) throw std::runtime_error("NPY header lacks '" + key + "'");
			return hdr.substr(hdr.find(':', k) + 1);
		};
		const std::string descr = field("descr"), order = field("fortran_order"), shape = field("shape");
		if(order.find("False") > 8) throw std::runtime_error("Fortran-ordered NPY arrays are not supported");
		const size_t q0 = descr.find('\''), q1 = descr.find('\'', q0 + 1);
		const std::string dt = descr.substr(q0 + 1, q1 - q0 - 1);
		ImageEx out;
		if(dt == "|u1" || dt == "<u1")
			out.format = SampleFormat::U8;
		else if(dt == "<u2")
			out.format = SampleFormat::U16;
		else if(dt == "<f2")
			out.format = SampleFormat::F16;
		else if(dt == "<f4")
			out.format = SampleFormat::F32;
		else
			throw std::runtime_error("NPY dtype " + dt + " is not supported (u1, <u2, <f2, <f4)");
		std::vector<unsigned long> dims;
		const size_t p0 = shape.find('('), p1 = shape.find(')');
		std::stringstream ss(shape.substr(p0 + 1, p1 - p0 - 1));
		std::string item;
		while(std::getline(ss, item, ','))
			if(item.find_first_of("0123456789") != std::string::npos) dims.push_back(std::stoul(item));
		if(dims.size() != 2 && dims.size() != 3) throw std::runtime_error("NPY array must have shape (H, W) or (H, W, C)");
		out.h = static_cast<unsigned>(dims[0]);
		out.w = static_cast<unsigned>(dims[1]);
		out.layout = layout_of(dims.size() == 3 ? static_cast<int>(dims[2]) : 1);
		const size_t n = size_t(out.w) * out.h * libcacaoimage::BytesPerPixel(out.layout, out.format);
		if(hstart + hlen + n > f.size()) throw std::runtime_error("NPY file is truncated");
		out.data.assign(f.begin() + hstart + hlen, f.begin() + hstart + hlen + n);
		return out;
	}

	void write_npy(const ImageEx& img, const std::filesystem::path& path) {
		const char* dt = img.format == SampleFormat::U8 ? "|u1" : img.format == SampleFormat::U16 ? "<u2" : img.format == SampleFormat::F16 ? "<f2" : "<f4";
		std::stringstream h;
		h << "{'descr': '" << dt << "', 'fortran_order': False, 'shape': (" << img.h << ", " << img.w << ", " << static_cast<int>(img.layout) << "), }";
		std::string hdr = h.str();
		while((10 + hdr.size() + 1) % 64 != 0) hdr.push_back(' ');
		hdr.push_back('\n');
		std::ofstream out(path, std::ios::binary);
		if(!out.is_open()) throw std::runtime_error("Failed to open output file stream for writing!");
		const unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, static_cast<unsigned char>(hdr.size() & 0xFF), static_cast<unsigned char>(hdr.size() >> 8)};
		out.write(reinterpret_cast<const char*>(pre), 10);
		out.write(hdr.data(), static_cast<std::streamsize>(hdr.size()));
		out.write(reinterpret_cast<const char*>(img.data.data()), static_cast<std::streamsize>(img.data.size()));
	}
}

namespace cubeio {
	ImageEx ReadImage(const std::filesystem::path& path) {
		return DecodeImage(slurp(path));
	}

	ImageEx DecodeImage(const std::vector<unsigned char>& f) {
		static const unsigned char pngSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
		if(f.size() >= 8 && std::memcmp(f.data(), pngSig, 8) == 0) return read_png(f);
		if(f.size() >= 6 && std::memcmp(f.data(), "\x93NUMPY", 6) == 0) return read_npy(f);
		if(f.size() >= 2 && f[0] == 'P' && (f[1] == 'F' || f[1] == 'f')) return read_pfm(f);
		if(f.size() >= 2 && f[0] == 'P' && (f[1] == '5' || f[1] == '6')) return read_pnm(f);
		if(vp8l::IsWebP(f.data(), f.size())) {
			vp8l::Decoded d = vp8l::Decode(f.data(), f.size());
			ImageEx img;
			img.w = d.w;
			img.h = d.h;
			img.layout = layout_of(d.channels);
			img.format = SampleFormat::U8;
			img.data = std::move(d.pixels);
			return img;
		}
		throw std::runtime_error("Unrecognised image format (this build reads PNG, lossless WebP, PFM, NPY and binary PGM/PPM)");
	}

	void WriteImage(const ImageEx& img, const std::filesystem::path& path, const std::string& format) {
		if(format == "png")
			write_png(img, path);
		else if(format == "pfm")
			write_pfm(img, path);
		else if(format == "npy")
			write_npy(img, path);
		else if(format == "webp") {
			if(img.format != SampleFormat::U8 || img.layout == Image::Layout::Grayscale) throw std::runtime_error("WebP output needs 8-bit RGB or RGBA samples");
			const std::vector<uint8_t> file = vp8l::Encode(img.data.data(), img.w, img.h, static_cast<int>(img.layout));
			std::ofstream out(path, std::ios::binary);
			if(!out.is_open()) throw std::runtime_error("Failed to open output file stream for writing!");
			out.write(reinterpret_cast<const char*>(file.data()), static_cast<std::streamsize>(file.size()));
		} else
			throw std::runtime_error("Unknown output format '" + format + "' (png, webp, pfm, npy)");
	}

	std::string DefaultFormat(const ImageEx& img) {
		return (img.format == SampleFormat::U8 || img.format == SampleFormat::U16) ? "png" : "npy";
	}
}
