// create.cpp -- `ce-cubetool create`: unpacked cubemap definition (.ajc) -> packed cubemap (.xjc).
//
// Same contract as the reference subcommand (tools/cubetool/create.cpp:9-102): positional `input`
// (an existing .ajc), required `-o` (a file to overwrite or a nonexistent path), face paths resolved
// relative to the .ajc, faces decoded by magic bytes, re-encoded as lossless WebP, bzip2-packed.
// Errors surface as CUBE_ERROR with the reference's texts.  One addition: --to-rgb8 runs the
// engine's load-time conversion (16 -> 8 bit through the sRGB table, then -> RGB;
// engine/src/core/Cubemap.cpp:28-31) on the GPU as ONE fused kernel per face before packing, so
// 16-bit or grayscale faces -- which the reference's WebP encoder rejects (WebP.cpp:76,78) -- can
// be packed, and the engine finds nothing left to convert.
#include <fstream>
#include <iterator>
#include <stdexcept>

#include "commands.hpp"
#include "libcacaoimage_ext.hpp"
#include "xjc.hpp"

namespace {
	std::vector<unsigned char> readAll(std::ifstream& in) {
		return std::vector<unsigned char>(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
	}
}

int RunCreate(const CreateOptions& opt) {
	//Load the input file
	VLOG_NONL("Opening input file " << opt.input << "... ");
	std::ifstream in(opt.input);
	if(!in.is_open()) {
		CUBE_ERROR("Failed to open input file stream for reading!")
	}
	VLOG("Done.")

	//Decode the definition and the six faces
	VLOG("Loading input files:")
	std::array<libcacaoimage::Image, 6> faces;
	std::array<std::vector<unsigned char>, 6> fileBytes;
	try {
		const std::vector<unsigned char> text = readAll(in);
		const std::array<std::string, 6> paths = xjc::ReadAjc(std::string(text.begin(), text.end()));
		static const char* streamNames[6] = {"Right", "Left", "Top", "Bottom", "Front", "Back"};
		for(int i = 0; i < 6; ++i) {
			VLOG_NONL("\tChecking file path " << paths[i] << "... ")
			std::filesystem::path p(paths[i]);
			if(p.is_relative()) p = opt.input.parent_path() / p;
			if(!std::filesystem::exists(p)) {
				CUBE_ERROR("Input file specifies a nonexistent file!")
			}
			VLOG("Done.")
			VLOG_NONL("\tOpening file stream... ")
			std::ifstream fs(p, std::ios::binary);
			if(!fs.is_open()) {
				CUBE_ERROR("Failed to open face file stream for reading!")
			}
			VLOG("Done.")
			fileBytes[i] = readAll(fs);
			if(fileBytes[i].empty()) throw std::runtime_error(std::string(streamNames[i]) + " face data input stream is invalid!");
		}
	} catch(const std::runtime_error& e) {
		CUBE_ERROR(e.what());
	}
	// the six decodes are independent (inflate / VP8L are sequential per image): one thread each
	const std::string derr = ForEachFace([&](int i) { faces[i] = xjc::DecodeFace(fileBytes[i]); });
	if(!derr.empty()) CUBE_ERROR(derr)

	//Optional GPU conversion to what the engine wants (one fused kernel per face)
	if(opt.toRgb8) {
		VLOG_NONL("Converting faces to 8-bit RGB on the GPU... ")
		try {
			for(auto& f : faces)
				if(f.bitsPerChannel != 8 || f.layout != libcacaoimage::Image::Layout::RGB) f = libcacaoimage::ToRGB8(f);
		} catch(const std::runtime_error& e) {
			CUBE_ERROR(e.what())
		}
		VLOG("Done.")
	}

	//Encode the file
	VLOG_NONL("Generating output... ")
	std::vector<unsigned char> packed;
	std::array<std::vector<unsigned char>, 6> blobs;
	const std::string eerr = ForEachFace([&](int i) {
		// validation of PackedEncode.cpp:20-23
		if(!(faces[i].w > 0 && faces[i].h > 0)) throw std::runtime_error("Cubemap face for packed encoding has invalid dimensions or channel count!");
		if(faces[i].data.empty()) throw std::runtime_error("Cubemap face for packed encoding has empty data buffer!");
		blobs[i] = xjc::EncodeFaceWebP(faces[i]);
	});
	if(!eerr.empty()) CUBE_ERROR(eerr)
	try {
		packed = xjc::WriteContainer(xjc::MakeCubemapContainer(blobs));
	} catch(const std::runtime_error& e) {
		CUBE_ERROR(e.what())
	}
	VLOG("Done.")

	//Write it out
	VLOG_NONL("Writing output file " << opt.output << "... ");
	std::ofstream out(opt.output, std::ios::binary);
	if(!out.is_open()) {
		CUBE_ERROR("Failed to open output file stream for writing!")
	}
	out.write(reinterpret_cast<const char*>(packed.data()), static_cast<std::streamsize>(packed.size()));
	out.flush();
	VLOG("Done.")
	return 0;
}
