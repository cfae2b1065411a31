// extract.cpp -- `ce-cubetool extract`: packed cubemap (.xjc) -> face images.
//
// Same contract as the reference subcommand (tools/cubetool/extract.cpp:9-101): positional `input`,
// -A/--all-faces or -f/--face NAME..., -o DIR (created when missing), output files
// <stem>_<face>.png.  One addition: --flip mirrors the rows on the GPU before writing (the
// orientation the OpenGL backend uploads, engine/src/opengl/impls/OpenGLCubemap.cpp:25).
#include <algorithm>
#include <fstream>
#include <iterator>
#include <sstream>
#include <stdexcept>

#include "commands.hpp"
#include "imageio.hpp"
#include "libcacaoimage_ext.hpp"
#include "xjc.hpp"

int RunExtract(const ExtractOptions& optIn) {
	ExtractOptions opt = optIn;
	if(opt.all) {
		for(int i = 0; i < 6; ++i) opt.faces.push_back(i);
	}
	if(opt.faces.empty()) {
		CUBE_ERROR("No faces selected for extraction!")
	}
	std::filesystem::create_directories(opt.outDir);

	//Load the input file
	VLOG_NONL("Opening input file " << opt.input << "... ");
	std::ifstream in(opt.input, std::ios::binary);
	if(!in.is_open()) {
		CUBE_ERROR("Failed to open input file stream for reading!")
	}
	VLOG("Done.")

	//Decode the file
	VLOG_NONL("Decoding input file... ")
	std::array<libcacaoimage::Image, 6> decoded;
	std::array<std::vector<unsigned char>, 6> blobs;
	try {
		const std::vector<unsigned char> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
		const xjc::Container pc = xjc::ReadContainer(file);
		blobs = xjc::SplitCubemapPayload(pc);
	} catch(const std::runtime_error& e) {
		CUBE_ERROR(e.what());
	}
	// all six are decoded, as PackedDecoder::DecodeCubemap does -- a damaged face fails the command even
	// when it was not asked for -- one thread per face
	const std::string derr = ForEachFace([&](int i) { decoded[i] = xjc::DecodeFace(blobs[i]); });
	if(!derr.empty()) CUBE_ERROR(derr)
	VLOG("Done.")

	//Extract the requested faces
	const std::string original = opt.input.filename().stem().string();
	for(int i : opt.faces) {
		VLOG("Opening output stream for " << kFaceNames[i] << ": " << (opt.outDir / (original + "_" + kFaceNames[i] + ".png")) << "... ")
	}
	VLOG("Writing output... ")
	const std::string werr = ForEachFace([&](int i) {
		if(std::find(opt.faces.begin(), opt.faces.end(), i) == opt.faces.end()) return;
		libcacaoimage::Image face = decoded[i];
		if(opt.flip) face = libcacaoimage::Flip(face);
		cubeio::WriteImage(libcacaoimage::FromImage(face), opt.outDir / (original + "_" + kFaceNames[i] + ".png"), "png");
	});
	if(!werr.empty()) CUBE_ERROR(werr)
	return 0;
}
