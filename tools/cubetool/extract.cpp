// extract.cpp -- `ce-cubetool extract`: packed cubemap (.xjc) -> face images.
//
// Same contract as the reference subcommand (tools/cubetool/extract.cpp:9-101): positional `input`,
// -A/--all-faces or -f/--face NAME..., -o DIR (created when missing), output files
// <stem>_<face>.png.  One addition: --flip mirrors the rows on the GPU before writing (the
// orientation the OpenGL backend uploads, engine/src/opengl/impls/OpenGLCubemap.cpp:25).
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
	try {
		const std::vector<unsigned char> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
		const xjc::Container pc = xjc::ReadContainer(file);
		const std::array<std::vector<unsigned char>, 6> blobs = xjc::SplitCubemapPayload(pc);
		for(int i = 0; i < 6; ++i) decoded[i] = xjc::DecodeFace(blobs[i]);
	} catch(const std::runtime_error& e) {
		CUBE_ERROR(e.what());
	}
	VLOG("Done.")

	//Extract the requested faces
	const std::string original = opt.input.filename().stem().string();
	for(int i : opt.faces) {
		std::stringstream fname;
		fname << original << "_" << kFaceNames[i] << ".png";
		const std::filesystem::path outfile = opt.outDir / fname.str();
		VLOG("Opening output stream for " << kFaceNames[i] << ": " << outfile << "... ")
		VLOG("Writing output... ")
		try {
			libcacaoimage::Image face = decoded[i];
			if(opt.flip) face = libcacaoimage::Flip(face);
			cubeio::WriteImage(libcacaoimage::FromImage(face), outfile, "png");
		} catch(const std::runtime_error& e) {
			CUBE_ERROR(e.what())
		}
	}
	return 0;
}
