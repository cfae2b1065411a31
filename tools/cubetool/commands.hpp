// commands.hpp -- ce-cubetool subcommands.  CLI contract follows the reference tool
// (tools/cubetool/main.cpp:15-38, commands.hpp:9-11,31): subcommand style, -v/--version,
// -V/--verbose, -o, and CUBE_ERROR = red "ERROR: " on stderr + exit(1).
#pragma once

#include <array>
#include <cstdlib>
#include <filesystem>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "toolutil.hpp"

#define CUBE_ERROR(...)                                                         \
	{                                                                           \
		std::cerr << "\x1b[0m\x1b[1;91mERROR: \x1b[0m" << __VA_ARGS__ << std::endl; \
		exit(1);                                                                \
	}

// face slot names, in slot order +X,-X,+Y,-Y,+Z,-Z (reference commands.hpp:31)
inline const std::array<const char*, 6> kFaceNames = {"right", "left", "up", "down", "front", "back"};
// keys of the unpacked cubemap definition (.ajc) in the same slot order (UnpackedDecode.cpp:29-40)
inline const std::array<const char*, 6> kAjcKeys = {"right", "left", "top", "bottom", "front", "back"};

// The six faces are independent and their codecs (zlib, VP8L) are sequential host code: run fn(0..5)
// on six threads.  Returns the first failure in face order ("" when all went well) so the caller can
// report it the usual way -- CUBE_ERROR must not be raised from a worker thread.
inline std::string ForEachFace(const std::function<void(int)>& fn) {
	std::array<std::string, 6> errors;
	std::vector<std::thread> pool;
	for(int i = 0; i < 6; ++i)
		pool.emplace_back([&errors, &fn, i] {
			try {
				fn(i);
			} catch(const std::exception& e) {
				errors[i] = e.what()[0] ? e.what() : "unknown error";
			}
		});
	for(auto& t : pool) t.join();
	for(const auto& e : errors)
		if(!e.empty()) return e;
	return "";
}

struct ProjectOptions {
	std::filesystem::path input;
	std::filesystem::path outDir;
	unsigned faceSize = 0;    // 0 = width / 4
	int gpus = 1;             // devices 0..gpus-1
	std::string format;       // png | pfm | npy ; empty = by sample type
	std::string layout;       // gray | rgb | rgba ; empty = keep
	std::string depth;        // 8 | 16 | f16 | f32 ; empty = keep
	std::vector<int> faces;   // empty = all six
	bool flip = false;        // bottom row first (for GL-style consumers, Flip.cpp's purpose)
	bool bench = false;       // print a JSON timing line
};

struct CreateOptions {
	std::filesystem::path input;  // .ajc, absolute
	std::filesystem::path output; // .xjc
	bool toRgb8 = false;          // GPU: 16 -> 8 bit + -> RGB per face before packing (Cubemap.cpp:28-31)
};

struct ExtractOptions {
	std::filesystem::path input;  // .xjc
	std::filesystem::path outDir; // absolute; default = current directory
	bool all = false;
	std::vector<int> faces;
	bool flip = false; // GPU: bottom row first
};

// ce-cubetool project: equirectangular panorama -> six cube faces (+ .ajc for `create`)
int RunProject(const ProjectOptions& opt);
// ce-cubetool create: .ajc + six face files -> packed .xjc (reference tools/cubetool/create.cpp)
int RunCreate(const CreateOptions& opt);
// ce-cubetool extract: packed .xjc -> <stem>_<face>.png (reference tools/cubetool/extract.cpp)
int RunExtract(const ExtractOptions& opt);
