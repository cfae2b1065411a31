// commands.hpp -- ce-cubetool subcommands.  CLI contract follows the reference tool
// (tools/cubetool/main.cpp:15-38, commands.hpp:9-11,31): subcommand style, -v/--version,
// -V/--verbose, -o, and CUBE_ERROR = red "ERROR: " on stderr + exit(1).
#pragma once

#include <array>
#include <cstdlib>
#include <filesystem>
#include <string>
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

// ce-cubetool project: equirectangular panorama -> six cube faces (+ .ajc for `create`)
int RunProject(const ProjectOptions& opt);
