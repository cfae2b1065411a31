// main.cpp -- ce-cubetool entry point.
//
// Keeps the reference tool's command-line contract (tools/cubetool/main.cpp:15-38): one required
// subcommand, -v/--version, -V/--verbose (silent by default), -h/--help, `-o` for outputs, failures
// as "ERROR: ..." on stderr with exit code 1.  create and extract keep the reference's options
// (create.cpp:9-30, extract.cpp:9-48); `project` is the new, GPU-backed subcommand.
#include <algorithm>
#include <cstring>
#include <filesystem>
#include <string>
#include <vector>

#include "commands.hpp"

#ifndef CACAO_VER
#define CACAO_VER "unknown"
#endif
#ifndef TOOL_VER
#define TOOL_VER "b200-r1"
#endif

namespace {
	void usage(const std::string& prog) {
		std::cout << "Cacao Engine Cubemap Tool \n\n\n"
				  << prog << " [OPTIONS] SUBCOMMAND\n\n\nOPTIONS:\n"
				  << "  -h,     --help              Print this help message and exit \n"
				  << "  -v,     --version           Show version info and exit \n"
				  << "  -V,     --verbose           Enable verbose output from the tool \n\n"
				  << "SUBCOMMANDS:\n"
				  << "  create                      Create a new cubemap \n"
				  << "  extract                     Extract face images from a cubemap \n"
				  << "  project                     Project an equirectangular panorama onto six cube faces (GPU) \n";
	}

	void usageProject(const std::string& prog) {
		std::cout << "Project an equirectangular panorama onto six cube faces (GPU) \n\n\n"
				  << prog << " project [OPTIONS] input\n\n\nPOSITIONALS:\n"
				  << "  input TEXT:FILE REQUIRED    Path to an equirectangular image (PNG, PFM, NPY, PGM/PPM) \n\nOPTIONS:\n"
				  << "  -h,     --help              Print this help message and exit \n"
				  << "  -o TEXT REQUIRED            Directory to place the face images and the .ajc definition in \n"
				  << "  -s,     --face-size UINT    Edge length of each face in pixels (default: input width / 4) \n"
				  << "  -f,     --face TEXT ...     Write only specific faces (left, right, up, down, front, back) \n"
				  << "  -g,     --gpus INT          Number of GPUs to spread face/row-band work units over (default 1) \n"
				  << "          --format TEXT       Face file format: png, pfm, npy (default: png for integer, npy for float samples) \n"
				  << "          --layout TEXT       Convert faces to gray, rgb or rgba \n"
				  << "          --depth TEXT        Convert faces to 8, 16, f16 or f32 samples \n"
				  << "          --flip              Store faces bottom row first \n"
				  << "          --bench             Print a JSON timing line \n";
	}

	void usageCreate(const std::string& prog) {
		std::cout << "Create a new cubemap \n\n\n"
				  << prog << " create [OPTIONS] input\n\n\nPOSITIONALS:\n"
				  << "  input TEXT:FILE REQUIRED    Path to an unpacked cubemap definition file (.ajc) to read as input \n\nOPTIONS:\n"
				  << "  -h,     --help              Print this help message and exit \n"
				  << "  -o TEXT REQUIRED            Output file path \n"
				  << "          --to-rgb8           Convert the faces to 8-bit RGB on the GPU before packing (what the engine does at load) \n";
	}

	void usageExtract(const std::string& prog) {
		std::cout << "Extract face images from a cubemap \n\n\n"
				  << prog << " extract [OPTIONS] input\n\n\nPOSITIONALS:\n"
				  << "  input TEXT:FILE REQUIRED    Path to an packed cubemap file (.xjc) to read as input \n\nOPTIONS:\n"
				  << "  -h,     --help              Print this help message and exit \n"
				  << "  -A,     --all-faces Excludes: --face \n                              Extract all face images from the cubemap \n"
				  << "  -f,     --face TEXT ... Excludes: --all-faces \n                              Extract specific faces from the cubemap (left, right, up, down, front, back) \n"
				  << "  -o TEXT                     Directory to place output files in \n"
				  << "          --flip              Write the faces bottom row first (mirrored on the GPU) \n";
	}

	[[noreturn]] void parseError(const std::string& msg) {
		// CLI11 prints the problem and a hint, and exits non-zero
		std::cerr << msg << "\nRun with --help for more information.\n";
		exit(106);
	}
}

int main(int argc, char* argv[]) {
	const std::string prog = std::filesystem::path(argv[0]).filename().string();
	std::vector<std::string> args(argv + 1, argv + argc);

	// global flags may appear anywhere (the reference app uses fallthrough())
	std::string sub;
	std::vector<std::string> rest;
	for(const std::string& a : args) {
		if(a == "-V" || a == "--verbose") {
			outputLvl = OutputLevel::Verbose;
		} else if(a == "-v" || a == "--version") {
			std::cout << "CubeTool v" << TOOL_VER << "\nFor Cacao Engine v" << CACAO_VER << std::endl;
			return 0;
		} else if(sub.empty() && (a == "-h" || a == "--help")) {
			usage(prog);
			return 0;
		} else if(sub.empty() && !a.empty() && a[0] != '-') {
			sub = a;
		} else {
			rest.push_back(a);
		}
	}
	if(sub.empty()) parseError("A subcommand is required");

	if(sub == "create") {
		CreateOptions opt;
		bool haveInput = false, haveOut = false;
		for(size_t i = 0; i < rest.size(); ++i) {
			const std::string& a = rest[i];
			if(a == "-h" || a == "--help") {
				usageCreate(prog);
				return 0;
			} else if(a == "-o") {
				if(i + 1 >= rest.size()) parseError("-o: 1 required TEXT missing");
				opt.output = rest[++i];
				haveOut = true;
			} else if(a == "--to-rgb8") {
				opt.toRgb8 = true;
			} else if(!a.empty() && a[0] == '-') {
				parseError("The following argument was not expected: " + a);
			} else if(!haveInput) {
				opt.input = std::filesystem::absolute(a);
				haveInput = true;
			} else {
				parseError("The following argument was not expected: " + a);
			}
		}
		if(!haveInput) parseError("input is required");
		if(!std::filesystem::is_regular_file(opt.input)) parseError("input: File does not exist: " + opt.input.string());
		if(!haveOut) parseError("-o is required");
		if(std::filesystem::exists(opt.output) && !std::filesystem::is_regular_file(opt.output)) parseError("-o: The output file must either be a file to overwrite or a nonexistent file!");
		return RunCreate(opt);
	}
	if(sub == "extract") {
		ExtractOptions opt;
		opt.outDir = std::filesystem::current_path();
		bool haveInput = false;
		for(size_t i = 0; i < rest.size(); ++i) {
			const std::string& a = rest[i];
			if(a == "-h" || a == "--help") {
				usageExtract(prog);
				return 0;
			} else if(a == "-o") {
				if(i + 1 >= rest.size()) parseError("-o: 1 required TEXT missing");
				opt.outDir = std::filesystem::absolute(rest[++i]);
			} else if(a == "-A" || a == "--all-faces") {
				opt.all = true;
			} else if(a == "-f" || a == "--face") {
				if(i + 1 >= rest.size()) parseError("--face: At least 1 required");
				// one or more names follow, as with the reference's vector option
				while(i + 1 < rest.size() && !rest[i + 1].empty() && rest[i + 1][0] != '-') {
					const std::string v = rest[i + 1];
					const auto it = std::find_if(kFaceNames.begin(), kFaceNames.end(), [&v](const char* n) { return v == n; });
					if(it == kFaceNames.end()) {
						// the first name must be a face; a later non-face token is the positional input
						if(opt.faces.empty() || haveInput) parseError("--face: Invalid face name!");
						break;
					}
					opt.faces.push_back(static_cast<int>(it - kFaceNames.begin()));
					++i;
				}
			} else if(a == "--flip") {
				opt.flip = true;
			} else if(!a.empty() && a[0] == '-') {
				parseError("The following argument was not expected: " + a);
			} else if(!haveInput) {
				opt.input = a;
				haveInput = true;
			} else {
				parseError("The following argument was not expected: " + a);
			}
		}
		if(!haveInput) parseError("input is required");
		if(!std::filesystem::is_regular_file(opt.input)) parseError("input: File does not exist: " + opt.input.string());
		if(opt.all && !opt.faces.empty()) parseError("--all-faces excludes --face");
		if(std::filesystem::exists(opt.outDir) && !std::filesystem::is_directory(opt.outDir)) parseError("-o: The output directory must either be a directory to write into or a nonexistent directory!");
		return RunExtract(opt);
	}
	if(sub != "project") parseError("The following argument was not expected: " + sub);

	ProjectOptions opt;
	bool haveInput = false, haveOut = false;
	for(size_t i = 0; i < rest.size(); ++i) {
		const std::string& a = rest[i];
		auto value = [&](const char* name) -> std::string {
			if(i + 1 >= rest.size()) parseError(std::string(name) + ": 1 required TEXT missing");
			return rest[++i];
		};
		if(a == "-h" || a == "--help") {
			usageProject(prog);
			return 0;
		} else if(a == "-o") {
			opt.outDir = std::filesystem::absolute(value("-o"));
			haveOut = true;
		} else if(a == "-s" || a == "--face-size") {
			const std::string v = value("--face-size");
			char* end = nullptr;
			const long n = std::strtol(v.c_str(), &end, 10);
			if(*end || n <= 0) parseError("--face-size: value must be a positive integer");
			opt.faceSize = static_cast<unsigned>(n);
		} else if(a == "-g" || a == "--gpus") {
			opt.gpus = std::atoi(value("--gpus").c_str());
			if(opt.gpus < 1) parseError("--gpus: value must be at least 1");
		} else if(a == "-f" || a == "--face") {
			const std::string v = value("--face");
			const auto it = std::find_if(kFaceNames.begin(), kFaceNames.end(), [&v](const char* n) { return v == n; });
			if(it == kFaceNames.end()) parseError("--face: Invalid face name!");
			opt.faces.push_back(static_cast<int>(it - kFaceNames.begin()));
		} else if(a == "--format") {
			opt.format = value("--format");
		} else if(a == "--layout") {
			opt.layout = value("--layout");
		} else if(a == "--depth") {
			opt.depth = value("--depth");
		} else if(a == "--flip") {
			opt.flip = true;
		} else if(a == "--bench") {
			opt.bench = true;
		} else if(!a.empty() && a[0] == '-') {
			parseError("The following argument was not expected: " + a);
		} else if(!haveInput) {
			opt.input = std::filesystem::absolute(a);
			haveInput = true;
		} else {
			parseError("The following argument was not expected: " + a);
		}
	}
	if(!haveInput) parseError("input is required");
	if(!std::filesystem::is_regular_file(opt.input)) parseError("input: File does not exist: " + opt.input.string());
	if(!haveOut) parseError("-o is required");
	if(std::filesystem::exists(opt.outDir) && !std::filesystem::is_directory(opt.outDir)) parseError("-o: The output directory must either be a directory to write into or a nonexistent directory!");
	return RunProject(opt);
}
