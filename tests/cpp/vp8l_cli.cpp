// vp8l_cli.cpp -- test helper around tools/cubetool/vp8l.cpp (built and driven by tests/test_xjc.py):
//   vp8l_cli enc <raw-in> <w> <h> <channels> <webp-out>
//   vp8l_cli dec <webp-in> <raw-out>        prints "w h channels"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

#include "vp8l.hpp"

static std::vector<uint8_t> slurp(const char* path) {
	std::ifstream f(path, std::ios::binary);
	if(!f) throw std::runtime_error(std::string("cannot open ") + path);
	return std::vector<uint8_t>(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
	try {
		const std::string mode = argc > 1 ? argv[1] : "";
		if(mode == "enc" && argc == 7) {
			const std::vector<uint8_t> raw = slurp(argv[2]);
			const unsigned w = std::atoi(argv[3]), h = std::atoi(argv[4]);
			const int ch = std::atoi(argv[5]);
			if(raw.size() != static_cast<size_t>(w) * h * ch) throw std::runtime_error("raw size does not match w*h*channels");
			const std::vector<uint8_t> out = vp8l::Encode(raw.data(), w, h, ch);
			std::ofstream(argv[6], std::ios::binary).write(reinterpret_cast<const char*>(out.data()), out.size());
			return 0;
		}
		if(mode == "dec" && argc == 4) {
			const std::vector<uint8_t> file = slurp(argv[2]);
			const vp8l::Decoded d = vp8l::Decode(file.data(), file.size());
			std::ofstream(argv[3], std::ios::binary).write(reinterpret_cast<const char*>(d.pixels.data()), d.pixels.size());
			std::cout << d.w << " " << d.h << " " << d.channels << std::endl;
			return 0;
		}
		std::cerr << "usage: vp8l_cli enc raw w h ch out.webp | dec in.webp out.raw" << std::endl;
		return 2;
	} catch(const std::exception& e) {
		std::cerr << e.what() << std::endl;
		return 1;
	}
}
