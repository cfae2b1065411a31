// toolutil.hpp -- output conventions shared by the tools, matching the reference's
// tools/toolutil.hpp: silent by default, -V turns VLOG lines on, errors go to stderr as a bold red
// "ERROR: " prefix followed by the message.
#pragma once

#include <iostream>

enum class OutputLevel { Silent, Normal, Verbose };

inline OutputLevel outputLvl = OutputLevel::Silent;

#define VLOG(...) \
	if(outputLvl == OutputLevel::Verbose) std::cout << __VA_ARGS__ << std::endl;
#define VLOG_NONL(...) \
	if(outputLvl == OutputLevel::Verbose) std::cout << __VA_ARGS__ << std::flush;
