// project.cpp -- `ce-cubetool project`: equirectangular panorama -> six cube faces on the GPU.
//
// New subcommand beside the reference's create/extract (tools/cubetool/main.cpp:31-33); it emits
// exactly what `create` consumes: six face files named <stem>_{right,left,up,down,front,back}
// (the extractor's naming, extract.cpp:83-85) plus an unpacked cubemap definition <stem>.ajc whose
// keys are right,left,top,bottom,front,back (UnpackedDecode.cpp:29-40).
#include <algorithm>
#include <chrono>
#include <fstream>
#include <stdexcept>

#include "commands.hpp"
#include "imageio.hpp"
#include "libcacaoimage_ext.hpp"

using namespace libcacaoimage;

int RunProject(const ProjectOptions& opt) {
	//Load the panorama
	VLOG_NONL("Opening input file " << opt.input << "... ");
	ImageEx pano;
	try {
		pano = cubeio::ReadImage(opt.input);
	} catch(const std::runtime_error& e) {
		CUBE_ERROR(e.what())
	}
	VLOG("Done. (" << pano.w << "x" << pano.h << ", " << static_cast<int>(pano.layout) << " channels)")
	if(pano.w != 2 * pano.h) {
		VLOG("Note: input is not 2:1; it is treated as a full 360x180 degree panorama anyway.")
	}
	const unsigned N = opt.faceSize ? opt.faceSize : pano.w / 4;
	if(N == 0) CUBE_ERROR("Face size must be at least 1 pixel!")

	//What the faces should look like in the end
	Image::Layout layout = pano.layout;
	SampleFormat fmt = pano.format;
	if(opt.layout == "gray") layout = Image::Layout::Grayscale;
	else if(opt.layout == "rgb") layout = Image::Layout::RGB;
	else if(opt.layout == "rgba") layout = Image::Layout::RGBA;
	else if(!opt.layout.empty()) CUBE_ERROR("Invalid layout! (gray, rgb, rgba)")
	if(opt.depth == "8") fmt = SampleFormat::U8;
	else if(opt.depth == "16") fmt = SampleFormat::U16;
	else if(opt.depth == "f16") fmt = SampleFormat::F16;
	else if(opt.depth == "f32") fmt = SampleFormat::F32;
	else if(!opt.depth.empty()) CUBE_ERROR("Invalid depth! (8, 16, f16, f32)")
	const bool post = layout != pano.layout || fmt != pano.format || opt.flip;

	//Project (+ convert + flip).  One GPU with post-processing: the panorama goes up once, the faces
	//stay on the device through projection, fused layout/depth conversion and row mirror, and come
	//back once.  Several GPUs (or nothing to post-process): the host-memory call, which shards the
	//(face, row band) units over the devices and pipelines upload / resampling / readback.
	std::array<ImageEx, 6> faces;
	auto run = [&]() {
		if(opt.gpus == 1 && post) {
			const DeviceImage dpano = DeviceImage::Upload(pano);
			//Float panoramas are resampled and converted in one pass where that is defined (Spec B.4)
			const std::array<DeviceImage, 6> dfaces = dpano.EquirectToCubemap(N, layout, fmt, opt.flip);
			for(int i = 0; i < 6; ++i) {
				DeviceImage f = dfaces[i];
				faces[i] = f.Download();
			}
			return;
		}
		// --gpus k deals the work units to k host threads; with fewer than k devices in the box the shares go
		// round robin over the devices there are (same result, the shares of one device run one after the other)
		const int have = std::max(1, cacao_img_device_count());
		std::vector<int> devices;
		for(int d = 0; d < opt.gpus; ++d) devices.push_back(d % have);
		if(devices.size() == 1) devices.clear(); // current device
		faces = EquirectToCubemap(pano, N, layout, fmt, devices, opt.flip);
	};
	VLOG_NONL("Projecting onto 6 faces of " << N << "x" << N << " on " << opt.gpus << " GPU(s)... ")
	const auto t0 = std::chrono::steady_clock::now();
	try {
		run();
	} catch(const std::runtime_error& e) {
		CUBE_ERROR(e.what())
	}
	const auto t1 = std::chrono::steady_clock::now();
	VLOG("Done.")
	double warm = 0.0;
	if(opt.bench) { // second pass: CUDA contexts, staging buffers and pinned rings exist now
		const auto w0 = std::chrono::steady_clock::now();
		try {
			run();
		} catch(const std::runtime_error& e) {
			CUBE_ERROR(e.what())
		}
		warm = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
	}

	//Write faces + definition
	std::filesystem::create_directories(opt.outDir);
	const std::string stem = opt.input.filename().stem().string();
	const std::string format = opt.format.empty() ? cubeio::DefaultFormat(faces[0]) : opt.format;
	std::array<std::string, 6> names;
	for(int i = 0; i < 6; ++i) names[i] = stem + "_" + kFaceNames[i] + "." + format;
	VLOG_NONL("Writing face files to " << opt.outDir << "... ")
	const std::string werr = ForEachFace([&](int i) {
		if(!opt.faces.empty() && std::find(opt.faces.begin(), opt.faces.end(), i) == opt.faces.end()) return;
		cubeio::WriteImage(faces[i], opt.outDir / names[i], format);
	});
	if(!werr.empty()) CUBE_ERROR(werr)
	VLOG("Done.")
	if(opt.faces.empty()) {
		const std::filesystem::path ajc = opt.outDir / (stem + ".ajc");
		std::ofstream def(ajc);
		if(!def.is_open()) CUBE_ERROR("Failed to open output file stream for writing!")
		for(int i = 0; i < 6; ++i) def << kAjcKeys[i] << ": " << names[i] << "\n";
		VLOG("Wrote cubemap definition " << ajc << ".")
	}
	if(opt.bench) {
		const double s = std::chrono::duration<double>(t1 - t0).count();
		std::cout << "{\"subcommand\": \"project\", \"faces\": 6, \"face_size\": " << N << ", \"seconds_first_call\": " << s << ", \"seconds_warm\": " << warm << ", \"output_mpixel_per_s_warm\": " << 6.0 * N * N / warm / 1e6 << ", \"gpus\": " << opt.gpus << "}" << std::endl;
	}
	return 0;
}
