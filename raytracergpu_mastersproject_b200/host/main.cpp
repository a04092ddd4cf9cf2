// Entry point, same shape as the reference's main.cpp:8-21: pick the program at compile time, construct, mainLoop().
// The BVH program is the hot path; the non-BVH program (Config::Programs::Raytracer) and the LogisticMap demo are the
// "next" rows N2 / N4.
#include "Config.hpp"
#include "LogisticMap.hpp"
#include "Raytracer.hpp"
#include "RaytracerBVH.hpp"

#include <cstdlib>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

namespace {
// additive: rtb200_main --gpus N <scene spec> [width height] -- one host thread and one RaytracerBVHRenderer::Raytracer per GPU,
// the frame shared by interleaved bands and assembled by rtb_gather_tiles (NCCL over NVLink); rank 0 writes the frame
int runMultiGpu(int n, const std::string& scene, u32 w, u32 h) {
	int have = 0;
	Device::check(rtb_device_count(&have), "failed to count devices");
	if (n < 1 || n > have) throw std::runtime_error("--gpus " + std::to_string(n) + ": " + std::to_string(have) + " device(s) present");
	unsigned char id[RTB_COMM_ID_BYTES];
	Device::check(rtb_comm_unique_id(id), "failed to create the communicator id");
	std::vector<std::thread> threads;
	std::vector<std::string> errors(static_cast<size_t>(n), std::string());
	for (int r = 0; r < n; r++)
		threads.emplace_back([&, r] {
			try {
				RaytracerBVHRenderer::Raytracer comp{ w, h, scene, r, RaytracerBVHRenderer::MultiGpu{ r, n, id } };
				comp.mainLoop();
			} catch (const std::exception& e) { errors[size_t(r)] = e.what(); }
		});
	for (auto& t : threads) t.join();
	for (int r = 0; r < n; r++)
		if (!errors[size_t(r)].empty()) throw std::runtime_error("rank " + std::to_string(r) + ": " + errors[size_t(r)]);
	return 0;
}
}  // namespace

int main(int argc, char** argv) {
	try {
		if constexpr (Config::CurrentProgram == Config::Programs::RaytracerBVH) {
			if (argc > 1 && std::string(argv[1]) == "--logistic") {   // additive: run the Config::Programs::LogisticMap host
				const u32 w = argc > 4 ? u32(std::atoi(argv[2])) : 1920, h = argc > 4 ? u32(std::atoi(argv[3])) : 1080;
				const u32 frames = argc > 4 ? u32(std::atoi(argv[4])) : Config::Headless::LogisticFrames;
				LogisticMapRenderer::LogisticMap comp{ w, h, frames };
				comp.mainLoop();
			} else if (argc > 1 && std::string(argv[1]) == "--non-bvh") {   // additive: run the Config::Programs::Raytracer host instead
				const u32 w = argc > 4 ? u32(std::atoi(argv[3])) : Config::Headless::Width;
				const u32 h = argc > 4 ? u32(std::atoi(argv[4])) : Config::Headless::Height;
				RaytracerRenderer::Raytracer comp{ w, h, argc > 2 ? argv[2] : "complexScene" };
				comp.mainLoop();
			} else if (argc > 3 && std::string(argv[1]) == "--gpus") {
				const u32 w = argc > 5 ? u32(std::atoi(argv[4])) : Config::Headless::Width;
				const u32 h = argc > 5 ? u32(std::atoi(argv[5])) : Config::Headless::Height;
				return runMultiGpu(std::atoi(argv[2]), argv[3], w, h);
			} else if (argc > 1) {   // additive: rtb200_main <scene spec> [width height]
				const u32 w = argc > 3 ? u32(std::atoi(argv[2])) : Config::Headless::Width;
				const u32 h = argc > 3 ? u32(std::atoi(argv[3])) : Config::Headless::Height;
				RaytracerBVHRenderer::Raytracer comp{ w, h, argv[1] };
				comp.mainLoop();
			} else {
				RaytracerBVHRenderer::Raytracer comp{};
				comp.mainLoop();
			}
		} else if constexpr (Config::CurrentProgram == Config::Programs::Raytracer) {
			RaytracerRenderer::Raytracer comp{};
			comp.mainLoop();
		} else {
			LogisticMapRenderer::LogisticMap comp{};
			comp.mainLoop();
		}
	} catch (const std::exception& e) {
		std::cerr << "error: " << e.what() << "\n";
		return 1;
	}
	return 0;
}
