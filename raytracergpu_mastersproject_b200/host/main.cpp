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
