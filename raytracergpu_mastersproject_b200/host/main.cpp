// Entry point, same shape as the reference's main.cpp:8-21: pick the program at compile time, construct, mainLoop().
// Only the BVH program is on this build's path; the other two Config::Programs values are out of scope.
#include "Config.hpp"
#include "RaytracerBVH.hpp"

#include <cstdlib>
#include <iostream>

int main(int argc, char** argv) {
	try {
		if constexpr (Config::CurrentProgram == Config::Programs::RaytracerBVH) {
			if (argc > 1) {   // additive: rtb200_main <scene spec> [width height]
				const u32 w = argc > 3 ? u32(std::atoi(argv[2])) : Config::Headless::Width;
				const u32 h = argc > 3 ? u32(std::atoi(argv[3])) : Config::Headless::Height;
				RaytracerBVHRenderer::Raytracer comp{ w, h, argv[1] };
				comp.mainLoop();
			} else {
				RaytracerBVHRenderer::Raytracer comp{};
				comp.mainLoop();
			}
		} else {
			std::cerr << "this build only contains Config::Programs::RaytracerBVH\n";
			return 2;
		}
	} catch (const std::exception& e) {
		std::cerr << "error: " << e.what() << "\n";
		return 1;
	}
	return 0;
}
