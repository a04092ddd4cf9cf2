// Scene builders.  The first five have the reference's names and signatures (Scenes.hpp:9-17) and build the same
// scenes (Scenes.cpp:20-398).  The synthetic generators below them are additive: they produce the benchmark
// configurations C2..C5 of BASELINE.md deterministically (no std::random_device), with primitives emitted in the
// reference's own Morton order so that its original-order-leaf LBVH (D8) is a usable tree.
#pragma once

#include <memory>
#include <string>

#include "VulkanWrapper/RaytraceScene.hpp"
#include "utils/PrimitiveTypes.hpp"

auto randomSpheres(std::unique_ptr<RaytraceScene>& scene) -> void;
auto cornellMixedScene(std::unique_ptr<RaytraceScene>& scene) -> void;
auto cornellBoxScene(std::unique_ptr<RaytraceScene>& scene) -> void;
auto simpleScene(std::unique_ptr<RaytraceScene>& scene) -> void;
auto complexScene(std::unique_ptr<RaytraceScene>& scene) -> void;

namespace SyntheticScenes {
	// C2: closed displaced-sphere mesh of 2*segments^2 triangles (segments = 660 -> 871 200) inside the simpleScene room
	auto meshRoomScene(std::unique_ptr<RaytraceScene>& scene, u32 segments = 660, u32 seed = 1) -> void;
	// C3: `count` spheres in [25,525]^3, radii [1,4], 70/15/15 % diffuse/metal/dielectric, + light and floor quads
	auto sphereFieldScene(std::unique_ptr<RaytraceScene>& scene, u32 count = 100000, u32 seed = 2) -> void;
	// C4 / C5: nx x nz cell fBm height field over [0,550]^2 (2 triangles per cell), light quad above, dummy sphere;
	// `dielectricPercent` of the Morton-contiguous triangle chunks are tagged DIELECTRIC (C5: 80)
	auto heightFieldScene(std::unique_ptr<RaytraceScene>& scene, u32 nx = 3162, u32 nz = 1581, u32 seed = 3, u32 dielectricPercent = 0) -> void;
	// by name: "complexScene", "simpleScene", "cornellBoxScene", "cornellMixedScene", "randomSpheres",
	// "meshRoom[:segments[:seed]]", "sphereField[:count[:seed]]", "heightField[:nx:nz[:seed[:dielectricPercent]]]"
	auto buildByName(std::unique_ptr<RaytraceScene>& scene, const std::string& spec) -> void;
}
