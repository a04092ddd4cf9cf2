// RaytracerRenderer::Raytracer -- the host of the NON-BVH program, Config::Programs::Raytracer (reference API:
// Raytracer.hpp:30-44,361-423).  Per frame: updateScene -> UBO -> one compute submission = ModelSpaceToWorldSpace, clear,
// raysPerPixel dispatches of raytrace.comp (Raytracer.cpp:394-538) -> present.  raytrace.comp's sceneHit scans every
// primitive, so this program is only practical for small scenes; it exists for API completeness and as an independent
// check of the BVH program's closest hits.
#pragma once

#include <memory>
#include <random>
#include <string>
#include <vector>

#include "Config.hpp"
#include "VulkanWrapper/Buffer.hpp"
#include "VulkanWrapper/Device.hpp"
#include "VulkanWrapper/RaytraceScene.hpp"
#include "utils/PrimitiveTypes.hpp"

namespace RaytracerRenderer {
	struct RaytracingUniformBufferObject {
		alignas(16) glm::vec3 camPos;
		alignas(16) glm::vec3 camLookAt;
		alignas(16) glm::vec3 camUpDir;
		alignas(16) f32 verticalFOV;
		u32 numTriangles;
		u32 numSpheres;
		u32 numMaterials;
		u32 numLights;
		u32 maxRayTraceDepth;
		u32 randomState;
	};
	static_assert(sizeof(RaytracingUniformBufferObject) == sizeof(rtb_ubo));
	struct FragmentUniformBufferObject { u32 raysPerPixel; };

	class Raytracer {
		Device device;
		u32 width, height;
		std::unique_ptr<RaytraceScene> scene;
		std::unique_ptr<Buffer> computeImage, presentImage;
		u32 iteration = 0;
		std::mt19937 gen;
		const f32 scratchSize = 20;
		std::vector<u8> lastFrame;

		auto doIteration(f32 frameTime) -> void;

	public:
		Raytracer();
		Raytracer(u32 width, u32 height, const std::string& sceneName, int deviceIndex = Config::Headless::DeviceIndex);   // additive
		~Raytracer();
		auto mainLoop() -> void;
		auto frameRGBA8() const -> const std::vector<u8>& { return lastFrame; }
	};
}
