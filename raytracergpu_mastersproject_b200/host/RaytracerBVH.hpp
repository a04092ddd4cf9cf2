// RaytracerBVHRenderer::Raytracer -- the host of the BVH program (reference API: RaytracerBVH.hpp:31-49,509-573):
// constructor builds the scene and the device buffers, mainLoop() runs frames, each frame = doIteration() =
// updateScene -> fill UBO -> S1 (BVH build) -> wait -> S2 (trace) -> wait -> resolve.  Headless: the swapchain
// present becomes an RGBA8 image in host memory (written to Config::Headless::OutputImage).
// Every Vulkan call of the reference's host is replaced by one C-ABI call of librtb200.so.
#pragma once

#include <chrono>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "Config.hpp"
#include "VulkanWrapper/Buffer.hpp"
#include "VulkanWrapper/Device.hpp"
#include "VulkanWrapper/RaytraceScene.hpp"
#include "VulkanWrapper/SceneTypes.hpp"
#include "utils/PrimitiveTypes.hpp"

namespace RaytracerBVHRenderer {
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
	static_assert(sizeof(RaytracingUniformBufferObject) == sizeof(rtb_ubo) && sizeof(rtb_ubo) == 80);
	struct EnclosingAABBBufferObject {
		alignas(16) glm::vec3 min;
		alignas(16) glm::vec3 max;
	};
	struct FragmentUniformBufferObject { u32 raysPerPixel; };

	struct FrameTimings { f32 updateSceneMs = 0, buildMs = 0, traceMs = 0, resolveMs = 0; };

	// additive: this renderer is rank `rank` of `ranks` renderers (one per GPU, one host thread or process each) that share a
	// frame by interleaved bands of Config::Headless::BandRows rows; commId = the 128 bytes of rtb_comm_unique_id().
	// The reference renders on the first device only (VulkanWrapper/Device.cpp:115-138).
	struct MultiGpu { int rank = 0; int ranks = 1; const void* commId = nullptr; };

	class Raytracer {
		Device device;
		u32 width, height;
		MultiGpu multi;
		u32 localRows;                            // rows of the accumulation image this rank renders (= height on one GPU)

		std::unique_ptr<RaytraceScene> scene;
		std::unique_ptr<Buffer> enclosingAABBBuffer, mortonPrimitiveBuffer1, mortonPrimitiveBuffer2;
		std::unique_ptr<Buffer> HLBVHNodesBuffer, HLBVHConstructionInfoBuffer;
		std::unique_ptr<Buffer> computeImage, presentImage;

		u32 iteration = 0;
		std::mt19937 gen;
		const f32 scratchSize = 20;
		std::vector<u8> lastFrame;                // RGBA8, row 0 = top
		FrameTimings lastTimings;
		std::vector<std::vector<std::chrono::microseconds>> times;

		auto createScene(const std::string& sceneName) -> void;
		auto doIteration(f32 frameTime) -> void;

	public:
		Raytracer();                                                      // 800 x 800, complexScene (RaytracerBVH.cpp:8,511)
		Raytracer(u32 width, u32 height, const std::string& sceneName, int deviceIndex = Config::Headless::DeviceIndex,
		          MultiGpu multi = {});                                   // additive
		~Raytracer();
		auto mainLoop() -> void;

		// additive accessors for the headless driver / tests
		auto getScene() -> RaytraceScene& { return *scene; }
		auto getDevice() -> Device& { return device; }
		auto frameRGBA8() const -> const std::vector<u8>& { return lastFrame; }
		auto timings() const -> const FrameTimings& { return lastTimings; }
		auto readAccumulationImage() -> std::vector<f32>;
		auto readNodes() -> std::vector<SceneTypes::GPU::BVHNode>;
		auto renderFrame() -> void { doIteration(0.0f); iteration++; }
	};
}
