// LogisticMapRenderer::LogisticMap -- the host of the third program, Config::Programs::LogisticMap (reference API:
// LogisticMap.hpp:25-38,150-230).  It is a bifurcation-diagram demo, unrelated to the path tracer (it is NOT the path
// tracer's RNG): 512 Ki points (x, r) are advanced one logistic-map step per frame and plotted into a 1920x1080 rgba8 image
// that is never cleared, so the attractor emerges over the frames.  Headless: the image is written as a PPM at the end.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "Config.hpp"
#include "VulkanWrapper/Buffer.hpp"
#include "VulkanWrapper/Device.hpp"
#include "glm/glm.hpp"
#include "utils/PrimitiveTypes.hpp"

namespace LogisticMapRenderer {
	constexpr const u32 SAMPLE_COUNT = 1024 * 512;
	constexpr const u32 KERNEL_SIZE = 1024;

	struct UniformBufferObject {
		glm::vec4 pixelColor;
		float iteration;
		float width;
		float height;
	};
	struct Logistic { float x; float r; };   // f(x; r) = x r (1 - x)

	class LogisticMap {
		Device device;
		u32 width, height, frames;
		std::unique_ptr<Buffer> shaderStorageBuffer, computeImage;
		u32 iteration = 0;
		std::vector<u8> lastFrame;

		auto doIteration() -> void;

	public:
		LogisticMap();                                                    // 1920 x 1080 (LogisticMap.cpp:9)
		LogisticMap(u32 width, u32 height, u32 frames, u32 seed = 1, int deviceIndex = Config::Headless::DeviceIndex);   // additive
		~LogisticMap();
		auto mainLoop() -> void;
		auto frameRGBA8() const -> const std::vector<u8>& { return lastFrame; }
		auto readPoints() -> std::vector<Logistic>;
	};
}
