#include "LogisticMap.hpp"
#include "utils/Png.hpp"

#include <fstream>
#include <iostream>
#include <random>

namespace LogisticMapRenderer {

LogisticMap::LogisticMap() : LogisticMap(1920, 1080, Config::Headless::LogisticFrames) {}

LogisticMap::LogisticMap(u32 w, u32 h, u32 nFrames, u32 seed, int deviceIndex) : device(deviceIndex), width(w), height(h), frames(nFrames) {
	std::cout << "physical device: " << device.name() << "\n";
	// createShaderStorageBuffers (LogisticMap.cpp:179-221): r uniform in [0,4), x uniform in [0,1); the reference seeds from
	// time(nullptr), here the seed is explicit
	std::mt19937 rndEngine(seed);
	std::uniform_real_distribution<float> rndDist(0.0f, 1.0f);
	std::vector<Logistic> points(SAMPLE_COUNT);
	for (auto& p : points) {
		p.r = rndDist(rndEngine) * 4.0f;
		p.x = rndDist(rndEngine);
	}
	shaderStorageBuffer = std::make_unique<Buffer>(device, sizeof(Logistic), SAMPLE_COUNT);
	shaderStorageBuffer->writeToBuffer(points.data(), points.size() * sizeof(Logistic));
	computeImage = std::make_unique<Buffer>(device, 4, width * height);     // rgba8 storage image (createComputeImage)
	Device::check(rtb_memset(device.context(), computeImage->getBuffer(), 0, computeImage->getBufferSize()), "failed to create image");
}

LogisticMap::~LogisticMap() { try { device.waitIdle(); } catch (...) {} }

auto LogisticMap::doIteration() -> void {                                // LogisticMap.hpp:150-195
	UniformBufferObject nUbo{};
	nUbo.iteration = f32(iteration);
	nUbo.pixelColor = glm::vec4(1.0f, 1.0f, 1.0f, 1.0f);
	nUbo.width = f32(width);
	nUbo.height = f32(height);
	const float color[4] = { nUbo.pixelColor.x, nUbo.pixelColor.y, nUbo.pixelColor.z, nUbo.pixelColor.w };
	// vkCmdDispatch(SAMPLE_COUNT / KERNEL_SIZE) of logistic.comp (LogisticMap.cpp:384) + fence wait
	Device::check(rtb_logistic_step(device.computeQueue(), shaderStorageBuffer->getBuffer(), SAMPLE_COUNT, computeImage->getBuffer(), width,
	                                height, color), "failed to submit compute command buffer!");
	device.waitIdle();
}

auto LogisticMap::mainLoop() -> void {                                   // LogisticMap.hpp:208-230
	for (; iteration < frames; iteration++) doIteration();
	device.waitIdle();
	lastFrame.resize(size_t(4) * width * height);
	computeImage->readFromBuffer(lastFrame.data(), lastFrame.size());
	std::ofstream f(Config::Headless::OutputImage, std::ios::binary);
	f << "P6\n" << width << " " << height << "\n255\n";
	for (size_t i = 0; i < size_t(width) * height; i++) f.write(reinterpret_cast<const char*>(&lastFrame[4 * i]), 3);
	png::writeRGB(Config::Headless::OutputPng, lastFrame.data(), width, height);
	std::cout << "LogisticMap: " << frames << " iterations of " << SAMPLE_COUNT << " points\n";
}

auto LogisticMap::readPoints() -> std::vector<Logistic> { return shaderStorageBuffer->readAs<Logistic>(SAMPLE_COUNT); }

}  // namespace LogisticMapRenderer
