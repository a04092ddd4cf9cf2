#include "Raytracer.hpp"
#include "utils/Png.hpp"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "Scenes.hpp"

namespace RaytracerRenderer {

Raytracer::Raytracer() : Raytracer(Config::Headless::Width, Config::Headless::Height, "complexScene") {}

Raytracer::Raytracer(u32 w, u32 h, const std::string& sceneName, int deviceIndex)
	: device(deviceIndex), width(w), height(h),
	  gen(Config::Headless::RandomState ? Config::Headless::RandomState
	                                    : static_cast<u32>(std::chrono::system_clock::now().time_since_epoch().count())) {
	std::cout << "physical device: " << device.name() << "\n";
	scene = std::make_unique<RaytraceScene>(device);                      // createScene, Raytracer.cpp:256-300
	SyntheticScenes::buildByName(scene, sceneName);
	computeImage = std::make_unique<Buffer>(device, 4 * sizeof(f32), width * height);
	presentImage = std::make_unique<Buffer>(device, 4, width * height);
}

Raytracer::~Raytracer() { try { device.waitIdle(); } catch (...) {} }

auto Raytracer::doIteration(f32) -> void {                               // Raytracer.hpp:208-352
	std::cout << "iteration: " << iteration << "\n";
	rtb_ctx* q = device.computeQueue();
	scene->updateScene();

	RaytracingUniformBufferObject rUbo{};                                 // Raytracer.hpp:288-297
	rUbo.camPos = glm::vec3(275.0f, 275.0f, -800.0f);
	rUbo.camLookAt = glm::vec3(275.0f, 275.0f, 0.0f);
	rUbo.camUpDir = glm::vec3(0.0f, 1.0f, 0.0f);
	rUbo.verticalFOV = scene->getCamera().getVerticalFOV();
	rUbo.numTriangles = scene->getTriangleCount();
	rUbo.numSpheres = scene->getSphereCount();
	rUbo.numMaterials = scene->getMaterialCount();
	rUbo.numLights = u32(scratchSize);
	rUbo.maxRayTraceDepth = scene->getMaxRaytraceDepth();
	rUbo.randomState = u32(gen());
	rtb_ubo ubo;
	std::memcpy(&ubo, &rUbo, sizeof(ubo));

	const auto t0 = std::chrono::high_resolution_clock::now();
	// the single compute command buffer (recordComputeCommandBuffer, Raytracer.cpp:394-538): K1, clear, K8 x raysPerPixel
	Device::check(rtb_model_to_world(q, &ubo, scene->getModelBuffer()->getBuffer(), scene->getTriangleBuffer()->getBuffer(),
	                                 scene->getSphereBuffer()->getBuffer()), "failed to submit compute command buffer!");
	Device::check(rtb_clear_image(q, computeImage->getBuffer(), width, height), "failed to clear the accumulation image");
	Device::check(rtb_bind_trace_buffers(q, &ubo, scene->getTriangleBuffer()->getBuffer(), scene->getSphereBuffer()->getBuffer(),
	                                     scene->getMaterialBuffer()->getBuffer(), nullptr), "failed to bind the raytrace set");
	rtb_trace_args args{};
	args.imageWidth = width; args.imageHeight = height; args.localRows = height;
	args.bandRows = height; args.bandFirst = 0; args.bandStep = 1;
	args.sampleSkip = 0; args.sampleCount = scene->getRaysPerPixel();
	args.flags = RTB_TRACE_LINEAR_SCAN;
	Device::check(rtb_raytrace(q, &ubo, computeImage->getBuffer(), &args), "failed to submit compute command buffer!");
	device.waitIdle();
	const f32 ms = std::chrono::duration<f32, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();

	Device::check(rtb_resolve_rgba8(q, computeImage->getBuffer(), width, height, scene->getRaysPerPixel(), presentImage->getBuffer()),
	              "failed to submit draw command buffer!");
	lastFrame.resize(size_t(4) * width * height);
	presentImage->readFromBuffer(lastFrame.data(), lastFrame.size());
	std::printf("TIMINGS:\n\tTotal Raytracing Time: %.0fus\n", 1e3 * ms);
}

auto Raytracer::mainLoop() -> void {                                     // Raytracer.hpp:363-421
	for (; iteration < Config::Headless::Frames; iteration++) {
		std::cout << "RaysPerPixel: " << scene->getRaysPerPixel() << " Depth: " << scene->getMaxRaytraceDepth() << std::endl;
		doIteration(0.0f);
	}
	device.waitIdle();
	if (!lastFrame.empty()) {
		std::ofstream f(Config::Headless::OutputImage, std::ios::binary);
		f << "P6\n" << width << " " << height << "\n255\n";
		for (size_t i = 0; i < size_t(width) * height; i++) f.write(reinterpret_cast<const char*>(&lastFrame[4 * i]), 3);
		png::writeRGB(Config::Headless::OutputPng, lastFrame.data(), width, height);
	}
}

}  // namespace RaytracerRenderer
