#include "RaytracerBVH.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "Scenes.hpp"
#include "utils/Png.hpp"

namespace RaytracerBVHRenderer {

namespace {
void writePPM(const std::string& path, const std::vector<u8>& rgba, u32 w, u32 h) {
	std::ofstream f(path, std::ios::binary);
	if (!f) throw std::runtime_error("failed to open " + path);
	f << "P6\n" << w << " " << h << "\n255\n";
	for (size_t i = 0; i < size_t(w) * h; i++) f.write(reinterpret_cast<const char*>(&rgba[4 * i]), 3);
}
f32 msSince(std::chrono::high_resolution_clock::time_point t0) {
	return std::chrono::duration<f32, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
}
}  // namespace

Raytracer::Raytracer() : Raytracer(Config::Headless::Width, Config::Headless::Height, "complexScene") {}

Raytracer::Raytracer(u32 w, u32 h, const std::string& sceneName, int deviceIndex, MultiGpu m)
	: device(deviceIndex), width(w), height(h), multi(m), localRows(h),
	  gen(Config::Headless::RandomState ? Config::Headless::RandomState
	                                    : static_cast<u32>(std::chrono::system_clock::now().time_since_epoch().count())) {
	if (multi.rank == 0) std::cout << "physical device: " << device.name() << "\n";           // Device.cpp:137
	if (multi.ranks > 1) {
		if (!Config::Headless::RandomState) throw std::runtime_error("multi-GPU rendering needs a fixed Config::Headless::RandomState (every rank must draw the same seeds)");
		Device::check(rtb_comm_init_rank(device.context(), multi.ranks, multi.rank, multi.commId), "failed to create the communicator");
		const u32 band = Config::Headless::BandRows, bands = (height + band - 1) / band;
		localRows = ((bands + u32(multi.ranks) - 1) / u32(multi.ranks)) * band;
	}
	createScene(sceneName);
}

Raytracer::~Raytracer() { try { device.waitIdle(); } catch (...) {} }

// createScene + the build / image buffers (reference: RaytracerBVH.cpp:479-550, 399-455)
auto Raytracer::createScene(const std::string& sceneName) -> void {
	scene = std::make_unique<RaytraceScene>(device);
	SyntheticScenes::buildByName(scene, sceneName);
	const u32 n = scene->getTriangleCount() + scene->getSphereCount();
	if (n == 0) throw std::runtime_error("failed to create scene: no primitives");
	enclosingAABBBuffer = std::make_unique<Buffer>(device, sizeof(EnclosingAABBBufferObject), 1);
	mortonPrimitiveBuffer1 = std::make_unique<Buffer>(device, sizeof(SceneTypes::GPU::MortonPrimitive), n);
	mortonPrimitiveBuffer2 = std::make_unique<Buffer>(device, sizeof(SceneTypes::GPU::MortonPrimitive), n);
	HLBVHNodesBuffer = std::make_unique<Buffer>(device, sizeof(SceneTypes::GPU::BVHNode), 2 * n - 1);
	HLBVHConstructionInfoBuffer = std::make_unique<Buffer>(device, sizeof(rtb_construction_info), 2 * n - 1);
	computeImage = std::make_unique<Buffer>(device, 4 * sizeof(f32), width * localRows);   // R32G32B32A32_SFLOAT (this rank's bands)
	presentImage = std::make_unique<Buffer>(device, 4, width * height);                 // B8G8R8A8_UNORM stand-in (RGBA8)
}

// one frame (reference: doIteration, RaytracerBVH.hpp:206-496)
auto Raytracer::doIteration(f32) -> void {
	if (multi.rank == 0) std::cout << "iteration: " << iteration << "\n";
	rtb_ctx* q = device.computeQueue();
	auto t0 = std::chrono::high_resolution_clock::now();
	scene->updateScene();                                                 // re-flatten + re-upload model-space arrays
	scene->getCamera().updateCameraForFrame(0.0f, f32(width) / f32(height));
	device.waitIdle();
	lastTimings.updateSceneMs = msSince(t0);

	RaytracingUniformBufferObject rUbo{};                                 // RaytracerBVH.hpp:340-352
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
	static_assert(sizeof(ubo) == sizeof(rUbo));
	std::memcpy(&ubo, &rUbo, sizeof(ubo));

	// S1: clear + K1..K6 (recordComputeS1CommandBuffer, RaytracerBVH.cpp:734-997), then the fence wait
	t0 = std::chrono::high_resolution_clock::now();
	Device::check(rtb_clear_image(q, computeImage->getBuffer(), width, localRows), "failed to clear the accumulation image");
	Device::check(rtb_build_bvh(q, &ubo, scene->getModelBuffer()->getBuffer(), scene->getTriangleBuffer()->getBuffer(),
	                            scene->getSphereBuffer()->getBuffer(), scene->getMaterialBuffer()->getBuffer(),
	                            enclosingAABBBuffer->getBuffer(), mortonPrimitiveBuffer1->getBuffer(), mortonPrimitiveBuffer2->getBuffer(),
	                            HLBVHNodesBuffer->getBuffer(), HLBVHConstructionInfoBuffer->getBuffer(), 0),
	              "failed to submit compute command buffer!");
	device.waitIdle();
	lastTimings.buildMs = msSince(t0);

	// S2: raysPerPixel samples (recordComputeS2CommandBuffer, RaytracerBVH.cpp:998-1050), then the fence wait
	t0 = std::chrono::high_resolution_clock::now();
	rtb_trace_args args{};
	args.imageWidth = width; args.imageHeight = height; args.localRows = localRows;
	args.bandRows = height; args.bandFirst = 0; args.bandStep = 1;
	if (multi.ranks > 1) { args.bandRows = Config::Headless::BandRows; args.bandFirst = u32(multi.rank); args.bandStep = u32(multi.ranks); }
	args.sampleSkip = 0; args.sampleCount = scene->getRaysPerPixel();
	Device::check(rtb_raytrace(q, &ubo, computeImage->getBuffer(), &args), "failed to submit compute command buffer!");
	device.waitIdle();
	lastTimings.traceMs = msSince(t0);

	// "present": the fullscreen fragment pass (SingleTriangleFullScreen.frag:13-21) into an RGBA8 host image
	t0 = std::chrono::high_resolution_clock::now();
	if (multi.ranks > 1)      // the bands of all ranks, resolved and re-assembled on every rank (NCCL over NVLink, in-stream)
		Device::check(rtb_gather_tiles(q, computeImage->getBuffer(), width, height, Config::Headless::BandRows, nullptr, scene->getRaysPerPixel(),
		                               presentImage->getBuffer()), "failed to gather the frame");
	else
		Device::check(rtb_resolve_rgba8(q, computeImage->getBuffer(), width, height, scene->getRaysPerPixel(), presentImage->getBuffer()),
		              "failed to submit draw command buffer!");
	lastFrame.resize(size_t(4) * width * height);
	presentImage->readFromBuffer(lastFrame.data(), lastFrame.size());
	lastTimings.resolveMs = msSince(t0);

	if (multi.rank == 0) std::printf("TIMINGS:\n\tupdateSceneTime: %.0fus, Total BVH Build Time: %.0fus, Total Raytracing Time: %.0fus, resolve: %.0fus\n",
	            1e3 * lastTimings.updateSceneMs, 1e3 * (lastTimings.updateSceneMs + lastTimings.buildMs), 1e3 * lastTimings.traceMs,
	            1e3 * lastTimings.resolveMs);
}

// reference: mainLoop, RaytracerBVH.hpp:511-571 -- "until the window closes" becomes Config::Headless::Frames
auto Raytracer::mainLoop() -> void {
	auto currentTime = std::chrono::high_resolution_clock::now();
	if constexpr (Config::RunRayPerPixelIncreasingDemo) scene->setRaysPerPixel(Config::RayPerPixelIncreasingDemoConfig::startRaysPerPixel);
	for (;;) {
		if constexpr (!Config::RunRayPerPixelIncreasingDemo) { if (iteration >= Config::Headless::Frames) break; }
		auto newTime = std::chrono::high_resolution_clock::now();
		auto frameTime = std::chrono::duration_cast<std::chrono::microseconds>(newTime - currentTime);
		currentTime = newTime;
		if (multi.rank == 0)
			std::cout << "Frame Time(us): " << frameTime.count() << " RaysPerPixel: " << scene->getRaysPerPixel()
			          << " Depth: " << scene->getMaxRaytraceDepth() << std::endl;
		doIteration(f32(frameTime.count()));
		if constexpr (Config::RunRayPerPixelIncreasingDemo) {               // the rays-per-pixel sweep -> runtimes.csv
			namespace D = Config::RayPerPixelIncreasingDemoConfig;
			if (iteration != 0) {
				const u32 index = (iteration - 1) / D::runsBeforeIncrease;
				if (scene->getRaysPerPixel() > D::maxRaysPerPixel) break;
				if (times.size() == index) times.push_back({ frameTime }); else times[index].push_back(frameTime);
				if (iteration % D::runsBeforeIncrease == 0) scene->setRaysPerPixel(scene->getRaysPerPixel() + D::increaseAmount);
			}
		}
		iteration++;
	}
	device.waitIdle();
	if (!lastFrame.empty() && multi.rank == 0) {
		writePPM(Config::Headless::OutputImage, lastFrame, width, height);
		png::writeRGB(Config::Headless::OutputPng, lastFrame.data(), width, height);
	}
	if (multi.rank != 0) return;
	if constexpr (Config::RunRayPerPixelIncreasingDemo) {
		std::ofstream out("runtimes.csv", std::ios::out | std::ios::trunc);
		for (size_t i = 0; i < times.size(); i++) {
			std::chrono::microseconds sum{ 0 };
			for (auto t : times[i]) sum += t;
			out << (i + 1) << ", " << (sum / Config::RayPerPixelIncreasingDemoConfig::runsBeforeIncrease).count() << ",\n";
		}
	}
}

auto Raytracer::readAccumulationImage() -> std::vector<f32> { return computeImage->readAs<f32>(size_t(4) * width * height); }
auto Raytracer::readNodes() -> std::vector<SceneTypes::GPU::BVHNode> {
	return HLBVHNodesBuffer->readAs<SceneTypes::GPU::BVHNode>(HLBVHNodesBuffer->getInstanceCount());
}

}  // namespace RaytracerBVHRenderer
