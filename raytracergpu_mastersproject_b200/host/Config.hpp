// Compile-time configuration (same names and meaning as the reference's Config.hpp:5-24).
// Additive fields for the headless build are grouped in Config::Headless; the reference has no resolution
// setting (the render size is the 800x800 window of RaytracerBVH.cpp:8) nor a frame limit (it runs until the
// window closes).
#pragma once

#include "utils/PrimitiveTypes.hpp"

namespace Config {
	enum struct Programs { LogisticMap, Raytracer, RaytracerBVH };

	constexpr const Programs CurrentProgram = Programs::RaytracerBVH;

	constexpr const bool ShowBufferDebug = 0;
	constexpr const bool Fake1SecondDelay = 0;

#ifdef RTB_CONFIG_RPP_DEMO                       // build flag of the `rtb200_rppdemo` target (the reference edits this line by hand)
	constexpr const bool RunRayPerPixelIncreasingDemo = 1;
#else
	constexpr const bool RunRayPerPixelIncreasingDemo = 0;
#endif
	namespace RayPerPixelIncreasingDemoConfig {
		constexpr const u32 runsBeforeIncrease = 4;
		constexpr const u32 startRaysPerPixel = 100;
		constexpr const u32 maxRaysPerPixel = 200;
		constexpr const u32 increaseAmount = 5;
	};

	namespace Headless {                        // additive (no counterpart in the reference)
		constexpr const u32 Width = 800;        // window{800, 800}
		constexpr const u32 Height = 800;
		constexpr const u32 Frames = 1;         // mainLoop() iterations before the "window closes"
		constexpr const u32 LogisticFrames = 200; // same, for the LogisticMap program (one map step per frame)
		constexpr const u32 RandomState = 12345; // deterministic stand-in for the clock-seeded mt19937 (D9); 0 = clock
		constexpr const int DeviceIndex = 0;
		constexpr const char* OutputImage = "frame.ppm";
		constexpr const char* OutputPng = "frame.png";     // the same frame as PNG (utils/Png.hpp)
		constexpr const u32 BandRows = 8;                  // multi-GPU tile mode: rows per interleaved band (rtb_trace_args)
	};
};
