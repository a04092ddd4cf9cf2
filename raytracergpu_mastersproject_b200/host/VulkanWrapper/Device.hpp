// Device: the compute device + queue the renderer submits to.  In the reference this class owns the Vulkan
// instance / surface / logical device / queues / command pools (Device.hpp:27-113); here it owns one rtb_ctx
// (CUDA device + stream) of librtb200.so.  Failures throw std::runtime_error like the reference.
#pragma once

#include <cstddef>
#include <stdexcept>
#include <string>

#include "rtb200.h"

class Device {
	rtb_ctx* ctx_ = nullptr;
	int index_ = 0;
public:
	explicit Device(int deviceIndex = 0);
	~Device();
	Device(const Device&) = delete;
	Device& operator=(const Device&) = delete;
	Device(Device&&) = delete;
	Device& operator=(Device&&) = delete;

	rtb_ctx* context() const { return ctx_; }
	rtb_ctx* computeQueue() const { return ctx_; }          // submissions are ordered on the context's stream
	int index() const { return index_; }
	std::string name() const;
	void waitIdle() const;                                  // vkWaitForFences / vkDeviceWaitIdle
	// Device::copyBuffer (Device.hpp:102): host <-> device copies go through Buffer; device-side copy:
	static void check(int rc, const char* what);            // throws std::runtime_error(what + rtb_last_error())
};
