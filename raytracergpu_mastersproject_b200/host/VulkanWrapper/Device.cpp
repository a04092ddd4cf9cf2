#include "Device.hpp"

void Device::check(int rc, const char* what) {
	if (rc != 0) throw std::runtime_error(std::string(what) + ": " + rtb_last_error());
}

Device::Device(int deviceIndex) : index_(deviceIndex) {
	check(rtb_ctx_create(deviceIndex, nullptr, &ctx_), "failed to create compute device");
}

Device::~Device() { rtb_ctx_destroy(ctx_); }

std::string Device::name() const {
	char buf[256];
	check(rtb_device_name(ctx_, buf, sizeof(buf)), "failed to query device name");
	return buf;
}

void Device::waitIdle() const { check(rtb_sync(ctx_), "failed to wait for the compute queue"); }
