// Buffer: RAII device-local storage of instanceSize x instanceCount bytes (reference API: Buffer.hpp:7-58).
// The staging-buffer dance of the reference (host-visible Buffer + map + writeToBuffer + Device::copyBuffer,
// RaytraceScene.hpp:139-178) collapses to writeToBuffer(): a host->device copy on the context's stream.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "Device.hpp"

class Buffer {
	Device& device_;
	void* dptr_ = nullptr;
	size_t instanceSize_;
	uint32_t instanceCount_;
public:
	Buffer(Device& device, size_t instanceSize, uint32_t instanceCount);
	~Buffer();
	Buffer(const Buffer&) = delete;
	Buffer& operator=(const Buffer&) = delete;

	int map() { return 0; }                         // kept for source compatibility; device memory is not host-mapped
	void unmap() {}
	int flush() { return 0; }
	void writeToBuffer(const void* data, size_t size = SIZE_MAX, size_t offset = 0);
	void readFromBuffer(void* out, size_t size = SIZE_MAX, size_t offset = 0) const;   // DEBUGgetDeployedBufferAs
	template <typename T> std::vector<T> readAs(size_t count) const {
		std::vector<T> v(count);
		readFromBuffer(v.data(), count * sizeof(T));
		return v;
	}
	void* getBuffer() const { return dptr_; }        // the VkBuffer handle's stand-in: a device pointer
	uint32_t getInstanceCount() const { return instanceCount_; }
	size_t getInstanceSize() const { return instanceSize_; }
	size_t getBufferSize() const { return instanceSize_ * instanceCount_; }
};
