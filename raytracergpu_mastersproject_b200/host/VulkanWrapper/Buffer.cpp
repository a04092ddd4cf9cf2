#include "Buffer.hpp"

Buffer::Buffer(Device& device, size_t instanceSize, uint32_t instanceCount)
	: device_(device), instanceSize_(instanceSize), instanceCount_(instanceCount) {
	Device::check(rtb_alloc(device_.context(), getBufferSize(), &dptr_), "failed to allocate buffer");
}

Buffer::~Buffer() { rtb_free(device_.context(), dptr_); }

void Buffer::writeToBuffer(const void* data, size_t size, size_t offset) {
	if (size == SIZE_MAX) size = getBufferSize() - offset;
	if (offset + size > getBufferSize()) throw std::runtime_error("Buffer::writeToBuffer out of range");
	Device::check(rtb_upload(device_.context(), static_cast<char*>(dptr_) + offset, data, size), "failed to write buffer");
}

void Buffer::readFromBuffer(void* out, size_t size, size_t offset) const {
	if (size == SIZE_MAX) size = getBufferSize() - offset;
	if (offset + size > getBufferSize()) throw std::runtime_error("Buffer::readFromBuffer out of range");
	Device::check(rtb_download(device_.context(), out, static_cast<const char*>(dptr_) + offset, size), "failed to read buffer");
}
