// Minimal PNG encoder for the headless "present" (resolve-to-PNG).  The reference presents to a swapchain
// (B8G8R8A8_UNORM, SwapChain.cpp:393) and writes no image; a headless build has to put the frame somewhere.  8-bit RGB, no
// filtering, zlib stream of stored (uncompressed) deflate blocks -- valid PNG, no dependency.
#pragma once

#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace png {
namespace detail {
inline uint32_t crc32(const uint8_t* p, size_t n, uint32_t crc = 0) {
	static uint32_t table[256];
	static bool init = false;
	if (!init) {
		for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
		init = true;
	}
	crc = ~crc;
	for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
	return ~crc;
}
inline void be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x)); }
inline void chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
	be32(out, uint32_t(data.size()));
	std::vector<uint8_t> body(type, type + 4);
	body.insert(body.end(), data.begin(), data.end());
	out.insert(out.end(), body.begin(), body.end());
	be32(out, crc32(body.data(), body.size()));
}
}  // namespace detail

// rgba: width * height * 4 bytes (alpha dropped); rows top to bottom
inline std::vector<uint8_t> encodeRGB(const uint8_t* rgba, uint32_t width, uint32_t height) {
	using namespace detail;
	std::vector<uint8_t> raw;                                   // filter byte 0 + RGB per scanline
	raw.reserve(size_t(height) * (1 + 3 * size_t(width)));
	for (uint32_t y = 0; y < height; y++) {
		raw.push_back(0);
		for (uint32_t x = 0; x < width; x++) { const uint8_t* p = rgba + 4 * (size_t(y) * width + x); raw.push_back(p[0]); raw.push_back(p[1]); raw.push_back(p[2]); }
	}
	std::vector<uint8_t> z = { 0x78, 0x01 };                    // zlib header: deflate, 32 K window, no preset dictionary
	uint32_t a = 1, b = 0;                                      // Adler-32 of the raw data
	for (size_t pos = 0; pos < raw.size() || pos == 0;) {
		const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
		const bool last = pos + n >= raw.size();
		z.push_back(last ? 1 : 0);                              // BFINAL, BTYPE = 00 (stored)
		z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8)); z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
		z.insert(z.end(), raw.begin() + long(pos), raw.begin() + long(pos + n));
		for (size_t i = pos; i < pos + n; i++) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
		pos += n;
		if (last) break;
	}
	be32(z, (b << 16) | a);
	std::vector<uint8_t> out = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
	std::vector<uint8_t> ihdr;
	be32(ihdr, width); be32(ihdr, height);
	ihdr.insert(ihdr.end(), { 8, 2, 0, 0, 0 });                 // 8 bits, colour type 2 (RGB), deflate, adaptive filtering, no interlace
	chunk(out, "IHDR", ihdr);
	chunk(out, "IDAT", z);
	chunk(out, "IEND", {});
	return out;
}
inline void writeRGB(const std::string& path, const uint8_t* rgba, uint32_t width, uint32_t height) {
	const std::vector<uint8_t> bytes = encodeRGB(rgba, width, height);
	std::ofstream f(path, std::ios::binary);
	if (!f) throw std::runtime_error("failed to open " + path);
	f.write(reinterpret_cast<const char*>(bytes.data()), long(bytes.size()));
}
}  // namespace png
