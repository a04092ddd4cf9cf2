#include "TransformComponent.hpp"

// The model matrix is T * Ry * Rx * Rz * S in column-major storage (reference:
// TransformComponent.cpp:9-45).  It is ON the hot path's input side: K1 multiplies every vertex by it, so the
// binary32 expression order below is part of the parity contract (tests/test_host_scene.py restates it).
namespace {
struct Euler { float cy, sy, cx, sx, cz, sz; };
Euler eulerOf(const glm::vec3& r) {
	return { glm::cos(r.y), glm::sin(r.y), glm::cos(r.x), glm::sin(r.x), glm::cos(r.z), glm::sin(r.z) };
}
// rotation columns of Ry*Rx*Rz; each entry keeps the left-to-right product order of the reference expressions
glm::vec3 rotCol(const Euler& e, int col) {
	switch (col) {
	case 0: return { e.cy * e.cz + e.sy * e.sx * e.sz, e.cx * e.sz, e.cy * e.sx * e.sz - e.cz * e.sy };
	case 1: return { e.cz * e.sy * e.sx - e.cy * e.sz, e.cx * e.cz, e.cy * e.cz * e.sx + e.sy * e.sz };
	default: return { e.cx * e.sy, -e.sx, e.cy * e.cx };
	}
}
}  // namespace

auto TransformComponent::mat4() const -> glm::mat4 {
	const Euler e = eulerOf(rotation);
	glm::mat4 m;
	for (int col = 0; col < 3; col++) {
		const glm::vec3 r = rotCol(e, col);
		const float s = scale[col];
		m[col] = glm::vec4(s * r.x, s * r.y, s * r.z, 0.0f);
	}
	m[3] = glm::vec4(translation.x, translation.y, translation.z, 1.0f);
	return m;
}

auto TransformComponent::normalMatrix() const -> glm::mat3 {
	const Euler e = eulerOf(rotation);
	const glm::vec3 inv = 1.0f / scale;
	glm::mat3 m;
	for (int col = 0; col < 3; col++) {
		const glm::vec3 r = rotCol(e, col);
		m[col] = glm::vec3(inv[col] * r.x, inv[col] * r.y, inv[col] * r.z);
	}
	return m;
}
