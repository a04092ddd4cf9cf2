// TransformComponent: translation / scale / Tait-Bryan YXZ rotation -> model matrix
// (reference API: VulkanWrapper/components/TransformComponent.hpp:8-32).
#pragma once

#include "../../glm/gtc/matrix_transform.hpp"
#include "../../utils/PrimitiveTypes.hpp"

struct TransformComponent {
	glm::vec3 translation{};
	glm::vec3 scale{ 1.f, 1.f, 1.f };
	glm::vec3 rotation{};   // radians

	TransformComponent() = default;
	TransformComponent(glm::vec3 t, glm::vec3 s, glm::vec3 r) : translation{ t }, scale{ s }, rotation{ r } {}

	auto mat4() const -> glm::mat4;
	auto normalMatrix() const -> glm::mat3;
};
