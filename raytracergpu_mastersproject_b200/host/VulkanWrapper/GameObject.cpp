#include "GameObject.hpp"

#include <cassert>

auto GameObject::createGameObject() -> GameObject {
	static GameObjectId nextId = 1;
	return GameObject{ nextId++ };
}

auto GameObject::makePointLight(f32 intensity, f32 radius, glm::vec3 color) -> GameObject {
	GameObject obj = createGameObject();
	obj.color = color;
	obj.transform.scale.x = radius;
	obj.getComponent<PointLightComponent>(obj.addComponent<PointLightComponent>())->lightIntensity = intensity;
	return obj;
}

auto GameObject::setModel(std::shared_ptr<RTModel> m, bool triangular) -> void {
	assert(triangular ? std::holds_alternative<RTModel_Triangles>(*m) : std::holds_alternative<RTModel_Sphere>(*m));
	isTriangleModel = triangular;
	isSphereModel = !triangular;
	model = std::move(m);
}

GameObject::GameObject(const GameObject& o)
	: id{ createGameObject().id }, model{ o.model }, color{ o.color }, transform{ o.transform },
	  isTriangleModel{ o.isTriangleModel }, isSphereModel{ o.isSphereModel } {
	components.reserve(o.components.size());
	for (const auto& c : o.components)
		std::visit([this](const auto& ptr) { components.emplace_back(std::make_unique<std::decay_t<decltype(*ptr)>>(*ptr)); }, c);
}
