// GameObject: id + transform + shared RTModel (+ optional components) (reference API: GameObject.hpp:52-104).
#pragma once

#include <memory>
#include <stdexcept>
#include <type_traits>
#include <unordered_map>
#include <variant>
#include <vector>

#include "../utils/PrimitiveTypes.hpp"
#include "RTModel.hpp"
#include "components/PointLightComponent.hpp"
#include "components/TransformComponent.hpp"

template <typename C>
concept ComponentType = std::is_same_v<C, TransformComponent> || std::is_same_v<C, PointLightComponent>;

using ComponentVariantType = std::variant<std::unique_ptr<TransformComponent>, std::unique_ptr<PointLightComponent>>;
using ComponentsVector = std::vector<ComponentVariantType>;
using GameObjectId = u32;

struct GameObject {
	using Map = std::unordered_map<GameObjectId, GameObject>;

protected:
	GameObjectId id;
	std::shared_ptr<RTModel> model{};

public:
	static auto createGameObject() -> GameObject;
	static auto makePointLight(f32 intensity = 10.0f, f32 radius = 0.1f, glm::vec3 color = glm::vec3(1.0f)) -> GameObject;

	template <ComponentType C> static auto getComponent(GameObject& obj, size_t index) -> C* { return obj.getComponent<C>(index); }
	template <ComponentType C> static auto addComponent(GameObject& obj) -> size_t { return obj.addComponent<C>(); }

	auto setModel(std::shared_ptr<RTModel> model, bool triangular) -> void;
	auto getModel() const -> std::shared_ptr<RTModel> { return model; }

	template <ComponentType C> auto getComponent(size_t index) const -> C* {
		if (index >= components.size()) throw std::out_of_range("Invalid index for component!");
		auto* p = std::get_if<std::unique_ptr<C>>(&components[index]);
		return p ? p->get() : nullptr;
	}
	template <ComponentType C> auto addComponent() -> size_t {
		components.emplace_back(std::make_unique<C>());
		return components.size() - 1;
	}
	template <ComponentType C> auto getComponent() const -> C* {
		for (auto& c : components)
			if (auto* p = std::get_if<std::unique_ptr<C>>(&c)) return p->get();
		return nullptr;
	}
	template <ComponentType C> auto getComponents() const -> std::vector<C*> {
		std::vector<C*> out;
		for (auto& c : components)
			if (auto* p = std::get_if<std::unique_ptr<C>>(&c)) out.push_back(p->get());
		return out;
	}

	GameObject(const GameObject&);               // deep-copies components, takes a fresh id
	GameObject& operator=(const GameObject&) = delete;
	GameObject(GameObject&&) = default;
	GameObject& operator=(GameObject&&) = default;

	auto getId() const -> GameObjectId { return id; }

	glm::vec3 color{};
	TransformComponent transform{};
	ComponentsVector components;
	bool isTriangleModel = false;
	bool isSphereModel = false;

protected:
	explicit GameObject(GameObjectId objId) : id{ objId } {}
};
