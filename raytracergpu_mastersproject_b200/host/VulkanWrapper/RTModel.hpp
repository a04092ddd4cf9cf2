// RTModel = variant<triangle mesh, sphere> + the three loadModel overloads (reference API: RTModel.hpp:56-85).
#pragma once

#include <memory>
#include <string>
#include <variant>
#include <vector>

#include "../glm/glm.hpp"
#include "../utils/PrimitiveTypes.hpp"
#include "SceneTypes.hpp"
#include "utils.hpp"

struct Vertex {
	glm::vec3 position{}, color{}, normal{};
	glm::vec2 uv{};
	bool operator==(const Vertex& o) const { return position == o.position && color == o.color && normal == o.normal && uv == o.uv; }
};

class RTModel_Triangles {
	std::vector<SceneTypes::CPU::Triangle> triangles;
	SceneTypes::GPU::Material material;
public:
	RTModel_Triangles(std::vector<SceneTypes::CPU::Triangle>&& t, SceneTypes::GPU::Material mat) : triangles{ std::move(t) }, material{ mat } {}
	auto getTriangles() const -> const std::vector<SceneTypes::CPU::Triangle>& { return triangles; }
	auto getMaterialType() const -> SceneTypes::GPU::Material { return material; }
};

class RTModel_Sphere {
	glm::vec3 center;
	f32 radius;
	SceneTypes::GPU::Material material;
public:
	RTModel_Sphere(f32 r, SceneTypes::GPU::Material mat) : center{ 0 }, radius{ r }, material{ mat } {}   // centre is the model origin
	auto getCenter() const -> const glm::vec3& { return center; }
	auto getRadius() const -> f32 { return radius; }
	auto getMaterialType() const -> SceneTypes::GPU::Material { return material; }
};

using RTModel = std::variant<RTModel_Triangles, RTModel_Sphere>;

auto loadModel(const std::string& filepath, glm::vec3 color) -> std::unique_ptr<RTModel>;
auto loadModel(const std::string& filepath, SceneTypes::GPU::Material mat) -> std::unique_ptr<RTModel>;
auto loadModel(f32 radius, SceneTypes::GPU::Material mat) -> std::unique_ptr<RTModel>;
// additive: a mesh that is already in memory (synthetic scenes)
auto loadModel(std::vector<SceneTypes::CPU::Triangle>&& triangles, SceneTypes::GPU::Material mat) -> std::unique_ptr<RTModel>;

// Where loadModel(path, ...) looks for relative paths ("models/quad.obj"): the working directory first, then this
// directory.  Defaults to the models/ directory shipped next to the host sources.
auto setModelSearchPath(const std::string& dir) -> void;
