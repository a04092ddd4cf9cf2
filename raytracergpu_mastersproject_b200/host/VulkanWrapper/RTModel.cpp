#include "RTModel.hpp"

#include <fstream>
#include <stdexcept>
#include <unordered_map>

#include "ObjLoader.hpp"

template <> struct std::hash<glm::vec3> {
	size_t operator()(const glm::vec3& v) const { size_t s = 0; hashCombine(s, v.x, v.y, v.z); return s; }
};
template <> struct std::hash<glm::vec2> {
	size_t operator()(const glm::vec2& v) const { size_t s = 0; hashCombine(s, v.x, v.y); return s; }
};
template <> struct std::hash<Vertex> {
	size_t operator()(const Vertex& v) const { size_t s = 0; hashCombine(s, v.position, v.color, v.normal, v.uv); return s; }
};

namespace {
std::string g_searchPath =
#ifdef RTB_MODEL_DIR
	RTB_MODEL_DIR;
#else
	".";
#endif

std::string locate(const std::string& filepath) {
	if (std::ifstream(filepath).good()) return filepath;
	const std::string base = filepath.substr(filepath.find_last_of('/') == std::string::npos ? 0 : filepath.find_last_of('/') + 1);
	for (const std::string& cand : { g_searchPath + "/" + filepath, g_searchPath + "/" + base })
		if (std::ifstream(cand).good()) return cand;
	return filepath;
}
}  // namespace

auto setModelSearchPath(const std::string& dir) -> void { g_searchPath = dir; }

auto loadModel(const std::string& filepath, glm::vec3 color) -> std::unique_ptr<RTModel> {
	return loadModel(filepath, SceneTypes::GPU::Material(color, SceneTypes::MaterialType::DIFFUSE));
}

// OBJ -> triangle list: corners are de-duplicated on (position, colour, normal, uv) and every 3 indices make a
// triangle of POSITIONS, in file order (reference: RTModel.cpp:47-108).
auto loadModel(const std::string& filepath, SceneTypes::GPU::Material mat) -> std::unique_ptr<RTModel> {
	obj::Mesh mesh;
	std::string err;
	if (!obj::load(locate(filepath), mesh, err)) throw std::runtime_error(err);

	std::vector<Vertex> vertices;
	std::vector<u32> indices;
	std::unordered_map<Vertex, u32> unique;
	indices.reserve(mesh.indices.size());
	for (const obj::Index& ix : mesh.indices) {
		Vertex v{};
		if (ix.vertex_index >= 0) {
			v.position = { mesh.vertices[3 * ix.vertex_index], mesh.vertices[3 * ix.vertex_index + 1], mesh.vertices[3 * ix.vertex_index + 2] };
			v.color = { mesh.colors[3 * ix.vertex_index], mesh.colors[3 * ix.vertex_index + 1], mesh.colors[3 * ix.vertex_index + 2] };
		}
		if (ix.normal_index >= 0)
			v.normal = { mesh.normals[3 * ix.normal_index], mesh.normals[3 * ix.normal_index + 1], mesh.normals[3 * ix.normal_index + 2] };
		if (ix.texcoord_index >= 0) v.uv = { mesh.texcoords[2 * ix.texcoord_index], mesh.texcoords[2 * ix.texcoord_index + 1] };
		auto [it, fresh] = unique.try_emplace(v, u32(vertices.size()));
		if (fresh) vertices.push_back(v);
		indices.push_back(it->second);
	}
	std::vector<SceneTypes::CPU::Triangle> tris;
	tris.reserve(indices.size() / 3);
	for (size_t i = 0; i + 2 < indices.size(); i += 3)
		tris.push_back({ vertices[indices[i]].position, vertices[indices[i + 1]].position, vertices[indices[i + 2]].position });
	return std::make_unique<RTModel>(RTModel_Triangles(std::move(tris), mat));
}

auto loadModel(f32 radius, SceneTypes::GPU::Material mat) -> std::unique_ptr<RTModel> {
	return std::make_unique<RTModel>(RTModel_Sphere(radius, mat));
}

auto loadModel(std::vector<SceneTypes::CPU::Triangle>&& triangles, SceneTypes::GPU::Material mat) -> std::unique_ptr<RTModel> {
	return std::make_unique<RTModel>(RTModel_Triangles(std::move(triangles), mat));
}
