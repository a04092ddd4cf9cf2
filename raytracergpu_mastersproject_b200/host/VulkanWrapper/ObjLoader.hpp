// In-tree Wavefront OBJ / MTL reader standing in for tinyobjloader's LoadObj (reference call site: RTModel.cpp:47-112, which
// passes triangulate = true by default; the library is neither vendored by the reference nor installed here, README.md:8 names
// no version).  It restates the PUBLISHED behaviour of tiny_obj_loader.h (v2.0.0rc10, the release current with the Vulkan SDK
// 1.3.250 the README names; function exportGroupsToShape) for what RTModel.cpp consumes -- per-corner (vertex, normal, texcoord)
// index triples in file order after triangulation:
//   * a 3-corner face is emitted as is;
//   * a QUAD is split along its SHORTER diagonal: |v2 - v0|^2 < |v3 - v1|^2 -> (0,1,2) (0,2,3), otherwise (0,1,3) (1,2,3);
//   * a polygon of 5+ corners is ear-clipped in the two axes of its dominant plane (first non-degenerate corner), ears taken in
//     corner order starting at corner 0, rejecting reflex corners (by the sign of the polygon's area) and ears that contain
//     another corner (pnpoly); the last three corners form the final triangle;
//   * faces with fewer than 3 corners are dropped with a warning; vertex colours default to 1.0; negative indices are relative;
//   * `mtllib` files are parsed into `materials` (newmtl / Ka / Kd / Ks / Ke / Ns / Ni / d / illum) and a missing file is only a
//     warning -- the reference never looks at them (RTModel.cpp:47-112 uses the Material passed to loadModel).
// Parity is pinned by restatement only (un-vendored third party): tests/test_host_scene.py holds quad / n-gon fixtures.
#pragma once

#include <string>
#include <vector>

namespace obj {
struct Index { int vertex_index = -1, normal_index = -1, texcoord_index = -1; };
struct Material {
	std::string name;
	float ambient[3] = { 0, 0, 0 }, diffuse[3] = { 0, 0, 0 }, specular[3] = { 0, 0, 0 }, emission[3] = { 0, 0, 0 };
	float shininess = 1.0f, ior = 1.0f, dissolve = 1.0f;
	int illum = 0;
};
struct Mesh {
	std::vector<float> vertices;   // xyz
	std::vector<float> colors;     // rgb per vertex (1,1,1 when the file has none)
	std::vector<float> normals;
	std::vector<float> texcoords;
	std::vector<Index> indices;    // 3 per triangle
	std::vector<int> materialIds;  // per triangle: index into `materials` of the active usemtl, -1 = none
	std::vector<Material> materials;
	std::string warnings;
};
// returns false and fills `err` on failure; `baseDir` is where mtllib files are looked up ("" = do not read them)
bool load(const std::string& path, Mesh& out, std::string& err);
bool parse(const std::string& text, Mesh& out, std::string& err, const std::string& baseDir = "");
bool parseMtl(const std::string& text, std::vector<Material>& out);
}
