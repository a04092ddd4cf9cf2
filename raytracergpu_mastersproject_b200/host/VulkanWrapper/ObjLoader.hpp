// In-tree Wavefront OBJ reader standing in for tinyobjloader's LoadObj (reference: RTModel.cpp:47-112; the
// library is not vendored there and not installed here).  Produces what RTModel.cpp consumes: per-corner
// (vertex, normal, texcoord) index triples after fan triangulation, in file order.
#pragma once

#include <string>
#include <vector>

namespace obj {
struct Index { int vertex_index = -1, normal_index = -1, texcoord_index = -1; };
struct Mesh {
	std::vector<float> vertices;   // xyz
	std::vector<float> colors;     // rgb per vertex (1,1,1 when the file has none)
	std::vector<float> normals;
	std::vector<float> texcoords;
	std::vector<Index> indices;    // 3 per triangle
};
// returns false and fills `err` on failure
bool load(const std::string& path, Mesh& out, std::string& err);
bool parse(const std::string& text, Mesh& out, std::string& err);
}
