#include "ObjLoader.hpp"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace obj {
namespace {
// OBJ indices are 1-based; negative values count back from the end of the list read so far
int resolve(long idx, size_t count) {
	if (idx > 0) return int(idx - 1);
	if (idx < 0) return int(long(count) + idx);
	return -1;
}
bool parseCorner(const char* tok, const Mesh& m, Index& out) {
	// v, v/vt, v//vn, v/vt/vn
	char* end = nullptr;
	long v = std::strtol(tok, &end, 10);
	if (end == tok) return false;
	out.vertex_index = resolve(v, m.vertices.size() / 3);
	if (*end == '/') {
		const char* p = end + 1;
		if (*p != '/') {
			long vt = std::strtol(p, &end, 10);
			if (end != p) out.texcoord_index = resolve(vt, m.texcoords.size() / 2);
		} else end = const_cast<char*>(p);
		if (*end == '/') {
			p = end + 1;
			long vn = std::strtol(p, &end, 10);
			if (end != p) out.normal_index = resolve(vn, m.normals.size() / 3);
		}
	}
	return out.vertex_index >= 0;
}
}  // namespace

bool parse(const std::string& text, Mesh& out, std::string& err) {
	std::istringstream in(text);
	std::string line;
	size_t lineNo = 0;
	while (std::getline(in, line)) {
		lineNo++;
		if (!line.empty() && line.back() == '\r') line.pop_back();
		std::istringstream ls(line);
		std::string tag;
		if (!(ls >> tag) || tag[0] == '#') continue;
		if (tag == "v") {
			float x = 0, y = 0, z = 0, r = 1, g = 1, b = 1;
			ls >> x >> y >> z;
			if (!(ls >> r >> g >> b)) { r = g = b = 1.0f; }
			out.vertices.insert(out.vertices.end(), { x, y, z });
			out.colors.insert(out.colors.end(), { r, g, b });
		} else if (tag == "vn") {
			float x = 0, y = 0, z = 0; ls >> x >> y >> z;
			out.normals.insert(out.normals.end(), { x, y, z });
		} else if (tag == "vt") {
			float u = 0, v = 0; ls >> u >> v;
			out.texcoords.insert(out.texcoords.end(), { u, v });
		} else if (tag == "f") {
			std::vector<Index> corners;
			std::string tok;
			while (ls >> tok) {
				Index ix;
				if (!parseCorner(tok.c_str(), out, ix)) { err = "bad face corner '" + tok + "' at line " + std::to_string(lineNo); return false; }
				if (size_t(ix.vertex_index) >= out.vertices.size() / 3) { err = "vertex index out of range at line " + std::to_string(lineNo); return false; }
				corners.push_back(ix);
			}
			if (corners.size() < 3) { err = "face with fewer than 3 corners at line " + std::to_string(lineNo); return false; }
			for (size_t k = 1; k + 1 < corners.size(); k++) {      // triangle fan (0, k, k+1)
				out.indices.push_back(corners[0]);
				out.indices.push_back(corners[k]);
				out.indices.push_back(corners[k + 1]);
			}
		}
		// o, g, s, usemtl, mtllib: no effect on the triangle list
	}
	return true;
}

bool load(const std::string& path, Mesh& out, std::string& err) {
	std::ifstream f(path, std::ios::binary);
	if (!f) { err = "cannot open " + path; return false; }
	std::stringstream ss;
	ss << f.rdbuf();
	return parse(ss.str(), out, err);
}
}  // namespace obj
