#include "ObjLoader.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
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

// the crossing-number test tiny_obj_loader.h uses for "does this ear contain another corner"
int pnpoly(int nvert, const float* vertx, const float* verty, float testx, float testy) {
	int c = 0;
	for (int i = 0, j = nvert - 1; i < nvert; j = i++) {
		if (((verty[i] > testy) != (verty[j] > testy)) &&
		    (testx < (vertx[j] - vertx[i]) * (testy - verty[i]) / (verty[j] - verty[i]) + vertx[i]))
			c = !c;
	}
	return c;
}

void emit(Mesh& m, const Index& a, const Index& b, const Index& c, int mat) {
	m.indices.push_back(a); m.indices.push_back(b); m.indices.push_back(c);
	m.materialIds.push_back(mat);
}

// exportGroupsToShape's triangulation of one face (see the header for the rules)
void triangulate(Mesh& m, const std::vector<Index>& face, int mat) {
	const std::vector<float>& v = m.vertices;
	size_t npolys = face.size();
	if (npolys == 3) { emit(m, face[0], face[1], face[2], mat); return; }
	if (npolys == 4) {
		const float* p0 = &v[3 * size_t(face[0].vertex_index)]; const float* p1 = &v[3 * size_t(face[1].vertex_index)];
		const float* p2 = &v[3 * size_t(face[2].vertex_index)]; const float* p3 = &v[3 * size_t(face[3].vertex_index)];
		const float e02x = p2[0] - p0[0], e02y = p2[1] - p0[1], e02z = p2[2] - p0[2];
		const float e13x = p3[0] - p1[0], e13y = p3[1] - p1[1], e13z = p3[2] - p1[2];
		const float sqr02 = e02x * e02x + e02y * e02y + e02z * e02z;
		const float sqr13 = e13x * e13x + e13y * e13y + e13z * e13z;
		if (sqr02 < sqr13) { emit(m, face[0], face[1], face[2], mat); emit(m, face[0], face[2], face[3], mat); }
		else { emit(m, face[0], face[1], face[3], mat); emit(m, face[1], face[2], face[3], mat); }
		return;
	}
	// 5+ corners: the two axes of the dominant plane, from the first corner that is not degenerate
	size_t axes[2] = { 1, 2 };
	for (size_t k = 0; k < npolys; ++k) {
		const float* a = &v[3 * size_t(face[(k + 0) % npolys].vertex_index)];
		const float* b = &v[3 * size_t(face[(k + 1) % npolys].vertex_index)];
		const float* c = &v[3 * size_t(face[(k + 2) % npolys].vertex_index)];
		const float e0x = b[0] - a[0], e0y = b[1] - a[1], e0z = b[2] - a[2];
		const float e1x = c[0] - b[0], e1y = c[1] - b[1], e1z = c[2] - b[2];
		const float cx = std::fabs(e0y * e1z - e0z * e1y), cy = std::fabs(e0z * e1x - e0x * e1z), cz = std::fabs(e0x * e1y - e0y * e1x);
		const float epsilon = std::numeric_limits<float>::epsilon();
		if (cx > epsilon || cy > epsilon || cz > epsilon) {
			if (cx > cy && cx > cz) {
			} else {
				axes[0] = 0;
				if (cz > cx && cz > cy) axes[1] = 1;
			}
			break;
		}
	}
	float area = 0;
	for (size_t k = 0; k < npolys; ++k) {
		const float* a = &v[3 * size_t(face[(k + 0) % npolys].vertex_index)];
		const float* b = &v[3 * size_t(face[(k + 1) % npolys].vertex_index)];
		area += (a[axes[0]] * b[axes[1]] - a[axes[1]] * b[axes[0]]) * 0.5f;
	}
	std::vector<Index> rest = face;
	size_t guess = 0;
	size_t remainingIterations = face.size();
	size_t previousRemaining = rest.size();
	while (rest.size() > 3 && remainingIterations > 0) {
		npolys = rest.size();
		if (guess >= npolys) guess -= npolys;
		if (previousRemaining != npolys) { previousRemaining = npolys; remainingIterations = npolys; }
		else remainingIterations--;
		Index ind[3];
		float vx[3], vy[3];
		for (size_t k = 0; k < 3; k++) {
			ind[k] = rest[(guess + k) % npolys];
			const size_t vi = size_t(ind[k].vertex_index);
			vx[k] = v[vi * 3 + axes[0]]; vy[k] = v[vi * 3 + axes[1]];
		}
		const float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0], e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
		const float cross = e0x * e1y - e0y * e1x;
		if (cross * area < 0.0f) { guess += 1; continue; }          // a reflex corner is no ear
		bool overlap = false;
		for (size_t other = 3; other < npolys; ++other) {
			const size_t ovi = size_t(rest[(guess + other) % npolys].vertex_index);
			if (pnpoly(3, vx, vy, v[ovi * 3 + axes[0]], v[ovi * 3 + axes[1]])) { overlap = true; break; }
		}
		if (overlap) { guess += 1; continue; }
		emit(m, ind[0], ind[1], ind[2], mat);                       // an ear: clip its middle corner
		size_t removed = (guess + 1) % npolys;
		while (removed + 1 < npolys) { rest[removed] = rest[removed + 1]; removed += 1; }
		rest.pop_back();
	}
	if (rest.size() == 3) emit(m, rest[0], rest[1], rest[2], mat);
}

void read3(std::istringstream& ls, float* d) { ls >> d[0] >> d[1] >> d[2]; }
}  // namespace

bool parseMtl(const std::string& text, std::vector<Material>& out) {
	std::istringstream in(text);
	std::string line;
	Material cur; bool open = false;
	while (std::getline(in, line)) {
		if (!line.empty() && line.back() == '\r') line.pop_back();
		std::istringstream ls(line);
		std::string tag;
		if (!(ls >> tag) || tag[0] == '#') continue;
		if (tag == "newmtl") {
			if (open) out.push_back(cur);
			cur = Material(); open = true;
			ls >> cur.name;
		} else if (tag == "Ka") read3(ls, cur.ambient);
		else if (tag == "Kd") read3(ls, cur.diffuse);
		else if (tag == "Ks") read3(ls, cur.specular);
		else if (tag == "Ke") read3(ls, cur.emission);
		else if (tag == "Ns") ls >> cur.shininess;
		else if (tag == "Ni") ls >> cur.ior;
		else if (tag == "d") ls >> cur.dissolve;
		else if (tag == "Tr") { float tr = 0; ls >> tr; cur.dissolve = 1.0f - tr; }
		else if (tag == "illum") ls >> cur.illum;
	}
	if (open) out.push_back(cur);
	return true;
}

bool parse(const std::string& text, Mesh& out, std::string& err, const std::string& baseDir) {
	std::istringstream in(text);
	std::string line;
	size_t lineNo = 0;
	int activeMaterial = -1;
	while (std::getline(in, line)) {
		lineNo++;
		if (!line.empty() && line.back() == '\r') line.pop_back();
		while (!line.empty() && line.back() == '\\') {              // line continuation
			std::string next;
			line.pop_back();
			if (!std::getline(in, next)) break;
			lineNo++;
			if (!next.empty() && next.back() == '\r') next.pop_back();
			line += " " + next;
		}
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
			if (corners.size() < 3) {                                 // tinyobjloader: "Degenerated face found" -> skipped
				out.warnings += "face with fewer than 3 corners skipped at line " + std::to_string(lineNo) + "\n";
				continue;
			}
			triangulate(out, corners, activeMaterial);
		} else if (tag == "mtllib") {
			std::string name;
			while (ls >> name) {
				if (baseDir.empty()) continue;
				std::ifstream f(baseDir + "/" + name, std::ios::binary);
				if (!f) { out.warnings += "material file [ " + name + " ] not found\n"; continue; }
				std::stringstream ss; ss << f.rdbuf();
				parseMtl(ss.str(), out.materials);
				break;
			}
		} else if (tag == "usemtl") {
			std::string name; ls >> name;
			activeMaterial = -1;
			for (size_t i = 0; i < out.materials.size(); i++) if (out.materials[i].name == name) activeMaterial = int(i);
			if (activeMaterial < 0) out.warnings += "material [ " + name + " ] not found in .mtl\n";
		}
		// o, g, s: shapes are concatenated in file order by RTModel.cpp:60-61, so they have no effect on the triangle list
	}
	return true;
}

bool load(const std::string& path, Mesh& out, std::string& err) {
	std::ifstream f(path, std::ios::binary);
	if (!f) { err = "cannot open " + path; return false; }
	std::stringstream ss;
	ss << f.rdbuf();
	const size_t slash = path.find_last_of('/');
	return parse(ss.str(), out, err, slash == std::string::npos ? "." : path.substr(0, slash));
}
}  // namespace obj
