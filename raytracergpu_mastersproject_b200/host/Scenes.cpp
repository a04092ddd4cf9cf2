#include "Scenes.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <sstream>
#include <stdexcept>

#include "VulkanWrapper/GameObject.hpp"
#include "VulkanWrapper/RTModel.hpp"

using SceneTypes::MaterialType;
using Mat = SceneTypes::GPU::Material;
using Tri = SceneTypes::CPU::Triangle;
using ScenePtr = std::unique_ptr<RaytraceScene>;

namespace {

// ---- deterministic scene RNG (PCG RXS-M-XS 32); the reference seeds mt19937 from std::random_device (D9) ----
struct SceneRng {
	u32 s;
	explicit SceneRng(u32 seed) : s(seed * 747796405u + 2891336453u) {}
	u32 next() { s = s * 747796405u + 2891336453u; u32 w = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u; return (w >> 22) ^ w; }
	f32 uniform() { return f32(next() >> 8) * (1.0f / 16777216.0f); }            // [0,1)
	f32 uniform(f32 lo, f32 hi) { return lo + (hi - lo) * uniform(); }
	glm::vec3 color(f32 lo = 0.0f, f32 hi = 1.0f) { f32 r = uniform(lo, hi), g = uniform(lo, hi), b = uniform(lo, hi); return { r, g, b }; }
};

const f32 kHalfPi = glm::pi<f32>() / 2;

// place a model in the scene: translation / scale / rotation, then hand the object over
void place(ScenePtr& scene, std::shared_ptr<RTModel> model, bool triangular, glm::vec3 at, glm::vec3 scale = { 1.0f, 1.0f, 1.0f },
           glm::vec3 rot = { 0.0f, 0.0f, 0.0f }) {
	GameObject obj = GameObject::createGameObject();
	obj.setModel(std::move(model), triangular);
	obj.transform.translation = at;
	obj.transform.scale = scale;
	obj.transform.rotation = rot;
	scene->addGameObject(std::move(obj));
}
std::shared_ptr<RTModel> quad(const Mat& m) { return loadModel("models/quad.obj", m); }
std::shared_ptr<RTModel> cube(const Mat& m) { return loadModel("models/cube.obj", m); }
std::shared_ptr<RTModel> ball(f32 radius, const Mat& m) { return loadModel(radius, m); }
Mat diffuse(f32 r, f32 g, f32 b) { return Mat({ r, g, b }, MaterialType::DIFFUSE); }
Mat emitter(f32 v) { return Mat({ v, v, v }, MaterialType::LIGHT); }

void finish(ScenePtr& scene, u32 depth, f32 fov, u32 raysPerPixel = 0) {
	if (raysPerPixel) scene->setRaysPerPixel(raysPerPixel);
	scene->setMaxRaytraceDepth(depth);
	scene->getCamera().setVerticalFOV(fov);
	scene->prepForRender();
}

// the 550-unit room shared by the Cornell-style scenes: wall half-extent 275, light panel 130 x 100 under the ceiling
const glm::vec3 kWallScale{ 275.0f, 1.0f, 275.0f };
void addRoomLight(ScenePtr& scene) { place(scene, quad(emitter(15.0f)), true, { 275.0f, 549.0f, 300.0f }, { 65.0f, 1.0f, 50.0f }); }
void addFloor(ScenePtr& scene, const Mat& m) { place(scene, quad(m), true, { 275.0f, 0.0f, 275.0f }, kWallScale); }
void addBackWall(ScenePtr& scene, const Mat& m) { place(scene, quad(m), true, { 275.0f, 275.0f, 550.0f }, kWallScale, { kHalfPi, 0.0f, 0.0f }); }

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// the reference's five scenes (Scenes.cpp:20-398): same objects, transforms, materials, insertion order, parameters
// ------------------------------------------------------------------------------------------------------------
auto randomSpheres(ScenePtr& scene) -> void {            // Scenes.cpp:20-114
	SceneRng rng(20);
	place(scene, ball(1000.0f, diffuse(0.5f, 0.5f, 0.5f)), false, { 0.0f, -1000.0f, 0.0f });
	const Mat glow = emitter(1.0f);
	const Mat tintA(rng.color(), MaterialType::DIFFUSE);
	const Mat mirror(rng.color(), MaterialType::METALLIC);
	const Mat tintB(rng.color(), MaterialType::DIFFUSE);
	struct Item { glm::vec3 at; f32 radius; const Mat* mat; };
	const Item items[] = {
		{ { 0.0f, 1.05f, 3.5f }, 1.0f, &glow },  { { -4.0f, 1.0f, 4.0f }, 2.0f, &tintA },  { { 4.0f, 1.0f, 4.0f }, 3.0f, &mirror },
		{ { 0.0f, 1.0f, -4.0f }, 1.0f, &tintB }, { { 4.0f, 1.0f, -4.0f }, 2.0f, &tintB },  { { -4.0f, 1.0f, -4.0f }, 3.0f, &tintB },
		{ { 4.0f, 5.0f, 4.0f }, 1.0f, &tintA },  { { -4.0f, 5.0f, 4.0f }, 1.0f, &tintA },
	};
	for (const Item& it : items) place(scene, ball(it.radius, *it.mat), false, it.at);
	place(scene, loadModel("models/quad.obj", glm::vec3(0.3f, 0.5f, 0.7f)), true, { 0.0f, 1.0f, 5.0f }, { 5.0f, 1.0f, 3.0f }, { 4.0f, 0.0f, 0.0f });
	finish(scene, 5, 90.0f);                             // raysPerPixel is never set by the reference here (U14 -> 1)
}

auto cornellMixedScene(ScenePtr& scene) -> void {        // Scenes.cpp:116-202
	const Mat red = diffuse(1.0f, 0.1f, 0.1f), green = diffuse(0.1f, 1.0f, 0.1f), blue = diffuse(0.1f, 0.1f, 1.0f), white = diffuse(1.0f, 1.0f, 1.0f);
	// (the reference also loads models/smooth_vase.obj here but never adds it to the scene, :121,:174)
	place(scene, quad(green), true, { 0.0f, -4.0f, 10.0f }, { 10.0f, 1.0f, 3.0f });
	place(scene, quad(white), true, { -10.0f, 1.0f, 10.0f }, { 5.0f, 1.0f, 3.0f }, { 0.0f, 0.0f, kHalfPi });
	place(scene, quad(white), true, { 10.0f, 1.0f, 10.0f }, { 5.0f, 1.0f, 3.0f }, { 0.0f, 0.0f, -kHalfPi });
	place(scene, quad(red), true, { 0.0f, 6.0f, 10.0f }, { 10.0f, 1.0f, 3.0f });
	place(scene, quad(blue), true, { 0.0f, 1.0f, 13.0f }, { 10.0f, 1.0f, 6.0f }, { kHalfPi, 0.0f, 0.0f });
	place(scene, ball(0.5f, emitter(1.0f)), false, { 0.0f, 6.0f, 10.0f });
	std::shared_ptr<RTModel> box = cube(white);
	place(scene, box, true, { -5.0f, -2.0f, 10.0f }, { 2.0f, 2.0f, 2.0f }, { 0.0f, glm::pi<f32>() / 3, 0.0f });
	place(scene, box, true, { 4.0f, -2.0f, 10.0f }, { 2.0f, 2.0f, 2.0f }, { 0.0f, -glm::pi<f32>() / 4, 0.0f });
	place(scene, ball(2.0f, diffuse(0.3f, 0.5f, 0.7f)), false, { -0.5f, 0.0f, 12.0f });
	finish(scene, 100, 80.0f);
}

auto cornellBoxScene(ScenePtr& scene) -> void {          // Scenes.cpp:204-288
	const Mat red = diffuse(0.65f, 0.05f, 0.05f), white = diffuse(0.73f, 0.73f, 0.73f), green = diffuse(0.12f, 0.45f, 0.15f);
	place(scene, quad(green), true, { 550.0f, 275.0f, 275.0f }, kWallScale, { 0.0f, 0.0f, kHalfPi });
	place(scene, quad(red), true, { 0.0f, 275.0f, 275.0f }, kWallScale, { 0.0f, 0.0f, kHalfPi });
	addRoomLight(scene);
	addFloor(scene, white);
	place(scene, quad(white), true, { 275.0f, 550.0f, 275.0f }, kWallScale);
	addBackWall(scene, white);
	std::shared_ptr<RTModel> box = cube(white);
	place(scene, box, true, { 350.0f, 160.0f, 395.0f }, { 80.0f, 160.0f, 80.0f }, { 0.0f, glm::radians(15.0f), 0.0f });
	place(scene, box, true, { 180.0f, 160.0f, 175.0f }, { 80.0f, 80.0f, 80.0f }, { 0.0f, glm::radians(-18.0f), 0.0f });
	place(scene, ball(1.0f, white), false, { -200.0f, 215.0f, -50.0f });     // the mandatory sphere, parked out of view
	finish(scene, 25, 40.0f, 128);
}

auto simpleScene(ScenePtr& scene) -> void {              // Scenes.cpp:290-335
	const Mat white = diffuse(0.73f, 0.73f, 0.73f), red = diffuse(0.65f, 0.05f, 0.05f);
	addRoomLight(scene);
	addFloor(scene, white);
	addBackWall(scene, red);
	place(scene, ball(200.0f, white), false, { 100.0f, 445.0f, 215.0f });
	finish(scene, 8, 40.0f, 16);
}

auto complexScene(ScenePtr& scene) -> void {             // Scenes.cpp:337-398 (the default scene, RaytracerBVH.cpp:511)
	const Mat wall = diffuse(0.33f, 0.73f, 0.33f), red = diffuse(0.65f, 0.05f, 0.05f), blue = diffuse(0.12f, 0.15f, 0.45f);
	addRoomLight(scene);
	addFloor(scene, wall);
	addBackWall(scene, wall);
	place(scene, loadModel("models/monkey.obj", blue), true, { 175.0f, 125.0f, 275.0f }, { 100.0f, 100.0f, 100.0f },
	      { glm::radians(-36.0f), glm::radians(180.0f), glm::radians(21.0f) });
	place(scene, ball(40.0f, red), false, { 100.0f, 215.0f, 50.0f });
	finish(scene, 8, 40.0f, 8);
}

// ------------------------------------------------------------------------------------------------------------
// synthetic generators (additive)
// ------------------------------------------------------------------------------------------------------------
namespace {

// lattice value noise + fBm from an integer hash: no libm, bit-reproducible everywhere
u32 hash2(u32 x, u32 y, u32 seed) {
	u32 h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u) ^ (seed * 0xC2B2AE3Du);
	h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
	return h;
}
f32 lattice(u32 x, u32 y, u32 seed) { return f32(hash2(x, y, seed) >> 8) * (1.0f / 16777216.0f); }
f32 valueNoise(f32 x, f32 y, u32 seed) {
	const f32 fx = std::floor(x), fy = std::floor(y);
	const u32 ix = u32(i32(fx)), iy = u32(i32(fy));
	f32 tx = x - fx, ty = y - fy;
	tx = tx * tx * (3.0f - 2.0f * tx); ty = ty * ty * (3.0f - 2.0f * ty);
	const f32 a = lattice(ix, iy, seed), b = lattice(ix + 1, iy, seed), c = lattice(ix, iy + 1, seed), d = lattice(ix + 1, iy + 1, seed);
	return (a + (b - a) * tx) + ((c + (d - c) * tx) - (a + (b - a) * tx)) * ty;
}
f32 fbm(f32 x, f32 y, u32 seed, int octaves) {
	f32 sum = 0.0f, amp = 0.5f;
	for (int o = 0; o < octaves; o++) { sum += amp * valueNoise(x, y, seed + u32(o)); x *= 2.0f; y *= 2.0f; amp *= 0.5f; }
	return sum;                                                        // in [0, 1)
}

// The reference's Morton code of a world-space centroid (GenerateMortonCodesOfPrimitives.comp:41-65), used ONLY to
// choose the emission order of generated primitives (D8) -- the device recomputes the real codes every frame.
struct MortonFrame {
	glm::vec3 lo{ 0.0f, 0.0f, 0.0f }, hi{ 0.0f, 0.0f, 0.0f };          // pin U4: the reduction starts from 0
	void add(glm::vec3 c) {
		lo = { std::min(lo.x, c.x), std::min(lo.y, c.y), std::min(lo.z, c.z) };
		hi = { std::max(hi.x, c.x), std::max(hi.y, c.y), std::max(hi.z, c.z) };
	}
	static u32 spread(u32 v) {
		if (v == 1024u) v--;
		v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu;
		v = (v | (v << 4)) & 0x030C30C3u;  v = (v | (v << 2)) & 0x09249249u;
		return v;
	}
	static u32 quant(f32 c, f32 span) { const f32 q = (c / span) * 1024.0f; return q > 0.0f ? (q >= 4294967296.0f ? 0xFFFFFFFFu : u32(q)) : 0u; }
	u32 code(glm::vec3 c) const {
		glm::vec3 l = lo, h = hi;
		for (int k = 0; k < 3; k++) if (h[k] - l[k] < 0.001f) { l[k] -= 0.0005f; h[k] += 0.0005f; }
		return spread(quant(c.z, h.z - l.z)) << 2 | spread(quant(c.y, h.y - l.y)) << 1 | spread(quant(c.x, h.x - l.x));
	}
};
glm::vec3 worldPoint(const glm::mat4& m, glm::vec3 p) { const glm::vec4 r = m * glm::vec4(p, 1.0f); return { r.x, r.y, r.z }; }
glm::vec3 centroid(glm::vec3 a, glm::vec3 b, glm::vec3 c) { return { ((a.x + b.x) + c.x) / 3.0f, ((a.y + b.y) + c.y) / 3.0f, ((a.z + b.z) + c.z) / 3.0f }; }

// centroids of the fixed room pieces so that the generator sees the same enclosing box as the device will
void addQuadCentroids(MortonFrame& f, glm::vec3 at, glm::vec3 scale, glm::vec3 rot = { 0.0f, 0.0f, 0.0f }) {
	const glm::mat4 m = TransformComponent(at, scale, rot).mat4();
	static const std::shared_ptr<RTModel> unitQuad = quad(diffuse(0.0f, 0.0f, 0.0f));
	for (const Tri& t : getVariantFromSharedPtr<RTModel_Triangles>(unitQuad)->getTriangles())
		f.add(centroid(worldPoint(m, t.v0), worldPoint(m, t.v1), worldPoint(m, t.v2)));
}

// stable order of `tris` by the reference Morton code of their world-space centroids
void sortTrianglesByMorton(std::vector<Tri>& tris, const glm::mat4& m, const MortonFrame& frame) {
	std::vector<u32> codes(tris.size());
	for (size_t i = 0; i < tris.size(); i++)
		codes[i] = frame.code(centroid(worldPoint(m, tris[i].v0), worldPoint(m, tris[i].v1), worldPoint(m, tris[i].v2)));
	std::vector<u32> order(tris.size());
	std::iota(order.begin(), order.end(), 0u);
	std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return codes[a] < codes[b]; });
	std::vector<Tri> sorted(tris.size());
	for (size_t i = 0; i < tris.size(); i++) sorted[i] = tris[order[i]];
	tris.swap(sorted);
}

}  // namespace

auto SyntheticScenes::meshRoomScene(ScenePtr& scene, u32 segments, u32 seed) -> void {
	if (segments < 3) throw std::runtime_error("meshRoomScene: segments must be >= 3");
	const Mat white = diffuse(0.73f, 0.73f, 0.73f), red = diffuse(0.65f, 0.05f, 0.05f), body = diffuse(0.45f, 0.55f, 0.25f);
	// a closed, bumpy blob: unit sphere sampled on a (segments+1) x segments lat/long grid that stops short of the
	// poles (no zero-area triangles), radius modulated by fBm; 2 * segments^2 triangles
	const u32 rings = segments + 1;
	std::vector<glm::vec3> pts(size_t(rings) * segments);
	for (u32 r = 0; r < rings; r++) {
		const f32 v = 0.02f + 0.96f * (f32(r) / f32(segments));            // latitude fraction in [0.02, 0.98]
		const f32 phi = v * glm::pi<f32>();
		for (u32 s = 0; s < segments; s++) {
			const f32 u = f32(s) / f32(segments);
			const f32 theta = u * 2.0f * glm::pi<f32>();
			const f32 wrap = std::fabs(2.0f * u - 1.0f);                     // seamless in longitude
			const f32 bump = 0.75f + 0.5f * fbm(6.0f * wrap + 3.0f, 8.0f * v + 1.0f, seed, 5);
			pts[size_t(r) * segments + s] = { bump * std::sin(phi) * std::cos(theta), bump * std::cos(phi), bump * std::sin(phi) * std::sin(theta) };
		}
	}
	std::vector<Tri> tris;
	tris.reserve(size_t(2) * segments * segments);
	for (u32 r = 0; r < segments; r++)
		for (u32 s = 0; s < segments; s++) {
			const u32 s1 = (s + 1) % segments;
			const glm::vec3 a = pts[size_t(r) * segments + s], b = pts[size_t(r) * segments + s1];
			const glm::vec3 c = pts[size_t(r + 1) * segments + s1], d = pts[size_t(r + 1) * segments + s];
			tris.push_back({ a, b, c });
			tris.push_back({ a, c, d });
		}
	const glm::vec3 at{ 275.0f, 200.0f, 275.0f }, scale{ 150.0f, 150.0f, 150.0f };
	const glm::mat4 m = TransformComponent(at, scale, { 0.0f, 0.0f, 0.0f }).mat4();
	MortonFrame frame;
	addQuadCentroids(frame, { 275.0f, 549.0f, 300.0f }, { 65.0f, 1.0f, 50.0f });
	addQuadCentroids(frame, { 275.0f, 0.0f, 275.0f }, kWallScale);
	addQuadCentroids(frame, { 275.0f, 275.0f, 550.0f }, kWallScale, { kHalfPi, 0.0f, 0.0f });
	frame.add({ -200.0f, 215.0f, -50.0f });
	for (const Tri& t : tris) frame.add(centroid(worldPoint(m, t.v0), worldPoint(m, t.v1), worldPoint(m, t.v2)));
	sortTrianglesByMorton(tris, m, frame);
	addRoomLight(scene);
	addFloor(scene, white);
	addBackWall(scene, red);
	place(scene, loadModel(std::move(tris), body), true, at, scale);
	place(scene, ball(1.0f, white), false, { -200.0f, 215.0f, -50.0f });
	finish(scene, 8, 40.0f, 64);
}

auto SyntheticScenes::sphereFieldScene(ScenePtr& scene, u32 count, u32 seed) -> void {
	SceneRng rng(seed);
	struct Item { glm::vec3 c; f32 r; Mat m; u32 code; };
	std::vector<Item> items(count);
	MortonFrame frame;
	addQuadCentroids(frame, { 275.0f, 549.0f, 300.0f }, { 65.0f, 1.0f, 50.0f });
	addQuadCentroids(frame, { 275.0f, 0.0f, 275.0f }, kWallScale);
	for (Item& it : items) {
		it.c = { rng.uniform(25.0f, 525.0f), rng.uniform(25.0f, 525.0f), rng.uniform(25.0f, 525.0f) };
		it.r = rng.uniform(1.0f, 4.0f);
		const f32 pick = rng.uniform();
		const MaterialType type = pick < 0.70f ? MaterialType::DIFFUSE : (pick < 0.85f ? MaterialType::METALLIC : MaterialType::DIELECTRIC);
		it.m = Mat(rng.color(0.1f, 0.9f), type);
		frame.add(it.c);
	}
	for (Item& it : items) it.code = frame.code(it.c);
	std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.code < b.code; });
	addRoomLight(scene);
	addFloor(scene, diffuse(0.73f, 0.73f, 0.73f));
	for (const Item& it : items) place(scene, ball(it.r, it.m), false, it.c);
	finish(scene, 8, 40.0f, 256);
}

auto SyntheticScenes::heightFieldScene(ScenePtr& scene, u32 nx, u32 nz, u32 seed, u32 dielectricPercent) -> void {
	if (nx == 0 || nz == 0) throw std::runtime_error("heightFieldScene: empty grid");
	const f32 extent = 550.0f, maxHeight = 100.0f;
	std::vector<f32> height(size_t(nx + 1) * (nz + 1));
	for (u32 z = 0; z <= nz; z++)
		for (u32 x = 0; x <= nx; x++)
			height[size_t(z) * (nx + 1) + x] = maxHeight * fbm(6.0f * f32(x) / f32(nx), 6.0f * f32(z) / f32(nz), seed, 6);
	auto vertex = [&](u32 x, u32 z) -> glm::vec3 { return { extent * (f32(x) / f32(nx)), height[size_t(z) * (nx + 1) + x], extent * (f32(z) / f32(nz)) }; };
	std::vector<Tri> tris;
	tris.reserve(size_t(2) * nx * nz);
	for (u32 z = 0; z < nz; z++)
		for (u32 x = 0; x < nx; x++) {
			const glm::vec3 a = vertex(x, z), b = vertex(x + 1, z), c = vertex(x + 1, z + 1), d = vertex(x, z + 1);
			tris.push_back({ a, c, b });
			tris.push_back({ a, d, c });
		}
	const glm::mat4 identity = TransformComponent().mat4();
	MortonFrame frame;
	addQuadCentroids(frame, { 275.0f, 549.0f, 300.0f }, { 65.0f, 1.0f, 50.0f });
	frame.add({ -200.0f, 215.0f, -50.0f });
	for (const Tri& t : tris) frame.add(centroid(t.v0, t.v1, t.v2));
	sortTrianglesByMorton(tris, identity, frame);
	addRoomLight(scene);
	// Materials are per GameObject, so "x % of the triangles are dielectric" is expressed by cutting the Morton-ordered
	// list into contiguous chunks (one GameObject each): the flattened order stays globally Morton-sorted.
	const Mat ground = diffuse(0.55f, 0.5f, 0.4f), glass(glm::vec3(0.9f, 0.95f, 1.0f), MaterialType::DIELECTRIC);
	const size_t chunk = dielectricPercent ? std::clamp<size_t>(tris.size() / 256, 64, 4096) : tris.size();
	SceneRng rng(seed ^ 0x5EEDu);
	for (size_t begin = 0; begin < tris.size(); begin += chunk) {
		const size_t end = std::min(tris.size(), begin + chunk);
		const bool isGlass = dielectricPercent && (rng.next() % 100u) < dielectricPercent;
		place(scene, loadModel(std::vector<Tri>(tris.begin() + begin, tris.begin() + end), isGlass ? glass : ground), true, { 0.0f, 0.0f, 0.0f });
	}
	place(scene, ball(1.0f, ground), false, { -200.0f, 215.0f, -50.0f });
	finish(scene, dielectricPercent ? 16 : 8, 40.0f, dielectricPercent ? 1024 : 256);
}

auto SyntheticScenes::buildByName(ScenePtr& scene, const std::string& spec) -> void {
	std::vector<std::string> parts;
	std::stringstream ss(spec);
	for (std::string p; std::getline(ss, p, ':');) parts.push_back(p);
	if (parts.empty()) throw std::runtime_error("empty scene name");
	auto arg = [&](size_t i, u32 dflt) -> u32 { return parts.size() > i && !parts[i].empty() ? u32(std::strtoul(parts[i].c_str(), nullptr, 10)) : dflt; };
	const std::string& n = parts[0];
	if (n == "complexScene") complexScene(scene);
	else if (n == "simpleScene") simpleScene(scene);
	else if (n == "cornellBoxScene") cornellBoxScene(scene);
	else if (n == "cornellMixedScene") cornellMixedScene(scene);
	else if (n == "randomSpheres") randomSpheres(scene);
	else if (n == "meshRoom") meshRoomScene(scene, arg(1, 660), arg(2, 1));
	else if (n == "sphereField") sphereFieldScene(scene, arg(1, 100000), arg(2, 2));
	else if (n == "heightField") heightFieldScene(scene, arg(1, 3162), arg(2, 1581), arg(3, 3), arg(4, 0));
	else throw std::runtime_error("unknown scene '" + n + "'");
}
