// host_capi.cpp -- tiny C interface over the C++ host mirror so that Python (tests, bench.py) gets its scene
// arrays from the SAME scene builders / flatten code the C++ host uses.  Host-only: no device is touched.
#include <cstring>
#include <memory>
#include <string>

#include "Scenes.hpp"
#include "utils/Png.hpp"
#include "VulkanWrapper/RTModel.hpp"
#include "VulkanWrapper/RaytraceScene.hpp"

namespace { thread_local std::string g_err; }

extern "C" {

const char* rtbh_last_error(void) { return g_err.c_str(); }
void rtbh_set_model_dir(const char* dir) { setModelSearchPath(dir); }

// Build a scene by spec (see SyntheticScenes::buildByName) and flatten it.  Returns an opaque handle or NULL.
void* rtbh_scene_create(const char* spec) {
	try {
		auto scene = std::make_unique<RaytraceScene>(RaytraceScene::HostOnly{});
		SyntheticScenes::buildByName(scene, spec);
		return scene.release();
	} catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void rtbh_scene_destroy(void* h) { delete static_cast<RaytraceScene*>(h); }

// counts[4] = models, triangles, spheres, materials ; params[2] = raysPerPixel, maxRaytraceDepth ; fov
void rtbh_scene_info(void* h, unsigned* counts, unsigned* params, float* fov) {
	auto* s = static_cast<RaytraceScene*>(h);
	counts[0] = unsigned(s->hostModels().size()); counts[1] = unsigned(s->hostTriangles().size());
	counts[2] = unsigned(s->hostSpheres().size()); counts[3] = unsigned(s->hostMaterials().size());
	params[0] = s->getRaysPerPixel(); params[1] = s->getMaxRaytraceDepth();
	*fov = s->getCamera().getVerticalFOV();
}
void rtbh_scene_copy(void* h, void* models, void* triangles, void* spheres, void* materials) {
	auto* s = static_cast<RaytraceScene*>(h);
	std::memcpy(models, s->hostModels().data(), s->hostModels().size() * sizeof(SceneTypes::GPU::Model));
	std::memcpy(triangles, s->hostTriangles().data(), s->hostTriangles().size() * sizeof(SceneTypes::GPU::Triangle));
	std::memcpy(spheres, s->hostSpheres().data(), s->hostSpheres().size() * sizeof(SceneTypes::GPU::Sphere));
	std::memcpy(materials, s->hostMaterials().data(), s->hostMaterials().size() * sizeof(SceneTypes::GPU::Material));
}
// TransformComponent::mat4 for (translation, scale, rotation) -> 16 floats column-major (parity test of A3)
void rtbh_transform_mat4(const float* t, const float* sc, const float* r, float* out) {
	TransformComponent tc({ t[0], t[1], t[2] }, { sc[0], sc[1], sc[2] }, { r[0], r[1], r[2] });
	const glm::mat4 m = tc.mat4();
	std::memcpy(out, &m, 64);
}
// triangle count of an OBJ file through loadModel (RTModel.cpp path) ; positions copied if out != NULL (9 floats each)
long rtbh_load_obj(const char* path, float* out, long maxTriangles) {
	try {
		std::shared_ptr<RTModel> m = loadModel(std::string(path), glm::vec3(1.0f, 1.0f, 1.0f));
		const auto& tris = getVariantFromSharedPtr<RTModel_Triangles>(m)->getTriangles();
		if (out) for (long i = 0; i < long(tris.size()) && i < maxTriangles; i++) std::memcpy(out + 9 * i, &tris[size_t(i)], 36);
		return long(tris.size());
	} catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// the headless hosts' PNG writer (utils/Png.hpp), for the CPU test of the encoder
int rtbh_write_png(const char* path, const unsigned char* rgba, unsigned width, unsigned height) {
	try { png::writeRGB(path, rgba, width, height); return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

}  // extern "C"
