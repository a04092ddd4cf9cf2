#include "RaytraceScene.hpp"

#include <cstring>
#include <stdexcept>

RaytraceScene::RaytraceScene(Device& d) : device(&d), camera{ CameraGameObject::makeCameraGameObject() } {}
RaytraceScene::RaytraceScene(HostOnly) : device(nullptr), camera{ CameraGameObject::makeCameraGameObject() } {}
RaytraceScene::~RaytraceScene() = default;

auto RaytraceScene::addGameObject(GameObject&& gameObject) -> void {
	if (buffersCreated) throw std::runtime_error("temp error: gameobject added after buffers deployed");
	if (gameObject.getModel() != nullptr) gameObjects.push_back(std::move(gameObject));
}

auto RaytraceScene::getGameObject(GameObjectId id) -> GameObject& {
	for (auto& g : gameObjects)
		if (g.getId() == id) return g;
	throw std::out_of_range("no game object with that id");
}

auto RaytraceScene::removeGameObject(GameObjectId) -> bool { return false; }   // unimplemented in the reference too
auto RaytraceScene::removeGameObject(size_t) -> bool { return false; }

// Flatten order (reference: RaytraceScene.cpp:157-200): objects in insertion order; each contributes one Model and
// one Material; its triangles are appended in mesh order, spheres to a separate array; both carry the indices.
auto RaytraceScene::moveGameObjectsToHostVectors() -> void {
	models.clear(); triangles.clear(); spheres.clear(); materials.clear();
	size_t triTotal = 0;
	for (const auto& g : gameObjects)
		if (auto* t = getVariantFromSharedPtr<RTModel_Triangles>(g.getModel())) triTotal += t->getTriangles().size();
	triangles.reserve(triTotal);
	for (const auto& g : gameObjects) {
		const u32 modelIndex = u32(models.size());
		SceneTypes::GPU::Model m;
		std::memset(static_cast<void*>(&m), 0, sizeof(m));
		m.modelMatrix = g.transform.mat4();
		models.push_back(m);
		const u32 materialIndex = u32(materials.size());
		if (auto* t = getVariantFromSharedPtr<RTModel_Triangles>(g.getModel())) {
			SceneTypes::GPU::Material mat;
			std::memset(static_cast<void*>(&mat), 0, sizeof(mat));
			mat.albedo = t->getMaterialType().albedo; mat.materialType = t->getMaterialType().materialType;
			materials.push_back(mat);
			for (const auto& tri : t->getTriangles()) {
				SceneTypes::GPU::Triangle r;
				std::memset(static_cast<void*>(&r), 0, sizeof(r));           // padding bytes are part of the uploaded record
				r.v0 = tri.v0; r.v1 = tri.v1; r.v2 = tri.v2; r.materialIndex = materialIndex; r.modelIndex = modelIndex;
				triangles.push_back(r);
			}
		} else if (auto* s = getVariantFromSharedPtr<RTModel_Sphere>(g.getModel())) {
			SceneTypes::GPU::Material mat;
			std::memset(static_cast<void*>(&mat), 0, sizeof(mat));
			mat.albedo = s->getMaterialType().albedo; mat.materialType = s->getMaterialType().materialType;
			materials.push_back(mat);
			SceneTypes::GPU::Sphere r;
			std::memset(static_cast<void*>(&r), 0, sizeof(r));
			r.center = s->getCenter(); r.radius = s->getRadius(); r.materialIndex = materialIndex; r.modelIndex = modelIndex;
			spheres.push_back(r);
		}
	}
	// Pin U13: the reference clamps the counts to >= 1 here and then overwrites them with the real lengths when
	// the buffers are (re)written; zero counts are reported as zero.
	modelCount = u32(models.size()); triangleCount = u32(triangles.size());
	sphereCount = u32(spheres.size()); materialCount = u32(materials.size());
}

auto RaytraceScene::uploadAll() -> void {
	if (!device) return;
	auto fit = [&](std::unique_ptr<Buffer>& b, size_t elem, size_t count) {
		if (!b || b->getInstanceCount() < count) b = std::make_unique<Buffer>(*device, elem, u32(count ? count : 1));
	};
	fit(modelBuffer, sizeof(SceneTypes::GPU::Model), models.size());
	fit(triangleBuffer, sizeof(SceneTypes::GPU::Triangle), triangles.size());
	fit(sphereBuffer, sizeof(SceneTypes::GPU::Sphere), spheres.size());
	fit(materialBuffer, sizeof(SceneTypes::GPU::Material), materials.size());
	modelBuffer->writeToBuffer(models.data(), models.size() * sizeof(models[0]));
	triangleBuffer->writeToBuffer(triangles.data(), triangles.size() * sizeof(triangles[0]));
	sphereBuffer->writeToBuffer(spheres.data(), spheres.size() * sizeof(spheres[0]));
	materialBuffer->writeToBuffer(materials.data(), materials.size() * sizeof(materials[0]));
}

auto RaytraceScene::prepForRender() -> void {
	moveGameObjectsToHostVectors();
	uploadAll();
	buffersCreated = true;
}

// Every frame the reference re-flattens and re-uploads the MODEL-SPACE arrays because K1 transforms them in
// place (RaytraceScene.cpp:78-113).  Same here.
auto RaytraceScene::updateScene() -> void {
	if (!buffersCreated) { prepForRender(); return; }
	moveGameObjectsToHostVectors();
	uploadAll();
}

namespace {
std::unique_ptr<Buffer>& need(std::unique_ptr<Buffer>& b) {
	if (!b) throw std::runtime_error("scene buffers are not created (host-only scene or prepForRender not called)");
	return b;
}
}
auto RaytraceScene::getModelBuffer() -> std::unique_ptr<Buffer>& { return need(modelBuffer); }
auto RaytraceScene::getTriangleBuffer() -> std::unique_ptr<Buffer>& { return need(triangleBuffer); }
auto RaytraceScene::getSphereBuffer() -> std::unique_ptr<Buffer>& { return need(sphereBuffer); }
auto RaytraceScene::getMaterialBuffer() -> std::unique_ptr<Buffer>& { return need(materialBuffer); }
