// RaytraceScene: owns the GameObjects, flattens them into the four device arrays (Model / Triangle / Sphere /
// Material) and keeps them on the device (reference API: RaytraceScene.hpp:28-137).
#pragma once

#include <memory>
#include <vector>

#include "../utils/PrimitiveTypes.hpp"
#include "Buffer.hpp"
#include "CameraGameObject.hpp"
#include "Device.hpp"
#include "GameObject.hpp"
#include "SceneTypes.hpp"

class RaytraceScene {
	Device* device;                              // null: host-only scene (flatten without device buffers)
	CameraGameObject camera;
	std::vector<GameObject> gameObjects;

	std::vector<SceneTypes::GPU::Model> models;
	std::vector<SceneTypes::GPU::Triangle> triangles;
	std::vector<SceneTypes::GPU::Sphere> spheres;
	std::vector<SceneTypes::GPU::Material> materials;

	std::unique_ptr<Buffer> modelBuffer, triangleBuffer, sphereBuffer, materialBuffer;
	u32 modelCount = 0, triangleCount = 0, sphereCount = 0, materialCount = 0;

	u32 maxRaytraceDepth = 1;                    // the reference leaves these three uninitialised (pin U14)
	u32 raysPerPixel = 1;
	bool buffersCreated = false;

public:
	explicit RaytraceScene(Device& device);
	struct HostOnly {};
	explicit RaytraceScene(HostOnly);            // additive: scene description without a device
	~RaytraceScene();
	RaytraceScene(const RaytraceScene&) = delete;
	RaytraceScene& operator=(const RaytraceScene&) = delete;

	auto setMaxRaytraceDepth(u32 d) -> void { maxRaytraceDepth = d; }
	auto setRaysPerPixel(u32 r) -> void { raysPerPixel = r; }
	auto getMaxRaytraceDepth() -> u32 { return maxRaytraceDepth; }
	auto getRaysPerPixel() -> u32 { return raysPerPixel; }
	auto getCamera() -> CameraGameObject& { return camera; }

	auto addGameObject(GameObject&& gameObject) -> void;
	auto getGameObject(GameObjectId id) -> GameObject&;
	auto removeGameObject(GameObjectId id) -> bool;
	auto removeGameObject(size_t index) -> bool;

	auto prepForRender() -> void;
	auto updateScene() -> void;

	auto getModelBuffer() -> std::unique_ptr<Buffer>&;
	auto getTriangleBuffer() -> std::unique_ptr<Buffer>&;
	auto getSphereBuffer() -> std::unique_ptr<Buffer>&;
	auto getMaterialBuffer() -> std::unique_ptr<Buffer>&;

	auto getModelCount() -> u32 { return modelCount; }
	auto getTriangleCount() -> u32 { return triangleCount; }
	auto getSphereCount() -> u32 { return sphereCount; }
	auto getMaterialCount() -> u32 { return materialCount; }

	// additive read access to the flattened host arrays (what updateScene uploads)
	auto hostModels() const -> const std::vector<SceneTypes::GPU::Model>& { return models; }
	auto hostTriangles() const -> const std::vector<SceneTypes::GPU::Triangle>& { return triangles; }
	auto hostSpheres() const -> const std::vector<SceneTypes::GPU::Sphere>& { return spheres; }
	auto hostMaterials() const -> const std::vector<SceneTypes::GPU::Material>& { return materials; }
	auto gameObjectCount() const -> size_t { return gameObjects.size(); }

private:
	auto moveGameObjectsToHostVectors() -> void;
	auto uploadAll() -> void;
};
