// Scene camera.  Only the vertical FOV reaches the kernels: the renderer hosts hard-code position / look-at / up in
// the UBO (RaytracerBVH.hpp:340-344) and ignore the rasteriser-style matrices and keyboard input, which need a
// window and are out of scope for the headless build (reference API: CameraGameObject.hpp:16-37).
#pragma once

#include "../utils/PrimitiveTypes.hpp"
#include "GameObject.hpp"

class Window;

class CameraGameObject : public GameObject {
	f32 fovy = 50.0f;
	explicit CameraGameObject(GameObjectId id) : GameObject(id) {}
public:
	static auto makeCameraGameObject() -> CameraGameObject { return CameraGameObject(GameObject::createGameObject().getId()); }
	auto getVerticalFOV() const -> f32 { return fovy; }
	auto setVerticalFOV(f32 f) -> void { fovy = f; }
	auto updateCameraForFrame(f32 /*dt*/, f32 /*aspectRatio*/) -> void {}   // results were unused by the shaders (D10)
};
