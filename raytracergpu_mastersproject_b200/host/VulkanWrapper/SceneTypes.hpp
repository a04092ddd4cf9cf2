// Host mirrors of the device record layouts (reference: VulkanWrapper/SceneTypes.hpp:13-127 ==
// shaders/include/definitions.glsl:6-77).  Sizes / offsets are asserted against the C-ABI records of rtb200.h,
// which the kernels consume byte for byte.
#pragma once

#include <cstddef>
#include <vector>

#include "../glm/glm.hpp"
#include "../utils/PrimitiveTypes.hpp"
#include "rtb200.h"

namespace SceneTypes {
	enum class MaterialType : u32 { LIGHT = 0, DIFFUSE = 1, METALLIC = 2, DIELECTRIC = 3 };

	namespace CPU {
		struct Triangle { glm::vec3 v0, v1, v2; };
		struct Sphere { glm::vec3 center; f32 radius; };
	}

	namespace GPU {
		struct Model {
			glm::mat4 modelMatrix;
			constexpr auto getSize() const -> size_t { return sizeof(Model); }
			bool operator==(const Model& o) const { return modelMatrix == o.modelMatrix; }
		};
		struct Triangle {
			alignas(16) glm::vec3 v0;
			alignas(16) glm::vec3 v1;
			alignas(16) glm::vec3 v2;
			alignas(16) u32 materialIndex;
			u32 modelIndex;
			constexpr auto getSize() const -> size_t { return sizeof(Triangle); }
			static auto convertFromCPUTriangle(const CPU::Triangle& t, u32 matIndex, u32 modIndex) -> Triangle {
				Triangle g{};
				g.v0 = t.v0; g.v1 = t.v1; g.v2 = t.v2; g.materialIndex = matIndex; g.modelIndex = modIndex;
				return g;
			}
			bool operator==(const Triangle& o) const {
				return v0 == o.v0 && v1 == o.v1 && v2 == o.v2 && materialIndex == o.materialIndex && modelIndex == o.modelIndex;
			}
		};
		struct Sphere {
			alignas(16) glm::vec3 center;
			alignas(16) f32 radius;
			u32 materialIndex;
			u32 modelIndex;
			constexpr auto getSize() const -> size_t { return sizeof(Sphere); }
			static auto convertFromCPUSphere(const CPU::Sphere& s, u32 matIndex, u32 modIndex) -> Sphere {
				Sphere g{};
				g.center = s.center; g.radius = s.radius; g.materialIndex = matIndex; g.modelIndex = modIndex;
				return g;
			}
			bool operator==(const Sphere& o) const {
				return center == o.center && radius == o.radius && materialIndex == o.materialIndex && modelIndex == o.modelIndex;
			}
		};
		struct Material {
			alignas(16) glm::vec3 albedo;
			alignas(8) MaterialType materialType;
			Material() : albedo{ 0 }, materialType{ MaterialType::DIFFUSE } {}
			Material(glm::vec3 color, MaterialType type) : albedo{ color }, materialType{ type } {}
			constexpr auto getSize() const -> size_t { return sizeof(Material); }
			bool operator==(const Material& o) const { return albedo == o.albedo && materialType == o.materialType; }
		};
		struct Light {
			f32 area; u32 triangleIndex;
			constexpr auto getSize() const -> size_t { return sizeof(Light); }
			bool operator==(const Light& o) const { return area == o.area && triangleIndex == o.triangleIndex; }
		};
		struct AABB { f32 minX, maxX, minY, maxY, minZ, maxZ; };
		struct MortonPrimitive { u32 code, primitiveIndex, primitiveType; };
		struct BVHNode { AABB aabb; u32 left, right, primitiveIndex, primitiveType; };

		static_assert(sizeof(Model) == sizeof(rtb_model) && sizeof(Model) == 64);
		static_assert(sizeof(Triangle) == sizeof(rtb_triangle) && offsetof(Triangle, v1) == 16 && offsetof(Triangle, v2) == 32 &&
		              offsetof(Triangle, materialIndex) == 48 && offsetof(Triangle, modelIndex) == 52);
		static_assert(sizeof(Sphere) == sizeof(rtb_sphere) && offsetof(Sphere, radius) == 16 && offsetof(Sphere, materialIndex) == 20 &&
		              offsetof(Sphere, modelIndex) == 24);
		static_assert(sizeof(Material) == sizeof(rtb_material) && offsetof(Material, materialType) == 16);
		static_assert(sizeof(BVHNode) == sizeof(rtb_bvh_node) && sizeof(BVHNode) == 40);
		static_assert(sizeof(MortonPrimitive) == sizeof(rtb_morton_primitive) && sizeof(MortonPrimitive) == 12);
	}

	using Material = GPU::Material;
}
