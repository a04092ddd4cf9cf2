#pragma once
#include "../../utils/PrimitiveTypes.hpp"
struct PointLightComponent { f32 lightIntensity = 1.0f; };
