// part of the minimal glm stand-in: everything lives in glm/glm.hpp
#pragma once
#include "../glm.hpp"
