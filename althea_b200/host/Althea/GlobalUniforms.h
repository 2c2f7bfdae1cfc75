// Include/Althea/GlobalUniforms.h:15-31: the 416-byte per-frame block, byte-for-byte (no glm dependency here: float[16]
// column-major where the reference has glm::mat4).
#pragma once
#include <althea_cuda.h>

namespace AltheaEngine {
using GlobalUniforms = althea_global_uniforms;
static_assert(sizeof(GlobalUniforms) == 416, "GlobalUniforms must match Shaders/Global/GlobalUniforms.glsl:8-24");
} // namespace AltheaEngine
