// Declarations the legacy DeferredPass.{vert,frag} use but no longer declare (SURVEY.md 8c-bis defects 1 and 3), supplied as
// rule R2 reads them: `globals` is the 416-byte GlobalUniforms block, `pointLightArr` the light SSBO and `shadowMapArray` the
// omni shadow cube array, bound as the up-to-date SSR.frag:14-36 binds them. (Ours, not reference text.)
#include <Global/GlobalUniforms.glsl>
#include <PointLights.glsl>
#define globals globalUniforms[0]
#define pointLightArr pointLights[0].pointLightArr
layout(set=0, binding=5) uniform samplerCubeArray shadowMapArray;
