// Include/Althea/PointLight.h:31-153: the light SSBO, the omni shadow cube array the deferred pass and SSR read, and
// drawShadowMaps (Src/PointLight.cpp:235-282), which renders the cubes from the models' primitives.
#pragma once
#include "Camera.h"
#include "CudaApplication.h"
#include "Model.h"

#include <cmath>
#include <cstring>

namespace AltheaEngine {

struct PointLight { // alignas(16) vec3 position; alignas(16) vec3 emission;  == althea_point_light
  float position[3];
  float _pad0 = 0.0f;
  float emission[3];
  float _pad1 = 0.0f;
};
static_assert(sizeof(PointLight) == 32 && sizeof(althea_point_light) == 32, "PointLight.h:31-34");

class PointLightCollection {
public:
  PointLightCollection() = default;
  // reference: (app, commandBuffer, heap, lightCount, createShadowMap, primConstants); shadow faces are 256^2 (PointLight.cpp:55-62)
  PointLightCollection(const CudaApplication& app, size_t lightCount, bool createShadowMap, uint32_t shadowRes = 256)
      : _app(&app), _lights(lightCount), _dirty(true) {
    if (lightCount) app.check(althea_cuda_create_buffer(app.ctx(), lightCount * sizeof(PointLight), &_buffer), "althea_cuda_create_buffer");
    if (createShadowMap && lightCount)
      _shadowMap = ImageResource(app, ALTHEA_FORMAT_R32_SFLOAT, shadowRes, shadowRes, 1, (uint32_t)(6 * lightCount));
  }
  ~PointLightCollection() {
    if (_buffer && _app) althea_cuda_release(_app->ctx(), _buffer);
  }
  PointLightCollection(PointLightCollection&& o) noexcept { *this = std::move(o); }
  PointLightCollection& operator=(PointLightCollection&& o) noexcept {
    std::swap(_app, o._app); std::swap(_lights, o._lights); std::swap(_dirty, o._dirty); std::swap(_buffer, o._buffer);
    _shadowMap = std::move(o._shadowMap);
    return *this;
  }

  void setLight(uint32_t lightId, const PointLight& light) { _lights.at(lightId) = light; _dirty = true; }
  const PointLight& getLight(uint32_t lightId) const { return _lights.at(lightId); }
  size_t getCount() const { return _lights.size(); }
  void updateResource() { // PointLight.cpp:193-203: memcpy the lights into this frame's SSBO slot
    if (_dirty && !_lights.empty()) {
      _app->check(althea_cuda_upload(_app->ctx(), _buffer, _lights.data(), _lights.size() * sizeof(PointLight), nullptr), "althea_cuda_upload(lights)");
      _dirty = false;
    }
  }
  // PointLight.cpp:235-282: one multiview pass per light over every primitive of every model
  void drawShadowMaps(const std::vector<Model>& models, const althea_sync* sync = nullptr) {
    if (_lights.empty() || !_shadowMap.handle()) return;
    updateResource();
    const std::vector<althea_primitive> prims = describeModels(models);
    const althea_point_light_constants c = pointLightConstants();
    _app->check(althea_cuda_draw_shadow_cubes(_app->ctx(), _buffer, (uint32_t)_lights.size(), &c, prims.data(), (uint32_t)prims.size(),
                                              _shadowMap.handle(), sync),
                "althea_cuda_draw_shadow_cubes");
  }
  // PointLightConstants as the constructor builds them (PointLight.cpp:72-118): a 90 degree, aspect 1, 0.01 / 1000 Camera at the
  // origin turned to X+ X- Y+ Y- Z+ Z- through setRotationDegrees, views = computeView(), inverses by glm::inverse (Camera.h)
  static althea_point_light_constants pointLightConstants() {
    althea_point_light_constants c;
    static_assert(sizeof c == 14 * 16 * sizeof(float), "projection, inverseProjection, views[6], inverseViews[6]");
    pointLightConstantMatrices(reinterpret_cast<float*>(&c));
    return c;
  }
  // texel = length(p - light) / 1000 (Shaders/ShadowMapBindless.frag:41), layer = 6 * light + face
  void uploadShadowMaps(const float* cubes, size_t floats) { _shadowMap.upload(cubes, floats * sizeof(float)); }
  uint64_t bufferHandle() const { return _buffer; }
  uint64_t shadowMapHandle() const { return _shadowMap.handle(); }

private:
  const CudaApplication* _app = nullptr;
  std::vector<PointLight> _lights;
  bool _dirty = false;
  uint64_t _buffer = 0;
  ImageResource _shadowMap;
};

} // namespace AltheaEngine
