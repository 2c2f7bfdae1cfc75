// Include/Althea/ScreenSpaceReflection.h:26-64.
#pragma once
#include "CudaApplication.h"
#include "GlobalUniforms.h"
#include "ImageBasedLighting.h"
#include "PointLight.h"
#include "ReflectionBuffer.h"

namespace AltheaEngine {
class GBufferResources;

class ScreenSpaceReflection {
public:
  ScreenSpaceReflection() = default;
  ScreenSpaceReflection(const CudaApplication& app, uint32_t width, uint32_t height) : _app(&app), _reflectionBuffer(app, width, height) {}

  // Src/ScreenSpaceReflection.cpp:60-83. The reference passes bindless indices of the uniforms / resources tables; here the
  // tables themselves are passed.
  void captureReflection(const GlobalUniforms& globals, const althea_gbuffer& gBuffer, const IBLResources& ibl, const PointLightCollection* lights,
                         const althea_sync* sync = nullptr) {
    const althea_ibl ib = ibl.getHandles();
    _app->check(althea_cuda_ssr_capture(_app->ctx(), &globals, &gBuffer, &ib, lights ? lights->bufferHandle() : 0, lights ? lights->shadowMapHandle() : 0,
                                        _reflectionBuffer.getHandle(), sync),
                "althea_cuda_ssr_capture");
  }
  void convolveReflectionBuffer(const althea_sync* sync = nullptr) { _reflectionBuffer.convolveReflectionBuffer(sync); } // :85-98
  const ReflectionBuffer& getReflectionBuffer() const { return _reflectionBuffer; }
  ReflectionBuffer& getReflectionBuffer() { return _reflectionBuffer; }

private:
  const CudaApplication* _app = nullptr;
  ReflectionBuffer _reflectionBuffer;
};

} // namespace AltheaEngine
