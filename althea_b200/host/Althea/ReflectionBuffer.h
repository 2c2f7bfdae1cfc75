// Include/Althea/ReflectionBuffer.h:27-73: one RGBA16F image with 5 mips (Src/ReflectionBuffer.cpp:28-39).
#pragma once
#include "CudaApplication.h"

namespace AltheaEngine {

class ReflectionBuffer {
public:
  static constexpr uint32_t kMipCount = 5;
  ReflectionBuffer() = default;
  ReflectionBuffer(const CudaApplication& app, uint32_t width, uint32_t height)
      : _app(&app), _reflectionBuffer(app, ALTHEA_FORMAT_R16G16B16A16_SFLOAT, width, height, kMipCount) {}
  uint64_t getHandle() const { return _reflectionBuffer.handle(); }
  ImageResource& getImage() { return _reflectionBuffer; }
  // Src/ReflectionBuffer.cpp:163-293: mips 1..4, 7-tap separable Gaussian, axis alternating V,H,V,H
  void convolveReflectionBuffer(const althea_sync* sync = nullptr) {
    _app->check(althea_cuda_glossy_convolve(_app->ctx(), _reflectionBuffer.handle(), sync), "althea_cuda_glossy_convolve");
  }

private:
  const CudaApplication* _app = nullptr;
  ImageResource _reflectionBuffer;
};

} // namespace AltheaEngine
