// Host-side C++ mirror of the reference's interface for the deferred screen-space path (header-only, C++17).
//
// The reference's classes on the path take `const Application&` (the Vulkan device owner, Include/Althea/Application.h) as
// their first argument. Here that role is played by AltheaEngine::CudaApplication, which owns one althea_cuda_ctx; every
// class forwards to the C ABI in include/althea_cuda.h and rethrows failures as std::runtime_error, which is how the
// reference reports errors (e.g. Src/Allocator.cpp:210, Src/ImageBasedLighting.cpp:424).
#pragma once
#include <althea_cuda.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace AltheaEngine {

class CudaApplication {
public:
  explicit CudaApplication(int cudaDevice = 0, const uint8_t* vkDeviceUuid = nullptr, uint32_t ctxFlags = 0) {
    if (althea_cuda_create(&_ctx, cudaDevice, vkDeviceUuid) != ALTHEA_OK)
      throw std::runtime_error(std::string("althea_cuda_create: ") + althea_cuda_last_error(nullptr));
    if (ctxFlags) check(althea_cuda_set_flags(_ctx, ctxFlags), "althea_cuda_set_flags");
  }
  ~CudaApplication() { althea_cuda_destroy(_ctx); }
  CudaApplication(const CudaApplication&) = delete;
  CudaApplication& operator=(const CudaApplication&) = delete;

  althea_cuda_ctx* ctx() const { return _ctx; }
  void check(int rc, const char* what) const {
    if (rc != ALTHEA_OK) throw std::runtime_error(std::string(what) + ": " + althea_cuda_last_error(_ctx));
  }
  void waitIdle() const { check(althea_cuda_synchronize(_ctx, nullptr), "althea_cuda_synchronize"); } // vkDeviceWaitIdle's role
  uint64_t launchCount() const { return althea_cuda_launch_count(_ctx); }

private:
  althea_cuda_ctx* _ctx = nullptr;
};

// ImageResource's role (Include/Althea/ImageResource.h): one image + its handle, released with the object (the
// reference's RAII deleters, Src/Allocator.cpp:132-134).
class ImageResource {
public:
  ImageResource() = default;
  ImageResource(const CudaApplication& app, uint32_t vkFormat, uint32_t w, uint32_t h, uint32_t mips = 1, uint32_t layers = 1)
      : _app(&app), _format(vkFormat), _w(w), _h(h), _mips(mips), _layers(layers) {
    app.check(althea_cuda_create_image(app.ctx(), vkFormat, w, h, mips, layers, &_handle), "althea_cuda_create_image");
  }
  // wraps memory imported from Vulkan: takes ownership of the handle
  ImageResource(const CudaApplication& app, uint64_t importedHandle, uint32_t vkFormat, uint32_t w, uint32_t h, uint32_t mips, uint32_t layers)
      : _app(&app), _handle(importedHandle), _format(vkFormat), _w(w), _h(h), _mips(mips), _layers(layers) {}
  ~ImageResource() { reset(); }
  ImageResource(ImageResource&& o) noexcept { *this = std::move(o); }
  ImageResource& operator=(ImageResource&& o) noexcept {
    if (this != &o) {
      reset();
      _app = o._app; _handle = o._handle; _format = o._format; _w = o._w; _h = o._h; _mips = o._mips; _layers = o._layers;
      o._handle = 0;
    }
    return *this;
  }
  ImageResource(const ImageResource&) = delete;
  ImageResource& operator=(const ImageResource&) = delete;

  uint64_t handle() const { return _handle; }
  uint32_t width() const { return _w; }
  uint32_t height() const { return _h; }
  uint32_t mipCount() const { return _mips; }
  size_t byteSize() const { return althea_cuda_image_bytes(_format, _w, _h, _mips, _layers); }
  void upload(const void* host, size_t bytes) const { _app->check(althea_cuda_upload(_app->ctx(), _handle, host, bytes, nullptr), "althea_cuda_upload"); }
  void download(void* host, size_t bytes) const { _app->check(althea_cuda_download(_app->ctx(), _handle, host, bytes, nullptr), "althea_cuda_download"); }

private:
  void reset() {
    if (_handle && _app) althea_cuda_release(_app->ctx(), _handle);
    _handle = 0;
  }
  const CudaApplication* _app = nullptr;
  uint64_t _handle = 0;
  uint32_t _format = 0, _w = 0, _h = 0, _mips = 1, _layers = 1;
};

} // namespace AltheaEngine
