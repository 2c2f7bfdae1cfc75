// Include/Althea/DeferredRendering.h:38-102 (GBufferResources) and the application-owned deferred lighting pass
// (Shaders/DeferredPass.vert/.frag) as a class with one draw() call.
#pragma once
#include "CudaApplication.h"
#include "GlobalUniforms.h"
#include "ImageBasedLighting.h"
#include "Model.h"
#include "PointLight.h"
#include "ScreenSpaceReflection.h"

namespace AltheaEngine {

class GBufferResources {
public:
  GBufferResources() = default;
  // formats of Src/DeferredRendering.cpp:42-99, plus the legacy RGBA32F position target DeferredPass.frag:18 reads
  GBufferResources(const CudaApplication& app, uint32_t width, uint32_t height)
      : _depthA(app, ALTHEA_FORMAT_R32_SFLOAT, width, height), _position(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, width, height),
        _normal(app, ALTHEA_FORMAT_R16G16B16A16_SFLOAT, width, height), _albedo(app, ALTHEA_FORMAT_R8G8B8A8_UNORM, width, height),
        _metallicRoughnessOcclusion(app, ALTHEA_FORMAT_R8G8B8A8_UNORM, width, height) {}
  ImageResource& getDepthA() { return _depthA; }
  ImageResource& getPosition() { return _position; }
  ImageResource& getNormal() { return _normal; }
  ImageResource& getAlbedo() { return _albedo; }
  ImageResource& getMetallicRoughnessOcclusion() { return _metallicRoughnessOcclusion; }
  althea_gbuffer getHandles() const {
    return althea_gbuffer{_depthA.handle(), _position.handle(), _normal.handle(), _albedo.handle(), _metallicRoughnessOcclusion.handle()};
  }

private:
  ImageResource _depthA, _position, _normal, _albedo, _metallicRoughnessOcclusion;
};

// SceneToGBufferPass (Include/Althea/DeferredRendering.h:132-155, Src/DeferredRendering.cpp:268-330) with its glTF subpass
// (Gltf.vert/.frag): rasterises the models' primitives into the G-buffer attachments.
class SceneToGBufferPass {
public:
  SceneToGBufferPass() = default;
  explicit SceneToGBufferPass(const CudaApplication& app) : _app(&app) {}
  void draw(const GlobalUniforms& globals, const std::vector<Model>& models, const GBufferResources& gBuffer, const althea_sync* sync = nullptr) {
    const std::vector<althea_primitive> prims = describeModels(models);
    const althea_gbuffer gb = gBuffer.getHandles();
    _app->check(althea_cuda_draw_gbuffer(_app->ctx(), &globals, prims.data(), (uint32_t)prims.size(), &gb, sync), "althea_cuda_draw_gbuffer");
  }

private:
  const CudaApplication* _app = nullptr;
};

class DeferredPass {
public:
  DeferredPass() = default;
  DeferredPass(const CudaApplication& app, uint32_t width, uint32_t height, uint32_t colorFormat = ALTHEA_FORMAT_R16G16B16A16_SFLOAT)
      : _app(&app), _color(app, colorFormat, width, height) {}
  // flags: ALTHEA_SHADE_SKIP_TONEMAP is the reference's SKIP_TONEMAP shader define (DeferredPass.frag:48-50,86-88)
  void draw(const GlobalUniforms& globals, const GBufferResources& gBuffer, const IBLResources& ibl, const PointLightCollection* lights,
            const ScreenSpaceReflection& ssr, uint32_t flags = 0, const althea_sync* sync = nullptr) {
    const althea_gbuffer gb = gBuffer.getHandles();
    const althea_ibl ib = ibl.getHandles();
    _app->check(althea_cuda_deferred_shade(_app->ctx(), &globals, &gb, &ib, lights ? lights->bufferHandle() : 0,
                                           lights ? lights->shadowMapHandle() : 0, ssr.getReflectionBuffer().getHandle(), _color.handle(),
                                           /*ao_counts*/ 0, flags, sync),
                "althea_cuda_deferred_shade");
  }
  ImageResource& getColorTarget() { return _color; }

private:
  const CudaApplication* _app = nullptr;
  ImageResource _color;
};

} // namespace AltheaEngine
