// One deferred frame through the C++ mirror of the reference's interface (althea_b200/host/Althea/*.h):
//   demo_frame <inputs.bin> <outputs.bin> [--parity]
// inputs.bin is written by tests/test_host_cpp.py (layout below); outputs.bin receives the reflection mip chain (RGBA16F)
// followed by the lit colour target (RGBA32F). The order of calls is the reference's per-frame order.
#include <Althea/DeferredRendering.h>
#include <Althea/ImageBasedLighting.h>
#include <Althea/PointLight.h>
#include <Althea/ScreenSpaceReflection.h>

#include <cstdio>
#include <cstring>
#include <fstream>

using namespace AltheaEngine;

namespace {
struct Header { int32_t W, H, nLights, shadowRes, envW, envH, preW, preH, preMips, irrW, irrH, lutW, lutH; };
std::vector<char> readBlock(std::ifstream& f, size_t n) {
  std::vector<char> v(n);
  f.read(v.data(), (std::streamsize)n);
  if (!f) throw std::runtime_error("inputs file truncated");
  return v;
}
} // namespace

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: demo_frame inputs.bin outputs.bin [--parity]\n"); return 2; }
  const bool parity = argc > 3 && std::strcmp(argv[3], "--parity") == 0;
  try {
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) throw std::runtime_error("cannot open inputs");
    Header h;
    f.read(reinterpret_cast<char*>(&h), sizeof h);
    GlobalUniforms globals;
    f.read(reinterpret_cast<char*>(&globals), sizeof globals);
    const size_t px = (size_t)h.W * h.H;

    CudaApplication app(0, nullptr, parity ? ALTHEA_CTX_PARITY_MATH : 0);
    GBufferResources gBuffer(app, h.W, h.H);
    gBuffer.getPosition().upload(readBlock(f, px * 16).data(), px * 16);
    gBuffer.getDepthA().upload(readBlock(f, px * 4).data(), px * 4);
    gBuffer.getNormal().upload(readBlock(f, px * 8).data(), px * 8);
    gBuffer.getAlbedo().upload(readBlock(f, px * 4).data(), px * 4);
    gBuffer.getMetallicRoughnessOcclusion().upload(readBlock(f, px * 4).data(), px * 4);

    PointLightCollection lights(app, (size_t)h.nLights, h.nLights > 0, (uint32_t)(h.shadowRes > 0 ? h.shadowRes : 1));
    if (h.nLights > 0) {
      std::vector<char> lb = readBlock(f, (size_t)h.nLights * sizeof(PointLight));
      for (int i = 0; i < h.nLights; ++i) {
        PointLight l;
        std::memcpy(&l, lb.data() + (size_t)i * sizeof(PointLight), sizeof l);
        lights.setLight((uint32_t)i, l);
      }
      lights.updateResource();
      const size_t cubeFloats = (size_t)h.nLights * 6 * h.shadowRes * h.shadowRes;
      std::vector<char> sb = readBlock(f, cubeFloats * 4);
      lights.uploadShadowMaps(reinterpret_cast<const float*>(sb.data()), cubeFloats);
    }

    // IBL maps are inputs of this demo (the precompute has its own entry point, ImageBasedLighting::createResources)
    IBLResources ibl;
    ibl.environmentMap = ImageResource(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, h.envW, h.envH);
    ibl.prefilteredMap = ImageResource(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, h.preW, h.preH, h.preMips);
    ibl.irradianceMap = ImageResource(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, h.irrW, h.irrH);
    ibl.brdfLut = ImageResource(app, ALTHEA_FORMAT_R8G8B8A8_UNORM, h.lutW, h.lutH);
    for (ImageResource* img : {&ibl.environmentMap, &ibl.prefilteredMap, &ibl.irradianceMap, &ibl.brdfLut})
      img->upload(readBlock(f, img->byteSize()).data(), img->byteSize());

    ScreenSpaceReflection ssr(app, h.W, h.H);
    DeferredPass deferred(app, h.W, h.H, ALTHEA_FORMAT_R32G32B32A32_SFLOAT);
    const PointLightCollection* pl = h.nLights > 0 ? &lights : nullptr;

    ssr.captureReflection(globals, gBuffer.getHandles(), ibl, pl);
    ssr.convolveReflectionBuffer();
    deferred.draw(globals, gBuffer, ibl, pl, ssr, ALTHEA_SHADE_SKIP_TONEMAP);
    app.waitIdle();

    std::vector<char> refl(ssr.getReflectionBuffer().getImage().byteSize()), color(deferred.getColorTarget().byteSize());
    ssr.getReflectionBuffer().getImage().download(refl.data(), refl.size());
    deferred.getColorTarget().download(color.data(), color.size());
    app.waitIdle();
    std::ofstream o(argv[2], std::ios::binary);
    o.write(refl.data(), (std::streamsize)refl.size());
    o.write(color.data(), (std::streamsize)color.size());
    std::printf("demo_frame ok: %dx%d, %d lights, %llu kernel launches\n", h.W, h.H, h.nLights, (unsigned long long)app.launchCount());

    // error behaviour: failures surface as std::runtime_error, as in the reference
    try {
      ReflectionBuffer small(app, 8, 8);
      ScreenSpaceReflection bad(app, 8, 8);
      bad.captureReflection(globals, gBuffer.getHandles(), ibl, pl);
      std::printf("demo_frame ERROR: size mismatch was not rejected\n");
      return 1;
    } catch (const std::runtime_error& e) {
      std::printf("expected failure: %s\n", e.what());
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "demo_frame failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
