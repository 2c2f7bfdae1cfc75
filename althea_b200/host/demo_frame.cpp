// One deferred frame through the C++ mirror of the reference's interface (althea_b200/host/Althea/*.h):
//   demo_frame <inputs.bin> <outputs.bin> [--parity]
//   demo_frame --raster <scene.bin> <outputs.bin>      the rasterising producers: shadow cubes + G-buffer from meshes
// inputs.bin is written by tests/test_host_cpp.py (layout below); outputs.bin receives the reflection mip chain (RGBA16F)
// followed by the lit colour target (RGBA32F). The order of calls is the reference's per-frame order.
#include <Althea/DeferredRendering.h>
#include <Althea/ImageBasedLighting.h>
#include <Althea/Model.h>
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

// scene.bin (tests/test_host_cpp.py): int32 W, H, nLights, shadowRes, nPrims; GlobalUniforms; PointLight[nLights]; per primitive:
// int32 nVerts, nIdx, frontCW, hasBaseTexture; float model[16]; float factors[8] (baseColorFactor, normalScale, metallic,
// roughness, alphaCutoff); Vertex[nVerts]; uint32[nIdx]; if hasBaseTexture: int32 w, h, mips, sampler; RGBA8 texels of all levels.
// outputs: depth, position, normal, albedo, MRO attachments, then the shadow cube array.
static int rasterDemo(const char* in, const char* out) {
  std::ifstream f(in, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open scene");
  int32_t hd[5];
  f.read(reinterpret_cast<char*>(hd), sizeof hd);
  const int W = hd[0], H = hd[1], nLights = hd[2], shadowRes = hd[3], nPrims = hd[4];
  GlobalUniforms globals;
  f.read(reinterpret_cast<char*>(&globals), sizeof globals);
  CudaApplication app(0);
  PointLightCollection lights(app, (size_t)nLights, nLights > 0, (uint32_t)shadowRes);
  for (int i = 0; i < nLights; ++i) {
    PointLight l;
    f.read(reinterpret_cast<char*>(&l), sizeof l);
    lights.setLight((uint32_t)i, l);
  }
  std::vector<Model> models(1);
  for (int p = 0; p < nPrims; ++p) {
    int32_t ph[4];
    f.read(reinterpret_cast<char*>(ph), sizeof ph);
    float model[16], factors[8];
    f.read(reinterpret_cast<char*>(model), sizeof model);
    f.read(reinterpret_cast<char*>(factors), sizeof factors);
    std::vector<Vertex> verts((size_t)ph[0]);
    std::vector<uint32_t> idx((size_t)ph[1]);
    f.read(reinterpret_cast<char*>(verts.data()), (std::streamsize)(verts.size() * sizeof(Vertex)));
    f.read(reinterpret_cast<char*>(idx.data()), (std::streamsize)(idx.size() * sizeof(uint32_t)));
    Material m;
    std::memcpy(m.baseColorFactor, factors, 16);
    m.normalScale = factors[4]; m.metallicFactor = factors[5]; m.roughnessFactor = factors[6]; m.alphaCutoff = factors[7];
    if (ph[3]) {
      int32_t th[4];
      f.read(reinterpret_cast<char*>(th), sizeof th);
      const size_t bytes = althea_cuda_image_bytes(ALTHEA_FORMAT_R8G8B8A8_UNORM, (uint32_t)th[0], (uint32_t)th[1], (uint32_t)th[2], 1);
      std::vector<char> texels = readBlock(f, bytes);
      m.baseTexture = std::make_shared<Texture>(app, texels.data(), (uint32_t)th[0], (uint32_t)th[1], (uint32_t)th[2], (uint32_t)th[3]);
    }
    if (!f) throw std::runtime_error("scene file truncated");
    models[0].addPrimitive(Primitive(app, verts, idx, model, std::move(m), ph[2] != 0));
  }
  GBufferResources gBuffer(app, (uint32_t)W, (uint32_t)H);
  SceneToGBufferPass gpass(app);
  lights.drawShadowMaps(models);
  gpass.draw(globals, models, gBuffer);  // step 5
  app.waitIdle();
  std::ofstream o(out, std::ios::binary);
  for (ImageResource* img : {&gBuffer.getDepthA(), &gBuffer.getPosition(), &gBuffer.getNormal(), &gBuffer.getAlbedo(), &gBuffer.getMetallicRoughnessOcclusion()}) {
    std::vector<char> b(img->byteSize());
    img->download(b.data(), b.size());
    app.waitIdle();
    o.write(b.data(), (std::streamsize)b.size());
  }
  if (nLights > 0) {
    std::vector<char> b((size_t)nLights * 6 * shadowRes * shadowRes * 4);
    app.check(althea_cuda_download(app.ctx(), lights.shadowMapHandle(), b.data(), b.size(), nullptr), "althea_cuda_download(shadow cubes)");
    app.waitIdle();
    o.write(b.data(), (std::streamsize)b.size());
  }
  std::printf("demo_frame raster ok: %dx%d, %d primitives, %d lights, %llu kernel launches\n", W, H, nPrims, nLights, (unsigned long long)app.launchCount());
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 4 && std::strcmp(argv[1], "--raster") == 0) {
    try {
      return rasterDemo(argv[2], argv[3]);
    } catch (const std::exception& e) {
      std::fprintf(stderr, "demo_frame failed: %s\n", e.what());
      return 1;
    }
  }
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
