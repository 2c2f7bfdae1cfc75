// Include/Althea/ImageBasedLighting.h:17-54: IBLResources + namespace ImageBasedLighting::createResources.
#pragma once
#include "CudaApplication.h"
#include "Utilities.h"

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

namespace AltheaEngine {

struct IBLResources {
  ImageResource environmentMap{};
  ImageResource prefilteredMap{};
  ImageResource irradianceMap{};
  ImageResource brdfLut{};
  althea_ibl getHandles() const { return althea_ibl{environmentMap.handle(), prefilteredMap.handle(), irradianceMap.handle(), brdfLut.handle()}; }
};

namespace ImageBasedLighting {
// Src/ImageBasedLighting.cpp:415-605 with the file IO lifted out: the caller hands over the decoded equirect env map
// (RGBA32F, as Utilities::loadHdri returns it) and optionally the BRDF LUT asset (RGBA8; generated when null).
// Shapes follow the reference: irradiance W x H, prefiltered (W/2 x H/2, 5 mips, roughness i/4), 10000 samples per texel.
inline IBLResources createResources(const CudaApplication& app, const float* envRgba, uint32_t W, uint32_t H, const uint8_t* brdfLutRgba8 = nullptr,
                                    uint32_t lutSize = 512, const althea_ibl_precompute_desc* desc = nullptr) {
  IBLResources r;
  const uint32_t mips = 1u + (uint32_t)std::floor(std::log2((double)(W > H ? W : H))); // Utilities.cpp:97-100
  ImageResource chain(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, W, H, mips);
  chain.upload(envRgba, (size_t)W * H * 16);
  app.check(althea_cuda_generate_mips(app.ctx(), chain.handle(), nullptr), "althea_cuda_generate_mips"); // Image.cpp:135-239
  r.environmentMap = ImageResource(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, W, H);
  r.environmentMap.upload(envRgba, (size_t)W * H * 16);
  r.irradianceMap = ImageResource(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, W, H);
  r.prefilteredMap = ImageResource(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, W / 2, H / 2, 5);
  althea_ibl_precompute_desc d{ALTHEA_IBL_LAYOUT_EQUIRECT, ALTHEA_IBL_SEQ_REFERENCE_HASH, 0, 0};
  if (desc) d = *desc;
  app.check(althea_cuda_ibl_precompute(app.ctx(), chain.handle(), &d, r.irradianceMap.handle(), r.prefilteredMap.handle(), nullptr),
            "althea_cuda_ibl_precompute");
  if (brdfLutRgba8) {
    r.brdfLut = ImageResource(app, ALTHEA_FORMAT_R8G8B8A8_UNORM, lutSize, lutSize);
    r.brdfLut.upload(brdfLutRgba8, (size_t)lutSize * lutSize * 4);
  } else {
    r.brdfLut = ImageResource(app, ALTHEA_FORMAT_R8G8B8A8_UNORM, lutSize, lutSize);
    app.check(althea_cuda_brdf_lut(app.ctx(), 1024, r.brdfLut.handle(), nullptr), "althea_cuda_brdf_lut");
  }
  app.waitIdle(); // `chain` is released on return
  return r;
}

inline bool fileExists(const std::string& path) {
  if (FILE* f = std::fopen(path.c_str(), "rb")) {
    std::fclose(f);
    return true;
  }
  return false;
}

// createResources(app, commandBuffer, envMapName) as the reference runs it (Src/ImageBasedLighting.cpp:415-605):
// <contentDir>/HDRI_Skybox/<envMapName>.hdr must exist; when any of PrecomputedMaps/<envMapName>/IrradianceMap.hdr,
// Prefiltered1..5.hdr is missing the maps are computed and saved (Utilities::saveHdri); then every map is loaded back from
// its file (Utilities::loadHdri), so the run-time data carries the files' RGBE quantisation on a cache hit and on a miss.
// The BRDF LUT asset is a PNG in the reference; this mirror has no PNG decoder, the caller passes the decoded texels or
// gets the generated table.
inline IBLResources createResources(const CudaApplication& app, const std::string& contentDir, const std::string& envMapName,
                                    const uint8_t* brdfLutRgba8 = nullptr, uint32_t lutSize = 512) {
  const std::string envFile = contentDir + "/HDRI_Skybox/" + envMapName + ".hdr";
  if (!fileExists(envFile)) throw std::runtime_error("Specified environment map does not exist!");
  const std::string dir = contentDir + "/PrecomputedMaps/" + envMapName + "/";
  std::string prefiltered[5];
  bool needToPrecomputeIBL = false;
  for (int i = 0; i < 5; ++i) {
    prefiltered[i] = dir + "Prefiltered" + std::to_string(i + 1) + ".hdr";
    needToPrecomputeIBL |= !fileExists(prefiltered[i]);
  }
  const std::string irradiance = dir + "IrradianceMap.hdr";
  needToPrecomputeIBL |= !fileExists(irradiance);

  Utilities::ImageFile env;
  Utilities::loadHdri(envFile, env);
  const uint32_t W = (uint32_t)env.width, H = (uint32_t)env.height;
  if (needToPrecomputeIBL) { // the directory must exist, as in the reference
    IBLResources fresh = createResources(app, reinterpret_cast<const float*>(env.data.data()), W, H, nullptr, 16);
    std::vector<std::byte> texels((size_t)W * H * 16);
    fresh.irradianceMap.download(texels.data(), texels.size());
    app.waitIdle();
    Utilities::saveHdri(irradiance, (int)W, (int)H, texels.data(), texels.size());
    std::vector<std::byte> chain(fresh.prefilteredMap.byteSize());
    fresh.prefilteredMap.download(chain.data(), chain.size());
    app.waitIdle();
    size_t offset = 0;
    for (int i = 0; i < 5; ++i) {
      const uint32_t w = (W / 2) >> i ? (W / 2) >> i : 1, h = (H / 2) >> i ? (H / 2) >> i : 1;
      Utilities::saveHdri(prefiltered[i], (int)w, (int)h, chain.data() + offset, (size_t)w * h * 16);
      offset += (size_t)w * h * 16;
    }
  }

  IBLResources r;
  r.environmentMap = ImageResource(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, W, H);
  r.environmentMap.upload(env.data.data(), env.data.size());
  Utilities::ImageFile irr;
  Utilities::loadHdri(irradiance, irr);
  r.irradianceMap = ImageResource(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, (uint32_t)irr.width, (uint32_t)irr.height);
  r.irradianceMap.upload(irr.data.data(), irr.data.size());
  std::vector<std::byte> levels;
  uint32_t w0 = 0, h0 = 0;
  for (int i = 0; i < 5; ++i) {
    Utilities::ImageFile level;
    Utilities::loadHdri(prefiltered[i], level);
    if (i == 0) { w0 = (uint32_t)level.width; h0 = (uint32_t)level.height; }
    levels.insert(levels.end(), level.data.begin(), level.data.end());
  }
  r.prefilteredMap = ImageResource(app, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, w0, h0, 5);
  if (levels.size() != r.prefilteredMap.byteSize()) throw std::runtime_error("Prefiltered1..5.hdr do not form a 5-level mip chain");
  r.prefilteredMap.upload(levels.data(), levels.size());
  r.brdfLut = ImageResource(app, ALTHEA_FORMAT_R8G8B8A8_UNORM, lutSize, lutSize);
  if (brdfLutRgba8)
    r.brdfLut.upload(brdfLutRgba8, (size_t)lutSize * lutSize * 4);
  else
    app.check(althea_cuda_brdf_lut(app.ctx(), 1024, r.brdfLut.handle(), nullptr), "althea_cuda_brdf_lut");
  app.waitIdle();
  return r;
}
} // namespace ImageBasedLighting
} // namespace AltheaEngine
