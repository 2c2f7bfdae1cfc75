// Include/Althea/ImageBasedLighting.h:17-54: IBLResources + namespace ImageBasedLighting::createResources.
#pragma once
#include "CudaApplication.h"

#include <cmath>

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
} // namespace ImageBasedLighting
} // namespace AltheaEngine
