// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h).
//
// IBL precompute of Althea, restated on the CPU. This half of the oracle IS pinned: the reference ships the
// outputs of exactly this code (Content/PrecomputedMaps/<env>/{IrradianceMap,Prefiltered1..5}.hdr made from
// Content/HDRI_Skybox/<env>.hdr) and tests/test_oracle_pin.py checks the restatement against them, and the
// BRDF LUT against Content/PrecomputedMaps/brdf_lut.png.
//
//   oracle_env_mip_chain   <- Src/Image.cpp:135-239 (LINEAR 2:1 blit chain), Src/Utilities.cpp:97-100 (level count)
//   oracle_ibl_irradiance  <- Shaders/IBL_Precompute/GenIrradianceMap.comp:19-31,78-102,106-155; host ImageBasedLighting.cpp:328-343
//   oracle_ibl_prefilter   <- Shaders/IBL_Precompute/PreFilterEnvMap.comp:36-42,78-92,98-122,126-177; host :376-395
//   oracle_brdf_lut        <- no generator in the reference; the standard split-sum integral its asset brdf_lut.png
//                             holds (consumer: Shaders/PBR/PBRMaterial.glsl:110; loader ImageBasedLighting.cpp:570-602)
// Two layouts: LAYOUT_EQUIRECT is the reference's own (output texel -> (yaw,pitch)); LAYOUT_CUBE is the BASELINE.json
// config-2 extension (output texel -> Vulkan cube-face direction) with the identical integrand.
// Two sample sequences for the prefilter: SEQ_HASH is the reference's per-texel hash RNG; SEQ_HAMMERSLEY is the
// BASELINE.json extension.
#include "oracle_math.h"
#include <omp.h>

using namespace oracle;

namespace {

enum { LAYOUT_EQUIRECT = 0, LAYOUT_CUBE = 1 };
enum { SEQ_HASH = 0, SEQ_HAMMERSLEY = 1 };

// GenIrradianceMap.comp:78-102 == PreFilterEnvMap.comp:98-122; sampler REPEAT + linear mips (ImageBasedLighting.cpp:168-174)
V3 sampleEnvMapPrecompute(const TexChain& env, V3 dir, float mip) {
  float pitch = 0.0f, yaw = 0.0f;
  float lenXz = sqrtf(dir.x * dir.x + dir.z * dir.z);
  if (lenXz > 0.001f) { yaw = atan2f(dir.z, dir.x); pitch = atanf(dir.y / lenXz); }
  else if (dir.y > 0.0f) pitch = 0.5f * kPi;
  else pitch = -0.5f * kPi;
  float u = yaw / (2.0f * kPi) + 0.5f, v = pitch / kPi + 0.5f;
  return xyz(trilinear(env, u, v, mip, ADDR_REPEAT));
}

// output texel -> surface normal
V3 texelNormal(int layout, int x, int y, int face, int outW, int outH) {
  if (layout == LAYOUT_EQUIRECT) { // GenIrradianceMap.comp:108-114, PreFilterEnvMap.comp:128-133
    float u = (float)x / (float)outW, v = (float)y / (float)outH;
    float yaw = kPi * (2.0f * u - 1.0f);
    float pitch = kPi * (v - 0.5f);
    return {cosf(pitch) * cosf(yaw), sinf(pitch), cosf(pitch) * sinf(yaw)};
  }
  // cube: inverse of the Vulkan face table (rule A8), texel centres
  float sc = 2.0f * (((float)x + 0.5f) / (float)outW) - 1.0f;
  float tc = 2.0f * (((float)y + 0.5f) / (float)outH) - 1.0f;
  V3 d;
  switch (face) {
  case 0: d = {1.0f, -tc, -sc}; break;
  case 1: d = {-1.0f, -tc, sc}; break;
  case 2: d = {sc, 1.0f, tc}; break;
  case 3: d = {sc, -1.0f, -tc}; break;
  case 4: d = {sc, -tc, 1.0f}; break;
  default: d = {-sc, -tc, -1.0f}; break;
  }
  return normalize(d);
}

float radicalInverse2(uint32_t bits) {
  bits = (bits << 16) | (bits >> 16);
  bits = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
  bits = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
  bits = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
  bits = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
  return (float)bits * 2.3283064365386963e-10f;
}

} // namespace

extern "C" {

int oracle_mip_count(int w, int h) { int m = w > h ? w : h; int n = 1; while (m > 1) { m >>= 1; ++n; } return n; }
size_t oracle_chain_texels(int w, int h, int mips) {
  size_t n = 0; for (int i = 0; i < mips; ++i) { n += (size_t)w * h; w = w > 1 ? w >> 1 : 1; h = h > 1 ? h >> 1 : 1; } return n;
}

// level 0 = env, level k+1 = LINEAR blit of level k to half size (Image.cpp:183-213): for even sizes the 2x2 box
void oracle_env_mip_chain(const float* env, int W, int H, int mips, float* chain) {
  memcpy(chain, env, (size_t)W * H * 16);
  TexChain ch{chain, W, H, mips, FMT_RGBA32F};
  for (int k = 1; k < mips; ++k) {
    Tex src = ch.level(k - 1), dst = ch.level(k);
    float* o = (float*)dst.data;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst.h; ++y)
      for (int x = 0; x < dst.w; ++x) {
        V4 c = bilinear(src, ((float)x + 0.5f) / (float)dst.w, ((float)y + 0.5f) / (float)dst.h, ADDR_CLAMP);
        float* p = o + ((size_t)y * dst.w + x) * 4;
        p[0] = c.x; p[1] = c.y; p[2] = c.z; p[3] = c.w;
      }
  }
}

// texels: n triplets (x, y, face). thetaSamples = 300 in the reference. Output n x RGBA.
void oracle_ibl_irradiance(const float* chain, int W, int H, int mips, int layout, int outW, int outH,
                           const int32_t* texels, int n, int thetaSamples, float* out) {
  TexChain env{chain, W, H, mips, FMT_RGBA32F};
  const int phiSamples = (int)((float)H * (float)thetaSamples / (float)W); // pushConstants.height * thetaSamples / width
  const float mip = log2f((float)W / (float)thetaSamples);
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < n; ++t) {
    V3 nor = texelNormal(layout, texels[3 * t], texels[3 * t + 1], texels[3 * t + 2], outW, outH);
    Frame tbn = localToWorld(nor);
    V3 irradiance{0.0f, 0.0f, 0.0f};
    for (int i = 0; i < thetaSamples; ++i) {
      float theta = (float)(i * 2) * kPi / (float)thetaSamples;
      float cosTheta = cosf(theta), sinTheta = sinf(theta);
      for (int j = 0; j < phiSamples; ++j) {
        float phi = (float)j * 0.5f * kPi / (float)phiSamples;
        float cosPhi = cosf(phi), sinPhi = sinf(phi);
        V3 sampleDir = tbn.apply(V3{cosTheta * sinPhi, sinTheta * sinPhi, cosPhi});
        irradiance = irradiance + (sampleEnvMapPrecompute(env, sampleDir, mip) * cosPhi) * sinPhi;
      }
    }
    V3 c = ((kPi * irradiance) / (float)thetaSamples) / (float)phiSamples;
    out[4 * t] = c.x; out[4 * t + 1] = c.y; out[4 * t + 2] = c.z; out[4 * t + 3] = 1.0f;
  }
}

// roughness is used directly as alpha (PreFilterEnvMap.comp:139-140). numSamples = 10000 in the reference.
void oracle_ibl_prefilter(const float* chain, int W, int H, int mips, int layout, int outW, int outH, float roughness,
                          int numSamples, int sequence, const int32_t* texels, int n, float* out) {
  TexChain env{chain, W, H, mips, FMT_RGBA32F};
  const float a2 = roughness * roughness;
  const float saTexel = 4.0f * kPi / (6.0f * (float)W * (float)H);
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < n; ++t) {
    const int x = texels[3 * t], y = texels[3 * t + 1], face = texels[3 * t + 2];
    Rng rng{(uint32_t)x, (uint32_t)y}; // seed = uvec2(gl_GlobalInvocationID.xy)
    V3 N = texelNormal(layout, x, y, face, outW, outH);
    V3 V = N;
    Frame tbn = localToWorld(N);
    V3 prefiltered{0.0f, 0.0f, 0.0f};
    float totalWeight = 0.0f;
    for (int i = 0; i < numSamples; ++i) {
      float xi0, xi1;
      if (sequence == SEQ_HASH) { xi0 = rng.next(); xi1 = rng.next(); }
      else { xi0 = (float)i / (float)numSamples; xi1 = radicalInverse2((uint32_t)i); }
      // sampleGGX (:84-92)
      float phi = 2.0f * kPi * xi0;
      float cosTheta = sqrtf((1.0f - xi1) / (1.0f + (a2 - 1.0f) * xi1));
      float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
      V3 H = tbn.apply(V3{cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta});
      V3 L = normalize((2.0f * dot(V, H)) * H - V);
      float NdotL = maxf(dot(N, L), 0.0f);
      float NdotH = maxf(dot(N, H), 0.0f);
      float HdotV = maxf(dot(H, V), 0.0f);
      float denom = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
      float D = a2 / (kPi * denom * denom);
      float pdf = D * NdotH / (4.0f * HdotV + 0.00001f);
      float saSample = 1.0f / ((float)numSamples * pdf + 0.0001f);
      float mipLevel = roughness == 0.0f ? 0.0f : 0.5f * log2f(saSample / saTexel);
      if (NdotL > 0.0f) {
        prefiltered = prefiltered + sampleEnvMapPrecompute(env, L, mipLevel) * NdotL;
        totalWeight += NdotL;
      }
    }
    V3 c = prefiltered / totalWeight;
    out[4 * t] = c.x; out[4 * t + 1] = c.y; out[4 * t + 2] = c.z; out[4 * t + 3] = 1.0f;
  }
}

// Split-sum BRDF LUT, texel (x,y) -> NdotV = (x+.5)/size, roughness = (y+.5)/size. kMode 0: k = roughness^2/2
// (the common IBL convention); kMode 1: k = roughness^4/2 (PBRMaterial.glsl:96 kIbl). out: size*size*(A,B), row y.
void oracle_brdf_lut(int size, int samples, int kMode, float* out) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < size; ++y)
    for (int x = 0; x < size; ++x) {
      float NdotV = ((float)x + 0.5f) / (float)size, roughness = ((float)y + 0.5f) / (float)size;
      V3 V{sqrtf(1.0f - NdotV * NdotV), 0.0f, NdotV};
      float a = roughness * roughness;
      float a2 = a * a;
      float k = (kMode == 0 ? a : a2) / 2.0f;
      float A = 0.0f, B = 0.0f;
      for (int i = 0; i < samples; ++i) {
        float xi0 = (float)i / (float)samples, xi1 = radicalInverse2((uint32_t)i);
        float phi = 2.0f * kPi * xi0;
        float cosTheta = sqrtf((1.0f - xi1) / (1.0f + (a2 - 1.0f) * xi1));
        float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
        V3 H{cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta};
        V3 L = (2.0f * dot(V, H)) * H - V;
        float NdotL = maxf(L.z, 0.0f), NdotH = maxf(H.z, 0.0f), VdotH = maxf(dot(V, H), 0.0f);
        if (NdotL > 0.0f) {
          float G = (NdotV / (NdotV * (1.0f - k) + k)) * (NdotL / (NdotL * (1.0f - k) + k));
          float GVis = (G * VdotH) / (NdotH * NdotV);
          float Fc = powf(1.0f - VdotH, 5.0f);
          A += (1.0f - Fc) * GVis;
          B += Fc * GVis;
        }
      }
      out[((size_t)y * size + x) * 2] = A / (float)samples;
      out[((size_t)y * size + x) * 2 + 1] = B / (float)samples;
    }
}

} // extern "C"
