// ORACLE — TEST INFRASTRUCTURE ONLY. Runs the REFERENCE'S OWN SHADER TEXT on the CPU: oracle/_ref/libshader_ref.so.
//
// The reference's per-frame path is GLSL that a Vulkan driver compiles at run time; this image has neither. What it does have
// is g++, and GLSL's expression language is (nearly) C++'s: `make -C oracle ref` lets oracle/glsl2cpp.py rewrite each shader,
// read where it lies under /root/reference/Shaders, into oracle/_ref/gen/*.inc (declarations and literals only, see its header; removed again once the library is linked),
// and this file includes those bodies into one struct per shader stage, with oracle/glsl_compat.h standing in for the GPU
// (types, built-ins, texture unit). The entry points below take the arguments of liboracle.so's, so a test hands both the same
// frame and compares: the restatement (oracle/althea_oracle.cpp) against the text it restates, executed.
//
// What this does and does not pin. It pins the READING of the shaders: control flow, operand order, which map at which LOD,
// every constant. It does not pin what GLSL leaves to the implementation (the rounding order inside dot / normalize / matrix
// products, the texture filter's arithmetic): those are oracle_math.h's in both, on purpose. The bit-rot of SURVEY.md 8(c-bis)
// is handled by the patches spelled out in oracle/Makefile (DeferredPass.frag's stale pbrMaterial call and missing
// declarations) and by the one overload SSR.frag calls but nobody defines, added below.
#include "glsl_compat.h"
#include <omp.h>
#include <vector>

extern "C" {
// the argument blocks of liboracle.so (oracle/althea_oracle.cpp), same layouts
struct OracleGlobalUniforms { float m[6][16]; float mouseUV[2]; int32_t lightCount; uint32_t lightBufferHandle; float time, exposure; uint32_t inputMask, frameCount; };
struct OracleGBuffer { int32_t W, H; const float* position; const float* depth; const uint16_t* normal; const uint8_t* albedo; const uint8_t* mro; };
struct OracleIBL { const float* env; int32_t envW, envH; const float* prefiltered; int32_t preW, preH, preMips; const float* irradiance; int32_t irrW, irrH; const uint8_t* lut; int32_t lutW, lutH; };
struct OracleLights { const float* lights; const float* shadow; int32_t shadowRes; };
}
static_assert(sizeof(OracleGlobalUniforms) == 416, "GlobalUniforms.h:15-31");

namespace glsl {

struct SsrVert : ShaderBase {
#include "_ref/gen/ssr_vert.inc"
};
struct SsrFrag : ShaderBase {
#include "_ref/gen/ssr_frag.inc"
  // SSR.frag:44 calls reconstructPosition(uv, dRaw); the tree defines (uv) and (uv, dRaw, inverseProjection, inverseView) only
  // (SURVEY.md 8c-bis defect 5). The missing overload can only mean the frame's own matrices:
  vec3 reconstructPosition(vec2 uv, float dRaw) {
    return reconstructPosition(uv, dRaw, globalUniforms[pushConstants.globalUniformsHandle].inverseProjection,
                               globalUniforms[pushConstants.globalUniformsHandle].inverseView);
  }
};
struct SsrFragFused : ShaderBase { // the same text with ReconstructPosition.glsl:8 contracted to one fma() (oracle/Makefile)
#include "_ref/gen/ssr_frag_fused.inc"
  vec3 reconstructPosition(vec2 uv, float dRaw) {
    return reconstructPosition(uv, dRaw, globalUniforms[pushConstants.globalUniformsHandle].inverseProjection,
                               globalUniforms[pushConstants.globalUniformsHandle].inverseView);
  }
};
struct DeferredVert : ShaderBase {
#include "_ref/gen/deferred_vert.inc"
};
struct DeferredFrag : ShaderBase {
#include "_ref/gen/deferred_frag.inc"
};
struct DeferredFragLinear : ShaderBase { // -DSKIP_TONEMAP
#include "_ref/gen/deferred_frag_linear.inc"
};
struct ConvolveComp : ShaderBase {
#include "_ref/gen/glossy_convolve_comp.inc"
};
struct IrradianceComp : ShaderBase {
#include "_ref/gen/gen_irradiance_comp.inc"
};
struct PrefilterComp : ShaderBase {
#include "_ref/gen/prefilter_comp.inc"
};

// the rasterising producers: G-buffer pass and omni shadow pass
struct GltfVert : ShaderBase {
#include "_ref/gen/gltf_vert.inc"
};
struct GltfFrag : ShaderBase {
#include "_ref/gen/gltf_frag.inc"
};
struct ShadowVert : ShaderBase {
#include "_ref/gen/shadow_vert.inc"
  MaterialConstants legacyMaterial; // see oracle/Makefile: the fields the legacy shader reads from `PrimitiveConstants` live here today
};
struct ShadowFrag : ShaderBase {
#include "_ref/gen/shadow_frag.inc"
  MaterialConstants legacyMaterial;
};

static_assert(sizeof(SsrFrag::GlobalUniforms) == 416, "the GLSL block is the C++ block (Global/GlobalUniforms.glsl:8-24)");

// texture handles of the bindless heap (any distinct numbers do)
enum { H_ENV, H_PRE, H_IRR, H_LUT, H_DEPTH, H_NORMAL, H_ALBEDO, H_MRO, H_POSITION, H_COUNT };

struct Bound { // one frame's resources in the shapes the shaders index
  sampler2D tex[H_COUNT];
  samplerCubeArray cubes;
  std::vector<SsrFrag::PointLight> lights;
  Bound(const OracleGBuffer& gb, const OracleIBL* ibl, const OracleLights* li, int lightCount) {
    using namespace oracle;
    auto one = [](const void* p, int w, int h, Format f) { sampler2D s; s.chain = TexChain{p, w, h, 1, f}; return s; };
    tex[H_DEPTH] = one(gb.depth, gb.W, gb.H, FMT_R32F);
    tex[H_NORMAL] = one(gb.normal, gb.W, gb.H, FMT_RGBA16F);
    tex[H_ALBEDO] = one(gb.albedo, gb.W, gb.H, FMT_RGBA8);
    tex[H_MRO] = one(gb.mro, gb.W, gb.H, FMT_RGBA8);
    tex[H_POSITION] = one(gb.position, gb.W, gb.H, FMT_RGBA32F);
    if (ibl) { // run-time samplers of the maps: CLAMP_TO_EDGE (Src/ImageBasedLighting.cpp:469-475)
      tex[H_ENV] = one(ibl->env, ibl->envW, ibl->envH, FMT_RGBA32F);
      tex[H_PRE].chain = TexChain{ibl->prefiltered, ibl->preW, ibl->preH, ibl->preMips, FMT_RGBA32F};
      tex[H_IRR] = one(ibl->irradiance, ibl->irrW, ibl->irrH, FMT_RGBA32F);
      tex[H_LUT] = one(ibl->lut, ibl->lutW, ibl->lutH, FMT_RGBA8);
    }
    if (li) {
      cubes.layers = li->shadow; cubes.res = li->shadowRes;
      for (int i = 0; i < lightCount; ++i) { // PointLight.h:31-34: 32-byte records {position, pad, emission, pad}
        SsrFrag::PointLight l;
        l.position = vec3(li->lights[i * 8 + 0], li->lights[i * 8 + 1], li->lights[i * 8 + 2]);
        l.emission = vec3(li->lights[i * 8 + 4], li->lights[i * 8 + 5], li->lights[i * 8 + 6]);
        lights.push_back(l);
      }
    }
  }
};

// direction at a pixel: the rasteriser's interpolation of the three vertices' outputs over the full-screen triangle
// (0,0) (2,0) (0,2) in uv: weights 1 - u/2 - v/2, u/2, v/2
template <class Vert, class Bind> static void vertexDirections(Bind&& bind, double d[3][3]) {
  for (int id = 0; id < 3; ++id) {
    Vert vs;
    bind(vs);
    vs.gl_VertexIndex = id;
    vs.main();
    for (int k = 0; k < 3; ++k) d[id][k] = vs.direction[k];
  }
}
static void announce(int x, int y, int W, int H, float u, float v) { gFrag.u = u; gFrag.v = v; gFrag.x = x; gFrag.y = y; gFrag.W = W; gFrag.H = H; gFrag.on = true; }
static vec3 interpolate(const double d[3][3], double u, double v) {
  const double l1 = 0.5 * u, l2 = 0.5 * v, l0 = 1.0 - l1 - l2;
  return vec3((float)(l0 * d[0][0] + l1 * d[1][0] + l2 * d[2][0]), (float)(l0 * d[0][1] + l1 * d[1][1] + l2 * d[2][1]),
              (float)(l0 * d[0][2] + l1 * d[1][2] + l2 * d[2][2]));
}

} // namespace glsl

using namespace glsl;

extern "C" {

// Misc/ReconstructPosition.glsl:4-22 through SSR.frag's own wrapper chain
void shaderref_reconstruct_position(const OracleGlobalUniforms* g, float u, float v, float dRaw, float* out3) {
  SsrFrag fs;
  SsrFrag::GlobalUniforms gu; memcpy(&gu, g, sizeof gu);
  fs.globalUniforms = &gu;
  fs.pushConstants.globalUniformsHandle = 0;
  const vec3 p = fs.reconstructPosition(vec2(u, v), dRaw);
  out3[0] = p.x; out3[1] = p.y; out3[2] = p.z;
}

// SSR.vert + SSR.frag main, then the blend-on-write onto a (0,0,0,0) clear and the RGBA16F store of the colour attachment
// (Src/ScreenSpaceReflection.cpp:37-44, Src/GraphicsPipeline.cpp:138-154). flags: 1 = the fused reading of ReconstructPosition.glsl:8;
// 2 = the varying evaluated per pixel from SSR.vert's own expression (screenPos := the pixel's uv) instead of interpolated from the
// three vertices: the two choices GLSL / the rasteriser leave open, aligned with the restatement's
} // extern "C"
template <class Frag> static void ssrCapture(const OracleGlobalUniforms* g, const OracleGBuffer* gb, const OracleIBL* ibl, const OracleLights* li,
                                             uint16_t* outReflection, uint8_t* outHit, bool closedForm) {
  const int W = gb->W, H = gb->H;
  Bound b(*gb, ibl, li, g->lightCount);
  typename Frag::GlobalUniforms gu; memcpy(&gu, g, sizeof gu);
  gu.lightBufferHandle = 0;
  typename Frag::GlobalResources res{};
  res.ibl.environmentMapHandle = H_ENV; res.ibl.prefilteredMapHandle = H_PRE; res.ibl.irradianceMapHandle = H_IRR; res.ibl.brdfLutHandle = H_LUT;
  res.gBuffer.depthAHandle = H_DEPTH; res.gBuffer.normalHandle = H_NORMAL; res.gBuffer.albedoHandle = H_ALBEDO;
  res.gBuffer.metallicRoughnessOcclusionHandle = H_MRO;
  res.shadowMapArray = 0;
  typename Frag::POINT_LIGHTS pl;
  pl.pointLightArr = reinterpret_cast<typename Frag::PointLight*>(b.lights.data());
  double vd[3][3];
  SsrVert::GlobalUniforms vgu; memcpy(&vgu, g, sizeof vgu);
  vertexDirections<SsrVert>([&](SsrVert& vs) { vs.globalUniforms = &vgu; vs.pushConstants.globalUniformsHandle = 0; }, vd);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      Frag fs;
      fs.globalUniforms = &gu; fs.globalResources = &res; fs.pointLights = &pl;
      fs.textureHeap = b.tex; fs.cubemapHeap = &b.cubes;
      fs.pushConstants.globalUniformsHandle = 0; fs.pushConstants.globalResourcesHandle = 0;
      const double u = (x + 0.5) / W, v = (y + 0.5) / H;
      fs.inUv = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
      announce(x, y, W, H, fs.inUv.x, fs.inUv.y);
      if (closedForm) { // SSR.vert:16-21 with screenPos := this pixel's uv
        const vec4 pos = vec4(fs.inUv * 2.0f - 1.0f, 0.0f, 1.0f);
        fs.inDirection = mat3(vgu.inverseView) * (vgu.inverseProjection * pos).xyz;
      } else {
        fs.inDirection = interpolate(vd, u, v);
      }
      fs.gl_FragCoord = vec4(x + 0.5f, y + 0.5f, 0.0f, 1.0f);
      fs.main();
      const vec4 c = fs.reflectedColor;
      const size_t idx = (size_t)y * W + x;
      outReflection[idx * 4 + 0] = oracle::floatToHalf(c.x * c.w);
      outReflection[idx * 4 + 1] = oracle::floatToHalf(c.y * c.w);
      outReflection[idx * 4 + 2] = oracle::floatToHalf(c.z * c.w);
      outReflection[idx * 4 + 3] = oracle::floatToHalf(c.w);
      if (outHit) outHit[idx] = c.w != 0.0f;
    }
}
extern "C" {
void shaderref_ssr_capture(const OracleGlobalUniforms* g, const OracleGBuffer* gb, const OracleIBL* ibl, const OracleLights* li,
                           uint16_t* outReflection, uint8_t* outHit) {
  ssrCapture<SsrFrag>(g, gb, ibl, li, outReflection, outHit, false);
}
void shaderref_ssr_capture_flags(const OracleGlobalUniforms* g, const OracleGBuffer* gb, const OracleIBL* ibl, const OracleLights* li,
                                 uint16_t* outReflection, uint8_t* outHit, uint32_t flags) {
  if (flags & 1u) ssrCapture<SsrFragFused>(g, gb, ibl, li, outReflection, outHit, (flags & 2u) != 0);
  else ssrCapture<SsrFrag>(g, gb, ibl, li, outReflection, outHit, (flags & 2u) != 0);
}

// SSRGlossyConvolve.comp main, dispatched as ReflectionBuffer::convolveReflectionBuffer does (Src/ReflectionBuffer.cpp:224-278):
// for mip L = 1.. : source = single-mip view of L - 1, target = L, width / height of L, direction (0,1) for odd L, (1,0) for even
void shaderref_glossy_convolve(uint16_t* mips, int W, int H, int mipCount) {
  oracle::TexChain ch{mips, W, H, mipCount, oracle::FMT_RGBA16F};
  for (int level = 1; level < mipCount; ++level) {
    const oracle::Tex src = ch.level(level - 1), dst = ch.level(level);
    sampler2D s; s.chain = oracle::TexChain{src.data, src.w, src.h, 1, oracle::FMT_RGBA16F};
    image2D img; img.texels = (uint16_t*)dst.data; img.w = dst.w; img.h = dst.h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst.h; ++y)
      for (int x = 0; x < dst.w; ++x) {
        ConvolveComp cs;
        cs.samplerHeap = &s; cs.imageHeap = &img;
        cs.pushConstants.srcMipTexHandle = 0; cs.pushConstants.targetMipImgHandle = 0;
        cs.pushConstants.width = (uint)dst.w; cs.pushConstants.height = (uint)dst.h;
        cs.pushConstants.direction = (level & 1) ? vec2(0.0f, 1.0f) : vec2(1.0f, 0.0f);
        cs.pushConstants.roughness = 0.0f;
        cs.gl_GlobalInvocationID = uvec3((uint)x, (uint)y, 0u);
        cs.main();
      }
  }
}

} // extern "C"

template <class Frag> static void bindDeferred(Frag& fs, Bound& b, typename Frag::GlobalUniforms* gu, typename Frag::POINT_LIGHTS* pl, const sampler2D* refl) {
  fs.globalUniforms = gu; fs.pointLights = pl;
  fs.environmentMap = b.tex[H_ENV]; fs.prefilteredMap = b.tex[H_PRE]; fs.irradianceMap = b.tex[H_IRR]; fs.brdfLut = b.tex[H_LUT];
  fs.gBufferPosition = b.tex[H_POSITION]; fs.gBufferNormal = b.tex[H_NORMAL]; fs.gBufferAlbedo = b.tex[H_ALBEDO];
  fs.gBufferMetallicRoughnessOcclusion = b.tex[H_MRO];
  if (refl) fs.reflectionBuffer = *refl;
  fs.shadowMapArray = b.cubes;
}

extern "C" {

// computeSSAO (SSAO.glsl:31-84) seeded as DeferredPass.frag:42 seeds it and called as :72 calls it; count = 24 (1 - result)
void shaderref_ssao(const OracleGlobalUniforms* g, const OracleGBuffer* gb, uint8_t* outCount) {
  const int W = gb->W, H = gb->H;
  Bound b(*gb, nullptr, nullptr, 0);
  DeferredFrag::GlobalUniforms gu; memcpy(&gu, g, sizeof gu);
  DeferredFrag::POINT_LIGHTS pl{nullptr};
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const size_t idx = (size_t)y * W + x;
      DeferredFrag fs;
      bindDeferred(fs, b, &gu, &pl, nullptr);
      fs.gl_FragCoord = vec4(x + 0.5f, y + 0.5f, 0.0f, 1.0f);
      fs.uv = vec2((float)((x + 0.5) / W), (float)((y + 0.5) / H));
      announce(x, y, W, H, fs.uv.x, fs.uv.y);
      fs.seed = uvec2(fs.gl_FragCoord.xy);                          // DeferredPass.frag:42
      const vec4 position = texture(fs.gBufferPosition, fs.uv).rgba; // :44
      if (position.a == 0.0f) { outCount[idx] = 255; continue; }      // :45 (the pass never calls computeSSAO there)
      const vec3 normal = normalize(texture(fs.gBufferNormal, fs.uv).xyz); // :55
      const float ao = fs.computeSSAO(fs.uv, position.xyz, normal);  // :72
      outCount[idx] = (uint8_t)lrintf((1.0f - ao) * 24.0f);
    }
}

} // extern "C"

// DeferredPass.vert + DeferredPass.frag main (patched per R1 / R2, oracle/Makefile); flags & 1 = SKIP_TONEMAP
template <class Frag> static void deferred(const OracleGlobalUniforms* g, const OracleGBuffer* gb, const OracleIBL* ibl, const OracleLights* li,
                                           const uint16_t* reflectionMips, int reflMipCount, float* outColor, bool closedForm) {
  const int W = gb->W, H = gb->H;
  Bound b(*gb, ibl, li, g->lightCount);
  typename Frag::GlobalUniforms gu; memcpy(&gu, g, sizeof gu);
  typename Frag::POINT_LIGHTS pl;
  static_assert(sizeof(typename Frag::PointLight) == sizeof(SsrFrag::PointLight), "same struct text");
  pl.pointLightArr = reinterpret_cast<typename Frag::PointLight*>(b.lights.data());
  sampler2D refl; refl.chain = oracle::TexChain{reflectionMips, W, H, reflMipCount, oracle::FMT_RGBA16F};
  double vd[3][3];
  DeferredVert::GlobalUniforms vgu; memcpy(&vgu, g, sizeof vgu);
  vertexDirections<DeferredVert>([&](DeferredVert& vs) { vs.globalUniforms = &vgu; }, vd);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      Frag fs;
      bindDeferred(fs, b, &gu, &pl, &refl);
      const double u = (x + 0.5) / W, v = (y + 0.5) / H;
      fs.uv = vec2((float)u, (float)v);
      announce(x, y, W, H, fs.uv.x, fs.uv.y);
      if (closedForm) { // DeferredPass.vert:11-19 with screenPos := this pixel's uv
        const vec4 pos = vec4(fs.uv * 2.0f - 1.0f, 0.0f, 1.0f);
        fs.direction = mat3(vgu.inverseView) * (vgu.inverseProjection * pos).xyz;
      } else {
        fs.direction = interpolate(vd, u, v);
      }
      fs.gl_FragCoord = vec4(x + 0.5f, y + 0.5f, 0.0f, 1.0f);
      fs.main();
      float* o = outColor + ((size_t)y * W + x) * 4;
      o[0] = fs.outColor.x; o[1] = fs.outColor.y; o[2] = fs.outColor.z; o[3] = fs.outColor.w;
    }
}
extern "C" {

void shaderref_deferred_shade(const OracleGlobalUniforms* g, const OracleGBuffer* gb, const OracleIBL* ibl, const OracleLights* li,
                              const uint16_t* reflectionMips, int reflMipCount, uint32_t flags, float* outColor) {
  // flags: 1 = SKIP_TONEMAP; 4 = the varying evaluated per pixel from the vertex stage's expression instead of interpolated
  if (flags & 1u) deferred<DeferredFragLinear>(g, gb, ibl, li, reflectionMips, reflMipCount, outColor, (flags & 4u) != 0);
  else deferred<DeferredFrag>(g, gb, ibl, li, reflectionMips, reflMipCount, outColor, (flags & 4u) != 0);
}

// DeferredPass.vert's direction at every pixel centre (3 floats per pixel)
void shaderref_view_directions(const OracleGlobalUniforms* g, int W, int H, float* out) {
  double vd[3][3];
  DeferredVert::GlobalUniforms vgu; memcpy(&vgu, g, sizeof vgu);
  vertexDirections<DeferredVert>([&](DeferredVert& vs) { vs.globalUniforms = &vgu; }, vd);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const vec3 d = interpolate(vd, (x + 0.5) / W, (y + 0.5) / H);
      float* o = out + ((size_t)y * W + x) * 3;
      o[0] = d.x; o[1] = d.y; o[2] = d.z;
    }
}

// IBL_Precompute/GenIrradianceMap.comp main at probe texels (x, y, face = 0) of an outW x outH equirect image, dispatched as
// ImageBasedLighting.cpp:315-343 does: environment map with its blit mip chain and a REPEAT sampler (:168-174), push constants =
// the output size
void shaderref_ibl_irradiance(const float* chain, int W, int H, int mips, int outW, int outH, const int32_t* texels, int n, float* out) {
  sampler2D env; env.chain = oracle::TexChain{chain, W, H, mips, oracle::FMT_RGBA32F}; env.address = oracle::ADDR_REPEAT;
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < n; ++t) {
    std::vector<float> dst((size_t)outW * outH * 4, 0.0f);
    IrradianceComp cs;
    cs.environmentMap = env;
    cs.irradianceMap.texels32 = dst.data(); cs.irradianceMap.w = outW; cs.irradianceMap.h = outH;
    cs.pushConstants.width = (float)outW; cs.pushConstants.height = (float)outH;
    const int x = texels[t * 3], y = texels[t * 3 + 1];
    cs.gl_GlobalInvocationID = uvec3((uint)x, (uint)y, 0u);
    cs.main();
    memcpy(out + (size_t)t * 4, dst.data() + ((size_t)y * outW + x) * 4, 16);
  }
}
// IBL_Precompute/PreFilterEnvMap.comp main at probe texels of one outW x outH level (ImageBasedLighting.cpp:351-402)
void shaderref_ibl_prefilter(const float* chain, int W, int H, int mips, int outW, int outH, float roughness, const int32_t* texels, int n, float* out) {
  sampler2D env; env.chain = oracle::TexChain{chain, W, H, mips, oracle::FMT_RGBA32F}; env.address = oracle::ADDR_REPEAT;
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < n; ++t) {
    std::vector<float> dst((size_t)outW * outH * 4, 0.0f);
    PrefilterComp cs;
    cs.environmentMap = env;
    cs.prefilteredMip.texels32 = dst.data(); cs.prefilteredMip.w = outW; cs.prefilteredMip.h = outH;
    cs.pushConstants.envMapWidth = (float)W; cs.pushConstants.envMapHeight = (float)H;
    cs.pushConstants.width = (float)outW; cs.pushConstants.height = (float)outH;
    cs.pushConstants.roughness = roughness;
    const int x = texels[t * 3], y = texels[t * 3 + 1];
    cs.gl_GlobalInvocationID = uvec3((uint)x, (uint)y, 0u);
    cs.main();
    memcpy(out + (size_t)t * 4, dst.data() + ((size_t)y * outW + x) * 4, 16);
  }
}

// ---- the programmable stages of the rasterising passes, as hooks of althea_oracle_raster.cpp ------------------------------------
struct OracleTex { const uint8_t* texels; int32_t w, h, mips; uint32_t sampler; };
struct OraclePrim { // oracle/althea_oracle_raster.cpp, same layout
  const float* verts; const uint32_t* idx; uint32_t triCount; uint32_t frontCW; float model[16]; float baseColorFactor[4];
  int32_t baseUv, mrUv; float normalScale, metallicFactor, roughnessFactor, alphaCutoff; OracleTex base, normal, mr;
};
struct OracleStageHooks {
  void (*gbufferVertex)(const OraclePrim*, uint32_t, const float*, const float*, float*, float*, float*);
  int (*gbufferFragment)(const OraclePrim*, const float*, const float (*)[2], const float (*)[2], const float (*)[2], float*, float*, float*);
  void (*shadowVertex)(const OraclePrim*, uint32_t, const float*, const float*, const float*, float*, float*);
  int (*shadowFragment)(const OraclePrim*, const float*, const float*, const float*, const float*, float*);
};
}
namespace {
constexpr int kVertexFloats = 26; // InstanceDataCommon.h:45-53: position 0, tangent 3, bitangent 6, normal 9, uvs 12
template <class M> void fillMaterial(M& m, const OraclePrim* p) { // MaterialConstants as Src/Primitive.cpp fills it
  memset(&m, 0, sizeof m);
  m.baseColorFactor = vec4(p->baseColorFactor[0], p->baseColorFactor[1], p->baseColorFactor[2], p->baseColorFactor[3]);
  m.baseTextureCoordinateIndex = p->baseUv;
  m.normalMapTextureCoordinateIndex = p->baseUv;
  m.metallicRoughnessTextureCoordinateIndex = p->mrUv;
  m.normalScale = p->normalScale; m.metallicFactor = p->metallicFactor; m.roughnessFactor = p->roughnessFactor;
  m.alphaCutoff = p->alphaCutoff;
  m.baseTextureHandle = 0; m.normalTextureHandle = 1; m.metallicRoughnessTextureHandle = 2;
}
void materialSamplers(sampler2D s[3], const OraclePrim* p) { // missing textures: the engine binds 1 x 1 defaults (white, flat normal, white)
  s[0].materialTex = &p->base; s[1].materialTex = &p->normal; s[2].materialTex = &p->mr;
  s[1].dflt[0] = s[1].dflt[1] = 128.0f / 255.0f;
}
template <class VS> void vertexInputs(VS& vs, const OraclePrim* p, uint32_t i) {
  const float* v = p->verts + (size_t)i * kVertexFloats;
  vs.position = vec3(v[0], v[1], v[2]);
  vs.tbn = mat3(vec3(v[3], v[4], v[5]), vec3(v[6], v[7], v[8]), vec3(v[9], v[10], v[11]));
  for (int k = 0; k < 4; ++k) vs.uvs[k] = vec2(v[12 + 2 * k], v[13 + 2 * k]);
}
void hookGbufferVertex(const OraclePrim* p, uint32_t i, const float* projection, const float* view, float clip[4], float world[3], float tbn[9]) {
  GltfVert vs;
  vertexInputs(vs, p, i);
  GltfVert::GlobalUniforms gu; memset(&gu, 0, sizeof gu);
  memcpy(&gu.projection, projection, 64); memcpy(&gu.view, view, 64);
  for (int c = 0; c < 4; ++c) gu.inverseView[c] = vec4(c == 0, c == 1, c == 2, c == 3); // only `direction` reads it, which no stage of the path consumes
  GltfVert::_primitiveConstants_BUFFER pc; memset(&pc, 0, sizeof pc); // nodeIdx 0, not skinned
  mat4 model; memcpy(&model, p->model, 64);
  GltfVert::_transformBuffer_BUFFER tb{&model};
  vs.globalUniforms = &gu; vs._primitiveConstantsHeap = &pc; vs._transformBufferHeap = &tb;
  vs.pushConstants.matrixBufferHandle = 0; vs.pushConstants.primConstantsBuffer = 0; vs.pushConstants.globalUniformsHandle = 0;
  vs.main();
  for (int c = 0; c < 4; ++c) clip[c] = vs.gl_Position[c];
  for (int c = 0; c < 3; ++c) world[c] = vs.worldPosition[c];
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) tbn[3 * c + r] = vs.vertTbn[c][r];
}
int hookGbufferFragment(const OraclePrim* p, const float tbn[9], const float (*uvs)[2], const float (*ddx)[2], const float (*ddy)[2],
                        float outN[4], float outA[4], float outM[4]) {
  GltfFrag fs;
  for (int k = 0; k < 4; ++k) fs.uvs[k] = vec2(uvs[k][0], uvs[k][1]);
  fs.fragTBN = mat3(vec3(tbn[0], tbn[1], tbn[2]), vec3(tbn[3], tbn[4], tbn[5]), vec3(tbn[6], tbn[7], tbn[8]));
  GltfFrag::_materialConstants_BUFFER mc; fillMaterial(mc.val, p);
  GltfFrag::_primitiveConstants_BUFFER pc; memset(&pc, 0, sizeof pc); // materialHandle 0
  sampler2D tex[3]; materialSamplers(tex, p);
  fs._materialConstantsHeap = &mc; fs._primitiveConstantsHeap = &pc; fs._materialSamplerHeap = tex;
  fs.pushConstants.primConstantsBuffer = 0;
  gImplicit.uvs = uvs; gImplicit.ddx = ddx; gImplicit.ddy = ddy; gImplicit.sets = 4;
  fs.main();
  gImplicit.sets = 0;
  if (fs.gl_Discarded) return 0;
  for (int c = 0; c < 4; ++c) { outN[c] = fs.GBuffer_Normal[c]; outA[c] = fs.GBuffer_Albedo[c]; outM[c] = fs.GBuffer_MetallicRoughnessOcclusion[c]; }
  return 1;
}
void hookShadowVertex(const OraclePrim* p, uint32_t i, const float* lightPos, const float* view, const float* projection, float clip[4], float cs[3]) {
  ShadowVert vs;
  vertexInputs(vs, p, i);
  ShadowVert::PointLight light; light.position = vec3(lightPos[0], lightPos[1], lightPos[2]);
  ShadowVert::POINT_LIGHTS pl{&light};
  ShadowVert::PointLightConstants lc; memset(&lc, 0, sizeof lc);
  memcpy(&lc.projection, projection, 64); memcpy(&lc.views[0], view, 64);
  fillMaterial(vs.legacyMaterial, p);
  vs.pointLights = &pl; vs.pointLightConstants = &lc;
  memcpy(&vs.pushConstants.model, p->model, 64);
  vs.pushConstants.lightIdx = 0; vs.pushConstants.pointLightBufferHandle = 0; vs.pushConstants.pointLightConstantsHandle = 0;
  vs.gl_ViewIndex = 0;
  vs.main();
  for (int c = 0; c < 4; ++c) clip[c] = vs.gl_Position[c];
  for (int c = 0; c < 3; ++c) cs[c] = vs.worldPosCS[c];
}
int hookShadowFragment(const OraclePrim* p, const float cs[3], const float uv[2], const float ddx[2], const float ddy[2], float* depth) {
  ShadowFrag fs;
  fs.worldPosCS = vec3(cs[0], cs[1], cs[2]);
  fs.baseColorUV = vec2(uv[0], uv[1]);
  fillMaterial(fs.legacyMaterial, p);
  sampler2D tex[3]; materialSamplers(tex, p);
  fs.textureHeap = tex;
  const float u1[1][2] = {{uv[0], uv[1]}}, dx1[1][2] = {{ddx[0], ddx[1]}}, dy1[1][2] = {{ddy[0], ddy[1]}};
  gImplicit.uvs = u1; gImplicit.ddx = dx1; gImplicit.ddy = dy1; gImplicit.sets = 1;
  fs.main();
  gImplicit.sets = 0;
  if (fs.gl_Discarded) return 0;
  *depth = fs.gl_FragDepth;
  return 1;
}
const OracleStageHooks kHooks{hookGbufferVertex, hookGbufferFragment, hookShadowVertex, hookShadowFragment};
}
extern "C" {
// `sampleMaterial` = liboracle.so's oracle_sample_texture (the passes' texture unit); returns what oracle_set_stage_hooks takes
const void* shaderref_raster_hooks(void (*sampleMaterial)(const void*, const float*, const float*, const float*, const float*, float*)) {
  gSampleMaterial = sampleMaterial;
  return &kHooks;
}

void shaderref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

} // extern "C"
