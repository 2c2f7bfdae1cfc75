// Per-frame kernels of the deferred screen-space path, sm_100a.
//
//   ssr_capture_kernel      <- Shaders/SSR.vert + SSR.frag (reflection mip 0)
//   glossy_convolve_kernel  <- Shaders/SSRGlossyConvolve.comp (one launch per mip, ReflectionBuffer.cpp:224-278)
//   ssao_kernel             <- Shaders/SSAO.glsl computeSSAO (occluded-ray count per pixel)
//   deferred_shade_kernel   <- Shaders/DeferredPass.vert/.frag + PBR/PBRMaterial.glsl pbrMaterial
//
// Compiled twice (fast / parity), see device_math.cuh. One thread per pixel; a warp covers a 16x2 pixel strip, so
// G-buffer reads are 128-byte coalesced rows (128-bit loads for position, 64-bit for normal/reflection).
#include <atomic>

#include "device_math.cuh"
#include "launchers.h"

namespace ALTHEA_NS {

// ---- run-time equirect IBL lookups: CLAMP_TO_EDGE (ImageBasedLighting.cpp:469-475,519-524,556-560) ----------------
ADEV V2 equirectUv(V3 d) { // PBRMaterial.glsl:5-7
  float yaw = atan2f(d.z, d.x);
  float pitch = -atan2f(d.y, fsqrt(d.x * d.x + d.z * d.z));
  V2 uv;
#ifdef ALTHEA_PARITY
  uv.x = (0.5f * yaw) / kPi + 0.5f;
  uv.y = pitch / kPi + 0.5f;
#else
  uv.x = fmaf(yaw, 0.5f / kPi, 0.5f);
  uv.y = fmaf(pitch, 1.0f / kPi, 0.5f);
#endif
  return uv;
}
ADEV V3 sampleEnvMapLod0(const FrameParams& P, V3 dir) { // DeferredPass.frag:33-39
  V2 uv = equirectUv(dir);
  return xyz(bilinear<FmtRGBA32F, AddrClamp, kFastTex>(P.env, uv.x, uv.y));
}
ADEV V3 sampleEnvMapRough(const FrameParams& P, V3 dir, float roughness) { // PBRMaterial.glsl:4-11
  V2 uv = equirectUv(dir);
  return xyz(trilinear<FmtRGBA32F, AddrClamp, kFastTex>(P.pre, uv.x, uv.y, 4.0f * roughness));
}
ADEV V3 sampleIrrMap(const FrameParams& P, V3 n) { // PBRMaterial.glsl:13-19
  V2 uv = equirectUv(n);
  return xyz(bilinear<FmtRGBA32F, AddrClamp, kFastTex>(P.irr, uv.x, uv.y));
}

// cube-array lookup (Vulkan face selection table; the bilinear footprint clamps inside the face)
ADEV float sampleShadowCube(const FrameParams& P, V3 q, int light) {
  float ax = fabsf(q.x), ay = fabsf(q.y), az = fabsf(q.z);
  int face;
  float sc, tc, ma;
  if (ax >= ay && ax >= az) { ma = ax; if (q.x >= 0.0f) { face = 0; sc = -q.z; tc = -q.y; } else { face = 1; sc = q.z; tc = -q.y; } }
  else if (ay >= az)        { ma = ay; if (q.y >= 0.0f) { face = 2; sc = q.x; tc = q.z; } else { face = 3; sc = q.x; tc = -q.z; } }
  else                      { ma = az; if (q.z >= 0.0f) { face = 4; sc = q.x; tc = -q.y; } else { face = 5; sc = -q.x; tc = -q.y; } }
#ifdef ALTHEA_PARITY
  float s = 0.5f * sc / ma + 0.5f, t = 0.5f * tc / ma + 0.5f;
#else
  const float hr = 0.5f * rcpf(ma);
  float s = fmaf(sc, hr, 0.5f), t = fmaf(tc, hr, 0.5f);
#endif
  ImgView layer = P.shadow;
  layer.ptr = static_cast<const char*>(P.shadow.ptr) + (size_t)(6 * light + face) * P.shadowLayerStride;
  return bilinearR32F<AddrClamp, kFastTex>(layer, s, t);
}

// ---- PBRMaterial.glsl:41-70 ---------------------------------------------------------------------------------------
ADEV float ndfGgx(float NdotH, float a2) {
  float tmp = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
  float denom = kPi * tmp * tmp;
  return fdiv(a2, denom);
}
ADEV float pow5(float x) {
#ifdef ALTHEA_PARITY
  return powf(x, 5.0f);
#else
  float x2 = x * x;
  return x2 * x2 * x;
#endif
}
ADEV V3 fresnelSchlick(float NdotV, V3 F0, float roughness) {
  float om = 1.0f - roughness;
  V3 m = mk3(max_glsl(om, F0.x), max_glsl(om, F0.y), max_glsl(om, F0.z));
  return F0 + (m - F0) * pow5(1.0f - NdotV);
}
ADEV float geometrySchlickGgx(float NdotV, float k) { return fdiv(NdotV, NdotV * (1.0f - k) + k); }
ADEV float geometrySmith(float NdotL, float NdotV, float k) { return geometrySchlickGgx(NdotV, k) * geometrySchlickGgx(NdotL, k); }

// PBRMaterial.glsl:72-162 with the current (worldPos, V, N, ...) signature
ADEV V3 pbrMaterial(const FrameParams& P, V3 worldPos, V3 V, V3 N, V3 baseColor, V3 reflectedColor, V3 irradianceColor, float metallic,
                    float roughness, float ambientOcclusion) {
  float NdotV = max_glsl(dot3(N, -V), 0.0f);
  V3 F0 = mix3(mk3(0.04f, 0.04f, 0.04f), baseColor, metallic);
  float a = roughness * roughness;
  float a2 = a * a;
  float kDirect = (a + 1.0f) * (a + 1.0f) / 8.0f;
  V3 color = mk3(0.0f, 0.0f, 0.0f);
  const V3 one = mk3(1.0f, 1.0f, 1.0f);
  V3 dielectricBase = mix3(baseColor, mk3(0.0f, 0.0f, 0.0f), metallic);
  {
    V3 F = fresnelSchlick(NdotV, F0, roughness);
    V3 diffuseColor = (one - F) * dielectricBase;
    V4 lut = bilinear<FmtRGBA8, AddrClamp, kFastTex>(P.lut, NdotV, roughness);
    V3 ambientSpecular = reflectedColor * (F * lut.x + mk3(lut.y, lut.y, lut.y));
    color = color + (irradianceColor * diffuseColor + ambientSpecular) * ambientOcclusion;
  }
  const int lightCount = P.g.lightCount;
  for (int i = 0; i < lightCount; ++i) {
    const float4* lp = reinterpret_cast<const float4*>(P.lights) + 2 * i;
    float4 l0 = __ldg(lp), l1 = __ldg(lp + 1);
    V3 L = mk3(l0.x, l0.y, l0.z) - worldPos;
    float LdistSq = dot3(L, L);
#ifdef ALTHEA_PARITY
    float Ldist = fsqrt(LdistSq);
    L = L / Ldist;
#else
    // a light behind the surface contributes (diffuse + specular) * radiance * max(dot(N, L), 0) = +0 (the specular term's
    // geometry factor vanishes with NdotL, its denominator does not): no shadow lookup, no BRDF. Neighbouring pixels share
    // their normals' side of a light, so warps skip together. One MUFU.RSQ serves the length and the normalisation.
    if (!(dot3(N, L) > 0.0f)) continue;
    const float invLdist = rsqrtf(LdistSq);
    const float Ldist = LdistSq * invLdist;
    L = L * invLdist;
#endif
    if (P.shadowRes > 0) {
      float closestDepth = sampleShadowCube(P, mk3(L.x, -L.y, -L.z), i);
      closestDepth *= 1000.0f;
      if (closestDepth < (Ldist - 0.5f)) continue;
    }
    V3 radiance = mk3(l1.x, l1.y, l1.z) / LdistSq;
    V3 H = normalize3(V + L);
    float NdotL = max_glsl(dot3(N, L), 0.0f);
    float NdotH = max_glsl(dot3(N, H), 0.0f);
    V3 F = fresnelSchlick(NdotH, F0, roughness);
#ifdef ALTHEA_PARITY
    V3 diffuseColor = ((one - F) * dielectricBase) / kPi;
#else
    V3 diffuseColor = ((one - F) * dielectricBase) * (1.0f / kPi);
#endif
    V3 specularColor = ((ndfGgx(NdotH, a2) * F) * geometrySmith(NdotL, NdotV, kDirect)) / (4.0f * NdotL * NdotV + 0.0001f);
    color = color + ((diffuseColor + specularColor) * radiance) * NdotL;
  }
  return color;
}

// DeferredPass.vert:10-22 == SSR.vert:15-23 at the pixel centre
ADEV V3 viewDirection(const FrameParams& P, float u, float v) {
  V4 p = mul44(P.g.inverseProjection, mk4(u * 2.0f - 1.0f, v * 2.0f - 1.0f, 0.0f, 1.0f));
  return mul33(P.g.inverseView, xyz(p));
}

// Misc/ReconstructPosition.glsl:4-22
ADEV V3 reconstructPosition(const FrameParams& P, float u, float v, float dRaw) {
  const float near = 0.01f, far = 1000.0f;
  // one fused multiply-add in BOTH builds: unfused, the cancellation costs ~3 digits of eye depth (see the oracle)
  float d = far * near / __fmaf_rn(dRaw, far - near, -far);
  V4 dirH = mul44(P.g.inverseProjection, mk4(2.0f * u - 1.0f, 2.0f * v - 1.0f, 2.0f, 1.0f));
  V3 cam = mk3(P.g.inverseView[12], P.g.inverseView[13], P.g.inverseView[14]);
  V3 zc = mk3(P.g.inverseView[8], P.g.inverseView[9], P.g.inverseView[10]);
#ifdef ALTHEA_PARITY
  V4 h = mk4(dirH.x / dirH.w, dirH.y / dirH.w, dirH.z / dirH.w, 0.0f);
  V4 wd = mul44(P.g.inverseView, h);
  V3 dir = normalize3(xyz(wd));
  float f = dot3(dir, zc);
  return cam + (d * dir) / f;
#else
  // normalize() cancels against the division by dot(dir, zAxis), and so does the 1/w: pos = cam + d * wd / dot(wd, z)
  V3 wd = mul33(P.g.inverseView, xyz(dirH));
  return cam + wd * (d / dot3(wd, zc));
#endif
}

ADEV bool outside01(float u, float v) { return u < 0.0f || u > 1.0f || v < 0.0f || v > 1.0f; }
ADEV V2 projectUv(const FrameParams& P, V3 p) { // 0.5 * clip.xy / clip.w + 0.5
  V4 pe = mul44(P.projView, mk4(p.x, p.y, p.z, 1.0f));
  V2 uv;
  uv.x = 0.5f * pe.x / pe.w + 0.5f;
  uv.y = 0.5f * pe.y / pe.w + 0.5f;
  return uv;
}

// ---- TMA bulk copies into shared memory, completion on an mbarrier (UBLKCP + SYNCS in SASS) -------------------------------
ADEV uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
ADEV void mbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
ADEV void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
ADEV void bulkCopyG2S(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)), "l"(src), "r"(bytes),
               "r"(smemAddr(bar))
               : "memory");
}
ADEV void mbarWait(uint64_t* bar, uint32_t phase) {
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra WAIT_%=;\n}\n" ::"r"(smemAddr(bar)), "r"(phase)
               : "memory");
}


// margins shared by the SSAO and SSR sign tests over plane records (derivation at ssao_cull_kernel)
constexpr float kNu = 1.7e-5f;  // sqrt(3) * 4e-6 (true model error) + 2^-20 (coefficients) + 64 ulp (the fp32 tap), rounded up
constexpr float kTiny = 1e-15f; // decisions on |value| <= kTiny are always re-evaluated (products must not underflow)

// ---- SSR ------------------------------------------------------------------------------------------------------------
// Padded depth (engine scratch, written by ssr_depth_pad_kernel once per frame): the depth image with a one-texel CLAMP_TO_EDGE
// border, (W + 2) x (H + 2) floats. The footprint whose top-left tap is texel (ix, iy), ix in [-1, W-1], iy in [-1, H-1], is the
// 2 x 2 block at padded (ix + 1, iy + 1): a march step needs no clamps and one 32-bit index for its four taps (the clamped
// four-tap form spends about 30 of its ~100 SASS instructions per step on clamping and 64-bit address arithmetic), and
// neighbouring lanes still read neighbouring floats. Same texels, same lerps: bit-identical to the clamped form in both builds.
__global__ void __launch_bounds__(256) ssr_depth_pad_kernel(const __grid_constant__ FrameParams P) {
  const int qx = blockIdx.x * 32 + (threadIdx.x & 31);
  const int qy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (qx == 0 && qy == 0 && P.ssrHitCount) { P.ssrHitCount[0] = 0u; P.ssrHitCount[1] = 0u; } // the hit list the march (next in the stream) appends to, and its chunk counter
  if (qx > P.W + 1 || qy > P.H + 1) return;
  const float d = __ldg(rowPtr<float>(P.depth, AddrClamp::wrap(qy - 1, P.H)) + AddrClamp::wrap(qx - 1, P.W));
  if (P.depthQuadOrigin) { // footprint records instead: this thread's texel is the footprint's top-left tap (qx - 1, qy - 1)
    if (qx > P.W || qy > P.H) return;
    const int i1 = AddrClamp::wrap(qx, P.W), j1 = AddrClamp::wrap(qy, P.H);
    const float* r0 = rowPtr<float>(P.depth, AddrClamp::wrap(qy - 1, P.H));
    const float* r1 = rowPtr<float>(P.depth, j1);
    const_cast<float4*>(P.depthQuadOrigin)[(qy - 1) * P.depthQuadRow + (qx - 1)] = make_float4(d, __ldg(r0 + i1), __ldg(r1 + AddrClamp::wrap(qx - 1, P.W)), __ldg(r1 + i1));
    return;
  }
  const_cast<float*>(P.depthPad)[(size_t)qy * P.depthPadRow + qx] = d;
}
struct DepthTap { float t00, t10, t01, t11, fx, fy; };
// (u, v) has passed outside01, so x is in [-0.5, W - 0.5] and floor(x) is in [-1, W-1] without clamping
ADEV DepthTap depthTapPadded(const FrameParams& P, float u, float v) {
  float x = __fsub_rn(__fmul_rn(u, P.Wf), 0.5f), y = __fsub_rn(__fmul_rn(v, P.Hf), 0.5f); // rule A1, never contracted
#ifdef ALTHEA_PARITY
  if (!(x == x)) x = 0.0f; // a NaN coordinate is not "outside": the restatement taps texel (0, 0) with zero weights
  if (!(y == y)) y = 0.0f;
#endif
  const int ix = __float2int_rd(x), iy = __float2int_rd(y); // |x| < 2^24: (float)ix == floorf(x) exactly
  DepthTap q;
  q.fx = x - (float)ix;
  q.fy = y - (float)iy;
  const float* r0 = P.depthPadOrigin + (iy * P.depthPadRow + ix); // origin = &padded(1, 1) = texel (0, 0)
  const float* r1 = r0 + P.depthPadRow;
  q.t00 = __ldg(r0); q.t10 = __ldg(r0 + 1); q.t01 = __ldg(r1); q.t11 = __ldg(r1 + 1);
  return q;
}

ADEV V4 environmentLitSample(const FrameParams& P, V3 currentPos, float u, float v, V3 rayDir, V3 normal) { // SSR.frag:55-78
  V3 baseColor = xyz(bilinear<FmtRGBA8, AddrClamp, kFastTex>(P.albedo, u, v));
  V3 mro = xyz(bilinear<FmtRGBA8, AddrClamp, kFastTex>(P.mro, u, v));
  V3 rd = normalize3(rayDir);
  V3 reflectedDirection = reflect3(rd, normal);
  V3 reflectedColor = sampleEnvMapRough(P, reflectedDirection, mro.y);
  V3 irradianceColor = sampleIrrMap(P, normal);
  V3 m = pbrMaterial(P, currentPos, rd, normal, baseColor, reflectedColor, irradianceColor, mro.x, mro.y, 1.0f);
  return mk4(m.x, m.y, m.z, 1.0f);
}

#ifndef ALTHEA_PARITY
// ---- fast build: the march of raymarchGBuffer (SSR.frag:97-130) with its per-step algebra folded ---------------------------
// The fast build's contract is the 0.1 % hit-mask bar, not the op order:
//  * reconstructPosition's matrix products are affine in (cu, cv): wd = W0 + Wu cu + Wv cv, dot(wd, zAxis) = S0 + Su cu + Sv cv
//    (FrameParams::ssr*, folded on the host), and along the march (cu, cv) = (u, v) + i step, so both are affine in the step
//    number. pos - worldPos = (cam - worldPos) + wd k, k = far near / den, den = (dRaw (far - near) - far) dot(wd, zAxis);
//  * dir = normalize(pos - worldPos) is never formed, and neither is k: the sign test of SSR.frag:118 only needs the sign of
//    projection = cP + k aP(i), and projection * den = cP den + far near aP(i) has that sign times sign(den), which does not
//    change along a ray (the first factor of den is negative for every depth in [0, 1], the second is the view-space z of the
//    texel's direction). One step is the bilinear depth tap, den, one FFMA and one product against the previous step's value;
//    the reciprocal, `dot(dir, rayDir) > 0.999` and the normal test run on the steps that flip (a handful per ray);
//  * the tap's footprint comes out of the mantissa of cu W + (1.5 * 2^23 - 1): no F2I / I2F round trip (quarter-rate on sm_100);
//    a tap exactly on a texel centre may take the footprint to its left with weight 1, the same value;
//  * the ray leaves [0, 1]^2 at a step number known to within a fraction of a step when it starts: the per-step test is one
//    compare against that number minus two, the exact outside01 test of SSR.frag:106 runs on the last steps only.
// One 16-byte record per bilinear footprint instead of four floats of the padded copy: a tap becomes one 128-bit load (the march's
// loads touch 10.4 sectors per warp-level request, its rays being as incoherent as the normals they reflect off), but the records
// are four times the bytes through L1 and L2: 1.76 ms against 1.37 ms at 4K. Off; round 1 had measured the same with the old step.
#ifndef ALTHEA_SSR_DEPTH_QUADS
#define ALTHEA_SSR_DEPTH_QUADS 0
#endif
struct SsrRayFast {
  float cu, cv, stepX, stepY;
  float fi, fSafe;       // steps taken so far; steps <= fSafe are inside the screen and below the 128-step cap for certain
  float S0, dS;          // dot(wd, zAxis) at step fi: S0 + fi dS
  float cP, aP0, daP;    // projection * den = cP den + aP0 + fi daP  (far near folded into aP)
  float prev;            // that value at the previous step; NaN before the first (`i > 0`, SSR.frag:118: NaN <= 0 is false)
};
constexpr float kFloorMagic = 12582912.0f; // 1.5 * 2^23: floats in [2^23, 2^24) are integers, their low mantissa bits the value
ADEV bool ssrRayFastSetup(const FrameParams& P, float u, float v, V3 worldPos, V3 rayDir, V3 perpRef, float stepX, float stepY, float dl, SsrRayFast& M) {
  // a ray whose end point projects to the pixel itself, to infinity or to NaN has NaN steps in the restatement: no tap of it can hit
  if (!(dl > 0.0f && dl < __int_as_float(0x7f800000))) return false;
  const V3 camMinusPos = mk3(P.g.inverseView[12], P.g.inverseView[13], P.g.inverseView[14]) - worldPos;
  const V3 W0 = mk3(P.ssrW0[0], P.ssrW0[1], P.ssrW0[2]), Wu = mk3(P.ssrWu[0], P.ssrWu[1], P.ssrWu[2]), Wv = mk3(P.ssrWv[0], P.ssrWv[1], P.ssrWv[2]);
  const V3 dW = mk3(fmaf(Wu.x, stepX, Wv.x * stepY), fmaf(Wu.y, stepX, Wv.y * stepY), fmaf(Wu.z, stepX, Wv.z * stepY));
  const V3 Wat = mk3(fmaf(Wu.x, u, fmaf(Wv.x, v, W0.x)), fmaf(Wu.y, u, fmaf(Wv.y, v, W0.y)), fmaf(Wu.z, u, fmaf(Wv.z, v, W0.z)));
  M.cu = u; M.cv = v; M.stepX = stepX; M.stepY = stepY;
  M.fi = 0.0f;
  // steps to the border the ray heads for (iterated fp32 sums drift by < 1e-5 of [0, 1] in 128 steps: a thousandth of a step)
  const float tX = (stepX > 0.0f ? 1.0f - u : u) / fabsf(stepX), tY = (stepY > 0.0f ? 1.0f - v : v) / fabsf(stepY);
  M.fSafe = fminf(fminf(tX, tY) - 2.0f, 128.0f);
  M.S0 = fmaf(P.ssrS[1], u, fmaf(P.ssrS[2], v, P.ssrS[0]));
  M.dS = fmaf(P.ssrS[1], stepX, P.ssrS[2] * stepY);
  M.cP = dot3(camMinusPos, perpRef);
  M.aP0 = (1000.0f * 0.01f) * dot3(Wat, perpRef);
  M.daP = (1000.0f * 0.01f) * dot3(dW, perpRef);
  M.prev = __int_as_float(0x7fc00000);
  return true;
}
// one step of the march: 0 = go on, 1 = the ray left the screen or took its 128 steps, 2 = hit (hitPos, hitNormal; M.cu, M.cv the tap)
ADEV int ssrRayFastStep(const FrameParams& P, SsrRayFast& M, V3 worldPos, V3 rayDir, V3& hitPos, V3& hitNormal) {
  M.cu += M.stepX;
  M.cv += M.stepY;
  M.fi += 1.0f;
  if (M.fi > M.fSafe) {
    if (M.fi > 128.0f || outside01(M.cu, M.cv)) return 1;
  }
  // bilinear depth tap, CLAMP_TO_EDGE (rule A1/A2) from the padded copy; a + t (b - a) lerps
  const float mx = fmaf(M.cu, P.Wf, kFloorMagic - 1.0f), my = fmaf(M.cv, P.Hf, kFloorMagic - 1.0f); // round(x - 0.5) + magic, x = cu W - 0.5
  const float fx = fmaf(M.cu, P.Wf, -0.5f) - (mx - kFloorMagic), fy = fmaf(M.cv, P.Hf, -0.5f) - (my - kFloorMagic);
  const int ix = __float_as_int(mx) - 0x4b400000, iy = __float_as_int(my) - 0x4b400000; // in [-1, W-1] x [-1, H-1]
  float t00, t10, t01, t11;
  if (ALTHEA_SSR_DEPTH_QUADS) { // one record per footprint (launch_ssr_capture's callers set it up whenever this build marches)
    const float4 q = __ldg(P.depthQuadOrigin + (iy * P.depthQuadRow + ix));
    t00 = q.x; t10 = q.y; t01 = q.z; t11 = q.w;
  } else {
    const float* r0 = P.depthPadOrigin + (iy * P.depthPadRow + ix);
    const float* r1 = r0 + P.depthPadRow;
    t00 = __ldg(r0); t10 = __ldg(r0 + 1); t01 = __ldg(r1); t11 = __ldg(r1 + 1);
  }
  const float top = fmaf(t10 - t00, fx, t00), bot = fmaf(t11 - t01, fx, t01);
  const float dRaw = fmaf(bot - top, fy, top);
  const float den = fmaf(dRaw, 1000.0f - 0.01f, -1000.0f) * fmaf(M.fi, M.dS, M.S0);
  const float projDen = fmaf(M.cP, den, fmaf(M.fi, M.daP, M.aP0));
  const bool flips = projDen * M.prev <= 0.0f;
  M.prev = projDen;
  if (flips) {
    // pos - worldPos = (cam - worldPos) + wd * (far near / den)
    const float k = (1000.0f * 0.01f) * rcpf(den);
    const V3 camMinusPos = mk3(P.g.inverseView[12], P.g.inverseView[13], P.g.inverseView[14]) - worldPos;
    const V3 wd = mk3(fmaf(P.ssrWu[0], M.cu, fmaf(P.ssrWv[0], M.cv, P.ssrW0[0])), fmaf(P.ssrWu[1], M.cu, fmaf(P.ssrWv[1], M.cv, P.ssrW0[1])),
                      fmaf(P.ssrWu[2], M.cu, fmaf(P.ssrWv[2], M.cv, P.ssrW0[2])));
    const V3 vv = mk3(fmaf(wd.x, k, camMinusPos.x), fmaf(wd.y, k, camMinusPos.y), fmaf(wd.z, k, camMinusPos.z));
    // dot(dir, rayDir) > 0.999  <=>  dot(vv, rayDir) > 0 and dot(vv, rayDir)^2 > 0.999^2 |vv|^2. A tap exactly AT worldPos (where the
    // restatement's normalize(0) is NaN) fails `along > 0` here too; only the step after it could differ, on nothing we render.
    const float along = dot3(vv, rayDir);
    if (along > 0.0f && along * along > (0.999f * 0.999f) * dot3(vv, vv)) {
      const V3 currentNormal = normalize3(xyz(bilinear<FmtRGBA16F, AddrClamp>(P.normal, M.cu, M.cv)));
      if (dot3(currentNormal, rayDir) < 0.0f) {
        hitPos = worldPos + vv;
        hitNormal = currentNormal;
        return 2;
      }
    }
  }
  return 0;
}
// Two steps per call: the depth taps of both are requested before either is evaluated, so the dependent chain
// tap -> den -> sign test of a step overlaps the next step's memory latency (the march's top stall is long_scoreboard on these taps).
// Same values and the same order of decisions as two calls of ssrRayFastStep.
struct SsrTap { float t00, t10, t01, t11, fx, fy; };
ADEV SsrTap ssrTapFast(const FrameParams& P, float cu, float cv) {
  const float mx = fmaf(cu, P.Wf, kFloorMagic - 1.0f), my = fmaf(cv, P.Hf, kFloorMagic - 1.0f);
  SsrTap q;
  q.fx = fmaf(cu, P.Wf, -0.5f) - (mx - kFloorMagic);
  q.fy = fmaf(cv, P.Hf, -0.5f) - (my - kFloorMagic);
  const int ix = __float_as_int(mx) - 0x4b400000, iy = __float_as_int(my) - 0x4b400000;
  if (ALTHEA_SSR_DEPTH_QUADS) {
    const float4 r = __ldg(P.depthQuadOrigin + (iy * P.depthQuadRow + ix));
    q.t00 = r.x; q.t10 = r.y; q.t01 = r.z; q.t11 = r.w;
    return q;
  }
  const float* r0 = P.depthPadOrigin + (iy * P.depthPadRow + ix);
  const float* r1 = r0 + P.depthPadRow;
  q.t00 = __ldg(r0); q.t10 = __ldg(r0 + 1); q.t01 = __ldg(r1); q.t11 = __ldg(r1 + 1);
  return q;
}
// the part of a step after its tap: 0 = go on, 2 = hit
ADEV int ssrEvalFast(const FrameParams& P, SsrRayFast& M, const SsrTap& q, float cu, float cv, float fi, V3 worldPos, V3 rayDir, V3& hitPos, V3& hitNormal) {
  const float top = fmaf(q.t10 - q.t00, q.fx, q.t00), bot = fmaf(q.t11 - q.t01, q.fx, q.t01);
  const float dRaw = fmaf(bot - top, q.fy, top);
  const float den = fmaf(dRaw, 1000.0f - 0.01f, -1000.0f) * fmaf(fi, M.dS, M.S0);
  const float projDen = fmaf(M.cP, den, fmaf(fi, M.daP, M.aP0));
  const bool flips = projDen * M.prev <= 0.0f;
  M.prev = projDen;
  if (flips) {
    const float k = (1000.0f * 0.01f) * rcpf(den);
    const V3 camMinusPos = mk3(P.g.inverseView[12], P.g.inverseView[13], P.g.inverseView[14]) - worldPos;
    const V3 wd = mk3(fmaf(P.ssrWu[0], cu, fmaf(P.ssrWv[0], cv, P.ssrW0[0])), fmaf(P.ssrWu[1], cu, fmaf(P.ssrWv[1], cv, P.ssrW0[1])),
                      fmaf(P.ssrWu[2], cu, fmaf(P.ssrWv[2], cv, P.ssrW0[2])));
    const V3 vv = mk3(fmaf(wd.x, k, camMinusPos.x), fmaf(wd.y, k, camMinusPos.y), fmaf(wd.z, k, camMinusPos.z));
    const float along = dot3(vv, rayDir);
    if (along > 0.0f && along * along > (0.999f * 0.999f) * dot3(vv, vv)) {
      const V3 currentNormal = normalize3(xyz(bilinear<FmtRGBA16F, AddrClamp>(P.normal, cu, cv)));
      if (dot3(currentNormal, rayDir) < 0.0f) {
        hitPos = worldPos + vv;
        hitNormal = currentNormal;
        return 2;
      }
    }
  }
  return 0;
}
ADEV int ssrRayFastStep2(const FrameParams& P, SsrRayFast& M, V3 worldPos, V3 rayDir, V3& hitPos, V3& hitNormal) {
  const float cuA = M.cu + M.stepX, cvA = M.cv + M.stepY, fiA = M.fi + 1.0f;
  const float cuB = cuA + M.stepX, cvB = cvA + M.stepY, fiB = fiA + 1.0f;
  M.cu = cuA; M.cv = cvA; M.fi = fiA;
  if (fiA > M.fSafe) {
    if (fiA > 128.0f || outside01(cuA, cvA)) return 1;
  }
  bool endB = false;
  if (fiB > M.fSafe) endB = fiB > 128.0f || outside01(cuB, cvB);
  const SsrTap qA = ssrTapFast(P, cuA, cvA);
  const SsrTap qB = ssrTapFast(P, endB ? cuA : cuB, endB ? cvA : cvB); // a step that will not be taken re-reads A's footprint
  int r = ssrEvalFast(P, M, qA, cuA, cvA, fiA, worldPos, rayDir, hitPos, hitNormal);
  if (r) return r;
  M.cu = cuB; M.cv = cvB; M.fi = fiB;
  if (endB) return 1;
  return ssrEvalFast(P, M, qB, cuB, cvB, fiB, worldPos, rayDir, hitPos, hitNormal);
}
#ifndef ALTHEA_SSR_STEPS2
#define ALTHEA_SSR_STEPS2 0 // measured: 1.43 ms (48 registers, spills) / 1.39 ms (64 registers) against 1.36 ms one step at a time
#endif
#endif

#ifndef ALTHEA_SSR_MIN_BLOCKS
#define ALTHEA_SSR_MIN_BLOCKS 4 // 64 registers: the march is latency-bound, a fourth resident CTA is worth more than the registers
#endif
// The march and the shading of its hits are two kernels. The march is bound by the latency of its dependent depth taps
// (long_scoreboard, profiles/r2_frame_full.md): without the 16-light shading of a hit (SSR.frag:55-78) in the same kernel it needs
// 48 registers instead of 64 (five CTAs per SM instead of four; six or eight were measured no faster). Hits (~18 % of the pixels, scattered over most
// warps) are appended to a list and shaded packed, every lane of every warp busy, by ssr_shade_hits_kernel.
#ifndef ALTHEA_SSR_MARCH_MIN_BLOCKS
#define ALTHEA_SSR_MARCH_MIN_BLOCKS 5
#endif
#ifndef ALTHEA_SSR_TILE_W
#define ALTHEA_SSR_TILE_W 16 // measured at 4K: 8 x 4 pixels per warp 1.40 ms, 16 x 2 1.37 ms, 32 x 1 1.47 ms, 64-wide tiles 1.50 ms
#endif
constexpr int kSsrTileW = ALTHEA_SSR_TILE_W, kSsrTileH = 256 / kSsrTileW;
__global__ void __launch_bounds__(256, ALTHEA_SSR_MARCH_MIN_BLOCKS) ssr_capture_kernel(const __grid_constant__ FrameParams P) {
  // A warp is min(kSsrTileW, 32) x (32 / kSsrTileW) pixels of a kSsrTileW x kSsrTileH tile. The march is bound by the L1 data pipe as
  // much as by issue (72 % / 65 %, profiles/): a depth tap of a warp costs a wavefront per 128-byte line it touches. Fewer image
  // rows per warp mean fewer lines per tap but lanes whose marches end further apart: 16 x 2 is the measured optimum.
  const int x = blockIdx.x * kSsrTileW + (threadIdx.x % kSsrTileW);
  const int y = P.y0 + blockIdx.y * kSsrTileH + (threadIdx.x / kSsrTileW); // rows [y0, y1): the scissor (whole frame by default)
  const bool inside = x < P.W && y < P.y1;
  const float u = ((float)x + 0.5f) / (float)P.W, v = ((float)y + 0.5f) / (float)P.H;
  V4 normal4 = mk4(0.0f, 0.0f, 0.0f, 0.0f);
  if (inside) normal4 = FmtRGBA16F::load(P.normal, x, y);
  bool hit = false;
  float hcu = 0.0f, hcv = 0.0f;
  V3 hitPos = mk3(0.0f, 0.0f, 0.0f), hitNormal = hitPos, hitRay = hitPos;
  if (normal4.w != 0.0f) { // SSR.frag:136-141
    const float dOwn = __ldg(rowPtr<float>(P.depth, y) + x);
    const V3 worldPos = reconstructPosition(P, u, v, dOwn);
    const V3 normal = normalize3(xyz(normal4));
    const V3 rayDir = reflect3(normalize3(viewDirection(P, u, v)), normal);
    // raymarchGBuffer, SSR.frag:80-133
    V2 uvEnd = projectUv(P, worldPos + rayDir * 10000.0f);
    float dx = uvEnd.x - u, dy = uvEnd.y - v;
    float dl = sqrtf(dx * dx + dy * dy);
    const float stepX = (dx / dl) * 0.005f, stepY = (dy / dl) * 0.005f;
    const V3 perpRef = normalize3(cross3(cross3(rayDir, normal), rayDir));
    float cu = u, cv = v;
#ifdef ALTHEA_PARITY
    float prevProjection = 0.0f;
    for (int i = 0; i < 128; ++i) {
      cu += stepX;
      cv += stepY;
      if (outside01(cu, cv)) break;
      const DepthTap q = depthTapPadded(P, cu, cv);
      float dRaw = mixf(mixf(q.t00, q.t10, q.fx), mixf(q.t01, q.t11, q.fx), q.fy); // == bilinearR32F<AddrClamp>(P.depth, cu, cv)
      V3 currentPos = reconstructPosition(P, cu, cv, dRaw);
      V3 dir = normalize3(currentPos - worldPos);
      float currentProjection = dot3(dir, perpRef);
      float f = dot3(dir, rayDir);
      if (currentProjection * prevProjection <= 0.0f && f > 0.999f && i > 0) {
        V3 currentNormal = normalize3(xyz(bilinear<FmtRGBA16F, AddrClamp>(P.normal, cu, cv)));
        if (dot3(currentNormal, rayDir) < 0.0f) {
          hit = true; hitPos = currentPos; hitNormal = currentNormal;
          break;
        }
      }
      prevProjection = currentProjection;
    }
#else
    SsrRayFast M;
    if (ssrRayFastSetup(P, u, v, worldPos, rayDir, perpRef, stepX, stepY, dl, M)) {
      for (;;) {
        const int r = ALTHEA_SSR_STEPS2 ? ssrRayFastStep2(P, M, worldPos, rayDir, hitPos, hitNormal) : ssrRayFastStep(P, M, worldPos, rayDir, hitPos, hitNormal);
        if (r) { hit = r == 2; break; }
      }
    }
    cu = M.cu; cv = M.cv;
#endif
    hcu = cu; hcv = cv; hitRay = rayDir;
  }
  // no hit: (0, 0, 0, 0), the clear the blend-on-write leaves (GraphicsPipeline.cpp:138-154)
  if (inside && !hit) rowPtrW<uint2>(P.refl.level[0], y)[x] = make_uint2(0u, 0u);
  // hits: one record each, appended with one atomic per warp
  const unsigned ballot = __ballot_sync(0xffffffffu, hit);
  if (ballot) {
    const int lane = threadIdx.x & 31, leader = __ffs(ballot) - 1;
    unsigned base = 0u;
    if (lane == leader) base = atomicAdd(P.ssrHitCount, (unsigned)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (hit) {
      float* r = P.ssrHits + (base + __popc(ballot & ((1u << lane) - 1u)));
      const unsigned cap = P.ssrHitCap;
      r[0] = __int_as_float(x | (y << 16));
      r[cap] = hcu; r[2 * cap] = hcv;
      r[3 * cap] = hitPos.x; r[4 * cap] = hitPos.y; r[5 * cap] = hitPos.z;
      r[6 * cap] = hitNormal.x; r[7 * cap] = hitNormal.y; r[8 * cap] = hitNormal.z;
      r[9 * cap] = hitRay.x; r[10 * cap] = hitRay.y; r[11 * cap] = hitRay.z;
    }
  }
}

#ifndef ALTHEA_PARITY
// ---- fast build: the same march on persistent warps whose idle lanes are refilled (warp ballots) ----------------------------
// The march of a pixel ends after 1 .. 128 steps (hit, screen border, sky pixels never start): with one pixel per thread a warp
// iterates as long as its slowest lane and the lanes that finished ride along. Here every warp pulls pixels from a queue: chunks
// of 16 x 8 pixels in tile order from one global counter, one pixel per idle lane, whenever a ballot shows at least
// kSsrRefillIdle lanes idle (the setup of a pixel runs with only the refilled lanes active: worth batching). A lane that hits parks
// its record in the warp's shared-memory slots; parked hits are appended to the hit list with one atomic per warp at the next refill.
// Pixels that do not hit are not written at all: the launcher clears the rows first (the blend-on-write over the clear,
// GraphicsPipeline.cpp:138-154, leaves (0, 0, 0, 0) there).
#ifndef ALTHEA_SSR_REFILL
#define ALTHEA_SSR_REFILL 0
#endif
#ifndef ALTHEA_SSR_REFILL_IDLE
#define ALTHEA_SSR_REFILL_IDLE 8
#endif
constexpr int kSsrRefillIdle = ALTHEA_SSR_REFILL_IDLE;
constexpr int kSsrChunkW = 16, kSsrChunkH = 8, kSsrChunk = kSsrChunkW * kSsrChunkH;
__global__ void __launch_bounds__(256, ALTHEA_SSR_MARCH_MIN_BLOCKS) ssr_capture_refill_kernel(const __grid_constant__ FrameParams P) {
  __shared__ float parked[8][11][32]; // per warp and lane: cu, cv, hitPos, hitNormal, rayDir of a hit waiting for the next append
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned ltMask = (1u << lane) - 1u;
  float (*park)[32] = parked[warp];
  const int chunksX = (P.W + kSsrChunkW - 1) / kSsrChunkW;
  const int chunks = chunksX * ((P.y1 - P.y0 + kSsrChunkH - 1) / kSsrChunkH);
  unsigned* const chunkCounter = P.ssrHitCount + 1; // zeroed with the hit count by ssr_depth_pad_kernel
  int next = 0, end = 0; // pixels [next, end) of the warp's current chunk are still to be handed out (warp-uniform)
  int chunkX = 0, chunkY = 0;
  bool exhausted = false;
  bool active = false, hitParked = false;
  int pixel = 0; // x | y << 16 of the lane's current (or parked) pixel
  SsrRayFast M;
  V3 worldPos = mk3(0.0f, 0.0f, 0.0f), rayDir = worldPos;
  for (;;) {
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (__popc(act) <= 32 - kSsrRefillIdle && !exhausted || act == 0u) {
      // ---- parked hits first: one atomic for the warp, every parked lane writes its own record
      const unsigned hits = __ballot_sync(0xffffffffu, hitParked);
      if (hits) {
        const int leader = __ffs(hits) - 1;
        unsigned base = 0u;
        if (lane == leader) base = atomicAdd(P.ssrHitCount, (unsigned)__popc(hits));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (hitParked) {
          float* r = P.ssrHits + (base + __popc(hits & ltMask));
          const unsigned cap = P.ssrHitCap;
          r[0] = __int_as_float(pixel);
#pragma unroll
          for (int k = 0; k < 11; ++k) r[(k + 1) * cap] = park[k][lane];
          hitParked = false;
        }
      }
      if (exhausted) {
        if (act == 0u) break;
      } else {
        if (next >= end) { // the warp's next chunk
          unsigned c = 0u;
          if (lane == 0) c = atomicAdd(chunkCounter, 1u);
          c = __shfl_sync(0xffffffffu, c, 0);
          if (c >= (unsigned)chunks) {
            exhausted = true;
            continue;
          }
          chunkX = (int)(c % (unsigned)chunksX) * kSsrChunkW;
          chunkY = P.y0 + (int)(c / (unsigned)chunksX) * kSsrChunkH;
          next = 0;
          end = kSsrChunk;
        }
        const unsigned idle = ~act;
        const int rank = __popc(idle & ltMask);
        const int take = min(__popc(idle), end - next);
        if (!active && rank < take) {
          const int j = next + rank;
          const int x = chunkX + (j & (kSsrChunkW - 1)), y = chunkY + (j / kSsrChunkW);
          if (x < P.W && y < P.y1) {
            const V4 normal4 = FmtRGBA16F::load(P.normal, x, y);
            if (normal4.w != 0.0f) { // SSR.frag:136-141
              const float u = ((float)x + 0.5f) / (float)P.W, v = ((float)y + 0.5f) / (float)P.H;
              const float dOwn = __ldg(rowPtr<float>(P.depth, y) + x);
              worldPos = reconstructPosition(P, u, v, dOwn);
              const V3 normal = normalize3(xyz(normal4));
              rayDir = reflect3(normalize3(viewDirection(P, u, v)), normal);
              const V2 uvEnd = projectUv(P, worldPos + rayDir * 10000.0f); // raymarchGBuffer, SSR.frag:80-95
              const float dx = uvEnd.x - u, dy = uvEnd.y - v;
              const float dl = sqrtf(dx * dx + dy * dy);
              const V3 perpRef = normalize3(cross3(cross3(rayDir, normal), rayDir));
              pixel = x | (y << 16);
              active = ssrRayFastSetup(P, u, v, worldPos, rayDir, perpRef, (dx / dl) * 0.005f, (dy / dl) * 0.005f, dl, M);
            }
          }
        }
        next += take;
        continue; // vote again: the new pixels may be sky, the chunk may have ended before every idle lane had one
      }
    }
    if (active) {
      V3 hitPos, hitNormal;
      const int r = ssrRayFastStep(P, M, worldPos, rayDir, hitPos, hitNormal);
      if (r) {
        active = false;
        if (r == 2) {
          hitParked = true;
          park[0][lane] = M.cu; park[1][lane] = M.cv;
          park[2][lane] = hitPos.x; park[3][lane] = hitPos.y; park[4][lane] = hitPos.z;
          park[5][lane] = hitNormal.x; park[6][lane] = hitNormal.y; park[7][lane] = hitNormal.z;
          park[8][lane] = rayDir.x; park[9][lane] = rayDir.y; park[10][lane] = rayDir.z;
        }
      }
    }
  }
}
#endif

#ifndef ALTHEA_SSR_SHADE_MIN_BLOCKS
#define ALTHEA_SSR_SHADE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, ALTHEA_SSR_SHADE_MIN_BLOCKS) ssr_shade_hits_kernel(const __grid_constant__ FrameParams P) {
  const unsigned count = *P.ssrHitCount, cap = P.ssrHitCap;
  for (unsigned k = blockIdx.x * 256u + threadIdx.x; k < count; k += gridDim.x * 256u) {
    const float* r = P.ssrHits + k;
    const int xy = __float_as_int(r[0]);
    const V4 out = environmentLitSample(P, mk3(r[3 * cap], r[4 * cap], r[5 * cap]), r[cap], r[2 * cap], mk3(r[9 * cap], r[10 * cap], r[11 * cap]),
                                        mk3(r[6 * cap], r[7 * cap], r[8 * cap]));
    // blend-on-write over the (0,0,0,0) clear: rgb*a, a
    rowPtrW<uint2>(P.refl.level[0], xy >> 16)[xy & 0xffff] = packHalf4(mk4(out.x * out.w, out.y * out.w, out.z * out.w, out.w));
  }
}

// ---- SSR, round 2: sign test over plane records of the depth buffer ------------------------------------------------------
// A march step can only hit when the projection of the tapped surface on perpRef changes sign (or vanishes) between two
// consecutive taps (SSR.frag:118). The tap is reconstructPosition(uv, bilinear depth) = cam + D(uv) t, t the eye depth, so its
// projection is t c0 (w - L(uv)) with w = 1 / t = (far - dRaw (far - near)) / (far near), AFFINE in the raw depth: the tap's w is
// the bilinear combination of its four texels' w, and bilinear weights reproduce affine functions. With one record
// {alpha, beta, gamma, r} per block of depth texels, |w_k - (alpha + beta x_k + gamma y_k)| <= r for every texel the block
// answers for, the tap's w lies within r of the plane AT THE TAP, and sign(w - L) is known whenever |w_plane - L| exceeds r plus
// the ray's slack (rounding of the fp32 evaluation, depth-buffer cancellation). Steps whose two taps have known, equal signs are
// skipped without touching the depth buffer; every other step is evaluated exactly as before (the previous tap's projection is
// re-evaluated when it was skipped), so the hit mask is that of the plain march. Neighbouring pixels reflect in neighbouring
// directions, so a warp's taps share a few records: they are read through L1 (the whole frame's records take 150 KB at 4K).
__global__ void __launch_bounds__(256) ssr_planes_kernel(const __grid_constant__ FrameParams P) {
  const int lane = threadIdx.x & 31;
  const int rec = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (rec >= kSsrPlaneStride * kSsrPlaneRows) return;
  const int S = 1 << P.ssrPlaneShift;
  // record (bx + 1, by + 1) of the array is block (bx, by): a tap up to half a texel left of / above the image lands in column /
  // row 0, which cannot decide (the march's masked block index must stay inside the array for every in-screen tap)
  const int bx = rec % kSsrPlaneStride - 1, by = rec / kSsrPlaneStride - 1;
  float4* out = const_cast<float4*>(P.ssrPlanes) + rec;
  const float inf = __int_as_float(0x7f800000);
  const int X0 = bx * S, Y0 = by * S;
  if (bx < 0 || by < 0 || X0 >= P.W || Y0 >= P.H) {
    if (lane == 0) *out = make_float4(0.0f, 0.0f, 0.0f, inf);
    return;
  }
  // texels the record answers for: the block and an apron: the march assigns a tap to a block by x / S rounded to 1 / 16
  // (S / 32 texels) and y / S to 2^-9 or finer, its tap coordinate is within 1e-3 texels of the restatement's, and the
  // footprint is the two texels from floor(x); inside the image (a footprint that clamps at the border repeats a covered texel)
  const int xlo = max(X0 - 2 - (S >> 5), 0), xhi = min(X0 + S + 1, P.W - 1), ylo = max(Y0 - 2, 0), yhi = min(Y0 + S + 1, P.H - 1);
  const int nx = xhi - xlo + 1, ny = yhi - ylo + 1;
  auto recip = [&](int x, int y) { // (far - dRaw (far - near)) / (far near), ReconstructPosition.glsl:6-8
    const float dRaw = __ldg(rowPtr<float>(P.depth, y) + x);
    return fmaf(dRaw, -(1000.0f - 0.01f), 1000.0f) * (1.0f / (1000.0f * 0.01f));
  };
  const int xc = (xlo + xhi) >> 1, yc = (ylo + yhi) >> 1;
  const float beta = nx > 1 ? (recip(xhi, yc) - recip(xlo, yc)) / (float)(xhi - xlo) : 0.0f;
  const float gamma = ny > 1 ? (recip(xc, yhi) - recip(xc, ylo)) / (float)(yhi - ylo) : 0.0f;
  const float alpha = recip(xc, yc) - (beta * (float)xc + gamma * (float)yc);
  float rlo = inf, rhi = -inf, wmax = 0.0f;
  bool ok = true;
  // texel k = lane, lane + 32, ... of the nx x ny box, row by row: the (column, row) pair is stepped, not divided out
  const int q32 = 32 / nx, r32 = 32 - q32 * nx;
  int kx = lane % nx, ky = lane / nx;
  for (int k = lane; k < nx * ny; k += 32) {
    const int x = xlo + kx, y = ylo + ky;
    kx += r32; ky += q32;
    if (kx >= nx) { kx -= nx; ky += 1; }
    const float w = recip(x, y);
    const float res = w - fmaf(beta, (float)x, fmaf(gamma, (float)y, alpha));
    ok = ok && (w > 0.0f) && (res == res) && (w < inf);
    rlo = fminf(rlo, res);
    rhi = fmaxf(rhi, res);
    wmax = fmaxf(wmax, w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rlo = fminf(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
    rhi = fmaxf(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
    wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    ok = __shfl_xor_sync(0xffffffffu, (int)ok, o) && ok;
  }
  if (lane == 0) {
    const float mid = 0.5f * (rlo + rhi);
    const float a2 = alpha + mid;
    float r = 0.5f * (rhi - rlo) + 4.0e-7f * (fabsf(mid) + fabsf(rlo) + fabsf(rhi));
    // a footprint that clamps at the image border repeats a texel: its weights no longer reproduce the plane, by up to half a
    // texel of slope; the march's tap coordinate is within 1e-3 texels of the restatement's
    const bool border = X0 == 0 || Y0 == 0 || X0 + S >= P.W || Y0 + S >= P.H;
    r += (border ? 0.51f : 0.001f) * (fabsf(beta) + fabsf(gamma));
    // rounding of the plane's evaluation here and in the march; the depth buffer's cancellation: the restatement's fp32 lerps
    // of the raw depth carry 4 ulp of 1.0, times (far - near) / (far near)
    r += 4.8e-7f * (fabsf(a2) + fabsf(alpha) + fabsf(beta) * (float)(xhi + 1) + fabsf(gamma) * (float)(yhi + 1));
    r += 2.4e-7f * wmax + 3.0e-5f;
    r *= 1.000001f;
    const bool fin = ok && isfinite(a2) && isfinite(beta) && isfinite(gamma) && isfinite(r);
    *out = fin ? make_float4(a2, beta, gamma, r) : make_float4(0.0f, 0.0f, 0.0f, inf);
  }
}

__global__ void __launch_bounds__(256, ALTHEA_SSR_MIN_BLOCKS) ssr_capture_skip_kernel(const __grid_constant__ FrameParams P) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15);
  const int y = P.y0 + blockIdx.y * 16 + (threadIdx.x >> 4); // rows [y0, y1): the scissor (whole frame by default)
  const bool inside = x < P.W && y < P.y1;
  const float u = ((float)x + 0.5f) / (float)P.W, v = ((float)y + 0.5f) / (float)P.H;
  V4 normal4 = mk4(0.0f, 0.0f, 0.0f, 0.0f);
  if (inside) normal4 = FmtRGBA16F::load(P.normal, x, y);
  V4 out = mk4(0.0f, 0.0f, 0.0f, 0.0f);
  if (normal4.w != 0.0f) { // SSR.frag:136-141
    const float dOwn = __ldg(rowPtr<float>(P.depth, y) + x);
    const V3 worldPos = reconstructPosition(P, u, v, dOwn);
    const V3 normal = normalize3(xyz(normal4));
    const V3 rayDir = reflect3(normalize3(viewDirection(P, u, v)), normal);
    // raymarchGBuffer, SSR.frag:80-133
    V2 uvEnd = projectUv(P, worldPos + rayDir * 10000.0f);
    float dx = uvEnd.x - u, dy = uvEnd.y - v;
    float dl = sqrtf(dx * dx + dy * dy);
    const float stepX = (dx / dl) * 0.005f, stepY = (dy / dl) * 0.005f;
    const V3 perpRef = normalize3(cross3(cross3(rayDir, normal), rayDir));
    // ---- constants of the sign test: L(x, y) = -dot(D(x, y), perpRef) / c0 in tap coordinates x = u W - 0.5, y = v H - 0.5
    const V3 camMinusPos = mk3(P.g.inverseView[12], P.g.inverseView[13], P.g.inverseView[14]) - worldPos;
    float Lc, Lx, Ly, rayConst;
    bool c0neg;
    {
      const float c0 = dot3(camMinusPos, perpRef);
      float invC0;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(invC0) : "f"(c0));
      const float aInv = fabsf(invC0);
      c0neg = c0 < 0.0f;
      Lc = -dot3(mk3(P.ssaoDc[0], P.ssaoDc[1], P.ssaoDc[2]), perpRef) * invC0;
      Lx = -dot3(mk3(P.ssaoDx[0], P.ssaoDx[1], P.ssaoDx[2]), perpRef) * invC0;
      Ly = -dot3(mk3(P.ssaoDy[0], P.ssaoDy[1], P.ssaoDy[2]), perpRef) * invC0;
      const float Labs = fabsf(Lc) + fabsf(Lx) * P.Wf + fabsf(Ly) * P.Hf;
      const float posL1 = fabsf(worldPos.x) + fabsf(worldPos.y) + fabsf(worldPos.z);
      const float kappa = (kNu * (2.0f * P.ssaoCamL1 + posL1) + kTiny) * aInv;
      // the margin rule of the SSAO sign test (same projection, same model), the rounding of L, the tap coordinate's 1e-3 texels
      rayConst = 1.0102f * fmaf(kappa, Labs, kNu * P.ssaoDmax1 * aInv) + 9.6e-7f * P.ssaoDmag * aInv + 4.8e-7f * Labs + 1e-3f * (fabsf(Lx) + fabsf(Ly));
      if (!(kappa <= 0.01f)) rayConst = __int_as_float(0x7f800000);
    }
    const float Sf = (float)(1 << P.ssrPlaneShift), invS = 1.0f / Sf;
    const float sxPx = stepX * P.Wf, syPx = stepY * P.Hf; // the step in texels, and taps per texel along each axis
    const float rsx = 1.0f / sxPx, rsy = 1.0f / syPx;
    const char* planeBytes = reinterpret_cast<const char*>(P.ssrPlanes);
    unsigned nSkip = 0u, nSpan = 0u, nExact = 0u, nUndecided = 0u, nWarpExact = 0u;
    bool hit = false; // shaded after the march, see ssr_capture_kernel
    V3 hitPos = worldPos, hitNormal = normal;
    float cu = u, cv = v, pu = u, pv = v;
    int prevCls = 0;       // sign class of the previous tap's projection: +-1 known, 0 unknown
    bool prevExact = true; // prevProjection is the previous tap's projection as the march computes it
#ifdef ALTHEA_PARITY
    float prevProjection = 0.0f;
#else
    float prevProjection = __int_as_float(0x7fc00000); // carries the i > 0 test, see ssr_capture_kernel
    const V3 W0 = mk3(P.ssrW0[0], P.ssrW0[1], P.ssrW0[2]), Wu = mk3(P.ssrWu[0], P.ssrWu[1], P.ssrWu[2]), Wv = mk3(P.ssrWv[0], P.ssrWv[1], P.ssrWv[2]);
    const float cR = dot3(camMinusPos, rayDir), cP = dot3(camMinusPos, perpRef);
    // the fast march's folded step (ssr_capture_kernel) at tap coordinates (tu, tv): projection, and what the cone test needs
    auto fastTap = [&](float tu, float tv, float& along, float& len2, float& k) {
      const DepthTap q = depthTapPadded(P, tu, tv);
      const float top = fmaf(q.t10 - q.t00, q.fx, q.t00), bot = fmaf(q.t11 - q.t01, q.fx, q.t01);
      const float dRaw = fmaf(bot - top, q.fy, top);
      const V3 wd = mk3(fmaf(Wu.x, tu, fmaf(Wv.x, tv, W0.x)), fmaf(Wu.y, tu, fmaf(Wv.y, tv, W0.y)), fmaf(Wu.z, tu, fmaf(Wv.z, tv, W0.z)));
      const float den = fmaf(dRaw, 1000.0f - 0.01f, -1000.0f) * fmaf(P.ssrS[1], tu, fmaf(P.ssrS[2], tv, P.ssrS[0]));
      k = (1000.0f * 0.01f) * rcpf(den);
      along = fmaf(k, dot3(wd, rayDir), cR);
      const V3 vv = mk3(fmaf(wd.x, k, camMinusPos.x), fmaf(wd.y, k, camMinusPos.y), fmaf(wd.z, k, camMinusPos.z));
      len2 = dot3(vv, vv);
      return fmaf(k, dot3(wd, perpRef), cP);
    };
#endif
    for (int i = 0; i < 128; ++i) {
      cu += stepX;
      cv += stepY;
      if (outside01(cu, cv)) break;
      // ---- sign class of this tap from the plane records
      int cls = 0;
      const float tx = fmaf(cu, P.Wf, -0.5f), ty = fmaf(cv, P.Hf, -0.5f);
      const float vx = fmaf(tx, invS, 786433.0f), vy = fmaf(ty, invS, 6145.0f); // 1.5 * 2^19 + 1, 1.5 * 2^12 + 1: (block + 1) << 4, << 11
      // the masked offset stays inside the 128 x 128 records whatever the coordinate (a NaN coordinate is not "outside": it
      // lands on some record, and NaN never compares as decided)
      const unsigned bxs = __float_as_uint(vx) & 0x7f0u, bys = __float_as_uint(vy) & 0x3f800u;
      const float4 rec = __ldg(reinterpret_cast<const float4*>(planeBytes + (bxs | bys)));
      const float d = fmaf(rec.y, tx, fmaf(rec.z, ty, rec.x)) - fmaf(Lx, tx, fmaf(Ly, ty, Lc));
      const float slack = rec.w + rayConst;
      if (fabsf(d) > slack) cls = ((d < 0.0f) != c0neg) ? -1 : 1;
      if (cls != 0 && cls == prevCls) { // both projections known, nonzero, of one sign: SSR.frag:118 cannot hold
        // Inside one block w_plane - L is affine along the ray: if the last tap before the ray leaves the block (and the image)
        // is decided with the same sign, so is every tap in between: they are skipped with their coordinate updates only.
        const float ex0 = (float)((int)(bxs >> 4) - 1) * Sf, ey0 = (float)((int)(bys >> 11) - 1) * Sf; // the record's own block
        const float toX = sxPx > 0.0f ? fminf(ex0 + Sf, P.Wf - 1.0f) - tx : fmaxf(ex0, 0.0f) - tx;
        const float toY = syPx > 0.0f ? fminf(ey0 + Sf, P.Hf - 1.0f) - ty : fmaxf(ey0, 0.0f) - ty;
        // taps that certainly stay inside: one less than fit (accumulated coordinates drift by < 0.05 texels over a march)
        const float mf = fminf(fminf(toX * rsx, toY * rsy), (float)(126 - i)) - 1.0f;
        int m = mf >= 1.0f ? (int)mf : 0; // NaN (a zero step component times an infinite reciprocal): no span
        if (m > 0) {
          const float dLast = fmaf((float)m, fmaf(rec.y - Lx, sxPx, (rec.z - Ly) * syPx), d);
          if (!(fabsf(dLast) > slack && dLast * d > 0.0f)) m = 0;
        }
        for (int j = 0; j < m; ++j) { cu += stepX; cv += stepY; }
        i += m;
        if (P.gatherCounter) { nSkip += 1u + (unsigned)m; nSpan += m > 0 ? 1u : 0u; }
        prevExact = false;
        pu = cu; pv = cv;
        continue;
      }
#ifdef ALTHEA_PARITY
      if (!prevExact) {
        const DepthTap q = depthTapPadded(P, pu, pv);
        const float dRaw = mixf(mixf(q.t00, q.t10, q.fx), mixf(q.t01, q.t11, q.fx), q.fy);
        prevProjection = dot3(normalize3(reconstructPosition(P, pu, pv, dRaw) - worldPos), perpRef);
      }
      const DepthTap q = depthTapPadded(P, cu, cv);
      float dRaw = mixf(mixf(q.t00, q.t10, q.fx), mixf(q.t01, q.t11, q.fx), q.fy); // == bilinearR32F<AddrClamp>(P.depth, cu, cv)
      V3 currentPos = reconstructPosition(P, cu, cv, dRaw);
      V3 dir = normalize3(currentPos - worldPos);
      float currentProjection = dot3(dir, perpRef);
      float f = dot3(dir, rayDir);
      if (currentProjection * prevProjection <= 0.0f && f > 0.999f && i > 0) {
        V3 currentNormal = normalize3(xyz(bilinear<FmtRGBA16F, AddrClamp>(P.normal, cu, cv)));
        if (dot3(currentNormal, rayDir) < 0.0f) {
          hit = true; hitPos = currentPos; hitNormal = currentNormal;
          break;
        }
      }
#else
      float along, len2, k;
      if (!prevExact) prevProjection = fastTap(pu, pv, along, len2, k);
      const float currentProjection = fastTap(cu, cv, along, len2, k);
      if (currentProjection * prevProjection <= 0.0f && along > 0.0f && along * along > (0.999f * 0.999f) * len2) {
        V3 currentNormal = normalize3(xyz(bilinear<FmtRGBA16F, AddrClamp>(P.normal, cu, cv)));
        if (dot3(currentNormal, rayDir) < 0.0f) {
          const V3 wd = mk3(fmaf(Wu.x, cu, fmaf(Wv.x, cv, W0.x)), fmaf(Wu.y, cu, fmaf(Wv.y, cv, W0.y)), fmaf(Wu.z, cu, fmaf(Wv.z, cv, W0.z)));
          const V3 vv = mk3(fmaf(wd.x, k, camMinusPos.x), fmaf(wd.y, k, camMinusPos.y), fmaf(wd.z, k, camMinusPos.z));
          hit = true; hitPos = worldPos + vv; hitNormal = currentNormal;
          break;
        }
      }
#endif
      if (P.gatherCounter) { nExact += 1u; nUndecided += cls == 0 ? 1u : 0u; nWarpExact += (threadIdx.x & 31) == (__ffs(__activemask()) - 1) ? 1u : 0u; }
      prevProjection = currentProjection;
      prevExact = true;
      // known sign of an exact value large enough that its product with a decided projection cannot underflow
      prevCls = currentProjection > 1e-15f ? 1 : currentProjection < -1e-15f ? -1 : 0;
      pu = cu; pv = cv;
    }
    if (hit) out = environmentLitSample(P, hitPos, cu, cv, rayDir, hitNormal);
    if (P.gatherCounter) { // diagnostics: taps skipped, spans, taps evaluated exactly, of those undecided by the records, warp-level exact steps
      atomicAdd(P.gatherCounter, (unsigned long long)nSkip); atomicAdd(P.gatherCounter + 1, (unsigned long long)nSpan);
      atomicAdd(P.gatherCounter + 2, (unsigned long long)nExact); atomicAdd(P.gatherCounter + 3, (unsigned long long)(nUndecided | ((unsigned long long)nWarpExact << 32)));
    }
  }
  // blend-on-write over the (0,0,0,0) clear (GraphicsPipeline.cpp:138-154): rgb*a, a
  if (inside) rowPtrW<uint2>(P.refl.level[0], y)[x] = packHalf4(mk4(out.x * out.w, out.y * out.w, out.z * out.w, out.w));
}

// ---- glossy convolve ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) glossy_convolve_kernel(const __grid_constant__ ConvolveParams C) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15);
  const int y = C.y0 + blockIdx.y * 16 + (threadIdx.x >> 4); // dst rows [y0, y1)
  const int w = C.dst.w, h = C.dst.h;
  if (x >= w || y >= C.y1) return;
  const float u = (float)x / (float)w, v = (float)y / (float)h; // texelPos / size, no half texel
  const float resolution = (float)w;                            // width for both axes (SSRGlossyConvolve.comp:41)
  const float dirx = C.vertical ? 0.0f : 1.0f, diry = C.vertical ? 1.0f : 0.0f;
  const float a1x = (1.411764705882353f * dirx) / resolution, a1y = (1.411764705882353f * diry) / resolution;
  const float a2x = (3.2941176470588234f * dirx) / resolution, a2y = (3.2941176470588234f * diry) / resolution;
  const float a3x = (5.176470588235294f * dirx) / resolution, a3y = (5.176470588235294f * diry) / resolution;
  V4 color = mk4(0.0f, 0.0f, 0.0f, 0.0f);
  color = color + bilinear<FmtRGBA16F, AddrClamp>(C.src, u, v) * 0.1964825501511404f;
  color = color + bilinear<FmtRGBA16F, AddrClamp>(C.src, u + a1x, v + a1y) * 0.2969069646728344f;
  color = color + bilinear<FmtRGBA16F, AddrClamp>(C.src, u - a1x, v - a1y) * 0.2969069646728344f;
  color = color + bilinear<FmtRGBA16F, AddrClamp>(C.src, u + a2x, v + a2y) * 0.09447039785044732f;
  color = color + bilinear<FmtRGBA16F, AddrClamp>(C.src, u - a2x, v - a2y) * 0.09447039785044732f;
  color = color + bilinear<FmtRGBA16F, AddrClamp>(C.src, u + a3x, v + a3y) * 0.010381362401148057f;
  color = color + bilinear<FmtRGBA16F, AddrClamp>(C.src, u - a3x, v - a3y) * 0.010381362401148057f;
  rowPtrW<uint2>(C.dst, y)[x] = packHalf4(color);
}

// Same pass with the source footprint of the tile staged in shared memory by the TMA engine (cp.async.bulk, one bulk copy per
// source row, completion on an mbarrier: UBLKCP + SYNCS in SASS). A 7-tap blur reads every source texel ~14 times; from shared
// memory those re-reads cost no L1 tag lookups, and the global side becomes a few hundred fully coalesced row copies per CTA.
// The arithmetic (coordinates, tap order, lerps) is the direct kernel's, so the results are identical bit for bit.
// Preconditions (checked by the launcher, which otherwise uses the direct kernel): source width even and level base 16-byte
// aligned, so that every staged row segment starts and ends on a 16-byte boundary.
struct StagedTile { // the staged source window: columns [x0, x0 + cols), rows [y0, y0 + rows) of the source level
  const uint2* texels;
  int x0, y0, cols, rows;
};
ADEV V4 stagedLoad(const StagedTile& t, int i, int j) { // (i, j) already clamped to the image by the CLAMP_TO_EDGE rule
  const int ci = min(max(i - t.x0, 0), t.cols - 1), cj = min(max(j - t.y0, 0), t.rows - 1); // memory safety only: the window covers every tap
  return unpackHalf4(t.texels[cj * t.cols + ci]);
}
ADEV V4 bilinearStaged(const StagedTile& t, int w, int h, float u, float v) {
  BilinearSetup s = bilinearSetup<AddrClamp>(w, h, u, v);
  V4 t00 = stagedLoad(t, s.i0, s.j0), t10 = stagedLoad(t, s.i1, s.j0);
  V4 t01 = stagedLoad(t, s.i0, s.j1), t11 = stagedLoad(t, s.i1, s.j1);
  return mix4(mix4(t00, t10, s.fx), mix4(t01, t11, s.fx), s.fy);
}

template <int TW, int TH> __global__ void __launch_bounds__(256) glossy_convolve_staged_kernel(const __grid_constant__ ConvolveParams C) {
  extern __shared__ __align__(128) unsigned char convolveSmem[];
  __shared__ __align__(8) uint64_t bar;
  const int w = C.dst.w, h = C.dst.h, ws = C.src.w, hs = C.src.h;
  const int tx0 = blockIdx.x * TW, ty0 = C.y0 + blockIdx.y * TH;
  const int tx1 = min(tx0 + TW, w) - 1, ty1 = min(ty0 + TH, C.y1) - 1; // inclusive dst range of this tile
  const float resolution = (float)w;
  const float dirx = C.vertical ? 0.0f : 1.0f, diry = C.vertical ? 1.0f : 0.0f;
  const float a1x = (1.411764705882353f * dirx) / resolution, a1y = (1.411764705882353f * diry) / resolution;
  const float a2x = (3.2941176470588234f * dirx) / resolution, a2y = (3.2941176470588234f * diry) / resolution;
  const float a3x = (5.176470588235294f * dirx) / resolution, a3y = (5.176470588235294f * diry) / resolution;
  // conservative source window: the outermost taps of the tile's corner texels, one texel of slack for rounding, clamped to
  // the image (taps beyond the edge clamp onto it), columns widened to even bounds (16-byte row segments)
  StagedTile T;
  {
    const float uLo = (float)tx0 / (float)w - a3x, uHi = (float)tx1 / (float)w + a3x;
    const float vLo = (float)ty0 / (float)h - a3y, vHi = (float)ty1 / (float)h + a3y;
    int xLo = (int)floorf(uLo * (float)ws - 0.5f) - 1, xHi = (int)floorf(uHi * (float)ws - 0.5f) + 2;
    int yLo = (int)floorf(vLo * (float)hs - 0.5f) - 1, yHi = (int)floorf(vHi * (float)hs - 0.5f) + 2;
    xLo = max(xLo, 0) & ~1;
    xHi = min(xHi, ws - 1) | 1; // ws is even, so xHi | 1 <= ws - 1
    yLo = max(yLo, 0);
    yHi = min(yHi, hs - 1);
    T.x0 = xLo; T.y0 = yLo; T.cols = xHi - xLo + 1; T.rows = yHi - yLo + 1;
    T.texels = reinterpret_cast<const uint2*>(convolveSmem);
  }
  if (threadIdx.x == 0) mbarInit(&bar, 1);
  __syncthreads();
  if (threadIdx.x < 32) { // one warp issues the row copies (a few per lane), lane 0 arms the barrier with the byte count first
    const uint32_t rowBytes = (uint32_t)T.cols * 8u;
    if (threadIdx.x == 0) mbarExpectTx(&bar, rowBytes * (uint32_t)T.rows);
    __syncwarp();
    for (int r = threadIdx.x; r < T.rows; r += 32)
      bulkCopyG2S(convolveSmem + (size_t)r * rowBytes, rowPtr<uint2>(C.src, T.y0 + r) + T.x0, rowBytes, &bar);
  }
  mbarWait(&bar, 0);
  for (int k = threadIdx.x; k < TW * TH; k += 256) {
    const int x = tx0 + (k % TW), y = ty0 + (k / TW);
    if (x >= w || y >= C.y1) continue;
    const float u = (float)x / (float)w, v = (float)y / (float)h; // texelPos / size, no half texel
    V4 color = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    color = color + bilinearStaged(T, ws, hs, u, v) * 0.1964825501511404f;
    color = color + bilinearStaged(T, ws, hs, u + a1x, v + a1y) * 0.2969069646728344f;
    color = color + bilinearStaged(T, ws, hs, u - a1x, v - a1y) * 0.2969069646728344f;
    color = color + bilinearStaged(T, ws, hs, u + a2x, v + a2y) * 0.09447039785044732f;
    color = color + bilinearStaged(T, ws, hs, u - a2x, v - a2y) * 0.09447039785044732f;
    color = color + bilinearStaged(T, ws, hs, u + a3x, v + a3y) * 0.010381362401148057f;
    color = color + bilinearStaged(T, ws, hs, u - a3x, v - a3y) * 0.010381362401148057f;
    rowPtrW<uint2>(C.dst, y)[x] = packHalf4(color);
  }
}

#ifndef ALTHEA_PARITY
// ---- glossy convolve, fast build, exact 2 : 1 levels --------------------------------------------------------------------------
// For a source of exactly twice the target's size the pass is a FIXED filter: texelPos / size has no half texel, so the centre tap
// falls on the corner between source texels 2 p - 1 and 2 p (weights 1/2, 1/2) and the six offset taps land a constant distance
// from it (+-off_k ws / w texels along x, +-off_k hs / w rows along y: `resolution` is the width for both axes), the same fractional
// weights for every output texel. Across the filter direction every tap averages the two texels around the corner. The host folds
// the seven taps' bilinear weights into one coefficient per source offset (FirTable, double precision); the kernels below apply
// them: ~140 instructions per output texel instead of ~680 for seven general bilinear taps, no staging (rows are read through
// L1, neighbouring outputs share them). Same sums in a different order: within half an ulp of the RGBA16F store of the general
// kernel's result (tests), which the parity build keeps.
struct FirTable {
  int lo, hi;    // source offsets (relative to 2 p) with a non-zero coefficient: [lo, hi], inside [-12, 12]
  float c[25];   // coefficient of offset o at c[o + 12]
  int n;         // the non-zero ones again as a list (seven taps x two texels: at most 14)
  int o[14];
  float w[14];
};
ADEV V4 firPairAverage(const uint2* row, int c0, int c1) { // 0.5 t(c0) + 0.5 t(c1)
  const V4 a = unpackHalf4(__ldg(row + c0)), b = unpackHalf4(__ldg(row + c1));
  return mk4(0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z), 0.5f * (a.w + b.w));
}
// vertical passes (levels 1, 3): a thread owns R consecutive target rows of one column; every source row it reads feeds up to R of them
template <int R> __global__ void __launch_bounds__(256) glossy_fir_vertical_kernel(const __grid_constant__ ConvolveParams C, const __grid_constant__ FirTable T) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = C.y0 + (blockIdx.y * 8 + (threadIdx.x >> 5)) * R;
  if (x >= C.dst.w || y0 >= C.y1) return;
  const int ws = C.src.w, hs = C.src.h;
  const int c0 = max(2 * x - 1, 0), c1 = min(2 * x, ws - 1);
  V4 acc[R];
#pragma unroll
  for (int k = 0; k < R; ++k) acc[k] = mk4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 4
  for (int r = T.lo; r <= T.hi + 2 * (R - 1); ++r) { // (unrolled by four: eight row loads in flight per thread)
    const int row = min(max(2 * y0 + r, 0), hs - 1); // CLAMP_TO_EDGE on the tap's row
    const V4 v = firPairAverage(rowPtr<uint2>(C.src, row), c0, c1);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int o = r - 2 * k;
      if (o >= T.lo && o <= T.hi) { // warp-uniform
        const float c = T.c[o + 12];
        acc[k] = mk4(fmaf(c, v.x, acc[k].x), fmaf(c, v.y, acc[k].y), fmaf(c, v.z, acc[k].z), fmaf(c, v.w, acc[k].w));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < R; ++k)
    if (y0 + k < C.y1) rowPtrW<uint2>(C.dst, y0 + k)[x] = packHalf4(acc[k]);
}
// horizontal passes (levels 2, 4: a sixteenth of the frame and less)
__global__ void __launch_bounds__(256) glossy_fir_horizontal_kernel(const __grid_constant__ ConvolveParams C, const __grid_constant__ FirTable T) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = C.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= C.dst.w || y >= C.y1) return;
  const int ws = C.src.w, hs = C.src.h;
  const uint2* ra = rowPtr<uint2>(C.src, max(2 * y - 1, 0));
  const uint2* rb = rowPtr<uint2>(C.src, min(2 * y, hs - 1));
  V4 acc = mk4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
  for (int i = 0; i < 14; ++i) { // unrolled over the list: the 28 loads of a texel are independent
    if (i < T.n) {               // warp-uniform
      const int col = min(max(2 * x + T.o[i], 0), ws - 1);
      const V4 a = unpackHalf4(__ldg(ra + col)), b = unpackHalf4(__ldg(rb + col));
      const float h = 0.5f * T.w[i];
      acc = mk4(fmaf(h, a.x + b.x, acc.x), fmaf(h, a.y + b.y, acc.y), fmaf(h, a.z + b.z, acc.z), fmaf(h, a.w + b.w, acc.w));
    }
  }
  rowPtrW<uint2>(C.dst, y)[x] = packHalf4(acc);
}
// the seven taps of SSRGlossyConvolve.comp:41-52 as coefficients per source offset; false when a tap falls outside [-12, 12]
static bool firTable(const ConvolveParams& C, FirTable* T) {
  const double off[7] = {0.0, 1.411764705882353, -1.411764705882353, 3.2941176470588234, -3.2941176470588234, 5.176470588235294, -5.176470588235294};
  const double wt[7] = {0.1964825501511404, 0.2969069646728344, 0.2969069646728344, 0.09447039785044732, 0.09447039785044732, 0.010381362401148057, 0.010381362401148057};
  const double scale = (C.vertical ? (double)C.src.h : (double)C.src.w) / (double)C.dst.w; // source texels per unit of off / resolution
  double c[25] = {0.0};
  for (int t = 0; t < 7; ++t) {
    const double p = off[t] * scale - 0.5;
    const double o = floor(p), f = p - o;
    if (o < -12.0 || o + 1.0 > 12.0) return false;
    c[(int)o + 12] += wt[t] * (1.0 - f);
    c[(int)o + 13] += wt[t] * f;
  }
  T->lo = 12; T->hi = -12; T->n = 0;
  for (int i = 0; i < 25; ++i) {
    T->c[i] = (float)c[i];
    if (T->c[i] != 0.0f) {
      T->lo = std::min(T->lo, i - 12); T->hi = std::max(T->hi, i - 12);
      if (T->n == 14) return false;
      T->o[T->n] = i - 12; T->w[T->n] = T->c[i]; ++T->n;
    }
  }
  return T->lo <= T->hi;
}
#endif // !ALTHEA_PARITY

// ---- SSAO -----------------------------------------------------------------------------------------------------------
// computeSSAO (SSAO.glsl:31-84) counts, per pixel, the rays (24) whose 12-step screen-space march over the POSITION
// G-buffer finds an occluder. Each march step is one bilinear RGBA32F tap whose location is effectively random over a
// window of up to ~300 px at 4K, so the stage is bound by the L1 data pipe: a warp-level LDG pays one wavefront per distinct
// 128-byte line and here nearly every lane has its own (profiles/r1a: 20.4 wavefronts per LDG.128, L1 pipe 98.8 % busy).
//
// ssao_exact_kernel: the straight restatement, 4 LDG.128 per tap.
// ssao_kernel: filtered exact predicates over a packed proxy. A pre-pass (ssao_quads_kernel) stores, for
//   every bilinear footprint (ix, iy), ONE 32-byte record {p00 (3 x fp32), p10-p00, p01-p00, p11-p00 (9 x fp16), E (fp16)}
//   where E is an upper bound of the per-component error of the record against the four fp32 texels. A tap is then ONE
//   256-bit load (one sector, one line) giving the interpolated position to within E. Every decision of the march is a
//   threshold test; whenever the approximate value is farther from its threshold than the rigorous error bound
//   (storage error E + fp32 rounding slop), the decision is the one the exact arithmetic takes. Otherwise that tap is
//   re-evaluated from the fp32 texels with the restatement's own expression. Counts are therefore identical to
//   ssao_exact_kernel's, bit for bit, in both builds (checked at 4K in tests/).
// march parameter t = i / 12 and 1 - t (SSAO.glsl:47-48), folded at compile time (correctly rounded, as at run time)
__constant__ float kSsaoT[12] = {0.0f / 12.0f, 1.0f / 12.0f, 2.0f / 12.0f, 3.0f / 12.0f, 4.0f / 12.0f, 5.0f / 12.0f,
                                 6.0f / 12.0f, 7.0f / 12.0f, 8.0f / 12.0f, 9.0f / 12.0f, 10.0f / 12.0f, 11.0f / 12.0f};
__constant__ float kSsaoOneMinusT[12] = {1.0f - 0.0f / 12.0f, 1.0f - 1.0f / 12.0f, 1.0f - 2.0f / 12.0f, 1.0f - 3.0f / 12.0f,
                                         1.0f - 4.0f / 12.0f, 1.0f - 5.0f / 12.0f, 1.0f - 6.0f / 12.0f, 1.0f - 7.0f / 12.0f,
                                         1.0f - 8.0f / 12.0f, 1.0f - 9.0f / 12.0f, 1.0f - 10.0f / 12.0f, 1.0f - 11.0f / 12.0f};
// == mixf(a, b, i / 12.0f); the contraction is spelled out so that both SSAO kernels of a build march the same coordinates
ADEV float marchCoord(float a, float b, int i) {
#ifdef ALTHEA_PARITY
  return __fadd_rn(__fmul_rn(a, kSsaoOneMinusT[i]), __fmul_rn(b, kSsaoT[i]));
#else
  return fmaf(b, kSsaoT[i], __fmul_rn(a, kSsaoOneMinusT[i]));
#endif
}

ADEV int ssaoCountExact(const FrameParams& P, int px, int py, float u0, float v0, V3 worldPos, V3 normal) {
  HashRng rng;
  rng.sx = (uint32_t)px;
  rng.sy = (uint32_t)py;
  const TangentFrame tbn = localToWorld(normal);
  int ao = 0;
  for (int ray = 0; ray < 24; ++ray) {
    float x0 = rng.next(), x1 = rng.next(), x2 = rng.next();
    V3 rayDir = frameApply(tbn, normalize3(mk3(2.0f * x0 - 1.0f, 2.0f * x1 - 1.0f, x2)));
    V2 uvEnd = projectUv(P, worldPos + rayDir * 0.5f);
    V3 perpRef = normalize3(cross3(cross3(rayDir, normal), rayDir));
    V3 prevPos = worldPos;
    float prevProjection = 0.0f; // i == 0 taps the pixel's own texel: currentProjection == 0 exactly
    for (int i = 1; i < 12; ++i) {
      float cu = marchCoord(u0, uvEnd.x, i), cv = marchCoord(v0, uvEnd.y, i);
      if (outside01(cu, cv)) break;
      V3 currentPos = xyz(bilinear<FmtRGBA32F, AddrClamp>(P.position, cu, cv));
      float currentProjection = dot3(currentPos - worldPos, perpRef);
      float worldStep = length3(currentPos - prevPos);
      if (currentProjection * prevProjection < 0.0f && worldStep <= 2.0f) {
        V3 currentNormal = normalize3(xyz(bilinear<FmtRGBA16F, AddrClamp>(P.normal, cu, cv)));
        if (dot3(currentNormal, rayDir) < 0.0f) {
          ao += 1;
          break;
        }
      }
      prevPos = currentPos;
      prevProjection = currentProjection;
    }
  }
  return ao;
}

__global__ void __launch_bounds__(256) ssao_exact_kernel(const __grid_constant__ FrameParams P) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15);
  const int y = P.y0 + blockIdx.y * 16 + (threadIdx.x >> 4);
  if (x >= P.W || y >= P.y1) return;
  V4 position = FmtRGBA32F::load(P.position, x, y);
  uint8_t count = 255;
  if (position.w != 0.0f) {
    const float u = ((float)x + 0.5f) / (float)P.W, v = ((float)y + 0.5f) / (float)P.H;
    V3 normal = normalize3(xyz(FmtRGBA16F::load(P.normal, x, y)));
    count = (uint8_t)ssaoCountExact(P, x, y, u, v, xyz(position), normal);
  }
  rowPtrW<uint8_t>(P.ao, y)[x] = count;
}

// -- proxy records ------------------------------------------------------------------------------------------------------
// 32 bytes per bilinear footprint: p00 (3 x fp32) and five words that each carry TWO coefficients of
//   bilinear(fx, fy) = p00 + A fx + B fy + C fx fy,  A = p10 - p00, B = p01 - p00, C = (p11 - p10) - (p01 - p00):
//   word 3: A.x | B.y   word 4: A.y | B.z   word 5: A.z | C.x   word 6: B.x | C.y   word 7: T | C.z
// The low half is an fp16 (one HADD2.F32 to unpack). The high-half coefficient needs NO unpacking: the march uses the whole
// 32-bit word as an fp32 value; the pre-pass picks the upper 16 bits so that this value (low half included) is the closest
// one to the coefficient, i.e. a bf16-grade coefficient whose exact decoded value is known when the error bound is computed.
// T (rounded up) = 2 (sqrt(3) E + slop * max|p|_1 + tiny): everything of the sign-decision threshold that does not depend
// on the shaded pixel, E being the bound on the per-component error of the decoded polynomial against the four fp32 texels.
struct __align__(32) QuadRecord { uint32_t w[8]; };
constexpr float kPlaneErr = 3.0e-6f; // SSAO plane records (below): model error allowed per texel of a decidable block, relative to |p|_1 + |cam|_1

constexpr float kSqrt3Up = 1.7320509f;
// fp32 rounding slop, as a multiple of the coordinate magnitude in play: both the restatement's evaluation and the
// proxy's are within a few ulp of real arithmetic (two lerp levels, a subtraction, a 3-term dot product); 64 ulp is generous
constexpr float kRoundSlop = 64.0f * 1.1920929e-7f;

ADEV uint32_t halfBits(float v) { return (uint32_t)__half_as_ushort(__float2half_rn(v)); }
ADEV float lowHalfToFloat(uint32_t w) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
// word with low half `lo` whose value as an fp32 is as close to `target` as the 16 free bits allow (floats of one sign are
// ordered like their bit patterns, so the nearest pattern is at most half a bf16 step away)
ADEV uint32_t hiWord(float target, uint32_t lo) {
  const uint32_t t = __float_as_uint(target);
  uint32_t w = (t & 0xffff0000u) | lo;
  const int d = (int)lo - (int)(t & 0xffffu);
  if (d > 0x8000 && (t & 0x7fff0000u) != 0u) w -= 0x10000u;
  else if (d < -0x8000) w += 0x10000u; // may reach inf / NaN: caught by the finite check of the decoded value
  return w;
}

// position of texel (x, y) in mode D, exactly as reconstruct_position_kernel writes it (coordinates clamped to the image first)
ADEV float4 reconstructedTexel(const FrameParams& P, int x, int y) {
  x = AddrClamp::wrap(x, P.W);
  y = AddrClamp::wrap(y, P.H);
  float4 out = make_float4(0.0f, 0.0f, 0.0f, 0.0f); // the attachment's clear colour where nothing was drawn
  const V4 normal4 = FmtRGBA16F::load(P.normal, x, y);
  if (normal4.w != 0.0f) {
    const float u = ((float)x + 0.5f) / (float)P.W, v = ((float)y + 0.5f) / (float)P.H;
    const V3 p = reconstructPosition(P, u, v, __ldg(rowPtr<float>(P.depth, y) + x));
    out = make_float4(p.x, p.y, p.z, 1.0f);
  }
  return out;
}
// (Reconstructing the mode-D positions inside this pass, through a shared tile, instead of reading back what
// reconstruct_position_kernel wrote was built and measured: 0.164 ms against 0.057 + 0.093 ms for the two passes at 4K with a cold
// L2. Both are bound by their writes (52 B/px here); the read-back it saves comes from L2.)
__global__ void __launch_bounds__(256) ssao_quads_kernel(const __grid_constant__ FrameParams P) {
  const int qx = blockIdx.x * 32 + (threadIdx.x & 31); // record column = ix + 1
  const int qy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (qx > P.W || qy > P.H) return;
  const int i0 = AddrClamp::wrap(qx - 1, P.W), i1 = AddrClamp::wrap(qx, P.W);
  const int j0 = AddrClamp::wrap(qy - 1, P.H), j1 = AddrClamp::wrap(qy, P.H);
  const V4 p00 = FmtRGBA32F::load(P.position, i0, j0), p10 = FmtRGBA32F::load(P.position, i1, j0);
  const V4 p01 = FmtRGBA32F::load(P.position, i0, j1), p11 = FmtRGBA32F::load(P.position, i1, j1);
  if (P.ssaoPlaneStats && qx == 0 && qy == 0) { P.ssaoPlaneStats[0] = 0u; P.ssaoPlaneStats[1] = 0u; P.ssaoTileList[0] = 0u; }
  if (P.ssaoRecip && qx < P.W && qy < P.H) { // this thread's p11 is texel (qx, qy): its reciprocal eye depth for ssao_planes_kernel
    const V3 cam = mk3(P.ssaoCam[0], P.ssaoCam[1], P.ssaoCam[2]);
    const float t = dot3(xyz(p11) - cam, mk3(P.ssaoFwd[0], P.ssaoFwd[1], P.ssaoFwd[2]));
    const float w = __frcp_rn(t);
    // the texel against its model point cam + D(x, y) t
    const float fx = (float)qx, fy = (float)qy;
    const V3 D = mk3(fmaf(P.ssaoDy[0], fy, fmaf(P.ssaoDx[0], fx, P.ssaoDc[0])), fmaf(P.ssaoDy[1], fy, fmaf(P.ssaoDx[1], fx, P.ssaoDc[1])),
                     fmaf(P.ssaoDy[2], fy, fmaf(P.ssaoDx[2], fx, P.ssaoDc[2])));
    const float err = fmaxf(fabsf(fmaf(D.x, t, cam.x) - p11.x), fmaxf(fabsf(fmaf(D.y, t, cam.y) - p11.y), fabsf(fmaf(D.z, t, cam.z) - p11.z)));
    const float mag = (fabsf(p11.x) + fabsf(p11.y) + fabsf(p11.z)) + P.ssaoCamL1;
    const bool ok = (t > 0.0f) && (w < __int_as_float(0x7f800000)) && (err <= kPlaneErr * mag); // NaN anywhere fails a comparison
    // -1: the attachment's clear colour (0, 0, 0), which is what empty pixels hold: not a point of a view ray, but a footprint of
    // four of them interpolates to exactly zero, which the plane records answer for as a class of its own
    const bool cleared = p11.x == 0.0f && p11.y == 0.0f && p11.z == 0.0f;
    P.ssaoRecip[(size_t)qy * P.W + qx] = ok ? w : cleared ? -1.0f : __int_as_float(0x7fc00000);
  }
  const float b[3] = {p00.x, p00.y, p00.z};
  const float t10[3] = {p10.x, p10.y, p10.z}, t01[3] = {p01.x, p01.y, p01.z}, t11[3] = {p11.x, p11.y, p11.z};
  float A[3], B[3], C[3], U[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    A[c] = __fsub_rn(t10[c], b[c]);
    B[c] = __fsub_rn(t01[c], b[c]);
    U[c] = __fsub_rn(t11[c], t10[c]);
    C[c] = __fsub_rn(U[c], B[c]);
  }
  // low halves first (their bits are part of the high-half values), then the high-half words around them
  const uint32_t lBy = halfBits(B[1]), lBz = halfBits(B[2]), lCx = halfBits(C[0]), lCy = halfBits(C[1]), lCz = halfBits(C[2]);
  const uint32_t w3 = hiWord(A[0], lBy), w4 = hiWord(A[1], lBz), w5 = hiWord(A[2], lCx), w6 = hiWord(B[0], lCy);
  const float dA[3] = {__uint_as_float(w3), __uint_as_float(w4), __uint_as_float(w5)};
  const float dB[3] = {__uint_as_float(w6), lowHalfToFloat(lBy), lowHalfToFloat(lBz)};
  const float dC[3] = {lowHalfToFloat(lCx), lowHalfToFloat(lCy), lowHalfToFloat(lCz)};
  float E = 0.0f;
  bool finite = true;
  constexpr float kUlp = 6.0e-8f; // 2^-24 rounded up: |fl(a - b) - (a - b)| <= 2^-24 |fl(a - b)|
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    finite = finite && isfinite(b[c]) && isfinite(t10[c]) && isfinite(t01[c]) && isfinite(t11[c]) && isfinite(dA[c]) && isfinite(dB[c]) && isfinite(dC[c]);
    // storage error of each coefficient + rounding of the fp32 differences themselves; fx, fy, fx*fy <= 1, so the sum
    // bounds the error anywhere in the footprint
    float e = fmaxf(fabsf(__fsub_ru(dA[c], A[c])), fabsf(__fsub_rd(dA[c], A[c])));
    e = __fadd_ru(e, fmaxf(fabsf(__fsub_ru(dB[c], B[c])), fabsf(__fsub_rd(dB[c], B[c]))));
    e = __fadd_ru(e, fmaxf(fabsf(__fsub_ru(dC[c], C[c])), fabsf(__fsub_rd(dC[c], C[c]))));
    const float mags = __fadd_ru(__fadd_ru(fabsf(A[c]), __fmul_ru(2.0f, fabsf(B[c]))), __fadd_ru(fabsf(U[c]), fabsf(C[c])));
    e = __fadd_ru(e, __fmul_ru(mags, kUlp));
    // the restatement's own fp32 lerps round relative to the texel MAGNITUDES; covered here for texels much larger than the
    // interpolated value (the tap-side slop scales with the interpolated value only)
    e = __fadd_ru(e, __fmul_ru(mags, 9.6e-7f));
    E = fmaxf(E, e);
  }
  E = __fadd_ru(__fmul_ru(E, 1.0001f), 1e-30f);
  // bound of |p|_1 anywhere in the footprint: a convex combination of the corners, plus the proxy's own deviation
  float M = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const V4 q = k == 0 ? p00 : k == 1 ? p10 : k == 2 ? p01 : p11;
    M = fmaxf(M, __fadd_ru(__fadd_ru(fabsf(q.x), fabsf(q.y)), fabsf(q.z)));
  }
  M = __fadd_ru(M, __fmul_ru(3.0f, E));
  const float T = __fmul_ru(2.0f, __fadd_ru(__fmul_ru(E, kSqrt3Up), __fadd_ru(__fmul_ru(M, kRoundSlop), kTiny)));
  finite = finite && isfinite(T);
  // T rounded UP to the next bf16 step; +inf (or NaN, with the low half) => the tap is always re-evaluated exactly
  const uint32_t w7 = (finite ? ((__float_as_uint(T) >> 16) + 1u) << 16 : 0x7f800000u) | lCz;
  QuadRecord* row = reinterpret_cast<QuadRecord*>(const_cast<char*>(static_cast<const char*>(P.quads)) + (size_t)qy * P.quadPitch);
  uint4* dst = reinterpret_cast<uint4*>(row + qx);
  dst[0] = make_uint4(__float_as_uint(b[0]), __float_as_uint(b[1]), __float_as_uint(b[2]), w3);
  dst[1] = make_uint4(w4, w5, w6, w7);
}

ADEV QuadRecord loadQuad(const void* p) { // one 256-bit load: LDG.E.ENL2.256 on sm_100a
  QuadRecord r;
  asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
      : "l"(p));
  return r;
}

struct ProxyTap {
  V3 pos;       // interpolated position, within E per component of the real-arithmetic bilinear value
  float twoTol; // the record's T
};
// (u, v) has passed outside01: x is in [-0.5, W - 0.5], so floor(x) + 1 is a valid record column without clamping. A NaN
// coordinate converts to column 1 and yields NaN weights, hence a NaN position, which the caller re-evaluates exactly.
struct ProxyAddr { const QuadRecord* rec; float fx, fy; };
ADEV ProxyAddr proxyAddr(const FrameParams& P, float u, float v) {
  // same coordinate arithmetic as the exact tap (rule A1): identical footprint and weights, approximate texel values
  const float x = __fsub_rn(__fmul_rn(u, P.Wf), 0.5f), y = __fsub_rn(__fmul_rn(v, P.Hf), 0.5f);
  const int ix = __float2int_rd(x), iy = __float2int_rd(y); // floor; |x| < 2^24, so (float)ix == floorf(x) exactly
  ProxyAddr a;
  a.fx = x - (float)ix;
  a.fy = y - (float)iy;
  // record (ix + 1, iy + 1); quadsOrigin points at record (1, 1), so the index is iy * (W + 1) + ix (>= -(W + 2))
  a.rec = static_cast<const QuadRecord*>(P.quadsOrigin) + (iy * P.quadRow + ix);
  return a;
}
ADEV ProxyTap proxyEval(const QuadRecord& r, float fx, float fy) {
  const float ax = __uint_as_float(r.w[3]), ay = __uint_as_float(r.w[4]), az = __uint_as_float(r.w[5]), bx = __uint_as_float(r.w[6]);
  const float by = lowHalfToFloat(r.w[3]), bz = lowHalfToFloat(r.w[4]), cx = lowHalfToFloat(r.w[5]), cy = lowHalfToFloat(r.w[6]),
              cz = lowHalfToFloat(r.w[7]);
  ProxyTap t;
  t.pos.x = fmaf(fmaf(cx, fy, ax), fx, fmaf(bx, fy, __uint_as_float(r.w[0])));
  t.pos.y = fmaf(fmaf(cy, fy, ay), fx, fmaf(by, fy, __uint_as_float(r.w[1])));
  t.pos.z = fmaf(fmaf(cz, fy, az), fx, fmaf(bz, fy, __uint_as_float(r.w[2])));
  t.twoTol = __uint_as_float(r.w[7]);
  return t;
}

// the restatement's own expressions for one tap; kept out of line: it runs for a few percent of the taps only
struct ExactTap { V3 pos; float projection; };
__device__ __noinline__ ExactTap exactTap(const FrameParams& P, float cu, float cv, V3 worldPos, V3 perpRef) {
  ExactTap t;
  t.pos = xyz(bilinear<FmtRGBA32F, AddrClamp>(P.position, cu, cv));
  t.projection = dot3(t.pos - worldPos, perpRef);
  return t;
}
__device__ __noinline__ bool facesRay(const FrameParams& P, float cu, float cv, V3 rayDir) {
  V3 currentNormal = normalize3(xyz(bilinear<FmtRGBA16F, AddrClamp>(P.normal, cu, cv)));
  return dot3(currentNormal, rayDir) < 0.0f;
}

#ifndef ALTHEA_SSAO_UNROLL
#define ALTHEA_SSAO_UNROLL 2
#endif
constexpr int kSsaoUnroll = ALTHEA_SSAO_UNROLL;
// sign-decided value: v shrunk towards zero by its tolerance; 0 <=> the sign of the exact value is not known
ADEV float decided(float v, float twoTol) { return copysignf(fmaxf(fabsf(v) - twoTol, 0.0f), v); }
ADEV float decidedExact(float v) { return fabsf(v) > kTiny ? v : 0.0f; }

// COUNT: diagnostic instantiation that also counts the proxy records gathered (bench.py's gather-rate roofline)
template <bool COUNT> ADEV int ssaoCountFiltered(const FrameParams& P, int px, int py, float u0, float v0, V3 worldPos, V3 normal, unsigned& gathers) {
  HashRng rng;
  rng.sx = (uint32_t)px;
  rng.sy = (uint32_t)py;
  const TangentFrame tbn = localToWorld(normal);
  // the shaded pixel's share of the threshold: 2 * slop * |worldPos|_1 (a hair more: the sum below is rounded to nearest)
  const float posSlop2 = (2.002f * kRoundSlop) * (fabsf(worldPos.x) + fabsf(worldPos.y) + fabsf(worldPos.z));
  int ao = 0;
  for (int ray = 0; ray < 24; ++ray) {
    float x0 = rng.next(), x1 = rng.next(), x2 = rng.next();
    V3 rayDir = frameApply(tbn, normalize3(mk3(2.0f * x0 - 1.0f, 2.0f * x1 - 1.0f, x2)));
    V2 uvEnd = projectUv(P, worldPos + rayDir * 0.5f);
    V3 perpRef = normalize3(cross3(cross3(rayDir, normal), rayDir));
    // taps before the ray leaves the screen (SSAO.glsl:50). Tap i sits at a(1 - t) + b t with t = i / 12 and a = this pixel's
    // centre, strictly inside (0, 1): when the end point b is inside [0, 1] too, so is every rounded tap coordinate
    // (a(1 - t) + b t <= 1 - (1 - a)/12, far more than the two roundings), and no tap needs the test.
    int n = 12;
    if (outside01(uvEnd.x, uvEnd.y)) {
      n = 1;
      while (n < 12 && !outside01(marchCoord(u0, uvEnd.x, n), marchCoord(v0, uvEnd.y, n))) ++n;
    }
    // projection of the proxy position: dot(p, perpRef) - dot(worldPos, perpRef), within a few ulp of |p|_1 + |worldPos|_1
    // of the restatement's dot(p - worldPos, perpRef) (covered by the slop in T and posSlop2)
    const float projBias = fmaf(worldPos.z, perpRef.z, fmaf(worldPos.y, perpRef.y, worldPos.x * perpRef.x));
    // previous step: position, projection, its threshold (0 <=> the value IS the restatement's fp32 value) and its
    // sign-decided form
    V3 prevPos = worldPos;
    float prevProjection = 0.0f, prevTwoTol = 0.0f, prevDecided = 0.0f;
#pragma unroll kSsaoUnroll
    for (int i = 1; i < n; ++i) {
      const float cu = marchCoord(u0, uvEnd.x, i), cv = marchCoord(v0, uvEnd.y, i);
      const ProxyAddr pa = proxyAddr(P, cu, cv);
      const ProxyTap tap = proxyEval(loadQuad(pa.rec), pa.fx, pa.fy);
      if (COUNT) gathers += 1u;
      V3 curPos = tap.pos;
      float curProjection = fmaf(curPos.z, perpRef.z, fmaf(curPos.y, perpRef.y, fmaf(curPos.x, perpRef.x, -projBias)));
      float curTwoTol = tap.twoTol + posSlop2; // inf / NaN when the record is flagged
      float curDecided = decided(curProjection, curTwoTol);
      // common case: both signs decided and equal (no crossing). Everything else: the first tap (previous projection exactly
      // 0: the product is +-0 or NaN, never < 0), a crossing, or a value too close to 0 for the proxy to call
      if (!(__fmul_rn(curDecided, prevDecided) > 0.0f)) {
        if (i > 1) {
          bool flip;
          if (curDecided != 0.0f && prevDecided != 0.0f) {
            flip = (curProjection < 0.0f) != (prevProjection < 0.0f); // also a same-sign product that underflowed
          } else {
            const float pu = marchCoord(u0, uvEnd.x, i - 1), pv = marchCoord(v0, uvEnd.y, i - 1);
            if (curDecided == 0.0f) {
              if (COUNT) gathers += 65536u;
              ExactTap e = exactTap(P, cu, cv, worldPos, perpRef);
              curPos = e.pos; curProjection = e.projection; curTwoTol = 0.0f; curDecided = decidedExact(curProjection);
            }
            if (prevDecided == 0.0f && prevTwoTol != 0.0f) {
              if (COUNT) gathers += 65536u;
              ExactTap e = exactTap(P, pu, pv, worldPos, perpRef);
              prevPos = e.pos; prevProjection = e.projection; prevTwoTol = 0.0f; prevDecided = decidedExact(prevProjection);
            }
            // a value at most kTiny in magnitude can push the fp32 product into underflow: then both factors must be exact
            if (curDecided == 0.0f && prevTwoTol != 0.0f) {
              if (COUNT) gathers += 65536u;
              ExactTap e = exactTap(P, pu, pv, worldPos, perpRef);
              prevPos = e.pos; prevProjection = e.projection; prevTwoTol = 0.0f; prevDecided = decidedExact(prevProjection);
            }
            if (prevDecided == 0.0f && curTwoTol != 0.0f) {
              if (COUNT) gathers += 65536u;
              ExactTap e = exactTap(P, cu, cv, worldPos, perpRef);
              curPos = e.pos; curProjection = e.projection; curTwoTol = 0.0f; curDecided = decidedExact(curProjection);
            }
            flip = __fmul_rn(curProjection, prevProjection) < 0.0f;
          }
          if (flip) {
            // worldStep = length(currentPos - prevPos) <= 2.0; each threshold covers twice its position tolerance
            const float worldStep = length3(curPos - prevPos);
            const float tol = curTwoTol + prevTwoTol;
            bool near;
            if (worldStep + tol <= 2.0f) near = true;
            else if (worldStep - tol > 2.0f) near = false;
            else {
              if (curTwoTol != 0.0f) {
                if (COUNT) gathers += 65536u;
              ExactTap e = exactTap(P, cu, cv, worldPos, perpRef);
                curPos = e.pos; curProjection = e.projection; curTwoTol = 0.0f; curDecided = decidedExact(curProjection);
              }
              if (prevTwoTol != 0.0f) prevPos = exactTap(P, marchCoord(u0, uvEnd.x, i - 1), marchCoord(v0, uvEnd.y, i - 1), worldPos, perpRef).pos;
              near = length3(curPos - prevPos) <= 2.0f;
            }
            if (near && facesRay(P, cu, cv, rayDir)) {
              ao += 1;
              break;
            }
          }
        }
      }
      prevPos = curPos; prevProjection = curProjection; prevTwoTol = curTwoTol; prevDecided = curDecided;
    }
  }
  return ao;
}

// -- ray-depth proxy (round 1e) -------------------------------------------------------------------------------------------
// A position G-buffer written by a perspective camera holds, at texel (x, y), a point of the view ray through that texel:
//   p(x, y) = cam + D(x, y) t,  D(x, y) = Dc + Dx x + Dy y (the ray scaled to view z = -1),  t = dot(p - cam, fwd) (eye depth).
// So a footprint's four texels are described by four SCALARS, and the record shrinks from 32 to 16 bytes: t00 (fp32), the
// coefficients a = t10 - t00, b = t01 - t00, c = (t11 - t10) - (t01 - t00) of its bilinear interpolant (3 x fp16) and the sign
// threshold T (bf16, rounded up). Half the bytes per tap is what the march needs most: it is bound by divergent gathers, whose
// rate the device sustains ~15-40 % higher for 16-byte records (tools/microbench/gather.cu), because twice as many footprints
// fit in L1. (a and b keep full fp32 precision in the record's spare words; only the second-order term c is fp16.)
// The bilinear position is  cam + D00 Tb + Dx Tx + Dy Ty  with Tb = bilinear t, Tx = fx ((1 - fy) t10 + fy t11),
// Ty = fy ((1 - fx) t01 + fx t11) (D is affine, so the weights regroup), hence its projection on perpRef is
//   c0 + Tb (ac + au ix + av iy) + Tx au + Ty av,   c0 = dot(cam - pos, perpRef), ac/au/av = dot(Dc/Dx/Dy, perpRef) per ray:
// no position is ever formed in the common case. Model and truth are both bilinear in (fx, fy), so their difference peaks at a
// corner: the pre-pass measures E there against the real texels (this also absorbs G-buffers that are NOT camera-consistent:
// E grows, the record cannot decide, the tap is re-evaluated exactly). Decisions follow the same rule as for the position
// records: a sign is taken from the proxy only when |value| exceeds T + the shaded pixel's slop; every flip candidate and every
// undecided tap is re-evaluated from the fp32 texels with the restatement's own expression, so counts are bit-identical.
struct __align__(16) RayRecord { uint32_t w[4]; }; // t00 | a (fp32) | c (fp16), T (bf16 in the high half) | b (fp32)
// rounding slop of the two evaluations (the restatement's lerps + dot product, the model's), as a multiple of the magnitudes in
// play: worst case ~16 ulp by operation count, 32 budgeted on each of the record's and the pixel's share
constexpr float kRaySlop = 32.0f * 1.1920929e-7f;

__global__ void __launch_bounds__(256) ssao_rayquads_kernel(const __grid_constant__ FrameParams P) {
  const int qx = blockIdx.x * 32 + (threadIdx.x & 31); // record column = ix + 1
  const int qy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (qx > P.W || qy > P.H) return;
  RayRecord* dst = reinterpret_cast<RayRecord*>(const_cast<char*>(static_cast<const char*>(P.quads)) + (size_t)qy * P.quadPitch) + qx;
  RayRecord rec;
  const int ix = qx - 1, iy = qy - 1;
  bool ok = ix >= 0 && iy >= 0 && ix + 1 < P.W && iy + 1 < P.H; // border footprints clamp onto one texel: not affine, always exact
  float tt[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  V4 tex[4];
  const V3 cam = mk3(P.ssaoCam[0], P.ssaoCam[1], P.ssaoCam[2]), fwd = mk3(P.ssaoFwd[0], P.ssaoFwd[1], P.ssaoFwd[2]);
  if (ok) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      tex[k] = FmtRGBA32F::load(P.position, ix + (k & 1), iy + (k >> 1));
      tt[k] = dot3(xyz(tex[k]) - cam, fwd);
      ok = ok && isfinite(tex[k].x) && isfinite(tex[k].y) && isfinite(tex[k].z) && isfinite(tt[k]);
    }
  }
  // Empty pixels hold the clear colour (0, 0, 0, 0) (Src/DeferredRendering.cpp:101-102), which is no point of a view ray. A
  // footprint of four zero texels interpolates to exactly zero: its record is a marker (t00 = NaN pattern) and the march
  // substitutes the model coordinates of the world origin (FrameParams::ssaoOrigin). Footprints that mix empty and covered
  // texels (silhouettes against the sky) measure a large E below and are re-evaluated exactly.
  bool sky = ok;
#pragma unroll
  for (int k = 0; k < 4; ++k) sky = sky && tex[k].x == 0.0f && tex[k].y == 0.0f && tex[k].z == 0.0f;
  if (sky) {
    const V3 o = mk3(P.ssaoOrigin[0], P.ssaoOrigin[1], P.ssaoOrigin[2]);
    const float ex = fmaf(o.z, P.ssaoDy[0], fmaf(o.y, P.ssaoDx[0], fmaf(o.x, P.ssaoDc[0], cam.x)));
    const float ey = fmaf(o.z, P.ssaoDy[1], fmaf(o.y, P.ssaoDx[1], fmaf(o.x, P.ssaoDc[1], cam.y)));
    const float ez = fmaf(o.z, P.ssaoDy[2], fmaf(o.y, P.ssaoDx[2], fmaf(o.x, P.ssaoDc[2], cam.z)));
    const float E0 = fmaxf(fabsf(ex), fmaxf(fabsf(ey), fabsf(ez))); // how far the model puts the origin from (0, 0, 0)
    const float camL1s = fabsf(cam.x) + fabsf(cam.y) + fabsf(cam.z);
    const float Ts = __fmul_ru(2.0f, __fadd_ru(__fmul_ru(__fmul_ru(E0, 1.001f), kSqrt3Up), __fadd_ru(__fmul_ru(camL1s, kRaySlop), kTiny)));
    rec.w[0] = 0x7fc00000u;
    rec.w[1] = 0u;
    rec.w[3] = 0u;
    rec.w[2] = isfinite(Ts) ? ((__float_as_uint(Ts) >> 16) + 1u) << 16 : 0x7f800000u;
    *dst = rec;
    return;
  }
  float a = __fsub_rn(tt[1], tt[0]), b = __fsub_rn(tt[2], tt[0]), c = __fsub_rn(__fsub_rn(tt[3], tt[1]), b);
  const uint32_t hc = halfBits(c);
  c = lowHalfToFloat(hc); // as the march decodes it
  ok = ok && isfinite(a) && isfinite(b) && isfinite(c);
  float E = 0.0f, M = 0.0f;
  if (ok) {
    // decoded corner depths and the model position at each corner against the texel itself
    const float td[4] = {tt[0], tt[0] + a, tt[0] + b, ((tt[0] + a) + b) + c};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float fxk = (float)(ix + (k & 1)), fyk = (float)(iy + (k >> 1));
      const V3 D = mk3(fmaf(P.ssaoDy[0], fyk, fmaf(P.ssaoDx[0], fxk, P.ssaoDc[0])), fmaf(P.ssaoDy[1], fyk, fmaf(P.ssaoDx[1], fxk, P.ssaoDc[1])),
                       fmaf(P.ssaoDy[2], fyk, fmaf(P.ssaoDx[2], fxk, P.ssaoDc[2])));
      const V3 m = mk3(fmaf(D.x, td[k], cam.x), fmaf(D.y, td[k], cam.y), fmaf(D.z, td[k], cam.z));
      E = fmaxf(E, fmaxf(fabsf(m.x - tex[k].x), fmaxf(fabsf(m.y - tex[k].y), fabsf(m.z - tex[k].z))));
      M = fmaxf(M, __fadd_ru(__fadd_ru(fabsf(tex[k].x), fabsf(tex[k].y)), fabsf(tex[k].z)));
    }
  }
  // everything of the sign threshold that does not depend on the shaded pixel: model error (inflated: it was itself measured in
  // fp32), rounding slop of both evaluations on the magnitudes in play (texels and camera position), and the underflow guard
  const float camL1 = fabsf(cam.x) + fabsf(cam.y) + fabsf(cam.z);
  const float T = __fmul_ru(2.0f, __fadd_ru(__fmul_ru(__fmul_ru(E, 1.001f), kSqrt3Up), __fadd_ru(__fmul_ru(__fadd_ru(M, camL1), kRaySlop), kTiny)));
  ok = ok && isfinite(T);
  rec.w[0] = __float_as_uint(tt[0]);
  rec.w[1] = __float_as_uint(a);
  rec.w[3] = __float_as_uint(b);
  rec.w[2] = hc | (ok ? ((__float_as_uint(T) >> 16) + 1u) << 16 : 0x7f800000u); // T rounded UP to bf16; +inf => always exact
  *dst = rec;
}

struct RayConsts { float c0, ac, au, av; }; // per ray: dot products of (cam - pos), Dc, Dx, Dy with perpRef
// A tap in MODEL coordinates (A, B, C): position = cam + Dc A + Dx B + Dy C, with A = Tb, B = ix Tb + Tx, C = iy Tb + Ty
// (D00 = Dc + Dx ix + Dy iy regrouped). Projection and step length are evaluated in these coordinates; no position is formed.
struct RayTap { V3 abc; float projection, twoTol; };
ADEV RayTap rayTap(const FrameParams& P, const RayConsts& rc, float u, float v) {
  // same coordinate arithmetic as the exact tap (rule A1): identical footprint and weights
  const float x = __fsub_rn(__fmul_rn(u, P.Wf), 0.5f), y = __fsub_rn(__fmul_rn(v, P.Hf), 0.5f);
  const int ix = __float2int_rd(x), iy = __float2int_rd(y);
  const float fix = (float)ix, fiy = (float)iy;
  const float fx = x - fix, fy = y - fiy;
  const uint4 r = __ldg(static_cast<const uint4*>(P.quadsOrigin) + (iy * P.quadRow + ix));
  const float t00 = __uint_as_float(r.x), a = __uint_as_float(r.y), b = __uint_as_float(r.w), c = lowHalfToFloat(r.z);
  const float Tb = fmaf(fmaf(c, fy, a), fx, fmaf(b, fy, t00));  // bilinear eye depth
  const float Tx = fx * fmaf(b + c, fy, t00 + a);              // fx ((1 - fy) t10 + fy t11)
  const float Ty = fy * fmaf(a + c, fx, t00 + b);              // fy ((1 - fx) t01 + fx t11)
  RayTap t;
  const bool sky = r.x == 0x7fc00000u; // four empty texels: the world origin, in model coordinates
  t.abc = mk3(sky ? P.ssaoOrigin[0] : Tb, sky ? P.ssaoOrigin[1] : fmaf(fix, Tb, Tx), sky ? P.ssaoOrigin[2] : fmaf(fiy, Tb, Ty));
  t.projection = fmaf(t.abc.z, rc.av, fmaf(t.abc.y, rc.au, fmaf(t.abc.x, rc.ac, rc.c0)));
  t.twoTol = __uint_as_float(r.z); // bf16 T with the fp16 of c below it: a hair above T, never below
  return t;
}
ADEV V3 modelToWorld(const FrameParams& P, V3 m) {
  return mk3(fmaf(m.z, P.ssaoDy[0], fmaf(m.y, P.ssaoDx[0], fmaf(m.x, P.ssaoDc[0], P.ssaoCam[0]))),
             fmaf(m.z, P.ssaoDy[1], fmaf(m.y, P.ssaoDx[1], fmaf(m.x, P.ssaoDc[1], P.ssaoCam[1]))),
             fmaf(m.z, P.ssaoDy[2], fmaf(m.y, P.ssaoDx[2], fmaf(m.x, P.ssaoDc[2], P.ssaoCam[2]))));
}
// |Dc a + Dx b + Dy c| from the Gram matrix of (Dc, Dx, Dy) (FrameParams::ssaoGram: cc, xx, yy, 2cx, 2cy, 2xy)
ADEV float modelLength(const FrameParams& P, V3 d) {
  const float q = fmaf(d.x, fmaf(d.x, P.ssaoGram[0], fmaf(d.y, P.ssaoGram[3], d.z * P.ssaoGram[4])), fmaf(d.y, fmaf(d.y, P.ssaoGram[1], d.z * P.ssaoGram[5]), d.z * d.z * P.ssaoGram[2]));
  return sqrtf(fmaxf(q, 0.0f));
}

__device__ __noinline__ V3 perpRefOf(V3 rayDir, V3 normal) { return normalize3(cross3(cross3(rayDir, normal), rayDir)); }

template <bool COUNT> ADEV int ssaoCountRayProxy(const FrameParams& P, int px, int py, float u0, float v0, V3 worldPos, V3 normal, unsigned& gathers) {
  HashRng rng;
  rng.sx = (uint32_t)px;
  rng.sy = (uint32_t)py;
  const TangentFrame tbn = localToWorld(normal);
  const V3 cam = mk3(P.ssaoCam[0], P.ssaoCam[1], P.ssaoCam[2]);
  const V3 camMinusPos = cam - worldPos;
  const V3 Dc = mk3(P.ssaoDc[0], P.ssaoDc[1], P.ssaoDc[2]), Dx = mk3(P.ssaoDx[0], P.ssaoDx[1], P.ssaoDx[2]), Dy = mk3(P.ssaoDy[0], P.ssaoDy[1], P.ssaoDy[2]);
  // the shaded pixel's share of the threshold (the record's share carries the texels' and the camera's magnitudes)
  const float posSlop2 = (2.002f * kRaySlop) * ((fabsf(worldPos.x) + fabsf(worldPos.y) + fabsf(worldPos.z)) + (fabsf(cam.x) + fabsf(cam.y) + fabsf(cam.z)));
  int ao = 0;
  for (int ray = 0; ray < 24; ++ray) {
    float x0 = rng.next(), x1 = rng.next(), x2 = rng.next();
    V3 rayDir = frameApply(tbn, normalize3(mk3(2.0f * x0 - 1.0f, 2.0f * x1 - 1.0f, x2)));
    V2 uvEnd = projectUv(P, worldPos + rayDir * 0.5f);
    int n = 12; // taps before the ray leaves the screen (SSAO.glsl:50), as in ssaoCountFiltered
    if (outside01(uvEnd.x, uvEnd.y)) {
      n = 1;
      while (n < 12 && !outside01(marchCoord(u0, uvEnd.x, n), marchCoord(v0, uvEnd.y, n))) ++n;
    }
    // perpRef only lives in the four dot products the fast path needs; the rare exact path gets it back from perpRefOf (kept out
    // of line so the compiler cannot keep the vector in registers across the march to share it)
    RayConsts rc;
    {
      const V3 perpRef = normalize3(cross3(cross3(rayDir, normal), rayDir));
      rc.c0 = dot3(camMinusPos, perpRef); rc.ac = dot3(Dc, perpRef); rc.au = dot3(Dx, perpRef); rc.av = dot3(Dy, perpRef);
    }
    // previous step: projection, its threshold (0 <=> the value IS the restatement's fp32 value, and prevPos is valid), its
    // sign-decided form
    V3 prevPos = worldPos;
    float prevProjection = 0.0f, prevTwoTol = 0.0f, prevDecided = 0.0f;
#pragma unroll kSsaoUnroll
    for (int i = 1; i < n; ++i) {
      const float cu = marchCoord(u0, uvEnd.x, i), cv = marchCoord(v0, uvEnd.y, i);
      const RayTap tap = rayTap(P, rc, cu, cv);
      if (COUNT) gathers += 1u;
      V3 curPos = tap.abc; // MODEL coordinates while curTwoTol != 0, the exact world position once it is 0
      float curProjection = tap.projection;
      float curTwoTol = tap.twoTol + posSlop2; // inf / NaN when the record is flagged
      float curDecided = decided(curProjection, curTwoTol);
      if (!(__fmul_rn(curDecided, prevDecided) > 0.0f)) {
        if (i > 1) {
          bool flip;
          if (curDecided != 0.0f && prevDecided != 0.0f) {
            flip = (curProjection < 0.0f) != (prevProjection < 0.0f); // also a same-sign product that underflowed
          } else {
            const float pu = marchCoord(u0, uvEnd.x, i - 1), pv = marchCoord(v0, uvEnd.y, i - 1);
            if (curDecided == 0.0f) {
              if (COUNT) gathers += 65536u;
              ExactTap e = exactTap(P, cu, cv, worldPos, perpRefOf(rayDir, normal));
              curPos = e.pos; curProjection = e.projection; curTwoTol = 0.0f; curDecided = decidedExact(curProjection);
            }
            if (prevDecided == 0.0f && prevTwoTol != 0.0f) {
              if (COUNT) gathers += 65536u;
              ExactTap e = exactTap(P, pu, pv, worldPos, perpRefOf(rayDir, normal));
              prevPos = e.pos; prevProjection = e.projection; prevTwoTol = 0.0f; prevDecided = decidedExact(prevProjection);
            }
            // a value at most kTiny in magnitude can push the fp32 product into underflow: then both factors must be exact
            if (curDecided == 0.0f && prevTwoTol != 0.0f) {
              if (COUNT) gathers += 65536u;
              ExactTap e = exactTap(P, pu, pv, worldPos, perpRefOf(rayDir, normal));
              prevPos = e.pos; prevProjection = e.projection; prevTwoTol = 0.0f; prevDecided = decidedExact(prevProjection);
            }
            if (prevDecided == 0.0f && curTwoTol != 0.0f) {
              if (COUNT) gathers += 65536u;
              ExactTap e = exactTap(P, cu, cv, worldPos, perpRefOf(rayDir, normal));
              curPos = e.pos; curProjection = e.projection; curTwoTol = 0.0f; curDecided = decidedExact(curProjection);
            }
            flip = __fmul_rn(curProjection, prevProjection) < 0.0f;
          }
          if (flip) {
            // worldStep = length(currentPos - prevPos) <= 2.0. Both taps in model coordinates (the usual case): the length of
            // their difference through the Gram matrix, no position formed. One of them already exact: the other's model position.
            const float pu = marchCoord(u0, uvEnd.x, i - 1), pv = marchCoord(v0, uvEnd.y, i - 1);
            float worldStep;
            if (curTwoTol != 0.0f && prevTwoTol != 0.0f) worldStep = modelLength(P, curPos - prevPos);
            else worldStep = length3((curTwoTol != 0.0f ? modelToWorld(P, curPos) : curPos) - (prevTwoTol != 0.0f ? modelToWorld(P, prevPos) : prevPos));
            // each threshold covers twice its tap's position error; the last term is the Gram form's own rounding
            const float tol = (curTwoTol + prevTwoTol) + 1e-5f * worldStep;
            bool near;
            if (worldStep + tol <= 2.0f) near = true;
            else if (worldStep - tol > 2.0f) near = false;
            else {
              if (COUNT) gathers += 65536u;
              if (curTwoTol != 0.0f) {
                ExactTap e = exactTap(P, cu, cv, worldPos, perpRefOf(rayDir, normal));
                curPos = e.pos; curProjection = e.projection; curTwoTol = 0.0f; curDecided = decidedExact(curProjection);
              }
              if (prevTwoTol != 0.0f) prevPos = exactTap(P, pu, pv, worldPos, perpRefOf(rayDir, normal)).pos;
              near = length3(curPos - prevPos) <= 2.0f;
            }
            if (near && facesRay(P, cu, cv, rayDir)) {
              ao += 1;
              break;
            }
          }
        }
      }
      prevPos = curPos; prevProjection = curProjection; prevTwoTol = curTwoTol; prevDecided = curDecided;
    }
  }
  return ao;
}

#ifndef ALTHEA_SSAO_MIN_BLOCKS
#define ALTHEA_SSAO_MIN_BLOCKS 4
#endif
#ifndef ALTHEA_SSAO_TILE_W
#define ALTHEA_SSAO_TILE_W 16
#endif
constexpr int kSsaoTileW = ALTHEA_SSAO_TILE_W, kSsaoTileH = 256 / ALTHEA_SSAO_TILE_W;
// One CTA per 16 x 16 tile in raster order (concurrently running tiles share L2 lines), or, after
// ssao_cull_kernel, four consecutive entries of the list of tiles it handed over (mostly-undecidable neighbourhoods) per CTA.
template <bool COUNT, int KIND> __global__ void __launch_bounds__(256, ALTHEA_SSAO_MIN_BLOCKS) ssao_kernel(const __grid_constant__ FrameParams P) {
  const int tilesX = (P.W + kSsaoTileW - 1) / kSsaoTileW;
  const unsigned listed = P.ssaoTileList ? __ldg(P.ssaoTileList) : 0u;
  unsigned gathers = 0u;
  for (unsigned k = P.ssaoTileList ? blockIdx.x * 4u : blockIdx.x, e = P.ssaoTileList ? min(k + 4u, listed) : k + 1u; k < e; ++k) {
    const int tile = P.ssaoTileList ? (int)__ldg(P.ssaoTileList + 1 + k) : (int)k;
    const int x = (tile % tilesX) * kSsaoTileW + (threadIdx.x % kSsaoTileW);
    const int y = P.y0 + (tile / tilesX) * kSsaoTileH + (threadIdx.x / kSsaoTileW);
    if (x >= P.W || y >= P.y1) continue;
    V4 position = FmtRGBA32F::load(P.position, x, y);
    uint8_t count = 255;
    if (position.w != 0.0f) {
      const float u = ((float)x + 0.5f) / (float)P.W, v = ((float)y + 0.5f) / (float)P.H;
      V3 normal = normalize3(xyz(FmtRGBA16F::load(P.normal, x, y)));
      count = (uint8_t)(KIND ? ssaoCountRayProxy<COUNT>(P, x, y, u, v, xyz(position), normal, gathers) : ssaoCountFiltered<COUNT>(P, x, y, u, v, xyz(position), normal, gathers));
    }
    rowPtrW<uint8_t>(P.ao, y)[x] = count;
  }
  if (COUNT) { // two device counters: records gathered, taps re-evaluated from the fp32 texels
    atomicAdd(P.gatherCounter, (unsigned long long)(gathers & 0xffffu));
    atomicAdd(P.gatherCounter + 1, (unsigned long long)(gathers >> 16));
  }
}

// ---- SSAO, round 2: coarse sign test over per-block plane records, exact steps compacted across the warp ------------------
// A march step can only score when dot(p - pos, perpRef) changes sign between two consecutive taps (SSAO.glsl:68). A position
// G-buffer written by a perspective camera holds p(x, y) = cam + D(x, y) t with D affine in the texel coordinates and t the eye
// depth, so a texel's projection is  c0 + t g(x, y) = t c0 (w - L(x, y)),  w = 1 / t,  c0 = dot(cam - pos, perpRef),
// g = dot(D, perpRef),  L = -g / c0: its SIGN is that of w - L up to the constant sign of c0, and L (the ray's plane in
// reciprocal depth) is affine on the screen. A planar surface has w affine on the screen too, so ONE 16-byte record per block of
// texels, {alpha, beta, gamma, r} with |w_k - (alpha + beta x_k + gamma y_k)| <= r for every texel the block covers, decides the
// sign at every texel of a tap's footprint (hence of the bilinear tap, a convex combination; and of its fp32 evaluation, by
// the margins below) whenever |w_plane - L| at the tap exceeds r plus the slack of the ray. The records of the tile's
// neighbourhood (32 x 32 blocks of 8, 16 or 32 texels, picked per tile from its nearest depth) are staged in shared memory by
// the TMA engine, so a tap costs one LDS.128 instead of a divergent 32-byte gather from L2.
//   phase 1 (per ray, all lanes in lockstep, no break): the 11 taps are classified (+, -, undecided); a step whose two taps are
//     decided and equal cannot flip and is dropped; every other step is pushed on the warp's queue as (lane, ray, step);
//   phase 2 (whenever 32 items wait): each lane takes one item, rebuilds that ray (the hash RNG is random access) and
//     evaluates the step with the filtered exact predicates of ssaoCountFiltered on the 32-byte position records.
// computeSSAO breaks at a ray's first scoring step and counts rays, so the count is the number of rays with ANY scoring step:
// steps are independent and can be evaluated in any order. Counts are those of ssao_exact_kernel, bit for bit (tests/).
// Tiles whose neighbourhood is mostly undecidable (random depth, foliage) are flagged and left to ssao_kernel.
//
// Margins of the coarse decision (DESIGN.md 4.1 spells out the derivation). Texel k's real-arithmetic projection is
//   pi_k = c0 + g_k / w_k + eps_k,  |eps_k| <= sqrt(3) err_k + 2^-20 (|cam|_1 + |pos|_1 + |p_k|_1)
// (err_k: distance of the texel from its model point, bounded by kPlaneErr (|p_k|_1 + |cam|_1) in every decidable block; the
// rest: rounding of c0, g), and the restatement's fp32 tap is within 64 ulp (|pos|_1 + |p|_1) of the bilinear combination.
// With |p_k|_1 <= |cam|_1 + Dmax1 / w_k and w_k <= |L_k| + |w_k - L_k| the sign of the tap is that of w - L whenever
//   |w_k - L_k| (1 - kappa) > kappa |L_k| + kNu Dmax1 / |c0|,  kappa = kNu (2 |cam|_1 + |pos|_1) / |c0|,
// at all four texels; rays with kappa > 0.01 take the exact path for all their steps (1 / (1 - kappa) <= 1.0102 otherwise).
#ifndef ALTHEA_PLANE_ETA_MAX
#define ALTHEA_PLANE_ETA_MAX 0.05f
#endif
constexpr float kPlaneEtaMax = ALTHEA_PLANE_ETA_MAX; // largest relative change of the reciprocal depth over one texel a decidable record may have
// constants of the margin rule that depend on it (kappa <= 0.01): |a| ((1 - eta) - kappa) > (1 - eta) S + kappa |L| + (1 + eta) kNu Dmax1 / |c0|
constexpr float kEtaDen = (1.0f - kPlaneEtaMax) - 0.01f;
constexpr float kEtaRec = 1.0005f * (1.0f - kPlaneEtaMax) / kEtaDen;                                   // factor on the record's slack S
constexpr float kEtaRay = kEtaRec * (1.001f * kPlaneEtaMax + 1e-3f) / (1.0f - kPlaneEtaMax) * 1.001f;  // the ray's share of the footprint term
constexpr float kEtaKappa = 1.001f / kEtaDen, kEtaDmax = 1.001f * (1.0f + kPlaneEtaMax) / kEtaDen;
constexpr int kPlaneWin = 32;        // blocks per side of the staged window
#ifndef ALTHEA_CULL_REACH
#define ALTHEA_CULL_REACH 0.5f
#endif
constexpr float kCullReach = ALTHEA_CULL_REACH; // screen reach of a tile's rays, in focal lengths per unit of (depth - 0.5): picks the window's block size

// The records of all-cleared blocks (r = -1) can answer for their taps as a class of their own (the projection of the clear colour
// is the same for every tap of a ray). Measured at 4K: 16 % fewer taps left to the exact path, but three more instructions per
// plane lookup in a kernel bound by instruction issue: 4.47 ms against 4.22 ms. Off; such records simply cannot decide.
#ifndef ALTHEA_CULL_SKY_CLASS
#define ALTHEA_CULL_SKY_CLASS 0
#endif
// A frame whose coarsest records almost never decide (random depth, foliage everywhere) is marched tile by tile anyway: the finer
// levels are not built and ssao_cull_kernel hands every tile over without staging anything.
ADEV bool planesHopeless(const FrameParams& P) { return P.ssaoPlaneStats[1] * 16u < P.ssaoPlaneStats[0]; }

// G lanes per record (8 for the 8-texel blocks of level 0: four records per warp, their reductions share the shuffles; a warp for
// the larger blocks). LEVEL 2 goes first and counts how many of its records can decide; levels 0 and 1 are skipped when hopeless.
template <int LEVEL, int G> ADEV void ssaoPlaneRecords(const FrameParams& P, unsigned block) {
  constexpr bool COARSEST = LEVEL == 2;
  const int lane = threadIdx.x & (G - 1);
  const long long rec = ((long long)block * 256 + threadIdx.x) / G;
  constexpr int level = LEVEL;
  const long long n = (long long)P.ssaoPlaneRow[level] * (P.ssaoPlaneNy[level] + 2 * kSsaoPlanePad + 1);
  // records beyond the level's end keep their lanes in the warp's shuffles (G < 32): clamped to the last record, not written
  const bool live = rec < n;
  const long long r = live ? rec : n - 1;
  if (!COARSEST && planesHopeless(P)) return;
  constexpr int S = 8 << level;
  const int row = P.ssaoPlaneRow[level];
  const int bx = (int)(r % row) - kSsaoPlanePad, by = (int)(r / row) - kSsaoPlanePad;
  float4* out = const_cast<float4*>(P.ssaoPlanes[level]) + ((long long)by * row + bx);
  const float inf = __int_as_float(0x7f800000);
  const bool padding = bx < 0 || by < 0 || bx >= P.ssaoPlaneNx[level] || by >= P.ssaoPlaneNy[level];
  // texels the record answers for: the block and an apron: the march assigns a tap to a block by x / S rounded to 1 / 16
  // (S / 32 texels) and y / S to 2^-9 or finer, its tap coordinate is within 1e-3 texels of the restatement's, and the
  // footprint is the two texels from floor(x); inside the image (a footprint that clamps at the border repeats a covered texel)
  const int X0 = bx * S, Y0 = by * S;
  const int xlo = max(X0 - 2 - (S >> 5), 0), xhi = min(X0 + S + 1, P.W - 1), ylo = max(Y0 - 2, 0), yhi = min(Y0 + S + 1, P.H - 1);
  const int nx = padding ? 1 : xhi - xlo + 1, ny = padding ? 0 : yhi - ylo + 1; // padding records read nothing
  auto recip = [&](int x, int y) { return __ldg(P.ssaoRecip + ((size_t)y * P.W + x)); }; // NaN: texel off the camera model
  // plane through the centre texel with the secant slopes of the middle row / column (for a quadratic surface these are the
  // least-squares slopes); the offset is re-centred on the residual range below
  const int xc = (xlo + xhi) >> 1, yc = (ylo + yhi) >> 1;
  float beta = 0.0f, gamma = 0.0f, alpha = 0.0f;
  if (!padding) {
    beta = nx > 1 ? (recip(xhi, yc) - recip(xlo, yc)) / (float)(xhi - xlo) : 0.0f;
    gamma = ny > 1 ? (recip(xc, yhi) - recip(xc, ylo)) / (float)(yhi - ylo) : 0.0f;
    alpha = recip(xc, yc) - (beta * (float)xc + gamma * (float)yc);
  }
  float rlo = inf, rhi = -inf, wmax = 0.0f;
  bool ok = true, cleared = true;
  // texel k = lane, lane + G, ... of the nx x ny box, row by row: the (column, row) pair is stepped, not divided out
  const int qG = G / nx, rG = G - qG * nx;
  int kx = lane % nx, ky = lane / nx;
  for (int k = lane; k < nx * ny; k += G) {
    const int x = xlo + kx, y = ylo + ky;
    kx += rG; ky += qG;
    if (kx >= nx) { kx -= nx; ky += 1; }
    const float w = recip(x, y);
    const float res = w - fmaf(beta, (float)x, fmaf(gamma, (float)y, alpha));
    ok = ok && (res == res) && (w > 0.0f);
    cleared = cleared && (w == -1.0f);
    rlo = fminf(rlo, res);
    rhi = fmaxf(rhi, res);
    wmax = fmaxf(wmax, w);
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    rlo = fminf(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
    rhi = fmaxf(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
    wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    ok = __shfl_xor_sync(0xffffffffu, (int)ok, o) && ok;
    cleared = __shfl_xor_sync(0xffffffffu, (int)cleared, o) && cleared;
  }
  if (lane != 0 || !live) return;
  if (padding) {
    *out = make_float4(0.0f, 0.0f, 0.0f, inf);
    return;
  }
  if (cleared) { // every texel the record answers for is the clear colour: r = -1 marks the class
    *out = make_float4(0.0f, 0.0f, 0.0f, ALTHEA_CULL_SKY_CLASS ? -1.0f : inf);
    return; // (not counted in the statistics: empty sky says nothing about whether surfaces decide)
  }
  {
    const float mid = 0.5f * (rlo + rhi);
    const float a2 = alpha + mid;
    // The tap is the bilinear combination of its footprint's texels, whose projections are t_k c0 (w_k - L_k): the weights
    // lambda_k t_k differ from lambda_k t by at most eta = dw / (w - dw) relatively, dw = what w can change by over one texel
    // (slope + residual), and bilinear weights reproduce the affine part of w - L exactly AT THE TAP. So the sign is decided
    // when |w_plane - L|(1 - eta) exceeds (1 + eta) r_raw + (1.001 eta + 1e-3)(|beta| + |gamma| + |Lx| + |Ly|) plus the
    // margins: only an eta-th of the slopes enters, not a whole texel of them. Records with eta above
    // kPlaneEtaMax, and blocks on the image border (a clamped footprint repeats a texel: no affine reproduction), cannot decide.
    const float rraw = 0.5f * (rhi - rlo) + 4.0e-7f * (fabsf(mid) + fabsf(rlo) + fabsf(rhi)) + 2.4e-7f * wmax;
    const float grec = fabsf(beta) + fabsf(gamma);
    const float dw = rraw + 1.001f * grec;
    // smallest plane value over the block: the plane is affine, its minimum sits at a corner
    const float wmin = a2 + fminf(beta * (float)xlo, beta * (float)xhi) + fminf(gamma * (float)ylo, gamma * (float)yhi);
    const float eta = wmin - dw > 0.0f ? 1.0001f * dw / (wmin - dw) : inf;
    const bool border = X0 == 0 || Y0 == 0 || X0 + S >= P.W || Y0 + S >= P.H;
    float r = kEtaRec * (rraw * (1.0f + eta) + (1.001f * eta + 1e-3f) * grec) / (1.0f - eta);
    // the rounding of the plane's evaluation here and in the march
    r += 4.8e-7f * (fabsf(a2) + fabsf(alpha) + fabsf(beta) * (float)(xhi + 1) + fabsf(gamma) * (float)(yhi + 1));
    r *= 1.000001f;
    const bool fin = ok && !border && eta <= kPlaneEtaMax && isfinite(a2) && isfinite(beta) && isfinite(gamma) && isfinite(r);
    *out = fin ? make_float4(a2, beta, gamma, r) : make_float4(0.0f, 0.0f, 0.0f, inf);
    if (COARSEST) { // blocks on the image border never decide: not counted either way
      if (!border) atomicAdd(P.ssaoPlaneStats, 1u);
      if (fin) atomicAdd(P.ssaoPlaneStats + 1, 1u);
    }
  }
}

// level 2 alone (its statistics decide whether the finer levels are built), then levels 0 and 1 in one launch: blocks [0, blocks0) build
// level 0, the rest level 1
__global__ void __launch_bounds__(256) ssao_planes_coarsest_kernel(const __grid_constant__ FrameParams P) { ssaoPlaneRecords<2, 32>(P, blockIdx.x); }
__global__ void __launch_bounds__(256) ssao_planes_finer_kernel(const __grid_constant__ FrameParams P, unsigned blocks0) {
  if (blockIdx.x < blocks0) ssaoPlaneRecords<0, 8>(P, blockIdx.x);
  else ssaoPlaneRecords<1, 32>(P, blockIdx.x - blocks0);
}

// the ray of one pixel: SSAO.glsl:36-45 (the hash RNG's state after k draws is seed + k, so ray r starts at seed + 3 r)
struct SsaoRay { V3 rayDir, perpRef; V2 uvEnd; };
// tangent-space direction of the ray whose RNG state starts at (sx, sy): three draws, SSAO.glsl:37-38
ADEV V3 ssaoLocalDir(uint32_t sx, uint32_t sy) {
  HashRng rng;
  rng.sx = sx;
  rng.sy = sy;
  const float x0 = rng.next(), x1 = rng.next(), x2 = rng.next();
  return normalize3(mk3(2.0f * x0 - 1.0f, 2.0f * x1 - 1.0f, x2));
}
// The same for every seed a frame of this size can reach, once (FrameParams::ssaoDirs): the values the inline code computes, in
// this build's arithmetic, so a kernel reading them counts exactly what it would count hashing.
__global__ void __launch_bounds__(256) ssao_dirs_kernel(float4* dirs, int row, int rows) {
  const int a = blockIdx.x * 32 + (threadIdx.x & 31), b = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (a >= row || b >= rows) return;
  const V3 d = ssaoLocalDir((uint32_t)a, (uint32_t)b);
  dirs[(size_t)b * row + a] = make_float4(d.x, d.y, d.z, 0.0f);
}
ADEV SsaoRay ssaoRay(const FrameParams& P, const TangentFrame& tbn, V3 worldPos, V3 normal, int px, int py, int ray) {
  V3 local;
  if (P.ssaoDirs) { // coalesced: the lanes of a warp read two runs of 16 neighbouring entries
    const float4 t = __ldg(P.ssaoDirs + ((size_t)(py + 3 * ray) * P.ssaoDirRow + (px + 3 * ray)));
    local = mk3(t.x, t.y, t.z);
  } else {
    local = ssaoLocalDir((uint32_t)px + 3u * (uint32_t)ray, (uint32_t)py + 3u * (uint32_t)ray);
  }
  SsaoRay r;
  r.rayDir = frameApply(tbn, local);
  r.uvEnd = projectUv(P, worldPos + r.rayDir * 0.5f);
  r.perpRef = normalize3(cross3(cross3(r.rayDir, normal), r.rayDir));
  return r;
}
// taps before the ray leaves the screen (SSAO.glsl:50), see ssaoCountFiltered. The tap coordinates move monotonically away
// from the pixel's centre by far more than an ulp per tap when the end point is outside, so "inside" holds for a prefix of
// the taps and the first tap outside is found by bisection (tap 0 is the pixel's own centre, inside).
ADEV int ssaoTapCount(float u0, float v0, V2 uvEnd) {
  if (!outside01(uvEnd.x, uvEnd.y)) return 12;
  int lo = 0, hi = 12; // tap lo is inside, tap hi is outside (tap 12 = the end point)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int mid = (lo + hi) >> 1;
    if (mid > lo) {
      if (outside01(marchCoord(u0, uvEnd.x, mid), marchCoord(v0, uvEnd.y, mid))) hi = mid;
      else lo = mid;
    }
  }
  return hi;
}

// Sign class of one tap the plane records could not call: +1 / -1 = sign of the restatement's fp32 projection (|value| > kTiny),
// 0 = a value so small (or NaN) that only the fp32 product of the step's two exact projections says whether it flips.
template <bool COUNT> ADEV int ssaoTapClass(const FrameParams& P, float cu, float cv, V3 worldPos, V3 perpRef, unsigned& gathers) {
  const float posSlop2 = (2.002f * kRoundSlop) * (fabsf(worldPos.x) + fabsf(worldPos.y) + fabsf(worldPos.z));
  const float projBias = fmaf(worldPos.z, perpRef.z, fmaf(worldPos.y, perpRef.y, worldPos.x * perpRef.x));
  const ProxyAddr pa = proxyAddr(P, cu, cv);
  const ProxyTap t = proxyEval(loadQuad(pa.rec), pa.fx, pa.fy);
  if (COUNT) gathers += 1u;
  const float projection = fmaf(t.pos.z, perpRef.z, fmaf(t.pos.y, perpRef.y, fmaf(t.pos.x, perpRef.x, -projBias)));
  if (fabsf(projection) > t.twoTol + posSlop2) return projection < 0.0f ? -1 : 1; // the sign of the exact value, which is > kTiny
  if (COUNT) gathers += 65536u;
  const float e = exactTap(P, cu, cv, worldPos, perpRef).projection;
  return fabsf(e) > kTiny ? (e < 0.0f ? -1 : 1) : 0;
}
// A step whose two taps have opposite sign classes (or, `exact`, a class-0 tap): the rest of SSAO.glsl:68-74
template <bool COUNT> ADEV bool ssaoFlipScores(const FrameParams& P, float u0, float v0, V3 worldPos, V3 rayDir, V3 perpRef, float uvEndX, float uvEndY, int i, bool exact,
                                               unsigned& gathers) {
  const float pu = marchCoord(u0, uvEndX, i - 1), pv = marchCoord(v0, uvEndY, i - 1);
  const float cu = marchCoord(u0, uvEndX, i), cv = marchCoord(v0, uvEndY, i);
  bool near;
  if (exact) {
    if (COUNT) gathers += 2u * 65536u;
    const ExactTap ec = exactTap(P, cu, cv, worldPos, perpRef), ep = exactTap(P, pu, pv, worldPos, perpRef);
    if (!(__fmul_rn(ec.projection, ep.projection) < 0.0f)) return false;
    near = length3(ec.pos - ep.pos) <= 2.0f;
  } else {
    const float posSlop2 = (2.002f * kRoundSlop) * (fabsf(worldPos.x) + fabsf(worldPos.y) + fabsf(worldPos.z));
    const ProxyAddr pa = proxyAddr(P, pu, pv), ca = proxyAddr(P, cu, cv);
    const ProxyTap pt = proxyEval(loadQuad(pa.rec), pa.fx, pa.fy), ct = proxyEval(loadQuad(ca.rec), ca.fx, ca.fy);
    if (COUNT) gathers += 2u;
    // worldStep = length(currentPos - prevPos) <= 2.0; each threshold covers twice its tap's position tolerance
    const float worldStep = length3(ct.pos - pt.pos);
    const float tol = (ct.twoTol + posSlop2) + (pt.twoTol + posSlop2);
    if (worldStep + tol <= 2.0f) near = true;
    else if (worldStep - tol > 2.0f) near = false;
    else {
      if (COUNT) gathers += 2u * 65536u;
      near = length3(exactTap(P, cu, cv, worldPos, perpRef).pos - exactTap(P, pu, pv, worldPos, perpRef).pos) <= 2.0f;
    }
  }
  return near && facesRay(P, cu, cv, rayDir);
}

ADEV float invSf(int level) { return level == 0 ? 0.125f : level == 1 ? 0.0625f : 0.03125f; }
constexpr int kFlipRing = 64; // flip items waiting per warp: at most 31 left over + one per lane
#ifndef ALTHEA_CULL_MIN_BLOCKS
#define ALTHEA_CULL_MIN_BLOCKS 4
#endif
#ifndef ALTHEA_CULL_RAY_PRETEST
#define ALTHEA_CULL_RAY_PRETEST 1
#endif
template <bool COUNT> __global__ void __launch_bounds__(256, ALTHEA_CULL_MIN_BLOCKS) ssao_cull_kernel(const __grid_constant__ FrameParams P) {
  // Per-thread arrays are indexed [field][thread of the CTA]: a lane's own slot is one register (its thread index) plus a
  // constant, another lane's slot is the warp's first thread + that lane: no per-warp base pointers to keep or rebuild.
  __shared__ __align__(128) float4 win[kPlaneWin * kPlaneWin];
  __shared__ uint16_t tapQueue[8][32 * 11];      // (lane, tap) of the current ray's undecided taps
  __shared__ unsigned tapResult[256];            // per lane: bit i = tap i negative, bit 16 + i = tap i of class 0
  __shared__ float flipRing[6][8 * kFlipRing];   // tag, uvEnd, rayDir of the steps that change sign: field, then warp, then slot
  __shared__ float rayState[8][256];             // the current ray of every lane: uvEnd, perpRef, rayDir (read across lanes)
  __shared__ unsigned hitMask[256];
  __shared__ float redMin[8];
  __shared__ int badSum[8];
  __shared__ float redDev[8];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wb = tid & ~31;           // the warp's first thread
  const int fw = warp * kFlipRing;    // the warp's first flip slot
  const int tileX = blockIdx.x * 16, tileY = P.y0 + blockIdx.y * 16;
  const int x = tileX + (lane & 15), y = tileY + warp * 2 + (lane >> 4);
  const bool inside = x < P.W && y < P.y1;
  V4 position = mk4(0.0f, 0.0f, 0.0f, 0.0f);
  if (inside) position = FmtRGBA32F::load(P.position, x, y);
  const bool covered = inside && position.w != 0.0f;
  const V3 worldPos = xyz(position);
  // nearest covered depth of the tile -> block size of the window (a ray is 0.5 world units long)
  float tmin = covered ? dot3(worldPos - mk3(P.ssaoCam[0], P.ssaoCam[1], P.ssaoCam[2]), mk3(P.ssaoFwd[0], P.ssaoFwd[1], P.ssaoFwd[2])) : __int_as_float(0x7f800000);
  if (!(tmin > 0.0f)) tmin = 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
  if (lane == 0) redMin[warp] = tmin;
  if (threadIdx.x == 0) mbarInit(&bar, 1);
  __syncthreads();
  tmin = fminf(fminf(fminf(redMin[0], redMin[1]), fminf(redMin[2], redMin[3])), fminf(fminf(redMin[4], redMin[5]), fminf(redMin[6], redMin[7])));
  if (tmin == __int_as_float(0x7f800000)) { // nothing to shade in this tile
    if (inside) rowPtrW<uint8_t>(P.ao, y)[x] = 255;
    return;
  }
  if (planesHopeless(P)) { // no plane records worth staging in this frame
    if (threadIdx.x == 0) P.ssaoTileList[1u + atomicAdd(P.ssaoTileList, 1u)] = blockIdx.y * gridDim.x + blockIdx.x;
    return;
  }
  const float reach = kCullReach * P.ssaoFocalPx / fmaxf(tmin - 0.5f, 1e-3f);
  const int level = reach <= 15.0f * 8.0f - 4.0f ? 0 : reach <= 15.0f * 16.0f - 4.0f ? 1 : 2;
  const int S = 8 << level;
  // window: blocks [wbx, wbx + 32) x [wby, wby + 32), the tile's first block in column / row 15
  const int wbx = (tileX >> (3 + level)) - 15, wby = (tileY >> (3 + level)) - 15;
  if (warp == 0) {
    if (lane == 0) mbarExpectTx(&bar, kPlaneWin * kPlaneWin * 16);
    __syncwarp();
    const float4* src = P.ssaoPlanes[level] + ((long long)(wby + lane) * P.ssaoPlaneRow[level] + wbx);
    bulkCopyG2S(&win[lane * kPlaneWin], src, kPlaneWin * 16, &bar);
  }
  hitMask[tid] = 0u;
  mbarWait(&bar, 0);
  float4 planeM = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  float planeR = __int_as_float(0x7f800000), planeRb = 0.0f;
  { // a neighbourhood that mostly cannot decide is marched by ssao_kernel instead: undecidable records among the blocks the
    // tile's rays can reach
    const int rb = min(15, (int)(reach * invSf(level)) + 2);
    const int row = threadIdx.x >> 3, col0 = (threadIdx.x & 7) * 4;
    int bad = 0;
    if (row >= 15 - rb && row <= 16 + rb) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (col0 + k >= 15 - rb && col0 + k <= 16 + rb) bad += win[row * kPlaneWin + col0 + k].w == __int_as_float(0x7f800000) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if (lane == 0) badSum[warp] = bad;
    __syncthreads();
    int all = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) all += badSum[k];
    if (all * 8 > 7 * (2 * rb + 2) * (2 * rb + 2)) { // the exact evaluations run at full lane occupancy here: worth it up to ~7 / 8 undecidable
      if (threadIdx.x == 0) P.ssaoTileList[1u + atomicAdd(P.ssaoTileList, 1u)] = blockIdx.y * gridDim.x + blockIdx.x;
      return;
    }
    // ---- a neighbourhood that is ONE plane (ground, a wall): every record the tile's rays can reach decides, and all of them lie
    // within planeR of the tile's own record planeM. A tap in block j is decided by record j's rule when |plane_j - L| > r_j + rayConst;
    // with planeR >= r_j + |plane_j - planeM| over block j's texels, |planeM - L| > planeR + rayConst implies that, with the sign of
    // planeM - L. planeM - L is affine along a ray, so a ray whose first and last taps clear planeR + rayConst on the same side has
    // no step that changes sign: it scores nothing and needs no lookup at all (the pre-test of the ray loop below).
    planeR = __int_as_float(0x7f800000);
    if (ALTHEA_CULL_RAY_PRETEST && all == 0) {
      planeM = win[15 * kPlaneWin + 15];
      float dev = 0.0f;
      if (row >= 15 - rb && row <= 16 + rb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int col = col0 + k;
          if (col >= 15 - rb && col <= 16 + rb) {
            const float4 rj = win[row * kPlaneWin + col];
            const float da = rj.x - planeM.x, db = rj.y - planeM.y, dc = rj.z - planeM.z;
            // the texels record (row, col) answers for (ssaoPlaneRecords: the block and its apron), as tap coordinates
            const float x0 = (float)((wbx + col) * S - 3 - (S >> 5)), x1 = (float)((wbx + col) * S + S + 2);
            const float y0 = (float)((wby + row) * S - 3), y1 = (float)((wby + row) * S + S + 2);
            const float ex = fmaxf(fabsf(x0), fabsf(x1)), ey = fmaxf(fabsf(y0), fabsf(y1));
            float d = fmaxf(fmaxf(fabsf(fmaf(db, x0, fmaf(dc, y0, da))), fabsf(fmaf(db, x1, fmaf(dc, y0, da)))),
                            fmaxf(fabsf(fmaf(db, x0, fmaf(dc, y1, da))), fabsf(fmaf(db, x1, fmaf(dc, y1, da)))));
            // rounding of the differences and of planeM's evaluation in the pre-test
            d += 1e-6f * (fabsf(rj.x) + fabsf(planeM.x) + (fabsf(rj.y) + fabsf(planeM.y)) * ex + (fabsf(rj.z) + fabsf(planeM.z)) * ey);
            dev = fmaxf(dev, d + rj.w);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dev = fmaxf(dev, __shfl_xor_sync(0xffffffffu, dev, o));
      if (lane == 0) redDev[warp] = dev;
      __syncthreads();
      dev = fmaxf(fmaxf(fmaxf(redDev[0], redDev[1]), fmaxf(redDev[2], redDev[3])), fmaxf(fmaxf(redDev[4], redDev[5]), fmaxf(redDev[6], redDev[7])));
      planeR = dev * 1.000001f; // +inf or NaN: no pre-test
      planeRb = (float)rb;
    }
  }
  // ---- per pixel
  const float u0 = ((float)x + 0.5f) / (float)P.W, v0 = ((float)y + 0.5f) / (float)P.H;
  V3 normal = mk3(0.0f, 0.0f, 1.0f);
  if (covered) normal = normalize3(xyz(FmtRGBA16F::load(P.normal, x, y)));
  const TangentFrame tbn = localToWorld(normal);
  const float posL1 = fabsf(worldPos.x) + fabsf(worldPos.y) + fabsf(worldPos.z);
  const float kappaNum = kNu * (2.0f * P.ssaoCamL1 + posL1) + kTiny;
  const float posSlop = kRoundSlop * posL1 + kTiny; // 64 ulp of the terms of dot(pos, perpRef)
  (void)posSlop;
  // block coordinates of a tap inside the window come out of the mantissa of  x / S + magic: with 4 (9) fraction bits the
  // low bits of the x (y) word are (block << 4) ((block << 9)): the byte offset of the record's column (row) in the window
  const float invS = 1.0f / (float)S;
  const float offX = 786432.0f - (float)wbx, offY = 24576.0f - (float)wby; // 1.5 * 2^19, 1.5 * 2^14
  const char* winBytes = reinterpret_cast<const char*>(win);
  const float xs0 = (float)x, ys0 = (float)y;
  uint16_t* tq = tapQueue[warp];
  int fhead = 0, ftail = 0; // FIFO of flip items, warp-uniform
  unsigned gathers = 0u, lookups = 0u, tapItems = 0u;
#if defined(ALTHEA_CULL_PROBE_LDS) || defined(ALTHEA_CULL_PROBE_ALU)
  float probe = 0.0f;
#endif
  // evaluates min(32, waiting) flip items, oldest first, one per lane
  auto drainFlips = [&]() {
    const int count = min(ftail - fhead, 32);
    const bool live = lane < count;
    const int e = (fhead + (live ? lane : 0)) & (kFlipRing - 1);
    const unsigned tag = __float_as_uint(flipRing[0][fw + e]); // lane | step << 5 | ray << 9 | exact << 14
    const int src = (int)(tag & 31u);
    const V3 sp = mk3(__shfl_sync(0xffffffffu, worldPos.x, src), __shfl_sync(0xffffffffu, worldPos.y, src), __shfl_sync(0xffffffffu, worldPos.z, src));
    const float su = __shfl_sync(0xffffffffu, u0, src), sv = __shfl_sync(0xffffffffu, v0, src);
    const V3 sn = mk3(__shfl_sync(0xffffffffu, normal.x, src), __shfl_sync(0xffffffffu, normal.y, src), __shfl_sync(0xffffffffu, normal.z, src));
    if (live) {
      const V3 rd = mk3(flipRing[3][fw + e], flipRing[4][fw + e], flipRing[5][fw + e]);
      const bool exact = (tag >> 14) & 1u;
      // perpRef only enters the projections of a class-0 step (rare): rebuilt as ssaoRay builds it
      const V3 pr = exact ? perpRefOf(rd, sn) : rd;
      if (ssaoFlipScores<COUNT>(P, su, sv, sp, rd, pr, flipRing[1][fw + e], flipRing[2][fw + e], (int)((tag >> 5) & 15u), exact, gathers)) atomicOr(&hitMask[wb + src], 1u << ((tag >> 9) & 31u));
    }
    fhead += count;
    __syncwarp();
  };
  // ---- what the coarse test needs of a ray: L(i) = L0 + i dL along its taps (xs0 + i dxs, ys0 + i dys), the slack rayConst
  struct RayCoarse { float c0, dxs, dys, L0, dL, rayConst, lastI; bool answerable; };
  auto rayCoarse = [&](const SsaoRay& R, int n) {
    RayCoarse C;
    const V3 cam = mk3(P.ssaoCam[0], P.ssaoCam[1], P.ssaoCam[2]);
    C.c0 = dot3(cam - worldPos, R.perpRef);
    const float au = dot3(mk3(P.ssaoDx[0], P.ssaoDx[1], P.ssaoDx[2]), R.perpRef), av = dot3(mk3(P.ssaoDy[0], P.ssaoDy[1], P.ssaoDy[2]), R.perpRef);
    const float ac = dot3(mk3(P.ssaoDc[0], P.ssaoDc[1], P.ssaoDc[2]), R.perpRef);
    float invC0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(invC0) : "f"(C.c0));
    const float aInv = fabsf(invC0);
    const float Lx = -au * invC0, Ly = -av * invC0;
    const float xEnd = fmaf(R.uvEnd.x, P.Wf, -0.5f), yEnd = fmaf(R.uvEnd.y, P.Hf, -0.5f);
    C.dxs = (xEnd - xs0) * (1.0f / 12.0f); C.dys = (yEnd - ys0) * (1.0f / 12.0f);
    C.L0 = -(fmaf(au, xs0, fmaf(av, ys0, ac))) * invC0;
    C.dL = fmaf(Lx, C.dxs, Ly * C.dys);
    const float Labs = fmaxf(fabsf(C.L0), fabsf(fmaf(11.0f, C.dL, C.L0))) + (fabsf(Lx) + fabsf(Ly));
    const float kappa = kappaNum * aInv;
    // slack of the ray: its share of the footprint term at the largest eta a record may have (kEtaRay), the margin rule with
    // its (1 - eta) and (1 - kappa) factors folded into the constants, the rounding of L (16 ulp of the largest term of
    // dot(D, perpRef) / c0) and of w - L
    C.rayConst = kEtaRay * (fabsf(Lx) + fabsf(Ly)) + kEtaKappa * kappa * Labs + kEtaDmax * kNu * P.ssaoDmax1 * aInv + 9.6e-7f * P.ssaoDmag * aInv + 4.8e-7f * Labs;
    // rays the records cannot answer for take the exact path for all their taps: a last tap outside the window, a plane
    // through (nearly) the camera. (A footprint that clamps at the image border repeats a texel the block covers.)
    C.lastI = (float)(max(n, 2) - 1);
    const float xl = fmaf(C.lastI, C.dxs, xs0), yl = fmaf(C.lastI, C.dys, ys0);
    const float wx = xl * invS - (float)wbx, wy = yl * invS - (float)wby; // window coordinates of the last tap, in blocks
    C.answerable = wx >= 0.25f && wx <= (float)kPlaneWin - 0.25f && wy >= 0.25f && wy <= (float)kPlaneWin - 0.25f && kappa <= 0.01f;
    return C;
  };
  auto tapCount = [&](const SsaoRay& R, bool active) {
    // the bisection runs for the warp only when some lane's ray ends off the screen (a vote: tiles away from the border skip it)
    int n = active ? 12 : 0;
    if (__any_sync(0xffffffffu, active && outside01(R.uvEnd.x, R.uvEnd.y))) n = active ? ssaoTapCount(u0, v0, R.uvEnd) : 0;
    return n;
  };
  // Rays still to be looked at, a bit per ray. A warp without a covered pixel has nothing to count (no block-wide barrier below
  // this point). In a planar neighbourhood (planeR finite) a first sweep generates every ray and keeps only those the tile's one
  // plane cannot clear (grazing rays, rays leaving the region: a few per cent); the loop below then runs as many times as the
  // busiest lane has rays left instead of 24 times.
  unsigned todo = covered ? 0xffffffu : 0u;
  if (COUNT && threadIdx.x == 0) atomicAdd(P.gatherCounter + (planeR < __int_as_float(0x7f800000) ? 77 : 78), 1ull); // planar / other tiles
  if (planeR < __int_as_float(0x7f800000) && __any_sync(0xffffffffu, covered)) {
    unsigned fail = 0u;
    const float regLo = 15.0f - planeRb + 0.25f, regHi = 17.0f + planeRb - 0.25f; // the blocks planeR answers for, in window coordinates
    for (int ray = 0; ray < 24; ++ray) {
      const SsaoRay R = ssaoRay(P, tbn, worldPos, normal, x, y, ray);
      const int n = tapCount(R, covered);
      const RayCoarse C = rayCoarse(R, n);
      const float xl = fmaf(C.lastI, C.dxs, xs0), yl = fmaf(C.lastI, C.dys, ys0);
      const float wx = xl * invS - (float)wbx, wy = yl * invS - (float)wby;
      const float x1 = xs0 + C.dxs, y1 = ys0 + C.dys;
      const float d1 = fmaf(planeM.y, x1, fmaf(planeM.z, y1, planeM.x)) - (C.L0 + C.dL);
      const float dl = fmaf(planeM.y, xl, fmaf(planeM.z, yl, planeM.x)) - fmaf(C.lastI, C.dL, C.L0);
      const float thr = planeR + C.rayConst;
      const bool clear = C.answerable && wx >= regLo && wx <= regHi && wy >= regLo && wy <= regHi && fabsf(d1) > thr && fabsf(dl) > thr && (d1 > 0.0f) == (dl > 0.0f);
      if (covered && n >= 3 && !clear) fail |= 1u << ray; // (n < 3: no step to test)
      if (COUNT && covered && n >= 3) atomicAdd(P.gatherCounter + (clear ? 76 : 79), 1ull); // diagnostics: rays of planar tiles cleared / kept
    }
    todo = fail;
  }
  while (__any_sync(0xffffffffu, todo != 0u)) {
    const bool active = todo != 0u;
    const int ray = active ? __ffs(todo) - 1 : 0;
    todo &= todo - 1u;
    const SsaoRay R = ssaoRay(P, tbn, worldPos, normal, x, y, ray);
    const int n = tapCount(R, active);
    rayState[0][tid] = R.uvEnd.x; rayState[1][tid] = R.uvEnd.y;
    rayState[2][tid] = R.perpRef.x; rayState[3][tid] = R.perpRef.y; rayState[4][tid] = R.perpRef.z;
    rayState[5][tid] = R.rayDir.x; rayState[6][tid] = R.rayDir.y; rayState[7][tid] = R.rayDir.z;
    unsigned decMask = 0u, negMask = 0u;
    { // the coarse test
      const RayCoarse C = rayCoarse(R, n);
      const float c0 = C.c0, dxs = C.dxs, dys = C.dys, L0 = C.L0, dL = C.dL;
      const bool answerable = C.answerable;
      float rayConst = C.rayConst;
      if (!answerable) rayConst = __int_as_float(0x7f800000);
      // a footprint of cleared texels (empty pixels) interpolates to exactly (0, 0, 0): its projection is -dot(pos, perpRef)
      // whatever the tap; known when it clears the rounding of the dot product. Stored with the sign the c0 flip below undoes.
#if ALTHEA_CULL_SKY_CLASS
      const float skyProj = dot3(worldPos, R.perpRef);
      const float dSky = fabsf(skyProj) > posSlop ? (((skyProj > 0.0f) != (c0 < 0.0f)) ? -1.0f : 1.0f) : 0.0f;
#endif
#ifndef ALTHEA_CULL_REUSE_RECORD
#define ALTHEA_CULL_REUSE_RECORD 1
#endif
      // The kernel is bound by the L1 data pipe as much as by instruction issue (one more LDS.128 per tap: +1.08 ms, sixteen more
      // FFMA per tap: +0.75 ms): 32 lanes reading 32 unrelated 16-byte records cost ~7 wavefronts. A tap that falls in the block of
      // the previous tap of its ray keeps that record in registers: the load is predicated off for those lanes (fewer lanes, fewer
      // bank conflicts).
      float4 rec = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0x7f800000));
      unsigned offPrev = 0xffffffffu;
#pragma unroll
      for (int i = 1; i < 12; ++i) {
        const float fi = (float)i;
        // unanswerable rays may have wild coordinates: the masked offset stays inside the window, the value is never used
        const float tx = fmaf(fi, dxs, xs0), ty = fmaf(fi, dys, ys0);
        const float vx = fmaf(tx, invS, offX), vy = fmaf(ty, invS, offY);
        const unsigned off = (__float_as_uint(vx) & 0x1f0u) | (__float_as_uint(vy) & 0x3e00u);
#if ALTHEA_CULL_REUSE_RECORD
        if (off != offPrev) rec = *reinterpret_cast<const float4*>(winBytes + off);
        offPrev = off;
#else
        rec = *reinterpret_cast<const float4*>(winBytes + off);
#endif
#if ALTHEA_CULL_SKY_CLASS
        const bool clearedRec = rec.w < 0.0f;
        const float d = clearedRec ? dSky : fmaf(rec.y, tx, fmaf(rec.z, ty, rec.x)) - fmaf(fi, dL, L0);
        if (fabsf(d) > (clearedRec ? 0.0f : rec.w + rayConst)) decMask |= 1u << i;
        if (d < 0.0f) negMask |= 1u << i;
#else
        // both masks are shifted in from the sign bits (one funnel shift each): tap i ends up at bit 11 - i. `decided` is
        // |d| > rec.w + rayConst, i.e. (rec.w + rayConst) - |d| negative; +inf thresholds give +inf, never negative
        const float d = fmaf(rec.y, tx, fmaf(rec.z, ty, rec.x)) - fmaf(fi, dL, L0);
#if defined(ALTHEA_CULL_PROBE_LDS)   // tuning probe: a second, equally conflicting record read per tap (what does the L1 data pipe cost?)
        { const float4 rec2 = *reinterpret_cast<const float4*>(winBytes + (off ^ 0x2010u)); probe += rec2.x + rec2.w; }
#elif defined(ALTHEA_CULL_PROBE_ALU) // tuning probe: sixteen more dependent FFMAs per tap (what does an issue slot cost?)
#pragma unroll
        for (int q = 0; q < 16; ++q) probe = fmaf(probe, d, tx);
#endif
        decMask = __funnelshift_l(__float_as_uint((rec.w + rayConst) - fabsf(d)), decMask, 1);
        negMask = __funnelshift_l(__float_as_uint(d), negMask, 1);
#endif
      }
#if !ALTHEA_CULL_SKY_CLASS
      decMask = __brev(decMask) >> 20; // bit 11 - i -> bit i
      negMask = __brev(negMask) >> 20;
      // a NaN difference (only possible on a ray the records cannot answer for: c0 = 0, wild coordinates) has no sign to trust
      if (!answerable) decMask = 0u;
#endif
      if (c0 < 0.0f) negMask = ~negMask; // the projection is t c0 (w - L): its sign, not that of w - L
    }
    // taps 1 .. n - 1 take part in steps 2 .. n - 1 (none when n < 3)
    const unsigned tapMask = n >= 3 ? (1u << n) - 2u : 0u;
    unsigned undecided = tapMask & ~decMask;
    negMask &= decMask;
    if (COUNT) {
      lookups += __popc(tapMask); tapItems += __popc(undecided);
      // diagnostics: per window level and tap index, the taps looked up / left undecided, and the rays no record answers for
      for (int i = 1; i < 12; ++i) {
        if ((tapMask >> i) & 1u) atomicAdd(P.gatherCounter + 4 + (level * 12 + i) * 2, 1ull);
        if ((undecided >> i) & 1u) atomicAdd(P.gatherCounter + 4 + (level * 12 + i) * 2 + 1, 1ull);
      }
      if (tapMask && decMask == 0u) atomicAdd(P.gatherCounter + 4 + (level * 12) * 2, 1ull); // rays with no decided tap at all
    }
    // ---- the undecided taps of this ray, compacted over the warp and classified from the position records
    tapResult[tid] = 0u;
    const int cnt = __popc(undecided);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    for (int pos = incl - cnt; undecided; ++pos) {
      const int i = __ffs(undecided) - 1;
      undecided &= undecided - 1u;
      tq[pos] = (uint16_t)(lane | (i << 5));
    }
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
      const bool live = base + lane < total;
      const unsigned item = live ? (unsigned)tq[base + lane] : (unsigned)lane;
      const int src = (int)(item & 31u), i = (int)(item >> 5);
      const V3 sp = mk3(__shfl_sync(0xffffffffu, worldPos.x, src), __shfl_sync(0xffffffffu, worldPos.y, src), __shfl_sync(0xffffffffu, worldPos.z, src));
      const float su = __shfl_sync(0xffffffffu, u0, src), sv = __shfl_sync(0xffffffffu, v0, src);
      if (live) {
        const int cls = ssaoTapClass<COUNT>(P, marchCoord(su, rayState[0][wb + src], i), marchCoord(sv, rayState[1][wb + src], i), sp,
                                            mk3(rayState[2][wb + src], rayState[3][wb + src], rayState[4][wb + src]), gathers);
        if (cls <= 0) atomicOr(&tapResult[wb + src], cls < 0 ? 1u << i : 0x10000u << i);
      }
    }
    __syncwarp();
    // ---- steps whose taps differ in sign (or have a class-0 tap) go on the flip ring with what their evaluation needs
    const unsigned res = tapResult[tid];
    negMask |= res & 0xffffu;
    const unsigned zero = res >> 16;
    const unsigned stepMask = tapMask & ~2u; // steps 2 .. n - 1
    const unsigned exactSteps = (zero | (zero << 1)) & stepMask;
    unsigned flips = (((negMask ^ (negMask << 1)) & stepMask) & ~exactSteps) | (exactSteps << 16);
    while (__any_sync(0xffffffffu, flips != 0u)) {
      const bool has = flips != 0u;
      const unsigned b = __ballot_sync(0xffffffffu, has);
      if (has) {
        const int bit = __ffs(flips) - 1;
        flips &= flips - 1u;
        const int e = (ftail + __popc(b & ((1u << lane) - 1u))) & (kFlipRing - 1);
        flipRing[0][fw + e] = __uint_as_float((unsigned)lane | ((unsigned)(bit & 15) << 5) | ((unsigned)ray << 9) | ((unsigned)(bit >> 4) << 14));
        flipRing[1][fw + e] = rayState[0][tid]; flipRing[2][fw + e] = rayState[1][tid];
        flipRing[3][fw + e] = rayState[5][tid]; flipRing[4][fw + e] = rayState[6][tid]; flipRing[5][fw + e] = rayState[7][tid];
      }
      ftail += __popc(b);
      __syncwarp();
      if (ftail - fhead >= 32) drainFlips();
    }
  }
  while (ftail > fhead) drainFlips();
  __syncwarp();
  if (inside) rowPtrW<uint8_t>(P.ao, y)[x] = covered ? (uint8_t)__popc(hitMask[tid]) : (uint8_t)255;
#if defined(ALTHEA_CULL_PROBE_LDS) || defined(ALTHEA_CULL_PROBE_ALU)
  if (probe == 123.456f && inside) rowPtrW<uint8_t>(P.ao, y)[x] = 7; // keeps the probe's work alive
#endif
  if (COUNT) {
    atomicAdd(P.gatherCounter, (unsigned long long)(gathers & 0xffffu));
    atomicAdd(P.gatherCounter + 1, (unsigned long long)(gathers >> 16));
    atomicAdd(P.gatherCounter + 2, (unsigned long long)lookups);
    atomicAdd(P.gatherCounter + 3, (unsigned long long)tapItems);
  }
}

// ---- mode D: positions from depth ------------------------------------------------------------------------------------
// Today's GBufferResources has no position attachment (Src/DeferredRendering.cpp:42-99); the lighting pass and SSAO then work on
// reconstructPosition(uv, depth) (Misc/ReconstructPosition.glsl:4-22), emptiness coming from normal.a == 0 as in SSR.frag:136-141.
// One pass writes them to engine scratch in the legacy attachment's layout, so everything downstream is the mode-P path.
__global__ void __launch_bounds__(256) reconstruct_position_kernel(const __grid_constant__ FrameParams P) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15);
  const int y = blockIdx.y * 16 + (threadIdx.x >> 4); // whole frame: SSAO taps reach outside a scissor band
  if (x >= P.W || y >= P.H) return;
  rowPtrW<float4>(P.position, y)[x] = reconstructedTexel(P, x, y);
}

// ---- deferred shading -----------------------------------------------------------------------------------------------
ADEV V3 tonemap(V3 c, float exposure) {
  return mk3(1.0f - expf(-c.x * exposure), 1.0f - expf(-c.y * exposure), 1.0f - expf(-c.z * exposure));
}

__global__ void __launch_bounds__(256) deferred_shade_kernel(const __grid_constant__ FrameParams P) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15);
  const int y = P.y0 + blockIdx.y * 16 + (threadIdx.x >> 4);
  if (x >= P.W || y >= P.y1) return;
  const float u = ((float)x + 0.5f) / (float)P.W, v = ((float)y + 0.5f) / (float)P.H;
  const V3 direction = viewDirection(P, u, v);
  V4 position = FmtRGBA32F::load(P.position, x, y);
  V3 outc;
  if (position.w == 0.0f) { // DeferredPass.frag:45-53
    outc = sampleEnvMapLod0(P, direction);
  } else {
    V3 normal = normalize3(xyz(FmtRGBA16F::load(P.normal, x, y)));
    V3 baseColor = xyz(FmtRGBA8::load(P.albedo, x, y));
    V3 mro = xyz(FmtRGBA8::load(P.mro, x, y));
    V3 vdir = normalize3(direction);
    V3 reflectedDirection = reflect3(vdir, normal);
    V4 reflectedColor = trilinear<FmtRGBA16F, AddrClamp, kFastTex>(P.refl, u, v, 4.0f * mro.y);
    V3 envReflected = sampleEnvMapRough(P, reflectedDirection, mro.y);
    V3 rc;
    if (reflectedColor.w < 0.01f) rc = envReflected;
    else rc = mix3(envReflected, xyz(reflectedColor) / reflectedColor.w, reflectedColor.w);
    V3 irradianceColor = sampleIrrMap(P, normal);
    if (!(P.flags & ALTHEA_SHADE_NO_SSAO)) {
      uint8_t cnt = __ldg(rowPtr<uint8_t>(P.ao, y) + x);
      mro.z = 1.0f - (float)cnt / 24.0f;
    }
    outc = pbrMaterial(P, xyz(position), vdir, normal, baseColor, rc, irradianceColor, mro.x, mro.y, mro.z);
  }
  if (!(P.flags & ALTHEA_SHADE_SKIP_TONEMAP)) outc = tonemap(outc, P.g.exposure);
  if (P.outIsF32) rowPtrW<float4>(P.out, y)[x] = make_float4(outc.x, outc.y, outc.z, 1.0f);
  else rowPtrW<uint2>(P.out, y)[x] = packHalf4(mk4(outc.x, outc.y, outc.z, 1.0f));
}

// ---- launchers ------------------------------------------------------------------------------------------------------
static inline dim3 tileGrid(int w, int h) { return dim3((unsigned)((w + 15) / 16), (unsigned)((h + 15) / 16)); }

void launch_ssr_capture(const FrameParams& P, cudaStream_t s) {
  if (P.ssrPlanes) { ssr_capture_skip_kernel<<<tileGrid(P.W, P.y1 - P.y0), 256, 0, s>>>(P); return; }
#if !defined(ALTHEA_PARITY) && ALTHEA_SSR_REFILL
  if (P.ssrHits) { // persistent warps, idle lanes refilled: pixels without a hit keep the clear
    cudaMemsetAsync(const_cast<char*>(static_cast<const char*>(P.refl.level[0].ptr)) + (size_t)P.y0 * P.refl.level[0].pitch, 0,
                    (size_t)(P.y1 - P.y0) * P.refl.level[0].pitch, s);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ssr_capture_refill_kernel<<<(unsigned)(sms * ALTHEA_SSR_MARCH_MIN_BLOCKS), 256, 0, s>>>(P);
    return;
  }
#endif
  ssr_capture_kernel<<<dim3((unsigned)((P.W + kSsrTileW - 1) / kSsrTileW), (unsigned)((P.y1 - P.y0 + kSsrTileH - 1) / kSsrTileH)), 256, 0, s>>>(P);
}
bool ssr_march_reads_depth_quads() {
#ifdef ALTHEA_PARITY
  return false;
#else
  return ALTHEA_SSR_DEPTH_QUADS != 0;
#endif
}
void launch_ssr_shade_hits(const FrameParams& P, cudaStream_t s) { ssr_shade_hits_kernel<<<148 * 8, 256, 0, s>>>(P); }
void launch_ssr_planes(const FrameParams& P, cudaStream_t s) { ssr_planes_kernel<<<(kSsrPlaneStride * kSsrPlaneRows + 7) / 8, 256, 0, s>>>(P); }
void launch_ssr_depth_pad(const FrameParams& P, cudaStream_t s) {
  ssr_depth_pad_kernel<<<dim3((unsigned)((P.W + 2 + 31) / 32), (unsigned)((P.H + 2 + 7) / 8)), 256, 0, s>>>(P);
}
// shared-memory bytes of the largest source window a TW x TH tile can need (mirrors the kernel's window computation)
static size_t stagedWindowBytes(const ConvolveParams& C, int TW, int TH) {
  const double rx = (double)C.src.w / C.dst.w, ry = (double)C.src.h / C.dst.h;
  const double offx = C.vertical ? 0.0 : 5.176470588235294 * C.src.w / C.dst.w, offy = C.vertical ? 5.176470588235294 * C.src.h / C.dst.w : 0.0;
  const size_t cols = (size_t)(TW * rx + 2.0 * offx) + 8, rows = (size_t)(TH * ry + 2.0 * offy) + 6;
  return cols * rows * 8;
}
void launch_glossy_convolve(const ConvolveParams& C, cudaStream_t s) {
#ifndef ALTHEA_PARITY
  { // exact 2 : 1 levels (every level of an even-sized chain): the fixed-filter kernels
    static const bool general = getenv("ALTHEA_CONVOLVE_GENERAL") != nullptr; // A/B switch: the seven general bilinear taps
    FirTable T;
    if (!general && C.src.w == 2 * C.dst.w && C.src.h == 2 * C.dst.h && firTable(C, &T)) {
      static const int R = getenv("ALTHEA_CONVOLVE_R") ? atoi(getenv("ALTHEA_CONVOLVE_R")) : 4; // tuning: target rows per thread
      const unsigned gx = (unsigned)((C.dst.w + 31) / 32);
      if (C.vertical && R == 2) glossy_fir_vertical_kernel<2><<<dim3(gx, (unsigned)((C.y1 - C.y0 + 15) / 16)), 256, 0, s>>>(C, T);
      else if (C.vertical && R == 8) glossy_fir_vertical_kernel<8><<<dim3(gx, (unsigned)((C.y1 - C.y0 + 63) / 64)), 256, 0, s>>>(C, T);
      else if (C.vertical) glossy_fir_vertical_kernel<4><<<dim3(gx, (unsigned)((C.y1 - C.y0 + 31) / 32)), 256, 0, s>>>(C, T);
      else glossy_fir_horizontal_kernel<<<dim3((unsigned)((C.dst.w + 31) / 32), (unsigned)((C.y1 - C.y0 + 7) / 8)), 256, 0, s>>>(C, T);
      return;
    }
  }
#endif
  constexpr int VW = 32, VH = 32, HW = 64, HH = 16; // vertical passes want tall tiles (halo in y), horizontal ones wide tiles
  const bool aligned = (C.src.w % 2 == 0) && ((uintptr_t)C.src.ptr % 16 == 0) && (C.src.pitch % 16 == 0);
  const size_t smem = C.vertical ? stagedWindowBytes(C, VW, VH) : stagedWindowBytes(C, HW, HH);
  const bool big = (long long)C.dst.w * (C.y1 - C.y0) >= 64 * 64; // tiny levels: launch latency dominates, nothing to stage
  if (aligned && big && smem <= 96 * 1024) {
    // the opt-in is per device (a process may hold one context per GPU) and idempotent: set it whenever a device is seen first
    static std::atomic<unsigned long long> optedIn{0ull};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !((optedIn.load(std::memory_order_relaxed) >> dev) & 1ull)) {
      cudaFuncSetAttribute(glossy_convolve_staged_kernel<VW, VH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(glossy_convolve_staged_kernel<HW, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (dev >= 0 && dev < 64) optedIn.fetch_or(1ull << dev, std::memory_order_relaxed);
    }
    if (C.vertical)
      glossy_convolve_staged_kernel<VW, VH><<<dim3((unsigned)((C.dst.w + VW - 1) / VW), (unsigned)((C.y1 - C.y0 + VH - 1) / VH)), 256, smem, s>>>(C);
    else
      glossy_convolve_staged_kernel<HW, HH><<<dim3((unsigned)((C.dst.w + HW - 1) / HW), (unsigned)((C.y1 - C.y0 + HH - 1) / HH)), 256, smem, s>>>(C);
    return;
  }
  glossy_convolve_kernel<<<tileGrid(C.dst.w, C.y1 - C.y0), 256, 0, s>>>(C);
}
void launch_ssao(const FrameParams& P, cudaStream_t s) {
  const int tiles = ((P.W + kSsaoTileW - 1) / kSsaoTileW) * ((P.y1 - P.y0 + kSsaoTileH - 1) / kSsaoTileH);
  const dim3 grid((unsigned)(P.ssaoTileList ? (tiles + 3) / 4 : tiles)); // listed tiles: four consecutive entries per CTA
  if (P.quadKind) {
    if (P.gatherCounter) ssao_kernel<true, 1><<<grid, 256, 0, s>>>(P);
    else ssao_kernel<false, 1><<<grid, 256, 0, s>>>(P);
  } else {
    if (P.gatherCounter) ssao_kernel<true, 0><<<grid, 256, 0, s>>>(P);
    else ssao_kernel<false, 0><<<grid, 256, 0, s>>>(P);
  }
}
static_assert(kSsaoTileW == 16, "ssao_cull_kernel and ssao_kernel share the 16 x 16 tile grid");
void launch_ssao_planes(const FrameParams& P, cudaStream_t s, bool coarsest) { // the coarsest level first, then the finer ones if worth it
  long long n[3];
  for (int l = 0; l < 3; ++l) n[l] = (long long)P.ssaoPlaneRow[l] * (P.ssaoPlaneNy[l] + 2 * kSsaoPlanePad + 1);
  if (coarsest) ssao_planes_coarsest_kernel<<<(unsigned)((n[2] + 7) / 8), 256, 0, s>>>(P);
  else {
    const unsigned blocks0 = (unsigned)((n[0] + 31) / 32), blocks1 = (unsigned)((n[1] + 7) / 8);
    ssao_planes_finer_kernel<<<blocks0 + blocks1, 256, 0, s>>>(P, blocks0);
  }
}
void launch_ssao_cull(const FrameParams& P, cudaStream_t s) {
  const dim3 grid((unsigned)((P.W + 15) / 16), (unsigned)((P.y1 - P.y0 + 15) / 16));
  if (P.gatherCounter) ssao_cull_kernel<true><<<grid, 256, 0, s>>>(P);
  else ssao_cull_kernel<false><<<grid, 256, 0, s>>>(P);
}
void launch_ssao_dirs(float4* dirs, int row, int rows, cudaStream_t s) {
  ssao_dirs_kernel<<<dim3((unsigned)((row + 31) / 32), (unsigned)((rows + 7) / 8)), 256, 0, s>>>(dirs, row, rows);
}
void launch_ssao_exact(const FrameParams& P, cudaStream_t s) { ssao_exact_kernel<<<tileGrid(P.W, P.y1 - P.y0), 256, 0, s>>>(P); }
void launch_ssao_quads(const FrameParams& P, cudaStream_t s) {
  const dim3 grid((unsigned)((P.W + 1 + 31) / 32), (unsigned)((P.H + 1 + 7) / 8));
  if (P.quadKind) ssao_rayquads_kernel<<<grid, 256, 0, s>>>(P);
  else ssao_quads_kernel<<<grid, 256, 0, s>>>(P);
}
void launch_reconstruct_position(const FrameParams& P, cudaStream_t s) { reconstruct_position_kernel<<<tileGrid(P.W, P.H), 256, 0, s>>>(P); }
void launch_deferred_shade(const FrameParams& P, cudaStream_t s) { deferred_shade_kernel<<<tileGrid(P.W, P.y1 - P.y0), 256, 0, s>>>(P); }

} // namespace ALTHEA_NS
