// IBL precompute kernels, sm_100a (FP32-FMA integrators, no tensor cores: nothing here is a dense contraction).
//
//   mip_downsample_kernel  <- Image::generateMipMaps, Src/Image.cpp:183-213 (LINEAR 2:1 blit)
//   ibl_irradiance_kernel  <- Shaders/IBL_Precompute/GenIrradianceMap.comp:106-155
//   ibl_prefilter_kernel   <- Shaders/IBL_Precompute/PreFilterEnvMap.comp:126-177
//   brdf_lut_kernel        <- the split-sum integral behind Content/PrecomputedMaps/brdf_lut.png (no generator in the reference)
//
// Work split: LANES threads cooperate on one output texel (LANES = 32: a warp strides over the sample index and reduces
// with shuffles; LANES = 1: one thread per texel, the reference's own summation order).
#define ALTHEA_NS althea_iblk
#include "device_math.cuh"
#include "launchers.h"

namespace althea_iblk {

// ---- fast scalar math for the integrators ------------------------------------------------------------------------------
// The precompute is a Monte-Carlo / Riemann integral checked against the oracle to 1e-3 (tests/test_ibl_parity.py), not a
// chain of threshold decisions, so IEEE division / sqrt / libm transcendentals (10-40 SASS instructions each) are replaced
// by MUFU forms and short polynomials with ~1e-7 error. This is where the instruction count per sample went from ~700 to
// ~200 (profiles/r2_ibl_full.md).
ADEV float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
ADEV float lg2_fast(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// atan on [0, 1]: Abramowitz & Stegun 4.4.49, |error| <= 2e-8 in exact arithmetic, 1.1e-7 in fp32
ADEV float atan01(float q) {
  const float s = q * q;
  float p = 0.0028662257f;
  p = fmaf(p, s, -0.0161657367f);
  p = fmaf(p, s, 0.0429096138f);
  p = fmaf(p, s, -0.0752896400f);
  p = fmaf(p, s, 0.1065626393f);
  p = fmaf(p, s, -0.1420889944f);
  p = fmaf(p, s, 0.1999355085f);
  p = fmaf(p, s, -0.3333314528f);
  return fmaf(p * s, q, q);
}
ADEV float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float r = atan01(mx > 0.0f ? mn * rcpf(mx) : 0.0f);
  if (ay > ax) r = 1.5707963267948966f - r;
  if (x < 0.0f) r = kPi - r;
  return y < 0.0f ? -r : r;
}

struct RepeatOnce { // REPEAT addressing for coordinates known to lie within one period of [0, n): no integer modulo
  static ADEV int wrap(int i, int n) { i += i < 0 ? n : 0; return i >= n ? i - n : i; }
};

// one bilinear RGB tap of one level (REPEAT), a + t (b - a) lerps
ADEV V3 bilinearRgbRepeat(const ImgView& im, float u, float v) {
  const float x = fmaf(u, (float)im.w, -0.5f), y = fmaf(v, (float)im.h, -0.5f);
  const float fx0 = floorf(x), fy0 = floorf(y);
  const float fx = x - fx0, fy = y - fy0;
  const int ix = (int)fx0, iy = (int)fy0;
  const int i0 = RepeatOnce::wrap(ix, im.w), i1 = RepeatOnce::wrap(ix + 1, im.w);
  const int j0 = RepeatOnce::wrap(iy, im.h), j1 = RepeatOnce::wrap(iy + 1, im.h);
  const float4* r0 = rowPtr<float4>(im, j0);
  const float4* r1 = rowPtr<float4>(im, j1);
  const float4 t00 = __ldg(r0 + i0), t10 = __ldg(r0 + i1), t01 = __ldg(r1 + i0), t11 = __ldg(r1 + i1);
  const float ax = fmaf(t10.x - t00.x, fx, t00.x), bx = fmaf(t11.x - t01.x, fx, t01.x);
  const float ay = fmaf(t10.y - t00.y, fx, t00.y), by = fmaf(t11.y - t01.y, fx, t01.y);
  const float az = fmaf(t10.z - t00.z, fx, t00.z), bz = fmaf(t11.z - t01.z, fx, t01.z);
  return mk3(fmaf(bx - ax, fy, ax), fmaf(by - ay, fy, ay), fmaf(bz - az, fy, az));
}

// GenIrradianceMap.comp:78-102 == PreFilterEnvMap.comp:98-122: direction -> equirect uv (REPEAT sampler, linear mips)
ADEV V2 equirectUvPrecompute(V3 dir) {
  float pitch, yaw = 0.0f;
  const float lenXz = fsqrt(dir.x * dir.x + dir.z * dir.z);
  if (lenXz > 0.001f) {
    yaw = atan2_fast(dir.z, dir.x);
    pitch = atan2_fast(dir.y, lenXz); // == atan(y / lenXz) for lenXz > 0
  } else pitch = dir.y > 0.0f ? 0.5f * kPi : -0.5f * kPi;
  V2 uv;
  uv.x = fmaf(yaw, 0.5f / kPi, 0.5f);
  uv.y = fmaf(pitch, 1.0f / kPi, 0.5f);
  return uv;
}
// explicit-LOD trilinear fetch, rule A5 (LOD clamped to the chain, LINEAR mip mode)
ADEV V3 sampleEnvMapPrecompute(const ChainView& env, V3 dir, float mip) {
  const V2 uv = equirectUvPrecompute(dir);
  float lod = mip == mip ? mip : 0.0f;
  lod = fminf(fmaxf(lod, 0.0f), (float)(env.mips - 1));
  const float l0f = floorf(lod);
  const int l0 = (int)l0f;
  const float f = lod - l0f;
  const V3 s0 = bilinearRgbRepeat(env.level[l0], uv.x, uv.y);
  if (f == 0.0f) return s0;
  const V3 s1 = bilinearRgbRepeat(env.level[min(l0 + 1, env.mips - 1)], uv.x, uv.y);
  return mk3(fmaf(s1.x - s0.x, f, s0.x), fmaf(s1.y - s0.y, f, s0.y), fmaf(s1.z - s0.z, f, s0.z));
}

ADEV V3 texelNormal(const IblParams& I, int face, int x, int y) {
  if (I.layout == ALTHEA_IBL_LAYOUT_EQUIRECT) { // texelPos / size -> (yaw, pitch)
    float u = (float)x / (float)I.out[0].w, v = (float)y / (float)I.out[0].h;
    float yaw = kPi * (2.0f * u - 1.0f);
    float pitch = kPi * (v - 0.5f);
    return mk3(cosf(pitch) * cosf(yaw), sinf(pitch), cosf(pitch) * sinf(yaw));
  }
  float sc = 2.0f * (((float)x + 0.5f) / (float)I.out[0].w) - 1.0f;
  float tc = 2.0f * (((float)y + 0.5f) / (float)I.out[0].h) - 1.0f;
  V3 d;
  switch (face) {
  case 0: d = mk3(1.0f, -tc, -sc); break;
  case 1: d = mk3(-1.0f, -tc, sc); break;
  case 2: d = mk3(sc, 1.0f, tc); break;
  case 3: d = mk3(sc, -1.0f, -tc); break;
  case 4: d = mk3(sc, -tc, 1.0f); break;
  default: d = mk3(-sc, -tc, -1.0f); break;
  }
  return d / sqrtf(dot3(d, d));
}

template <int LANES> ADEV V4 laneReduce(V4 a) {
  if (LANES > 1) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
      a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
      a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
      a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
    }
  }
  return a;
}

__global__ void __launch_bounds__(256) mip_downsample_kernel(const __grid_constant__ MipGenParams M) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15);
  const int y = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (x >= M.dst.w || y >= M.dst.h) return;
  V4 c = bilinear<FmtRGBA32F, AddrClamp>(M.src, ((float)x + 0.5f) / (float)M.dst.w, ((float)y + 0.5f) / (float)M.dst.h);
  rowPtrW<float4>(M.dst, y)[x] = make_float4(c.x, c.y, c.z, c.w);
}

constexpr int kMaxPhiTable = 512;

// LANES threads per texel: 1, 32 (a warp strides over theta), or 256 (a CTA: 64 theta lanes x 4 phi lanes, for outputs of a
// few thousand texels such as the 32^2 cube of BASELINE configs[1], which a warp per texel leaves 148 SMs mostly idle on).
// blockIdx.y = cube face.
template <int LANES> __global__ void __launch_bounds__(256) ibl_irradiance_kernel(const __grid_constant__ IblParams I) {
  __shared__ float2 phiTable[kMaxPhiTable]; // (cosPhi, sinPhi), identical for every texel
  __shared__ float4 partial[8];
  for (int j = threadIdx.x; j < I.phiSamples && j < kMaxPhiTable; j += blockDim.x) {
    float phi = (float)j * 0.5f * kPi / (float)I.phiSamples;
    phiTable[j] = make_float2(cosf(phi), sinf(phi));
  }
  __syncthreads();
  const int face = blockIdx.y;
  const ImgView& out = I.out[face];
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long texel = gtid / LANES;
  const int lane = (int)(gtid % LANES);
  const long long total = (long long)out.w * out.h;
  if (texel >= total) return; // whole LANES-group exits together
  const int x = (int)(texel % out.w), y = (int)(texel / out.w);
  const V3 nor = texelNormal(I, face, x, y);
  const TangentFrame tbn = localToWorld(nor);
  // theta lanes x phi lanes of the group (lanes over phi instead, i.e. neighbouring taps per warp, measured slower: 69.9 vs
  // 53.6 ms at 512 x 256: the integrator is bound by instruction issue, not by its loads, profiles/r2_ibl_full.md)
  constexpr int TL = LANES == 256 ? 64 : LANES, PL = LANES == 256 ? 4 : 1;
  const int tl = lane % TL, pl = lane / TL;
  V3 irradiance = mk3(0.0f, 0.0f, 0.0f);
  for (int i = tl; i < I.thetaSamples; i += TL) {
    float theta = (float)(i * 2) * kPi / (float)I.thetaSamples;
    float cosTheta = cosf(theta), sinTheta = sinf(theta);
    for (int j = pl; j < I.phiSamples; j += PL) {
      float cosPhi, sinPhi;
      if (j < kMaxPhiTable) { float2 cs = phiTable[j]; cosPhi = cs.x; sinPhi = cs.y; }
      else { float phi = (float)j * 0.5f * kPi / (float)I.phiSamples; cosPhi = cosf(phi); sinPhi = sinf(phi); }
      V3 sampleDir = frameApply(tbn, mk3(cosTheta * sinPhi, sinTheta * sinPhi, cosPhi));
      irradiance = irradiance + (sampleEnvMapPrecompute(I.env, sampleDir, I.mip) * cosPhi) * sinPhi;
    }
  }
  V4 r = laneReduce<(LANES > 32 ? 32 : LANES)>(mk4(irradiance.x, irradiance.y, irradiance.z, 0.0f));
  if (LANES == 256) { // the CTA is one texel: add the eight warps' sums
    if ((threadIdx.x & 31) == 0) partial[threadIdx.x >> 5] = make_float4(r.x, r.y, r.z, 0.0f);
    __syncthreads();
    if (threadIdx.x == 0) {
      r = mk4(0.0f, 0.0f, 0.0f, 0.0f);
      for (int k = 0; k < 8; ++k) r = r + mk4(partial[k].x, partial[k].y, partial[k].z, 0.0f);
    }
  }
  if (lane == 0) {
    V3 c = ((kPi * xyz(r)) / (float)I.thetaSamples) / (float)I.phiSamples;
    rowPtrW<float4>(out, y)[x] = make_float4(c.x, c.y, c.z, 1.0f);
  }
}

template <int LANES> __global__ void __launch_bounds__(256) ibl_prefilter_kernel(const __grid_constant__ IblParams I) {
  const int face = blockIdx.y;
  const ImgView& out = I.out[face];
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long texel = gtid / LANES;
  const int lane = (int)(gtid % LANES);
  const long long total = (long long)out.w * out.h;
  if (texel >= total) return;
  const int x = (int)(texel % out.w), y = (int)(texel / out.w);
  const V3 N = texelNormal(I, face, x, y);
  const V3 V = N;
  const TangentFrame tbn = localToWorld(N);
  const float a2 = I.roughness * I.roughness;
  const float saTexel = 4.0f * kPi / (6.0f * (float)I.env.level[0].w * (float)I.env.level[0].h);
  const float fN = (float)I.numSamples;
  const float lg2SaTexel = log2f(saTexel);
  const float rcpN = 1.0f / fN;
  V3 acc = mk3(0.0f, 0.0f, 0.0f);
  float totalWeight = 0.0f;
  // roughness 0: cosTheta = sqrt(x / x) = 1 and sinTheta = 0 exactly, so H == N and every sample fetches the same texels
  // with the same weight (PreFilterEnvMap.comp:139-161 runs all 10000 regardless): the quotient acc / totalWeight is that
  // one sample. A sample whose xi1 is exactly 1 is NaN and skipped as in the reference, so each lane stops after its first
  // VALID sample.
  const int sampleEnd = I.numSamples;
  for (int i = lane; i < sampleEnd; i += LANES) {
    float xi0, xi1;
    if (I.sequence == ALTHEA_IBL_SEQ_REFERENCE_HASH) {
      HashRng rng; // the reference's RNG state after k draws is seed + k, so sample i starts at seed + 2i
      rng.sx = (uint32_t)x + 2u * (uint32_t)i;
      rng.sy = (uint32_t)y + 2u * (uint32_t)i;
      xi0 = rng.next();
      xi1 = rng.next();
    } else {
      xi0 = (float)i * rcpN;
      xi1 = (float)__brev((uint32_t)i) * 2.3283064365386963e-10f;
    }
    // phi = 2 pi xi0 in [0, 2 pi]: evaluate sin/cos at phi - pi in [-pi, pi] (MUFU range) and flip the signs
    float sinPhi, cosPhi;
    __sincosf(fmaf(xi0, 2.0f * kPi, -kPi), &sinPhi, &cosPhi);
    sinPhi = -sinPhi;
    cosPhi = -cosPhi;
    // cos^2 = (1 - xi) / (1 + (a2 - 1) xi), hence sin^2 = a2 xi / (1 + (a2 - 1) xi): formed directly, because
    // sqrt(1 - cos^2) (PreFilterEnvMap.comp:88-89) would amplify the reciprocal's last-bit error near cos = 1
    const float rden = rcpf(fmaf(a2 - 1.0f, xi1, 1.0f));
    const float cosTheta = fsqrt((1.0f - xi1) * rden); // 0 * inf = NaN at roughness 0, xi1 == 1: sample skipped below, as the reference's 0/0
    const float sinTheta = fsqrt(a2 * xi1 * rden);
    const V3 H = frameApply(tbn, mk3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
    const float VdotH = dot3(V, H); // V == N (PreFilterEnvMap.comp:133-135), so NdotH == HdotV == max(VdotH, 0)
    const V3 Lraw = (2.0f * VdotH) * H - V;
    const V3 L = Lraw * rsqrt_fast(dot3(Lraw, Lraw));
    const float NdotL = fmaxf(dot3(N, L), 0.0f); // NaN -> 0: skipped
    const float NdotH = fmaxf(VdotH, 0.0f);
    const float denom = fmaf(NdotH * NdotH, a2 - 1.0f, 1.0f);
    const float D = a2 * rcpf(kPi * denom * denom);
    const float pdf = D * NdotH * rcpf(fmaf(4.0f, NdotH, 0.00001f));
    const float saSample = rcpf(fmaf(fN, pdf, 0.0001f));
    const float mipLevel = I.roughness == 0.0f ? 0.0f : 0.5f * (lg2_fast(saSample) - lg2SaTexel);
    if (NdotL > 0.0f) {
      const V3 e = sampleEnvMapPrecompute(I.env, L, mipLevel);
      acc = mk3(fmaf(e.x, NdotL, acc.x), fmaf(e.y, NdotL, acc.y), fmaf(e.z, NdotL, acc.z));
      totalWeight += NdotL;
      if (I.roughness == 0.0f) break;
    }
  }
  V4 r = laneReduce<LANES>(mk4(acc.x, acc.y, acc.z, totalWeight));
  if (lane == 0) rowPtrW<float4>(out, y)[x] = make_float4(r.x / r.w, r.y / r.w, r.z / r.w, 1.0f);
}

__global__ void __launch_bounds__(256) brdf_lut_kernel(const __grid_constant__ LutParams Lp) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15);
  const int y = blockIdx.y * 16 + (threadIdx.x >> 4);
  const int size = Lp.out.w;
  if (x >= size || y >= Lp.out.h) return;
  // row 0 holds roughness ~ 1: the orientation of the reference's asset
  const int ry = Lp.out.h - 1 - y;
  const float NdotV = ((float)x + 0.5f) / (float)size, roughness = ((float)ry + 0.5f) / (float)Lp.out.h;
  const V3 V = mk3(sqrtf(1.0f - NdotV * NdotV), 0.0f, NdotV);
  const float a = roughness * roughness;
  const float a2 = a * a;
  const float k = a / 2.0f;
  const float fN = (float)Lp.samples;
  float A = 0.0f, B = 0.0f;
  for (int i = 0; i < Lp.samples; ++i) {
    float xi0 = (float)i / fN, xi1 = (float)__brev((uint32_t)i) * 2.3283064365386963e-10f;
    float phi = 2.0f * kPi * xi0;
    float cosTheta = sqrtf((1.0f - xi1) / (1.0f + (a2 - 1.0f) * xi1));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    V3 H = mk3(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta);
    V3 L = (2.0f * dot3(V, H)) * H - V;
    float NdotL = fmaxf(L.z, 0.0f), NdotH = fmaxf(H.z, 0.0f), VdotH = fmaxf(dot3(V, H), 0.0f);
    if (NdotL > 0.0f) {
      float G = (NdotV / (NdotV * (1.0f - k) + k)) * (NdotL / (NdotL * (1.0f - k) + k));
      float GVis = (G * VdotH) / (NdotH * NdotV);
      float om = 1.0f - VdotH;
      float om2 = om * om;
      float Fc = om2 * om2 * om;
      A += (1.0f - Fc) * GVis;
      B += Fc * GVis;
    }
  }
  A /= fN;
  B /= fN;
  if (Lp.outIsF32) rowPtrW<float4>(Lp.out, y)[x] = make_float4(A, B, 0.0f, 1.0f);
  else {
    // UNORM8 store: round(clamp(v,0,1)*255)
    uint32_t r = (uint32_t)__float2int_rn(fminf(fmaxf(A, 0.0f), 1.0f) * 255.0f);
    uint32_t g = (uint32_t)__float2int_rn(fminf(fmaxf(B, 0.0f), 1.0f) * 255.0f);
    rowPtrW<uint32_t>(Lp.out, y)[x] = r | (g << 8) | (255u << 24);
  }
}

static inline dim3 tileGrid(int w, int h) { return dim3((unsigned)((w + 15) / 16), (unsigned)((h + 15) / 16)); }
static inline unsigned linearGrid(long long threads) { return (unsigned)((threads + 255) / 256); }
constexpr long long kThreadPerTexelMin = 148LL * 2048LL * 2LL; // below ~2 resident waves of texels, give each texel a warp

void launch_mip_downsample(const MipGenParams& M, cudaStream_t s) { mip_downsample_kernel<<<tileGrid(M.dst.w, M.dst.h), 256, 0, s>>>(M); }

void launch_ibl_irradiance(const IblParams& I, cudaStream_t s) {
  const long long texels = (long long)I.out[0].w * I.out[0].h, all = texels * I.faces;
  const unsigned faces = (unsigned)I.faces;
  if (all >= kThreadPerTexelMin) ibl_irradiance_kernel<1><<<dim3(linearGrid(texels), faces), 256, 0, s>>>(I);
  else if (all * 32 >= kThreadPerTexelMin) ibl_irradiance_kernel<32><<<dim3(linearGrid(texels * 32), faces), 256, 0, s>>>(I);
  else ibl_irradiance_kernel<256><<<dim3((unsigned)texels, faces), 256, 0, s>>>(I);
}

void launch_ibl_prefilter(const IblParams& I, cudaStream_t s) {
  const long long texels = (long long)I.out[0].w * I.out[0].h * I.faces;
  // Hash RNG: neighbouring texels draw unrelated directions, so there is no coherence to lose by giving every texel a
  // warp. Hammersley: all texels share the sequence, so one thread per texel keeps a warp's taps adjacent.
  if (I.roughness == 0.0f || (I.sequence == ALTHEA_IBL_SEQ_HAMMERSLEY && texels >= kThreadPerTexelMin)) ibl_prefilter_kernel<1><<<dim3(linearGrid(texels / I.faces), (unsigned)I.faces), 256, 0, s>>>(I);
  else ibl_prefilter_kernel<32><<<dim3(linearGrid(texels / I.faces * 32), (unsigned)I.faces), 256, 0, s>>>(I);
}

void launch_brdf_lut(const LutParams& L, cudaStream_t s) { brdf_lut_kernel<<<tileGrid(L.out.w, L.out.h), 256, 0, s>>>(L); }

} // namespace althea_iblk
