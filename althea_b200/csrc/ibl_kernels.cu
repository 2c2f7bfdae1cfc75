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

// GenIrradianceMap.comp:78-102 == PreFilterEnvMap.comp:98-122. Precompute sampler: REPEAT, linear mips.
ADEV V3 sampleEnvMapPrecompute(const ChainView& env, V3 dir, float mip) {
  float pitch = 0.0f, yaw = 0.0f;
  float lenXz = sqrtf(dir.x * dir.x + dir.z * dir.z);
  if (lenXz > 0.001f) {
    yaw = atan2f(dir.z, dir.x);
    pitch = atanf(dir.y / lenXz);
  } else if (dir.y > 0.0f) pitch = 0.5f * kPi;
  else pitch = -0.5f * kPi;
  float u = yaw / (2.0f * kPi) + 0.5f, v = pitch / kPi + 0.5f;
  return xyz(trilinear<FmtRGBA32F, AddrRepeat>(env, u, v, mip));
}

ADEV V3 texelNormal(const IblParams& I, int x, int y) {
  if (I.layout == ALTHEA_IBL_LAYOUT_EQUIRECT) { // texelPos / size -> (yaw, pitch)
    float u = (float)x / (float)I.out.w, v = (float)y / (float)I.out.h;
    float yaw = kPi * (2.0f * u - 1.0f);
    float pitch = kPi * (v - 0.5f);
    return mk3(cosf(pitch) * cosf(yaw), sinf(pitch), cosf(pitch) * sinf(yaw));
  }
  float sc = 2.0f * (((float)x + 0.5f) / (float)I.out.w) - 1.0f;
  float tc = 2.0f * (((float)y + 0.5f) / (float)I.out.h) - 1.0f;
  V3 d;
  switch (I.face) {
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

template <int LANES> __global__ void __launch_bounds__(256) ibl_irradiance_kernel(const __grid_constant__ IblParams I) {
  __shared__ float2 phiTable[kMaxPhiTable]; // (cosPhi, sinPhi), identical for every texel
  for (int j = threadIdx.x; j < I.phiSamples && j < kMaxPhiTable; j += blockDim.x) {
    float phi = (float)j * 0.5f * kPi / (float)I.phiSamples;
    phiTable[j] = make_float2(cosf(phi), sinf(phi));
  }
  __syncthreads();
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long texel = gtid / LANES;
  const int lane = (int)(gtid % LANES);
  const long long total = (long long)I.out.w * I.out.h;
  if (texel >= total) return; // whole LANES-group exits together
  const int x = (int)(texel % I.out.w), y = (int)(texel / I.out.w);
  const V3 nor = texelNormal(I, x, y);
  const TangentFrame tbn = localToWorld(nor);
  V3 irradiance = mk3(0.0f, 0.0f, 0.0f);
  for (int i = lane; i < I.thetaSamples; i += LANES) {
    float theta = (float)(i * 2) * kPi / (float)I.thetaSamples;
    float cosTheta = cosf(theta), sinTheta = sinf(theta);
    for (int j = 0; j < I.phiSamples; ++j) {
      float cosPhi, sinPhi;
      if (j < kMaxPhiTable) { float2 cs = phiTable[j]; cosPhi = cs.x; sinPhi = cs.y; }
      else { float phi = (float)j * 0.5f * kPi / (float)I.phiSamples; cosPhi = cosf(phi); sinPhi = sinf(phi); }
      V3 sampleDir = frameApply(tbn, mk3(cosTheta * sinPhi, sinTheta * sinPhi, cosPhi));
      irradiance = irradiance + (sampleEnvMapPrecompute(I.env, sampleDir, I.mip) * cosPhi) * sinPhi;
    }
  }
  V4 r = laneReduce<LANES>(mk4(irradiance.x, irradiance.y, irradiance.z, 0.0f));
  if (lane == 0) {
    V3 c = ((kPi * xyz(r)) / (float)I.thetaSamples) / (float)I.phiSamples;
    rowPtrW<float4>(I.out, y)[x] = make_float4(c.x, c.y, c.z, 1.0f);
  }
}

template <int LANES> __global__ void __launch_bounds__(256) ibl_prefilter_kernel(const __grid_constant__ IblParams I) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long texel = gtid / LANES;
  const int lane = (int)(gtid % LANES);
  const long long total = (long long)I.out.w * I.out.h;
  if (texel >= total) return;
  const int x = (int)(texel % I.out.w), y = (int)(texel / I.out.w);
  const V3 N = texelNormal(I, x, y);
  const V3 V = N;
  const TangentFrame tbn = localToWorld(N);
  const float a2 = I.roughness * I.roughness;
  const float saTexel = 4.0f * kPi / (6.0f * (float)I.env.level[0].w * (float)I.env.level[0].h);
  const float fN = (float)I.numSamples;
  V3 acc = mk3(0.0f, 0.0f, 0.0f);
  float totalWeight = 0.0f;
  // roughness 0: cosTheta = sqrt(x / x) = 1 and sinTheta = 0 exactly, so H == N and every sample fetches the same texels
  // with the same weight (PreFilterEnvMap.comp:139-161 runs all 10000 regardless). One pass of the lanes gives the same
  // quotient acc / totalWeight to within fp32 summation noise; a sample whose xi1 is exactly 1 is NaN and skipped as in
  // the reference.
  const int sampleEnd = (I.roughness == 0.0f) ? min(I.numSamples, LANES) : I.numSamples;
  for (int i = lane; i < sampleEnd; i += LANES) {
    float xi0, xi1;
    if (I.sequence == ALTHEA_IBL_SEQ_REFERENCE_HASH) {
      HashRng rng; // the reference's RNG state after k draws is seed + k, so sample i starts at seed + 2i
      rng.sx = (uint32_t)x + 2u * (uint32_t)i;
      rng.sy = (uint32_t)y + 2u * (uint32_t)i;
      xi0 = rng.next();
      xi1 = rng.next();
    } else {
      xi0 = (float)i / fN;
      xi1 = (float)__brev((uint32_t)i) * 2.3283064365386963e-10f;
    }
    float phi = 2.0f * kPi * xi0;
    float cosTheta = sqrtf((1.0f - xi1) / (1.0f + (a2 - 1.0f) * xi1));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    V3 H = frameApply(tbn, mk3(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta));
    V3 Lraw = (2.0f * dot3(V, H)) * H - V;
    V3 L = Lraw / sqrtf(dot3(Lraw, Lraw));
    float NdotL = fmaxf(dot3(N, L), 0.0f); // NaN (xi1 == 1 at roughness 0) -> 0: the sample is skipped
    float NdotH = fmaxf(dot3(N, H), 0.0f);
    float HdotV = fmaxf(dot3(H, V), 0.0f);
    float denom = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
    float D = a2 / (kPi * denom * denom);
    float pdf = D * NdotH / (4.0f * HdotV + 0.00001f);
    float saSample = 1.0f / (fN * pdf + 0.0001f);
    float mipLevel = I.roughness == 0.0f ? 0.0f : 0.5f * log2f(saSample / saTexel);
    if (NdotL > 0.0f) {
      acc = acc + sampleEnvMapPrecompute(I.env, L, mipLevel) * NdotL;
      totalWeight += NdotL;
    }
  }
  V4 r = laneReduce<LANES>(mk4(acc.x, acc.y, acc.z, totalWeight));
  if (lane == 0) rowPtrW<float4>(I.out, y)[x] = make_float4(r.x / r.w, r.y / r.w, r.z / r.w, 1.0f);
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
  long long texels = (long long)I.out.w * I.out.h;
  if (texels >= kThreadPerTexelMin) ibl_irradiance_kernel<1><<<linearGrid(texels), 256, 0, s>>>(I);
  else ibl_irradiance_kernel<32><<<linearGrid(texels * 32), 256, 0, s>>>(I);
}

void launch_ibl_prefilter(const IblParams& I, cudaStream_t s) {
  long long texels = (long long)I.out.w * I.out.h;
  // Hash RNG: neighbouring texels draw unrelated directions, so there is no coherence to lose by giving every texel a
  // warp. Hammersley: all texels share the sequence, so one thread per texel keeps a warp's taps adjacent.
  if (I.sequence == ALTHEA_IBL_SEQ_HAMMERSLEY && texels >= kThreadPerTexelMin) ibl_prefilter_kernel<1><<<linearGrid(texels), 256, 0, s>>>(I);
  else ibl_prefilter_kernel<32><<<linearGrid(texels * 32), 256, 0, s>>>(I);
}

void launch_brdf_lut(const LutParams& L, cudaStream_t s) { brdf_lut_kernel<<<tileGrid(L.out.w, L.out.h), 256, 0, s>>>(L); }

} // namespace althea_iblk
