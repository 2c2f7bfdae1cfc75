// Device-side vector math and the "Vulkan texture unit" rules for the sm_100a kernels.
//
// The per-frame kernels are compiled twice from the same source (see build.py):
//   * fast   : default nvcc FP contraction (FFMA), rsqrt-based normalisation        -> namespace althea_fast
//   * parity : -fmad=false -DALTHEA_PARITY, IEEE div/sqrt everywhere, so every + - * / sqrt is the same correctly
//              rounded operation, in the same order, as the GLSL restatement evaluates it   -> namespace althea_parity
// Texture filtering is done in FP32 in the kernels (not by the TEX unit, whose 8-bit fractional weights would move
// the SSAO/SSR threshold tests): unnormalised coordinate = u*size - 0.5, two taps per axis, lerp of lerps.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "params.h"

#ifndef ALTHEA_NS
#define ALTHEA_NS althea_fast
#endif

namespace ALTHEA_NS {

#define ADEV __device__ __forceinline__

constexpr float kPi = 3.14159265359f; // the literal every shader on the path uses (Shaders/Misc/Constants.glsl:4)

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

// Division and square root: IEEE (correctly rounded, ~10 SASS instructions with the slow-path check) in the parity build,
// one MUFU + one multiply in the fast build. Used for colour arithmetic and directions, never for texel addressing.
ADEV float rcpf(float x) {
#ifdef ALTHEA_PARITY
  return 1.0f / x;
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
ADEV float fdiv(float a, float b) {
#ifdef ALTHEA_PARITY
  return a / b;
#else
  return a * rcpf(b);
#endif
}
ADEV float fsqrt(float x) {
#ifdef ALTHEA_PARITY
  return sqrtf(x);
#else
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}

ADEV V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
ADEV V4 mk4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
ADEV V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
ADEV V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
ADEV V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
ADEV V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
ADEV V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
ADEV V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
ADEV V3 operator/(V3 a, float s) {
#ifdef ALTHEA_PARITY
  return mk3(a.x / s, a.y / s, a.z / s);
#else
  const float r = rcpf(s);
  return mk3(a.x * r, a.y * r, a.z * r);
#endif
}
ADEV V4 operator+(V4 a, V4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
ADEV V4 operator*(V4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }
ADEV float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
ADEV V3 cross3(V3 a, V3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
ADEV V3 xyz(V4 a) { return mk3(a.x, a.y, a.z); }
ADEV float length3(V3 a) { return sqrtf(dot3(a, a)); }
ADEV V3 normalize3(V3 a) {
#ifdef ALTHEA_PARITY
  return a / sqrtf(dot3(a, a));
#else
  return a * rsqrtf(dot3(a, a));
#endif
}
ADEV V3 reflect3(V3 i, V3 n) { return i - (2.0f * dot3(n, i)) * n; }
ADEV float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
ADEV V3 mix3(V3 a, V3 b, float t) { return mk3(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }
ADEV V4 mix4(V4 a, V4 b, float t) { return mk4(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t), mixf(a.w, b.w, t)); }
ADEV float max_glsl(float a, float b) { return a > b ? a : b; }

// column-major mat4 * vec4, summed left to right
ADEV V4 mul44(const float* m, V4 v) {
  return mk4(((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w, ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w,
             ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w, ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w);
}
ADEV V3 mul33(const float* m, V3 v) { // mat3(m) * v
  return mk3((m[0] * v.x + m[4] * v.y) + m[8] * v.z, (m[1] * v.x + m[5] * v.y) + m[9] * v.z, (m[2] * v.x + m[6] * v.y) + m[10] * v.z);
}

// ---- images (ImgView / ChainView come from params.h) ----------------------------------------------------------
template <typename T> ADEV const T* rowPtr(const ImgView& im, int y) {
  return reinterpret_cast<const T*>(static_cast<const char*>(im.ptr) + (size_t)y * im.pitch);
}
template <typename T> ADEV T* rowPtrW(const ImgView& im, int y) {
  return reinterpret_cast<T*>(const_cast<char*>(static_cast<const char*>(im.ptr)) + (size_t)y * im.pitch);
}

ADEV V4 unpackHalf4(uint2 raw) {
  __half2 lo = *reinterpret_cast<__half2*>(&raw.x), hi = *reinterpret_cast<__half2*>(&raw.y);
  float2 a = __half22float2(lo), b = __half22float2(hi);
  return mk4(a.x, a.y, b.x, b.y);
}
ADEV uint2 packHalf4(V4 v) { // round-to-nearest-even, overflow -> inf (rule A7)
  __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&lo);
  r.y = *reinterpret_cast<uint32_t*>(&hi);
  return r;
}
ADEV V4 unpackUnorm4(uint32_t raw) {
  return mk4((float)(raw & 0xffu) / 255.0f, (float)((raw >> 8) & 0xffu) / 255.0f, (float)((raw >> 16) & 0xffu) / 255.0f,
             (float)(raw >> 24) / 255.0f);
}

struct FmtRGBA32F { static ADEV V4 load(const ImgView& im, int x, int y) { float4 t = __ldg(rowPtr<float4>(im, y) + x); return mk4(t.x, t.y, t.z, t.w); } };
struct FmtRGBA16F { static ADEV V4 load(const ImgView& im, int x, int y) { return unpackHalf4(__ldg(rowPtr<uint2>(im, y) + x)); } };
struct FmtRGBA8   { static ADEV V4 load(const ImgView& im, int x, int y) { return unpackUnorm4(__ldg(rowPtr<uint32_t>(im, y) + x)); } };
struct FmtR32F    { static ADEV V4 load(const ImgView& im, int x, int y) { return mk4(__ldg(rowPtr<float>(im, y) + x), 0.0f, 0.0f, 1.0f); } };

struct AddrClamp { static ADEV int wrap(int i, int n) { return min(max(i, 0), n - 1); } };
struct AddrRepeat { static ADEV int wrap(int i, int n) { int m = i % n; return m < 0 ? m + n : m; } };

struct BilinearSetup { int i0, i1, j0, j1; float fx, fy; };
// FAST (only ever true in the fast build, see kFastTex) is for the shading passes: their bar is 1e-3 on colours and no threshold
// test hangs on a footprint.
#ifdef ALTHEA_PARITY
constexpr bool kFastTex = false;
#else
constexpr bool kFastTex = true;
#endif
template <typename Addr, bool FAST = false> ADEV BilinearSetup bilinearSetup(int w, int h, float u, float v) {
  // FAST:
  // floor(x) comes out of the mantissa of u w - 1 + 1.5 * 2^23 (round to nearest of x - 0.5) instead of FRND + F2I on the
  // quarter-rate XU pipe. A coordinate exactly on a texel centre may take the footprint to its left with weight 1 (the same
  // value); NaN / huge coordinates give an arbitrary index, which Addr::wrap brings inside the image like any other.
  if (FAST) {
    const float kMagic = 12582912.0f;
    const float fw = (float)w, fh = (float)h;
    const float mx = fmaf(u, fw, kMagic - 1.0f), my = fmaf(v, fh, kMagic - 1.0f);
    BilinearSetup s;
    s.fx = fmaf(u, fw, -0.5f) - (mx - kMagic);
    s.fy = fmaf(v, fh, -0.5f) - (my - kMagic);
    const int ix = __float_as_int(mx) - 0x4b400000, iy = __float_as_int(my) - 0x4b400000;
    s.i0 = Addr::wrap(ix, w); s.i1 = Addr::wrap(ix + 1, w);
    s.j0 = Addr::wrap(iy, h); s.j1 = Addr::wrap(iy + 1, h);
    return s;
  }
  // rule A1. __fmul_rn/__fsub_rn are never contracted: the coordinate must round the same way in both builds.
  float x = __fsub_rn(__fmul_rn(u, (float)w), 0.5f), y = __fsub_rn(__fmul_rn(v, (float)h), 0.5f);
  if (!(x == x)) x = 0.0f;
  if (!(y == y)) y = 0.0f;
  float fx0 = floorf(x), fy0 = floorf(y);
  BilinearSetup s;
  s.fx = x - fx0;
  s.fy = y - fy0;
  int ix = (int)fx0, iy = (int)fy0;
  s.i0 = Addr::wrap(ix, w); s.i1 = Addr::wrap(ix + 1, w);
  s.j0 = Addr::wrap(iy, h); s.j1 = Addr::wrap(iy + 1, h);
  return s;
}
template <typename Fmt, typename Addr, bool FAST = false> ADEV V4 bilinear(const ImgView& im, float u, float v) {
  BilinearSetup s = bilinearSetup<Addr, FAST>(im.w, im.h, u, v);
  V4 t00 = Fmt::load(im, s.i0, s.j0), t10 = Fmt::load(im, s.i1, s.j0);
  V4 t01 = Fmt::load(im, s.i0, s.j1), t11 = Fmt::load(im, s.i1, s.j1);
  return mix4(mix4(t00, t10, s.fx), mix4(t01, t11, s.fx), s.fy);
}
template <typename Addr, bool FAST = false> ADEV float bilinearR32F(const ImgView& im, float u, float v) {
  BilinearSetup s = bilinearSetup<Addr, FAST>(im.w, im.h, u, v);
  const float* r0 = rowPtr<float>(im, s.j0);
  const float* r1 = rowPtr<float>(im, s.j1);
  float t00 = __ldg(r0 + s.i0), t10 = __ldg(r0 + s.i1), t01 = __ldg(r1 + s.i0), t11 = __ldg(r1 + s.i1);
  return mixf(mixf(t00, t10, s.fx), mixf(t01, t11, s.fx), s.fy);
}
// rule A5: explicit LOD clamped to [0, mips-1], LINEAR mip mode
template <typename Fmt, typename Addr, bool FAST = false> ADEV V4 trilinear(const ChainView& c, float u, float v, float lod) {
  float maxLod = (float)(c.mips - 1);
  if (!(lod == lod)) lod = 0.0f;
  lod = fminf(fmaxf(lod, 0.0f), maxLod);
  float l0f = floorf(lod);
  int l0 = (int)l0f;
  float f = lod - l0f;
  V4 s0 = bilinear<Fmt, Addr, FAST>(c.level[l0], u, v);
  if (f == 0.0f) return s0;
  int l1 = min(l0 + 1, c.mips - 1);
  V4 s1 = bilinear<Fmt, Addr, FAST>(c.level[l1], u, v);
  return mix4(s0, s1, f);
}

// ---- hash RNG of the reference (SSAO.glsl:5-11, PreFilterEnvMap.comp:36-42): state after k draws is seed + k ----
struct HashRng {
  uint32_t sx, sy;
  ADEV uint32_t nextU() {
    sx += 1u; sy += 1u;
    uint32_t qx = 1103515245u * ((sx >> 1) ^ sy);
    uint32_t qy = 1103515245u * ((sy >> 1) ^ sx);
    return 1103515245u * (qx ^ (qy >> 3));
  }
  ADEV float next() { return __fmul_rn((float)nextU(), 2.3283064365386963e-10f); } // 1.0/float(0xffffffff) == 2^-32
};

struct TangentFrame { V3 tan, bit, nor; };
ADEV TangentFrame localToWorld(V3 n) { // coordinateSystem + LocalToWorld (SSAO.glsl:15-27)
  TangentFrame f;
  f.nor = n;
  if (fabsf(n.x) > fabsf(n.y)) f.tan = mk3(-n.z, 0.0f, n.x) / sqrtf(n.x * n.x + n.z * n.z);
  else f.tan = mk3(0.0f, n.z, -n.y) / sqrtf(n.y * n.y + n.z * n.z);
  f.bit = cross3(n, f.tan);
  return f;
}
ADEV V3 frameApply(const TangentFrame& f, V3 v) { return (f.tan * v.x + f.bit * v.y) + f.nor * v.z; }

} // namespace ALTHEA_NS
