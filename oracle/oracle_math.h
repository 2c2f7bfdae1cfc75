// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the GLSL on Althea's deferred screen-space path. Nothing
// under althea_b200/ may include, link or call this; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// This header is the "Vulkan texture unit" + GLSL vector helpers the shaders lean
// on (rules A1-A8). Everything is plain IEEE binary32: the file
// is compiled with -ffp-contract=off and without fast-math, so every + - * / sqrt
// below is one correctly rounded operation, in the order written.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle {

constexpr float kPi = 3.14159265359f; // Shaders/Misc/Constants.glsl:4 (same literal in every shader on the path)

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline V3 normalize(V3 a) { return a / length(a); }
inline V3 reflect(V3 i, V3 n) { return i - (2.0f * dot(n, i)) * n; }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline V3 mix(V3 a, V3 b, float t) { return {mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)}; }
inline V4 mix(V4 a, V4 b, float t) { return {mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t), mixf(a.w, b.w, t)}; }
inline V3 xyz(V4 a) { return {a.x, a.y, a.z}; }
inline float maxf(float a, float b) { return a > b ? a : b; } // GLSL max(): returns b only if a < b
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// column-major mat4 (glm / GLSL layout): m[c*4 + r]
struct M4 { float m[16]; };
inline V4 mul(const M4& M, V4 v) {
  V4 r;
  r.x = ((M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z) + M.m[12] * v.w;
  r.y = ((M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z) + M.m[13] * v.w;
  r.z = ((M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z) + M.m[14] * v.w;
  r.w = ((M.m[3] * v.x + M.m[7] * v.y) + M.m[11] * v.z) + M.m[15] * v.w;
  return r;
}
inline V3 mul3(const M4& M, V3 v) { // mat3(M) * v
  V3 r;
  r.x = (M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z;
  r.y = (M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z;
  r.z = (M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z;
  return r;
}
inline M4 matmul(const M4& A, const M4& B) { // A*B, each column = A * B.col
  M4 R;
  for (int c = 0; c < 4; ++c) {
    V4 col = mul(A, V4{B.m[c * 4 + 0], B.m[c * 4 + 1], B.m[c * 4 + 2], B.m[c * 4 + 3]});
    R.m[c * 4 + 0] = col.x; R.m[c * 4 + 1] = col.y; R.m[c * 4 + 2] = col.z; R.m[c * 4 + 3] = col.w;
  }
  return R;
}

// --- storage formats (rule A7) -------------------------------------------------------------
inline float halfToFloat(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else { // subnormal half -> normal float
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
  else bits = sign | ((exp + 112u) << 23) | (man << 13);
  float f; memcpy(&f, &bits, 4); return f;
}
inline uint16_t floatToHalf(float f) { // round-to-nearest-even, overflow -> inf
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t ax = x & 0x7fffffffu;
  if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));
  if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); // >= 65520 rounds to inf
  if (ax < 0x33000001u) return (uint16_t)sign;              // <= 2^-25 rounds to 0
  int e = (int)(ax >> 23) - 127;
  uint32_t man = (ax & 0x7fffffu) | 0x800000u;
  int shift;
  uint32_t hexp;
  if (e < -14) { shift = 13 + (-14 - e); hexp = 0; }
  else { shift = 13; hexp = (uint32_t)(e + 15); }
  uint32_t q = man >> shift;
  uint32_t rem = man & ((1u << shift) - 1u);
  uint32_t halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (q & 1u))) ++q;
  uint32_t out = (hexp == 0) ? q : (((hexp - 1u) << 10) + q); // carry propagates into exponent
  return (uint16_t)(sign | out);
}

// --- texture unit ----------------------------------------------------------------------------
enum Format { FMT_RGBA32F = 0, FMT_RGBA16F = 1, FMT_RGBA8 = 2, FMT_R32F = 3 };
enum Address { ADDR_CLAMP = 0, ADDR_REPEAT = 1 };

struct Tex { // one mip level, tightly packed, row 0 = top
  const void* data; int w, h; Format fmt;
};

inline V4 texel(const Tex& t, int x, int y) {
  size_t i = (size_t)y * t.w + x;
  switch (t.fmt) {
  case FMT_RGBA32F: { const float* p = (const float*)t.data + i * 4; return {p[0], p[1], p[2], p[3]}; }
  case FMT_RGBA16F: { const uint16_t* p = (const uint16_t*)t.data + i * 4;
    return {halfToFloat(p[0]), halfToFloat(p[1]), halfToFloat(p[2]), halfToFloat(p[3])}; }
  case FMT_RGBA8: { const uint8_t* p = (const uint8_t*)t.data + i * 4;
    return {p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f}; }
  default: { float d = ((const float*)t.data)[i]; return {d, 0.0f, 0.0f, 1.0f}; }
  }
}
inline int wrapIndex(int i, int n, Address a) {
  if (a == ADDR_CLAMP) return i < 0 ? 0 : (i >= n ? n - 1 : i);
  int m = i % n; return m < 0 ? m + n : m;
}
// rule A1: unnormalised = u*size - 0.5, two taps per axis, lerp of lerps in FP32
inline V4 bilinear(const Tex& t, float u, float v, Address a) {
  float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
  if (!(x == x)) x = 0.0f;
  if (!(y == y)) y = 0.0f;
  float fx0 = floorf(x), fy0 = floorf(y);
  float fx = x - fx0, fy = y - fy0;
  int i0 = wrapIndex((int)fx0, t.w, a), i1 = wrapIndex((int)fx0 + 1, t.w, a);
  int j0 = wrapIndex((int)fy0, t.h, a), j1 = wrapIndex((int)fy0 + 1, t.h, a);
  V4 top = mix(texel(t, i0, j0), texel(t, i1, j0), fx);
  V4 bot = mix(texel(t, i0, j1), texel(t, i1, j1), fx);
  return mix(top, bot, fy);
}

struct TexChain { // tightly packed mip chain, level k has size max(1, w>>k) x max(1, h>>k)
  const void* data; int w, h, mips; Format fmt;
  Tex level(int k) const {
    size_t bpp = (fmt == FMT_RGBA32F) ? 16 : (fmt == FMT_RGBA16F ? 8 : 4);
    size_t off = 0; int lw = w, lh = h;
    for (int i = 0; i < k; ++i) { off += (size_t)lw * lh * bpp; lw = lw > 1 ? lw >> 1 : 1; lh = lh > 1 ? lh >> 1 : 1; }
    return Tex{(const uint8_t*)data + off, lw, lh, fmt};
  }
};
// rule A5: explicit LOD, clamp to [0, mips-1], LINEAR mip mode
inline V4 trilinear(const TexChain& c, float u, float v, float lod, Address a) {
  float maxLod = (float)(c.mips - 1);
  if (!(lod == lod)) lod = 0.0f;
  lod = clampf(lod, 0.0f, maxLod);
  float l0f = floorf(lod);
  int l0 = (int)l0f;
  float f = lod - l0f;
  V4 s0 = bilinear(c.level(l0), u, v, a);
  if (f == 0.0f) return s0;
  int l1 = l0 + 1 < c.mips ? l0 + 1 : c.mips - 1;
  V4 s1 = bilinear(c.level(l1), u, v, a);
  return mix(s0, s1, f);
}

// --- hash RNG (SSAO.glsl:5-11, PreFilterEnvMap.comp:36-42, Misc/Sampling.glsl:13-18) ---------
struct Rng {
  uint32_t sx, sy;
  uint32_t nextU() {
    sx += 1u; sy += 1u;
    uint32_t qx = 1103515245u * ((sx >> 1) ^ sy);
    uint32_t qy = 1103515245u * ((sy >> 1) ^ sx);
    return 1103515245u * (qx ^ (qy >> 3));
  }
  // float(n) * (1.0 / float(0xffffffffU)): float(0xffffffff) rounds to 2^32, so the scale is 2^-32
  float next() { return (float)nextU() * (1.0f / 4294967296.0f); }
};

// coordinateSystem / LocalToWorld (SSAO.glsl:15-27, GenIrradianceMap.comp:19-31)
struct Frame { V3 tan, bit, nor; V3 apply(V3 v) const { return (tan * v.x + bit * v.y) + nor * v.z; } };
inline Frame localToWorld(V3 n) {
  Frame f; f.nor = n;
  if (fabsf(n.x) > fabsf(n.y)) f.tan = V3{-n.z, 0.0f, n.x} / sqrtf(n.x * n.x + n.z * n.z);
  else f.tan = V3{0.0f, n.z, -n.y} / sqrtf(n.y * n.y + n.z * n.z);
  f.bit = cross(n, f.tan);
  return f;
}

} // namespace oracle
