// ORACLE — TEST INFRASTRUCTURE ONLY. The GLSL the reference's shaders are written in, as a C++17 library, so that the shader TEXT
// (translated token for token by oracle/glsl2cpp.py, never copied into this repository) compiles with g++ and runs on the CPU:
// oracle/_ref/libshader_ref.so. What lives here is the part of GLSL a GPU driver supplies: vector / matrix types with swizzles,
// the built-in functions the path's shaders call, and a texture unit. Where GLSL leaves the order of a built-in's arithmetic to
// the implementation (dot, normalize, reflect, mix, matrix products) this header takes the order oracle_math.h takes, and the
// texture unit IS oracle_math.h's: differences between the restatement and the executed shader text are then differences in
// reading the text, which is what this library exists to find.
#pragma once
#include "oracle_math.h"
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef uint32_t uint;

// ---- swizzle proxies: a view of some components of the vector they are a union member of ---------------------------
template <class V, class T, int N, int... I> struct Swz {
  T d[N];
  operator V() const { return V(d[I]...); }
  Swz& operator=(const V& v) { int k = 0; ((d[I] = v[k++]), ...); return *this; }
  Swz& operator=(const Swz& o) { return *this = (V)o; } // a.rgb = b.rgb writes three components, not the union's storage
  Swz& operator+=(const V& v) { int k = 0; ((d[I] = d[I] + v[k++]), ...); return *this; }
  Swz& operator-=(const V& v) { int k = 0; ((d[I] = d[I] - v[k++]), ...); return *this; }
  Swz& operator*=(T s) { ((d[I] = d[I] * s), ...); return *this; }
  Swz& operator/=(T s) { ((d[I] = d[I] / s), ...); return *this; }
};

// Constructors take their scalars as templates because glsl2cpp.py writes every GLSL constructor call with braces: GLSL evaluates
// arguments left to right (vec3(rng(), rng(), rng()) draws x first), C++ guarantees that order only inside a braced list, and a
// braced list refuses narrowing conversions (an int where a float is declared).
template <class T> struct tvec2;
template <class T> struct tvec3;
template <class T> struct tvec4;

template <class T> struct tvec2 {
  union {
    struct { T x, y; };
    struct { T r, g; };
    Swz<tvec2<T>, T, 2, 0, 1> xy, rg;
    Swz<tvec2<T>, T, 2, 1, 0> yx;
  };
  tvec2(const tvec2& o) { x = o.x; y = o.y; }
  tvec2& operator=(const tvec2& o) { x = o.x; y = o.y; return *this; }
  tvec2() { x = 0; y = 0; }
  explicit tvec2(T s) { x = s; y = s; }
  template <class A, class B> tvec2(A a, B b) { x = (T)a; y = (T)b; }
  template <class U> explicit tvec2(const tvec2<U>& o) { x = (T)o.x; y = (T)o.y; }
  template <class U, int N, int A, int B> explicit tvec2(const Swz<tvec2<U>, U, N, A, B>& s) { const tvec2<U> o = s; x = (T)o.x; y = (T)o.y; } // uvec2(v.xy)
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};
template <class T> struct tvec3 {
  union {
    struct { T x, y, z; };
    struct { T r, g, b; };
    Swz<tvec2<T>, T, 3, 0, 1> xy, rg;
    Swz<tvec2<T>, T, 3, 0, 2> xz;
    Swz<tvec2<T>, T, 3, 1, 2> yz;
    Swz<tvec3<T>, T, 3, 0, 1, 2> xyz, rgb;
  };
  tvec3(const tvec3& o) { x = o.x; y = o.y; z = o.z; }
  tvec3& operator=(const tvec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
  tvec3() { x = 0; y = 0; z = 0; }
  explicit tvec3(T s) { x = s; y = s; z = s; }
  template <class A, class B, class C> tvec3(A a, B b, C c) { x = (T)a; y = (T)b; z = (T)c; }
  template <class C> tvec3(const tvec2<T>& a, C c) { x = a.x; y = a.y; z = (T)c; }
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};
template <class T> struct tvec4 {
  union {
    struct { T x, y, z, w; };
    struct { T r, g, b, a; };
    Swz<tvec2<T>, T, 4, 0, 1> xy, rg;
    Swz<tvec2<T>, T, 4, 0, 2> xz;
    Swz<tvec3<T>, T, 4, 0, 1, 2> xyz, rgb;
    Swz<tvec4<T>, T, 4, 0, 1, 2, 3> xyzw, rgba;
    Swz<tvec2<T>, T, 4, 2, 1> bg;
  };
  tvec4(const tvec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; }
  tvec4& operator=(const tvec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
  tvec4() { x = 0; y = 0; z = 0; w = 0; }
  explicit tvec4(T s) { x = s; y = s; z = s; w = s; }
  template <class A, class B, class C, class D> tvec4(A a, B b, C c, D d) { x = (T)a; y = (T)b; z = (T)c; w = (T)d; }
  template <class D> tvec4(const tvec3<T>& a, D d) { x = a.x; y = a.y; z = a.z; w = (T)d; }
  template <class C, class D> tvec4(const tvec2<T>& a, C c, D d) { x = a.x; y = a.y; z = (T)c; w = (T)d; }
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};
template <class T, int N> struct arr { // a GLSL array: a value
  T v[N];
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
};
typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<uint> uvec2;
typedef tvec3<uint> uvec3;
typedef tvec4<uint> uvec4;
typedef tvec2<int> ivec2;

// ---- operators: concrete (non-template) overloads, so that swizzle proxies convert ----------------------------------
#define GLSL_VEC_OPS(V, T, N)                                                                                          \
  inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }        \
  inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }        \
  inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; }        \
  inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; }        \
  inline V operator+(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + s; return r; }                  \
  inline V operator-(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - s; return r; }                  \
  inline V operator*(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }                  \
  inline V operator/(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }                  \
  inline V operator+(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s + a[i]; return r; }                  \
  inline V operator-(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s - a[i]; return r; }                  \
  inline V operator*(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r; }                  \
  inline V operator/(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s / a[i]; return r; }                  \
  inline V& operator+=(V& a, const V& b) { a = a + b; return a; }                                                      \
  inline V& operator-=(V& a, const V& b) { a = a - b; return a; }                                                      \
  inline V& operator*=(V& a, const V& b) { a = a * b; return a; }                                                      \
  inline V& operator/=(V& a, const V& b) { a = a / b; return a; }                                                      \
  inline V& operator+=(V& a, T s) { a = a + s; return a; }                                                             \
  inline V& operator-=(V& a, T s) { a = a - s; return a; }                                                             \
  inline V& operator*=(V& a, T s) { a = a * s; return a; }                                                             \
  inline V& operator/=(V& a, T s) { a = a / s; return a; }                                                             \
  inline bool operator==(const V& a, const V& b) { for (int i = 0; i < N; ++i) if (!(a[i] == b[i])) return false; return true; } \
  inline bool operator!=(const V& a, const V& b) { return !(a == b); }
GLSL_VEC_OPS(vec2, float, 2)
GLSL_VEC_OPS(vec3, float, 3)
GLSL_VEC_OPS(vec4, float, 4)
GLSL_VEC_OPS(uvec2, uint, 2)
GLSL_VEC_OPS(uvec3, uint, 3)
GLSL_VEC_OPS(uvec4, uint, 4)
GLSL_VEC_OPS(ivec2, int, 2)
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }
#define GLSL_UINT_OPS(V, N)                                                                                            \
  inline V operator>>(const V& a, uint s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] >> s; return r; }             \
  inline V operator<<(const V& a, uint s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] << s; return r; }             \
  inline V operator>>(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] >> b[i]; return r; }      \
  inline V operator^(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] ^ b[i]; return r; }        \
  inline V operator&(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] & b[i]; return r; }        \
  inline V& operator^=(V& a, const V& b) { a = a ^ b; return a; }
GLSL_UINT_OPS(uvec2, 2)
GLSL_UINT_OPS(uvec3, 3)
GLSL_UINT_OPS(uvec4, 4)

// ---- built-in functions (GLSL 4.60 section 8), scalar arithmetic in fp32 --------------------------------------------
inline float abs(float x) { return ::fabsf(x); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float tan(float x) { return ::tanf(x); }
inline float atan(float y, float x) { return ::atan2f(y, x); }
inline float atan(float x) { return ::atanf(x); }
inline float acos(float x) { return ::acosf(x); }
inline float asin(float x) { return ::asinf(x); }
inline float exp(float x) { return ::expf(x); }
inline float log2(float x) { return ::log2f(x); }
inline float log(float x) { return ::logf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float floor(float x) { return ::floorf(x); }
inline float fract(float x) { return x - ::floorf(x); }
inline float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
inline float radians(float d) { return d * (3.14159265358979323846f / 180.0f); }
inline bool isinf(float x) { return __builtin_isinf(x); }
inline bool isnan(float x) { return x != x; }
inline float max(float a, float b) { return a < b ? b : a; } // GLSL: y if x < y, otherwise x
inline float min(float a, float b) { return b < a ? b : a; } // GLSL: y if y < x, otherwise x
inline int max(int a, int b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; } // x (1 - a) + y a, as the spec writes it
inline vec2 mix(const vec2& a, const vec2& b, float t) { return vec2(mix(a.x, b.x, t), mix(a.y, b.y, t)); }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return vec3(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)); }
inline vec4 mix(const vec4& a, const vec4& b, float t) { return vec4(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t), mix(a.w, b.w, t)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 exp(const vec3& a) { return vec3(exp(a.x), exp(a.y), exp(a.z)); }
inline vec3 abs(const vec3& a) { return vec3(abs(a.x), abs(a.y), abs(a.z)); }
inline vec3 pow(const vec3& a, const vec3& b) { return vec3(pow(a.x, b.x), pow(a.y, b.y), pow(a.z, b.z)); }
inline vec3 clamp(const vec3& a, float lo, float hi) { return vec3(clamp(a.x, lo, hi), clamp(a.y, lo, hi), clamp(a.z, lo, hi)); }
// the orders of oracle_math.h (dot: (x x' + y y') + z z'; normalize: v / length(v); reflect: I - (2 dot(N, I)) N)
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(const vec2& a) { return sqrt(dot(a, a)); }
inline float length(const vec3& a) { return sqrt(dot(a, a)); }
inline float length(const vec4& a) { return sqrt(dot(a, a)); }
inline vec2 normalize(const vec2& a) { return a / length(a); }
inline vec3 normalize(const vec3& a) { return a / length(a); }
inline vec3 reflect(const vec3& i, const vec3& n) { return i - (2.0f * dot(n, i)) * n; }

// ---- matrices: column-major, m[c] is a column -------------------------------------------------------------------------
struct mat4;
struct mat3 {
  vec3 c[3];
  mat3() {}
  mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
  explicit mat3(const mat4& m);
  vec3& operator[](int i) { return c[i]; }
  const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
  vec4 c[4];
  mat4() {}
  explicit mat4(float d) { c[0] = vec4(d, 0.0f, 0.0f, 0.0f); c[1] = vec4(0.0f, d, 0.0f, 0.0f); c[2] = vec4(0.0f, 0.0f, d, 0.0f); c[3] = vec4(0.0f, 0.0f, 0.0f, d); }
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};
inline mat4 operator*(float s, const mat4& m) { mat4 r; for (int i = 0; i < 4; ++i) r[i] = s * m[i]; return r; }
inline mat4 operator+(const mat4& a, const mat4& b) { mat4 r; for (int i = 0; i < 4; ++i) r[i] = a[i] + b[i]; return r; }
inline mat4& operator+=(mat4& a, const mat4& b) { a = a + b; return a; }
inline mat3::mat3(const mat4& m) { for (int i = 0; i < 3; ++i) c[i] = vec3(m[i].x, m[i].y, m[i].z); }
static_assert(sizeof(mat4) == 64 && sizeof(vec2) == 8 && sizeof(vec3) == 12 && sizeof(vec4) == 16, "tight layouts");
// oracle_math.h's orders: mat4 * vec4 = ((c0 x + c1 y) + c2 z) + c3 w per row; mat3 * vec3 = (c0 x + c1 y) + c2 z
inline vec4 operator*(const mat4& m, const vec4& v) {
  vec4 r;
  for (int i = 0; i < 4; ++i) r[i] = ((m[0][i] * v.x + m[1][i] * v.y) + m[2][i] * v.z) + m[3][i] * v.w;
  return r;
}
inline vec3 operator*(const mat3& m, const vec3& v) {
  vec3 r;
  for (int i = 0; i < 3; ++i) r[i] = (m[0][i] * v.x + m[1][i] * v.y) + m[2][i] * v.z;
  return r;
}
inline mat4 operator*(const mat4& a, const mat4& b) { mat4 r; for (int j = 0; j < 4; ++j) r[j] = a * b[j]; return r; }
inline mat3 operator*(const mat3& a, const mat3& b) { mat3 r; for (int j = 0; j < 3; ++j) r[j] = a * b[j]; return r; }
inline mat3 transpose(const mat3& m) { mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i][j] = m[j][i]; return r; }

// ---- texture unit: oracle_math.h's (Vulkan's LINEAR filter, CLAMP_TO_EDGE / REPEAT, explicit LOD, cube face selection) -----
struct sampler2D {
  oracle::TexChain chain{nullptr, 0, 0, 1, oracle::FMT_RGBA32F};
  oracle::Address address = oracle::ADDR_CLAMP;
  // material textures of the rasterising passes: an OracleTex of althea_oracle_raster.cpp (RGBA8 mip chain + sampler word; null texels
  // = the engine's 1 x 1 default of that slot), sampled by that file's texture unit at the implicit level of detail
  const void* materialTex = nullptr;
  float dflt[4] = {1.0f, 1.0f, 1.0f, 1.0f};
};
// implicit derivatives: a fragment stage samples its material textures at one of its interpolated uv sets, whose screen-space
// derivatives the rasteriser knows; the stage driver announces them, texture() picks the set its coordinate is
struct ImplicitLod { const float (*uvs)[2] = nullptr; const float (*ddx)[2] = nullptr; const float (*ddy)[2] = nullptr; int sets = 0; };
inline thread_local ImplicitLod gImplicit;
inline void (*gSampleMaterial)(const void* tex, const float dflt[4], const float uv[2], const float ddx[2], const float ddy[2], float out[4]) = nullptr;
struct samplerCubeArray { const float* layers = nullptr; int res = 0; }; // R32F, layer = 6 * cube + face
struct image2D { uint16_t* texels = nullptr; float* texels32 = nullptr; int w = 0, h = 0; }; // RGBA16F (or RGBA32F) storage image
inline vec4 fromV4(oracle::V4 v) { return vec4(v.x, v.y, v.z, v.w); }
// Rule A3: a fragment's uv is its pixel centre ((x + .5) / W, (y + .5) / H), and a fetch AT THAT uv from an image of
// the frame's size returns the pixel's own texel exactly: a texture unit's fixed-point weights (8 sub-texel bits) are 0 there,
// while u W - 0.5 in fp32 would leave a weight of a few 1e-5 on a neighbour. The stage driver announces the fragment it runs.
struct FragCtx { float u = 0.0f, v = 0.0f; int x = 0, y = 0, W = 0, H = 0; bool on = false; };
inline thread_local FragCtx gFrag;
inline oracle::V4 fetchLevel0(const sampler2D& s, const vec2& uv) {
  const oracle::Tex t = s.chain.level(0);
  if (gFrag.on && t.w == gFrag.W && t.h == gFrag.H && uv.x == gFrag.u && uv.y == gFrag.v) return oracle::texel(t, gFrag.x, gFrag.y);
  return oracle::bilinear(t, uv.x, uv.y, s.address);
}
// rule A5 (oracle_math.h's trilinear): explicit LOD clamped to the chain, LINEAR between the two levels
inline vec4 textureLod(const sampler2D& s, const vec2& uv, float lod) {
  if (!(lod == lod)) lod = 0.0f;
  lod = oracle::clampf(lod, 0.0f, (float)(s.chain.mips - 1));
  const float l0f = ::floorf(lod);
  const int l0 = (int)l0f;
  const float f = lod - l0f;
  const oracle::V4 s0 = l0 == 0 ? fetchLevel0(s, uv) : oracle::bilinear(s.chain.level(l0), uv.x, uv.y, s.address);
  if (f == 0.0f) return fromV4(s0);
  const int l1 = l0 + 1 < s.chain.mips ? l0 + 1 : s.chain.mips - 1;
  return fromV4(oracle::mix(s0, oracle::bilinear(s.chain.level(l1), uv.x, uv.y, s.address), f));
}
// implicit-LOD fetch of a fragment shader: every image the path samples this way has one level (rule A10)
inline vec4 texture(const sampler2D& s, const vec2& uv) {
  if (s.materialTex) {
    const float zero[2] = {0.0f, 0.0f}, c[2] = {uv.x, uv.y};
    const float *dx = zero, *dy = zero;
    for (int k = 0; k < gImplicit.sets; ++k)
      if (gImplicit.uvs[k][0] == uv.x && gImplicit.uvs[k][1] == uv.y) { dx = gImplicit.ddx[k]; dy = gImplicit.ddy[k]; break; }
    float out[4];
    gSampleMaterial(s.materialTex, s.dflt, c, dx, dy, out);
    return vec4(out[0], out[1], out[2], out[3]);
  }
  return fromV4(fetchLevel0(s, uv));
}
inline vec4 texture(const samplerCubeArray& s, const vec4& q) { // Vulkan 1.3 spec 16.5.1 (cube map face selection), bilinear inside the face
  const float ax = ::fabsf(q.x), ay = ::fabsf(q.y), az = ::fabsf(q.z);
  int face; float sc, tc, ma;
  if (ax >= ay && ax >= az) { ma = ax; if (q.x >= 0.0f) { face = 0; sc = -q.z; tc = -q.y; } else { face = 1; sc = q.z; tc = -q.y; } }
  else if (ay >= az)        { ma = ay; if (q.y >= 0.0f) { face = 2; sc = q.x; tc = q.z; } else { face = 3; sc = q.x; tc = -q.z; } }
  else                      { ma = az; if (q.z >= 0.0f) { face = 4; sc = q.x; tc = -q.y; } else { face = 5; sc = -q.x; tc = -q.y; } }
  const float u = 0.5f * sc / ma + 0.5f, v = 0.5f * tc / ma + 0.5f;
  const int layer = 6 * (int)q.w + face;
  const float* p = s.layers + (size_t)layer * s.res * s.res;
  return fromV4(oracle::bilinear(oracle::Tex{p, s.res, s.res, oracle::FMT_R32F}, u, v, oracle::ADDR_CLAMP));
}
inline void imageStore(const image2D& img, const ivec2& p, const vec4& c) { // RGBA16F: round to nearest even
  if (img.texels32) { float* t = img.texels32 + ((size_t)p.y * img.w + p.x) * 4; t[0] = c.x; t[1] = c.y; t[2] = c.z; t[3] = c.w; return; }
  uint16_t* t = img.texels + ((size_t)p.y * img.w + p.x) * 4;
  t[0] = oracle::floatToHalf(c.x); t[1] = oracle::floatToHalf(c.y); t[2] = oracle::floatToHalf(c.z); t[3] = oracle::floatToHalf(c.w);
}

// what every shader stage inherits
struct ShaderBase {
  vec4 gl_FragCoord, gl_Position;
  int gl_VertexIndex = 0, gl_ViewIndex = 0;
  uvec3 gl_GlobalInvocationID;
  bool gl_Discarded = false;
};
#define discard do { gl_Discarded = true; return; } while (0)

} // namespace glsl
