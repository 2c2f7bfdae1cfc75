// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the two rasterising producers of the deferred path's inputs:
//   oracle_draw_gbuffer       <- Shaders/Gltf/Gltf.vert:37-60, Shaders/Gltf/Gltf.frag:28-51,
//                                Shaders/InstanceData/InstanceData.glsl:28-69 (fetchMaterial),
//                                pipeline state: Src/Primitive.cpp:23-48, Include/Althea/GraphicsPipeline.h:171-172
//                                (cull BACK, front CCW, dynamic front face), Src/GraphicsPipeline.cpp:176 (depth LESS),
//                                clears Src/DeferredRendering.cpp:101-106 (colour 0, depth 1)
//   oracle_draw_shadow_cubes  <- Shaders/ShadowMapBindless.vert:33-50, Shaders/ShadowMapBindless.frag:22-42,
//                                Src/PointLight.cpp:235-282 (one multiview pass per light, 6 views)
// PINNING. The PROGRAMMABLE stages below (vertex transform, TBN, fetchMaterial, alpha test, the three outputs; the shadow pass's
// light-space position and depth) are pinned by the reference's own shader text, executed: with the stage hooks further down
// installed (oracle/ref_shader_driver.cpp runs Gltf.vert / Gltf.frag / ShadowMapBindless.vert / .frag from their text) a draw
// reproduces the plain draw bit for bit on every attachment (tests/test_shader_ref.py, vectors in tests/golden/shader_ref.npz).
// The FIXED-FUNCTION part stays unpinned: the reference ships no golden G-buffer or shadow map, and rasterisation (sub-pixel
// snapping, derivative quads, anisotropic filtering) is implementation-defined in Vulkan. This file states the rules the CUDA path
// implements, in scalar form and in draw order, so that the two can be compared bit for bit on coverage, triangle ids and depth:
//   * a fragment exists where the pixel centre is inside the triangle (top-left rule), evaluated with edge functions in
//     2-D homogeneous coordinates (x, y, w) so triangles that cross the eye plane need no geometric clipping;
//   * depth clip 0 <= z_c <= w_c, back faces culled by the sign of the homogeneous determinant;
//   * triangles are processed in draw order with a LESS depth test against a 1.0 clear (first fragment wins ties); colour
//     attachments are alpha-blended as every pipeline of the reference does (Src/GraphicsPipeline.cpp:138-154);
//   * attributes are interpolated perspective-correctly; texture level of detail comes from analytic uv derivatives,
//     isotropic (rho = max(|d uv/dx * size|, |d uv/dy * size|)), trilinear.
// Build: g++ -O2 -fopenmp -ffp-contract=off -fno-fast-math (oracle/Makefile): no a*b+c is ever fused unless std::fmaf says so.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct V4 { float x, y, z, w; };

// column-major mat4 * vec4, summed left to right
V4 mulMV(const float* m, float x, float y, float z, float w) {
  V4 r;
  r.x = ((m[0] * x + m[4] * y) + m[8] * z) + m[12] * w;
  r.y = ((m[1] * x + m[5] * y) + m[9] * z) + m[13] * w;
  r.z = ((m[2] * x + m[6] * y) + m[10] * z) + m[14] * w;
  r.w = ((m[3] * x + m[7] * y) + m[11] * z) + m[15] * w;
  return r;
}
void mulM3V(const float* m, const float* v, float* out) {
  out[0] = (m[0] * v[0] + m[4] * v[1]) + m[8] * v[2];
  out[1] = (m[1] * v[0] + m[5] * v[1]) + m[9] * v[2];
  out[2] = (m[2] * v[0] + m[6] * v[1]) + m[10] * v[2];
}
void matmul44(const float* A, const float* B, float* R) { // A * B, column by column
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float s = A[0 * 4 + r] * B[c * 4 + 0];
      s = s + A[1 * 4 + r] * B[c * 4 + 1];
      s = s + A[2 * 4 + r] * B[c * 4 + 2];
      s = s + A[3 * 4 + r] * B[c * 4 + 3];
      R[c * 4 + r] = s;
    }
}

} // namespace

extern "C" {

struct OracleTex { // RGBA8, levels packed one after another; sampler word as in include/althea_cuda.h
  const uint8_t* texels;
  int32_t w, h, mips;
  uint32_t sampler;
};
struct OraclePrim {
  const float* verts; // the engine's Vertex, 26 floats apart (InstanceDataCommon.h:45-53): position 0, tangent 3, bitangent 6, normal 9, uvs 12
  const uint32_t* idx;
  uint32_t triCount;
  uint32_t frontCW;
  float model[16];
  float baseColorFactor[4];
  int32_t baseUv, mrUv;
  float normalScale, metallicFactor, roughnessFactor, alphaCutoff;
  OracleTex base, normal, mr;
};

// Stage hooks (oracle/ref_shader_driver.cpp installs them): the programmable stages of the two passes taken from the reference's
// SHADER TEXT, executed, while the fixed-function part (coverage, depth test, interpolation, derivatives, blending) stays this
// file's. A draw with hooks must reproduce the draw without them: that is the pin of the stages' restatement below.
struct OracleStageHooks {
  // Gltf/Gltf.vert main for vertex i of p: gl_Position, worldPosition, vertTbn (columns tangent, bitangent, normal)
  void (*gbufferVertex)(const OraclePrim* p, uint32_t i, const float* projection, const float* view, float clip[4], float world[3], float tbn[9]);
  // Gltf/Gltf.frag main on interpolated inputs (uv sets with their screen-space derivatives); returns 0 when the fragment is discarded
  int (*gbufferFragment)(const OraclePrim* p, const float tbn[9], const float uvs[4][2], const float ddx[4][2], const float ddy[4][2],
                         float outNormal[4], float outAlbedo[4], float outMro[4]);
  // ShadowMapBindless.vert main for view `view` of the light at lightPos: gl_Position and worldPosCS
  void (*shadowVertex)(const OraclePrim* p, uint32_t i, const float* lightPos, const float* view, const float* projection, float clip[4], float cs[3]);
  // ShadowMapBindless.frag main: gl_FragDepth, or 0 returned on discard
  int (*shadowFragment)(const OraclePrim* p, const float cs[3], const float uv[2], const float ddx[2], const float ddy[2], float* depth);
};

} // extern "C"

namespace {

const OracleStageHooks* g_hooks = nullptr;
constexpr int kVertexFloats = 26;

float srgbToLinear(int i) {
  const double c = i / 255.0;
  return (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
}
const float* srgbTable() {
  static float t[256];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 256; ++i) t[i] = srgbToLinear(i);
    init = true;
  }
  return t;
}

int wrapIndex(int i, int n, uint32_t mode) {
  if (mode == 1u) return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
  if (mode == 2u) {
    int m = i % (2 * n);
    if (m < 0) m += 2 * n;
    m -= n;
    if (m < 0) m = -(1 + m);
    return (n - 1) - m;
  }
  int m = i % n;
  return m < 0 ? m + n : m;
}
V4 texel(const OracleTex& t, const uint8_t* level, int w, int x, int y) {
  const uint8_t* p = level + ((size_t)y * w + x) * 4;
  V4 r;
  if (t.sampler & 0x100u) {
    const float* lut = srgbTable();
    r.x = lut[p[0]]; r.y = lut[p[1]]; r.z = lut[p[2]];
  } else {
    r.x = (float)p[0] / 255.0f; r.y = (float)p[1] / 255.0f; r.z = (float)p[2] / 255.0f;
  }
  r.w = (float)p[3] / 255.0f;
  return r;
}
V4 lerp4(V4 a, V4 b, float t) {
  const float o = 1.0f - t;
  return V4{a.x * o + b.x * t, a.y * o + b.y * t, a.z * o + b.z * t, a.w * o + b.w * t};
}
V4 sampleLevel(const OracleTex& t, int level, float u, float v, bool nearest) {
  const uint8_t* base = t.texels;
  int w = t.w, h = t.h;
  for (int k = 0; k < level; ++k) {
    base += (size_t)w * h * 4;
    w = (w >> 1) > 1 ? (w >> 1) : 1;
    h = (h >> 1) > 1 ? (h >> 1) : 1;
  }
  const uint32_t wu = t.sampler & 3u, wv = (t.sampler >> 2) & 3u;
  if (nearest) return texel(t, base, w, wrapIndex((int)std::floor(u * (float)w), w, wu), wrapIndex((int)std::floor(v * (float)h), h, wv));
  const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
  const float fx0 = std::floor(x), fy0 = std::floor(y);
  const float fx = x - fx0, fy = y - fy0;
  const int ix = (int)fx0, iy = (int)fy0;
  const int i0 = wrapIndex(ix, w, wu), i1 = wrapIndex(ix + 1, w, wu), j0 = wrapIndex(iy, h, wv), j1 = wrapIndex(iy + 1, h, wv);
  return lerp4(lerp4(texel(t, base, w, i0, j0), texel(t, base, w, i1, j0), fx), lerp4(texel(t, base, w, i0, j1), texel(t, base, w, i1, j1), fx), fy);
}
V4 sampleTexture(const OracleTex& t, V4 dflt, const float uv[2], const float ddx[2], const float ddy[2]) {
  if (!t.texels) return dflt;
  const float ax = ddx[0] * (float)t.w, ay = ddx[1] * (float)t.h, bx = ddy[0] * (float)t.w, by = ddy[1] * (float)t.h;
  const float rho2 = std::fmax(ax * ax + ay * ay, bx * bx + by * by);
  float lod = 0.5f * std::log2(rho2);
  if (!(lod == lod)) lod = 0.0f;
  const bool magNearest = t.sampler & 0x10u, minNearest = t.sampler & 0x20u;
  const uint32_t mipMode = (t.sampler >> 6) & 3u;
  if (lod <= 0.0f) return sampleLevel(t, 0, uv[0], uv[1], magNearest);
  if (mipMode == 0u || t.mips <= 1) return sampleLevel(t, 0, uv[0], uv[1], minNearest);
  lod = std::fmin(lod, (float)(t.mips - 1));
  if (mipMode == 1u) {
    int l = (int)std::ceil(lod + 0.5f) - 1;
    l = l < 0 ? 0 : (l > t.mips - 1 ? t.mips - 1 : l);
    return sampleLevel(t, l, uv[0], uv[1], minNearest);
  }
  const float l0f = std::floor(lod);
  const int l0 = (int)l0f;
  const float f = lod - l0f;
  const V4 s0 = sampleLevel(t, l0, uv[0], uv[1], minNearest);
  if (f == 0.0f || l0 + 1 >= t.mips) return s0;
  return lerp4(s0, sampleLevel(t, l0 + 1, uv[0], uv[1], minNearest), f);
}

struct Tri {
  uint32_t vi[3];
  V4 world[3], clip[3];
  float cs[3][3];
  float tbn[3][9]; // G-buffer pass: mat3(model) * tbn per vertex (Gltf.vert:59), columns tangent, bitangent, normal
  float A[3], B[3], C[3], rdet;
  bool ok;
};

// vertex stage + triangle setup for one view. viewA/viewB/off: clip = viewB ? viewB * (viewA * (world - off)) : viewA * world
// hookProj / hookView: the unmultiplied matrices of the G-buffer pass, for the vertex hook (viewA is their product there)
Tri setupTriangle(const OraclePrim& p, uint32_t t, const float* viewA, const float* viewB, const float* off, int W, int H,
                  const float* hookProj = nullptr, const float* hookView = nullptr) {
  Tri g;
  g.ok = false;
  for (int k = 0; k < 3; ++k) {
    const uint32_t i = p.idx[3u * t + k];
    g.vi[k] = i;
    const float* pos = p.verts + (size_t)i * kVertexFloats;
    if (g_hooks && !viewB && g_hooks->gbufferVertex && hookProj) {
      float c[4], w3[3];
      g_hooks->gbufferVertex(&p, i, hookProj, hookView, c, w3, g.tbn[k]);
      g.clip[k] = V4{c[0], c[1], c[2], c[3]};
      g.world[k] = V4{w3[0], w3[1], w3[2], 1.0f};
      g.cs[k][0] = g.cs[k][1] = g.cs[k][2] = 0.0f;
      continue;
    }
    if (g_hooks && viewB && g_hooks->shadowVertex) {
      float c[4];
      g_hooks->shadowVertex(&p, i, off, viewA, viewB, c, g.cs[k]);
      g.clip[k] = V4{c[0], c[1], c[2], c[3]};
      g.world[k] = V4{0, 0, 0, 1};
      continue;
    }
    const V4 w = mulMV(p.model, pos[0], pos[1], pos[2], 1.0f);
    g.world[k] = w;
    if (!viewB) { mulM3V(p.model, pos + 3, g.tbn[k]); mulM3V(p.model, pos + 6, g.tbn[k] + 3); mulM3V(p.model, pos + 9, g.tbn[k] + 6); }
    if (viewB) {
      const V4 c = mulMV(viewA, w.x - off[0], w.y - off[1], w.z - off[2], w.w);
      g.clip[k] = mulMV(viewB, c.x, c.y, c.z, c.w);
      g.cs[k][0] = c.x / c.w; g.cs[k][1] = c.y / c.w; g.cs[k][2] = c.z / c.w;
    } else {
      g.clip[k] = mulMV(viewA, w.x, w.y, w.z, w.w);
      g.cs[k][0] = g.cs[k][1] = g.cs[k][2] = 0.0f;
    }
  }
  // adjugate rows of M = [x; y; w]: r_i = v_{i+1} x v_{i+2}
  float rx[3], ry[3], rz[3];
  for (int i = 0; i < 3; ++i) {
    const V4 a = g.clip[(i + 1) % 3], b = g.clip[(i + 2) % 3];
    rx[i] = a.y * b.w - a.w * b.y;
    ry[i] = a.w * b.x - a.x * b.w;
    rz[i] = a.x * b.y - a.y * b.x;
  }
  const float det = (g.clip[0].x * rx[0] + g.clip[0].y * ry[0]) + g.clip[0].w * rz[0];
  if (!(det != 0.0f) || !std::isfinite(det)) return g;
  // y-down framebuffer: det > 0 is clockwise on screen; the default front face is counter-clockwise, back faces are culled
  const bool front = p.frontCW ? det > 0.0f : det < 0.0f;
  if (!front) return g;
  const float s = det > 0.0f ? 1.0f : -1.0f;
  const float sx = 2.0f / (float)W, sy = 2.0f / (float)H;
  for (int i = 0; i < 3; ++i) { // lambda_i in pixel units: x_ndc = x_p * 2/W - 1
    g.A[i] = s * (rx[i] * sx);
    g.B[i] = s * (ry[i] * sy);
    g.C[i] = s * ((rz[i] - rx[i]) - ry[i]);
  }
  g.rdet = 1.0f / std::fabs(det);
  g.ok = true;
  return g;
}

bool edgeInside(float e, float A, float B) { return e > 0.0f || (e == 0.0f && (A > 0.0f || (A == 0.0f && B > 0.0f))); }

struct Frag { float e[3], S, z; };
// coverage + depth clip of the pixel centre (px + 0.5, py + 0.5)
bool fragment(const Tri& g, int px, int py, Frag& f) {
  const float x = (float)px + 0.5f, y = (float)py + 0.5f;
  for (int i = 0; i < 3; ++i) f.e[i] = std::fmaf(g.A[i], x, std::fmaf(g.B[i], y, g.C[i]));
  if (!(edgeInside(f.e[0], g.A[0], g.B[0]) && edgeInside(f.e[1], g.A[1], g.B[1]) && edgeInside(f.e[2], g.A[2], g.B[2]))) return false;
  const float zn = std::fmaf(f.e[0], g.clip[0].z, std::fmaf(f.e[1], g.clip[1].z, f.e[2] * g.clip[2].z));
  // normalised by the same edge values (sum e_i w_i = |det| in exact arithmetic), so that their rounding errors cancel
  const float wn = std::fmaf(f.e[0], g.clip[0].w, std::fmaf(f.e[1], g.clip[1].w, f.e[2] * g.clip[2].w));
  f.z = zn / wn;
  if (!(f.z >= 0.0f && f.z <= 1.0f)) return false;
  f.S = (f.e[0] + f.e[1]) + f.e[2];
  return f.S > 0.0f;
}

struct UvSample { float uv[2], ddx[2], ddy[2]; };
UvSample interpUv(const Tri& g, const Frag& f, const OraclePrim& p, int set) {
  float u[3][2];
  for (int k = 0; k < 3; ++k) {
    const float* uv = p.verts + (size_t)g.vi[k] * kVertexFloats + 12 + 2 * (set & 3);
    u[k][0] = uv[0]; u[k][1] = uv[1];
  }
  const float rs = 1.0f / f.S;
  UvSample r;
  const float sA = g.A[0] + g.A[1] + g.A[2], sB = g.B[0] + g.B[1] + g.B[2];
  for (int c = 0; c < 2; ++c) {
    r.uv[c] = (f.e[0] * u[0][c] + f.e[1] * u[1][c] + f.e[2] * u[2][c]) * rs;
    r.ddx[c] = ((g.A[0] * u[0][c] + g.A[1] * u[1][c] + g.A[2] * u[2][c]) - r.uv[c] * sA) * rs;
    r.ddy[c] = ((g.B[0] * u[0][c] + g.B[1] * u[1][c] + g.B[2] * u[2][c]) - r.uv[c] * sB) * rs;
  }
  return r;
}
float fragmentAlpha(const Tri& g, const Frag& f, const OraclePrim& p) {
  if (!p.base.texels) return p.baseColorFactor[3];
  const UvSample u = interpUv(g, f, p, p.baseUv);
  return sampleTexture(p.base, V4{1, 1, 1, 1}, u.uv, u.ddx, u.ddy).w * p.baseColorFactor[3];
}

// what the rasteriser hands Gltf.frag: the interpolated TBN columns and the four uv sets with their derivatives
void fragmentInputs(const Tri& g, const Frag& f, const OraclePrim& p, float tbn[9], float uvs[4][2], float ddx[4][2], float ddy[4][2]) {
  const float b[3] = {f.e[0] / f.S, f.e[1] / f.S, f.e[2] / f.S};
  for (int c = 0; c < 9; ++c) tbn[c] = 0.0f;
  for (int q = 0; q < 3; ++q)
    for (int c = 0; c < 9; ++c) tbn[c] += b[q] * g.tbn[q][c];
  for (int set = 0; set < 4; ++set) {
    const UvSample u = interpUv(g, f, p, set);
    for (int c = 0; c < 2; ++c) { uvs[set][c] = u.uv[c]; ddx[set][c] = u.ddx[c]; ddy[set][c] = u.ddy[c]; }
  }
}
bool hookedFragment(const Tri& g, const Frag& f, const OraclePrim& p, float n[4], float a[4], float m[4]) {
  float tbn[9], uvs[4][2], ddx[4][2], ddy[4][2];
  fragmentInputs(g, f, p, tbn, uvs, ddx, ddy);
  return g_hooks->gbufferFragment(&p, tbn, uvs, ddx, ddy, n, a, m) != 0;
}

uint8_t unorm8(float v) {
  v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
  return (uint8_t)std::nearbyintf(v * 255.0f); // round to nearest even (default rounding mode)
}

} // namespace

extern "C" {

void oracle_set_stage_hooks(const OracleStageHooks* h) { g_hooks = h; }
// the texture unit of the two passes, for the hooks (same filter, same level of detail from the derivatives)
void oracle_sample_texture(const OracleTex* t, const float dflt[4], const float uv[2], const float ddx[2], const float ddy[2], float out[4]) {
  const V4 r = sampleTexture(*t, V4{dflt[0], dflt[1], dflt[2], dflt[3]}, uv, ddx, ddy);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// Outputs (any may be null): depth W*H floats, position W*H*4 floats, normal W*H*4 floats (before the RGBA16F store),
// albedo / mro W*H*4 bytes, triId W*H uint32 (global triangle ordinal in draw order, 0xffffffff = none).
void oracle_draw_gbuffer(const float* projection, const float* view, const OraclePrim* prims, int nPrims, int W, int H, float* depth,
                         float* position, float* normal, uint8_t* albedo, uint8_t* mro, uint32_t* triId) {
  float pv[16];
  matmul44(projection, view, pv); // Gltf.vert:55: (projection * view) * worldPos
  const size_t n = (size_t)W * H;
  std::vector<float> zbuf(n, 1.0f);
  std::vector<uint32_t> ids(n, 0xffffffffu);
  std::vector<std::vector<uint32_t>> chain(n); // per pixel: the triangles whose fragment passed the depth test, in draw order
  std::vector<uint32_t> offsets(nPrims + 1, 0u);
  for (int i = 0; i < nPrims; ++i) offsets[i + 1] = offsets[i] + prims[i].triCount;
  // in draw order; rows in parallel inside a triangle would be pointless for small triangles, so: serial triangles, LESS test
  for (int pi = 0; pi < nPrims; ++pi) {
    const OraclePrim& p = prims[pi];
    for (uint32_t t = 0; t < p.triCount; ++t) {
      const Tri g = setupTriangle(p, t, pv, nullptr, nullptr, W, H, projection, view);
      if (!g.ok) continue;
      // screen bounds when the whole triangle is in front of the eye (a superset, padded); the whole viewport otherwise
      int x0 = 0, y0 = 0, x1 = W - 1, y1 = H - 1;
      if (g.clip[0].w > 0.0f && g.clip[1].w > 0.0f && g.clip[2].w > 0.0f) {
        float xmin = 3e38f, xmax = -3e38f, ymin = 3e38f, ymax = -3e38f;
        for (int k = 0; k < 3; ++k) {
          const float fx = (g.clip[k].x / g.clip[k].w + 1.0f) * (0.5f * (float)W), fy = (g.clip[k].y / g.clip[k].w + 1.0f) * (0.5f * (float)H);
          xmin = std::fmin(xmin, fx); xmax = std::fmax(xmax, fx); ymin = std::fmin(ymin, fy); ymax = std::fmax(ymax, fy);
        }
        if (std::isfinite(xmin) && std::isfinite(xmax) && std::isfinite(ymin) && std::isfinite(ymax)) {
          x0 = (int)std::fmax(std::floor(xmin) - 2.0f, 0.0f); y0 = (int)std::fmax(std::floor(ymin) - 2.0f, 0.0f);
          x1 = (int)std::fmin(std::ceil(xmax) + 2.0f, (float)(W - 1)); y1 = (int)std::fmin(std::ceil(ymax) + 2.0f, (float)(H - 1));
        }
      }
      for (int py = y0; py <= y1; ++py)
        for (int px = x0; px <= x1; ++px) {
          Frag f;
          if (!fragment(g, px, py, f)) continue;
          if (p.alphaCutoff > 0.0f) { // discard (Gltf.frag:38-40)
            float hn[4], ha[4], hm[4];
            if (g_hooks && g_hooks->gbufferFragment ? !hookedFragment(g, f, p, hn, ha, hm) : fragmentAlpha(g, f, p) < p.alphaCutoff) continue;
          }
          const size_t o = (size_t)py * W + px;
          if (!(f.z < zbuf[o])) continue; // VK_COMPARE_OP_LESS
          zbuf[o] = f.z;
          ids[o] = offsets[pi] + t;
          chain[o].push_back(offsets[pi] + t);
        }
    }
  }
  // Colour attachments are BLENDED, always: every pipeline of the reference enables blending with SRC_ALPHA / ONE_MINUS_SRC_ALPHA
  // on colour and ONE / ZERO on alpha (Src/GraphicsPipeline.cpp:138-154), so each fragment that passes the depth test leaves
  //   rgb = src.rgb * a + dst.rgb * (1 - a),  a = src.a   in the attachment's format (RGBA16F normal, UNORM8 albedo / MRO).
  // The fragments that pass at a pixel are `chain[pixel]`, in draw order with strictly decreasing depth; an opaque one (a == 1)
  // erases what was below it, so only the tail of the chain from the last opaque fragment on needs shading.
#pragma omp parallel for schedule(dynamic, 4)
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      const size_t o = (size_t)py * W + px;
      float outN[4] = {0, 0, 0, 0}, outP[4] = {0, 0, 0, 0}; // the clears (Src/DeferredRendering.cpp:101-104)
      uint8_t outA[4] = {0, 0, 0, 0}, outM[4] = {0, 0, 0, 0};
      const std::vector<uint32_t>& ch = chain[o];
      struct Layer { float n[4], a[4], m[4], pos[3]; };
      std::vector<Layer> layers; // top first
      for (size_t k = ch.size(); k-- > 0;) {
        int pi = 0;
        while (offsets[pi + 1] <= ch[k]) ++pi;
        const OraclePrim& p = prims[pi];
        const Tri g = setupTriangle(p, ch[k] - offsets[pi], pv, nullptr, nullptr, W, H, projection, view);
        Frag f;
        fragment(g, px, py, f);
        const float b[3] = {f.e[0] / f.S, f.e[1] / f.S, f.e[2] / f.S};
        float T[3] = {0, 0, 0}, Bt[3] = {0, 0, 0}, N[3] = {0, 0, 0};
        for (int q = 0; q < 3; ++q) // Gltf.vert:59: vertTbn = mat3(model) * tbn, interpolated
          for (int c = 0; c < 3; ++c) { T[c] += b[q] * g.tbn[q][c]; Bt[c] += b[q] * g.tbn[q][3 + c]; N[c] += b[q] * g.tbn[q][6 + c]; }
        Layer L;
        L.pos[0] = b[0] * g.world[0].x + b[1] * g.world[1].x + b[2] * g.world[2].x;
        L.pos[1] = b[0] * g.world[0].y + b[1] * g.world[1].y + b[2] * g.world[2].y;
        L.pos[2] = b[0] * g.world[0].z + b[1] * g.world[1].z + b[2] * g.world[2].z;
        if (g_hooks && g_hooks->gbufferFragment) { // the same fragment through Gltf.frag's own text
          hookedFragment(g, f, p, L.n, L.a, L.m);
          layers.push_back(L);
          if (L.a[3] == 1.0f) break;
          continue;
        }
        const UvSample ub = interpUv(g, f, p, p.baseUv), um = interpUv(g, f, p, p.mrUv);
        V4 base = sampleTexture(p.base, V4{1, 1, 1, 1}, ub.uv, ub.ddx, ub.ddy);
        base.x *= p.baseColorFactor[0]; base.y *= p.baseColorFactor[1]; base.z *= p.baseColorFactor[2]; base.w *= p.baseColorFactor[3];
        const V4 nm = sampleTexture(p.normal, V4{128.0f / 255.0f, 128.0f / 255.0f, 1.0f, 1.0f}, ub.uv, ub.ddx, ub.ddy); // normal map uses the BASE uv set (InstanceData.glsl:41)
        const float ts[3] = {(2.0f * nm.x - 1.0f) * p.normalScale, (2.0f * nm.y - 1.0f) * p.normalScale, 2.0f * nm.z - 1.0f};
        float nn[3];
        for (int c = 0; c < 3; ++c) nn[c] = ts[0] * T[c] + ts[1] * Bt[c] + ts[2] * N[c];
        const float nl = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
        const V4 mr = sampleTexture(p.mr, V4{1, 1, 1, 1}, um.uv, um.ddx, um.ddy);
        const float metallic = mr.z * p.metallicFactor, roughness = mr.y * p.roughnessFactor; // .bg
        // Gltf.frag:44-49: the three colour outputs of this fragment
        L.n[0] = nn[0] / nl; L.n[1] = nn[1] / nl; L.n[2] = nn[2] / nl; L.n[3] = base.w;
        L.a[0] = base.x; L.a[1] = base.y; L.a[2] = base.z; L.a[3] = base.w;
        L.m[0] = metallic; L.m[1] = roughness; L.m[2] = 0.0f; L.m[3] = base.w; // desc.ao = 0.0 (InstanceData.glsl:60)
        layers.push_back(L);
        if (base.w == 1.0f) break; // opaque: nothing below it survives the blend
      }
      if (!layers.empty()) {
        float dN[4] = {0, 0, 0, 0}, dA[4] = {0, 0, 0, 0}, dM[4] = {0, 0, 0, 0}; // attachment contents, as read back from their formats
        for (size_t k = layers.size(); k-- > 0;) { // bottom-up, each step stored in the attachment's format
          const Layer& L = layers[k];
          const float a = L.n[3], ia = 1.0f - a;
          for (int c = 0; c < 3; ++c) {
            dN[c] = (float)(_Float16)(L.n[c] * a + dN[c] * ia);
            dA[c] = (float)unorm8(L.a[c] * a + dA[c] * ia) / 255.0f;
            dM[c] = (float)unorm8(L.m[c] * a + dM[c] * ia) / 255.0f;
          }
          dN[3] = (float)(_Float16)a;
          dA[3] = dM[3] = (float)unorm8(a) / 255.0f;
        }
        for (int c = 0; c < 4; ++c) {
          outN[c] = dN[c];
          outA[c] = unorm8(dA[c]);
          outM[c] = unorm8(dM[c]);
        }
        outP[0] = layers[0].pos[0]; outP[1] = layers[0].pos[1]; outP[2] = layers[0].pos[2]; outP[3] = 1.0f; // our extension: the nearest fragment
      }
      if (depth) depth[o] = zbuf[o];
      if (triId) triId[o] = ids[o];
      if (position) std::memcpy(position + 4 * o, outP, sizeof outP);
      if (normal) std::memcpy(normal + 4 * o, outN, sizeof outN);
      if (albedo) std::memcpy(albedo + 4 * o, outA, sizeof outA);
      if (mro) std::memcpy(mro + 4 * o, outM, sizeof outM);
    }
}

// out: nLights * 6 layers of res x res floats. lights: 8 floats per light (PointLight, position first). views: 6 x 16 floats.
void oracle_draw_shadow_cubes(const float* lights, int nLights, const float* projection, const float* views, const OraclePrim* prims, int nPrims,
                              int res, float* out) {
  const size_t face = (size_t)res * res;
  for (size_t i = 0; i < face * 6 * (size_t)nLights; ++i) out[i] = 1.0f;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int l = 0; l < nLights; ++l)
    for (int v = 0; v < 6; ++v) {
      float* layer = out + ((size_t)l * 6 + v) * face;
      const float* off = lights + 8 * l;
      for (int pi = 0; pi < nPrims; ++pi) {
        const OraclePrim& p = prims[pi];
        for (uint32_t t = 0; t < p.triCount; ++t) {
          const Tri g = setupTriangle(p, t, views + 16 * v, projection, off, res, res);
          if (!g.ok) continue;
          if (g.clip[0].z < 0.0f && g.clip[1].z < 0.0f && g.clip[2].z < 0.0f) continue; // wholly behind the near plane
          for (int py = 0; py < res; ++py)
            for (int px = 0; px < res; ++px) {
              Frag f;
              if (!fragment(g, px, py, f)) continue;
              const bool hooked = g_hooks && g_hooks->shadowFragment;
              if (!hooked && p.alphaCutoff > 0.0f && fragmentAlpha(g, f, p) < p.alphaCutoff) continue; // ShadowMapBindless.frag:33-38
              const float b0 = f.e[0] / f.S, b1 = f.e[1] / f.S, b2 = f.e[2] / f.S;
              const float cx = (b0 * g.cs[0][0] + b1 * g.cs[1][0]) + b2 * g.cs[2][0];
              const float cy = (b0 * g.cs[0][1] + b1 * g.cs[1][1]) + b2 * g.cs[2][1];
              const float cz = (b0 * g.cs[0][2] + b1 * g.cs[1][2]) + b2 * g.cs[2][2];
              float d = std::sqrt((cx * cx + cy * cy) + cz * cz) / 1000.0f; // gl_FragDepth = length(worldPosCS) / zFar
              if (hooked) {
                const float cs3[3] = {cx, cy, cz};
                const UvSample u = interpUv(g, f, p, p.baseUv);
                if (!g_hooks->shadowFragment(&p, cs3, u.uv, u.ddx, u.ddy, &d)) continue;
              }
              float& dst = layer[(size_t)py * res + px];
              if (d >= 0.0f && d < dst) dst = d; // LESS
            }
        }
      }
    }
}

} // extern "C"
