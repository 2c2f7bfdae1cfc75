// Rasterising producers of the deferred path's inputs, sm_100a:
//
//   G-buffer pass      <- Shaders/Gltf/Gltf.vert + Gltf.frag + InstanceData/InstanceData.glsl fetchMaterial
//                         (SceneToGBufferPass, Src/DeferredRendering.cpp:268-330; Primitive::buildPipeline, Src/Primitive.cpp:23-48)
//   omni shadow cubes  <- Shaders/ShadowMapBindless.vert + .frag (PointLightCollection::drawShadowMaps, Src/PointLight.cpp:235-282)
//
// There is no ROP here, so the design is a visibility buffer: raster_setup_kernel turns every (triangle, view) into edge functions
// in 2-D homogeneous form (no near-plane clipping of geometry: a triangle that crosses w = 0 keeps exact edge functions, the
// clipped polygon is only used for its bounding box) and queues one work item per 64 x 64 pixel tile of its bounding box;
// raster_fill_kernel gives each work item to one warp, which tests pixel centres and resolves depth with a 64-bit atomicMin on
// (depth bits << 32 | triangle ordinal): LESS against a 1.0 clear, the first-drawn triangle winning ties exactly as in-order
// rasterisation does. The shadow pass min-reduces gl_FragDepth = |view-space position| / 1000 straight into the cube layers.
// gbuffer_resolve_kernel then shades each pixel once from its winning triangle (perspective-correct attributes, analytic uv
// derivatives for the mip level), so texture work is never spent on hidden fragments.
// The reference alpha-blends its colour attachments in draw order; pixels whose winner is translucent are peeled layer by layer
// (gbuffer_peel_kernel) and composited with the blend equation and the attachment formats' roundings.
//
// Coverage rules (the restatement in oracle/althea_oracle_raster.cpp spells out the same operations): pixel centres; top-left
// tie rule; 0 <= z <= w depth clip; back faces culled (VK_CULL_MODE_BACK_BIT, Include/Althea/GraphicsPipeline.h:171) with the
// primitive's dynamic front face. Every operation that decides coverage or depth is an explicit round-to-nearest intrinsic or an
// explicit fmaf, so the compiler's contraction choices cannot move a pixel: ids and depths match the restatement bit for bit.
// A shared edge evaluates to exactly opposite values in the two triangles that share it (the cross products negate exactly),
// so meshes are watertight. Texture filtering is isotropic trilinear (the reference asks for the device's maximum anisotropy,
// which Vulkan leaves implementation-defined). Skinned primitives are not supported.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "raster_launchers.h"

namespace althea_raster {

#define RDEV __device__ __forceinline__

__constant__ float kSrgbToLinear[256];

RDEV float mulr(float a, float b) { return __fmul_rn(a, b); }
RDEV float addr(float a, float b) { return __fadd_rn(a, b); }
RDEV float subr(float a, float b) { return __fsub_rn(a, b); }

struct F4 { float x, y, z, w; };
struct F3 { float x, y, z; };
struct F2 { float x, y; };

// column-major mat4 * vec4, summed left to right, never contracted
RDEV F4 mulMV(const float* m, float x, float y, float z, float w) {
  F4 r;
  r.x = addr(addr(addr(mulr(m[0], x), mulr(m[4], y)), mulr(m[8], z)), mulr(m[12], w));
  r.y = addr(addr(addr(mulr(m[1], x), mulr(m[5], y)), mulr(m[9], z)), mulr(m[13], w));
  r.z = addr(addr(addr(mulr(m[2], x), mulr(m[6], y)), mulr(m[10], z)), mulr(m[14], w));
  r.w = addr(addr(addr(mulr(m[3], x), mulr(m[7], y)), mulr(m[11], z)), mulr(m[15], w));
  return r;
}
RDEV F3 mulM3V(const float* m, const float* v) { // mat3(m) * v
  F3 r;
  r.x = addr(addr(mulr(m[0], v[0]), mulr(m[4], v[1])), mulr(m[8], v[2]));
  r.y = addr(addr(mulr(m[1], v[0]), mulr(m[5], v[1])), mulr(m[9], v[2]));
  r.z = addr(addr(mulr(m[2], v[0]), mulr(m[6], v[1])), mulr(m[10], v[2]));
  return r;
}

struct TriGeom {
  uint32_t vi[3];
  F4 world[3]; // model * (position, 1)
  F4 clip[3];
  F3 cs[3]; // shadow pass: view-space position / w
};

// Gltf.vert:52-53 / ShadowMapBindless.vert:44: world position of the triangle's vertices (view-independent)
RDEV void triWorld(const RasterPrim& p, uint32_t t, TriGeom& g) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t i = min(__ldg(p.idx + 3u * t + k), p.vertCount - 1u); // an index past the buffer reads its last vertex
    g.vi[k] = i;
    const float* pos = p.verts[i].position;
    const float px = __ldg(pos), py = __ldg(pos + 1), pz = __ldg(pos + 2);
    g.world[k] = mulMV(p.model, px, py, pz, 1.0f);
  }
}
// Gltf.vert:55 (clip = (projection * view) * worldPos, the product formed on the host) and ShadowMapBindless.vert:45-49
RDEV void triClip(const RasterView& v, TriGeom& g) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const F4 w = g.world[k];
    if (v.hasB) {
      const F4 c = mulMV(v.a, subr(w.x, v.off[0]), subr(w.y, v.off[1]), subr(w.z, v.off[2]), w.w);
      g.clip[k] = mulMV(v.b, c.x, c.y, c.z, c.w);
      g.cs[k].x = __fdiv_rn(c.x, c.w);
      g.cs[k].y = __fdiv_rn(c.y, c.w);
      g.cs[k].z = __fdiv_rn(c.z, c.w);
    } else {
      g.clip[k] = mulMV(v.a, w.x, w.y, w.z, w.w);
      g.cs[k].x = g.cs[k].y = g.cs[k].z = 0.0f;
    }
  }
}
RDEV void triTransform(const RasterPrim& p, uint32_t t, const RasterView& v, TriGeom& g) {
  triWorld(p, t, g);
  triClip(v, g);
}

struct TriEdges {
  float A[3], B[3], C[3];
  float rdet;
};

// Edge functions of the triangle in 2-D homogeneous coordinates (x, y, w): lambda_i = row i of adj(M) . (x_ndc, y_ndc, 1), which
// equals det * b_i / w for the perspective-correct barycentric b_i, so all three carry the sign of det exactly where the pixel
// sees the triangle in front of the eye. Returns false for culled or degenerate triangles.
RDEV bool triEdges(const TriGeom& g, int W, int H, bool frontCW, TriEdges& e) {
  float rx[3], ry[3], rz[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const F4 a = g.clip[(i + 1) % 3], b = g.clip[(i + 2) % 3];
    rx[i] = subr(mulr(a.y, b.w), mulr(a.w, b.y));
    ry[i] = subr(mulr(a.w, b.x), mulr(a.x, b.w));
    rz[i] = subr(mulr(a.x, b.y), mulr(a.y, b.x));
  }
  const float det = addr(addr(mulr(g.clip[0].x, rx[0]), mulr(g.clip[0].y, ry[0])), mulr(g.clip[0].w, rz[0]));
  if (!(det != 0.0f) || !(fabsf(det) < __int_as_float(0x7f800000))) return false; // zero, NaN or inf
  // In a y-down framebuffer det > 0 is a clockwise triangle; Vulkan's area sign makes counter-clockwise the front by default.
  const bool front = frontCW ? det > 0.0f : det < 0.0f;
  if (!front) return false;
  const float s = det > 0.0f ? 1.0f : -1.0f;
  const float sx = __fdiv_rn(2.0f, (float)W), sy = __fdiv_rn(2.0f, (float)H);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    e.A[i] = mulr(s, mulr(rx[i], sx));
    e.B[i] = mulr(s, mulr(ry[i], sy));
    e.C[i] = mulr(s, subr(subr(rz[i], rx[i]), ry[i]));
  }
  e.rdet = __fdiv_rn(1.0f, fabsf(det));
  return true;
}

constexpr int kSmallBoxPixels = 16;
RDEV float edgeAt(float A, float B, float C, float x, float y) { return __fmaf_rn(A, x, __fmaf_rn(B, y, C)); }
RDEV bool edgeInside(float e, float A, float B) { return e > 0.0f || (e == 0.0f && (A > 0.0f || (A == 0.0f && B > 0.0f))); } // top-left rule

// ---- textures ---------------------------------------------------------------------------------------------------------
RDEV int wrapIndex(int i, int n, uint32_t mode) {
  if (mode == 1u) return min(max(i, 0), n - 1);
  if (mode == 2u) { // MIRRORED_REPEAT: (n - 1) - mirror((i mod 2n) - n), mirror(k) = k >= 0 ? k : -(1 + k)
    int m = i % (2 * n);
    if (m < 0) m += 2 * n;
    m -= n;
    if (m < 0) m = -(1 + m);
    return (n - 1) - m;
  }
  int m = i % n;
  return m < 0 ? m + n : m;
}
RDEV F4 texel(const RasterTex& t, const uint8_t* level, int w, int x, int y) {
  const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(level) + (size_t)y * w + x);
  F4 r;
  if (t.sampler & 0x100u) { // sRGB decode happens before filtering
    r.x = kSrgbToLinear[raw & 0xffu];
    r.y = kSrgbToLinear[(raw >> 8) & 0xffu];
    r.z = kSrgbToLinear[(raw >> 16) & 0xffu];
  } else {
    r.x = (float)(raw & 0xffu) / 255.0f;
    r.y = (float)((raw >> 8) & 0xffu) / 255.0f;
    r.z = (float)((raw >> 16) & 0xffu) / 255.0f;
  }
  r.w = (float)(raw >> 24) / 255.0f;
  return r;
}
RDEV F4 lerp4(F4 a, F4 b, float t) {
  const float o = 1.0f - t;
  F4 r;
  r.x = a.x * o + b.x * t; r.y = a.y * o + b.y * t; r.z = a.z * o + b.z * t; r.w = a.w * o + b.w * t;
  return r;
}
RDEV F4 sampleLevel(const RasterTex& t, int level, float u, float v, bool nearest) {
  const uint8_t* base = t.texels;
  int w = t.w, h = t.h;
  for (int k = 0; k < level; ++k) {
    base += (size_t)w * h * 4;
    w = max(w >> 1, 1);
    h = max(h >> 1, 1);
  }
  const uint32_t wu = t.sampler & 3u, wv = (t.sampler >> 2) & 3u;
  if (nearest) {
    const int i = wrapIndex((int)floorf(mulr(u, (float)w)), w, wu), j = wrapIndex((int)floorf(mulr(v, (float)h)), h, wv);
    return texel(t, base, w, i, j);
  }
  const float x = subr(mulr(u, (float)w), 0.5f), y = subr(mulr(v, (float)h), 0.5f);
  const float fx0 = floorf(x), fy0 = floorf(y);
  const float fx = x - fx0, fy = y - fy0;
  const int ix = (int)fx0, iy = (int)fy0;
  const int i0 = wrapIndex(ix, w, wu), i1 = wrapIndex(ix + 1, w, wu), j0 = wrapIndex(iy, h, wv), j1 = wrapIndex(iy + 1, h, wv);
  return lerp4(lerp4(texel(t, base, w, i0, j0), texel(t, base, w, i1, j0), fx), lerp4(texel(t, base, w, i0, j1), texel(t, base, w, i1, j1), fx), fy);
}
// texture(): isotropic level of detail from the uv derivatives (Vulkan 16.5.7 with no anisotropy), then the sampler's filters
RDEV F4 sampleTexture(const RasterTex& t, F4 dflt, F2 uv, F2 ddx, F2 ddy) {
  if (!t.texels) return dflt;
  const float ax = ddx.x * (float)t.w, ay = ddx.y * (float)t.h, bx = ddy.x * (float)t.w, by = ddy.y * (float)t.h;
  const float rho2 = fmaxf(ax * ax + ay * ay, bx * bx + by * by);
  float lod = 0.5f * log2f(rho2); // log2(rho); -inf for constant uv
  if (!(lod == lod)) lod = 0.0f;
  const bool magNearest = t.sampler & 0x10u, minNearest = t.sampler & 0x20u;
  const uint32_t mipMode = (t.sampler >> 6) & 3u;
  if (lod <= 0.0f) return sampleLevel(t, 0, uv.x, uv.y, magNearest);
  if (mipMode == 0u || t.mips <= 1) return sampleLevel(t, 0, uv.x, uv.y, minNearest);
  lod = fminf(lod, (float)(t.mips - 1));
  if (mipMode == 1u) { // NEAREST mip: level = ceil(lod + 0.5) - 1
    int l = (int)ceilf(lod + 0.5f) - 1;
    l = min(max(l, 0), t.mips - 1);
    return sampleLevel(t, l, uv.x, uv.y, minNearest);
  }
  const float l0f = floorf(lod);
  const int l0 = (int)l0f;
  const float f = lod - l0f;
  const F4 s0 = sampleLevel(t, l0, uv.x, uv.y, minNearest);
  if (f == 0.0f || l0 + 1 >= t.mips) return s0;
  return lerp4(s0, sampleLevel(t, l0 + 1, uv.x, uv.y, minNearest), f);
}

// perspective-correct interpolation of a uv set and its screen-space derivatives from the edge functions:
// uv = N / S with N = sum e_i uv_i, S = sum e_i, so d uv / dx = (sum A_i uv_i - uv sum A_i) / S
struct UvSample { F2 uv, ddx, ddy; };
RDEV UvSample interpUv(const float e[3], const float A[3], const float B[3], float S, const F2 uvs[3]) {
  const float rs = 1.0f / S;
  UvSample r;
  r.uv.x = (e[0] * uvs[0].x + e[1] * uvs[1].x + e[2] * uvs[2].x) * rs;
  r.uv.y = (e[0] * uvs[0].y + e[1] * uvs[1].y + e[2] * uvs[2].y) * rs;
  const float sA = A[0] + A[1] + A[2], sB = B[0] + B[1] + B[2];
  r.ddx.x = ((A[0] * uvs[0].x + A[1] * uvs[1].x + A[2] * uvs[2].x) - r.uv.x * sA) * rs;
  r.ddx.y = ((A[0] * uvs[0].y + A[1] * uvs[1].y + A[2] * uvs[2].y) - r.uv.y * sA) * rs;
  r.ddy.x = ((B[0] * uvs[0].x + B[1] * uvs[1].x + B[2] * uvs[2].x) - r.uv.x * sB) * rs;
  r.ddy.y = ((B[0] * uvs[0].y + B[1] * uvs[1].y + B[2] * uvs[2].y) - r.uv.y * sB) * rs;
  return r;
}
RDEV F2 vertexUv(const RasterPrim& p, uint32_t vi, int set) {
  const float* uv = p.verts[vi].uvs[set & 3];
  F2 r;
  r.x = __ldg(uv);
  r.y = __ldg(uv + 1);
  return r;
}
// base colour alpha of the fragment (InstanceData.glsl:37-39,66; ShadowMapBindless.frag:33-38)
RDEV float fragmentAlpha(const RasterPrim& p, const uint32_t vi[3], const float e[3], const float A[3], const float B[3], float S) {
  if (!p.mat.base.texels) return p.mat.baseColorFactor[3];
  F2 uvs[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) uvs[k] = vertexUv(p, vi[k], p.mat.baseUv);
  const UvSample u = interpUv(e, A, B, S, uvs);
  F4 one;
  one.x = one.y = one.z = one.w = 1.0f;
  return sampleTexture(p.mat.base, one, u.uv, u.ddx, u.ddy).w * p.mat.baseColorFactor[3];
}

// ---- setup ------------------------------------------------------------------------------------------------------------
RDEV int primOfTriangle(const RasterJob& J, uint32_t tri) { // binary search over the draw-ordered triangle offsets
  int lo = 0, hi = J.nPrims - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (J.prims[mid].triOffset <= tri) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// One pixel of one (triangle, view): coverage, depth clip, optional alpha discard, depth resolve. Shared by the warp-per-tile
// fill and by the setup thread that rasterises small triangles itself.
RDEV void rasterPixel(const RasterJob& J, const RasterPrim& p, const float A[3], const float B[3], const float C[3], const float Z[3], const float Wc[3],
                      const float* attr, uint32_t tri, uint32_t view, bool alphaTest, const uint32_t vi[3], int px, int py) {
  const float x = (float)px + 0.5f, y = (float)py + 0.5f;
  float e[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) e[i] = edgeAt(A[i], B[i], C[i], x, y);
  if (!(edgeInside(e[0], A[0], B[0]) && edgeInside(e[1], A[1], B[1]) && edgeInside(e[2], A[2], B[2]))) return;
  if (J.bound && !(tri + 1u < J.bound[(size_t)py * J.W + px])) return; // peel pass: only what was drawn before the layer found last (bound = its ordinal + 1; 0 = closed pixel)
  const float zn = __fmaf_rn(e[0], Z[0], __fmaf_rn(e[1], Z[1], mulr(e[2], Z[2])));
  // normalised by the SAME edge values (sum e_i W_i = |det| in exact arithmetic): their rounding errors, which grow as the
  // triangle shrinks, then cancel between numerator and denominator instead of landing on the depth
  const float wn = __fmaf_rn(e[0], Wc[0], __fmaf_rn(e[1], Wc[1], mulr(e[2], Wc[2])));
  const float z = __fdiv_rn(zn, wn);
  if (!(z >= 0.0f && z <= 1.0f)) return; // depth clip 0 <= z_c <= w_c
  const float S = addr(addr(e[0], e[1]), e[2]);
  if (!(S > 0.0f)) return;
  if (alphaTest && fragmentAlpha(p, vi, e, A, B, S) < p.mat.alphaCutoff) return; // discard
  if (J.mode == RASTER_MODE_SHADOW) {
    // gl_FragDepth = length(worldPosCS) / zFar with worldPosCS interpolated perspective-correctly (ShadowMapBindless.frag:28,41)
    const float b0 = __fdiv_rn(e[0], S), b1 = __fdiv_rn(e[1], S), b2 = __fdiv_rn(e[2], S);
    const float cx = addr(addr(mulr(b0, attr[0]), mulr(b1, attr[3])), mulr(b2, attr[6]));
    const float cy = addr(addr(mulr(b0, attr[1]), mulr(b1, attr[4])), mulr(b2, attr[7]));
    const float cz = addr(addr(mulr(b0, attr[2]), mulr(b1, attr[5])), mulr(b2, attr[8]));
    const float d = __fdiv_rn(__fsqrt_rn(addr(addr(mulr(cx, cx), mulr(cy, cy)), mulr(cz, cz))), 1000.0f);
    if (!(d >= 0.0f && d < 1.0f)) return; // LESS against the 1.0 clear; depth writes are clamped to [0, 1]
    atomicMin(reinterpret_cast<unsigned int*>(J.shadowBase + view * J.shadowLayerStride + (size_t)py * J.W + px), __float_as_uint(d));
  } else {
    if (!(z < 1.0f)) return; // LESS against the 1.0 clear
    const unsigned long long key = ((unsigned long long)__float_as_uint(z) << 32) | tri;
    atomicMin(J.vis + (size_t)py * J.W + px, key);
  }
}

// one (triangle, view) after the vertex stage: rejection, edge functions, bounding box, record, tile work items
RDEV void setupTriangleView(const RasterJob& J, const RasterPrim& p, int pi, uint32_t tri, uint32_t view, const TriGeom& g) {
  // trivial rejection against the clip volume -w <= x, y <= w, 0 <= z <= w
  bool allOut[6] = {true, true, true, true, true, true};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const F4 c = g.clip[k];
    allOut[0] = allOut[0] && c.x < -c.w; allOut[1] = allOut[1] && c.x > c.w;
    allOut[2] = allOut[2] && c.y < -c.w; allOut[3] = allOut[3] && c.y > c.w;
    allOut[4] = allOut[4] && c.z < 0.0f; allOut[5] = allOut[5] && c.z > c.w;
  }
  if (allOut[0] || allOut[1] || allOut[2] || allOut[3] || allOut[4] || allOut[5]) return;
  TriEdges e;
  if (!triEdges(g, J.W, J.H, p.frontCW != 0u, e)) return;
  // bounding box of the part in front of the near plane (conservative: it only limits which pixels are tested)
  float xmin = 3.0e38f, xmax = -3.0e38f, ymin = 3.0e38f, ymax = -3.0e38f;
  bool whole = false;
  auto addPoint = [&](float x, float y, float w) {
    if (!(w > 0.0f)) { whole = true; return; }
    const float fx = (x / w + 1.0f) * (0.5f * (float)J.W), fy = (y / w + 1.0f) * (0.5f * (float)J.H);
    if (!(fx == fx) || !(fy == fy)) { whole = true; return; }
    xmin = fminf(xmin, fx); xmax = fmaxf(xmax, fx); ymin = fminf(ymin, fy); ymax = fmaxf(ymax, fy);
  };
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const F4 a = g.clip[k], b = g.clip[(k + 1) % 3];
    const bool ain = a.z >= 0.0f, bin = b.z >= 0.0f;
    if (ain) addPoint(a.x, a.y, a.w);
    if (ain != bin) {
      const float t = a.z / (a.z - b.z);
      addPoint(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.w + t * (b.w - a.w));
    }
  }
  int x0, y0, x1, y1;
  if (whole) { x0 = 0; y0 = 0; x1 = J.W - 1; y1 = J.H - 1; }
  else {
    x0 = (int)fmaxf(floorf(xmin) - 1.0f, 0.0f);
    y0 = (int)fmaxf(floorf(ymin) - 1.0f, 0.0f);
    x1 = (int)fminf(ceilf(xmax) + 1.0f, (float)(J.W - 1));
    y1 = (int)fminf(ceilf(ymax) + 1.0f, (float)(J.H - 1));
  }
  if (x1 < x0 || y1 < y0) return;
  if ((x1 - x0 + 1) * (y1 - y0 + 1) <= kSmallBoxPixels) {
    // small triangles (most of a dense mesh, nearly all of it in a 256^2 shadow face) are rasterised right here: no record,
    // no work item, no warp spent on a handful of pixels. Same per-pixel code, and the depth resolve is order-independent.
    float Z[3], Wc[3], attr[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) { Z[k] = g.clip[k].z; Wc[k] = g.clip[k].w; attr[3 * k] = g.cs[k].x; attr[3 * k + 1] = g.cs[k].y; attr[3 * k + 2] = g.cs[k].z; }
    const bool alphaTest = !p.opaque;
    for (int py = y0; py <= y1; ++py)
      for (int px = x0; px <= x1; ++px) rasterPixel(J, p, e.A, e.B, e.C, Z, Wc, attr, tri, view, alphaTest, g.vi, px, py);
    return;
  }
  const uint32_t ri = atomicAdd(&J.counters[0], 1u);
  if (ri >= J.recCap) { J.counters[2] = 1u; atomicMax(&J.counters[4], ri + 1u); return; } // sticky demand: the shim redoes the call with a larger list
  RasterRecord r;
#pragma unroll
  for (int i = 0; i < 3; ++i) { r.A[i] = e.A[i]; r.B[i] = e.B[i]; r.C[i] = e.C[i]; r.Z[i] = g.clip[i].z; r.Wc[i] = g.clip[i].w; }
  r.rdet = e.rdet;
  r.tri = tri;
  r.view = view;
  r.prim = (uint32_t)pi;
  r.bbox[0] = (uint16_t)x0; r.bbox[1] = (uint16_t)y0; r.bbox[2] = (uint16_t)x1; r.bbox[3] = (uint16_t)y1;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.attr[3 * k] = g.cs[k].x; r.attr[3 * k + 1] = g.cs[k].y; r.attr[3 * k + 2] = g.cs[k].z; }
  J.recs[ri] = r;
  const int tx0 = x0 / J.tile, tx1 = x1 / J.tile, ty0 = y0 / J.tile, ty1 = y1 / J.tile;
  const uint32_t n = (uint32_t)((tx1 - tx0 + 1) * (ty1 - ty0 + 1));
  const uint32_t w0 = atomicAdd(&J.counters[1], n);
  if (w0 + n > J.workCap) { J.counters[2] = 1u; atomicMax(&J.counters[3], w0 + n); return; } // sticky: the shim redoes the call with a larger list
  uint32_t k = 0;
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) J.work[w0 + k++] = make_uint2(ri, (uint32_t)tx | ((uint32_t)ty << 16));
}

// Cheap, conservative pre-test for one cube face (shadow pass): the face transform evaluated with plain contracted arithmetic on
// the light-relative positions; a triangle is skipped only when all three vertices are outside the same clip plane by a margin
// a thousand times the difference between this arithmetic and the exact one, so it never changes what is drawn. A triangle
// typically survives for one or two of the six faces, and only those pay for the exact transforms and the edge setup.
RDEV bool faceCannotSee(const RasterView& v, const TriGeom& g) {
  bool out[6] = {true, true, true, true, true, true};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float dx = g.world[k].x - v.off[0], dy = g.world[k].y - v.off[1], dz = g.world[k].z - v.off[2], dw = g.world[k].w;
    const float cx = v.a[0] * dx + v.a[4] * dy + v.a[8] * dz + v.a[12] * dw, cy = v.a[1] * dx + v.a[5] * dy + v.a[9] * dz + v.a[13] * dw;
    const float cz = v.a[2] * dx + v.a[6] * dy + v.a[10] * dz + v.a[14] * dw, cw = v.a[3] * dx + v.a[7] * dy + v.a[11] * dz + v.a[15] * dw;
    const float X = v.b[0] * cx + v.b[4] * cy + v.b[8] * cz + v.b[12] * cw, Y = v.b[1] * cx + v.b[5] * cy + v.b[9] * cz + v.b[13] * cw;
    const float Z = v.b[2] * cx + v.b[6] * cy + v.b[10] * cz + v.b[14] * cw, Wc = v.b[3] * cx + v.b[7] * cy + v.b[11] * cz + v.b[15] * cw;
    const float m = 1e-3f * (fabsf(X) + fabsf(Y) + fabsf(Z) + fabsf(Wc)) + 1e-6f;
    out[0] = out[0] && X < -Wc - m; out[1] = out[1] && X > Wc + m;
    out[2] = out[2] && Y < -Wc - m; out[3] = out[3] && Y > Wc + m;
    out[4] = out[4] && Z < -m;      out[5] = out[5] && Z > Wc + m;
  }
  return out[0] || out[1] || out[2] || out[3] || out[4] || out[5];
}

__global__ void __launch_bounds__(256) raster_setup_kernel(const __grid_constant__ RasterJob J) {
  const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x; // one thread per triangle; its views in a loop
  if (tri >= J.triTotal) return;
  const int pi = primOfTriangle(J, tri);
  const RasterPrim& p = J.prims[pi];
  TriGeom g;
  triWorld(p, tri - p.triOffset, g);
  if (J.cubeFaces) {
    // Shadow cubes: views come six per light (view = 6 light + face) and the faces' frusta are the six 90-degree pyramids around
    // the light's axes (J.faceOfAxis: which face looks along +X, -X, +Y, -Y, +Z, -Z; the shim only sets cubeFaces when the face
    // cameras ARE axis-aligned). A triangle whose three vertices lie strictly inside ONE pyramid, by a margin far above the float
    // noise of the face matrices, lies inside it entirely (the pyramid is convex) and outside the other five: only that face is
    // set up. Everything else (triangles across pyramid boundaries, or next to the light) takes the six conservative tests.
    for (int light = 0; light < J.nViews / 6; ++light) {
      const RasterView& v0 = J.views[6 * light];
      int only = -1;
      {
        int axis = -2;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float dx = g.world[k].x - v0.off[0], dy = g.world[k].y - v0.off[1], dz = g.world[k].z - v0.off[2];
          const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
          int a;
          float major, m1, m2;
          if (ax >= ay && ax >= az) { a = dx >= 0.0f ? 0 : 1; major = ax; m1 = ay; m2 = az; }
          else if (ay >= az) { a = dy >= 0.0f ? 2 : 3; major = ay; m1 = ax; m2 = az; }
          else { a = dz >= 0.0f ? 4 : 5; major = az; m1 = ax; m2 = ay; }
          const bool strict = fmaxf(m1, m2) <= major * 0.998f && major > 1e-6f && g.world[k].w == 1.0f;
          if (!strict) a = -1;
          axis = (axis == -2) ? a : (axis == a ? axis : -1);
        }
        if (axis >= 0) only = J.faceOfAxis[axis];
      }
      for (int f = 0; f < 6; ++f) {
        if (only >= 0 && f != only) continue;
        const RasterView& v = J.views[6 * light + f];
        if (faceCannotSee(v, g)) continue;
        triClip(v, g);
        setupTriangleView(J, p, pi, tri, (uint32_t)(6 * light + f), g);
      }
    }
    return;
  }
  for (int view = 0; view < J.nViews; ++view) {
    const RasterView& v = J.views[view];
    if (J.nViews > 1 && faceCannotSee(v, g)) continue;
    triClip(v, g);
    setupTriangleView(J, p, pi, tri, (uint32_t)view, g);
  }
}

// ---- fill -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raster_fill_kernel(const __grid_constant__ RasterJob J) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
  if (J.counters[2]) return; // the work list overflowed: nothing in it can be trusted (the shim retries)
  const uint32_t nWork = min(J.counters[1], J.workCap);
  for (uint32_t wi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wi < nWork; wi += warpsTotal) {
    const uint2 item = J.work[wi];
    const RasterRecord& rr = J.recs[item.x];
    // the record, once per warp (two 64-byte halves, broadcast loads)
    float A[3], B[3], C[3], Z[3], Wc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { A[i] = rr.A[i]; B[i] = rr.B[i]; C[i] = rr.C[i]; Z[i] = rr.Z[i]; Wc[i] = rr.Wc[i]; }
    const uint32_t tri = rr.tri, view = rr.view;
    const int tx = (int)(item.y & 0xffffu) * J.tile, ty = (int)(item.y >> 16) * J.tile;
    const int x0 = max((int)rr.bbox[0], tx), y0 = max((int)rr.bbox[1], ty);
    const int x1 = min((int)rr.bbox[2], tx + J.tile - 1), y1 = min((int)rr.bbox[3], ty + J.tile - 1);
    // whole-tile rejection: an edge function that is negative at the corner where it is largest is negative everywhere
    bool reject = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float cx = A[i] > 0.0f ? (float)x1 + 0.5f : (float)x0 + 0.5f, cy = B[i] > 0.0f ? (float)y1 + 0.5f : (float)y0 + 0.5f;
      // exact: each fmaf is monotone in the coordinate it multiplies (rounding is monotone), so this corner bounds every pixel
      reject = reject || edgeAt(A[i], B[i], C[i], cx, cy) < 0.0f;
    }
    if (reject) continue;
    const RasterPrim& p = J.prims[rr.prim];
    const bool alphaTest = !p.opaque;
    uint32_t vi[3] = {0u, 0u, 0u};
    if (alphaTest) {
      const uint32_t t = tri - p.triOffset;
#pragma unroll
      for (int k = 0; k < 3; ++k) vi[k] = min(__ldg(p.idx + 3u * t + k), p.vertCount - 1u);
    }
    // 8 x 4 pixel blocks, one pixel per lane
    const int bw = (x1 - x0 + 8) >> 3, bh = (y1 - y0 + 4) >> 2;
    for (int b = 0; b < bw * bh; ++b) {
      const int px = x0 + (b % bw) * 8 + (int)(lane & 7u), py = y0 + (b / bw) * 4 + (int)(lane >> 3);
      if (px > x1 || py > y1) continue;
      rasterPixel(J, p, A, B, C, Z, Wc, rr.attr, tri, view, alphaTest, vi, px, py);
    }
  }
}

__global__ void __launch_bounds__(256) raster_clear_kernel(unsigned long long* vis, float* depth, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (vis) vis[i] = ~0ull;
  if (depth) depth[i] = 1.0f;
}

// ---- G-buffer resolve -------------------------------------------------------------------------------------------------
RDEV uint32_t packUnorm4(float x, float y, float z, float w) { // UNORM8 conversion: clamp, scale, round to nearest even
  auto q = [](float v) { return (uint32_t)__float2int_rn(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f); };
  return q(x) | (q(y) << 8) | (q(z) << 16) | (q(w) << 24);
}

// What one fragment's shader invocation outputs (Gltf.frag:44-49), before the blend and the attachment formats
struct Shaded {
  float n[3], alpha, albedo[3], metallic, roughness;
  float4 position;
};
// vertex stage, interpolation and fetchMaterial for triangle `tri` at pixel (px, py)
RDEV Shaded shadeFragment(const RasterJob& J, uint32_t tri, int px, int py) {
  const RasterPrim& p = J.prims[primOfTriangle(J, tri)];
  TriGeom g;
  triTransform(p, tri - p.triOffset, J.views[0], g);
  TriEdges E;
  triEdges(g, J.W, J.H, p.frontCW != 0u, E); // it passed the depth test once already: same values
  const float x = (float)px + 0.5f, y = (float)py + 0.5f;
  float e[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) e[i] = edgeAt(E.A[i], E.B[i], E.C[i], x, y);
  const float S = addr(addr(e[0], e[1]), e[2]);
  const float b0 = e[0] / S, b1 = e[1] / S, b2 = e[2] / S;
  // Gltf.vert:52-59: world position and mat3(model) * tbn per vertex, interpolated perspective-correctly
  F3 T = {0.0f, 0.0f, 0.0f}, Bt = T, N = T;
  const float bw[3] = {b0, b1, b2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const althea_vertex& vtx = p.verts[g.vi[k]];
    float t3[3], b3[3], n3[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { t3[c] = __ldg(vtx.tangent + c); b3[c] = __ldg(vtx.bitangent + c); n3[c] = __ldg(vtx.normal + c); }
    const F3 tw = mulM3V(p.model, t3), bwv = mulM3V(p.model, b3), nw = mulM3V(p.model, n3);
    T.x += bw[k] * tw.x; T.y += bw[k] * tw.y; T.z += bw[k] * tw.z;
    Bt.x += bw[k] * bwv.x; Bt.y += bw[k] * bwv.y; Bt.z += bw[k] * bwv.z;
    N.x += bw[k] * nw.x; N.y += bw[k] * nw.y; N.z += bw[k] * nw.z;
  }
  Shaded out;
  out.position = make_float4(b0 * g.world[0].x + b1 * g.world[1].x + b2 * g.world[2].x, b0 * g.world[0].y + b1 * g.world[1].y + b2 * g.world[2].y,
                             b0 * g.world[0].z + b1 * g.world[1].z + b2 * g.world[2].z, 1.0f);
  // fetchMaterial, InstanceData.glsl:28-69
  const RasterMaterial& m = p.mat;
  F2 uvb[3], uvm[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { uvb[k] = vertexUv(p, g.vi[k], m.baseUv); uvm[k] = vertexUv(p, g.vi[k], m.mrUv); }
  const UvSample ub = interpUv(e, E.A, E.B, S, uvb), um = interpUv(e, E.A, E.B, S, uvm);
  F4 one; one.x = one.y = one.z = one.w = 1.0f;
  F4 flat; flat.x = flat.y = 128.0f / 255.0f; flat.z = 1.0f; flat.w = 1.0f; // Content/Engine/Textures/normal1x1.png
  F4 base = sampleTexture(m.base, one, ub.uv, ub.ddx, ub.ddy);
  base.x *= m.baseColorFactor[0]; base.y *= m.baseColorFactor[1]; base.z *= m.baseColorFactor[2]; base.w *= m.baseColorFactor[3];
  const F4 nm = sampleTexture(m.normal, flat, ub.uv, ub.ddx, ub.ddy);
  const float tsx = (2.0f * nm.x - 1.0f) * m.normalScale, tsy = (2.0f * nm.y - 1.0f) * m.normalScale, tsz = 2.0f * nm.z - 1.0f;
  float nx = tsx * T.x + tsy * Bt.x + tsz * N.x, ny = tsx * T.y + tsy * Bt.y + tsz * N.y, nz = tsx * T.z + tsy * Bt.z + tsz * N.z;
  const float nl = sqrtf(nx * nx + ny * ny + nz * nz);
  out.n[0] = nx / nl; out.n[1] = ny / nl; out.n[2] = nz / nl;
  const F4 mr = sampleTexture(m.mr, one, um.uv, um.ddx, um.ddy);
  out.metallic = mr.z * m.metallicFactor; // .bg (InstanceData.glsl:52-55)
  out.roughness = mr.y * m.roughnessFactor;
  out.albedo[0] = base.x; out.albedo[1] = base.y; out.albedo[2] = base.z;
  out.alpha = base.w;
  return out;
}

// The reference blends EVERY colour attachment (SRC_ALPHA, ONE_MINUS_SRC_ALPHA on colour; ONE, ZERO on alpha,
// Src/GraphicsPipeline.cpp:138-154): a fragment that passes the depth test leaves rgb = src.rgb a + dst.rgb (1 - a), a = src.a,
// rounded to the attachment's format. One blend step on the three attachments' contents (kept as the floats their formats hold):
struct Attach { float n[4], a[4], m[4]; };
RDEV float roundHalf(float v) { return __half2float(__float2half_rn(v)); }
RDEV float roundUnorm(float v) { return __fdiv_rn((float)__float2int_rn(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f), 255.0f); }
RDEV void blendOver(Attach& d, const Shaded& s) {
  const float a = s.alpha, ia = subr(1.0f, a);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    d.n[c] = roundHalf(addr(mulr(s.n[c], a), mulr(d.n[c], ia)));
    d.a[c] = roundUnorm(addr(mulr(s.albedo[c], a), mulr(d.a[c], ia)));
  }
  d.m[0] = roundUnorm(addr(mulr(s.metallic, a), mulr(d.m[0], ia)));
  d.m[1] = roundUnorm(addr(mulr(s.roughness, a), mulr(d.m[1], ia)));
  d.m[2] = roundUnorm(mulr(d.m[2], ia)); // desc.ao = 0.0 (InstanceData.glsl:60)
  d.n[3] = roundHalf(a);
  d.a[3] = d.m[3] = roundUnorm(a);
}
RDEV void writeAttachments(const RasterJob& J, int px, int py, const Attach& d) {
  if (J.outNormal) {
    __half2 lo = __floats2half2_rn(d.n[0], d.n[1]), hi = __floats2half2_rn(d.n[2], d.n[3]);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&lo);
    v.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<char*>(J.outNormal) + (size_t)py * J.pitchNormal + (size_t)px * 8) = v;
  }
  if (J.outAlbedo) *reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(J.outAlbedo) + (size_t)py * J.pitchAlbedo + (size_t)px * 4) = packUnorm4(d.a[0], d.a[1], d.a[2], d.a[3]);
  if (J.outMro) *reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(J.outMro) + (size_t)py * J.pitchMro + (size_t)px * 4) = packUnorm4(d.m[0], d.m[1], d.m[2], d.m[3]);
}

// Pass 0: every pixel from its depth winner. An opaque winner (alpha == 1, the overwhelmingly common case) erases whatever was
// drawn below it, so the pixel is final. A translucent winner is written blended over the clear colour for now and the pixel is
// left OPEN (openBound = its triangle ordinal + 1): what lies under it is peeled in draw order by the passes below.
__global__ void __launch_bounds__(256, 4) gbuffer_resolve_kernel(const __grid_constant__ RasterJob J) { // 64 registers: four CTAs per SM, as before the blend
  const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (px >= J.W || py >= J.H) return;
  const size_t pix = (size_t)py * J.W + px;
  const unsigned long long key = J.vis[pix];
  float depth = 1.0f;
  float4 position = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  Attach d;
#pragma unroll
  for (int c = 0; c < 4; ++c) d.n[c] = d.a[c] = d.m[c] = 0.0f; // the clears (Src/DeferredRendering.cpp:101-104)
  uint32_t open = 0u;
  if (key != ~0ull) {
    const uint32_t tri = (uint32_t)(key & 0xffffffffu);
    depth = __uint_as_float((uint32_t)(key >> 32));
    const Shaded s = shadeFragment(J, tri, px, py);
    position = s.position;
    blendOver(d, s);
    if (s.alpha != 1.0f) open = tri + 1u; // ordinal + 1: 0 means final
  }
  if (J.outDepth) *reinterpret_cast<float*>(reinterpret_cast<char*>(J.outDepth) + (size_t)py * J.pitchDepth + (size_t)px * 4) = depth;
  if (J.outPosition) *reinterpret_cast<float4*>(reinterpret_cast<char*>(J.outPosition) + (size_t)py * J.pitchPosition + (size_t)px * 16) = position;
  writeAttachments(J, px, py, d);
  J.openBound[pix] = open;
  if (open) atomicAdd(&J.counters[5], 1u);
}

// Peel pass k >= 1 (only when pass 0 left pixels open). J.vis now holds, for every open pixel, the depth winner among the
// triangles drawn BEFORE the layer found last (the fill ran with J.bound = openBound): the fragment that was on top when that
// layer was blended. It is pushed on the pixel's stack; the pixel closes when the new layer is opaque, when nothing lies below
// (the clear colour), or at the last pass; closing composites the stack bottom-up with blendOver and writes the attachments.
__global__ void __launch_bounds__(256) gbuffer_peel_kernel(const __grid_constant__ RasterJob J, int pass, int lastPass) {
  const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (px >= J.W || py >= J.H) return;
  const size_t pix = (size_t)py * J.W + px, npx = (size_t)J.W * J.H;
  const uint32_t bound = J.openBound[pix];
  if (!bound && pass > 0) return; // closed
  // layer stack: layerTri[k * npx + pix] = triangle ordinal of the k-th layer from the top
  if (pass == 0) { // registers the pass-0 winner of an open pixel as layer 0
    if (bound) J.layerTri[pix] = bound - 1u;
    return;
  }
  const unsigned long long key = J.vis[pix];
  bool close = lastPass != 0;
  int count = pass; // layers 0 .. pass-1 are on the stack
  if (key == ~0ull) close = true; // nothing was drawn below: the clear colour
  else {
    const uint32_t tri = (uint32_t)(key & 0xffffffffu);
    J.layerTri[(size_t)pass * npx + pix] = tri;
    count = pass + 1;
    const Shaded s = shadeFragment(J, tri, px, py);
    if (s.alpha == 1.0f) close = true;
    else J.openBound[pix] = tri + 1u;
  }
  if (!close) { atomicAdd(&J.counters[5], 1u); return; }
  Attach d;
#pragma unroll
  for (int c = 0; c < 4; ++c) d.n[c] = d.a[c] = d.m[c] = 0.0f;
  for (int k = count - 1; k >= 0; --k) blendOver(d, shadeFragment(J, J.layerTri[(size_t)k * npx + pix], px, py));
  writeAttachments(J, px, py, d);
  J.openBound[pix] = 0u;
}

// min alpha of a texture's level 0 (cached per image by the shim: a material whose alpha can never fall below its cutoff skips
// the per-fragment alpha test in raster_fill_kernel)
__global__ void __launch_bounds__(256) texture_min_alpha_kernel(const uint32_t* texels, size_t n, unsigned int* out) {
  unsigned int m = 255u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = min(m, __ldg(texels + i) >> 24);
  m = __reduce_min_sync(0xffffffffu, m);
  if ((threadIdx.x & 31u) == 0u) atomicMin(out, m);
}

// ---- launchers --------------------------------------------------------------------------------------------------------
void upload_srgb_table(const float* table256, cudaStream_t s) { cudaMemcpyToSymbolAsync(kSrgbToLinear, table256, 256 * sizeof(float), 0, cudaMemcpyHostToDevice, s); }
void launch_raster_clear(unsigned long long* vis, float* depth, size_t n, cudaStream_t s) {
  raster_clear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(vis, depth, n);
}
void launch_raster_setup(const RasterJob& J, cudaStream_t s) {
  if (J.triTotal) raster_setup_kernel<<<(unsigned)((J.triTotal + 255u) / 256u), 256, 0, s>>>(J);
}
void launch_raster_fill(const RasterJob& J, int sms, cudaStream_t s) { raster_fill_kernel<<<(unsigned)(sms * 8), 256, 0, s>>>(J); }
void launch_gbuffer_resolve(const RasterJob& J, cudaStream_t s) {
  gbuffer_resolve_kernel<<<dim3((unsigned)((J.W + 15) / 16), (unsigned)((J.H + 15) / 16)), 256, 0, s>>>(J);
}
void launch_gbuffer_peel(const RasterJob& J, int pass, int lastPass, cudaStream_t s) {
  gbuffer_peel_kernel<<<dim3((unsigned)((J.W + 15) / 16), (unsigned)((J.H + 15) / 16)), 256, 0, s>>>(J, pass, lastPass);
}
void launch_texture_min_alpha(const uint32_t* texels, size_t n, unsigned int* out, cudaStream_t s) {
  const unsigned blocks = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  texture_min_alpha_kernel<<<blocks ? blocks : 1u, 256, 0, s>>>(texels, n, out);
}

} // namespace althea_raster
