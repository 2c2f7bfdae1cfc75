// Parameter blocks of the rasterising producers (raster_kernels.cu): the G-buffer pass (Shaders/Gltf/Gltf.vert/.frag through
// SceneToGBufferPass, Src/DeferredRendering.cpp:268-330) and the omni shadow-cube pass (Shaders/ShadowMapBindless.vert/.frag
// through PointLightCollection::drawShadowMaps, Src/PointLight.cpp:235-282). Plain structs shared by the shim and the kernels.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/althea_cuda.h"

// sampler word of a texture: wrapU | wrapV << 2 | magNearest << 4 | minNearest << 5 | mipMode << 6 | srgb << 8
// (wrap: 0 REPEAT, 1 CLAMP_TO_EDGE, 2 MIRRORED_REPEAT; mipMode: 0 none, 1 NEAREST, 2 LINEAR), Src/Sampler.cpp:11-88
struct RasterTex {
  const uint8_t* texels; // RGBA8, mip levels tightly packed one after another; null => the material's default constant
  int w, h, mips;
  uint32_t sampler;
};

struct RasterMaterial { // the MaterialConstants fields fetchMaterial reads (Shaders/InstanceData/InstanceData.glsl:28-69)
  float baseColorFactor[4];
  int baseUv, mrUv; // the normal map is sampled with the BASE colour's uv set (InstanceData.glsl:41)
  float normalScale, metallicFactor, roughnessFactor, alphaCutoff;
  RasterTex base, normal, mr;
};

struct RasterPrim { // one Primitive::draw
  const althea_vertex* verts;
  const uint32_t* idx;
  uint32_t triCount, triOffset; // triOffset: triangles of the primitives drawn before this one (draw order breaks depth ties)
  uint32_t vertCount, pad0;     // vertices in the buffer: indices are clamped to it (robustBufferAccess-style, never out of bounds)
  float model[16];
  uint32_t frontCW; // VK_FRONT_FACE_CLOCKWISE (Primitive::getFrontFace)
  uint32_t opaque;  // the shim proved alpha >= alphaCutoff everywhere: no per-fragment alpha test
  RasterMaterial mat;
};

struct RasterView { // clip = hasB ? b * (a * (world - (off, 0))) : a * world
  float a[16], b[16];
  float off[3];
  int hasB;
};

// one surviving (triangle, view): edge functions in pixel units, pre-multiplied by sign(det) so that inside <=> e_i >= 0
struct __align__(16) RasterRecord {
  float A[3], B[3], C[3]; // e_i(x, y) = A_i x + B_i y + C_i at pixel centres (x, y) = (px + 0.5, py + 0.5)
  float Z[3];             // clip-space z of the three vertices: z_ndc = (e0 Z0 + e1 Z1 + e2 Z2) / (e0 W0 + e1 W1 + e2 W2)
  float rdet;             // 1 / |det|
  uint32_t tri;           // global triangle ordinal (draw order)
  uint32_t view;
  uint32_t prim;
  uint16_t bbox[4]; // x0, y0, x1, y1 inclusive
  float attr[9];    // shadow pass: view-space positions of the three vertices (ShadowMapBindless.vert:45-49)
  float Wc[3];      // clip-space w of the three vertices
};
static_assert(sizeof(RasterRecord) == 128, "RasterRecord is two 64-byte halves");

enum { RASTER_MODE_GBUFFER = 0, RASTER_MODE_SHADOW = 1 };


struct RasterJob {
  const RasterPrim* prims;
  int nPrims;
  uint32_t triTotal;
  const RasterView* views;
  int nViews;
  int W, H;
  int mode;
  int cubeFaces;     // shadow pass with axis-aligned face cameras: views are six per light and faceOfAxis is valid
  int faceOfAxis[6]; // face index (0..5) whose camera looks along +X, -X, +Y, -Y, +Z, -Z
  int tile; // pixels per side of a fill work item: 64 for frame-sized targets, 32 for shadow faces (a 256^2 face would otherwise
            // be 16 work items, each a 128-iteration warp: too few warps, too long a tail)
  RasterRecord* recs;
  uint32_t recCap;
  uint2* work; // (record index, tileX | tileY << 16)
  uint32_t workCap;
  uint32_t* counters; // [0] records, [1] work items, [2] overflow flag of this pass, [3] / [4] sticky: largest work-item / record demand that overflowed, [5] pixels left open by the last resolve / peel pass
  // G-buffer: visibility buffer, one 64-bit key per pixel = depth bits << 32 | global triangle ordinal (atomicMin)
  unsigned long long* vis;
  // alpha blending of the G-buffer attachments (raster_kernels.cu, gbuffer_peel_kernel): per pixel, 1 + the triangle ordinal below
  // which the next peel pass looks (0 = the pixel is final); `bound` is that array while a peel pass fills, null in pass 0;
  // layerTri[k * W * H + pixel] = the k-th translucent layer from the top
  uint32_t* openBound;
  const uint32_t* bound;
  uint32_t* layerTri;
  // shadow: depth layers of the current light, layer `view` at shadowBase + view * shadowLayerStride floats
  float* shadowBase;
  size_t shadowLayerStride;
  // G-buffer resolve targets (null => not written)
  float* outDepth;     // R32F
  float4* outPosition; // RGBA32F (world position, 1) / 0 when empty
  uint2* outNormal;    // RGBA16F
  uint32_t* outAlbedo; // RGBA8
  uint32_t* outMro;    // RGBA8
  int pitchDepth, pitchPosition, pitchNormal, pitchAlbedo, pitchMro; // bytes
};
