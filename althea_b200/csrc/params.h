// Kernel parameter blocks shared by the host shim (althea_cuda.cu) and the kernels. Plain structs, passed by value
// as __grid_constant__ kernel parameters.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector_types.h>

#include "../../include/althea_cuda.h"

struct ImgView { // one mip level of one layer of a linear image
  const void* ptr;
  int w, h;
  int pitch; // bytes
};
constexpr int kMaxMips = 14;
constexpr int kSsrPlaneStride = 128, kSsrPlaneRows = 128; // records of the SSR plane image (global memory, L1 / L2 resident)
constexpr int kSsaoPlanePad = 16; // blocks of padding around every level of the SSAO plane records
struct ChainView {
  ImgView level[kMaxMips];
  int mips;
};

struct FrameParams {
  althea_global_uniforms g; // byte-for-byte the reference's GlobalUniforms (416 B)
  float projView[16];       // projection * view, computed once on the host in the oracle's op order
  // fast-build SSR march: mat3(inverseView) * (inverseProjection * (2u-1, 2v-1, 2, 1)).xyz = W0 + Wu u + Wv v, and its dot
  // product with the camera z axis inverseView[2].xyz = S[0] + S[1] u + S[2] v (Misc/ReconstructPosition.glsl:9-21)
  float ssrW0[3], ssrWu[3], ssrWv[3], ssrS[3];
  int W, H;
  int y0, y1; // rows this launch shades (the scissor; [0, H) by default)
  ImgView depth, position, normal, albedo, mro; // GBufferResources
  ImgView env, irr, lut;                        // IBLResources
  ChainView pre;                                // prefiltered env, 5 mips
  ChainView refl;                               // reflection buffer mips
  const float* lights;                          // althea_point_light[lightCount]
  ImgView shadow;                               // layer 0 of the cube array; layer l at ptr + l*shadowLayerStride
  size_t shadowLayerStride;
  int shadowRes; // 0 => unshadowed
  ImgView out;   // deferred colour target
  int outIsF32;
  ImgView ao; // R8_UINT occluded-ray counts
  uint32_t flags;
  // SSAO position-quad proxy (engine scratch): (W+1) x (H+1) records of 32 bytes, record (ix+1, iy+1)
  // describes the bilinear footprint whose top-left tap is texel (ix, iy), ix in [-1, W-1], iy in [-1, H-1]
  const void* quads;
  size_t quadPitch; // bytes per record row = (W+1)*32
  const void* quadsOrigin; // &record(1, 1): the footprint whose top-left tap is texel (0, 0)
  int quadRow;             // records per row = W + 1
  float Wf, Hf;            // (float)W, (float)H
  // SSAO ray-depth proxy: position of texel (x, y) modelled as cam + (Dc + Dx x + Dy y) * t with t the eye depth
  // dot(p - cam, fwd); the 16-byte records hold t of a footprint's four texels. quadKind: 0 = 32-byte position records, 1 = these
  float ssaoCam[3], ssaoFwd[3], ssaoDc[3], ssaoDx[3], ssaoDy[3];
  float ssaoGram[6]; // Dc.Dc, Dx.Dx, Dy.Dy, 2 Dc.Dx, 2 Dc.Dy, 2 Dx.Dy
  float ssaoOrigin[3]; // model coordinates of the world origin: Dc o0 + Dx o1 + Dy o2 = -cam (where empty pixels' cleared positions lie)
  int quadKind;
  // SSAO coarse sign test: per level l (blocks of 8 << l texels) one 16-byte record per block:
  // {alpha, beta, gamma, r}: reciprocal eye depth of every texel the block covers lies within r of alpha + beta x + gamma y
  // (texel coordinates); r = +inf where the block cannot decide. Padded by kSsaoPlanePad blocks of r = +inf on every side.
  const float4* ssaoPlanes[3]; // record of block (0, 0) of each level
  int ssaoPlaneRow[3];         // records per padded row
  int ssaoPlaneNx[3], ssaoPlaneNy[3]; // blocks per level (unpadded)
  float ssaoFocalPx;           // focal length in pixels (the larger axis): level choice per tile
  float ssaoDmax1;             // max |D(x, y)|_1 over the image
  float ssaoDmag;              // |Dc|_1 + |Dx|_1 W + |Dy|_1 H: bounds the rounding of dot(D(x, y), v)
  float ssaoCamL1;             // |cam|_1
  unsigned* ssaoTileList;       // [0]: number of 16 x 16 tiles the cull kernel handed over to the march kernel, [1 + k]: tile ids
  unsigned* ssaoPlaneStats;     // [0]: level-2 records inside the image, [1]: those that can decide (zeroed by ssao_quads_kernel)
  float* ssaoRecip;             // W x H: reciprocal eye depth of every position texel (NaN: off the camera model), by ssao_quads_kernel
  // SSAO ray directions (engine scratch, built once per frame size): computeSSAO seeds its hash RNG with the pixel's coordinates
  // (DeferredPass.frag:42) and ray r starts at seed + 3 r, so the tangent-space direction normalize(2 xi.xy - 1, xi.z) of
  // (pixel, ray) is a function of (px + 3 r, py + 3 r) alone, the same in every frame: entry (a, b) holds it for seed (a, b). Null: hashed inline
  const float4* ssaoDirs;
  int ssaoDirRow;               // entries per row
  unsigned long long* gatherCounter; // diagnostics (ALTHEA_CTX_SSAO_COUNT_TAPS): proxy records gathered by the SSAO march; else null
  // SSR sign test: one plane record {alpha, beta, gamma, r} of reciprocal eye depth per block of
  // (1 << ssrPlaneShift)^2 depth texels, the whole frame in kSsrPlaneStride x kSsrPlaneRows records (r = +inf outside). Null: plain march
  const float4* ssrPlanes;
  int ssrPlaneShift;
  // SSR hit list (engine scratch): the march appends one record per hit pixel (12 floats, structure of arrays with stride
  // ssrHitCap), a second kernel shades them packed. ssrHitCount is zeroed by ssr_depth_pad_kernel
  float* ssrHits;
  unsigned* ssrHitCount;
  unsigned ssrHitCap;
  // SSR padded depth (engine scratch): the depth image with a one-texel CLAMP_TO_EDGE border, (W+2) x (H+2) floats
  const float* depthPad;
  const float* depthPadOrigin; // &padded(1, 1), i.e. texel (0, 0)
  int depthPadRow;             // floats per padded row = W + 2
  // fast-build march: the same depths as one 16-byte record {t00, t10, t01, t11} per bilinear footprint (ix, iy), ix in [-1, W-1],
  // iy in [-1, H-1]: a tap is ONE 128-bit load. (W + 1) x (H + 1) records; null: the march reads the padded floats
  const float4* depthQuadOrigin; // record of footprint (0, 0)
  int depthQuadRow;              // records per row = W + 1
};

struct ConvolveParams {
  ImgView src, dst;
  int vertical; // (level & 1): direction = (0,1) else (1,0)  (ReflectionBuffer.cpp:257-258)
  int y0, y1;   // dst rows to write
};

struct MipGenParams {
  ImgView src, dst;
  int channels; // 4 (RGBA32F) only
};

struct IblParams {
  ChainView env; // RGBA32F equirect with full mip chain
  ImgView out[6]; // one mip level of every layer (cube face) of the output (RGBA32F); blockIdx.y picks the face
  int faces;      // 1 (equirect) or 6
  int layout;     // ALTHEA_IBL_LAYOUT_*
  int sequence;  // ALTHEA_IBL_SEQ_*
  int numSamples;
  int thetaSamples, phiSamples;
  float roughness;
  float mip; // irradiance: log2(envW / thetaSamples)
};

struct LutParams {
  ImgView out;
  int outIsF32;
  int samples;
};
