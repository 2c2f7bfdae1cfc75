/* althea_cuda.h — C ABI of the B200-native engine for Althea's deferred screen-space path.
 *
 * This is the drop-in boundary. The reference (nithinp7/Althea) has no FFI today: the path is a set of C++
 * methods that record Vulkan commands. Each entry point below replaces the BODY of one of those methods
 * (the C++ signatures stay as they are, see INTEGRATION.md for the host-side patch):
 *
 *   althea_cuda_ssr_capture      <- ScreenSpaceReflection::captureReflection      Src/ScreenSpaceReflection.cpp:60-83
 *                                   (Shaders/SSR.vert, Shaders/SSR.frag)
 *   althea_cuda_glossy_convolve  <- ReflectionBuffer::convolveReflectionBuffer     Src/ReflectionBuffer.cpp:163-293
 *                                   (Shaders/SSRGlossyConvolve.comp)
 *   althea_cuda_deferred_shade   <- the app-owned deferred lighting pass that binds IBLResources::bind
 *                                   (Src/ImageBasedLighting.cpp:20-26), GBufferResources::bindTextures
 *                                   (Src/DeferredRendering.cpp:157-162) and ScreenSpaceReflection::bindTexture
 *                                   (Shaders/DeferredPass.vert/.frag, Shaders/SSAO.glsl, Shaders/PBR/PBRMaterial.glsl)
 *   althea_cuda_ibl_precompute   <- ImageBasedLighting.cpp:137-412 precomputeResources (called from
 *                                   ImageBasedLighting::createResources :415-446)
 *                                   (Shaders/IBL_Precompute/GenIrradianceMap.comp, PreFilterEnvMap.comp)
 *   althea_cuda_generate_mips    <- Image::generateMipMaps                         Src/Image.cpp:135-239
 *   althea_cuda_brdf_lut         <- (no generator in the reference; replaces loading Content/PrecomputedMaps/brdf_lut.png,
 *                                   ImageBasedLighting.cpp:570-602)
 *   althea_cuda_import_*         <- the VMA-backed resources of GBufferResources (Src/DeferredRendering.cpp:37-155),
 *                                   ReflectionBuffer (Src/ReflectionBuffer.cpp:20-99), PointLightCollection
 *                                   (Src/PointLight.cpp:31-185), IBLResources (Src/ImageBasedLighting.cpp:448-602)
 *
 * Conventions: plain pointers and sizes, no C++/torch types. Every call returns 0 on success and a negative
 * althea_status otherwise; althea_cuda_last_error() gives the message (the reference throws std::runtime_error —
 * the C++ mirror in althea_b200/host rethrows). Calls on one ctx are made from one thread (the reference's render
 * thread); work is stream-ordered on the stream named in althea_sync (or the ctx stream). There is NO CPU fallback:
 * every compute entry point launches sm_100a kernels or fails.
 */
#ifndef ALTHEA_CUDA_H
#define ALTHEA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALTHEA_CUDA_ABI_VERSION 1

typedef struct althea_cuda_ctx althea_cuda_ctx;

typedef enum althea_status {
  ALTHEA_OK = 0,
  ALTHEA_ERR_INVALID_ARGUMENT = -1,
  ALTHEA_ERR_CUDA = -2,
  ALTHEA_ERR_UNSUPPORTED = -3,
  ALTHEA_ERR_BAD_HANDLE = -4,
  ALTHEA_ERR_OUT_OF_MEMORY = -5
} althea_status;

/* VkFormat values of the formats on the path (numeric values from vulkan_core.h, so the host passes VkFormat as is) */
enum {
  ALTHEA_FORMAT_R8G8B8A8_UNORM = 37,       /* albedo, MRO (DeferredRendering.cpp:79-99), BRDF LUT */
  ALTHEA_FORMAT_R16G16B16A16_SFLOAT = 97,  /* normal (DeferredRendering.cpp:62-77), reflection buffer (ReflectionBuffer.cpp:36) */
  ALTHEA_FORMAT_R32_SFLOAT = 100,          /* depth aspect copied out of D32_SFLOAT(_S8); shadow cube array */
  ALTHEA_FORMAT_R32G32B32A32_SFLOAT = 109, /* position (legacy), IBL maps (ImageBasedLighting.cpp:456,496,542) */
  ALTHEA_FORMAT_D32_SFLOAT = 126,          /* accepted as an alias of R32_SFLOAT for linear depth copies */
  ALTHEA_FORMAT_R8_UINT = 13               /* SSAO occluded-ray counts (engine-internal scratch) */
};

/* ---- context ------------------------------------------------------------------------------------------------- */
/* vk_device_uuid: VkPhysicalDeviceIDProperties::deviceUUID of the Vulkan device, or NULL to skip the check.        */
int althea_cuda_create(althea_cuda_ctx** out_ctx, int cuda_device, const uint8_t vk_device_uuid[16]);
void althea_cuda_destroy(althea_cuda_ctx* ctx);
const char* althea_cuda_last_error(const althea_cuda_ctx* ctx); /* ctx may be NULL: error of a failed create */
int althea_cuda_abi_version(void);

#define ALTHEA_CTX_PARITY_MATH 1u /* run the -fmad=false build of the per-frame kernels (bit-matches the CPU oracle's IEEE op order) */
#define ALTHEA_CTX_SSAO_EXACT_TAPS 2u /* SSAO marches the fp32 position texels directly (4 loads per tap) instead of the packed proxy with exact re-evaluation; same counts, slower: A/B switch for tests and profiling */
#define ALTHEA_CTX_SSAO_RAY_DEPTH_PROXY 8u /* SSAO marches 16-byte ray-depth records (eye depth of a footprint's four texels along the camera's view rays) instead of the 32-byte position records; same counts bit for bit; faster on frames without sky, slightly slower with it: opt-in */
#define ALTHEA_CTX_SSAO_COUNT_TAPS 4u /* diagnostics: the SSAO march also counts the proxy records it gathers (read with althea_cuda_diag_ssao_gathers); slower */
#define ALTHEA_CTX_SSAO_NO_CULL 16u /* SSAO without the coarse sign test (per-block plane records in shared memory that drop the march steps which cannot flip): every tap gathers its position record, as in round 1; same counts bit for bit: A/B switch */
#define ALTHEA_CTX_SSR_PLANE_SKIP 32u /* SSR skips the march steps whose two taps the plane records of the depth buffer (one {alpha, beta, gamma, r} of reciprocal eye depth per 32 x 32 texels) prove to project on one side of the ray, and whole runs of them inside a block; same hit mask bit for bit in the parity build. Opt-in: 95 % of the taps are skipped at 4K, but a skipped tap costs ~45 instructions against the fast march's ~55, so it only pays in the parity build */
#define ALTHEA_CTX_BAND_EXCHANGE_HALO 64u /* row bands: ssr_capture writes only the scissor's own rows of reflection mip 0; the host fills the halo rows the band's glossy mips read (althea_cuda_band_rows) from the ranks that own them before glossy_convolve, instead of every rank recomputing them */
int althea_cuda_set_flags(althea_cuda_ctx* ctx, uint32_t flags);

/* Row bands (multi-GPU split of ONE frame, BASELINE configs[3]): restricts ssr_capture / glossy_convolve / deferred_shade on
 * this ctx to what rows [y0, y1) of the final image need, the Vulkan scissor's role (RenderPass::begin sets viewport and
 * scissor to the full extent, Src/RenderPass.cpp:247-288). ssr_capture and glossy_convolve widen the band by the halo of
 * reflection rows the band's glossy mips depend on (recomputed locally, no halo exchange), so the band's pixels come out
 * bit-identical to a whole-frame run. y1 == 0 restores the whole frame. Inputs are always whole-frame images. */
int althea_cuda_set_scissor_rows(althea_cuda_ctx* ctx, uint32_t y0, uint32_t y1);
/* The rows [out_lo[L], out_hi[L]) of reflection mip L that a band [y0, y1) of a w x h frame reads or writes (L < mips). */
int althea_cuda_band_rows(uint32_t w, uint32_t h, uint32_t mips, uint32_t y0, uint32_t y1, uint32_t* out_lo, uint32_t* out_hi);

/* Per-kernel device timing (CUDA events on the launching stream around every kernel this library launches). */
int althea_cuda_enable_timing(althea_cuda_ctx* ctx, int enable);
/* Sums since the last reset. names[i] points at static strings. Returns the number of distinct kernels (<= cap). Synchronises the recorded events. */
int althea_cuda_get_timings(althea_cuda_ctx* ctx, const char** names, float* total_ms, uint32_t* launches, int cap);
int althea_cuda_reset_timings(althea_cuda_ctx* ctx);
/* Number of kernels this library has launched on ctx since creation (bench.py's gpu_launches). */
uint64_t althea_cuda_launch_count(const althea_cuda_ctx* ctx);

/* Diagnostics for bench.py's roofline of the SSAO march, which is bound by the rate of divergent 32-byte gathers (L1 data pipe),
 * not by HBM. Neither call is on the rendering path. _ssao_gathers: proxy records gathered by the last SSAO launch made with
 * ALTHEA_CTX_SSAO_COUNT_TAPS set. _gather_ceiling: measures the device's records/s for the same access pattern (one 256-bit
 * load per lane at random positions within +-radius records of the lane's 16x16 tile, over a (w+1) x (h+1) record grid). */
int althea_cuda_diag_ssao_gathers(althea_cuda_ctx* ctx, uint64_t* out_records);
/* ... and the taps of that launch that had to be re-evaluated from the fp32 texels (ray-depth proxy only). */
int althea_cuda_diag_ssao_exact_fallbacks(althea_cuda_ctx* ctx, uint64_t* out_taps);
/* All four counters of that launch: position records gathered, taps re-evaluated from the fp32 texels, plane-record lookups
 * of the coarse sign test (shared memory), and march steps it could not drop (evaluated exactly). */
int althea_cuda_diag_ssao_cull(althea_cuda_ctx* ctx, uint64_t out_counts[4]);
int althea_cuda_diag_gather_ceiling(althea_cuda_ctx* ctx, uint32_t w, uint32_t h, uint32_t radius, uint32_t taps_per_pixel,
                                    double* out_records_per_second);

/* ---- resources ------------------------------------------------------------------------------------------------- */
/* Linear image layout used by every entry point: row-major, row 0 = top; for mips > 1 the levels are tightly packed one
 * after another (level k is max(1,w>>k) x max(1,h>>k), row pitch = width*bytes_per_texel); layers (cube array: 6*cube+face,
 * faces +X,-X,+Y,-Y,+Z,-Z) are whole mip chains one after another. `pitch` (bytes) applies to mips == 1 only; 0 = tight. */
size_t althea_cuda_image_bytes(uint32_t vk_format, uint32_t w, uint32_t h, uint32_t mips, uint32_t layers);

#define ALTHEA_IMAGE_CUBE 1u
#define ALTHEA_IMAGE_OPTIMAL_TILING 2u /* not mappable as a linear buffer: rejected with ALTHEA_ERR_UNSUPPORTED */

/* Vulkan owns the memory, CUDA maps it (cudaImportExternalMemory, opaque fd). Release before the Vulkan object dies. */
int althea_cuda_import_image(althea_cuda_ctx* ctx, int fd, uint64_t alloc_size, uint64_t offset, uint32_t vk_format,
                             uint32_t w, uint32_t h, uint32_t mips, uint32_t layers, uint32_t flags, uint64_t pitch,
                             uint64_t* out_handle);
int althea_cuda_import_buffer(althea_cuda_ctx* ctx, int fd, uint64_t size, uint64_t offset, uint64_t* out_handle);
int althea_cuda_import_semaphore(althea_cuda_ctx* ctx, int fd, int is_timeline, uint64_t* out_handle);
/* Twins that wrap plain device pointers (tests, bench, CUDA-only hosts). The caller keeps ownership. */
int althea_cuda_wrap_linear_image(althea_cuda_ctx* ctx, void* dptr, size_t pitch, uint32_t vk_format, uint32_t w, uint32_t h,
                                  uint32_t mips, uint32_t layers, uint64_t* out_handle);
int althea_cuda_wrap_buffer(althea_cuda_ctx* ctx, void* dptr, size_t size, uint64_t* out_handle);
/* Context-owned device images/buffers + host transfers (what the C++ mirror classes use when no Vulkan device exists). */
int althea_cuda_create_image(althea_cuda_ctx* ctx, uint32_t vk_format, uint32_t w, uint32_t h, uint32_t mips, uint32_t layers,
                             uint64_t* out_handle);
int althea_cuda_create_buffer(althea_cuda_ctx* ctx, size_t size, uint64_t* out_handle);
/* host may be pageable or pinned; copies are asynchronous on `stream` (0 = ctx stream) when host is pinned. */
int althea_cuda_upload(althea_cuda_ctx* ctx, uint64_t handle, const void* host, size_t bytes, void* stream);
int althea_cuda_download(althea_cuda_ctx* ctx, uint64_t handle, void* host, size_t bytes, void* stream);
int althea_cuda_device_pointer(althea_cuda_ctx* ctx, uint64_t handle, void** out_dptr, size_t* out_bytes);
int althea_cuda_release(althea_cuda_ctx* ctx, uint64_t handle);
int althea_cuda_synchronize(althea_cuda_ctx* ctx, void* stream);

/* ---- parameter blocks (byte-for-byte the reference's) ------------------------------------------------------------ */
typedef struct althea_global_uniforms { /* 416 B == Include/Althea/GlobalUniforms.h:15-31 == Shaders/Global/GlobalUniforms.glsl:8-24 */
  float projection[16], inverseProjection[16], view[16], prevView[16], inverseView[16], prevInverseView[16]; /* column-major */
  float mouseUV[2];
  int32_t lightCount;
  uint32_t lightBufferHandle; /* bindless index in the reference; ignored here (lights_buf is passed explicitly) */
  float time, exposure;
  uint32_t inputMask, frameCount;
} althea_global_uniforms;

typedef struct althea_point_light { /* 32 B == Include/Althea/PointLight.h:31-34 == Shaders/PointLights.glsl:7-10 */
  float position[3], _pad0, emission[3], _pad1;
} althea_point_light;

typedef struct althea_gbuffer { /* image handles; GBufferResources, Src/DeferredRendering.cpp:37-155 */
  uint64_t depth;    /* R32_SFLOAT: the depth image written THIS frame (SSR.frag:29 hard-wires depthA) */
  uint64_t position; /* R32G32B32A32_SFLOAT, .a == 0 => empty (legacy DeferredPass.frag:18,44); 0 => the lighting pass and SSAO use
                        positions reconstructed from `depth` (Misc/ReconstructPosition.glsl), empty where normal.a == 0: what a
                        host without the legacy attachment passes (today's GBufferResources) */
  uint64_t normal;   /* R16G16B16A16_SFLOAT, .a == 0 => empty (SSR.frag:136-141) */
  uint64_t albedo;   /* R8G8B8A8_UNORM */
  uint64_t mro;      /* R8G8B8A8_UNORM: metallic, roughness, occlusion */
} althea_gbuffer;

typedef struct althea_ibl { /* IBLResources, Include/Althea/ImageBasedLighting.h:25-43 */
  uint64_t env;         /* RGBA32F equirect, 1 mip */
  uint64_t prefiltered; /* RGBA32F equirect, 5 mips, level k = roughness k/4 */
  uint64_t irradiance;  /* RGBA32F equirect */
  uint64_t brdf_lut;    /* RGBA8, sampled at (NdotV, roughness) (PBRMaterial.glsl:110) */
} althea_ibl;

typedef struct althea_sync { /* all zero => plain stream order on the ctx stream */
  uint64_t wait_sem, wait_value;     /* imported semaphore the work waits on before starting (0 = none) */
  uint64_t signal_sem, signal_value; /* imported semaphore signalled when the work is done (0 = none) */
  void* cuda_stream;                 /* cudaStream_t to launch on; NULL => the ctx's own stream */
} althea_sync;

/* ---- per-frame stages -------------------------------------------------------------------------------------------- */
/* Writes mip 0 of `reflection` (RGBA16F, >= 1 mip, frame-sized). lights_buf: althea_point_light[lightCount];
 * shadow_cube_array: R32_SFLOAT, layers = 6*lightCount, texel = length(p-light)/1000 (ShadowMapBindless.frag:41); 0 => unshadowed. */
int althea_cuda_ssr_capture(althea_cuda_ctx* ctx, const althea_global_uniforms* uniforms, const althea_gbuffer* gbuffer,
                            const althea_ibl* ibl, uint64_t lights_buf, uint64_t shadow_cube_array, uint64_t reflection,
                            const althea_sync* sync);
/* Writes mips 1..mips-1 of `reflection` from mip 0 (7-tap separable Gaussian, axis alternating V,H,V,H). */
int althea_cuda_glossy_convolve(althea_cuda_ctx* ctx, uint64_t reflection, const althea_sync* sync);

#define ALTHEA_SHADE_SKIP_TONEMAP 1u /* DeferredPass.frag:48-50,86-88 */
#define ALTHEA_SHADE_NO_SSAO 2u      /* keep the G-buffer occlusion channel instead of computeSSAO */
#define ALTHEA_SHADE_AO_FROM_IMAGE 4u /* read occluded-ray counts from ao_counts instead of computing them (tests; second half of a split frame) */
#define ALTHEA_SHADE_AO_ONLY 8u       /* compute the occluded-ray counts into ao_counts (required) and stop: no shading, out_color and reflection are not touched. With AO_FROM_IMAGE on a later call this splits the stage, so that a row-band host can run SSAO while the reflection halo is exchanged */
/* out_color: RGBA16F or RGBA32F, frame-sized. ao_counts: optional R8_UINT frame-sized image; when non-zero the SSAO
 * kernel writes its per-pixel occluded-ray counts there (or, with AO_FROM_IMAGE, reads them). 0 => internal scratch. */
int althea_cuda_deferred_shade(althea_cuda_ctx* ctx, const althea_global_uniforms* uniforms, const althea_gbuffer* gbuffer,
                               const althea_ibl* ibl, uint64_t lights_buf, uint64_t shadow_cube_array, uint64_t reflection,
                               uint64_t out_color, uint64_t ao_counts, uint32_t flags, const althea_sync* sync);

/* ---- rasterising producers of the path's inputs ------------------------------------------ */
/* The engine's own vertex, Include/Althea/Common/InstanceDataCommon.h:45-53 (what Primitive's VertexBuffer<Vertex> holds). */
typedef struct althea_vertex {
  float position[3], tangent[3], bitangent[3], normal[3];
  float uvs[4][2];
  float weights[4];
  uint16_t joints[4];
} althea_vertex; /* 104 bytes */

/* sampler word: wrapU | wrapV << 2 | magNearest << 4 | minNearest << 5 | mipMode << 6 | srgb << 8, the SamplerOptions the
 * reference derives from the glTF sampler (Src/Sampler.cpp:11-88); wrap: 0 REPEAT, 1 CLAMP_TO_EDGE, 2 MIRRORED_REPEAT;
 * mipMode: 0 none, 1 NEAREST, 2 LINEAR. Filtering is isotropic (the reference enables the device's maximum anisotropy, which
 * Vulkan leaves implementation-defined). */
#define ALTHEA_SAMPLER_WORD(wrapU, wrapV, magNearest, minNearest, mipMode, srgb) \
  ((uint32_t)(wrapU) | ((uint32_t)(wrapV) << 2) | ((uint32_t)(magNearest) << 4) | ((uint32_t)(minNearest) << 5) | ((uint32_t)(mipMode) << 6) | ((uint32_t)(srgb) << 8))
typedef struct althea_texture_ref {
  uint64_t image; /* R8G8B8A8_UNORM image handle with its mip chain; 0 => the reference's 1x1 default for the slot
                     (white / normal1x1 / white, Src/Material.cpp:36-75, Src/DefaultTextures.cpp:24-31) */
  uint32_t sampler;
  uint32_t _pad;
} althea_texture_ref;

typedef struct althea_material { /* the MaterialConstants fields fetchMaterial reads, InstanceDataCommon.h:8-33 */
  float baseColorFactor[4];
  int32_t baseTextureCoordinateIndex, metallicRoughnessTextureCoordinateIndex;
  float normalScale, metallicFactor, roughnessFactor, alphaCutoff;
  althea_texture_ref baseTexture, normalTexture, metallicRoughnessTexture;
} althea_material;

typedef struct althea_primitive { /* one Primitive of a Model, as Model::draw / drawShadowMaps issue it */
  uint64_t vertices; /* buffer of althea_vertex */
  uint64_t indices;  /* buffer of uint32 (Primitive's IndexBuffer) */
  uint32_t index_count;
  uint32_t front_face_clockwise; /* Primitive::getFrontFace() == VK_FRONT_FACE_CLOCKWISE */
  float model[16];               /* the node's transform, transformBuffer[nodeIdx] (Gltf.vert:49) */
  althea_material material;
} althea_primitive;

typedef struct althea_point_light_constants { /* PointLightConstants, Src/PointLight.cpp:72-118 / Shaders/PointLights.glsl */
  float projection[16], inverseProjection[16];
  float views[6][16], inverseViews[6][16];
} althea_point_light_constants;

/* SceneToGBufferPass + Gltf.vert/.frag: rasterises the primitives in order (depth LESS, back faces culled, alpha-cutoff
 * discard) into the G-buffer attachments; any of gbuffer's handles may be 0 (not written). position (legacy attachment)
 * receives (world position, 1); uncovered pixels get the reference's clears (colour 0, depth 1). The colour attachments are
 * alpha-blended in draw order as the reference's pipelines do (Src/GraphicsPipeline.cpp:138-154; up to four translucent layers
 * per pixel). Skinned primitives:
 * ALTHEA_ERR_UNSUPPORTED is never raised here because skinning data is not part of althea_primitive; pass skinned
 * geometry pre-transformed. */
int althea_cuda_draw_gbuffer(althea_cuda_ctx* ctx, const althea_global_uniforms* uniforms, const althea_primitive* primitives,
                             uint32_t primitive_count, const althea_gbuffer* gbuffer, const althea_sync* sync);
/* PointLightCollection::drawShadowMaps + ShadowMapBindless.vert/.frag: for every light, the 6 cube faces of
 * shadow_cube_array (R32_SFLOAT, square, >= 6*light_count layers) receive min over fragments of length(p - light) / 1000,
 * cleared to 1. Face cameras come from `constants` exactly as the reference's vertex shader reads them. */
int althea_cuda_draw_shadow_cubes(althea_cuda_ctx* ctx, uint64_t lights_buf, uint32_t light_count,
                                  const althea_point_light_constants* constants, const althea_primitive* primitives,
                                  uint32_t primitive_count, uint64_t shadow_cube_array, const althea_sync* sync);

/* ---- IBL precompute ------------------------------------------------------------------------------------------------ */
#define ALTHEA_IBL_LAYOUT_EQUIRECT 0u   /* the reference's layout: outputs are equirect images */
#define ALTHEA_IBL_LAYOUT_CUBE 1u       /* BASELINE config 2: outputs are 6-layer cube images */
#define ALTHEA_IBL_SEQ_REFERENCE_HASH 0u /* PreFilterEnvMap.comp:36-42 per-texel hash RNG */
#define ALTHEA_IBL_SEQ_HAMMERSLEY 1u

typedef struct althea_ibl_precompute_desc {
  uint32_t layout;            /* ALTHEA_IBL_LAYOUT_* */
  uint32_t sequence;          /* ALTHEA_IBL_SEQ_* (prefilter only) */
  uint32_t prefilter_samples; /* 0 => 10000 (PreFilterEnvMap.comp:5) */
  uint32_t theta_samples;     /* 0 => 300 (GenIrradianceMap.comp:119) */
} althea_ibl_precompute_desc;

/* Builds the full mip chain of `image` in place from level 0 (LINEAR 2:1 blit chain, Src/Image.cpp:183-213). */
int althea_cuda_generate_mips(althea_cuda_ctx* ctx, uint64_t image, const althea_sync* sync);
/* env_with_mips: RGBA32F equirect WITH its full mip chain (ImageBasedLighting.cpp:153-166).
 * out_irradiance (0 = skip): RGBA32F; equirect layout: any size (the reference uses the env size); cube: layers = 6.
 * out_prefiltered (0 = skip): RGBA32F with n mips. Equirect layout: the reference's 5 images env>>1..env>>5 ARE the 5 mips
 *   of this image (level k: roughness k/4). Cube layout: layers = 6, level k: roughness k/(n-1). */
int althea_cuda_ibl_precompute(althea_cuda_ctx* ctx, uint64_t env_with_mips, const althea_ibl_precompute_desc* desc,
                               uint64_t out_irradiance, uint64_t out_prefiltered, const althea_sync* sync);
/* out_lut: RGBA8 (R = scale, G = bias, B = 0, A = 255) or RGBA32F, square. Row 0 holds roughness = 1 (the orientation
 * of the reference's asset, which PBRMaterial.glsl:110 samples un-flipped). */
int althea_cuda_brdf_lut(althea_cuda_ctx* ctx, uint32_t samples, uint64_t out_lut, const althea_sync* sync);

#ifdef __cplusplus
}
#endif
#endif /* ALTHEA_CUDA_H */
