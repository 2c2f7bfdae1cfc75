/* althea_host.h — C ABI of the host-side (CPU, no CUDA) geometry preparation that feeds the G-buffer producer.
 *
 * These are the vertex-stream fix-ups the reference performs while it loads a glTF primitive, before any vertex
 * reaches the GPU (Src/Primitive.cpp:147-191). They are host work in the reference too; they live in their own
 * library (althea_b200/lib/libalthea_host.so, plain g++) so that tools without a GPU can prepare models.
 *
 *   althea_host_compute_flat_normals   <- GeometryUtilities::computeFlatNormals    Include/Althea/GeometryUtilities.h:32-48
 *   althea_host_compute_tangent_space  <- GeometryUtilities::computeTangentSpace   Include/Althea/GeometryUtilities.h:51-70,137-155
 *                                         (MikkTSpace genTangSpaceDefault, Extern/MikkTSpace/mikktspace.c, through the
 *                                         m_setTSpaceBasic callback)
 *   althea_host_save_hdri              <- Utilities::saveHdri                       Src/Utilities.cpp:244-255
 *                                         (stb_image_write.h stbi_write_hdr: Radiance RGBE, run-length coded rows)
 *   althea_host_save_exr               <- Utilities::saveExr                        Src/Utilities.cpp:258-271
 *                                         (tinyexr SaveEXR, 4 x FLOAT, scan lines)
 *   althea_host_load_hdri(_info)       <- Utilities::loadHdri                       Src/Utilities.cpp:189-213
 *                                         (stb_image.h stbi_loadf_from_memory, 4 channels requested)
 *     the pair ImageBasedLighting::createResources uses for its on-disk cache, Src/ImageBasedLighting.cpp:415-446
 *   althea_host_camera                 <- Camera (Include/Althea/Camera.h, Src/Camera.cpp:7-110): projection, transform, view
 *   althea_host_point_light_constants  <- the PointLightConstants PointLightCollection's constructor derives from six
 *                                         Cameras, Src/PointLight.cpp:72-118 (the `constants` argument of
 *                                         althea_cuda_draw_shadow_cubes)
 *
 * Geometry inputs are de-indexed triangle lists, three consecutive vertices per face, tightly packed floats, exactly what
 * Primitive.cpp hands over after duplicating vertices (:147). Returns 0 on success, -1 on a null pointer or an impossible
 * size, -2 when memory runs out.
 */
#ifndef ALTHEA_HOST_H
#define ALTHEA_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALTHEA_HOST_ABI_VERSION 1

int althea_host_abi_version(void);

/* normal_out[9 * f + 3 * k ..] = normalize(cross(p1 - p0, p2 - p0)) of face f for its three vertices k */
int althea_host_compute_flat_normals(const float* position /* 9 floats per face */, uint64_t face_count,
                                     float* normal_out /* 9 floats per face */);

/* tangent_out and bitangent_out: 9 floats per face each. bitangent = sign * cross(normal, tangent) with the sign the
 * tangent-space generator reports for the corner; corners nothing can be derived for get tangent (1, 0, 0), sign -1. */
int althea_host_compute_tangent_space(const float* position /* 9 per face */, const float* normal /* 9 per face */,
                                      const float* uv /* 6 per face: the normal map's uv set */, uint64_t face_count,
                                      float* tangent_out, float* bitangent_out);

/* Writes width x height RGBA32F texels (row 0 first; alpha is not stored) as a Radiance .hdr file, byte-identical to
 * stbi_write_hdr's output except for the header's comment line. -1: bad argument, -2: the file cannot be written. */
int althea_host_save_hdri(const char* path, int32_t width, int32_t height, const float* rgba);

/* Writes width x height RGBA32F texels as an OpenEXR file with the layout Utilities::saveExr produces (Src/Utilities.cpp:258-271,
 * tinyexr SaveEXR with four FLOAT channels A, B, G, R, scan lines, increasing y), uncompressed; the reference's LoadEXR decodes it
 * to the same bits (tests/test_hdr_cache.py). The dump format of BASELINE configs[0]'s golden frame. -1: bad argument, -2: the
 * file cannot be written. */
int althea_host_save_exr(const char* path, int32_t width, int32_t height, const float* rgba);

/* Size of a .hdr file's image. -1: bad argument, -2: cannot be read, -3: not a Radiance RGBE file stbi_loadf would accept. */
int althea_host_load_hdri_info(const char* path, int32_t* width, int32_t* height);

/* Decodes into rgba_out (4 floats per texel, alpha = 1, as loadHdri returns it). capacity_floats must be at least
 * 4 * width * height of althea_host_load_hdri_info; -4 if it is not. */
int althea_host_load_hdri(const char* path, float* rgba_out, uint64_t capacity_floats);

/* One Camera: constructed with (fov, aspect, near, far), then setPosition and setRotationRadians. Each output is 16 floats,
 * column-major as glm::mat4: getProjection(), getTransform(), computeView(), and glm::inverse of the projection. Any output
 * pointer may be NULL. Bit-identical to the reference's class built against its GLM (tests/test_camera_pin.py). */
int althea_host_camera(float fov_degrees, float aspect, float near_plane, float far_plane, const float position[3],
                       float yaw_radians, float pitch_radians, float* projection16, float* transform16, float* view16,
                       float* inverse_projection16);

/* 14 matrices of 16 floats: projection, inverseProjection, views[6], inverseViews[6] == althea_point_light_constants. */
int althea_host_point_light_constants(float* matrices224);

#ifdef __cplusplus
}
#endif
#endif /* ALTHEA_HOST_H */
