// C ABI over the header-only host mirrors (include/althea_host.h). Built by build_host.build_lib() with plain g++.
#include "althea_host.h"

#include <exception>

#include "Althea/Camera.h"
#include "Althea/GeometryUtilities.h"
#include "Althea/Utilities.h"

using namespace AltheaEngine::tangent_space_detail;

extern "C" {

int althea_host_abi_version(void) { return ALTHEA_HOST_ABI_VERSION; }

int althea_host_compute_flat_normals(const float* position, uint64_t face_count, float* normal_out) {
  if (!position || !normal_out) return -1;
  for (uint64_t f = 0; f < face_count; ++f) {
    const float* p = position + 9 * f;
    const F3 a{p[0], p[1], p[2]}, b{p[3], p[4], p[5]}, c{p[6], p[7], p[8]};
    const F3 ab = b - a, ac = c - a;
    const F3 n = unit(F3{ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x});
    for (int k = 0; k < 3; ++k) {
      normal_out[9 * f + 3 * k] = n.x;
      normal_out[9 * f + 3 * k + 1] = n.y;
      normal_out[9 * f + 3 * k + 2] = n.z;
    }
  }
  return 0;
}

int althea_host_compute_tangent_space(const float* position, const float* normal, const float* uv, uint64_t face_count,
                                      float* tangent_out, float* bitangent_out) {
  if (!position || !normal || !uv || !tangent_out || !bitangent_out) return -1;
  if (face_count == 0) return 0;
  if (face_count > 0x55555555ull) return -1; // corner indices are 32-bit
  std::vector<float> sign;
  try {
    sign.resize(face_count * 3);
    generate(position, normal, uv, (size_t)face_count, tangent_out, sign.data());
  } catch (const std::exception&) { // out of memory: nothing may cross the C boundary
    return -2;
  }
  for (uint64_t c = 0; c < face_count * 3; ++c) {
    const float* N = normal + 3 * c;
    const float* T = tangent_out + 3 * c;
    bitangent_out[3 * c] = sign[c] * (N[1] * T[2] - N[2] * T[1]);
    bitangent_out[3 * c + 1] = sign[c] * (N[2] * T[0] - N[0] * T[2]);
    bitangent_out[3 * c + 2] = sign[c] * (N[0] * T[1] - N[1] * T[0]);
  }
  return 0;
}

int althea_host_save_hdri(const char* path, int32_t width, int32_t height, const float* rgba) {
  if (!path || !rgba || width <= 0 || height <= 0) return -1;
  try {
    AltheaEngine::Utilities::saveHdri(path, width, height, reinterpret_cast<const std::byte*>(rgba), (size_t)width * height * 16);
  } catch (const std::exception&) {
    return -2;
  }
  return 0;
}

int althea_host_save_exr(const char* path, int32_t width, int32_t height, const float* rgba) {
  if (!path || !rgba || width <= 0 || height <= 0) return -1;
  try {
    AltheaEngine::Utilities::saveExr(path, width, height, reinterpret_cast<const std::byte*>(rgba), (size_t)width * height * 16);
  } catch (const std::exception&) {
    return -2;
  }
  return 0;
}

static int loadHdri(const char* path, int32_t* width, int32_t* height, std::vector<float>& rgba) {
  std::vector<uint8_t> file;
  try {
    file = AltheaEngine::Utilities::readFile(path);
  } catch (const std::exception&) {
    return -2;
  }
  int w = 0, h = 0;
  try {
    if (!AltheaEngine::Utilities::decodeHdri(file.data(), file.size(), w, h, rgba)) return -3;
  } catch (const std::exception&) { // a header announcing more texels than memory holds
    return -3;
  }
  *width = w;
  *height = h;
  return 0;
}

int althea_host_load_hdri_info(const char* path, int32_t* width, int32_t* height) {
  if (!path || !width || !height) return -1;
  std::vector<float> rgba;
  return loadHdri(path, width, height, rgba);
}

int althea_host_load_hdri(const char* path, float* rgba_out, uint64_t capacity_floats) {
  if (!path || !rgba_out) return -1;
  std::vector<float> rgba;
  int32_t w = 0, h = 0;
  const int rc = loadHdri(path, &w, &h, rgba);
  if (rc != 0) return rc;
  if (capacity_floats < rgba.size()) return -4;
  std::memcpy(rgba_out, rgba.data(), rgba.size() * sizeof(float));
  return 0;
}

int althea_host_camera(float fov_degrees, float aspect, float near_plane, float far_plane, const float position[3], float yaw_radians,
                       float pitch_radians, float* projection16, float* transform16, float* view16, float* inverse_projection16) {
  if (!position) return -1;
  AltheaEngine::Camera camera(fov_degrees, aspect, near_plane, far_plane);
  camera.setPosition(position[0], position[1], position[2]);
  camera.setRotationRadians(yaw_radians, pitch_radians);
  if (projection16) std::memcpy(projection16, camera.getProjection().data(), 64);
  if (transform16) std::memcpy(transform16, camera.getTransform().data(), 64);
  if (view16) {
    const AltheaEngine::Mat4 v = camera.computeView();
    std::memcpy(view16, v.data(), 64);
  }
  if (inverse_projection16) {
    const AltheaEngine::Mat4 ip = AltheaEngine::inverse(camera.getProjection());
    std::memcpy(inverse_projection16, ip.data(), 64);
  }
  return 0;
}

int althea_host_point_light_constants(float* matrices224) {
  if (!matrices224) return -1;
  AltheaEngine::pointLightConstantMatrices(matrices224);
  return 0;
}

} // extern "C"
