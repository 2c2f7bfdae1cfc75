// ORACLE — TEST INFRASTRUCTURE ONLY. Driver over the REFERENCE's own Camera class: compiled together with
// /root/reference/Src/Camera.cpp (where it lies; never copied) and the reference's vendored GLM, with the GLM switches of
// its CMakeLists.txt:89-96, into oracle/_ref/libcamera_ref.so by oracle/Makefile's `ref` target.
//   ref_camera                 one camera as the engine's CameraController sets it up (Src/Camera.cpp:7-110)
//   ref_point_light_constants  the six cube-face cameras exactly as PointLightCollection's constructor derives them
//                              (Src/PointLight.cpp:72-118): projection, inverse, views[6], inverseViews[6]
// tests/test_camera_pin.py holds scene.make_uniforms, model.point_light_constants and the C++ mirror's pointLightConstants to it.
#include "Camera.h"

#include <glm/gtc/matrix_inverse.hpp>

#include <cstring>

using AltheaEngine::Camera;

extern "C" {

void ref_camera(float fovDegrees, float aspect, float nearPlane, float farPlane, const float* position, float yawRadians,
                float pitchRadians, float* projection16, float* transform16, float* view16) {
  Camera camera(fovDegrees, aspect, nearPlane, farPlane);
  camera.setPosition(glm::vec3(position[0], position[1], position[2]));
  camera.setRotationRadians(yawRadians, pitchRadians);
  const glm::mat4 view = camera.computeView();
  std::memcpy(projection16, &camera.getProjection(), 64);
  std::memcpy(transform16, &camera.getTransform(), 64);
  std::memcpy(view16, &view, 64);
}

void ref_point_light_constants(float* out /* 14 matrices: projection, inverseProjection, views[6], inverseViews[6] */) {
  Camera camera(90.0f, 1.0f, 0.01f, 1000.0f);
  glm::mat4 m[14];
  m[0] = camera.getProjection();
  m[1] = glm::inverse(m[0]);
  camera.setPosition(glm::vec3(0.0f));
  const float yawPitch[6][2] = {{90.0f, 0.0f}, {-90.0f, 0.0f}, {180.0f, 90.0f}, {180.0f, -90.0f}, {180.0f, 0.0f}, {0.0f, 0.0f}};
  for (int f = 0; f < 6; ++f) {
    camera.setRotationDegrees(yawPitch[f][0], yawPitch[f][1]);
    m[2 + f] = camera.computeView();
    m[8 + f] = glm::inverse(m[2 + f]);
  }
  std::memcpy(out, m, sizeof m);
}

} // extern "C"
