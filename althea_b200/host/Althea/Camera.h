// Host-side mirror of Include/Althea/Camera.h / Src/Camera.cpp:7-110: the matrices the path consumes (GlobalUniforms'
// projection / view and their inverses, and the six cube-face views of PointLightConstants) come out of this class in the
// reference. It is written against GLM there; here the handful of GLM operations it uses are spelled out in fp32 in GLM's
// evaluation order (column-major storage, m[column][row]):
//   perspective  = glm::perspective under GLM_FORCE_DEPTH_ZERO_TO_ONE, right-handed (CMakeLists.txt:89-96), Y flipped
//   computeView  = glm::affineInverse: cofactor inverse of the 3x3 block, translation -(inv * t)
//   inverse      = glm::inverse(mat4): 2x2 sub-determinants, cofactor columns, one division by the determinant
// so that the bits agree with the reference's build of the class (tests/test_camera_pin.py, oracle/_ref/libcamera_ref.so),
// down to the rounding noise of sin/cos at 180 and +-90 degrees that decides how the +-Y shadow-cube faces are rotated.
#pragma once

#include <cmath>
#include <cstring>

namespace AltheaEngine {

struct Mat4 {
  float m[4][4]; // m[column][row], as glm::mat4
  static Mat4 identity() {
    Mat4 r{};
    r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f;
    return r;
  }
  const float* data() const { return &m[0][0]; }
};

// glm::inverse(mat4)
inline Mat4 inverse(const Mat4& a) {
  const float(*m)[4] = a.m;
  const float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
  const float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
  const float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
  const float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
  const float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
  const float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
  const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  const float v0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, v1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
  const float v2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, v3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
  const float sa[4] = {1.0f, -1.0f, 1.0f, -1.0f}, sb[4] = {-1.0f, 1.0f, -1.0f, 1.0f};
  Mat4 r;
  for (int i = 0; i < 4; ++i) {
    r.m[0][i] = (v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i]) * sa[i];
    r.m[1][i] = (v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i]) * sb[i];
    r.m[2][i] = (v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i]) * sa[i];
    r.m[3][i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb[i];
  }
  const float d0 = m[0][0] * r.m[0][0], d1 = m[0][1] * r.m[1][0], d2 = m[0][2] * r.m[2][0], d3 = m[0][3] * r.m[3][0];
  const float oneOverDet = 1.0f / ((d0 + d1) + (d2 + d3));
  for (int c = 0; c < 4; ++c)
    for (int i = 0; i < 4; ++i) r.m[c][i] *= oneOverDet;
  return r;
}

// glm::affineInverse(mat4)
inline Mat4 affineInverse(const Mat4& a) {
  const float(*m)[4] = a.m;
  const float oneOverDet = 1.0f / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) +
                                   m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
  float i[3][3];
  i[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * oneOverDet;
  i[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * oneOverDet;
  i[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * oneOverDet;
  i[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * oneOverDet;
  i[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * oneOverDet;
  i[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * oneOverDet;
  i[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * oneOverDet;
  i[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * oneOverDet;
  i[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * oneOverDet;
  Mat4 r{};
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < 3; ++k) r.m[c][k] = i[c][k];
  const float tx = m[3][0], ty = m[3][1], tz = m[3][2];
  for (int k = 0; k < 3; ++k) r.m[3][k] = (-i[0][k]) * tx + (-i[1][k]) * ty + (-i[2][k]) * tz;
  r.m[3][3] = 1.0f;
  return r;
}

class Camera {
public:
  Camera() = default;
  Camera(float fovDegrees, float aspectRatio, float nearPlane, float farPlane)
      : _fov(radians(fovDegrees)), _aspectRatio(aspectRatio), _nearPlane(nearPlane), _farPlane(farPlane) {
    _recomputeProjection();
  }
  void setFovDegrees(float fovDegrees) { _fov = radians(fovDegrees); _recomputeProjection(); }
  void setAspectRatio(float aspectRatio) { _aspectRatio = aspectRatio; _recomputeProjection(); }
  void setClippingPlanes(float nearPlane, float farPlane) { _nearPlane = nearPlane; _farPlane = farPlane; _recomputeProjection(); }
  void setPosition(float x, float y, float z) {
    _transform.m[3][0] = x; _transform.m[3][1] = y; _transform.m[3][2] = z; _transform.m[3][3] = 1.0f;
  }
  void setRotationDegrees(float yawDegrees, float pitchDegrees) { setRotationRadians(radians(yawDegrees), radians(pitchDegrees)); }
  void setRotationRadians(float yawRadians, float pitchRadians) {
    const float pitchLimit = 3.14159265358979323846264338327950288f - 0.01f;
    pitchRadians = pitchRadians < -pitchLimit ? -pitchLimit : (pitchRadians > pitchLimit ? pitchLimit : pitchRadians);
    // the reference calls the unqualified C cos / sin on floats: double-precision functions, and the products with cosPitch
    // are formed in double before they are narrowed into the glm::vec3
    const float cosPitch = (float)std::cos((double)pitchRadians);
    const float z[3] = {(float)(std::sin((double)yawRadians) * (double)cosPitch), (float)(-std::sin((double)pitchRadians)),
                        (float)(std::cos((double)yawRadians) * (double)cosPitch)};
    // cross((0, 1, 0), z), all terms kept: the products with 0 decide the signs of the zeros
    float x[3] = {1.0f * z[2] - z[1] * 0.0f, 0.0f * z[0] - z[2] * 0.0f, 0.0f * z[1] - z[0] * 1.0f};
    const float inv = 1.0f / std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); // glm::normalize = v * inversesqrt(dot(v, v))
    for (float& v : x) v *= inv;
    const float y[3] = {z[1] * x[2] - x[1] * z[2], z[2] * x[0] - x[2] * z[0], z[0] * x[1] - x[0] * z[1]};
    for (int k = 0; k < 3; ++k) { _transform.m[0][k] = x[k]; _transform.m[1][k] = y[k]; _transform.m[2][k] = z[k]; }
    _transform.m[0][3] = _transform.m[1][3] = _transform.m[2][3] = 0.0f;
  }
  const Mat4& getTransform() const { return _transform; }
  const Mat4& getProjection() const { return _projection; }
  Mat4 computeView() const { return affineInverse(_transform); }
  float computeYaw() const { return std::atan2(_transform.m[2][0], _transform.m[2][2]); }
  float computePitch() const {
    const float* z = _transform.m[2];
    return -std::atan2(z[1], std::sqrt(z[0] * z[0] + z[2] * z[2]));
  }

  static float radians(float degrees) { return degrees * 0.01745329251994329576923690768489f; }

private:
  void _recomputeProjection() {
    const float tanHalfFovy = std::tan(_fov / 2.0f);
    Mat4 p{};
    p.m[0][0] = 1.0f / (_aspectRatio * tanHalfFovy);
    p.m[1][1] = 1.0f / tanHalfFovy;
    p.m[2][2] = _farPlane / (_nearPlane - _farPlane);
    p.m[2][3] = -1.0f;
    p.m[3][2] = -(_farPlane * _nearPlane) / (_farPlane - _nearPlane);
    p.m[1][1] *= -1.0f; // Vulkan screen-Y convention (Camera.cpp:108)
    _projection = p;
  }

  float _fov = 0.0f, _aspectRatio = 1.0f, _nearPlane = 0.01f, _farPlane = 1000.0f;
  Mat4 _transform = Mat4::identity();
  Mat4 _projection = Mat4::identity();
};

// The six cube-face cameras of PointLightCollection's constructor (Src/PointLight.cpp:72-118), as 14 matrices:
// projection, inverseProjection, views[6], inverseViews[6] == the layout of althea_point_light_constants.
inline void pointLightConstantMatrices(float out[14 * 16]) {
  Camera camera(90.0f, 1.0f, 0.01f, 1000.0f);
  Mat4 m[14];
  m[0] = camera.getProjection();
  m[1] = inverse(m[0]);
  camera.setPosition(0.0f, 0.0f, 0.0f);
  const float yawPitch[6][2] = {{90.0f, 0.0f}, {-90.0f, 0.0f}, {180.0f, 90.0f}, {180.0f, -90.0f}, {180.0f, 0.0f}, {0.0f, 0.0f}};
  for (int f = 0; f < 6; ++f) { // X+ X- Y+ Y- Z+ Z-
    camera.setRotationDegrees(yawPitch[f][0], yawPitch[f][1]);
    m[2 + f] = camera.computeView();
    m[8 + f] = inverse(m[2 + f]);
  }
  std::memcpy(out, m, sizeof m);
}

} // namespace AltheaEngine
