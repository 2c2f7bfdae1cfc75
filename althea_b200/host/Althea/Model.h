// Include/Althea/Model.h, Primitive.h, Material.h, Texture.h: what the rasterising producers need of them. A Model is a list of
// Primitives (vertex buffer of the engine's Vertex, uint32 index buffer, node transform, front face) each with a Material
// (the MaterialConstants fetchMaterial reads + its three textures). The glTF parsing of the reference (cesium-native) is not part
// of the path; primitives are built from finished vertex / index arrays, exactly what Primitive's constructor ends up holding
// (Src/Primitive.cpp:367-443).
#pragma once
#include "CudaApplication.h"

#include <cstring>
#include <memory>

namespace AltheaEngine {

struct Vertex { // Include/Althea/Common/InstanceDataCommon.h:45-53 == althea_vertex
  float position[3], tangent[3], bitangent[3], normal[3];
  float uvs[4][2];
  float weights[4];
  uint16_t joints[4];
};
static_assert(sizeof(Vertex) == 104 && sizeof(althea_vertex) == 104, "InstanceDataCommon.h:45-53");

class BufferResource { // VertexBuffer<T> / IndexBuffer's role: device buffer + handle, released with the object
public:
  BufferResource() = default;
  BufferResource(const CudaApplication& app, const void* data, size_t bytes) : _app(&app), _bytes(bytes) {
    app.check(althea_cuda_create_buffer(app.ctx(), bytes, &_handle), "althea_cuda_create_buffer");
    app.check(althea_cuda_upload(app.ctx(), _handle, data, bytes, nullptr), "althea_cuda_upload");
    app.waitIdle();
  }
  ~BufferResource() {
    if (_handle && _app) althea_cuda_release(_app->ctx(), _handle);
  }
  BufferResource(BufferResource&& o) noexcept { std::swap(_app, o._app); std::swap(_handle, o._handle); std::swap(_bytes, o._bytes); }
  BufferResource& operator=(BufferResource&& o) noexcept { std::swap(_app, o._app); std::swap(_handle, o._handle); std::swap(_bytes, o._bytes); return *this; }
  BufferResource(const BufferResource&) = delete;
  BufferResource& operator=(const BufferResource&) = delete;
  uint64_t handle() const { return _handle; }

private:
  const CudaApplication* _app = nullptr;
  uint64_t _handle = 0;
  size_t _bytes = 0;
};

class Texture { // Src/Texture.cpp:47-99: RGBA8 image with its mip chain + the sampler derived from the glTF sampler
public:
  Texture() = default;
  Texture(const CudaApplication& app, const void* texelsAllLevels, uint32_t w, uint32_t h, uint32_t mips, uint32_t samplerWord)
      : _image(app, ALTHEA_FORMAT_R8G8B8A8_UNORM, w, h, mips, 1), _sampler(samplerWord) {
    _image.upload(texelsAllLevels, _image.byteSize());
    app.waitIdle();
  }
  althea_texture_ref ref() const { return althea_texture_ref{_image.handle(), _sampler, 0}; }

private:
  ImageResource _image;
  uint32_t _sampler = 0;
};

struct Material { // Src/Material.cpp:14-110 (defaults of a primitive without a glTF material)
  float baseColorFactor[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  int32_t baseTextureCoordinateIndex = 0, metallicRoughnessTextureCoordinateIndex = 0;
  float normalScale = 1.0f, metallicFactor = 0.0f, roughnessFactor = 1.0f, alphaCutoff = 0.5f;
  std::shared_ptr<Texture> baseTexture, normalTexture, metallicRoughnessTexture; // null => the reference's 1x1 defaults
};

class Primitive {
public:
  Primitive(const CudaApplication& app, const std::vector<Vertex>& vertices, const std::vector<uint32_t>& indices, const float nodeTransform[16],
            Material material, bool flipFrontFace = false)
      : _vertexBuffer(app, vertices.data(), vertices.size() * sizeof(Vertex)), _indexBuffer(app, indices.data(), indices.size() * sizeof(uint32_t)),
        _indexCount((uint32_t)indices.size()), _material(std::move(material)), _flipFrontFace(flipFrontFace) {
    if (vertices.empty()) throw std::runtime_error("Attempting to create a primitive with no vertices!"); // Src/Primitive.cpp:424
    std::memcpy(_transform, nodeTransform, sizeof _transform);
  }
  bool isFrontFaceClockwise() const { return _flipFrontFace; } // getFrontFace() == VK_FRONT_FACE_CLOCKWISE (Primitive.h:93)
  althea_primitive describe() const {
    althea_primitive a;
    std::memset(&a, 0, sizeof a);
    a.vertices = _vertexBuffer.handle();
    a.indices = _indexBuffer.handle();
    a.index_count = _indexCount;
    a.front_face_clockwise = _flipFrontFace ? 1u : 0u;
    std::memcpy(a.model, _transform, sizeof a.model);
    std::memcpy(a.material.baseColorFactor, _material.baseColorFactor, sizeof a.material.baseColorFactor);
    a.material.baseTextureCoordinateIndex = _material.baseTextureCoordinateIndex;
    a.material.metallicRoughnessTextureCoordinateIndex = _material.metallicRoughnessTextureCoordinateIndex;
    a.material.normalScale = _material.normalScale;
    a.material.metallicFactor = _material.metallicFactor;
    a.material.roughnessFactor = _material.roughnessFactor;
    a.material.alphaCutoff = _material.alphaCutoff;
    if (_material.baseTexture) a.material.baseTexture = _material.baseTexture->ref();
    if (_material.normalTexture) a.material.normalTexture = _material.normalTexture->ref();
    if (_material.metallicRoughnessTexture) a.material.metallicRoughnessTexture = _material.metallicRoughnessTexture->ref();
    return a;
  }

private:
  BufferResource _vertexBuffer, _indexBuffer;
  uint32_t _indexCount = 0;
  float _transform[16];
  Material _material;
  bool _flipFrontFace = false;
};

class Model {
public:
  void addPrimitive(Primitive&& p) { _primitives.emplace_back(std::move(p)); }
  const std::vector<Primitive>& getPrimitives() const { return _primitives; }

private:
  std::vector<Primitive> _primitives;
};

// every primitive of every model, in draw order (the loops of PointLight.cpp:262-276 and of the G-buffer subpass)
inline std::vector<althea_primitive> describeModels(const std::vector<Model>& models) {
  std::vector<althea_primitive> out;
  for (const Model& m : models)
    for (const Primitive& p : m.getPrimitives()) out.push_back(p.describe());
  return out;
}

} // namespace AltheaEngine
