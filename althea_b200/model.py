"""Host-side mirror of the reference's Model / Primitive / Material / Texture for the rasterising producers
(althea_cuda_draw_gbuffer, althea_cuda_draw_shadow_cubes): numpy descriptions of primitives, their upload into engine buffers,
a GLB (binary glTF 2.0) loader that builds the engine's `Vertex` buffers the way Src/Primitive.cpp:51-367 does, and small
procedural meshes for tests. Host plumbing only: nothing here computes pixels.

Differences from the reference's loader, all upstream of the C ABI (which takes finished Vertex buffers):
  * tangents missing from the file come from libalthea_host.so (include/althea_host.h, host/Althea/GeometryUtilities.h), a
    from-scratch generator held bit for bit to the reference's MikkTSpace build (tests/test_tangent_space.py);
  * skinned primitives are transformed on the host exactly as Gltf.vert:37-54 would (weights x joint matrices, bind pose: no
    animation is played) and handed over with an identity transform, because the producers take pre-transformed geometry.
"""
from __future__ import annotations

import ctypes as C
import io
import json
import os
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _capi

VERTEX_FLOATS = 26  # sizeof(Vertex) = 104 bytes (InstanceDataCommon.h:45-53); floats 24-25 hold the u16vec4 joints
WRAP_REPEAT, WRAP_CLAMP, WRAP_MIRROR = 0, 1, 2
MIP_NONE, MIP_NEAREST, MIP_LINEAR = 0, 1, 2


def sampler_word(wrap_u=WRAP_REPEAT, wrap_v=WRAP_REPEAT, mag_nearest=False, min_nearest=False, mip_mode=MIP_LINEAR, srgb=False) -> int:
    """ALTHEA_SAMPLER_WORD (include/althea_cuda.h): the SamplerOptions of Src/Sampler.cpp:11-88 in one word."""
    return wrap_u | (wrap_v << 2) | (int(mag_nearest) << 4) | (int(min_nearest) << 5) | (mip_mode << 6) | (int(srgb) << 8)


@dataclass
class TextureData:
    """RGBA8 texels with their mip chain (level k is max(1, w >> k) x max(1, h >> k)), as Src/Texture.cpp:47-99 creates them."""
    levels: List[np.ndarray]  # each (h, w, 4) uint8
    sampler: int = sampler_word()

    @property
    def width(self) -> int:
        return self.levels[0].shape[1]

    @property
    def height(self) -> int:
        return self.levels[0].shape[0]

    def packed(self) -> np.ndarray:
        return np.concatenate([np.ascontiguousarray(l, np.uint8).reshape(-1) for l in self.levels])

    @staticmethod
    def from_rgba8(rgba: np.ndarray, sampler: int, mips: bool = True) -> "TextureData":
        """Builds the mip chain the way Image.cpp:183-213 does (LINEAR 2:1 blits: a box of the 2x2 parents; sRGB images are
        averaged in linear light, as a blit between SRGB formats does)."""
        rgba = np.ascontiguousarray(rgba, np.uint8)
        levels = [rgba]
        use_mips = mips and ((sampler >> 6) & 3) != MIP_NONE
        srgb = bool(sampler & 0x100)
        cur = rgba
        while use_mips and (cur.shape[0] > 1 or cur.shape[1] > 1):
            h, w = cur.shape[:2]
            nh, nw = max(1, h >> 1), max(1, w >> 1)
            f = cur.astype(np.float64) / 255.0
            if srgb:
                f[..., :3] = np.where(f[..., :3] <= 0.04045, f[..., :3] / 12.92, ((f[..., :3] + 0.055) / 1.055) ** 2.4)
            # dst texel centre -> src coordinate, LINEAR filter (exactly the 2x2 average when the size halves evenly)
            ys = (np.arange(nh) + 0.5) * (h / nh) - 0.5
            xs = (np.arange(nw) + 0.5) * (w / nw) - 0.5
            y0 = np.clip(np.floor(ys).astype(int), 0, h - 1); y1 = np.clip(y0 + 1, 0, h - 1); fy = (ys - np.floor(ys))[:, None, None]
            x0 = np.clip(np.floor(xs).astype(int), 0, w - 1); x1 = np.clip(x0 + 1, 0, w - 1); fx = (xs - np.floor(xs))[None, :, None]
            top = f[y0][:, x0] * (1 - fx) + f[y0][:, x1] * fx
            bot = f[y1][:, x0] * (1 - fx) + f[y1][:, x1] * fx
            g = top * (1 - fy) + bot * fy
            if srgb:
                g[..., :3] = np.where(g[..., :3] <= 0.0031308, g[..., :3] * 12.92, 1.055 * np.maximum(g[..., :3], 0) ** (1 / 2.4) - 0.055)
            cur = np.clip(np.rint(g * 255.0), 0, 255).astype(np.uint8)
            levels.append(cur)
        return TextureData(levels, sampler)


@dataclass
class MaterialData:
    """MaterialConstants as Src/Material.cpp:14-110 fills them (defaults of a material-less primitive)."""
    baseColorFactor: tuple = (1.0, 1.0, 1.0, 1.0)
    baseTextureCoordinateIndex: int = 0
    metallicRoughnessTextureCoordinateIndex: int = 0
    normalScale: float = 1.0
    metallicFactor: float = 0.0
    roughnessFactor: float = 1.0
    alphaCutoff: float = 0.5
    baseTexture: Optional[TextureData] = None
    normalTexture: Optional[TextureData] = None
    metallicRoughnessTexture: Optional[TextureData] = None


@dataclass
class PrimitiveData:
    vertices: np.ndarray  # (n, 26) float32, the engine's Vertex
    indices: np.ndarray   # (3 t,) uint32
    model: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))  # row-major 4x4
    material: MaterialData = field(default_factory=MaterialData)
    front_face_clockwise: bool = False

    @property
    def triangle_count(self) -> int:
        return len(self.indices) // 3


def make_vertices(position, normal, uv0=None, tangent=None, bitangent=None) -> np.ndarray:
    n = len(position)
    v = np.zeros((n, VERTEX_FLOATS), np.float32)
    v[:, 0:3] = position
    v[:, 9:12] = normal
    if uv0 is not None:
        v[:, 12:14] = uv0
    if tangent is None:
        # any unit vector orthogonal to the normal
        nrm = v[:, 9:12]
        a = np.where(np.abs(nrm[:, 0:1]) > 0.9, np.array([[0.0, 1.0, 0.0]], np.float32), np.array([[1.0, 0.0, 0.0]], np.float32))
        t = np.cross(a, nrm)
        t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-20)
        tangent = t
        bitangent = np.cross(nrm, t)
    v[:, 3:6] = tangent
    v[:, 6:9] = bitangent
    return v


# ---- procedural meshes (tests, smoke) -----------------------------------------------------------------------------------
def uv_sphere(radius=1.0, centre=(0.0, 0.0, 0.0), stacks=12, slices=24) -> PrimitiveData:
    """Counter-clockwise (seen from outside) triangles, uv = (slice, stack) / counts."""
    ph = np.linspace(0.0, np.pi, stacks + 1)
    th = np.linspace(0.0, 2.0 * np.pi, slices + 1)
    P, T = np.meshgrid(ph, th, indexing="ij")
    n = np.stack([np.sin(P) * np.cos(T), np.cos(P), np.sin(P) * np.sin(T)], -1).reshape(-1, 3)
    pos = n * radius + np.asarray(centre)
    uv = np.stack([T / (2 * np.pi), P / np.pi], -1).reshape(-1, 2)
    tang = np.stack([-np.sin(T), np.zeros_like(T), np.cos(T)], -1).reshape(-1, 3)
    idx = []
    for i in range(stacks):
        for j in range(slices):
            a, b = i * (slices + 1) + j, i * (slices + 1) + j + 1
            c, d = a + slices + 1, b + slices + 1
            if i > 0:
                idx += [a, b, c]
            if i < stacks - 1:
                idx += [b, d, c]
    v = make_vertices(pos.astype(np.float32), n.astype(np.float32), uv.astype(np.float32), tang.astype(np.float32),
                      np.cross(n, tang).astype(np.float32))
    return PrimitiveData(v, np.asarray(idx, np.uint32))


def quad(corners, uv_scale=1.0) -> PrimitiveData:
    """Two triangles over 4 corners given counter-clockwise as seen from the front."""
    c = np.asarray(corners, np.float32)
    nrm = np.cross(c[1] - c[0], c[3] - c[0])
    nrm /= np.linalg.norm(nrm)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32) * uv_scale
    t = (c[1] - c[0]) / np.linalg.norm(c[1] - c[0])
    v = make_vertices(c, np.tile(nrm, (4, 1)), uv, np.tile(t, (4, 1)), np.tile(np.cross(nrm, t), (4, 1)))
    return PrimitiveData(v, np.array([0, 1, 2, 0, 2, 3], np.uint32))


def checker_texture(size=64, cells=8, a=(230, 230, 230, 255), b=(40, 60, 200, 255), sampler=None, alpha_holes=False) -> TextureData:
    y, x = np.mgrid[0:size, 0:size]
    m = ((x * cells // size) + (y * cells // size)) & 1
    img = np.where(m[..., None] == 0, np.array(a, np.uint8), np.array(b, np.uint8)).astype(np.uint8)
    if alpha_holes:
        img[..., 3] = np.where(m == 0, 255, 0)
    return TextureData.from_rgba8(img, sampler if sampler is not None else sampler_word(srgb=True))


# ---- GLB loader ---------------------------------------------------------------------------------------------------------
_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


class _Buffers:
    """The buffers of a glTF asset, loaded on first use: the GLB's BIN chunk (a buffer without uri), files next to the .gltf,
    or base64 data: URIs."""

    def __init__(self, gltf, base_dir, bin_chunk=b""):
        self.gltf, self.base_dir, self.bin_chunk, self.cache = gltf, base_dir, bin_chunk, {}

    def uri(self, uri: str) -> bytes:
        if uri.startswith("data:"):
            import base64
            return base64.b64decode(uri.split(",", 1)[1])
        from urllib.parse import unquote
        return open(os.path.join(self.base_dir, unquote(uri)), "rb").read()

    def __getitem__(self, i: int) -> bytes:
        if i not in self.cache:
            b = self.gltf["buffers"][i]
            self.cache[i] = self.uri(b["uri"]) if "uri" in b else self.bin_chunk
        return self.cache[i]

    def view(self, bv_index: int) -> bytes:
        bv = self.gltf["bufferViews"][bv_index]
        start = bv.get("byteOffset", 0)
        return self[bv.get("buffer", 0)][start: start + bv["byteLength"]]


def _accessor(gltf, buffers, idx):
    """Accessor -> (count, components) array. `buffers`: a _Buffers, or the bytes of buffer 0 (a GLB's BIN chunk)."""
    acc = gltf["accessors"][idx]
    dt = np.dtype(_COMPONENT[acc["componentType"]])
    nc = _NCOMP[acc["type"]]
    count = acc["count"]
    if "sparse" in acc:
        raise ValueError("sparse accessors are not supported")
    if "bufferView" not in acc:
        out = np.zeros((count, nc), dt)
    else:
        bv = gltf["bufferViews"][acc["bufferView"]]
        data = buffers if isinstance(buffers, (bytes, bytearray, memoryview)) else buffers[bv.get("buffer", 0)]
        start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * nc
        raw = np.frombuffer(data, np.uint8, count=(count - 1) * stride + dt.itemsize * nc, offset=start)
        out = np.lib.stride_tricks.as_strided(raw, shape=(count, dt.itemsize * nc), strides=(stride, 1)).copy().view(dt).reshape(count, nc)
    if acc.get("normalized") and dt != np.float32:
        out = np.maximum(out.astype(np.float32) / float(np.iinfo(dt).max), -1.0)
    return out


def _node_matrix(node) -> np.ndarray:
    if "matrix" in node:
        return np.asarray(node["matrix"], np.float64).reshape(4, 4).T
    m = np.eye(4)
    if "scale" in node:
        m = np.diag(list(node["scale"]) + [1.0]) @ m
    if "rotation" in node:
        x, y, z, w = node["rotation"]
        r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 0],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 0],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y), 0], [0, 0, 0, 1]])
        m = r @ m
    if "translation" in node:
        t = np.eye(4)
        t[:3, 3] = node["translation"]
        m = t @ m
    return m


def _host():
    """The host-side library (include/althea_host.h), bound in _hostapi.py."""
    from . import _hostapi
    return _hostapi.load()


def compute_flat_normals(pos) -> np.ndarray:
    """GeometryUtilities::computeFlatNormals (Include/Althea/GeometryUtilities.h:32-48) on a de-indexed triangle list."""
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    faces = len(pos) // 3
    out = np.zeros((faces * 3, 3), np.float32)
    if _host().althea_host_compute_flat_normals(pos.ctypes.data, faces, out.ctypes.data) != 0:
        raise RuntimeError("althea_host_compute_flat_normals failed")
    return out


def compute_tangent_space(pos, nrm, uv):
    """GeometryUtilities::computeTangentSpace (Include/Althea/GeometryUtilities.h:51-70,137-155): MikkTSpace-equivalent tangents
    and sign * cross(normal, tangent) bitangents for a de-indexed triangle list. Returns (tangent, bitangent), (3 * faces, 3)."""
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1, 3)
    uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    faces = len(pos) // 3
    if len(nrm) != len(pos) or len(uv) != len(pos):
        raise ValueError("position, normal and uv must have one row per corner")
    tang, bit = np.zeros((faces * 3, 3), np.float32), np.zeros((faces * 3, 3), np.float32)
    if _host().althea_host_compute_tangent_space(pos.ctypes.data, nrm.ctypes.data, uv.ctypes.data, faces, tang.ctypes.data,
                                                 bit.ctypes.data) != 0:
        raise RuntimeError("althea_host_compute_tangent_space failed")
    return tang, bit


def load_glb(path: str, max_texture_size: Optional[int] = None) -> List[PrimitiveData]:
    """Kept name of load_gltf (both containers go through it)."""
    return load_gltf(path, max_texture_size)


def load_gltf(path: str, max_texture_size: Optional[int] = None) -> List[PrimitiveData]:
    """glTF 2.0, binary (.glb) or JSON (.gltf with external / embedded buffers and images) -> the primitives Model::Model builds
    (Src/Model.cpp, Src/Primitive.cpp:51-367, Src/Material.cpp:14-110). Images are decoded with Pillow; the reference decodes
    them with stb_image through cesium-native (GltfReader.cpp:640-670): PNG is lossless either way, JPEG differs by the two
    decoders' IDCT / upsampling rounding (measured on DamagedHelmet's five 2048^2 JPEGs against the reference's stb build:
    at most 3 levels, 0.1-2 % of the texels off by one)."""
    from PIL import Image as PILImage
    data = open(path, "rb").read()
    bin_chunk = b""
    if data[:4] == b"glTF":
        magic, version, length = struct.unpack_from("<III", data, 0)
        off, gltf = 12, None
        while off < length:
            clen, ctype = struct.unpack_from("<II", data, off)
            chunk = data[off + 8: off + 8 + clen]
            if ctype == 0x4E4F534A:
                gltf = json.loads(chunk.decode("utf-8"))
            elif ctype == 0x004E4942:
                bin_chunk = chunk
            off += 8 + clen
        if gltf is None:
            raise ValueError("%s: GLB without a JSON chunk" % path)
    else:
        try:
            gltf = json.loads(data.decode("utf-8"))
        except (UnicodeDecodeError, json.JSONDecodeError):
            raise ValueError("%s is neither a GLB nor a glTF JSON file" % path)
        if "asset" not in gltf:
            raise ValueError("%s is not a glTF asset" % path)
    buffers = _Buffers(gltf, os.path.dirname(os.path.abspath(path)), bin_chunk)

    tex_cache = {}

    def texture(info, srgb):
        if info is None:
            return None
        key = (info["index"], srgb)
        if key in tex_cache:
            return tex_cache[key]
        tex = gltf["textures"][info["index"]]
        img = gltf["images"][tex["source"]]
        raw = buffers.view(img["bufferView"]) if "bufferView" in img else buffers.uri(img["uri"])
        pil = PILImage.open(io.BytesIO(raw)).convert("RGBA")
        if max_texture_size and max(pil.size) > max_texture_size:
            pil = pil.resize((max(1, pil.size[0] * max_texture_size // max(pil.size)), max(1, pil.size[1] * max_texture_size // max(pil.size))), PILImage.BILINEAR)
        smp = gltf.get("samplers", [{}])[tex["sampler"]] if "sampler" in tex else {}
        wrap = {33071: WRAP_CLAMP, 33648: WRAP_MIRROR, 10497: WRAP_REPEAT}
        wu = wrap.get(smp.get("wrapS", 10497), WRAP_CLAMP)
        wv = wrap.get(smp.get("wrapT", 10497), WRAP_CLAMP)
        minf = smp.get("minFilter", 9987)
        mip_mode = MIP_LINEAR if minf in (9987, 9986) else MIP_NEAREST if minf in (9984, 9985) else MIP_NONE
        word = sampler_word(wu, wv, smp.get("magFilter", 9729) == 9728, minf in (9728, 9984, 9986), mip_mode, srgb)
        t = TextureData.from_rgba8(np.asarray(pil, np.uint8), word)
        tex_cache[key] = t
        return t

    def material(idx) -> MaterialData:
        if idx is None:
            return MaterialData()
        m = gltf["materials"][idx]
        pbr = m.get("pbrMetallicRoughness", {})
        bt, mt, nt = pbr.get("baseColorTexture"), pbr.get("metallicRoughnessTexture"), m.get("normalTexture")
        return MaterialData(
            baseColorFactor=tuple(pbr.get("baseColorFactor", [1, 1, 1, 1])), baseTextureCoordinateIndex=(bt or {}).get("texCoord", 0),
            metallicRoughnessTextureCoordinateIndex=(mt or {}).get("texCoord", 0), normalScale=float((nt or {}).get("scale", 1.0)),
            metallicFactor=float(pbr.get("metallicFactor", 1.0)) if "pbrMetallicRoughness" in m else 0.0,
            roughnessFactor=float(pbr.get("roughnessFactor", 1.0)), alphaCutoff=float(m.get("alphaCutoff", 0.5)),
            baseTexture=texture(bt, True), normalTexture=texture(nt, False), metallicRoughnessTexture=texture(mt, False))

    out: List[PrimitiveData] = []

    # Node transforms as Model::_updateTransforms leaves them in the matrix buffer (Src/Model.cpp:341-356): global transform times
    # the node's inverse bind pose (identity unless a skin names the node as a joint, Model.cpp:193-216). No animation is applied,
    # so every node sits at the transform the file gives it.
    nodes = gltf.get("nodes", [])
    node_global = [None] * len(nodes)

    def globals_of(node_idx, parent):
        node_global[node_idx] = parent @ _node_matrix(nodes[node_idx])
        for c in nodes[node_idx].get("children", []):
            globals_of(c, node_global[node_idx])

    for n in gltf["scenes"][gltf.get("scene", 0)]["nodes"]:
        globals_of(n, np.eye(4))
    inverse_bind = [np.eye(4) for _ in nodes]
    for skin in gltf.get("skins", []):
        if "inverseBindMatrices" in skin:
            ibm = _accessor(gltf, buffers, skin["inverseBindMatrices"]).astype(np.float64).reshape(-1, 4, 4)
            for j, joint in enumerate(skin["joints"]):
                inverse_bind[joint] = ibm[j].T  # accessor matrices are column-major
    node_matrix = [None if g is None else (g @ inverse_bind[i]).astype(np.float32) for i, g in enumerate(node_global)]

    def skin_vertices(v, skin_idx, joints, weights):
        """Gltf.vert:37-54 for a skinned primitive, applied on the host (the producers take pre-transformed geometry):
        model = sum over the four influences with weight > 0 of weight * matrix[jointMap[joint]]; position and the TBN columns go
        through it; the primitive is then drawn with an identity transform. fp32, in the shader's order of operations."""
        joint_nodes = np.asarray(gltf["skins"][skin_idx]["joints"], np.int64)
        mats = np.stack([node_matrix[j] if node_matrix[j] is not None else np.eye(4, dtype=np.float32) for j in joint_nodes])
        model_m = np.zeros((len(v), 4, 4), np.float32)
        for i in range(4):
            w = weights[:, i].astype(np.float32)
            use = w > 0
            model_m[use] += w[use, None, None] * mats[joints[use, i].astype(np.int64)]
        p = v[:, 0:3]
        v[:, 0:3] = ((model_m[:, :3, 0] * p[:, 0:1] + model_m[:, :3, 1] * p[:, 1:2]) + model_m[:, :3, 2] * p[:, 2:3]) + model_m[:, :3, 3]
        for a in (3, 6, 9):  # tangent, bitangent, normal: mat3(model) * column
            c = v[:, a:a + 3].copy()
            v[:, a:a + 3] = (model_m[:, :3, 0] * c[:, 0:1] + model_m[:, :3, 1] * c[:, 1:2]) + model_m[:, :3, 2] * c[:, 2:3]
        return v

    def visit(node_idx, parent):
        node = gltf["nodes"][node_idx]
        world = parent @ _node_matrix(node)
        if "mesh" in node:
            for prim in gltf["meshes"][node["mesh"]]["primitives"]:
                if prim.get("mode", 4) != 4 or "POSITION" not in prim["attributes"]:
                    continue
                at = prim["attributes"]
                pos = _accessor(gltf, buffers, at["POSITION"]).astype(np.float32)
                nrm = _accessor(gltf, buffers, at["NORMAL"]).astype(np.float32) if "NORMAL" in at else None
                tan4 = _accessor(gltf, buffers, at["TANGENT"]).astype(np.float32) if "TANGENT" in at else None
                uvs = []  # Primitive.cpp:126-134: the sets stop at the first missing TEXCOORD_i (no compaction past a gap)
                for k in range(4):
                    if "TEXCOORD_%d" % k not in at:
                        break
                    uvs.append(_accessor(gltf, buffers, at["TEXCOORD_%d" % k]).astype(np.float32))
                idx = _accessor(gltf, buffers, prim["indices"]).reshape(-1).astype(np.uint32) if "indices" in prim else np.arange(len(pos), dtype=np.uint32)
                mat = material(prim.get("material"))
                had_normals = nrm is not None
                duplicate = nrm is None or tan4 is None  # Primitive.cpp:147: flat normals / generated tangents need unshared vertices
                idx = idx[: len(idx) // 3 * 3]
                src_index = idx.copy() if duplicate else None
                if duplicate:
                    pos, uvs = pos[idx], [u[idx] for u in uvs]
                    nrm = nrm[idx] if nrm is not None else None
                    tan4 = tan4[idx] if tan4 is not None else None
                    idx = np.arange(len(pos), dtype=np.uint32)
                    if nrm is None:  # Primitive.cpp:185-187
                        nrm = compute_flat_normals(pos)
                    if tan4 is None:  # Primitive.cpp:189-191, on the NORMAL MAP's uv set (Material.cpp:69)
                        nuv = int((gltf["materials"][prim["material"]].get("normalTexture") or {}).get("texCoord", 0)) if "material" in prim else 0
                        uvn = uvs[nuv] if len(uvs) > nuv else np.zeros((len(pos), 2), np.float32)
                        tang, bit = compute_tangent_space(pos, nrm, uvn)
                    elif had_normals:
                        tang = tan4[:, :3]
                        bit = tan4[:, 3:4] * np.cross(nrm, tang)
                    else:  # TANGENT without NORMAL: Primitive.cpp:336-345 reads tangents only inside `if (hasNormals)`, and
                        tang = np.zeros_like(pos)  # :189 generates none either: both stay zero-initialised
                        bit = np.zeros_like(pos)
                else:
                    tang = tan4[:, :3]
                    bit = tan4[:, 3:4] * np.cross(nrm, tang)  # Primitive.cpp:341-343
                v = np.zeros((len(pos), VERTEX_FLOATS), np.float32)
                v[:, 0:3], v[:, 3:6], v[:, 6:9], v[:, 9:12] = pos, tang, bit, nrm
                for k, u in enumerate(uvs[:4]):
                    v[:, 12 + 2 * k: 14 + 2 * k] = u
                # Model.cpp:349: the matrix a primitive is drawn with is global * inverseBindPose of ITS node (identity unless the node is
                # also a joint of some skin)
                prim_world = node_matrix[node_idx] if node_matrix[node_idx] is not None else world.astype(np.float32)
                if "JOINTS_0" in at and "WEIGHTS_0" in at and "skin" in node:  # Primitive.cpp:320: isSkinned
                    joints = _accessor(gltf, buffers, at["JOINTS_0"])
                    wacc = gltf["accessors"][at["WEIGHTS_0"]]
                    weights = _accessor(gltf, buffers, at["WEIGHTS_0"]).astype(np.float32)
                    if wacc["componentType"] in (5121, 5123) and not wacc.get("normalized"):
                        # Primitive.cpp:354-359 divides u8 / u16 weights by 255 / 65535 whether or not the accessor says `normalized`
                        weights = weights / np.float32(255.0 if wacc["componentType"] == 5121 else 65535.0)
                    if src_index is not None:
                        joints, weights = joints[src_index], weights[src_index]
                    v[:, 20:24] = weights
                    v[:, 24:26] = np.ascontiguousarray(joints.astype(np.uint16)).view(np.float32)
                    v = skin_vertices(v, node["skin"], joints, weights)
                    prim_world = np.eye(4, dtype=np.float32)
                out.append(PrimitiveData(v, idx, prim_world, mat, False))
        for c in node.get("children", []):
            visit(c, world)

    scene = gltf["scenes"][gltf.get("scene", 0)]
    for n in scene["nodes"]:
        visit(n, np.eye(4))
    return out


# ---- C-ABI structs ------------------------------------------------------------------------------------------------------
def point_light_constants() -> _capi.PointLightConstants:
    """PointLightConstants as the PointLightCollection constructor builds them (Src/PointLight.cpp:72-118): a 90 degree, aspect 1,
    near 0.01 / far 1000 Camera at the origin turned to the six axes, views = computeView(), inverses by glm::inverse. Computed by
    the C++ mirror of the reference's Camera (host/Althea/Camera.h through libalthea_host.so), which is bit-identical to the
    reference's class built against its GLM, including the rounding noise of sin/cos at 180 and +-90 degrees that orients the
    +-Y faces (tests/test_camera_pin.py)."""
    lib = _host()
    pc = _capi.PointLightConstants()
    assert C.sizeof(pc) == 14 * 64
    if lib.althea_host_point_light_constants(C.addressof(pc)) != 0:
        raise RuntimeError("althea_host_point_light_constants failed")
    return pc


def camera_matrices(fov_degrees: float, aspect: float, near: float, far: float, position, yaw: float, pitch: float):
    """The reference's Camera (Src/Camera.cpp:7-110) for one pose: (projection, transform, view, inverse projection), each a
    (4, 4) float32 array in glm's column-major storage (row i of the array is column i of the matrix)."""
    lib = _host()
    pos = np.ascontiguousarray(position, np.float32)
    out = [np.zeros((4, 4), np.float32) for _ in range(4)]
    if lib.althea_host_camera(fov_degrees, aspect, near, far, pos.ctypes.data, yaw, pitch, *[o.ctypes.data for o in out]) != 0:
        raise RuntimeError("althea_host_camera failed")
    return tuple(out)


class UploadedModel:
    """Primitives resident on the device: what a Model holds after construction (vertex / index buffers, textures)."""

    def __init__(self, ctx, prims: List[PrimitiveData]):
        import torch
        self.ctx = ctx
        self.prims = prims
        self._keep = []
        dev = "cuda:%d" % ctx.device
        self.array = (_capi.Primitive * max(1, len(prims)))()
        tex_handles = {}

        def tex_ref(t: Optional[TextureData]) -> _capi.TextureRef:
            r = _capi.TextureRef()
            if t is None:
                return r
            if id(t) not in tex_handles:
                packed = torch.from_numpy(t.packed()).to(dev)
                tex_handles[id(t)] = ctx.wrap_tensor(packed, _capi.FORMAT_R8G8B8A8_UNORM, t.width, t.height, len(t.levels), 1)
            r.image = tex_handles[id(t)].handle
            r.sampler = t.sampler
            return r

        for i, p in enumerate(prims):
            vb = ctx.wrap_buffer(torch.from_numpy(np.ascontiguousarray(p.vertices, np.float32).reshape(-1)).to(dev))
            ib = ctx.wrap_buffer(torch.from_numpy(np.ascontiguousarray(p.indices, np.uint32).view(np.int32)).to(dev))
            self._keep += [vb, ib]
            a = self.array[i]
            a.vertices, a.indices = vb.handle, ib.handle
            a.index_count = len(p.indices)
            a.front_face_clockwise = int(p.front_face_clockwise)
            col_major = np.asarray(p.model, np.float32).T.reshape(-1)
            for k in range(16):
                a.model[k] = float(col_major[k])
            m = p.material
            for k in range(4):
                a.material.baseColorFactor[k] = float(m.baseColorFactor[k])
            a.material.baseTextureCoordinateIndex = m.baseTextureCoordinateIndex
            a.material.metallicRoughnessTextureCoordinateIndex = m.metallicRoughnessTextureCoordinateIndex
            a.material.normalScale, a.material.metallicFactor = m.normalScale, m.metallicFactor
            a.material.roughnessFactor, a.material.alphaCutoff = m.roughnessFactor, m.alphaCutoff
            a.material.baseTexture = tex_ref(m.baseTexture)
            a.material.normalTexture = tex_ref(m.normalTexture)
            a.material.metallicRoughnessTexture = tex_ref(m.metallicRoughnessTexture)
        self._keep += list(tex_handles.values())
        self.count = len(prims)

    @property
    def triangle_count(self) -> int:
        return sum(p.triangle_count for p in self.prims)
