"""BASELINE configs[0]'s asset as a test fixture: rebuilds the DamagedHelmet primitive (engine Vertex buffer, material, textures)
from tests/golden/helmet.npz (written by tests/golden/make_helmet_golden.py)."""
from __future__ import annotations

import os

import numpy as np

from althea_b200 import model, scene

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "helmet.npz")
SIZE = (1280, 720)


def make_camera(W, H):
    """The view the fixture was generated with (only make_helmet_golden.py calls this; tests read the stored block)."""
    return scene.make_uniforms(W, H, pos=(0.6, 0.35, 2.4), yaw=0.25, pitch=-0.12, light_count=0)


def CAMERA(path: str = GOLDEN):
    """The GlobalUniforms block stored in the fixture (bytes, so the matrices are the same on every machine)."""
    import ctypes as C

    from althea_b200._capi import GlobalUniforms
    g = GlobalUniforms()
    raw = np.load(path)["uniforms"].tobytes()
    C.memmove(C.addressof(g), raw, C.sizeof(g))
    return g


def helmet_primitives(path: str = GOLDEN):
    """The single primitive as Primitive.cpp builds it for a mesh without tangents: vertices de-indexed, tangents generated."""
    d = np.load(path)
    idx = d["idx"].astype(np.uint32)
    pos, nrm, uv = d["pos"][idx], d["nrm"].astype(np.float32)[idx], d["uv"][idx]
    nrm = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    tang, bit = model.compute_tangent_space(pos, nrm, uv)  # GeometryUtilities::computeTangentSpace, as Primitive.cpp:189-191
    v = np.zeros((len(pos), model.VERTEX_FLOATS), np.float32)
    v[:, 0:3], v[:, 3:6], v[:, 6:9], v[:, 9:12], v[:, 12:14] = pos, tang, bit, nrm, uv
    f = d["factors"]
    mat = model.MaterialData(
        baseColorFactor=tuple(float(x) for x in f[:4]), normalScale=float(f[4]), metallicFactor=float(f[5]), roughnessFactor=float(f[6]),
        alphaCutoff=float(f[7]), baseTexture=model.TextureData.from_rgba8(d["base"], model.sampler_word(srgb=True)),
        normalTexture=model.TextureData.from_rgba8(d["normal"], model.sampler_word()),
        metallicRoughnessTexture=model.TextureData.from_rgba8(d["mr"], model.sampler_word()))
    return [model.PrimitiveData(v, np.arange(len(pos), dtype=np.uint32), d["model"].astype(np.float32), mat, False)]
