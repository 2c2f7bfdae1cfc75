"""Generates tests/golden/helmet.npz. Run in the BUILD container (needs /root/reference):

    python tests/golden/make_helmet_golden.py

BASELINE configs[0] is "DamagedHelmet glTF ... one 1280x720 deferred frame". The asset itself (3.7 MB GLB with 2048^2 JPEGs)
stays in the reference tree; this fixture keeps what the G-buffer pass needs to render it on the GPU box: the indexed mesh of
its single primitive (positions, normals, uv0, indices: Content/Models/DamagedHelmet.glb, Khronos sample asset, CC-BY 4.0),
its node transform and material factors, its three material textures downsampled to 256^2, plus the restatement's answer for the
1280x720 frame (CRC of the depth image, coverage, a 1/8-scale depth thumbnail) so the GPU test has a committed vector as well as
the live oracle."""
from __future__ import annotations

import json
import os
import struct
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from althea_b200 import model  # noqa: E402
from helmet_fixture import SIZE, helmet_primitives, make_camera  # noqa: E402
from oracle import oracle as O  # noqa: E402

GLB = "/root/reference/Content/Models/DamagedHelmet.glb"


def main():
    data = open(GLB, "rb").read()
    clen, _ = struct.unpack_from("<II", data, 12)
    gltf = json.loads(data[20:20 + clen].decode())
    bin_chunk = data[20 + clen + 8:]
    prim = gltf["meshes"][0]["primitives"][0]
    at = prim["attributes"]
    pos = model._accessor(gltf, bin_chunk, at["POSITION"]).astype(np.float32)
    nrm = model._accessor(gltf, bin_chunk, at["NORMAL"]).astype(np.float32)
    uv = model._accessor(gltf, bin_chunk, at["TEXCOORD_0"]).astype(np.float32)
    idx = model._accessor(gltf, bin_chunk, prim["indices"]).reshape(-1).astype(np.uint16)
    assert "TANGENT" not in at and len(idx) == 46356 and len(pos) == 14556
    full = model.load_glb(GLB, max_texture_size=256)
    assert len(full) == 1
    m = full[0].material
    out = os.path.join(HERE, "helmet.npz")
    np.savez_compressed(
        out, pos=pos, nrm=nrm.astype(np.float16), uv=uv, idx=idx, model=full[0].model.astype(np.float32),
        base=m.baseTexture.levels[0], normal=m.normalTexture.levels[0], mr=m.metallicRoughnessTexture.levels[0],
        factors=np.array(list(m.baseColorFactor) + [m.normalScale, m.metallicFactor, m.roughnessFactor, m.alphaCutoff], np.float32))
    # the restatement's answer at the configs[0] size
    prims = helmet_primitives(out)
    W, H = SIZE
    g = make_camera(W, H)
    want = O.draw_gbuffer(list(g.projection), list(g.view), prims, W, H)
    cov = want["tri"] != 0xFFFFFFFF
    import ctypes as C
    golden = dict(np.load(out))
    golden["uniforms"] = np.frombuffer(C.string_at(C.addressof(g), C.sizeof(g)), np.uint8).copy()
    golden["depth_crc"] = np.array([zlib.crc32(want["depth"].tobytes())], np.uint32)
    golden["tri_crc"] = np.array([zlib.crc32(want["tri"].tobytes())], np.uint32)
    golden["coverage"] = np.array([int(cov.sum())], np.int64)
    golden["depth_thumb"] = want["depth"][4::8, 4::8].copy()
    golden["albedo_thumb"] = want["albedo"][4::8, 4::8].copy()
    np.savez_compressed(out, **golden)
    print("helmet.npz", os.path.getsize(out), "bytes; coverage", cov.mean(), "depth crc", hex(int(golden["depth_crc"][0])))


if __name__ == "__main__":
    main()
