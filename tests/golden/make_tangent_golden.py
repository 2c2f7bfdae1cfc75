"""Generates tests/golden/tangent_space.npz. Run in the BUILD container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_tangent_golden.py

Runs the REFERENCE's tangent-space generator (Extern/MikkTSpace/mikktspace.c, compiled where it lies into
oracle/_ref/libmikktspace_ref.so and fed the way Include/Althea/GeometryUtilities.h:51-155 feeds it) on the meshes of
tests/tangent_cases.py and on the reference's own DamagedHelmet primitive, and stores its answers: tangent and sign per corner
for the small cases, CRCs for the helmet (whose mesh is already in tests/golden/helmet.npz)."""
import ctypes as C
import os
import subprocess
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tangent_cases import CASES  # noqa: E402


def reference_tangents(pos, nrm, uv):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmikktspace_ref.so"))
    faces = len(pos) // 3
    tang, sign = np.zeros((3 * faces, 3), np.float32), np.zeros(3 * faces, np.float32)
    ok = ref.ref_mikktspace(pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), faces,
                            tang.ctypes.data_as(C.c_void_p), sign.ctypes.data_as(C.c_void_p))
    assert ok == 1
    return tang, sign


def helmet_soup():
    d = np.load(os.path.join(HERE, "helmet.npz"))
    idx = d["idx"].astype(np.uint32)
    pos, nrm, uv = d["pos"][idx], d["nrm"].astype(np.float32)[idx], d["uv"][idx]
    nrm = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    return np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nrm, np.float32), np.ascontiguousarray(uv, np.float32)


def main():
    out = {}
    for name, make in CASES.items():
        pos, nrm, uv = make()
        tang, sign = reference_tangents(pos, nrm, uv)
        out[name + "_tangent"], out[name + "_sign"] = tang, sign.astype(np.int8)
        out[name + "_input_crc"] = np.array([zlib.crc32(pos.tobytes() + nrm.tobytes() + uv.tobytes())], np.uint32)
    pos, nrm, uv = helmet_soup()
    tang, sign = reference_tangents(pos, nrm, uv)
    out["helmet_tangent_crc"] = np.array([zlib.crc32(tang.tobytes())], np.uint32)
    out["helmet_sign_crc"] = np.array([zlib.crc32(sign.tobytes())], np.uint32)
    out["helmet_thumb"] = tang[::97].copy()
    np.savez_compressed(os.path.join(HERE, "tangent_space.npz"), **out)
    print({k: (v.shape if v.ndim > 1 else v[:1]) for k, v in out.items()})


if __name__ == "__main__":
    main()
