"""Generates tests/golden/stb_written.hdr + stb_written.npz. Run in the BUILD container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_hdr_golden.py

A small float image with every feature of the container (runs, literals longer than 128, tiny values, zero texels) is written by
the REFERENCE's stbi_write_hdr (oracle/_ref/libstb_ref.so, compiled from /root/reference/Extern/stb where it lies) and decoded
by its stbi_loadf; the input floats, the file and the decoded floats are committed so the codec stays pinned where the reference
is not mounted."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def golden_image():
    rs = np.random.default_rng(11)
    h, w = 24, 300
    img = np.exp(rs.uniform(-12, 9, (h, w, 4))).astype(np.float32)
    img[:, 40:200] = img[:, 40:41]          # long runs (> 127)
    img[::3, 200:] = 0.0                    # zero texels
    img[1, 1, :3] = (1e-35, 0.0, 0.0)       # below the 1e-32 cut-off
    img[2, 2, :3] = (1000.0, 1e-3, 1.0)     # shared exponent truncates the small channels
    img[5, :, :3] = img[5, :1, :3]          # a whole constant row
    return img


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libstb_ref.so"))
    ref.ref_stbi_loadf_from_memory.restype = C.POINTER(C.c_float)
    img = golden_image()
    path = os.path.join(HERE, "stb_written.hdr")
    assert ref.ref_stbi_write_hdr(path.encode(), img.shape[1], img.shape[0], img.ctypes.data_as(C.c_void_p)) == 1
    buf = open(path, "rb").read()
    w, h = C.c_int(), C.c_int()
    p = ref.ref_stbi_loadf_from_memory(buf, len(buf), C.byref(w), C.byref(h))
    decoded = np.ctypeslib.as_array(p, (h.value, w.value, 4)).copy()
    ref.ref_stbi_free(p)
    np.savez_compressed(os.path.join(HERE, "stb_written.npz"), image=img, decoded=decoded)
    print(len(buf), decoded.shape)


if __name__ == "__main__":
    main()
