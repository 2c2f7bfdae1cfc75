"""Builds althea_b200/host/demo_frame (g++, C++17) against include/althea_cuda.h and the in-tree libalthea_cuda.so."""
from __future__ import annotations

import os
import subprocess

from .. import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
BIN = os.path.join(_HERE, "demo_frame")


def build(force: bool = False) -> str:
    lib = _build.build()
    srcs = [os.path.join(_HERE, "demo_frame.cpp")] + [os.path.join(_HERE, "Althea", f) for f in sorted(os.listdir(os.path.join(_HERE, "Althea")))]
    if not force and os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(p) for p in srcs + [lib]):
        return BIN
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-I", _HERE,
           os.path.join(_HERE, "demo_frame.cpp"), "-L", _build.LIB_DIR, "-lalthea_cuda", "-Wl,-rpath,$ORIGIN/../lib", "-ldl", "-lpthread", "-lrt",
           "-o", BIN]
    subprocess.check_call(cmd)
    return BIN


HOST_LIB = os.path.join(_build.LIB_DIR, "libalthea_host.so")


def build_lib(force: bool = False) -> str:
    """althea_b200/lib/libalthea_host.so: the CPU-side geometry preparation (include/althea_host.h). Plain g++, no CUDA;
    fp contraction off so the fp32 arithmetic is the same on every host."""
    srcs = [os.path.join(_HERE, "host_abi.cpp"), os.path.join(_HERE, "Althea", "GeometryUtilities.h"),
            os.path.join(_HERE, "Althea", "Utilities.h"), os.path.join(_HERE, "Althea", "Camera.h"),
            os.path.join(ROOT, "include", "althea_host.h")]
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(HOST_LIB) >= os.path.getmtime(p) for p in srcs):
        return HOST_LIB
    os.makedirs(_build.LIB_DIR, exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra",
           "-I", os.path.join(ROOT, "include"), "-I", _HERE, srcs[0], "-o", HOST_LIB]
    subprocess.check_call(cmd)
    return HOST_LIB
