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
