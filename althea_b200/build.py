"""Builds althea_b200/lib/libalthea_cuda.so with nvcc for sm_100a, in-tree (the .so travels to the GPU box).

frame_kernels.cu is compiled twice: the fast build (FFMA contraction) and the parity build (-fmad=false, IEEE
division/sqrt) that bit-matches the CPU oracle's operation order; both live in the one shared library and the context
flag ALTHEA_CTX_PARITY_MATH picks between them at run time.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ_DIR = os.path.join(_HERE, "build")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libalthea_cuda.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "--expt-relaxed-constexpr"]

UNITS = [
    # (source, object, extra flags)
    ("frame_kernels.cu", "frame_fast.o", ["-DALTHEA_NS=althea_fast"]),
    ("frame_kernels.cu", "frame_parity.o", ["-DALTHEA_NS=althea_parity", "-DALTHEA_PARITY", "-fmad=false"]),
    ("ibl_kernels.cu", "ibl_kernels.o", []),
    ("raster_kernels.cu", "raster_kernels.o", []),
    ("althea_cuda.cu", "althea_cuda.o", []),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the sm_100a kernels")


def _sources_digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(_HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    with open(os.path.abspath(__file__), "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def is_stale() -> bool:
    stamp = os.path.join(LIB_DIR, ".digest")
    if not (os.path.exists(LIB_PATH) and os.path.exists(stamp)):
        return True
    return open(stamp).read().strip() != _sources_digest()


def build_variant(name: str, extra_flags) -> str:
    """A/B builds for kernel tuning: lib/libalthea_cuda_<name>.so with extra -D flags; selected at run time with the
    ALTHEA_CUDA_LIB environment variable (see _capi.library_path). Not part of the default build."""
    nvcc = _nvcc()
    obj_dir = os.path.join(OBJ_DIR, name)
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    objs, procs = [], []
    for src, obj, extra in UNITS:
        out = os.path.join(obj_dir, obj)
        cmd = [nvcc, *ARCH, *COMMON, *extra, *extra_flags, "-c", os.path.join(CSRC, src), "-o", out]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(out)
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + out)
    lib = os.path.join(LIB_DIR, "libalthea_cuda_%s.so" % name)
    subprocess.check_call([nvcc, *ARCH, "-shared", "-o", lib, *objs, "-cudart", "static"])
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    objs = []
    procs = []
    for src, obj, extra in UNITS:
        out = os.path.join(OBJ_DIR, obj)
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", out]
        procs.append((cmd, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(out)
    log = []
    for cmd, p in procs:
        out, _ = p.communicate()
        log.append("$ " + " ".join(cmd) + "\n" + out)
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError("nvcc failed for " + cmd[-3])
    link = [nvcc, *ARCH, "-shared", "-o", LIB_PATH, *objs, "-cudart", "static"]
    r = subprocess.run(link, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + " ".join(link) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("link failed")
    with open(os.path.join(OBJ_DIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    with open(os.path.join(LIB_DIR, ".digest"), "w") as f:
        f.write(_sources_digest())
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
