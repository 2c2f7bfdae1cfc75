"""ctypes binding of include/althea_host.h (althea_b200/lib/libalthea_host.so): the host-side, CUDA-free preparation the
reference does on the CPU as well: tangent spaces and flat normals, Radiance .hdr IO, Camera / PointLightConstants matrices.
One entry per symbol of the header; tests/test_tangent_space.py checks header, table and exports against each other."""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 1

_P, _F, _I32P = C.c_void_p, C.c_float, C.POINTER(C.c_int32)
SIGNATURES = {
    "althea_host_abi_version": (C.c_int, []),
    "althea_host_compute_flat_normals": (C.c_int, [_P, C.c_uint64, _P]),
    "althea_host_compute_tangent_space": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P]),
    "althea_host_save_hdri": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, _P]),
    "althea_host_save_exr": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, _P]),
    "althea_host_load_hdri_info": (C.c_int, [C.c_char_p, _I32P, _I32P]),
    "althea_host_load_hdri": (C.c_int, [C.c_char_p, _P, C.c_uint64]),
    "althea_host_camera": (C.c_int, [_F, _F, _F, _F, _P, _F, _F, _P, _P, _P, _P]),
    "althea_host_point_light_constants": (C.c_int, [_P]),
}

_lib = None


def load():
    """Loads (building on first use: plain g++) the library and binds every symbol. A missing compiler or a stale ABI is an
    error, never a silent fallback."""
    global _lib
    if _lib is None:
        from .host import build_host
        lib = C.CDLL(build_host.build_lib())
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.althea_host_abi_version() != ABI_VERSION:
            raise RuntimeError("libalthea_host.so: ABI version %d, binding expects %d" % (lib.althea_host_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib
