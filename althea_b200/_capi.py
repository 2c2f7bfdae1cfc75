"""ctypes binding of libalthea_cuda.so (the C ABI in include/althea_cuda.h).

Fails loudly when the library is missing: there is no CPU or PyTorch fallback for any stage.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

# VkFormat values (include/althea_cuda.h)
FORMAT_R8_UINT = 13
FORMAT_R8G8B8A8_UNORM = 37
FORMAT_R16G16B16A16_SFLOAT = 97
FORMAT_R32_SFLOAT = 100
FORMAT_R32G32B32A32_SFLOAT = 109
FORMAT_D32_SFLOAT = 126

BYTES_PER_TEXEL = {FORMAT_R8_UINT: 1, FORMAT_R8G8B8A8_UNORM: 4, FORMAT_R16G16B16A16_SFLOAT: 8, FORMAT_R32_SFLOAT: 4,
                   FORMAT_R32G32B32A32_SFLOAT: 16, FORMAT_D32_SFLOAT: 4}

IMAGE_CUBE = 1
IMAGE_OPTIMAL_TILING = 2
CTX_PARITY_MATH = 1
CTX_SSAO_EXACT_TAPS = 2
CTX_SSAO_COUNT_TAPS = 4
CTX_SSAO_RAY_DEPTH_PROXY = 8
CTX_SSAO_NO_CULL = 16
CTX_SSR_PLANE_SKIP = 32
CTX_BAND_EXCHANGE_HALO = 64
SHADE_SKIP_TONEMAP = 1
SHADE_NO_SSAO = 2
SHADE_AO_ONLY = 8
SHADE_AO_FROM_IMAGE = 4
IBL_LAYOUT_EQUIRECT, IBL_LAYOUT_CUBE = 0, 1
IBL_SEQ_REFERENCE_HASH, IBL_SEQ_HAMMERSLEY = 0, 1


class GlobalUniforms(C.Structure):
    """althea_global_uniforms == Include/Althea/GlobalUniforms.h:15-31 (416 bytes)."""
    _fields_ = [
        ("projection", C.c_float * 16), ("inverseProjection", C.c_float * 16), ("view", C.c_float * 16),
        ("prevView", C.c_float * 16), ("inverseView", C.c_float * 16), ("prevInverseView", C.c_float * 16),
        ("mouseUV", C.c_float * 2), ("lightCount", C.c_int32), ("lightBufferHandle", C.c_uint32),
        ("time", C.c_float), ("exposure", C.c_float), ("inputMask", C.c_uint32), ("frameCount", C.c_uint32),
    ]


class GBuffer(C.Structure):
    _fields_ = [("depth", C.c_uint64), ("position", C.c_uint64), ("normal", C.c_uint64), ("albedo", C.c_uint64),
                ("mro", C.c_uint64)]


class IBL(C.Structure):
    _fields_ = [("env", C.c_uint64), ("prefiltered", C.c_uint64), ("irradiance", C.c_uint64), ("brdf_lut", C.c_uint64)]


class Sync(C.Structure):
    _fields_ = [("wait_sem", C.c_uint64), ("wait_value", C.c_uint64), ("signal_sem", C.c_uint64),
                ("signal_value", C.c_uint64), ("cuda_stream", C.c_void_p)]


class IblPrecomputeDesc(C.Structure):
    _fields_ = [("layout", C.c_uint32), ("sequence", C.c_uint32), ("prefilter_samples", C.c_uint32),
                ("theta_samples", C.c_uint32)]


assert C.sizeof(GlobalUniforms) == 416


class TextureRef(C.Structure):
    _fields_ = [("image", C.c_uint64), ("sampler", C.c_uint32), ("_pad", C.c_uint32)]


class Material(C.Structure):
    """althea_material: the MaterialConstants fields fetchMaterial reads (InstanceDataCommon.h:8-33)."""
    _fields_ = [("baseColorFactor", C.c_float * 4), ("baseTextureCoordinateIndex", C.c_int32),
                ("metallicRoughnessTextureCoordinateIndex", C.c_int32), ("normalScale", C.c_float), ("metallicFactor", C.c_float),
                ("roughnessFactor", C.c_float), ("alphaCutoff", C.c_float), ("baseTexture", TextureRef), ("normalTexture", TextureRef),
                ("metallicRoughnessTexture", TextureRef)]


class Primitive(C.Structure):
    _fields_ = [("vertices", C.c_uint64), ("indices", C.c_uint64), ("index_count", C.c_uint32), ("front_face_clockwise", C.c_uint32),
                ("model", C.c_float * 16), ("material", Material)]


class PointLightConstants(C.Structure):
    """althea_point_light_constants == PointLightConstants (Src/PointLight.cpp:72-118)."""
    _fields_ = [("projection", C.c_float * 16), ("inverseProjection", C.c_float * 16), ("views", (C.c_float * 16) * 6),
                ("inverseViews", (C.c_float * 16) * 6)]


VERTEX_BYTES = 104  # sizeof(althea_vertex)

# every symbol include/althea_cuda.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = {
    "althea_cuda_abi_version": (C.c_int, []),
    "althea_cuda_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "althea_cuda_destroy": (None, [C.c_void_p]),
    "althea_cuda_last_error": (C.c_char_p, [C.c_void_p]),
    "althea_cuda_set_flags": (C.c_int, [C.c_void_p, C.c_uint32]),
    "althea_cuda_set_scissor_rows": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "althea_cuda_band_rows": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "althea_cuda_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "althea_cuda_get_timings": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_int]),
    "althea_cuda_reset_timings": (C.c_int, [C.c_void_p]),
    "althea_cuda_launch_count": (C.c_uint64, [C.c_void_p]),
    "althea_cuda_diag_ssao_gathers": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "althea_cuda_diag_ssao_exact_fallbacks": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "althea_cuda_diag_ssao_cull": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "althea_cuda_diag_gather_ceiling": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]),
    "althea_cuda_image_bytes": (C.c_size_t, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "althea_cuda_import_image": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]),
    "althea_cuda_import_buffer": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "althea_cuda_import_semaphore": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "althea_cuda_wrap_linear_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "althea_cuda_wrap_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]),
    "althea_cuda_create_image": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                           C.POINTER(C.c_uint64)]),
    "althea_cuda_create_buffer": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]),
    "althea_cuda_upload": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "althea_cuda_download": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "althea_cuda_device_pointer": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "althea_cuda_release": (C.c_int, [C.c_void_p, C.c_uint64]),
    "althea_cuda_synchronize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "althea_cuda_ssr_capture": (C.c_int, [C.c_void_p, C.POINTER(GlobalUniforms), C.POINTER(GBuffer), C.POINTER(IBL), C.c_uint64,
                                          C.c_uint64, C.c_uint64, C.POINTER(Sync)]),
    "althea_cuda_glossy_convolve": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(Sync)]),
    "althea_cuda_deferred_shade": (C.c_int, [C.c_void_p, C.POINTER(GlobalUniforms), C.POINTER(GBuffer), C.POINTER(IBL), C.c_uint64,
                                             C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(Sync)]),
    "althea_cuda_draw_gbuffer": (C.c_int, [C.c_void_p, C.POINTER(GlobalUniforms), C.POINTER(Primitive), C.c_uint32, C.POINTER(GBuffer), C.POINTER(Sync)]),
    "althea_cuda_draw_shadow_cubes": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(PointLightConstants), C.POINTER(Primitive), C.c_uint32,
                                                C.c_uint64, C.POINTER(Sync)]),
    "althea_cuda_generate_mips": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(Sync)]),
    "althea_cuda_ibl_precompute": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(IblPrecomputeDesc), C.c_uint64, C.c_uint64,
                                             C.POINTER(Sync)]),
    "althea_cuda_brdf_lut": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(Sync)]),
}

_lib = None


def library_path() -> str:
    """The in-tree library; ALTHEA_CUDA_LIB names an A/B variant built by build.build_variant (kernel tuning only)."""
    return os.environ.get("ALTHEA_CUDA_LIB") or _build.LIB_PATH


def load() -> C.CDLL:
    """Loads the in-tree shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "althea_b200: %s is missing. Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU fallback." % path)
        lib = C.CDLL(path)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib
