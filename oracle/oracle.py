"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Nothing under althea_b200/ does: the product path has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

SKIP_TONEMAP = 1
NO_SSAO = 2
FMT_RGBA32F, FMT_RGBA16F, FMT_RGBA8, FMT_R32F = 0, 1, 2, 3
ADDR_CLAMP, ADDR_REPEAT = 0, 1
LAYOUT_EQUIRECT, LAYOUT_CUBE = 0, 1
SEQ_HASH, SEQ_HAMMERSLEY = 0, 1


def build(force: bool = False) -> str:
    """Compiles the restatement with oracle/Makefile (gcc only; building the checker is not using it)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("althea_oracle.cpp", "althea_oracle_ibl.cpp", "althea_oracle_raster.cpp", "oracle_math.h", "Makefile")
    ):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def build_ref():
    """oracle/_ref/libmikktspace_ref.so: the reference's vendored tangent-space generator, compiled from the sources where they
    lie (`make ref`). Only possible where /root/reference is mounted; elsewhere the prebuilt file (if it travelled) is used.
    Returns the path or None."""
    out = os.path.join(_HERE, "_ref", "libmikktspace_ref.so")
    if os.path.isfile("/root/reference/Extern/MikkTSpace/mikktspace.c"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return out if os.path.exists(out) else None


class GlobalUniforms(C.Structure):
    """Include/Althea/GlobalUniforms.h:15-31 (416 bytes)."""
    _fields_ = [
        ("projection", C.c_float * 16), ("inverseProjection", C.c_float * 16), ("view", C.c_float * 16),
        ("prevView", C.c_float * 16), ("inverseView", C.c_float * 16), ("prevInverseView", C.c_float * 16),
        ("mouseUV", C.c_float * 2), ("lightCount", C.c_int32), ("lightBufferHandle", C.c_uint32),
        ("time", C.c_float), ("exposure", C.c_float), ("inputMask", C.c_uint32), ("frameCount", C.c_uint32),
    ]


assert C.sizeof(GlobalUniforms) == 416


class _GBuffer(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("position", C.c_void_p), ("depth", C.c_void_p),
                ("normal", C.c_void_p), ("albedo", C.c_void_p), ("mro", C.c_void_p)]


class _IBL(C.Structure):
    _fields_ = [("env", C.c_void_p), ("envW", C.c_int32), ("envH", C.c_int32),
                ("prefiltered", C.c_void_p), ("preW", C.c_int32), ("preH", C.c_int32), ("preMips", C.c_int32),
                ("irradiance", C.c_void_p), ("irrW", C.c_int32), ("irrH", C.c_int32),
                ("lut", C.c_void_p), ("lutW", C.c_int32), ("lutH", C.c_int32)]


class _Lights(C.Structure):
    _fields_ = [("lights", C.c_void_p), ("shadow", C.c_void_p), ("shadowRes", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_chain_texels.restype = C.c_size_t
        _lib.oracle_sample_cube.restype = C.c_float
        _lib.oracle_sample_cube.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        _lib.oracle_reconstruct_position.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib.oracle_reconstruct_positions.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_ibl_prefilter.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                              C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    """OpenMP threads of the restatement (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    lib().oracle_set_num_threads(int(n))


def rng(sx: int, sy: int, n: int):
    u = np.empty(n, np.uint32)
    f = np.empty(n, np.float32)
    lib().oracle_rng(C.c_uint32(sx), C.c_uint32(sy), n, _p(u), _p(f))
    return u, f


def half_roundtrip(x):
    x = _c(x, np.float32).ravel()
    h = np.empty(x.size, np.uint16)
    f = np.empty(x.size, np.float32)
    lib().oracle_half_roundtrip(_p(x), x.size, _p(h), _p(f))
    return h, f


def sample(data, w, h, mips, fmt, addr, uvl):
    uvl = _c(uvl, np.float32).reshape(-1, 3)
    out = np.empty((uvl.shape[0], 4), np.float32)
    lib().oracle_sample(_p(data), w, h, mips, fmt, addr, _p(uvl), uvl.shape[0], _p(out))
    return out


def reconstruct_positions(g: GlobalUniforms, W, H, depth, normal_u16):
    """(H, W, 4) float32 position image of mode D: reconstructPosition at every pixel centre, empty where normal.a == 0."""
    depth, normal_u16 = _c(depth, np.float32), _c(normal_u16, np.uint16)
    out = np.zeros((H, W, 4), np.float32)
    lib().oracle_reconstruct_positions(C.addressof(g), int(W), int(H), _p(depth), _p(normal_u16), _p(out))
    return out


def sample_cube(shadow, res, light, q):
    shadow = _c(shadow, np.float32)
    return float(lib().oracle_sample_cube(_p(shadow), res, light, float(q[0]), float(q[1]), float(q[2])))


def reconstruct_position(g: GlobalUniforms, u, v, d_raw):
    out = np.empty(3, np.float32)
    lib().oracle_reconstruct_position(C.addressof(g), float(u), float(v), float(d_raw), _p(out))
    return out


class Frame:
    """Host-side (numpy) bundle of one frame's inputs; keeps the arrays alive for the C calls."""

    def __init__(self, uniforms: GlobalUniforms, W, H, position, depth, normal, albedo, mro, env, prefiltered,
                 pre_size, pre_mips, irradiance, lut, lights=None, shadow=None, shadow_res=0):
        self.g = uniforms
        self.W, self.H = W, H
        self.position = _c(position, np.float32)
        self.depth = _c(depth, np.float32)
        self.normal = _c(normal, np.uint16)      # RGBA16F bit patterns
        self.albedo = _c(albedo, np.uint8)
        self.mro = _c(mro, np.uint8)
        self.env = _c(env, np.float32)
        self.prefiltered = _c(prefiltered, np.float32)  # tight mip chain, flat
        self.pre_size, self.pre_mips = pre_size, pre_mips
        self.irradiance = _c(irradiance, np.float32)
        self.lut = _c(lut, np.uint8)
        self.lights = _c(lights, np.float32)
        self.shadow = _c(shadow, np.float32)
        self.shadow_res = shadow_res if shadow is not None else 0

    def _structs(self):
        gb = _GBuffer(self.W, self.H, _p(self.position), _p(self.depth), _p(self.normal), _p(self.albedo), _p(self.mro))
        ibl = _IBL(_p(self.env), self.env.shape[1], self.env.shape[0], _p(self.prefiltered), self.pre_size[0],
                   self.pre_size[1], self.pre_mips, _p(self.irradiance), self.irradiance.shape[1],
                   self.irradiance.shape[0], _p(self.lut), self.lut.shape[1], self.lut.shape[0])
        li = _Lights(_p(self.lights), _p(self.shadow), self.shadow_res)
        return gb, ibl, li


def chain_texels(w, h, mips):
    return int(lib().oracle_chain_texels(w, h, mips))


def ssr_capture(fr: Frame):
    gb, ibl, li = fr._structs()
    refl = np.zeros((fr.H, fr.W, 4), np.uint16)
    hit = np.zeros((fr.H, fr.W), np.uint8)
    steps = np.zeros((fr.H, fr.W), np.uint8)
    lib().oracle_ssr_capture(C.byref(fr.g), C.byref(gb), C.byref(ibl), C.byref(li), _p(refl), _p(hit), _p(steps))
    return refl, hit, steps


def glossy_convolve(mip0_u16, mip_count=5):
    H, W = mip0_u16.shape[:2]
    chain = np.zeros(chain_texels(W, H, mip_count) * 4, np.uint16)
    chain[: W * H * 4] = np.ascontiguousarray(mip0_u16, np.uint16).ravel()
    lib().oracle_glossy_convolve(_p(chain), W, H, mip_count)
    return chain


def ssao(fr: Frame):
    gb, _, _ = fr._structs()
    out = np.zeros((fr.H, fr.W), np.uint8)
    lib().oracle_ssao(C.byref(fr.g), C.byref(gb), _p(out))
    return out


def deferred_shade(fr: Frame, refl_chain_u16, refl_mips=5, flags=SKIP_TONEMAP, ao_count=None):
    gb, ibl, li = fr._structs()
    refl_chain_u16 = _c(refl_chain_u16, np.uint16)
    ao_count = _c(ao_count, np.uint8)
    out = np.zeros((fr.H, fr.W, 4), np.float32)
    lib().oracle_deferred_shade(C.byref(fr.g), C.byref(gb), C.byref(ibl), C.byref(li), _p(refl_chain_u16), refl_mips,
                                C.c_uint32(flags), _p(ao_count), _p(out))
    return out


def mip_count(w, h):
    return int(lib().oracle_mip_count(w, h))


def env_mip_chain(env_rgba):
    env_rgba = _c(env_rgba, np.float32)
    H, W = env_rgba.shape[:2]
    mips = mip_count(W, H)
    chain = np.empty(chain_texels(W, H, mips) * 4, np.float32)
    lib().oracle_env_mip_chain(_p(env_rgba), W, H, mips, _p(chain))
    return chain, mips


def _texels(texels):
    t = np.ascontiguousarray(texels, np.int32).reshape(-1, 3)
    return t, t.shape[0]


def ibl_irradiance(chain, W, H, mips, out_w, out_h, texels, layout=LAYOUT_EQUIRECT, theta_samples=300):
    t, n = _texels(texels)
    out = np.empty((n, 4), np.float32)
    lib().oracle_ibl_irradiance(_p(chain), W, H, mips, layout, out_w, out_h, _p(t), n, theta_samples, _p(out))
    return out


def ibl_prefilter(chain, W, H, mips, out_w, out_h, roughness, texels, layout=LAYOUT_EQUIRECT, num_samples=10000,
                  sequence=SEQ_HASH):
    t, n = _texels(texels)
    out = np.empty((n, 4), np.float32)
    lib().oracle_ibl_prefilter(_p(chain), W, H, mips, layout, out_w, out_h, float(roughness), num_samples, sequence,
                               _p(t), n, _p(out))
    return out


def brdf_lut(size, samples=1024, k_mode=0):
    out = np.empty((size, size, 2), np.float32)
    lib().oracle_brdf_lut(size, samples, k_mode, _p(out))
    return out


# ---- rasterising producers (althea_oracle_raster.cpp) ---------------------------------------------------------------------
class _Tex(C.Structure):
    _fields_ = [("texels", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("mips", C.c_int32), ("sampler", C.c_uint32)]


class _Prim(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("idx", C.c_void_p), ("triCount", C.c_uint32), ("frontCW", C.c_uint32), ("model", C.c_float * 16),
                ("baseColorFactor", C.c_float * 4), ("baseUv", C.c_int32), ("mrUv", C.c_int32), ("normalScale", C.c_float),
                ("metallicFactor", C.c_float), ("roughnessFactor", C.c_float), ("alphaCutoff", C.c_float), ("base", _Tex), ("normal", _Tex),
                ("mr", _Tex)]


def _prim_array(prims):
    """prims: objects with .vertices (n, 26) f32, .indices u32, .model (row-major 4x4), .front_face_clockwise, .material (the
    fields of MaterialConstants; textures with .levels (list of (h, w, 4) u8) and .sampler). Returns (ctypes array, keep-alive)."""
    arr = (_Prim * max(1, len(prims)))()
    keep = []

    def tex(t):
        r = _Tex()
        if t is None:
            return r
        packed = np.concatenate([np.ascontiguousarray(l, np.uint8).reshape(-1) for l in t.levels])
        keep.append(packed)
        r.texels = packed.ctypes.data
        r.w, r.h, r.mips, r.sampler = t.levels[0].shape[1], t.levels[0].shape[0], len(t.levels), t.sampler
        return r

    for i, p in enumerate(prims):
        v = np.ascontiguousarray(p.vertices, np.float32)
        ix = np.ascontiguousarray(p.indices, np.uint32)
        keep += [v, ix]
        a = arr[i]
        a.verts, a.idx = v.ctypes.data, ix.ctypes.data
        a.triCount = len(ix) // 3
        a.frontCW = int(p.front_face_clockwise)
        cm = np.asarray(p.model, np.float32).T.reshape(-1)
        m = p.material
        for k in range(16):
            a.model[k] = float(cm[k])
        for k in range(4):
            a.baseColorFactor[k] = float(m.baseColorFactor[k])
        a.baseUv, a.mrUv = m.baseTextureCoordinateIndex, m.metallicRoughnessTextureCoordinateIndex
        a.normalScale, a.metallicFactor, a.roughnessFactor, a.alphaCutoff = m.normalScale, m.metallicFactor, m.roughnessFactor, m.alphaCutoff
        a.base, a.normal, a.mr = tex(m.baseTexture), tex(m.normalTexture), tex(m.metallicRoughnessTexture)
    return arr, keep


def draw_gbuffer(projection, view, prims, W: int, H: int) -> dict:
    """Gltf.vert/.frag through a LESS depth test. projection / view: 16 floats, column-major. Returns depth (H, W) f32,
    position / normal (H, W, 4) f32, albedo / mro (H, W, 4) u8, tri (H, W) u32 (draw-order triangle ordinal, 0xffffffff = none)."""
    arr, keep = _prim_array(prims)
    pj = np.ascontiguousarray(projection, np.float32).reshape(-1)
    vw = np.ascontiguousarray(view, np.float32).reshape(-1)
    out = {"depth": np.empty((H, W), np.float32), "position": np.empty((H, W, 4), np.float32), "normal": np.empty((H, W, 4), np.float32),
           "albedo": np.empty((H, W, 4), np.uint8), "mro": np.empty((H, W, 4), np.uint8), "tri": np.empty((H, W), np.uint32)}
    f = lib().oracle_draw_gbuffer
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
    f(pj.ctypes.data, vw.ctypes.data, C.addressof(arr), len(prims), W, H, out["depth"].ctypes.data, out["position"].ctypes.data,
      out["normal"].ctypes.data, out["albedo"].ctypes.data, out["mro"].ctypes.data, out["tri"].ctypes.data)
    return out


def draw_shadow_cubes(lights, projection, views, prims, res: int) -> np.ndarray:
    """ShadowMapBindless.vert/.frag. lights: (n, 8) f32 PointLight records; projection 16 floats; views (6, 16), column-major.
    Returns (n, 6, res, res) f32 holding min length(p - light) / 1000, 1.0 where nothing was drawn."""
    arr, keep = _prim_array(prims)
    lt = np.ascontiguousarray(lights, np.float32).reshape(-1, 8)
    pj = np.ascontiguousarray(projection, np.float32).reshape(-1)
    vw = np.ascontiguousarray(views, np.float32).reshape(-1)
    out = np.empty((lt.shape[0], 6, res, res), np.float32)
    f = lib().oracle_draw_shadow_cubes
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    f(lt.ctypes.data, lt.shape[0], pj.ctypes.data, vw.ctypes.data, C.addressof(arr), len(prims), res, out.ctypes.data)
    return out
