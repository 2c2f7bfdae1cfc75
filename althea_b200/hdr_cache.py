"""On-disk IBL cache in the reference's own format.

`ImageBasedLighting::createResources` (Src/ImageBasedLighting.cpp:415-446) skips the precompute when
Content/PrecomputedMaps/<env>/{IrradianceMap,Prefiltered1..5}.hdr all exist, and otherwise writes them with
`Utilities::saveHdri` (Src/Utilities.cpp:244-255 -> stb_image_write.h stbi_write_hdr: Radiance RGBE, RLE scanlines,
frexp-normalised TRUNCATED mantissas) and re-loads them with `Utilities::loadHdri` (stbi_loadf: mantissa * 2^(e-136)).
This module writes and reads the same files from the CUDA-generated maps, so the untouched engine loader picks them up and
the run-time data carries the same RGBE quantisation as the reference's.

Host-side IO only: the codec is the C++ mirror of Utilities::saveHdri / loadHdri (host/Althea/Utilities.h behind
include/althea_host.h), held byte for byte to the reference's stb build (tests/test_hdr_cache.py); float_to_rgbe /
rgbe_to_float below are its numpy twins for array-level work. The maps themselves come from althea_cuda_ibl_precompute.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np

PREFILTERED_COUNT = 5  # ImageBasedLighting.cpp:432-437


def float_to_rgbe(rgb: np.ndarray) -> np.ndarray:
    """(H, W, >=3) float32 -> (H, W, 4) uint8, stbiw__linear_to_rgbe: e = exponent + 128, mantissa = (uchar)(c * 256 / 2^exponent)."""
    rgb = np.asarray(rgb, np.float32)[..., :3]
    maxc = rgb.max(axis=-1)
    out = np.zeros(rgb.shape[:2] + (4,), np.uint8)
    ok = maxc >= 1e-32
    mant, expo = np.frexp(maxc.astype(np.float32))
    scale = np.where(ok, mant * 256.0 / np.where(ok, maxc, 1.0), 0.0).astype(np.float32)
    out[..., :3] = np.clip((rgb * scale[..., None]), 0, 255).astype(np.uint8)  # C cast: truncation
    out[..., 3] = np.where(ok, expo + 128, 0).astype(np.uint8)
    out[~ok] = 0
    return out


def rgbe_to_float(rgbe: np.ndarray) -> np.ndarray:
    """stbi__hdr_convert: mantissa * 2^(e - 136); e == 0 -> 0. Returns (H, W, 3) float32."""
    e = rgbe[..., 3].astype(np.int32)
    f = np.ldexp(np.float32(1.0), e - 136).astype(np.float32)
    out = rgbe[..., :3].astype(np.float32) * f[..., None]
    out[e == 0] = 0
    return out


def _host():
    from . import _hostapi
    return _hostapi.load()


def write_hdr(path: str, rgba: np.ndarray) -> None:
    """Utilities::saveHdri through libalthea_host.so (host/Althea/Utilities.h): the file stbi_write_hdr would write, byte for
    byte after the header's comment line. rgba: (H, W, 3 or 4) float; alpha is not stored."""
    a = np.asarray(rgba, np.float32)
    if a.ndim != 3 or a.shape[2] not in (3, 4):
        raise ValueError("write_hdr takes an (H, W, 3|4) image")
    if a.shape[2] == 3:
        a = np.concatenate([a, np.ones(a.shape[:2] + (1,), np.float32)], -1)
    a = np.ascontiguousarray(a)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    rc = _host().althea_host_save_hdri(os.fsencode(path), a.shape[1], a.shape[0], a.ctypes.data)
    if rc != 0:
        raise OSError("althea_host_save_hdri(%s) failed: %d" % (path, rc))


def read_hdr(path: str) -> np.ndarray:
    """Utilities::loadHdri through libalthea_host.so: Radiance RGBE -> (H, W, 3) float32 exactly as stbi_loadf decodes it."""
    import ctypes as C
    w, h = C.c_int32(), C.c_int32()
    rc = _host().althea_host_load_hdri_info(os.fsencode(path), C.byref(w), C.byref(h))
    if rc == -2:
        raise FileNotFoundError(path)
    if rc != 0:
        raise ValueError("%s is not a Radiance HDR file (%d)" % (path, rc))
    out = np.empty((h.value, w.value, 4), np.float32)
    rc = _host().althea_host_load_hdri(os.fsencode(path), out.ctypes.data, out.size)
    if rc != 0:
        raise ValueError("%s: decode failed (%d)" % (path, rc))
    return np.ascontiguousarray(out[..., :3])


def cache_paths(content_dir: str, env_name: str) -> Tuple[str, list]:
    """(IrradianceMap.hdr, [Prefiltered1.hdr .. Prefiltered5.hdr]) under <content_dir>/PrecomputedMaps/<env_name>/."""
    d = os.path.join(content_dir, "PrecomputedMaps", env_name)
    return os.path.join(d, "IrradianceMap.hdr"), [os.path.join(d, "Prefiltered%d.hdr" % (i + 1)) for i in range(PREFILTERED_COUNT)]


def cache_complete(content_dir: str, env_name: str) -> bool:
    """The reference's needToPrecomputeIBL test, negated (ImageBasedLighting.cpp:427-442)."""
    irr, pre = cache_paths(content_dir, env_name)
    return os.path.exists(irr) and all(os.path.exists(p) for p in pre)


def environment_map_path(content_dir: str, env_name: str) -> str:
    """<content_dir>/HDRI_Skybox/<env_name>.hdr; missing => the reference's error (ImageBasedLighting.cpp:421-425)."""
    p = os.path.join(content_dir, "HDRI_Skybox", env_name + ".hdr")
    if not os.path.exists(p):
        raise RuntimeError("Specified environment map does not exist!")
    return p


def save_precomputed_maps(content_dir: str, env_name: str, irradiance, prefiltered) -> None:
    """Downloads the CUDA-generated maps (engine.Image objects, RGBA32F; prefiltered = 5 mips) and writes the six files."""
    irr_path, pre_paths = cache_paths(content_dir, env_name)
    write_hdr(irr_path, irradiance.level_numpy(0).view(np.float32).reshape(irradiance.h, irradiance.w, 4))
    if prefiltered.mips != PREFILTERED_COUNT:
        raise ValueError("the reference's cache holds exactly %d prefiltered levels" % PREFILTERED_COUNT)
    for k, p in enumerate(pre_paths):
        lw, lh = max(1, prefiltered.w >> k), max(1, prefiltered.h >> k)
        write_hdr(p, prefiltered.level_numpy(k).view(np.float32).reshape(lh, lw, 4))


def load_precomputed_maps(content_dir: str, env_name: str):
    """(irradiance (H, W, 4) float32, [5 prefiltered levels (h, w, 4) float32]) with alpha = 1, as loadHdri returns them."""
    irr_path, pre_paths = cache_paths(content_dir, env_name)

    def rgba(a):
        return np.concatenate([a, np.ones(a.shape[:2] + (1,), np.float32)], -1)

    return rgba(read_hdr(irr_path)), [rgba(read_hdr(p)) for p in pre_paths]


def save_exr(path: str, rgba: np.ndarray):
    """Utilities::saveExr through libalthea_host.so: (H, W, 4) float32 -> an OpenEXR file in the layout the reference writes
    (Src/Utilities.cpp:258-271), the dump format of a golden frame."""
    a = np.ascontiguousarray(rgba, np.float32)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("save_exr needs an (H, W, 4) float32 image")
    rc = _host().althea_host_save_exr(os.fsencode(path), a.shape[1], a.shape[0], a.ctypes.data)
    if rc != 0:
        raise OSError("althea_host_save_exr(%s) failed: %d" % (path, rc))
