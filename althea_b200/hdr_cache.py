"""On-disk IBL cache in the reference's own format.

`ImageBasedLighting::createResources` (Src/ImageBasedLighting.cpp:415-446) skips the precompute when
Content/PrecomputedMaps/<env>/{IrradianceMap,Prefiltered1..5}.hdr all exist, and otherwise writes them with
`Utilities::saveHdri` (Src/Utilities.cpp:244-255 -> stb_image_write.h stbi_write_hdr: Radiance RGBE, RLE scanlines,
frexp-normalised TRUNCATED mantissas) and re-loads them with `Utilities::loadHdri` (stbi_loadf: mantissa * 2^(e-136)).
This module writes and reads the same files from the CUDA-generated maps, so the untouched engine loader picks them up and
the run-time data carries the same RGBE quantisation as the reference's.

Host-side IO only (numpy); the maps themselves come from althea_cuda_ibl_precompute.
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


def write_hdr(path: str, rgba: np.ndarray) -> None:
    """stbi_write_hdr's container: '#?RADIANCE' header, -Y H +X W, one RLE-framed scanline per row (channel-planar)."""
    rgbe = float_to_rgbe(rgba)
    h, w = rgbe.shape[:2]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\n# Written by althea_b200 (stb_image_write.h layout)\nFORMAT=32-bit_rle_rgbe\n")
        f.write(("EXPOSURE=          1.0000000000000\n\n-Y %d +X %d\n" % (h, w)).encode())
        if w < 8 or w >= 32768:  # stb writes such images flat
            f.write(rgbe.tobytes())
            return
        head = bytes([2, 2, (w >> 8) & 0xFF, w & 0xFF])
        for y in range(h):
            f.write(head)
            for c in range(4):
                row = rgbe[y, :, c]
                for x0 in range(0, w, 128):  # literal packets only (count <= 128): valid RLE framing, no run detection needed
                    chunk = row[x0:x0 + 128]
                    f.write(bytes([len(chunk)]))
                    f.write(chunk.tobytes())


def read_hdr(path: str) -> np.ndarray:
    """Radiance RGBE -> (H, W, 3) float32 as stbi_loadf decodes it (flat or RLE scanlines)."""
    buf = open(path, "rb").read()
    pos = 0
    if not (buf.startswith(b"#?RADIANCE") or buf.startswith(b"#?RGBE")):
        raise ValueError("%s is not a Radiance HDR file" % path)
    while True:
        end = buf.index(b"\n", pos)
        line = buf[pos:end]
        pos = end + 1
        if line == b"":
            break
    end = buf.index(b"\n", pos)
    tok = buf[pos:end].split()
    pos = end + 1
    if len(tok) != 4 or tok[0] != b"-Y" or tok[2] != b"+X":
        raise ValueError("unsupported HDR orientation in %s" % path)
    h, w = int(tok[1]), int(tok[3])
    data = np.frombuffer(buf, np.uint8)
    out = np.empty((h, w, 4), np.uint8)
    if w < 8 or w >= 32768 or not (data[pos] == 2 and data[pos + 1] == 2 and not (data[pos + 2] & 0x80)):
        out[:] = data[pos:pos + h * w * 4].reshape(h, w, 4)
        return rgbe_to_float(out)
    for y in range(h):
        if data[pos] != 2 or data[pos + 1] != 2 or ((int(data[pos + 2]) << 8) | int(data[pos + 3])) != w:
            raise ValueError("corrupt RLE scanline %d in %s" % (y, path))
        pos += 4
        for c in range(4):
            x = 0
            while x < w:
                count = int(data[pos])
                pos += 1
                if count > 128:
                    count -= 128
                    out[y, x:x + count, c] = data[pos]
                    pos += 1
                else:
                    out[y, x:x + count, c] = data[pos:pos + count]
                    pos += count
                x += count
    return rgbe_to_float(out)


def cache_paths(content_dir: str, env_name: str) -> Tuple[str, list]:
    """(IrradianceMap.hdr, [Prefiltered1.hdr .. Prefiltered5.hdr]) under <content_dir>/PrecomputedMaps/<env_name>/."""
    d = os.path.join(content_dir, "PrecomputedMaps", env_name)
    return os.path.join(d, "IrradianceMap.hdr"), [os.path.join(d, "Prefiltered%d.hdr" % (i + 1)) for i in range(PREFILTERED_COUNT)]


def cache_complete(content_dir: str, env_name: str) -> bool:
    """The reference's needToPrecomputeIBL test, negated (ImageBasedLighting.cpp:427-442)."""
    irr, pre = cache_paths(content_dir, env_name)
    return os.path.exists(irr) and all(os.path.exists(p) for p in pre)


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
