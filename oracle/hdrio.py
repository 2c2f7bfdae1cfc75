"""ORACLE-SIDE TEST INFRASTRUCTURE: Radiance RGBE (.hdr) codec in numpy.

Follows what the reference's IO does (Src/Utilities.cpp:216-255 -> stb_image.h `stbi_loadf` /
stb_image_write.h `stbi_write_hdr`): decode is mantissa * 2^(e-136) (no +0.5), encode is
frexp-normalised truncation. Used by the tests to read the reference's shipped fixtures and to
write/read the small committed fixtures under tests/golden/.
"""
from __future__ import annotations

import numpy as np


def _read_header(buf: bytes):
    pos = 0
    first = True
    while True:
        end = buf.index(b"\n", pos)
        line = buf[pos:end]
        pos = end + 1
        if first:
            if not (line.startswith(b"#?RADIANCE") or line.startswith(b"#?RGBE")):
                raise ValueError("not a Radiance HDR file")
            first = False
        if line == b"":
            break
    end = buf.index(b"\n", pos)
    tok = buf[pos:end].split()
    pos = end + 1
    if len(tok) != 4 or tok[0] != b"-Y" or tok[2] != b"+X":
        raise ValueError("unsupported HDR orientation: %r" % (tok,))
    return int(tok[3]), int(tok[1]), pos


def read_hdr_rgbe(path: str) -> np.ndarray:
    """Returns the raw RGBE bytes, shape (H, W, 4) uint8."""
    buf = open(path, "rb").read()
    w, h, pos = _read_header(buf)
    data = np.frombuffer(buf, dtype=np.uint8)
    out = np.empty((h, w, 4), dtype=np.uint8)
    if w < 8 or w >= 32768 or not (data[pos] == 2 and data[pos + 1] == 2 and not (data[pos + 2] & 0x80)):
        out[:] = data[pos:pos + h * w * 4].reshape(h, w, 4)
        return out
    for y in range(h):
        if data[pos] != 2 or data[pos + 1] != 2 or ((int(data[pos + 2]) << 8) | int(data[pos + 3])) != w:
            raise ValueError("corrupt RLE scanline %d" % y)
        pos += 4
        for c in range(4):
            x = 0
            row = out[y, :, c]
            while x < w:
                count = int(data[pos])
                pos += 1
                if count > 128:
                    count -= 128
                    row[x:x + count] = data[pos]
                    pos += 1
                else:
                    row[x:x + count] = data[pos:pos + count]
                    pos += count
                x += count
    return out


def rgbe_to_float(rgbe: np.ndarray) -> np.ndarray:
    """stbi__hdr_convert: value = mantissa * 2^(e - 136); e == 0 -> 0. Returns (H, W, 3) float32."""
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e != 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0.0)).astype(np.float32)
    return (rgbe[..., :3].astype(np.float32) * scale[..., None]).astype(np.float32)


def read_hdr(path: str) -> np.ndarray:
    return rgbe_to_float(read_hdr_rgbe(path))


def float_to_rgbe(rgb: np.ndarray) -> np.ndarray:
    """stbiw__linear_to_rgbe: truncating conversion with the shared exponent of the largest channel."""
    rgb = np.asarray(rgb, dtype=np.float32)
    maxc = rgb.max(axis=-1)
    out = np.zeros(rgb.shape[:-1] + (4,), dtype=np.uint8)
    ok = maxc >= 1e-32
    mant, expo = np.frexp(np.where(ok, maxc, 1.0).astype(np.float32))
    norm = (mant * np.float32(256.0) / np.where(ok, maxc, 1.0)).astype(np.float32)
    q = (rgb * norm[..., None]).astype(np.int64)
    out[..., :3] = np.where(ok[..., None], np.clip(q, 0, 255), 0).astype(np.uint8)
    out[..., 3] = np.where(ok, expo + 128, 0).astype(np.uint8)
    return out


def write_hdr(path: str, rgb: np.ndarray) -> None:
    """Writes an RLE-framed (all-literal runs) RGBE file that stb_image and read_hdr both accept."""
    rgbe = float_to_rgbe(rgb)
    h, w = rgbe.shape[:2]
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\n# Written by althea_b200 test tooling\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n")
        f.write(b"-Y %d +X %d\n" % (h, w))
        if w < 8 or w >= 32768:
            f.write(rgbe.tobytes())
            return
        for y in range(h):
            f.write(bytes([2, 2, (w >> 8) & 0xFF, w & 0xFF]))
            for c in range(4):
                row = rgbe[y, :, c].tobytes()
                for x in range(0, w, 128):
                    chunk = row[x:x + 128]
                    f.write(bytes([len(chunk)]))
                    f.write(chunk)
