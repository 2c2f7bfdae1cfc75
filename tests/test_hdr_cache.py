"""The on-disk IBL cache in the reference's format: RGBE codec properties on CPU (against the
oracle-side codec and, when mounted, the reference's shipped files), and the GPU round trip precompute -> .hdr -> reload."""
import os

import numpy as np
import pytest

from althea_b200 import hdr_cache
from helpers import GOLDEN, REFERENCE, golden_env
from oracle import hdrio


def test_rgbe_codec_matches_stb_semantics(tmp_path):
    rs = np.random.default_rng(3)
    img = np.exp(rs.uniform(-12, 9, (37, 53, 3))).astype(np.float32)
    img[0, 0] = 0.0
    img[1, 1] = (1e-35, 0.0, 0.0)      # below stb's 1e-32 cut-off -> (0,0,0,0)
    img[2, 2] = (1000.0, 1e-3, 1.0)    # shared exponent: small channels truncate to 0
    p = str(tmp_path / "a" / "t.hdr")
    hdr_cache.write_hdr(p, img)
    back = hdr_cache.read_hdr(p)
    assert back.shape == img.shape
    assert np.array_equal(back, hdrio.read_hdr(p))            # same decode as the oracle-side reader
    assert (back <= img + 1e-30).all()                         # truncation never rounds up
    step = img.max(axis=-1, keepdims=True) / 128.0             # one 8-bit mantissa step of the largest channel
    assert (img - back <= step * 1.0001 + 1e-30).all()
    assert not back[0, 0].any() and not back[1, 1].any()
    # idempotent: re-encoding decoded values reproduces the same bytes
    assert np.array_equal(hdr_cache.float_to_rgbe(back), hdr_cache.float_to_rgbe(hdr_cache.rgbe_to_float(hdr_cache.float_to_rgbe(img))))
    # narrow images are written flat, as stb does
    hdr_cache.write_hdr(str(tmp_path / "n.hdr"), img[:, :5])
    assert np.array_equal(hdr_cache.read_hdr(str(tmp_path / "n.hdr")), hdr_cache.rgbe_to_float(hdr_cache.float_to_rgbe(img[:, :5])))


def test_reads_the_committed_fixture_like_the_oracle_reader():
    p = os.path.join(GOLDEN, "env_512x256.hdr")
    assert np.array_equal(hdr_cache.read_hdr(p), hdrio.read_hdr(p))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference not mounted (GPU box)")
def test_reads_the_reference_shipped_cache_and_rewrites_it_losslessly(tmp_path):
    src = os.path.join(REFERENCE, "Content/PrecomputedMaps/LuxuryRoom/Prefiltered4.hdr")
    a = hdr_cache.read_hdr(src)
    assert a.shape == (128, 256, 3)
    assert np.array_equal(a, hdrio.read_hdr(src))
    out = str(tmp_path / "again.hdr")
    hdr_cache.write_hdr(out, a)
    assert np.array_equal(hdr_cache.read_hdr(out), a)          # decoded RGBE values survive a rewrite exactly
    assert hdr_cache.cache_complete(os.path.join(REFERENCE, "Content"), "LuxuryRoom")
    assert not hdr_cache.cache_complete(os.path.join(REFERENCE, "Content"), "NoSuchEnv")


@pytest.mark.gpu
def test_cache_round_trip_on_gpu(tmp_path, ctx_fast):
    from althea_b200 import _capi, engine
    env = golden_env()[::2, ::2].copy()  # 256 x 128
    res = engine.ImageBasedLighting.createResources(ctx_fast, env, lut_size=32)
    content = str(tmp_path / "Content")
    assert not hdr_cache.cache_complete(content, "TestEnv")
    hdr_cache.save_precomputed_maps(content, "TestEnv", res.irradianceMap, res.prefilteredMap)
    assert hdr_cache.cache_complete(content, "TestEnv")
    irr, pre = hdr_cache.load_precomputed_maps(content, "TestEnv")
    gi = res.irradianceMap.level_numpy(0).view(np.float32).reshape(128, 256, 4)
    assert irr.shape == gi.shape and (irr[..., 3] == 1).all()
    assert (np.abs(irr[..., :3] - gi[..., :3]) <= gi[..., :3].max(axis=-1, keepdims=True) / 128.0 + 1e-12).all()
    assert [p.shape[:2] for p in pre] == [(64, 128), (32, 64), (16, 32), (8, 16), (4, 8)]
    for k, p in enumerate(pre):
        g = res.prefilteredMap.level_numpy(k).view(np.float32).reshape(p.shape)
        assert (np.abs(p[..., :3] - g[..., :3]) <= g[..., :3].max(axis=-1, keepdims=True) / 128.0 + 1e-12).all()
    # what the engine would re-load goes back to the device as the run-time IBL set
    F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
    flat = np.concatenate([p.reshape(-1) for p in pre])
    reloaded = ctx_fast.image_from_numpy(flat, F32, 128, 64, 5)
    assert reloaded.mips == 5


# ---- pinned by the reference's own IO library ----------------------------------------------------------------------------
def _payload(file_bytes: bytes) -> bytes:
    """Everything after the free-text comment line of the header (ours names this repo, stb's names stb)."""
    return file_bytes[file_bytes.index(b"FORMAT="):]


def test_writer_and_reader_equal_the_references_stb_on_the_committed_vector(tmp_path):
    """tests/golden/stb_written.hdr was written by the reference's stbi_write_hdr and decoded by its stbi_loadf
    (tests/golden/make_hdr_golden.py): our writer must produce the same bytes, our reader the same floats."""
    g = np.load(os.path.join(GOLDEN, "stb_written.npz"))
    theirs = open(os.path.join(GOLDEN, "stb_written.hdr"), "rb").read()
    out = str(tmp_path / "ours.hdr")
    hdr_cache.write_hdr(out, g["image"])
    ours = open(out, "rb").read()
    assert ours.startswith(b"#?RADIANCE\n") and _payload(ours) == _payload(theirs)
    assert np.array_equal(hdr_cache.read_hdr(os.path.join(GOLDEN, "stb_written.hdr")), g["decoded"][..., :3])
    assert (g["decoded"][..., 3] == 1).all()
    # the numpy twins agree with the C++ codec
    assert np.array_equal(hdr_cache.rgbe_to_float(hdr_cache.float_to_rgbe(g["image"])), g["decoded"][..., :3])
    # runs longer than 127 and literal stretches longer than 128 are both in the vector
    assert len(theirs) < g["image"].shape[0] * g["image"].shape[1] * 4


def _stb():
    import ctypes as C
    import subprocess

    from helpers import ROOT
    if os.path.isdir(os.path.join(REFERENCE, "Extern", "stb")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    path = os.path.join(ROOT, "oracle", "_ref", "libstb_ref.so")
    if not os.path.exists(path):
        pytest.skip("reference sources not mounted and oracle/_ref not built")
    lib = C.CDLL(path)
    lib.ref_stbi_loadf_from_memory.restype = C.POINTER(C.c_float)
    return lib


def _stb_loadf(lib, path):
    import ctypes as C
    buf = open(path, "rb").read()
    w, h = C.c_int(), C.c_int()
    p = lib.ref_stbi_loadf_from_memory(buf, len(buf), C.byref(w), C.byref(h))
    if not p:
        return None
    a = np.ctypeslib.as_array(p, (h.value, w.value, 4)).copy()
    lib.ref_stbi_free(p)
    return a


def test_live_against_the_references_stb(tmp_path):
    """Random images of awkward sizes: identical payload bytes, identical decode in both directions; malformed files are
    rejected by both."""
    import ctypes as C
    lib = _stb()
    rs = np.random.default_rng(0)
    for k, (w, h) in enumerate([(53, 37), (5, 9), (8, 3), (300, 20), (1024, 8), (129, 4), (7, 7), (32767, 1), (128, 2), (131, 3)]):
        img = np.exp(rs.uniform(-14, 10, (h, w, 4))).astype(np.float32)
        if k % 2 == 0:
            img[:, : w // 2] = img[:1, :1]
            img[1::2] = 0
        if k % 3 == 0:
            img = np.round(img * 4) / 4          # coarse values: many short runs of two and three
        img = np.ascontiguousarray(img, np.float32)
        a, b = str(tmp_path / ("a%d.hdr" % k)), str(tmp_path / ("b%d.hdr" % k))
        hdr_cache.write_hdr(a, img)
        assert lib.ref_stbi_write_hdr(b.encode(), w, h, img.ctypes.data_as(C.c_void_p)) == 1
        assert _payload(open(a, "rb").read()) == _payload(open(b, "rb").read()), (w, h)
        theirs = _stb_loadf(lib, a)
        assert np.array_equal(hdr_cache.read_hdr(b), theirs[..., :3]) and np.array_equal(hdr_cache.read_hdr(a), theirs[..., :3])
    good = open(str(tmp_path / "a0.hdr"), "rb").read()
    for name, bad in (("no_format", good.replace(b"FORMAT=32-bit_rle_rgbe", b"FORMAT=32-bit_rle_xyze")),
                      ("not_hdr", b"#?RADIANT\nnothing of the kind\n"), ("bad_layout", good.replace(b"-Y ", b"+Y "))):
        p = str(tmp_path / (name + ".hdr"))
        open(p, "wb").write(bad)
        assert _stb_loadf(lib, p) is None, name
        with pytest.raises(ValueError):
            hdr_cache.read_hdr(p)
    with pytest.raises(FileNotFoundError):
        hdr_cache.read_hdr(str(tmp_path / "missing.hdr"))
    # stricter than stb on purpose: a file cut short is an error here (stb pads the missing bytes with zeros and carries on)
    open(str(tmp_path / "cut.hdr"), "wb").write(good[: len(good) // 2])
    with pytest.raises(ValueError):
        hdr_cache.read_hdr(str(tmp_path / "cut.hdr"))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference not mounted (GPU box)")
def test_the_references_shipped_cache_files_survive_a_rewrite_byte_for_byte(tmp_path):
    """Content/PrecomputedMaps/*/Prefiltered5.hdr were written by the engine itself (saveHdri). Decoding them and writing the
    floats back with our writer reproduces the engine's files byte for byte (after the comment line), and both decoders agree
    on every file of the shipped cache."""
    lib = _stb()
    for env in ("LuxuryRoom", "NeoclassicalInterior", "ThatchChapel"):
        for name in ("Prefiltered5.hdr", "Prefiltered3.hdr"):
            src = os.path.join(REFERENCE, "Content/PrecomputedMaps", env, name)
            a = hdr_cache.read_hdr(src)
            assert np.array_equal(a, _stb_loadf(lib, src)[..., :3])
            out = str(tmp_path / (env + name))
            hdr_cache.write_hdr(out, a)
            assert _payload(open(out, "rb").read()) == _payload(open(src, "rb").read()), (env, name)


def test_cpp_mirror_utilities_load_and_save(tmp_path):
    """Utilities::loadHdri / saveHdri with the reference's names and result struct (Include/Althea/Utilities.h:29-53), driven from
    C++: load the committed stb-written file, save it again, same payload; a missing file throws as the reference's does."""
    import subprocess

    from helpers import ROOT
    src = tmp_path / "hdr_main.cpp"
    src.write_text(r'''
#include "Althea/Utilities.h"
#include <cstdio>
using namespace AltheaEngine;
int main(int argc, char** argv) {
  if (argc != 3) return 1;
  Utilities::ImageFile img;
  Utilities::loadHdri(argv[1], img);
  if (img.channels != 4 || img.bytesPerChannel != 4 || img.data.size() != (size_t)img.width * img.height * 16) return 2;
  Utilities::saveHdri(argv[2], img.width, img.height, img.data.data(), img.data.size());
  try { Utilities::loadHdri("/nonexistent/none.hdr", img); return 3; } catch (const std::runtime_error&) {}
  std::printf("%d %d\n", img.width, img.height);
  return 0;
}
''')
    exe = tmp_path / "hdr_main"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "althea_b200", "host"),
                    str(src), "-o", str(exe)], check=True)
    golden = os.path.join(GOLDEN, "stb_written.hdr")
    out = str(tmp_path / "again.hdr")
    r = subprocess.run([str(exe), golden, out], check=True, capture_output=True, text=True)
    assert r.stdout.split() == ["300", "24"]
    assert _payload(open(out, "rb").read()) == _payload(open(golden, "rb").read())


# ---- createResources(envMapName): the reference's cache-or-compute flow, host logic on CPU --------------------------------
class _FakeImage:
    def __init__(self, arr, fmt, w, h, mips):
        self.arr, self.format, self.w, self.h, self.mips = np.asarray(arr), fmt, w, h, mips

    def level_numpy(self, k):
        off = sum(max(1, self.w >> i) * max(1, self.h >> i) * 4 for i in range(k))
        n = max(1, self.w >> k) * max(1, self.h >> k) * 4
        return self.arr.reshape(-1)[off:off + n].copy()


class _FakeCtx:
    """Stands in for engine.Context where only host-side bookkeeping is exercised (no CUDA on this machine)."""

    def __init__(self):
        self.images, self.syncs = [], 0

    def image_from_numpy(self, arr, fmt, w, h, mips=1, layers=1):
        self.images.append(_FakeImage(arr, fmt, w, h, mips))
        return self.images[-1]

    def synchronize(self, stream=0):
        self.syncs += 1


def test_create_resources_from_content_follows_the_reference_flow(tmp_path, monkeypatch):
    """ImageBasedLighting.cpp:415-446: missing env map -> the reference's error; incomplete cache -> compute, write the six files,
    then load everything back from disk; complete cache -> no compute at all. The GPU work is stubbed out here (it is covered by
    test_cache_round_trip_on_gpu and test_ibl_parity); what is checked is which files are read and written and what reaches the
    device."""
    from PIL import Image as PILImage

    from althea_b200 import _capi, engine
    content = str(tmp_path / "Content")
    ctx = _FakeCtx()
    with pytest.raises(RuntimeError, match="Specified environment map does not exist"):
        engine.ImageBasedLighting.createResourcesFromContent(ctx, content, "Studio")
    rs = np.random.default_rng(2)
    env = np.exp(rs.uniform(-3, 3, (32, 64, 3))).astype(np.float32)
    hdr_cache.write_hdr(os.path.join(content, "HDRI_Skybox", "Studio.hdr"), env)
    lut = rs.integers(0, 256, (8, 8, 4), dtype=np.uint8)
    lut[..., 3] = 255
    os.makedirs(os.path.join(content, "PrecomputedMaps"), exist_ok=True)
    PILImage.fromarray(lut, "RGBA").save(os.path.join(content, "PrecomputedMaps", "brdf_lut.png"))
    computed = []

    def fake_create(c, env_rgba, brdf_lut_rgba8=None, lut_size=512, stream=0):
        computed.append(env_rgba.shape)
        H, W = env_rgba.shape[:2]
        irr = _FakeImage(np.full((H, W, 4), 0.25, np.float32), 0, W, H, 1)
        pre = _FakeImage(np.concatenate([np.full((H >> (k + 1)) * (W >> (k + 1)) * 4, 0.5 + k, np.float32) for k in range(5)]), 0, W >> 1, H >> 1, 5)
        return engine.IBLResources(None, pre, irr, None)

    monkeypatch.setattr(engine.ImageBasedLighting, "createResources", staticmethod(fake_create))
    res = engine.ImageBasedLighting.createResourcesFromContent(ctx, content, "Studio")
    assert computed == [(32, 64, 4)] and ctx.syncs == 1 and hdr_cache.cache_complete(content, "Studio")
    env_img, irr_img, pre_img, lut_img = ctx.images
    assert (env_img.w, env_img.h, env_img.mips) == (64, 32, 1) and (irr_img.w, irr_img.h) == (64, 32)
    assert (pre_img.w, pre_img.h, pre_img.mips) == (32, 16, 5) and pre_img.arr.size == sum((16 >> k) * (32 >> k) * 4 for k in range(5))
    assert np.array_equal(env_img.arr[..., :3], hdr_cache.read_hdr(os.path.join(content, "HDRI_Skybox", "Studio.hdr")))
    assert np.allclose(irr_img.arr[..., :3], 0.25) and (irr_img.arr[..., 3] == 1).all()          # 0.25 is exact in RGBE
    assert pre_img.arr[0] == 0.5 and pre_img.arr[-4] == 4.5
    assert lut_img.format == _capi.FORMAT_R8G8B8A8_UNORM and np.array_equal(lut_img.arr, lut)
    assert res.irradianceMap is irr_img and res.prefilteredMap is pre_img
    # second call: the cache is complete, nothing is computed
    ctx2 = _FakeCtx()
    engine.ImageBasedLighting.createResourcesFromContent(ctx2, content, "Studio")
    assert computed == [(32, 64, 4)] and ctx2.syncs == 0 and len(ctx2.images) == 4


def test_cpp_mirror_create_resources_by_name_compiles(tmp_path):
    """The C++ twin of createResourcesFromContent (host/Althea/ImageBasedLighting.h): reference signature shape
    createResources(app, ..., envMapName), built on Utilities::loadHdri / saveHdri. Compile check (running it needs a GPU; the
    pieces it calls are covered by demo_frame's GPU test and by the Utilities test above)."""
    import subprocess

    from helpers import ROOT
    src = tmp_path / "ibl_probe.cpp"
    src.write_text('''
#include "Althea/ImageBasedLighting.h"
using namespace AltheaEngine;
uint64_t probe(const CudaApplication& app) {
  IBLResources r = ImageBasedLighting::createResources(app, "/srv/Content", "LuxuryRoom");
  return r.getHandles().irradiance;
}
''')
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "althea_b200", "host"), str(src)], check=True)


# ---- Utilities::saveExr (Src/Utilities.cpp:258-271): the dump format of a golden frame ------------------------------------
def _exr_header(path):
    """(attributes {name: (type, bytes)}, header end offset) of a single-part OpenEXR file."""
    import struct
    b = open(path, "rb").read()
    assert struct.unpack_from("<I", b, 0)[0] == 20000630 and (struct.unpack_from("<I", b, 4)[0] & 0xFF) == 2
    o, attrs = 8, {}
    while b[o] != 0:
        e = b.index(b"\0", o); name = b[o:e].decode(); o = e + 1
        e = b.index(b"\0", o); typ = b[o:e].decode(); o = e + 1
        size = struct.unpack_from("<i", b, o)[0]; o += 4
        attrs[name] = (typ, b[o:o + size]); o += size
    return attrs, o + 1


def _tinyexr():
    import ctypes as C
    import subprocess

    from helpers import ROOT
    if os.path.isdir(os.path.join(REFERENCE, "Extern", "tinyexr")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    path = os.path.join(ROOT, "oracle", "_ref", "libtinyexr_ref.so")
    if not os.path.exists(path):
        pytest.skip("reference sources not mounted and oracle/_ref not built")
    lib = C.CDLL(path)
    lib.ref_load_exr.argtypes = [C.c_char_p, C.c_void_p, C.c_ulonglong, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_save_exr.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    return lib


def test_exr_writer_has_the_layout_the_reference_writes(tmp_path):
    """An uncompressed scan-line file with the channel list, windows and line order of tinyexr's SaveEXR(data, w, h, 4, 0, path);
    every line decodes back to the bits that went in (own reader here: the file is plain enough to parse in ten lines)."""
    import struct
    rs = np.random.default_rng(3)
    for (w, h) in ((1, 1), (5, 3), (64, 17), (130, 33)):
        img = rs.normal(0, 50, (h, w, 4)).astype(np.float32)
        img[0, 0] = [np.inf, -0.0, 1e-42, 3.4e38]  # specials survive: the payload is raw fp32
        path = str(tmp_path / ("o_%dx%d.exr" % (w, h)))
        hdr_cache.save_exr(path, img)
        attrs, end = _exr_header(path)
        assert attrs["channels"][0] == "chlist"
        names, o, c = [], 0, attrs["channels"][1]
        while c[o] != 0:
            e = c.index(b"\0", o); names.append(c[o:e].decode())
            ptype, plinear, xs, ys = struct.unpack_from("<iB3xii", c, e + 1)
            assert (ptype, xs, ys) == (2, 1, 1)  # FLOAT, no subsampling
            o = e + 1 + 16
        assert names == ["A", "B", "G", "R"]
        assert attrs["compression"][1] == b"\0" and attrs["lineOrder"][1] == b"\0"
        assert struct.unpack("<4i", attrs["dataWindow"][1]) == (0, 0, w - 1, h - 1) == struct.unpack("<4i", attrs["displayWindow"][1])
        b = open(path, "rb").read()
        offs = struct.unpack_from("<%dQ" % h, b, end)
        got = np.zeros_like(img)
        for y, o in enumerate(offs):
            yy, size = struct.unpack_from("<ii", b, o)
            assert (yy, size) == (y, w * 16)
            planes = np.frombuffer(b, np.float32, 4 * w, o + 8).reshape(4, w)  # A, B, G, R
            got[y] = planes[::-1].T
        assert np.array_equal(got.view(np.uint32), img.view(np.uint32))
    with pytest.raises(ValueError):
        hdr_cache.save_exr(str(tmp_path / "bad.exr"), np.zeros((4, 4, 3), np.float32))
    with pytest.raises(OSError):
        hdr_cache.save_exr(str(tmp_path / "no" / "such" / "dir.exr"), np.zeros((4, 4, 4), np.float32))


def test_exr_against_the_references_tinyexr(tmp_path):
    """The reference's own library (Extern/tinyexr, compiled into oracle/_ref) reads the product's file back to the same bits,
    and the file the reference writes for the same image has the same channel list, windows and line order (it deflates its
    blocks, the product does not: the one attribute allowed to differ)."""
    import ctypes as C
    lib = _tinyexr()
    rs = np.random.default_rng(4)
    for (w, h) in ((7, 5), (64, 40), (257, 19)):
        img = np.exp(rs.uniform(-12, 9, (h, w, 4))).astype(np.float32)
        ours, theirs = str(tmp_path / "ours.exr"), str(tmp_path / "theirs.exr")
        hdr_cache.save_exr(ours, img)
        out = np.zeros_like(img)
        ww, hh = C.c_int(), C.c_int()
        assert lib.ref_load_exr(os.fsencode(ours), out.ctypes.data, out.size, C.byref(ww), C.byref(hh)) == 0
        assert (ww.value, hh.value) == (w, h) and np.array_equal(out.view(np.uint32), img.view(np.uint32))
        assert lib.ref_save_exr(os.fsencode(theirs), w, h, img.ctypes.data) == 0
        a, _ = _exr_header(ours)
        b, _ = _exr_header(theirs)
        for key in ("channels", "dataWindow", "displayWindow", "lineOrder", "pixelAspectRatio", "screenWindowCenter", "screenWindowWidth"):
            assert a[key] == b[key], key
        assert set(a) == set(b)
