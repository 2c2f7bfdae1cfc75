"""Pins the oracle to the reference's OWN outputs: the shipped Content/PrecomputedMaps. The full-size
inputs cannot be committed (3 x 25 MB) and /root/reference does not exist on the GPU box, so:
  * test_pin_record_*  check the committed record tests/golden/ibl_pin.json (shipped value vs oracle value per probe texel,
    written by tests/golden/make_golden.py) against the RGBE-quantisation bound;
  * test_pin_live_*    recompute a subset from /root/reference when it is mounted (the build container), else skip.
The BRDF LUT (a10) golden IS committed (brdf_lut.npz = the reference's brdf_lut.png) and is checked directly."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, REFERENCE
from oracle import hdrio


def _rgbe_bound(shipped):
    # stb writes truncated 8-bit mantissas sharing the exponent of the largest channel: the stored value is at most one
    # mantissa step below the true one; allow 1.5 steps for the oracle's own fp32 summation-order noise.
    shipped = np.asarray(shipped, np.float64)
    step = np.max(shipped) / 128.0
    return 1.5 * step + 1e-6


def test_pin_record_irradiance_and_prefilter():
    """Bounds: the irradiance (a smooth integrand) must sit inside the RGBE truncation interval for every probe; the
    prefilter mostly does, with outliers of a few mantissa steps at high-contrast HDR texels, where the reference GPU's
    texture unit (8-bit filter weights, its own atan/log2) and FP32 differ most (ThatchChapel has the sun in frame)."""
    rec = json.load(open(os.path.join(GOLDEN, "ibl_pin.json")))
    assert set(rec["envs"]) == {"NeoclassicalInterior", "LuxuryRoom", "ThatchChapel"}
    for name, e in rec["envs"].items():
        for p in e["irradiance"]:
            s, o = np.array(p["shipped"]), np.array(p["oracle"])
            assert np.all(np.abs(o - s) <= _rgbe_bound(s)), (name, p)
        steps, rels = [], []
        for k in e["prefiltered"].values():
            assert len(k["probes"]) >= 24
            for p in k["probes"]:
                s, o = np.array(p["shipped"]), np.array(p["oracle"])
                steps.append(np.abs(o - s).max() / (s.max() / 128.0))
                rels.append(np.abs(o - s).max() / s.max())
        steps, rels = np.array(steps), np.array(rels)
        assert np.median(rels) < 0.007, name
        assert (steps <= 1.5).mean() >= 0.75, name
        assert np.percentile(steps, 95) < 3.0, name
        assert rels.max() < 0.08, name


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference not mounted (GPU box)")
def test_pin_live_prefilter_and_irradiance(oracle):
    name = "LuxuryRoom"
    env = hdrio.read_hdr(os.path.join(REFERENCE, "Content/HDRI_Skybox/%s.hdr" % name))
    H, W = env.shape[:2]
    chain, mips = oracle.env_mip_chain(np.concatenate([env, np.ones((H, W, 1), np.float32)], -1))
    assert mips == 13
    irr = hdrio.read_hdr(os.path.join(REFERENCE, "Content/PrecomputedMaps/%s/IrradianceMap.hdr" % name))
    pts = [(123, 456), (2500, 1500), (4095, 2047), (0, 0)]
    out = oracle.ibl_irradiance(chain, W, H, mips, W, H, [(x, y, 0) for x, y in pts])
    for (x, y), o in zip(pts, out):
        assert np.all(np.abs(o[:3] - irr[y, x]) <= _rgbe_bound(irr[y, x]))
    # Prefiltered1 = roughness 0 = corner-aligned bilinear decimation: check a whole row band cheaply (8 samples suffice,
    # every sample is identical at roughness 0)
    pf1 = hdrio.read_hdr(os.path.join(REFERENCE, "Content/PrecomputedMaps/%s/Prefiltered1.hdr" % name))
    tex = [(x, 300, 0) for x in range(0, 2048, 7)]
    out = oracle.ibl_prefilter(chain, W, H, mips, 2048, 1024, 0.0, tex, num_samples=8)
    ref = np.array([pf1[300, x] for x, _, _ in tex])
    rel = np.abs(out[:, :3] - ref) / np.maximum(ref, 1e-3)
    assert np.median(rel) < 0.006 and np.percentile(rel, 99) < 0.03
    for i, rough in ((3, 0.5), (5, 1.0)):
        pf = hdrio.read_hdr(os.path.join(REFERENCE, "Content/PrecomputedMaps/%s/Prefiltered%d.hdr" % (name, i)))
        h, w = pf.shape[:2]
        pts = [(w // 3, h // 2, 0), (w - 1, 0, 0), (5, h - 2, 0)]
        out = oracle.ibl_prefilter(chain, W, H, mips, w, h, rough, pts)
        for (x, y, _), o in zip(pts, out):
            assert np.all(np.abs(o[:3] - pf[y, x]) <= 2.0 * _rgbe_bound(pf[y, x]))


def test_brdf_lut_matches_reference_asset(oracle):
    # the reference's asset stores roughness increasing UPWARD and is sampled un-flipped: compare bottom-up
    ref = np.load(os.path.join(GOLDEN, "brdf_lut.npz"))["lut"][..., :2].astype(np.float32) / 255.0
    lut = oracle.brdf_lut(450, 256, 0)[::-1]
    d = np.abs(lut - ref)
    assert d.mean() < 0.01          # palette (256 colours) quantisation
    assert d[8:-8, 8:-8].max() < 0.15
    wrong = np.abs(oracle.brdf_lut(450, 64, 0) - ref).mean()
    assert wrong > 10 * d.mean()    # the un-flipped reading is clearly not what the asset holds


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference not mounted (GPU box)")
def test_committed_brdf_lut_is_what_the_engine_decodes():
    """tests/golden/brdf_lut.npz was expanded from the palette PNG with Pillow; the engine decodes it with stb
    (Src/Utilities.cpp:104-150 -> stbi_load_from_memory, 4 channels). Same texels from the reference's own stb build."""
    import ctypes as C
    import subprocess

    from helpers import ROOT, golden_lut
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libstb_ref.so"))
    ref.ref_stbi_load_from_memory.restype = C.POINTER(C.c_ubyte)
    buf = open(os.path.join(REFERENCE, "Content/PrecomputedMaps/brdf_lut.png"), "rb").read()
    w, h = C.c_int(), C.c_int()
    p = ref.ref_stbi_load_from_memory(buf, len(buf), C.byref(w), C.byref(h))
    assert p
    texels = np.ctypeslib.as_array(p, (h.value, w.value, 4)).copy()
    ref.ref_stbi_free(p)
    assert np.array_equal(texels, golden_lut())
