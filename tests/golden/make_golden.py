"""Generates the committed fixtures under tests/golden/. Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

Outputs (all small):
  brdf_lut.npz      the reference's Content/PrecomputedMaps/brdf_lut.png, palette expanded to RGBA8 (the a10 golden)
  env_512x256.hdr   Content/HDRI_Skybox/NeoclassicalInterior.hdr box-downsampled 8x, RGBE (test environment)
  ibl_pin.json      probe texels of the reference's shipped IrradianceMap.hdr / Prefiltered1..5.hdr for all three
                    environments, with the oracle's value at the same texel computed from the full 4096x2048 input:
                    the record that the oracle is pinned to the reference's own outputs (the inputs are too large to commit)
  frame_golden.npz  oracle outputs for two small synthetic frames (SSR hit mask + reflection, AO counts, shaded colour):
                    lets the GPU tests compare against committed vectors as well as against the live oracle
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import hdrio  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/Content"
ENVS = ["NeoclassicalInterior", "LuxuryRoom", "ThatchChapel"]


def make_lut():
    from PIL import Image
    a = np.array(Image.open(os.path.join(REF, "PrecomputedMaps/brdf_lut.png")).convert("RGBA"))
    np.savez_compressed(os.path.join(HERE, "brdf_lut.npz"), lut=a)
    print("brdf_lut", a.shape)


def make_env():
    env = hdrio.read_hdr(os.path.join(REF, "HDRI_Skybox/NeoclassicalInterior.hdr"))
    H, W = env.shape[:2]
    small = env.reshape(H // 8, 8, W // 8, 8, 3).astype(np.float64).mean(axis=(1, 3)).astype(np.float32)
    hdrio.write_hdr(os.path.join(HERE, "env_512x256.hdr"), small)
    back = hdrio.read_hdr(os.path.join(HERE, "env_512x256.hdr"))
    print("env", small.shape, "rgbe roundtrip max rel", float(np.max(np.abs(back - small) / np.maximum(small, 1e-3))))


def make_ibl_pin():
    rng = np.random.default_rng(0xA17EA)
    report = {"note": "oracle vs the reference's shipped precompute outputs; residual = RGBE truncation (<= 1/128 of the "
                      "largest channel, more on small channels)", "envs": {}}
    for name in ENVS:
        env = hdrio.read_hdr(os.path.join(REF, "HDRI_Skybox/%s.hdr" % name))
        H, W = env.shape[:2]
        rgba = np.concatenate([env, np.ones((H, W, 1), np.float32)], -1)
        chain, mips = O.env_mip_chain(rgba)
        entry = {"size": [W, H], "mips": mips, "irradiance": [], "prefiltered": {}}
        irr = hdrio.read_hdr(os.path.join(REF, "PrecomputedMaps/%s/IrradianceMap.hdr" % name))
        probes = [(100, 100), (2048, 1024), (3000, 500), (1000, 1800), (4000, 1024), (17, 2000)] + [
            (int(rng.integers(0, W)), int(rng.integers(0, H))) for _ in range(10)]
        out = O.ibl_irradiance(chain, W, H, mips, W, H, [(x, y, 0) for x, y in probes])
        for (x, y), o in zip(probes, out):
            entry["irradiance"].append({"texel": [x, y], "shipped": irr[y, x].tolist(), "oracle": o[:3].tolist()})
        for i in range(1, 6):
            pf = hdrio.read_hdr(os.path.join(REF, "PrecomputedMaps/%s/Prefiltered%d.hdr" % (name, i)))
            h, w = pf.shape[:2]
            pts = [(int(rng.integers(0, w)), int(rng.integers(0, h)), 0) for _ in range(24)]
            out = O.ibl_prefilter(chain, W, H, mips, w, h, (i - 1) / 4.0, pts)
            entry["prefiltered"][str(i)] = {"size": [w, h], "roughness": (i - 1) / 4.0, "probes": [
                {"texel": [p[0], p[1]], "shipped": pf[p[1], p[0]].tolist(), "oracle": o[:3].tolist()} for p, o in zip(pts, out)]}
        report["envs"][name] = entry
        print("pinned", name)
    with open(os.path.join(HERE, "ibl_pin.json"), "w") as f:
        json.dump(report, f, indent=1)


def make_frame_golden():
    import helpers
    out = {}
    for tag, kind, W, H in (("scene", "scene", 160, 90), ("rand", "rand", 96, 54)):
        fd = helpers.FrameData(kind, W, H, n_lights=4, shadow_res=32)
        fr = fd.oracle_frame()
        refl, hit, steps = O.ssr_capture(fr)
        chain = O.glossy_convolve(refl)
        ao = O.ssao(fr)
        col = O.deferred_shade(fr, chain, 5, O.SKIP_TONEMAP, ao)
        out[tag + "_refl_chain"] = chain
        out[tag + "_hit"] = hit
        out[tag + "_ao"] = ao
        out[tag + "_color"] = col
        print(tag, "hit frac", hit.mean(), "ao mean", ao[ao < 255].mean() if (ao < 255).any() else 0)
    np.savez_compressed(os.path.join(HERE, "frame_golden.npz"), **out)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (run in the build container)")
    make_lut()
    make_env()
    make_ibl_pin()
    make_frame_golden()
