#!/usr/bin/env python
"""Generates tests/golden/shader_ref.npz: what the REFERENCE'S OWN SHADER TEXT yields on two seeded frames, executed on the CPU
through oracle/_ref/libshader_ref.so (oracle/glsl2cpp.py + oracle/glsl_compat.h + oracle/ref_shader_driver.cpp).

Run in the build container, where /root/reference is mounted (the library cannot be built elsewhere):
    python tests/golden/make_shader_golden.py
The frames are rebuilt from their seeds by tests/test_shader_ref.py (helpers.FrameData), so only outputs are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import FrameData, golden_env  # noqa: E402
from oracle import oracle as O, shader_ref as S  # noqa: E402

FRAMES = {"scene": ("scene", 128, 72), "rand": ("rand", 96, 54)}  # name -> FrameData(kind, W, H), 4 lights, 64^2 shadow cubes


def main():
    out = {}
    for name, (kind, W, H) in FRAMES.items():
        fr = FrameData(kind, W, H, n_lights=4, shadow_res=64).oracle_frame()
        out[name + "_ao"] = S.ssao(fr)
        refl, hit = S.ssr_capture(fr)
        out[name + "_refl"], out[name + "_hit"] = refl, hit
        # the same stage with the two choices GLSL / the rasteriser leave open aligned with the restatement's (one fma in
        # ReconstructPosition.glsl:8, the varying evaluated per pixel): what must then agree bit for bit
        out[name + "_refl_aligned"], _ = S.ssr_capture(fr, fused_reconstruct=True, closed_form_direction=True)
        chain = S.glossy_convolve(refl)
        out[name + "_chain"] = chain
        out[name + "_color_linear"] = S.deferred_shade(fr, chain, flags=O.SKIP_TONEMAP)
        out[name + "_color_linear_aligned"] = S.deferred_shade(fr, chain, flags=O.SKIP_TONEMAP, closed_form_direction=True)
        if name == "scene":
            out[name + "_color_tonemapped"] = S.deferred_shade(fr, chain, flags=0)
        out[name + "_dir"] = S.view_directions(fr.g, W, H)
    # IBL_Precompute: GenIrradianceMap.comp and PreFilterEnvMap.comp at probe texels, in the reference's own layout (equirect output
    # of the environment's size / of level i + 1, hash RNG, 300 x 150 and 10 000 samples) on the committed 512 x 256 environment
    env = golden_env()
    H, W = env.shape[:2]
    chain, mips = O.env_mip_chain(env)
    rng = np.random.default_rng(3)
    tex = np.stack([rng.integers(0, W, 12), rng.integers(0, H, 12), np.zeros(12, int)], -1)
    tex[0], tex[1], tex[2] = (0, 0, 0), (W - 1, H - 1, 0), (W // 2, 0, 0)
    out["irr_texels"], out["irr"] = tex, S.ibl_irradiance(chain, W, H, mips, W, H, tex)
    for i, r in enumerate((0.0, 0.25, 0.5, 0.75, 1.0)):
        ow, oh = W >> (i + 1), H >> (i + 1)
        tx = np.stack([rng.integers(0, ow, 8), rng.integers(0, oh, 8), np.zeros(8, int)], -1)
        out["pre%d_texels" % i], out["pre%d" % i] = tx, S.ibl_prefilter(chain, W, H, mips, ow, oh, r, tx)
    # the rasterising producers: liboracle.so's fixed-function rasteriser with the PROGRAMMABLE stages taken from the shader text
    # (Gltf/Gltf.vert + .frag, ShadowMapBindless.vert + .frag) through its stage hooks
    from producer_scene import gbuffer_case, shadow_case
    S.set_raster_stage_hooks(True)
    try:
        proj, view, prims, W, H = gbuffer_case()
        gb = O.draw_gbuffer(proj, view, prims, W, H)
        for k, v in gb.items():
            out["gbuffer_" + k] = v
        lights, projection, views, sc, res = shadow_case()
        out["shadow_cubes"] = O.draw_shadow_cubes(lights, projection, views, sc, res)
    finally:
        S.set_raster_stage_hooks(False)
    path = os.path.join(HERE, "shader_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
