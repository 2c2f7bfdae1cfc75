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

from helpers import FrameData  # noqa: E402
from oracle import oracle as O, shader_ref as S  # noqa: E402

FRAMES = {"scene": ("scene", 128, 72), "rand": ("rand", 96, 54)}  # name -> FrameData(kind, W, H), 4 lights, 64^2 shadow cubes


def main():
    out = {}
    for name, (kind, W, H) in FRAMES.items():
        fr = FrameData(kind, W, H, n_lights=4, shadow_res=64).oracle_frame()
        out[name + "_ao"] = S.ssao(fr)
        refl, hit = S.ssr_capture(fr)
        out[name + "_refl"], out[name + "_hit"] = refl, hit
        chain = S.glossy_convolve(refl)
        out[name + "_chain"] = chain
        out[name + "_color_linear"] = S.deferred_shade(fr, chain, flags=O.SKIP_TONEMAP)
        if name == "scene":
            out[name + "_color_tonemapped"] = S.deferred_shade(fr, chain, flags=0)
        out[name + "_dir"] = S.view_directions(fr.g, W, H)
    path = os.path.join(HERE, "shader_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
