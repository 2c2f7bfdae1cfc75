#!/usr/bin/env python
"""IBL precompute timings (device events per kernel): the reference's shape (equirect, hash RNG) and BASELINE configs[1]
(32^2 irradiance cube + 512^2 6-mip GGX prefilter cube, Hammersley + 512^2 BRDF LUT).
  python tools/ibl_bench.py [--env 4096] [--samples 10000] [--which ref|cube|both]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from althea_b200 import _capi, engine, scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=10000)
    ap.add_argument("--which", default="both")
    ap.add_argument("--full-irradiance", action="store_true")
    args = ap.parse_args()
    W, H = args.env, args.env // 2
    ctx = engine.Context(0)
    F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
    env_t = scene.procedural_env(W, H, "cuda:0")
    mips = 1 + int(torch.log2(torch.tensor(float(W))).floor())
    chain = ctx.new_image(F32, W, H, mips)
    chain.tensor[: W * H * 16].copy_(env_t.view(-1).view(torch.uint8))
    engine.ImageBasedLighting.generateMipMaps(ctx, chain)
    out = {}

    def timed(name, fn):
        fn()  # warm-up
        ctx.synchronize(); torch.cuda.synchronize()
        ctx.enable_timing(True); ctx.reset_timings()
        fn()
        torch.cuda.synchronize()
        t = ctx.timings()
        ctx.enable_timing(False)
        out[name] = {k: round(v["total_ms"], 3) for k, v in t.items()}

    if args.which in ("ref", "both"):
        pre = ctx.new_image(F32, W >> 1, H >> 1, 5)
        irr = ctx.new_image(F32, 512, 256)
        timed("reference_shape_prefilter_5mips_hash", lambda: engine.ImageBasedLighting.precomputeResources(ctx, chain, None, pre, prefilter_samples=args.samples))
        timed("irradiance_equirect_512x256", lambda: engine.ImageBasedLighting.precomputeResources(ctx, chain, irr, None))
        if args.full_irradiance:  # the reference's own shape: an irradiance image the size of the environment map (ImageBasedLighting.cpp:315-343)
            irrf = ctx.new_image(F32, W, H)
            timed("irradiance_equirect_%dx%d_reference_shape" % (W, H), lambda: engine.ImageBasedLighting.precomputeResources(ctx, chain, irrf, None))
    if args.which in ("cube", "both"):
        prec = ctx.new_image(F32, 512, 512, 6, 6)
        irrc = ctx.new_image(F32, 32, 32, 1, 6)
        lut = ctx.new_image(_capi.FORMAT_R8G8B8A8_UNORM, 512, 512)
        timed("config1_prefilter_cube512_6mips_hammersley", lambda: engine.ImageBasedLighting.precomputeResources(
            ctx, chain, None, prec, layout=_capi.IBL_LAYOUT_CUBE, sequence=_capi.IBL_SEQ_HAMMERSLEY, prefilter_samples=args.samples))
        timed("config1_irradiance_cube32", lambda: engine.ImageBasedLighting.precomputeResources(ctx, chain, irrc, None, layout=_capi.IBL_LAYOUT_CUBE))
        timed("config1_brdf_lut_512", lambda: engine.ImageBasedLighting.generateBrdfLut(ctx, lut, 1024))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
