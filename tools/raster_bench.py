#!/usr/bin/env python
"""Device timings of the rasterising producers on a synthetic mesh scene at 4K (tuning helper, not the contract bench).

  python tools/raster_bench.py [--stacks 512] [--iters 5] [--lights 16]
Scene: one finely tessellated textured sphere (2 * stacks^2 * 2 triangles), a small sphere, and a room of six large quads
(large-triangle path: one work item per 64 x 64 tile)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from althea_b200 import engine, model, scene  # noqa: E402


def build_scene(stacks):
    rng = np.random.default_rng(5)
    big = model.uv_sphere(1.0, (0.0, 0.0, 0.0), stacks, 2 * stacks)
    big.material = model.MaterialData(metallicFactor=0.7, roughnessFactor=0.9)
    big.material.baseTexture = model.checker_texture(1024, 64)
    nm = np.zeros((512, 512, 4), np.uint8)
    nm[..., :2] = rng.integers(96, 160, (512, 512, 2))
    nm[..., 2:] = 255
    big.material.normalTexture = model.TextureData.from_rgba8(nm, model.sampler_word())
    big.material.metallicRoughnessTexture = model.TextureData.from_rgba8(rng.integers(0, 256, (256, 256, 4)).astype(np.uint8), model.sampler_word())
    small = model.uv_sphere(0.5, (1.6, -0.5, 0.8), 64, 128)
    room = [model.quad(c, 8.0) for c in (
        [(-6, -1, 6), (6, -1, 6), (6, -1, -6), (-6, -1, -6)], [(-6, 5, -6), (6, 5, -6), (6, 5, 6), (-6, 5, 6)],
        [(-6, -1, -5), (6, -1, -5), (6, 6, -5), (-6, 6, -5)], [(6, -1, 6), (-6, -1, 6), (-6, 6, 6), (6, 6, 6)],
        [(-6, -1, 6), (-6, -1, -6), (-6, 6, -6), (-6, 6, 6)], [(6, -1, -6), (6, -1, 6), (6, 6, 6), (6, 6, -6)])]
    for q in room:
        q.material.baseTexture = big.material.baseTexture
    return [big, small] + room


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stacks", type=int, default=512)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--lights", type=int, default=16)
    ap.add_argument("--size", type=int, nargs=2, default=[3840, 2160])
    args = ap.parse_args()
    W, H = args.size
    ctx = engine.Context(0)
    prims = build_scene(args.stacks)
    up = model.UploadedModel(ctx, prims)
    g = scene.make_uniforms(W, H, pos=(0.3, 0.8, 3.5), yaw=0.1, pitch=-0.2, light_count=args.lights)
    gb = engine.GBufferResources(ctx, W, H)
    gpass = engine.SceneToGBufferPass(ctx)
    lights = engine.PointLightCollection(ctx, args.lights, shadow_res=256)
    rng = np.random.default_rng(1)
    for i in range(args.lights):
        lights.setLight(i, engine.PointLight((rng.uniform(-4, 4), rng.uniform(0, 4), rng.uniform(-3, 4)), (10.0, 10.0, 10.0)))
    for _ in range(2):
        gpass.draw(g, up, gb)
        lights.drawShadowMaps([up])
    torch.cuda.synchronize()
    out = {"triangles": up.triangle_count, "size": [W, H], "lights": args.lights}
    for name, fn in (("gbuffer", lambda: gpass.draw(g, up, gb)), ("shadow_cubes", lambda: lights.drawShadowMaps([up]))):
        ctx.enable_timing(True)
        ctx.reset_timings()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = ctx.timings()
        ctx.enable_timing(False)
        out[name] = {"ms_per_call_wall": round(e0.elapsed_time(e1) / args.iters, 3),
                     "kernels_ms": {k: round(v["total_ms"] / args.iters, 3) for k, v in t.items()}}
    cov = float((gb.depth.tensor.view(torch.float32) < 1).float().mean())
    out["coverage"] = round(cov, 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
