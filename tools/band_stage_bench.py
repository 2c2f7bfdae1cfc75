#!/usr/bin/env python
"""Per-kernel timings of ONE band of the 8K S-rand frame (BASELINE configs[3]) on one GPU: what a rank of an N-way split runs.
  python tools/band_stage_bench.py [world] [rank]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from althea_b200 import _capi, bands, engine, scene

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = engine.Context(0)
ibl, lights, _, _ = bench.build_rank_inputs(ctx, 0, 0, "cuda:0", quick_ibl=True)
W8, H8 = 7680, 4320
g = scene.make_uniforms(W8, H8, pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0, light_count=bench.N_LIGHTS)
gbd = scene.s_rand(g, W8, H8, device="cuda:0")
gb = engine.GBufferResources(ctx, W8, H8, with_position=False)
gb.upload(depth=gbd.depth, normal=gbd.normal, albedo=gbd.albedo, mro=gbd.mro)
del gbd
out = {}
for w, r in ((1, 0), (world, rank)):
    bf = bands.BandedFrame(ctx, W8, H8, rank=r, world=w)
    stream = engine.current_stream_ptr(0)
    bf.render(g, gb, ibl, lights, _capi.SHADE_SKIP_TONEMAP, stream)
    torch.cuda.synchronize()
    ctx.enable_timing(True); ctx.reset_timings()
    for _ in range(2):
        bf.render(g, gb, ibl, lights, _capi.SHADE_SKIP_TONEMAP, stream)
    torch.cuda.synchronize()
    t = {k: round(v["total_ms"] / 2, 3) for k, v in ctx.timings().items()}
    ctx.enable_timing(False)
    t["total"] = round(sum(t.values()), 3)
    out["world%d_rank%d" % (w, r)] = t
    del bf
print(json.dumps(out, indent=1))
