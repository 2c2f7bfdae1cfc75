#!/usr/bin/env python
"""Diagnostic: SSR with the plane-record sign test vs the plain march on one 4K view, both builds: reflection mip 0 compared
bit for bit, per-kernel timings.   python tools/ssr_ab.py [bench|test|rand]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from althea_b200 import _capi, engine, scene

which = sys.argv[1] if len(sys.argv) > 1 else "bench"
ctx = engine.Context(0)
ibl, lights, views, _ = bench.build_rank_inputs(ctx, 0, 1, "cuda:0", quick_ibl=True, with_position=False)
stream = engine.current_stream_ptr(0)
g, gb, ssr, dp = views[0]
W, H = 3840, 2160
if which != "bench":
    if which == "test":
        g = scene.make_uniforms(W, H, pos=(0.0, 2.0, 6.0), yaw=0.0, pitch=-0.25, light_count=0)
        gbd = scene.s_scene(g, W, H, scene.make_scene(64, device="cuda:0"), device="cuda:0")
    else:
        g = scene.make_uniforms(W, H, pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0, light_count=0)
        gbd = scene.s_rand(g, W, H, device="cuda:0")
    gb.upload(depth=gbd.depth, normal=gbd.normal, albedo=gbd.albedo, mro=gbd.mro)
    lights = None
res = {}
n0 = W * H * 8
for name, flags in (("fast_skip", _capi.CTX_SSR_PLANE_SKIP), ("fast_march", 0), ("parity_skip", _capi.CTX_PARITY_MATH | _capi.CTX_SSR_PLANE_SKIP), ("parity_march", _capi.CTX_PARITY_MATH)):
    ctx.set_flags(flags)
    for it in range(3):
        if it == 2:
            ctx.enable_timing(True)
            ctx.reset_timings()
        ssr.captureReflection(g, gb, ibl, lights, stream)
        torch.cuda.synchronize()
    t = {k: round(v["total_ms"], 4) for k, v in ctx.timings().items() if k.startswith("ssr")}
    ctx.enable_timing(False)
    res[name] = ssr.getReflectionBuffer().image.tensor[:n0].clone().view(torch.int16).view(H, W, 4)
    hit = (res[name][..., 3] != 0).float().mean().item()
    print(name, t, "hit %.4f" % hit, flush=True)
for a, b in (("fast_skip", "fast_march"), ("parity_skip", "parity_march"), ("fast_skip", "parity_skip")):
    d = (res[a] != res[b]).any(-1)
    m = ((res[a][..., 3] != 0) != (res[b][..., 3] != 0))
    idx = d.view(-1).nonzero().flatten()[:6].tolist()
    print(a, "vs", b, ": pixels differing", int(d.sum()), "hit mask differing", int(m.sum()), "of", W * H, [(i % W, i // W) for i in idx])
ctx.set_flags(_capi.CTX_SSAO_COUNT_TAPS | _capi.CTX_SSR_PLANE_SKIP)
ssr.captureReflection(g, gb, ibl, lights, stream)
torch.cuda.synchronize()
c = ctx.ssao_cull_counts()
und = c["exact_steps"] & 0xffffffff
print("taps skipped %d (spans %d), exact %d (undecided by the records %d), warp-level exact steps %d" % (c["records"], c["exact_taps"], c["plane_lookups"], und, c["exact_steps"] >> 32))
print("lib", os.path.basename(_capi.library_path()))
