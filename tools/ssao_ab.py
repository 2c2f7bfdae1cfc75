#!/usr/bin/env python
"""Diagnostic: SSAO counts of the packed-proxy march vs the fp32-texel march on one 4K view, both builds.
  python tools/ssao_ab.py [bench|test]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from althea_b200 import _capi, engine, scene

which = sys.argv[1] if len(sys.argv) > 1 else "bench"
ctx = engine.Context(0)
ibl, lights, views, _ = bench.build_rank_inputs(ctx, 0, 1, "cuda:0", quick_ibl=True)
stream = engine.current_stream_ptr(0)
g, gb, ssr, dp = views[0]
if which == "test":
    W, H = 3840, 2160
    g = scene.make_uniforms(W, H, pos=(0.0, 2.0, 6.0), yaw=0.0, pitch=-0.25, light_count=0)
    gbd = scene.s_scene(g, W, H, scene.make_scene(64, device="cuda:0"), device="cuda:0")
    gb.upload(position=gbd.position, depth=gbd.depth, normal=gbd.normal, albedo=gbd.albedo, mro=gbd.mro)
    lights = None
res = {}
for name, flags in (("fast", 0), ("fast_ray", _capi.CTX_SSAO_RAY_DEPTH_PROXY), ("parity_ray", _capi.CTX_PARITY_MATH | _capi.CTX_SSAO_RAY_DEPTH_PROXY), ("fast_exact", _capi.CTX_SSAO_EXACT_TAPS), ("parity", _capi.CTX_PARITY_MATH),
                    ("parity_exact", _capi.CTX_PARITY_MATH | _capi.CTX_SSAO_EXACT_TAPS)):
    ctx.set_flags(flags)
    dp.aoCounts.tensor.zero_()
    dp.draw(g, gb, ibl, lights, ssr, _capi.SHADE_SKIP_TONEMAP, stream)
    torch.cuda.synchronize()
    res[name] = dp.aoCounts.tensor.clone()
for a, b in (("fast", "fast_exact"), ("parity", "parity_exact"), ("fast_ray", "fast_exact"), ("parity_ray", "parity_exact"), ("fast", "parity")):
    d = (res[a] != res[b])
    idx = d.nonzero().flatten()[:8].tolist()
    print(a, "vs", b, ": mismatching pixels", int(d.sum()), "of", d.numel(), "max |diff|", int((res[a].int() - res[b].int()).abs().max()),
          [(i % 3840, i // 3840, int(res[a][i]), int(res[b][i])) for i in idx])
print("lib", os.path.basename(_capi.library_path()))
