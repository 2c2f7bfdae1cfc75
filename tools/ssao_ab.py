#!/usr/bin/env python
"""Diagnostic: SSAO counts of the coarse-sign-test path vs the round-1 march vs the fp32-texel march on 4K views, both builds,
with per-kernel timings and the cull counters.
  python tools/ssao_ab.py [bench|test|rand]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from althea_b200 import _capi, engine, scene

which = sys.argv[1] if len(sys.argv) > 1 else "bench"
mode_p = "modep" in sys.argv
ctx = engine.Context(0)
ibl, lights, views, _ = bench.build_rank_inputs(ctx, 0, 1, "cuda:0", quick_ibl=True, with_position=(mode_p or which != "bench"))
stream = engine.current_stream_ptr(0)
g, gb, ssr, dp = views[0]
W, H = 3840, 2160
if which == "test":
    g = scene.make_uniforms(W, H, pos=(0.0, 2.0, 6.0), yaw=0.0, pitch=-0.25, light_count=0)
    gbd = scene.s_scene(g, W, H, scene.make_scene(64, device="cuda:0"), device="cuda:0")
    gb.upload(position=gbd.position, depth=gbd.depth, normal=gbd.normal, albedo=gbd.albedo, mro=gbd.mro)
    lights = None
elif which == "rand":
    g = scene.make_uniforms(W, H, pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0, light_count=0)
    gbd = scene.s_rand(g, W, H, device="cuda:0")
    gb.upload(position=gbd.position, depth=gbd.depth, normal=gbd.normal, albedo=gbd.albedo, mro=gbd.mro)
    lights = None
res, times = {}, {}
NC = _capi.CTX_SSAO_NO_CULL
for name, flags in (("fast_cull", 0), ("fast_march", NC), ("fast_exact", _capi.CTX_SSAO_EXACT_TAPS), ("parity_cull", _capi.CTX_PARITY_MATH),
                    ("parity_march", _capi.CTX_PARITY_MATH | NC), ("parity_exact", _capi.CTX_PARITY_MATH | _capi.CTX_SSAO_EXACT_TAPS)):
    ctx.set_flags(flags)
    dp.aoCounts.tensor.zero_()
    for it in range(3):
        if it == 2:
            ctx.enable_timing(True)
            ctx.reset_timings()
        dp.draw(g, gb, ibl, lights, ssr, _capi.SHADE_SKIP_TONEMAP, stream)
        torch.cuda.synchronize()
    times[name] = {k: round(v["total_ms"], 4) for k, v in ctx.timings().items() if k.startswith("ssao") or k.startswith("reconstruct")}
    ctx.enable_timing(False)
    res[name] = dp.aoCounts.tensor.clone()
    print(name, times[name], flush=True)
for a, b in (("fast_cull", "fast_exact"), ("fast_march", "fast_exact"), ("parity_cull", "parity_exact"), ("parity_march", "parity_exact"), ("fast_cull", "parity_cull")):
    d = (res[a] != res[b])
    idx = d.nonzero().flatten()[:8].tolist()
    print(a, "vs", b, ": mismatching pixels", int(d.sum()), "of", d.numel(), "max |diff|", int((res[a].int() - res[b].int()).abs().max()),
          [(i % W, i // W, int(res[a].view(-1)[i]), int(res[b].view(-1)[i])) for i in idx])
ctx.set_flags(_capi.CTX_SSAO_COUNT_TAPS)
dp.draw(g, gb, ibl, lights, ssr, _capi.SHADE_SKIP_TONEMAP, stream)
torch.cuda.synchronize()
print("cull counters", ctx.ssao_cull_counts())
ctx.set_flags(_capi.CTX_SSAO_COUNT_TAPS | NC)
dp.draw(g, gb, ibl, lights, ssr, _capi.SHADE_SKIP_TONEMAP, stream)
torch.cuda.synchronize()
print("march counters", ctx.ssao_cull_counts())
cov = (res["fast_exact"] != 255)
print("covered", float(cov.float().mean()), "mean count", float(res["fast_exact"][cov].float().mean()))
print("lib", os.path.basename(_capi.library_path()))
try:
    import ctypes as C
    ctx.set_flags(_capi.CTX_SSAO_COUNT_TAPS)
    dp.draw(g, gb, ibl, lights, ssr, _capi.SHADE_SKIP_TONEMAP, stream)
    torch.cuda.synchronize()
    hist = (C.c_uint64 * 76)()
    fn = C.CDLL(_capi.library_path()).althea_cuda_diag_ssao_cull_histogram
    fn.argtypes = [C.c_void_p, C.c_void_p]
    if fn(ctx._ptr, hist) == 0:
        for lv in range(3):
            row = [(hist[(lv * 12 + i) * 2], hist[(lv * 12 + i) * 2 + 1]) for i in range(1, 12)]
            tot = sum(a for a, _ in row)
            if tot:
                print("level %d: %5.1f %% of the lookups; undecided per tap 1..11: %s; rays with no decided tap: %d" % (
                    lv, 100.0 * tot / max(1, sum(hist[(l * 12 + i) * 2] for l in range(3) for i in range(1, 12))),
                    " ".join("%.1f" % (100.0 * b / max(a, 1)) for a, b in row), hist[lv * 24]))
except Exception as e:  # tuning aid only
    print("histogram unavailable:", e)
