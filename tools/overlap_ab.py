#!/usr/bin/env python
"""A/B: one 4K frame's stages in stream order against the SSR chain and the SSAO pass on two streams (the reflection chain and the
occlusion counts are independent until the lighting pass reads both). Wall time per frame by CUDA events over a batch of frames.
  python tools/overlap_ab.py [--iters 10] [--views 4]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from althea_b200 import _capi, engine

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--views", type=int, default=4)
args = ap.parse_args()
ctx = engine.Context(0)
ibl, lights, views, _ = bench.build_rank_inputs(ctx, 0, args.views, "cuda:0", quick_ibl=True)
sA = torch.cuda.current_stream()
sB = torch.cuda.Stream()
T = _capi.SHADE_SKIP_TONEMAP


def serial(v):
    bench.run_frame(v, ibl, lights, sA.cuda_stream)


def overlapped(v):
    g, gb, ssr, dp = v
    sB.wait_stream(sA)
    dp.draw(g, gb, ibl, lights, ssr, T | _capi.SHADE_AO_ONLY, sB.cuda_stream)
    ssr.captureReflection(g, gb, ibl, lights, sA.cuda_stream)
    ssr.convolveReflectionBuffer(sA.cuda_stream)
    sA.wait_stream(sB)
    dp.draw(g, gb, ibl, lights, ssr, T | _capi.SHADE_AO_FROM_IMAGE, sA.cuda_stream)


def interleaved():  # view k's SSAO + lighting on stream B while view k + 1's reflection chain runs on stream A
    ev = []
    for v in views:
        g, gb, ssr, dp = v
        ssr.captureReflection(g, gb, ibl, lights, sA.cuda_stream)
        ssr.convolveReflectionBuffer(sA.cuda_stream)
        e = torch.cuda.Event()
        e.record(sA)
        sB.wait_event(e)
        dp.draw(g, gb, ibl, lights, ssr, T, sB.cuda_stream)
    sA.wait_stream(sB)


out = {}
for name, fn in (("serial", lambda: [serial(v) for v in views]), ("two_streams", lambda: [overlapped(v) for v in views]), ("interleaved", interleaved)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sA)
    for _ in range(args.iters):
        fn()
    e1.record(sA)
    torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) / (args.iters * len(views)), 4)
    col = views[0][3].colorTarget.tensor.clone()
    out[name + "_sum"] = float(col.view(torch.float16).float().nan_to_num().sum())
print(json.dumps(out))
