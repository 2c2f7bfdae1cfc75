#!/usr/bin/env python
"""Per-stage device timings of one 4K view (kernel tuning helper; bench.py is the contract benchmark).

  python tools/stage_bench.py [--iters 5] [--exact-taps] [--parity]
  ALTHEA_CUDA_LIB=althea_b200/lib/libalthea_cuda_<variant>.so python tools/stage_bench.py     # A/B a build variant
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from althea_b200 import _capi, engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--exact-taps", action="store_true")
    ap.add_argument("--parity", action="store_true")
    ap.add_argument("--mode-p", action="store_true", help="legacy position attachment as an input (default: mode D, what bench.py times)")
    args = ap.parse_args()
    ctx = engine.Context(0)
    flags = (_capi.CTX_PARITY_MATH if args.parity else 0) | (_capi.CTX_SSAO_EXACT_TAPS if args.exact_taps else 0)
    ibl, lights, views, _ = bench.build_rank_inputs(ctx, 0, args.views, "cuda:0", quick_ibl=True, with_position=args.mode_p)
    ctx.set_flags(flags)
    stream = engine.current_stream_ptr(0)
    for _ in range(2):
        for v in views:
            bench.run_frame(v, ibl, lights, stream)
    torch.cuda.synchronize()
    ctx.enable_timing(True)
    ctx.reset_timings()
    for _ in range(args.iters):
        for v in views:
            bench.run_frame(v, ibl, lights, stream)
    torch.cuda.synchronize()
    t = ctx.timings()
    n = args.iters * len(views)
    out = {k: round(v["total_ms"] / n, 4) for k, v in t.items()}
    out["frame_ms"] = round(sum(out.values()), 4)
    out["lib"] = os.path.basename(_capi.library_path())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
