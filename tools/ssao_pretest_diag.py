#!/usr/bin/env python
"""How often the planar-neighbourhood pre-test of ssao_cull_kernel applies on the bench views (tuning aid): tiles whose reachable plane
records all decide and agree with the tile's own plane, and the rays of those tiles the one plane clears."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from althea_b200 import _capi, engine  # noqa: E402

ctx = engine.Context(0)
ibl, lights, views, _ = bench.build_rank_inputs(ctx, 0, 1, "cuda:0", quick_ibl=True)
stream = engine.current_stream_ptr(0)
ctx.set_flags(_capi.CTX_SSAO_COUNT_TAPS)
bench.run_frame(views[0], ibl, lights, stream)
torch.cuda.synchronize()
hist = (C.c_uint64 * 76)()
fn = C.CDLL(_capi.library_path()).althea_cuda_diag_ssao_cull_histogram
fn.argtypes = [C.c_void_p, C.c_void_p]
assert fn(ctx._ptr, hist) == 0
cleared, planar, other, kept = hist[72], hist[73], hist[74], hist[75]
print("S-scene bench view: planar tiles %d of %d (%.1f %%); rays of planar tiles cleared %d, kept %d (%.1f %% cleared)" % (
    planar, planar + other, 100.0 * planar / max(1, planar + other), cleared, kept, 100.0 * cleared / max(1, cleared + kept)))
print("cull counters", ctx.ssao_cull_counts())
