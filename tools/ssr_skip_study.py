#!/usr/bin/env python
"""Offline feasibility study for DESIGN.md section 8 item 2 (CPU, numpy float64; not part of the product).

SSR's hit test at a step needs `f = dot(normalize(tap - origin), rayDir) > 0.999`: the tapped surface must lie inside a narrow
cone around the reflected ray. For a tap position on the screen that confines the surface to an interval of distances along the
view ray through the tap, computable from the ray alone (a quadratic in the distance). If the scene's depth range over the
footprint of a SPAN of consecutive steps misses those intervals, none of the span's steps can hit and the span's taps can be
skipped without changing the hit mask. This script counts, on the test scenes, how many 8-step spans a conservative min/max
depth filter proves hit-free, and checks the proof against the exact march (no span declared hit-free may contain a step
with f > 0.999).

  python tools/ssr_skip_study.py [W H]
"""
import os
import sys

import numpy as np
from scipy.ndimage import maximum_filter, minimum_filter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_second_evaluation as T  # noqa: E402
from helpers import FrameData  # noqa: E402

SPAN = 8
C = 0.999
PYRAMID = False   # True: bounds from a min/max mip pyramid (2 x 2 blocks of one level) instead of an exact sliding window


def view_ray(g, u, v):
    """Unit direction through (u, v) and the factor that turns the reconstructed eye depth into a distance along it
    (Shaders/Misc/ReconstructPosition.glsl:4-22): position = cam + (d / f) * dir."""
    invP, invV = T._mat(g.inverseProjection), T._mat(g.inverseView)
    plane = np.stack([2 * u - 1, 2 * v - 1, np.full_like(u, 2.0), np.ones_like(u)], -1)
    dh = plane @ invP.T
    d = T.normalize((dh[..., :3] / dh[..., 3:4]) @ invV[:3, :3].T)
    f = d @ invV[:3, 2]
    # the z = 2 plane point has w < 0, so the shader's direction points away from the scene and d / f comes out negative;
    # flip both so that the distance along the ray is positive
    back = f > 0
    return np.where(back[..., None], -d, d), np.where(back, -f, f)


def distance_along_ray(d_raw, f):
    near, far = 0.01, 1000.0
    return far * near / (d_raw * (far - near) - far) / f


def cone_interval(a, d, rd):
    """Distances t along the view ray (origin cam, unit direction d) at which q = a + t d satisfies dot(normalize(q), rd) > C.
    a = cam - origin. Returns (lo, hi), empty where lo >= hi."""
    al, be = np.einsum("...i,...i->...", a, rd), np.einsum("...i,...i->...", d, rd)
    ad, aa = np.einsum("...i,...i->...", a, d), np.einsum("...i,...i->...", a, a)
    A, B, Cc = be * be - C * C, 2 * (al * be - C * C * ad), al * al - C * C * aa
    disc = B * B - 4 * A * Cc
    lo, hi = np.full(al.shape, np.inf), np.full(al.shape, -np.inf)
    ok = disc > 0
    sq = np.sqrt(np.where(ok, disc, 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        r0, r1 = (-B - sq) / (2 * A), (-B + sq) / (2 * A)
    r0, r1 = np.minimum(r0, r1), np.maximum(r0, r1)
    # A < 0 (view ray not inside the cone's opening): the quadratic is positive between the roots
    inside = ok & (A < 0)
    lo, hi = np.where(inside, r0, lo), np.where(inside, r1, hi)
    # A >= 0: positive outside the roots; keep it simple and conservative: anything in front may be inside
    wide = ok & (A >= 0)
    lo, hi = np.where(wide, 0.0, lo), np.where(wide, np.inf, hi)
    # the half-space dot(q, rd) > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        tz = -al / be
    lo = np.where(be > 0, np.maximum(lo, tz), lo)
    hi = np.where(be < 0, np.minimum(hi, tz), hi)
    return np.maximum(lo, 0.0), hi


def study(kind, W, H):
    fd = FrameData(kind, W, H, n_lights=0)
    g = fd.uniforms
    valid, hit, steps, *_ = T.ssr_march64(fd)
    invV = T._mat(g.inverseView)
    cam = invV[:3, 3]
    depth = fd.depth.astype(np.float64)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    u, v = (xx + 0.5) / W, (yy + 0.5) / H
    P = T.reconstruct_position64(g, u, v, depth)
    nrm = T._f16(fd.normal)
    N = T.normalize(np.where(valid[..., None], nrm[..., :3], [0.0, 0.0, 1.0]))
    vd, _ = view_ray(g, u, v)
    # the march's own direction uses the z = 0 plane point (SSR.vert), same direction up to normalisation
    rd = vd - 2 * np.einsum("hwi,hwi->hw", N, vd)[..., None] * N
    PV = T._mat(g.projection) @ T._mat(g.view)
    pe = np.concatenate([P + rd * 10000.0, np.ones((H, W, 1))], -1) @ PV.T
    uv_end = 0.5 * pe[..., :2] / pe[..., 3:4] + 0.5
    uv0 = np.stack([u, v], -1)
    step = T.normalize(uv_end - uv0) * 0.005
    # conservative scene depth range over a span's footprint: a square filter as large as the span can reach, plus the
    # bilinear footprint
    size = int(np.ceil(SPAN * 0.005 * max(W, H))) + 3
    if PYRAMID:
        # what a kernel would have: a min/max mip pyramid; a window of `size` texels is covered by the 2 x 2 blocks of the level
        # whose blocks are at least that large, i.e. a sliding (2 * block)-sized window aligned to the block grid
        level = int(np.ceil(np.log2(size)))
        blk = 1 << level
        hb, wb = -(-H // blk), -(-W // blk)
        pad = np.pad(depth, ((0, hb * blk - H), (0, wb * blk - W)), mode="edge")
        bmin = pad.reshape(hb, blk, wb, blk).min(axis=(1, 3))
        bmax = pad.reshape(hb, blk, wb, blk).max(axis=(1, 3))
        # block pair starting at the block that holds (centre - blk / 2)
        def lookup(b, my, mx, red):
            y0 = np.clip((my - blk // 2) // blk, 0, hb - 1)
            x0 = np.clip((mx - blk // 2) // blk, 0, wb - 1)
            y1, x1 = np.minimum(y0 + 1, hb - 1), np.minimum(x0 + 1, wb - 1)
            return red(red(b[y0, x0], b[y0, x1]), red(b[y1, x0], b[y1, x1]))
    else:
        dmin, dmax = minimum_filter(depth, size=size, mode="nearest"), maximum_filter(depth, size=size, mode="nearest")
    spans = skipped = wrong = taps_total = taps_skipped = 0
    for s0 in range(0, 128, SPAN):
        alive = valid & (steps > s0)                      # pixels whose march reaches this span
        if not alive.any():
            break
        mid = uv0 + step * (s0 + 1 + (SPAN - 1) / 2.0)
        mx = np.clip((mid[..., 0] * W).astype(int), 0, W - 1)
        my = np.clip((mid[..., 1] * H).astype(int), 0, H - 1)
        if PYRAMID:
            lo_d, hi_d = lookup(bmin, my, mx, np.minimum), lookup(bmax, my, mx, np.maximum)
        else:
            lo_d, hi_d = dmin[my, mx], dmax[my, mx]
        hit_free = alive.copy()
        any_candidate = np.zeros((H, W), bool)
        for k in range(SPAN):
            i = s0 + k
            uv = uv0 + step * (i + 1)
            inside = (uv[..., 0] >= 0) & (uv[..., 0] <= 1) & (uv[..., 1] >= 0) & (uv[..., 1] <= 1)
            d, f = view_ray(g, uv[..., 0], uv[..., 1])
            t_lo, t_hi = cone_interval(cam - P, d, rd)
            s_lo, s_hi = distance_along_ray(lo_d, f), distance_along_ray(hi_d, f)
            s_lo, s_hi = np.minimum(s_lo, s_hi), np.maximum(s_lo, s_hi)
            overlap = (t_lo <= s_hi) & (t_hi >= s_lo)
            hit_free &= ~(overlap & inside)
            # ground truth for this step
            d_raw = T.bilinear_clamp(depth[..., None], uv[..., 0], uv[..., 1])[..., 0]
            cur = T.reconstruct_position64(g, uv[..., 0], uv[..., 1], d_raw)
            fval = np.einsum("hwi,hwi->hw", T.normalize(cur - P), rd)
            any_candidate |= alive & inside & (steps > i) & (fval > C)
        n_taps = np.clip(steps - s0, 0, SPAN)
        spans += alive.sum()
        skipped += hit_free.sum()
        wrong += (hit_free & any_candidate).sum()
        taps_total += n_taps[alive].sum()
        taps_skipped += n_taps[hit_free].sum()
    print("%s%s %dx%d: %d-step spans proven hit-free %.1f %% (taps saved %.1f %%, minus one re-seed tap per skipped span: %.1f %%); "
          "unsound proofs: %d" % (kind, " (pyramid)" if PYRAMID else "", W, H, SPAN, 100.0 * skipped / spans, 100.0 * taps_skipped / taps_total,
                                  100.0 * (taps_skipped - skipped) / taps_total, wrong))
    return wrong


if __name__ == "__main__":
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (384, 216)
    bad = study("scene", W, H) + study("rand", W // 2, H // 2)
    PYRAMID = True
    bad += study("scene", W, H)
    sys.exit(1 if bad else 0)
