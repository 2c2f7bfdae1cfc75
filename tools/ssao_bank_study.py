#!/usr/bin/env python
"""Offline count of the shared-memory wavefronts of ssao_cull_kernel's plane-record lookups (CPU, numpy; not part of the product).

A lookup is one LDS.128 per tap: 32 lanes read the 16-byte record of the block their tap falls in, out of a 32 x 32 window
(row pitch 512 bytes). The hardware serves a 128-bit load a quarter-warp at a time; two lanes of a quarter that read DIFFERENT
records in the same 16-byte bank group (record column mod 8) take separate wavefronts. This script replays the taps of the
bench view for sampled tiles and counts wavefronts per lookup for record layouts (plain rows; the column XOR-ed with bits of the
row), with and without the reuse predicate (a lane whose tap stays in the block of its previous tap does not load).

  python tools/ssao_bank_study.py [scene|room] [tiles]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ssao_cull_study as CS  # noqa: E402
import test_second_evaluation as T  # noqa: E402

W, H = CS.W, CS.H
REACH = 0.75


def wavefronts(bx, by, active, swz):
    """bx, by: (..., 32) block coordinates inside the window; active: (..., 32). Returns wavefronts per warp-level load."""
    col = swz(bx, by) & 7
    addr = by * 32 + bx
    total = np.zeros(bx.shape[:-1], np.int64)
    for q in range(4):
        sl = slice(8 * q, 8 * q + 8)
        a, c, m = addr[..., sl], col[..., sl], active[..., sl]
        worst = np.zeros(bx.shape[:-1], np.int64)
        for g in range(8):
            in_g = m & (c == g)
            # distinct addresses among the lanes of this bank group
            aa = np.where(in_g, a, -1)
            aa = np.sort(aa, -1)
            distinct = ((aa[..., 1:] != aa[..., :-1]) & (aa[..., 1:] >= 0)).sum(-1) + (aa[..., 0] >= 0)
            worst = np.maximum(worst, distinct)
        total += worst
    return total


def main(kind="scene", ntiles=300, seed=1):
    pos, nrm_u16, gm = CS.make_view(kind)
    proj, view, invP, invV = gm
    PV = proj @ view
    cam, fwd = invV[:3, 3], -invV[:3, 2]
    pos_img = pos[..., :3].astype(np.float64)
    nrm_img = T._f16(nrm_u16)[..., :3]
    rng = np.random.default_rng(seed)
    tx = rng.integers(0, W // 16, ntiles)
    ty = rng.integers(0, H // 16, ntiles)
    yy = (ty[:, None, None] * 16 + np.arange(16)[None, :, None]) + 0 * np.arange(16)[None, None, :]
    xx = (tx[:, None, None] * 16 + np.arange(16)[None, None, :]) + 0 * np.arange(16)[None, :, None]
    P = pos_img[yy, xx]
    covered = pos[yy, xx, 3] != 0
    N = T.normalize(np.where(covered[..., None], nrm_img[yy, xx], [0.0, 0.0, 1.0]))
    wide = np.abs(N[..., 0]) > np.abs(N[..., 1])
    with np.errstate(invalid="ignore", divide="ignore"):
        t_a = np.stack([-N[..., 2], 0 * N[..., 0], N[..., 0]], -1) / np.sqrt(N[..., 0] ** 2 + N[..., 2] ** 2)[..., None]
        t_b = np.stack([0 * N[..., 0], N[..., 2], -N[..., 1]], -1) / np.sqrt(N[..., 1] ** 2 + N[..., 2] ** 2)[..., None]
    Tn = np.where(wide[..., None], t_a, t_b)
    Bn = np.cross(N, Tn)
    xi = T.shader_rng(xx, yy, 72).reshape(24, 3, *xx.shape)
    tdepth = np.where(covered, (P - cam) @ fwd, np.inf)
    fpx = 0.5 * W * abs(proj[0, 0])
    tmin = tdepth.min((1, 2))
    keep = np.isfinite(tmin)
    reach = REACH * fpx / np.maximum(tmin - 0.5, 1e-3)
    level = np.where(reach <= 15 * 8 - 4, 0, np.where(reach <= 15 * 16 - 4, 1, 2))
    S = (8 << level)[:, None, None].astype(np.float64)
    wbx = ((tx * 16) >> (3 + level)) - 15
    wby = ((ty * 16) >> (3 + level)) - 15
    layouts = {"plain": lambda bx, by: bx, "col ^ row": lambda bx, by: bx ^ by, "col ^ 2 row": lambda bx, by: bx ^ (by << 1),
               "col + 3 row": lambda bx, by: bx + 3 * by, "col ^ row ^ row>>3": lambda bx, by: bx ^ by ^ (by >> 3)}
    tot = {(k, r): np.zeros(12, np.int64) for k in layouts for r in (False, True)}
    cnt = np.zeros(12, np.int64)
    act = np.zeros((12, 2), np.int64)
    uv0 = np.stack([(xx + 0.5) / W, (yy + 0.5) / H], -1)
    for ray in range(24):
        loc = T.normalize(np.stack([2 * xi[ray, 0] - 1, 2 * xi[ray, 1] - 1, xi[ray, 2]], -1))
        rd = Tn * loc[..., 0:1] + Bn * loc[..., 1:2] + N * loc[..., 2:3]
        pe = np.concatenate([P + 0.5 * rd, np.ones(xx.shape + (1,))], -1) @ PV.T
        with np.errstate(invalid="ignore", divide="ignore"):
            uv1 = 0.5 * pe[..., :2] / pe[..., 3:4] + 0.5
        prev = None
        for i in range(1, 12):
            t = i / 12.0
            uv = uv0 * (1 - t) + uv1 * t
            with np.errstate(invalid="ignore"):
                x, y = uv[..., 0] * W - 0.5, uv[..., 1] * H - 0.5
                bx = np.floor(x / S).astype(np.int64) - wbx[:, None, None]
                by = np.floor(y / S).astype(np.int64) - wby[:, None, None]
            bx, by = bx & 31, by & 31  # the kernel masks the offset; lanes of uncovered pixels load as well
            # a warp = two rows of 16 pixels
            bxw = bx.reshape(ntiles, 8, 32)
            byw = by.reshape(ntiles, 8, 32)
            live = np.broadcast_to(keep[:, None, None], bxw.shape)
            same = np.zeros_like(live) if prev is None else (bxw == prev[0]) & (byw == prev[1])
            prev = (bxw, byw)
            anyc = covered.reshape(ntiles, 8, 32).any(-1)  # warps without a covered pixel do not run
            for k, f in layouts.items():
                for reuse in (False, True):
                    a = live & ~same if reuse else live
                    w = wavefronts(bxw, byw, a, f)
                    tot[(k, reuse)][i] += int(w[anyc & keep[:, None]].sum())
            cnt[i] += int((anyc & keep[:, None]).sum())
            act[i, 0] += int((live & ~same)[anyc & keep[:, None]].sum())
    print("kind %s: %d tiles, %d warp-level lookups per tap index" % (kind, ntiles, cnt[1]))
    print("active lanes per lookup with the reuse predicate, taps 1..11: " + " ".join("%.1f" % (act[i, 0] / max(cnt[i], 1)) for i in range(1, 12)))
    for (k, reuse), v in tot.items():
        per = [v[i] / max(cnt[i], 1) for i in range(1, 12)]
        print("%-20s reuse %d: mean %.2f wavefronts per lookup   per tap: %s" % (k, reuse, sum(v[1:]) / max(sum(cnt[1:]), 1), " ".join("%.1f" % p for p in per)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "scene", int(sys.argv[2]) if len(sys.argv) > 2 else 300)
