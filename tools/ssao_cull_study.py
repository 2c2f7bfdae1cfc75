#!/usr/bin/env python
"""Offline feasibility study for the SSAO coarse sign test (CPU, numpy float64; not part of the product).

A march step of computeSSAO (Shaders/SSAO.glsl:46-79) can only score when dot(p - pos, perpRef) changes sign between two
consecutive taps. For a G-buffer written by a perspective camera, texel (x, y) holds p = cam + D(x, y) t with D affine in
(x, y) and t the eye depth, so the projection is t (c0 w + g(x, y)) with w = 1 / t, c0 = dot(cam - pos, perpRef) and g
affine: its SIGN is the sign of w - L(x, y), L = -g / c0 affine on the screen (the ray's plane in reciprocal depth). A plane
surface has w affine on the screen as well, so a per-block plane fit of w with a residual range decides the sign of every
footprint texel of a tap from ONE small record (shared-memory resident), and the 32-byte position record only has to be
gathered where the block cannot decide or the sign changes.

This script measures, on the bench view, which share of the taps such a record decides, how many gathers remain, and checks
the decisions against the exact march (no decided sign may differ).

  python tools/ssao_cull_study.py [scene|room|rand] [block] [tiles]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_second_evaluation as T  # noqa: E402
from althea_b200 import scene  # noqa: E402

W, H = 3840, 2160
LOOSE = os.environ.get("LOOSE", "0") == "1"
MULTI = os.environ.get("MULTI", "")  # "ray": finest level whose window holds the ray, "tap": per tap, "tile": per tile (kernel)
UND = {}
PRE = os.environ.get("PRE", "0") == "1"
PRE_OVERLAP = os.environ.get("PRE_OVERLAP", "1") == "1"
PRE_BLOCKS = (64, 128, 256)
PRESTAT, PREOK, PREBAD = {}, {}, {}
REACH = float(os.environ.get("REACH", "0.75"))
ETA = os.environ.get("ETA", "0") == "1"
KERNEL = os.environ.get("KERNEL", "0") == "1"
SKY = os.environ.get("SKY", "1") == "1"


def make_view(kind):
    cache = "/tmp/ssao_cull_%s.npz" % kind
    if os.path.exists(cache):
        z = np.load(cache)
        return z["position"], z["normal"], z["g"]
    if kind == "rand":
        g = scene.make_uniforms(W, H, pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0)
        gb = scene.s_rand(g, W, H)
    else:
        g = scene.make_uniforms(W, H, pos=(0.0, 2.0, 0.0), yaw=0.0, pitch=-0.2)
        gb = scene.s_scene(g, W, H, scene.make_ring_scene(160))
    d = gb.numpy()
    gm = np.stack([T._mat(g.projection), T._mat(g.view), T._mat(g.inverseProjection), T._mat(g.inverseView)])
    np.savez(cache, position=d["position"], normal=d["normal"], g=gm)
    return d["position"], d["normal"], gm


def block_records(pos, cam, fwd, B, ox=0, oy=0):
    """Per B x B block (+ one texel of apron to the right / below): plane fit of w = 1 / eye depth and the residual range."""
    covered = pos[..., 3] != 0
    t = (pos[..., :3].astype(np.float64) - cam) @ fwd
    with np.errstate(divide="ignore"):
        w = np.where(covered, 1.0 / t, 0.0)
    nbx, nby = (W + ox + B - 1) // B, (H + oy + B - 1) // B
    # gather the (B + 1)^2 texels of every block (clamped at the image edge); block b covers texels b * B - o .. b * B - o + B
    ys = np.clip(np.arange(nby)[:, None] * B - oy + np.arange(B + 1)[None, :], 0, H - 1)   # (nby, B+1)
    xs = np.clip(np.arange(nbx)[:, None] * B - ox + np.arange(B + 1)[None, :], 0, W - 1)
    wb = w[ys[:, None, :, None], xs[None, :, None, :]]                                # (nby, nbx, B+1, B+1)
    cb = covered[ys[:, None, :, None], xs[None, :, None, :]]
    allc, anyc = cb.all((2, 3)), cb.any((2, 3))
    dx = np.arange(B + 1, dtype=np.float64)
    # least squares on the regular grid, coordinates relative to the block's corner texel
    mx = dx.mean()
    sxx = ((dx - mx) ** 2).sum() * (B + 1)
    wm = wb.mean((2, 3))
    beta = ((wb - wm[..., None, None]) * (dx - mx)[None, None, None, :]).sum((2, 3)) / sxx
    gamma = ((wb - wm[..., None, None]) * (dx - mx)[None, None, :, None]).sum((2, 3)) / sxx
    w0 = wm - beta * mx - gamma * mx
    res = wb - (w0[..., None, None] + beta[..., None, None] * dx[None, None, None, :] + gamma[..., None, None] * dx[None, None, :, None])
    rlo, rhi = res.min((2, 3)), res.max((2, 3))
    kind = np.where(allc, 1, np.where(anyc, 0, 2))   # 1 = decidable surface, 0 = mixed (never decides), 2 = all empty
    return dict(w0=w0, beta=beta, gamma=gamma, rlo=rlo, rhi=rhi, kind=kind, wmax=wb.max((2, 3)))


def study(kind="scene", B=8, ntiles=400, seed=1):
    pos, nrm_u16, gm = make_view(kind)
    proj, view, invP, invV = gm
    PV = proj @ view
    cam, fwd = invV[:3, 3], -invV[:3, 2]
    recs = {b: block_records(pos, cam, fwd, b) for b in ((8, 16, 32) if MULTI else (B,))}
    global recs_pre
    recs_pre = {pb: {(ox, oy): block_records(pos, cam, fwd, pb, ox, oy) for ox, oy in ((0, 0), (pb // 2, 0), (0, pb // 2), (pb // 2, pb // 2))} for pb in PRE_BLOCKS} if PRE else {}
    pos_img = pos[..., :3].astype(np.float64)
    nrm_img = T._f16(nrm_u16)[..., :3]
    # D(x, y): direction through texel (x, y) scaled to eye depth 1, affine in (x, y)
    def Dray(x, y):
        ndc = np.stack([2 * (x + 0.5) / W - 1, 2 * (y + 0.5) / H - 1, np.full_like(x, 2.0), np.ones_like(x)], -1)
        dh = ndc @ invP.T
        d = (dh[..., :3] / dh[..., 3:4]) @ invV[:3, :3].T
        return d / (d @ fwd)[..., None]
    z = np.zeros(1)
    Dc = Dray(z, z)[0]
    Dx = Dray(z + 1, z)[0] - Dc
    Dy = Dray(z, z + 1)[0] - Dc
    global DMAX1, DMAG
    DMAX1 = max(np.abs(Dc + Dx * cx + Dy * cy).sum() for cx in (0, W - 1) for cy in (0, H - 1))
    DMAG = np.abs(Dc).sum() + np.abs(Dx).sum() * W + np.abs(Dy).sum() * H
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
    tile_reach = REACH * fpx / np.maximum(tdepth.min((1, 2)) - 0.5, 1e-3)
    uv0 = np.stack([(xx + 0.5) / W, (yy + 0.5) / H], -1)
    shp = xx.shape
    stats = dict(taps_old=0, taps_all=0, decided=0, unsound=0, gathers_new=0, rays=0, rays_all_decided=0, hits=0,
                 warp_iters_old=0, warp_iters_coarse=0, warp_iters_detail=0, detail_steps=0)
    for ray in range(24):
        loc = T.normalize(np.stack([2 * xi[ray, 0] - 1, 2 * xi[ray, 1] - 1, xi[ray, 2]], -1))
        rd = Tn * loc[..., 0:1] + Bn * loc[..., 1:2] + N * loc[..., 2:3]
        end = P + 0.5 * rd
        pe = np.concatenate([end, np.ones(shp + (1,))], -1) @ PV.T
        uv1 = 0.5 * pe[..., :2] / pe[..., 3:4] + 0.5
        perp = T.normalize(np.cross(np.cross(rd, N), rd))
        c0 = np.einsum("...i,...i->...", cam - P, perp)
        ac, au, av = perp @ Dc, perp @ Dx, perp @ Dy
        sky_proj = np.einsum("...i,...i->...", -P, perp)
        with np.errstate(divide="ignore", invalid="ignore"):
            x11, y11 = (uv0[..., 0] / 12 + uv1[..., 0] * 11 / 12) * W - 0.5, (uv0[..., 1] / 12 + uv1[..., 1] * 11 / 12) * H - 0.5
            L11 = -(ac + au * x11 + av * y11) / c0
        # ---- per-ray pre-test: taps 1 and 11 inside ONE coarse block whose record decides both with the same sign
        if PRE:
            for PB in PRE_BLOCKS:
                rp = recs_pre[PB]
                xa, ya = (uv0[..., 0] * 11 / 12 + uv1[..., 0] / 12) * W - 0.5, (uv0[..., 1] * 11 / 12 + uv1[..., 1] / 12) * H - 0.5
                ok_any = np.zeros(shp, bool)
                for ox, oy in (((0, 0), (PB // 2, 0), (0, PB // 2), (PB // 2, PB // 2)) if PRE_OVERLAP else ((0, 0),)):
                    r_ = rp[(ox, oy)]
                    with np.errstate(invalid="ignore"):
                        bxa, bya = np.floor((xa + ox) / PB), np.floor((ya + oy) / PB)
                        bxb, byb = np.floor((x11 + ox) / PB), np.floor((y11 + oy) / PB)
                        same = (bxa == bxb) & (bya == byb) & np.isfinite(x11) & np.isfinite(y11) & (x11 >= 0) & (x11 <= W - 1) & (y11 >= 0) & (y11 <= H - 1)
                    bxi = np.clip(np.where(same, bxa, 0), 0, r_["kind"].shape[1] - 1).astype(int)
                    byi = np.clip(np.where(same, bya, 0), 0, r_["kind"].shape[0] - 1).astype(int)
                    k_ = r_["kind"][byi, bxi]
                    def dval(xq, yq):
                        wpl = r_["w0"][byi, bxi] + r_["beta"][byi, bxi] * (xq + ox - bxi * PB) + r_["gamma"][byi, bxi] * (yq + oy - byi * PB)
                        with np.errstate(divide="ignore", invalid="ignore"):
                            return wpl + 0.5 * (r_["rhi"][byi, bxi] + r_["rlo"][byi, bxi]) + (ac + au * xq + av * yq) / c0
                    rraw = 0.5 * (r_["rhi"][byi, bxi] - r_["rlo"][byi, bxi])
                    grec = np.abs(r_["beta"][byi, bxi]) + np.abs(r_["gamma"][byi, bxi])
                    with np.errstate(divide="ignore", invalid="ignore"):
                        gray = np.abs(au / c0) + np.abs(av / c0)
                        slack = 1.02 * rraw + 0.06 * (grec + gray) + 3e-5 * np.abs(r_["wmax"][byi, bxi]) / np.maximum(np.abs(c0), 1e-30) + 1e-6 * r_["wmax"][byi, bxi]
                        da, db = dval(xa, ya), dval(x11, y11)
                        ok = same & (k_ == 1) & (np.abs(da) > slack) & (np.abs(db) > slack) & (da * db > 0)
                    ok_any |= ok
                PRESTAT.setdefault(PB, [0, 0])
                PRESTAT[PB][0] += int((covered & ok_any).sum())
                PRESTAT[PB][1] += int(covered.sum())
                PREOK[PB] = ok_any
        prev_pos, prev_proj = P.copy(), np.zeros(shp)
        live = covered.copy()
        cls_prev = np.zeros(shp, np.int8)          # coarse class of the previous tap: +1 / -1 / 0 undecided
        have_prev = np.zeros(shp, bool)            # previous tap's record already gathered by the new scheme
        visited = np.zeros((12,) + shp, bool)
        detail = np.zeros((12,) + shp, bool)
        for i in range(1, 12):
            t = i / 12.0
            uv = uv0 * (1 - t) + uv1 * t
            inside = (uv[..., 0] >= 0) & (uv[..., 0] <= 1) & (uv[..., 1] >= 0) & (uv[..., 1] <= 1)
            live &= inside
            cur = T.bilinear_clamp(pos_img, uv[..., 0], uv[..., 1])
            pr = np.einsum("...i,...i->...", cur - P, perp)
            step = np.linalg.norm(cur - prev_pos, axis=-1)
            # --- coarse class of this tap
            x, y = uv[..., 0] * W - 0.5, uv[..., 1] * H - 0.5
            def classify(rec, B):
                ix, iy = np.clip(np.floor(x), 0, W - 1).astype(int), np.clip(np.floor(y), 0, H - 1).astype(int)
                bx, by = ix // B, iy // B
                k = rec["kind"][by, bx]
                lx, ly = x - bx * B, y - by * B
                wpl = rec["w0"][by, bx] + rec["beta"][by, bx] * lx + rec["gamma"][by, bx] * ly
                with np.errstate(divide="ignore", invalid="ignore"):
                    L = -(ac + au * x + av * y) / c0
                    Lx, Ly = -au / c0, -av / c0
                slack = (np.abs(rec["beta"][by, bx]) + np.abs(rec["gamma"][by, bx]) + np.abs(Lx) + np.abs(Ly)) if LOOSE else (np.abs(rec["beta"][by, bx] - Lx) + np.abs(rec["gamma"][by, bx] - Ly))
                pad = 1e-6 * (np.abs(wpl) + np.abs(L)) + 2e-6 * rec["wmax"][by, bx] / np.maximum(np.abs(c0), 1e-30)
                if KERNEL:  # the margins of ssao_cull_kernel (frame_kernels.cu)
                    kNu = 1.7e-5
                    aInv = 1.0 / np.maximum(np.abs(c0), 1e-30)
                    posL1, camL1 = np.abs(P).sum(-1), np.abs(cam).sum()
                    kappa = kNu * (2 * camL1 + posL1) * aInv
                    Labs = np.maximum(np.abs(-(ac + au * xx + av * yy) / c0), np.abs(L11)) + np.abs(Lx) + np.abs(Ly)
                    rayc = 1.01 * (np.abs(Lx) + np.abs(Ly)) + 1.0102 * (kappa * Labs + kNu * DMAX1 * aInv) + 9.6e-7 * DMAG * aInv + 4.8e-7 * Labs
                    rayc = np.where(kappa <= 0.01, rayc, np.inf)
                    rraw = 0.5 * (rec["rhi"][by, bx] - rec["rlo"][by, bx])
                    grec = np.abs(rec["beta"][by, bx]) + np.abs(rec["gamma"][by, bx])
                    mid = wpl + 0.5 * (rec["rhi"][by, bx] + rec["rlo"][by, bx]) - L
                    if ETA:  # the footprint's texels enter with weights lambda_k t_k: sharper rule
                        wbar = wpl + 0.5 * (rec["rhi"][by, bx] + rec["rlo"][by, bx])
                        dw = grec + rraw
                        eta = dw / np.maximum(wbar - dw, 1e-30) + 2e-3
                        eta = np.where((wbar - dw > 0) & (eta < 0.05), eta, np.inf)
                        border = (bx == 0) | (by == 0) | (bx == rec["kind"].shape[1] - 1) | (by == rec["kind"].shape[0] - 1)
                        eta = np.where(border, np.inf, eta)
                        gray = np.abs(Lx) + np.abs(Ly)
                        rr = (rraw * (1 + eta) + eta * (grec + gray)) / (1 - eta) + 1e-6 * rec["wmax"][by, bx]
                        rayc = rayc - 1.01 * gray
                    else:
                        rr = rraw + 1.01 * grec + 1e-6 * rec["wmax"][by, bx]
                    lo, hi = mid - rr - rayc, mid + rr + rayc
                else:
                    lo = wpl + rec["rlo"][by, bx] - L - slack - pad
                    hi = wpl + rec["rhi"][by, bx] - L + slack + pad
                sgn = np.where(lo > 0, 1, np.where(hi < 0, -1, 0)) * np.sign(c0).astype(int)
                sgn = np.where(k == 1, sgn, 0)
                sky_s = np.where(np.abs(sky_proj) > 1e-5, np.sign(sky_proj), 0).astype(int)
                sgn = np.where(k == 2, sky_s if SKY else 0, sgn).astype(np.int8)
                sgn = np.where(np.isfinite(lo) & np.isfinite(hi) | (k == 2), sgn, 0).astype(np.int8)
                if i not in UND: UND[i] = np.zeros(4, np.int64)
                und = live & (sgn == 0)
                UND[i] += np.array([int((und & (k == 0)).sum()), int((und & (k == 2)).sum()), int((und & (k == 1)).sum()), int(live.sum())])
                return sgn
            if not MULTI:
                sgn = classify(recs[B], B)
            else:
                with np.errstate(invalid="ignore"):
                    if MULTI == "tap":
                        dist = np.maximum(np.abs(x - xx), np.abs(y - yy))
                    elif MULTI == "ray":
                        dist = np.maximum(np.abs(x11 - xx), np.abs(y11 - yy))
                    else:
                        dist = np.broadcast_to(tile_reach[:, None, None], xx.shape)
                s8, s16, s32 = classify(recs[8], 8), classify(recs[16], 16), classify(recs[32], 32)
                sgn = np.where(dist <= 116, s8, np.where(dist <= 236, s16, np.where(dist <= 476, s32, 0))).astype(np.int8)
            stats["taps_all"] += int((covered & inside).sum())
            stats["taps_old"] += int(live.sum())
            visited[i] = live
            dec = live & (sgn != 0)
            stats["decided"] += int(dec.sum())
            stats["unsound"] += int((dec & (np.sign(pr) != sgn)).sum())
            # --- new scheme: step i needs the detailed path unless both classes are decided and equal
            need = live & (i > 1) & ~((sgn != 0) & (sgn == cls_prev))
            detail[i] = need
            stats["gathers_new"] += int(need.sum()) + int((need & ~have_prev).sum())
            have_prev = need
            cls_prev = sgn
            cand = live & (pr * prev_proj < 0) & (step <= 2.0) & (i > 1)
            if cand.any():
                with np.errstate(invalid="ignore", divide="ignore"):
                    cn = T.normalize(T.bilinear_clamp(nrm_img, uv[..., 0], uv[..., 1]))
                hit = cand & (np.einsum("...i,...i->...", cn, rd) < 0)
                stats["hits"] += int(hit.sum())
                live &= ~hit
            prev_pos = np.where(live[..., None], cur, prev_pos)
            prev_proj = np.where(live, pr, prev_proj)
        if PRE:
            flipped = (covered & ~live)  # rays that scored (a flip happened) ...
            for PB in PRE_BLOCKS:
                PREBAD[PB] = PREBAD.get(PB, 0) + int((PREOK[PB] & flipped).sum())
        stats["rays"] += int(covered.sum())
        stats["rays_all_decided"] += int((covered & ~detail.any(0)).sum())
        stats["detail_steps"] += int(detail.sum())
        # warp-level iteration counts: a warp is a 16 x 2 strip of the tile
        vw = visited.reshape(12, ntiles, 8, 32)
        dw = detail.reshape(12, ntiles, 8, 32)
        stats["warp_iters_old"] += int(vw.any(-1).sum())                 # old: one iteration per step any lane visits
        # two-phase: coarse loop runs to the longest lane's n (no breaks), detail loop max over lanes of their detail steps
        stats["warp_iters_coarse"] += int(vw.any(-1).sum())
        stats["warp_iters_detail"] += int(dw.sum(0).max(-1).sum())
    s = stats
    print("kind %s block %d: rays %d, taps visited by the march %d (%.1f / covered px)" % (kind, B, s["rays"], s["taps_old"], s["taps_old"] / max(1, s["rays"] / 24)))
    print("  coarse-decided taps: %.1f %%   unsound: %d" % (100.0 * s["decided"] / s["taps_old"], s["unsound"]))
    print("  rays with no detailed step: %.1f %%   hits: %.1f %% of rays" % (100.0 * s["rays_all_decided"] / s["rays"], 100.0 * s["hits"] / s["rays"]))
    print("  gathers: old %d -> new %d (%.1f %%), detailed steps %d" % (s["taps_old"], s["gathers_new"], 100.0 * s["gathers_new"] / s["taps_old"], s["detail_steps"]))
    for PB in PRESTAT: print("   pre-test, %3d-texel blocks%s: %.1f %% of the rays pass; passed rays that scored (unsound): %d" % (PB, " (overlapping)" if PRE_OVERLAP else "", 100.0 * PRESTAT[PB][0] / max(1, PRESTAT[PB][1]), PREBAD.get(PB, 0)))
    for i in sorted(UND): print("   tap %2d: undecided mixed %6d sky %6d surface %6d of %d" % (i, *UND[i]))
    print("  warp iterations: old march %d; two-phase coarse %d + detail %d" % (s["warp_iters_old"], s["warp_iters_coarse"], s["warp_iters_detail"]))
    return s


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "scene"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    nt = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    study(kind, B, nt)
