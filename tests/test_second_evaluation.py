"""Independent second evaluations of the per-frame stages, written in numpy float64 straight from the reference's GLSL
(not from oracle/): the reference ships no golden output for these stages, so two implementations
written separately from the same published text agreeing is the strongest check available here for the restatement
the GPU path is held to. CPU only."""
import numpy as np
import pytest

from helpers import FrameData
from oracle import oracle as O


def _mat(field):
    return np.array(list(field), np.float64).reshape(4, 4).T


def _f16(a_u16):
    return a_u16.view(np.float16).astype(np.float64)


def bilinear_clamp(img, u, v):
    """Vulkan linear filter with CLAMP_TO_EDGE on a (h, w, c) float64 image; u, v arrays of any one shape."""
    h, w = img.shape[:2]
    x, y = u * w - 0.5, v * h - 0.5
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = (x - x0)[..., None], (y - y0)[..., None]
    xi0, xi1 = np.clip(x0, 0, w - 1).astype(int), np.clip(x0 + 1, 0, w - 1).astype(int)
    yi0, yi1 = np.clip(y0, 0, h - 1).astype(int), np.clip(y0 + 1, 0, h - 1).astype(int)
    top = img[yi0, xi0] * (1 - fx) + img[yi0, xi1] * fx
    bot = img[yi1, xi0] * (1 - fx) + img[yi1, xi1] * fx
    return top * (1 - fy) + bot * fy


def shader_rng(seed_x, seed_y, n):
    """Shaders/SSAO.glsl:4-10 for arrays of seeds: returns (n, ...) float64 draws; uint32 wrap-around arithmetic, the
    float(n) conversion rounded to fp32 as the shader's is, 1/float(0xffffffff) == 2^-32 in fp32."""
    sx, sy = seed_x.astype(np.uint32), seed_y.astype(np.uint32)
    out = []
    m = np.uint32(1103515245)
    with np.errstate(over="ignore"):
        for _ in range(n):
            sx, sy = sx + np.uint32(1), sy + np.uint32(1)
            qx = m * ((sx >> np.uint32(1)) ^ sy)
            qy = m * ((sy >> np.uint32(1)) ^ sx)
            r = m * (qx ^ (qy >> np.uint32(3)))
            out.append(r.astype(np.float32).astype(np.float64) * 2.0 ** -32)
    return np.stack(out)


def normalize(a):
    return a / np.linalg.norm(a, axis=-1, keepdims=True)


def ssao_counts64(fd: FrameData):
    """Shaders/SSAO.glsl:31-84 with seed = uvec2(gl_FragCoord.xy) (DeferredPass.frag:42), all pixels at once."""
    W, H = fd.W, fd.H
    PV = _mat(fd.uniforms.projection) @ _mat(fd.uniforms.view)
    pos_img = fd.position[..., :3].astype(np.float64)
    nrm_img = _f16(fd.normal)[..., :3]
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    uv0 = np.stack([(xx + 0.5) / W, (yy + 0.5) / H], -1)                     # (H, W, 2)
    P = pos_img
    covered = fd.position[..., 3] != 0
    N = normalize(np.where(covered[..., None], nrm_img, [0.0, 0.0, 1.0]))        # empty pixels are not shaded
    # coordinateSystem / LocalToWorld
    wide = np.abs(N[..., 0]) > np.abs(N[..., 1])
    with np.errstate(invalid="ignore", divide="ignore"):
        t_a = np.stack([-N[..., 2], 0 * N[..., 0], N[..., 0]], -1) / np.sqrt(N[..., 0] ** 2 + N[..., 2] ** 2)[..., None]
        t_b = np.stack([0 * N[..., 0], N[..., 2], -N[..., 1]], -1) / np.sqrt(N[..., 1] ** 2 + N[..., 2] ** 2)[..., None]
    T = np.where(wide[..., None], t_a, t_b)
    B = np.cross(N, T)
    xi = shader_rng(xx, yy, 72).reshape(24, 3, H, W)
    ao = np.zeros((H, W), np.int32)
    for ray in range(24):
        loc = normalize(np.stack([2 * xi[ray, 0] - 1, 2 * xi[ray, 1] - 1, xi[ray, 2]], -1))
        rd = T * loc[..., 0:1] + B * loc[..., 1:2] + N * loc[..., 2:3]
        end = P + 0.5 * rd
        pe = np.einsum("ij,hwj->hwi", PV, np.concatenate([end, np.ones((H, W, 1))], -1))
        uv1 = 0.5 * pe[..., :2] / pe[..., 3:4] + 0.5
        perp = normalize(np.cross(np.cross(rd, N), rd))
        prev_pos, prev_proj = P.copy(), np.zeros((H, W))
        live = np.ones((H, W), bool)
        for i in range(12):
            t = i / 12.0
            uv = uv0 * (1 - t) + uv1 * t
            inside = (uv[..., 0] >= 0) & (uv[..., 0] <= 1) & (uv[..., 1] >= 0) & (uv[..., 1] <= 1)
            live &= inside
            # the first tap is the pixel's own centre: a texture unit's fixed-point weights are exactly (1, 0) there
            cur = P if i == 0 else bilinear_clamp(pos_img, uv[..., 0], uv[..., 1])
            proj = np.einsum("hwi,hwi->hw", cur - P, perp)
            step = np.linalg.norm(cur - prev_pos, axis=-1)
            cand = live & (proj * prev_proj < 0) & (step <= 2.0) & (i > 0)
            if cand.any():
                with np.errstate(invalid="ignore", divide="ignore"):
                    cn = normalize(bilinear_clamp(nrm_img, uv[..., 0], uv[..., 1]))
                hit = cand & (np.einsum("hwi,hwi->hw", cn, rd) < 0)
                ao += hit
                live &= ~hit
            prev_pos = np.where(live[..., None], cur, prev_pos)
            prev_proj = np.where(live, proj, prev_proj)
    return ao


@pytest.mark.parametrize("kind,W,H", [("rand", 48, 32), ("scene", 64, 40)])
def test_ssao_counts_against_a_float64_evaluation(kind, W, H):
    fd = FrameData(kind, W, H, n_lights=0)
    got = O.ssao(fd.oracle_frame()).astype(np.int32)
    want = ssao_counts64(fd)
    covered = fd.position[..., 3] != 0
    assert covered.sum() > 0.5 * W * H
    assert np.all(got[~covered] == 255)
    diff = np.abs(got - want)[covered]
    assert want[covered].max() > 3                       # the case exercises occlusion at all
    # fp32 and float64 disagree only where a product of projections sits at the sign boundary
    # (observed: none at these sizes)
    assert (diff == 0).mean() >= 0.999, (diff == 0).mean()
    assert diff.max() <= 1, diff.max()


def reconstruct_position64(g, u, v, d_raw):
    """Shaders/Misc/ReconstructPosition.glsl:4-22 for arrays."""
    invP, invV = _mat(g.inverseProjection), _mat(g.inverseView)
    near, far = 0.01, 1000.0
    d = far * near / (d_raw * (far - near) - far)
    plane = np.stack([2 * u - 1, 2 * v - 1, np.full_like(u, 2.0), np.ones_like(u)], -1)
    dh = plane @ invP.T
    dirv = normalize((dh[..., :3] / dh[..., 3:4]) @ invV[:3, :3].T)
    f = dirv @ invV[:3, 2]
    return invV[:3, 3] + d[..., None] * dirv / f[..., None]


def ssr_march64(fd: FrameData):
    """Shaders/SSR.frag:80-149 with SSR.vert:15-23: per pixel (hit, steps taken, hit uv, hit position, hit normal)."""
    W, H = fd.W, fd.H
    g = fd.uniforms
    PV = _mat(g.projection) @ _mat(g.view)
    invP, invV = _mat(g.inverseProjection), _mat(g.inverseView)
    depth_img = fd.depth.astype(np.float64)[..., None]
    nrm_img = _f16(fd.normal)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    u, v = (xx + 0.5) / W, (yy + 0.5) / H
    valid = nrm_img[..., 3] != 0
    P = reconstruct_position64(g, u, v, depth_img[..., 0])
    N = normalize(np.where(valid[..., None], nrm_img[..., :3], [0.0, 0.0, 1.0]))
    ndc = np.stack([2 * u - 1, 2 * v - 1, np.zeros_like(u), np.ones_like(u)], -1)
    view_dir = normalize((ndc @ invP.T)[..., :3] @ invV[:3, :3].T)
    rd = view_dir - 2 * np.einsum("hwi,hwi->hw", N, view_dir)[..., None] * N            # reflect()
    end = P + rd * 10000.0
    pe = np.concatenate([end, np.ones((H, W, 1))], -1) @ PV.T
    uv_end = 0.5 * pe[..., :2] / pe[..., 3:4] + 0.5
    uv = np.stack([u, v], -1)
    step = normalize(uv_end - uv) * 0.005
    perp = normalize(np.cross(np.cross(rd, N), rd))
    prev_proj = np.zeros((H, W))
    live = valid.copy()
    hit = np.zeros((H, W), bool)
    steps = np.zeros((H, W), np.int32)
    hit_uv, hit_pos, hit_nrm = np.zeros((H, W, 2)), np.zeros((H, W, 3)), np.zeros((H, W, 3))
    for i in range(128):
        steps[live] = i + 1
        uv = uv + step
        live &= (uv[..., 0] >= 0) & (uv[..., 0] <= 1) & (uv[..., 1] >= 0) & (uv[..., 1] <= 1)
        d_raw = bilinear_clamp(depth_img, uv[..., 0], uv[..., 1])[..., 0]
        cur = reconstruct_position64(g, uv[..., 0], uv[..., 1], d_raw)
        dirv = normalize(cur - P)
        proj = np.einsum("hwi,hwi->hw", dirv, perp)
        f = np.einsum("hwi,hwi->hw", dirv, rd)
        cand = live & (proj * prev_proj <= 0) & (f > 0.999) & (i > 0)
        if cand.any():
            with np.errstate(invalid="ignore", divide="ignore"):
                cn = normalize(bilinear_clamp(nrm_img[..., :3], uv[..., 0], uv[..., 1]))
            h = cand & (np.einsum("hwi,hwi->hw", cn, rd) < 0)
            hit |= h
            hit_uv[h], hit_pos[h], hit_nrm[h] = uv[h], cur[h], cn[h]
            live &= ~h
        prev_proj = np.where(live, proj, prev_proj)
    return valid, hit, steps, hit_uv, hit_pos, hit_nrm, rd


@pytest.mark.parametrize("kind,W,H", [("scene", 96, 60), ("rand", 48, 32)])
def test_ssr_march_against_a_float64_evaluation(kind, W, H):
    fd = FrameData(kind, W, H, n_lights=0)
    _, got_hit, got_steps = O.ssr_capture(fd.oracle_frame())
    valid, hit, steps, *_ = ssr_march64(fd)
    got_hit = got_hit.astype(bool)
    assert not got_hit[~valid].any() and np.all(got_steps[~valid] == 0)
    if kind == "scene":
        assert hit.mean() > 0.02                         # the spheres-on-a-plane scene does produce reflections
    agree = (got_hit == hit)[valid]
    same_steps = (got_steps == steps)[valid]
    print(kind, "hit agreement", agree.mean(), "step agreement", same_steps.mean(), "hit fraction", hit.mean())
    # the tests f > 0.999 and proj * prevProj <= 0 sit on fp32 noise for a few pixels
    assert agree.mean() > 0.99 and same_steps.mean() > 0.99


def equirect_uv(d):
    """Shaders/PBR/PBRMaterial.glsl:5-7."""
    yaw = np.arctan2(d[..., 2], d[..., 0])
    pitch = -np.arctan2(d[..., 1], np.hypot(d[..., 0], d[..., 2]))
    return 0.5 * yaw / np.pi + 0.5, pitch / np.pi + 0.5


def trilinear_clamp(flat_chain, w, h, mips, u, v, lod):
    """LINEAR mipmap mode over a tight RGBA32F chain: lod clamped to [0, mips - 1], the two nearest levels blended."""
    levels, off = [], 0
    for l in range(mips):
        lw, lh = max(w >> l, 1), max(h >> l, 1)
        levels.append(flat_chain[off:off + lw * lh * 4].reshape(lh, lw, 4).astype(np.float64))
        off += lw * lh * 4
    lod = np.clip(lod, 0, mips - 1)
    l0 = np.minimum(np.floor(lod).astype(int), mips - 1)
    l1 = np.minimum(l0 + 1, mips - 1)
    fr = (lod - l0)[..., None]
    out = np.zeros(u.shape + (4,))
    for l in range(mips):
        s = bilinear_clamp(levels[l], u, v)
        out += np.where((l0 == l)[..., None], s * (1 - fr), 0) + np.where((l1 == l)[..., None] & (l1 != l0)[..., None], s * fr, 0)
    return out


def environment_term64(fd, V, N, base, metallic, rough, refl, occl):
    """Shaders/PBR/PBRMaterial.glsl:82-114 (no point lights) for arrays of pixels."""
    NdotV = np.maximum(np.einsum("...i,...i->...", N, -V), 0.0)
    F0 = 0.04 * (1 - metallic[..., None]) + base * metallic[..., None]
    F = F0 + (np.maximum(1 - rough[..., None], F0) - F0) * ((1 - NdotV) ** 5)[..., None]
    diffuse = (1 - F) * base * (1 - metallic[..., None])
    lut = bilinear_clamp(fd.lut.astype(np.float64) / 255.0, NdotV, rough)
    iu, iv = equirect_uv(N)
    irr = bilinear_clamp(fd.irr.astype(np.float64), iu, iv)[..., :3]
    return (irr * diffuse + refl * (F * lut[..., 0:1] + lut[..., 1:2])) * occl[..., None]


def test_ssr_hit_colour_against_a_float64_evaluation():
    """Shaders/SSR.frag:55-78: the colour written for a hit, IBL terms only (no lights), then the blend-on-write
    (Src/GraphicsPipeline.cpp:138-154: rgb * a, a) and the RGBA16F store."""
    W, H = 96, 60
    fd = FrameData("scene", W, H, n_lights=0)
    refl, got_hit, got_steps = O.ssr_capture(fd.oracle_frame())
    valid, hit, steps, huv, hpos, hn, rd = ssr_march64(fd)
    both = hit & got_hit.astype(bool) & (steps == got_steps)          # the same hit, not a neighbouring step's
    assert both.sum() > 300
    u, v = huv[both, 0], huv[both, 1]
    base = bilinear_clamp(fd.albedo.astype(np.float64) / 255.0, u, v)[..., :3]
    mro = bilinear_clamp(fd.mro.astype(np.float64) / 255.0, u, v)
    Vd, N = normalize(rd[both]), hn[both]
    rdir = Vd - 2 * np.einsum("ni,ni->n", N, Vd)[:, None] * N
    eu, ev = equirect_uv(rdir)
    pre = trilinear_clamp(fd.pre, fd.pre_size[0], fd.pre_size[1], 5, eu, ev, 4.0 * mro[:, 1])[..., :3]
    want = environment_term64(fd, Vd, N, base, mro[:, 0], mro[:, 1], pre, np.ones(len(u)))
    got = refl.view(np.float16).astype(np.float64)[both]
    assert np.all(got[:, 3] == 1.0)
    err = np.abs(got[:, :3] - want) / (np.abs(want) + 1e-2)
    # half storage: 2^-11 relative; the hit position differs in the last fp32 bits, which moves the lookups a little
    assert np.quantile(err, 0.99) < 2e-3 and err.max() < 2e-2, (np.quantile(err, 0.99), err.max())
    # and no-hit pixels hold (0, 0, 0, 0)
    assert not refl[~got_hit.astype(bool)].any()


def shadow_cube_lookup64(cubes, res, light, q):
    """samplerCubeArray lookup of direction q (n, 3) in layer 6 * light + face: major-axis face selection and (s, t) per
    the Vulkan cube map table, linear filter kept inside the selected face."""
    ax = np.abs(q)
    face = np.where((ax[:, 0] >= ax[:, 1]) & (ax[:, 0] >= ax[:, 2]), np.where(q[:, 0] >= 0, 0, 1),
                    np.where(ax[:, 1] >= ax[:, 2], np.where(q[:, 1] >= 0, 2, 3), np.where(q[:, 2] >= 0, 4, 5)))
    x, y, z = q[:, 0], q[:, 1], q[:, 2]
    sc = np.choose(face, [-z, z, x, x, x, -x])
    tc = np.choose(face, [-y, -y, z, -z, -y, -y])
    ma = np.choose(face, [ax[:, 0], ax[:, 0], ax[:, 1], ax[:, 1], ax[:, 2], ax[:, 2]])
    s, t = 0.5 * sc / ma + 0.5, 0.5 * tc / ma + 0.5
    out = np.zeros(len(q))
    cubes = np.asarray(cubes).reshape(-1, res, res)
    for f in range(6):
        m = face == f
        if m.any():
            layer = cubes[6 * light + f].astype(np.float64)[..., None]
            out[m] = bilinear_clamp(layer, s[m], t[m])[..., 0]
    return out


def test_deferred_frame_against_a_float64_evaluation():
    """Shaders/DeferredPass.frag:41-93 + PBR/PBRMaterial.glsl:72-162 on the spheres-on-a-plane scene with the real maps,
    four shadowed point lights, the reflection chain and the tone map; inputs of the stage (reflection chain, AO counts)
    come from the stages before it."""
    W, H = 96, 60
    fd = FrameData("scene", W, H, n_lights=4, shadow_res=32)
    fr = fd.oracle_frame()
    refl, _, _ = O.ssr_capture(fr)
    chain = O.glossy_convolve(refl, 5)
    ao = O.ssao(fr)
    got = O.deferred_shade(fr, chain, 5, 0, ao).astype(np.float64)

    g = fd.uniforms
    invP, invV = _mat(g.inverseProjection), _mat(g.inverseView)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    u, v = (xx + 0.5) / W, (yy + 0.5) / H
    ndc = np.stack([2 * u - 1, 2 * v - 1, np.zeros_like(u), np.ones_like(u)], -1)
    direction = (ndc @ invP.T)[..., :3] @ invV[:3, :3].T
    covered = fd.position[..., 3] != 0
    want = np.zeros((H, W, 3))
    eu, ev = equirect_uv(direction)
    want[~covered] = bilinear_clamp(fd.env.astype(np.float64), eu, ev)[~covered][:, :3]

    c = covered
    Vd = normalize(direction)[c]
    N = normalize(_f16(fd.normal)[..., :3][c])
    base = fd.albedo[c][:, :3] / 255.0
    metallic, rough = fd.mro[c][:, 0] / 255.0, fd.mro[c][:, 1] / 255.0
    occl = 1.0 - ao[c] / 24.0
    P = fd.position[c][:, :3].astype(np.float64)
    rdir = Vd - 2 * np.einsum("ni,ni->n", N, Vd)[:, None] * N
    ru, rv = equirect_uv(rdir)
    env_refl = trilinear_clamp(fd.pre, fd.pre_size[0], fd.pre_size[1], 5, ru, rv, 4.0 * rough)[:, :3]
    chain32 = chain.view(np.float16).astype(np.float32)
    rc = trilinear_clamp(chain32, W, H, 5, u[c], v[c], 4.0 * rough)
    a = rc[:, 3:4]
    with np.errstate(invalid="ignore", divide="ignore"):
        mixed = env_refl * (1 - a) + (rc[:, :3] / a) * a
    refl_c = np.where(a < 0.01, env_refl, mixed)
    col = environment_term64(fd, Vd, N, base, metallic, rough, refl_c, occl)
    NdotV = np.maximum(np.einsum("ni,ni->n", N, -Vd), 0.0)
    F0 = 0.04 * (1 - metallic[:, None]) + base * metallic[:, None]
    al = rough * rough
    a2 = al * al
    k = (al + 1) ** 2 / 8
    g1 = lambda cs: cs / (cs * (1 - k) + k)  # noqa: E731
    near_threshold = np.zeros(len(P), bool)
    for i, li in enumerate(fd.lights.astype(np.float64)):
        L = li[0:3] - P
        d2 = np.einsum("ni,ni->n", L, L)
        dist = np.sqrt(d2)
        L = L / dist[:, None]
        closest = shadow_cube_lookup64(fd.shadow, fd.shadow_res, i, L * [1.0, -1.0, -1.0]) * 1000.0
        lit = ~(closest < dist - 0.5)
        near_threshold |= np.abs(closest - (dist - 0.5)) < 1e-3
        Hh = normalize(Vd + L)
        NdotL = np.maximum(np.einsum("ni,ni->n", N, L), 0.0)
        NdotH = np.maximum(np.einsum("ni,ni->n", N, Hh), 0.0)
        F = F0 + (np.maximum(1 - rough[:, None], F0) - F0) * ((1 - NdotH) ** 5)[:, None]
        diff = (1 - F) * base * (1 - metallic[:, None]) / np.pi
        t = NdotH * NdotH * (a2 - 1) + 1
        spec = (a2 / (np.pi * t * t) * g1(NdotL) * g1(NdotV) / (4 * NdotL * NdotV + 0.0001))[:, None] * F
        col = col + np.where(lit[:, None], (diff + spec) * (li[4:7] / d2[:, None]) * NdotL[:, None], 0.0)
    want[c] = col
    want = 1.0 - np.exp(-want * g.exposure)

    err = np.abs(got[..., :3] - want)
    ok = np.ones((H, W), bool)
    ok[c] = ~near_threshold
    assert ok.mean() > 0.99
    assert np.all(got[..., 3] == 1.0)
    # tone-mapped values live in [0, 1): absolute error. fp32 atan2 / pow against float64 leaves ~1e-6.
    assert err[ok].max() < 2e-5, err[ok].max()
    assert (fd.shadow.min() * 1000 < 50) and covered.mean() > 0.5       # shadows and geometry are really in play


# ---- IBL precompute (the pinned half as well: a tight check next to the RGBE-bounded pin of test_oracle_pin.py) --------------
def _chain_levels(flat_chain, w, h, mips):
    levels, off = [], 0
    for l in range(mips):
        lw, lh = max(w >> l, 1), max(h >> l, 1)
        levels.append(flat_chain[off:off + lw * lh * 4].reshape(lh, lw, 4).astype(np.float64))
        off += lw * lh * 4
    return levels


def _bilinear_repeat(img, u, v):
    h, w = img.shape[:2]
    x, y = u * w - 0.5, v * h - 0.5
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = (x - x0)[..., None], (y - y0)[..., None]
    xi0, xi1 = x0.astype(int) % w, (x0.astype(int) + 1) % w
    yi0, yi1 = y0.astype(int) % h, (y0.astype(int) + 1) % h
    top = img[yi0, xi0] * (1 - fx) + img[yi0, xi1] * fx
    bot = img[yi1, xi0] * (1 - fx) + img[yi1, xi1] * fx
    return top * (1 - fy) + bot * fy


def _env_lookup64(levels, d, mip):
    """sampleEnvMap of the precompute shaders (GenIrradianceMap.comp:78-102): REPEAT sampler, LINEAR mip filter."""
    len_xz = np.hypot(d[..., 0], d[..., 2])
    safe = len_xz > 0.001
    yaw = np.where(safe, np.arctan2(d[..., 2], d[..., 0]), 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        pitch = np.where(safe, np.arctan(d[..., 1] / len_xz), np.where(d[..., 1] > 0, 0.5 * np.pi, -0.5 * np.pi))
    u, v = yaw / (2 * np.pi) + 0.5, pitch / np.pi + 0.5
    mip = np.clip(np.broadcast_to(mip, u.shape), 0, len(levels) - 1)
    l0 = np.minimum(np.floor(mip).astype(int), len(levels) - 1)
    l1 = np.minimum(l0 + 1, len(levels) - 1)
    fr = (mip - l0)[..., None]
    out = np.zeros(u.shape + (4,))
    for l in range(len(levels)):
        m0, m1 = l0 == l, (l1 == l) & (l1 != l0)
        if m0.any() or m1.any():
            s = _bilinear_repeat(levels[l], u, v)
            out += np.where(m0[..., None], s * (1 - fr), 0) + np.where(m1[..., None], s * fr, 0)
    return out[..., :3]


def _local_to_world(n):
    if abs(n[0]) > abs(n[1]):
        t = np.array([-n[2], 0.0, n[0]]) / np.sqrt(n[0] ** 2 + n[2] ** 2)
    else:
        t = np.array([0.0, n[2], -n[1]]) / np.sqrt(n[1] ** 2 + n[2] ** 2)
    return t, np.cross(n, t), n


def _texel_normal(x, y, w, h):
    yaw, pitch = np.pi * (2.0 * x / w - 1.0), np.pi * (y / h - 0.5)
    return np.array([np.cos(pitch) * np.cos(yaw), np.sin(pitch), np.cos(pitch) * np.sin(yaw)])


def test_irradiance_against_a_float64_evaluation():
    """Shaders/IBL_Precompute/GenIrradianceMap.comp:106-155 at a few texels of a 512 x 256 map, all 300 x 150 samples."""
    from helpers import golden_env
    env = golden_env()
    H, W = env.shape[:2]
    chain, mips = O.env_mip_chain(env)
    levels = _chain_levels(chain, W, H, mips)
    texels = [(0, 128, 0), (100, 40, 0), (511, 255, 0), (256, 1, 0), (333, 200, 0)]
    got = O.ibl_irradiance(chain, W, H, mips, W, H, texels)
    theta_n, phi_n = 300, int(H * 300 / W)
    mip = np.log2(W / 300.0)
    th = np.arange(theta_n) * 2 * np.pi / theta_n
    ph = np.arange(phi_n) * 0.5 * np.pi / phi_n
    T, P = np.meshgrid(th, ph, indexing="ij")
    local = np.stack([np.cos(T) * np.sin(P), np.sin(T) * np.sin(P), np.cos(P)], -1)
    for (x, y, _), g in zip(texels, got):
        t, b, n = _local_to_world(_texel_normal(x, y, W, H))
        d = local[..., 0:1] * t + local[..., 1:2] * b + local[..., 2:3] * n
        s = _env_lookup64(levels, d, mip) * (np.cos(P) * np.sin(P))[..., None]
        want = np.pi * s.sum(axis=(0, 1)) / theta_n / phi_n
        assert np.abs(g[:3] - want).max() < 5e-5 * np.abs(want).max(), (x, y, g[:3], want)   # observed 4e-6
        assert g[3] == 1.0


@pytest.mark.parametrize("roughness", [0.0, 0.25, 0.75, 1.0])
def test_prefilter_against_a_float64_evaluation(roughness):
    """Shaders/IBL_Precompute/PreFilterEnvMap.comp:126-177 with the shader's own hash sequence (:38-44) and GGX sampling
    (:84-92), 10000 samples per texel. The sample directions depend only on integer hashing, so the two evaluations see the
    same samples; what differs is fp32 against float64 in the warps, the mip choice and the sum."""
    from helpers import golden_env
    env = golden_env()
    H, W = env.shape[:2]
    chain, mips = O.env_mip_chain(env)
    levels = _chain_levels(chain, W, H, mips)
    ow, oh = W >> 1, H >> 1
    texels = [(17, 64, 0), (200, 100, 0), (128, 5, 0)]
    got = O.ibl_prefilter(chain, W, H, mips, ow, oh, roughness, texels)
    n_samples = 10000
    a2 = roughness * roughness
    for (x, y, _), g in zip(texels, got):
        xi = shader_rng(np.array([x]), np.array([y]), 2 * n_samples)[:, 0].reshape(n_samples, 2)
        phi = 2 * np.pi * xi[:, 0]
        cos_t = np.sqrt((1 - xi[:, 1]) / (1 + (a2 - 1) * xi[:, 1]))
        sin_t = np.sqrt(1 - cos_t * cos_t)
        t, b, n = _local_to_world(_texel_normal(x, y, ow, oh))
        Hh = (np.cos(phi) * sin_t)[:, None] * t + (np.sin(phi) * sin_t)[:, None] * b + cos_t[:, None] * n
        V = n
        L = normalize(2.0 * (Hh @ V)[:, None] * Hh - V)
        NdotL, NdotH, HdotV = np.maximum(L @ n, 0), np.maximum(Hh @ n, 0), np.maximum(Hh @ V, 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            D = a2 / (np.pi * (NdotH * NdotH * (a2 - 1) + 1) ** 2)   # 0 / 0 at roughness 0, where the mip is fixed anyway
        pdf = D * NdotH / (4 * HdotV + 0.00001)
        sa_texel = 4 * np.pi / (6.0 * W * H)
        sa_sample = 1.0 / (n_samples * pdf + 0.0001)
        with np.errstate(divide="ignore"):
            mip = np.zeros(n_samples) if roughness == 0.0 else 0.5 * np.log2(sa_sample / sa_texel)
        use = NdotL > 0
        col = _env_lookup64(levels, L[use], mip[use]) * NdotL[use, None]
        want = col.sum(0) / NdotL[use].sum()
        assert np.abs(g[:3] - want).max() < 5e-5 * np.abs(want).max(), (x, y, roughness, g[:3], want)   # observed 3e-6


def test_ssr_span_skip_proof_is_sound_on_a_small_frame():
    """tools/ssr_skip_study.py (the plan of DESIGN.md section 8 item 2): the conservative proof never declares a span hit-free that
    contains a step with f > 0.999, with window bounds and with pyramid bounds."""
    import importlib
    import os
    import sys

    from helpers import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    study = importlib.import_module("ssr_skip_study")
    try:
        study.PYRAMID = False
        assert study.study("scene", 96, 54) == 0
        study.PYRAMID = True
        assert study.study("scene", 96, 54) == 0
    finally:
        study.PYRAMID = False
