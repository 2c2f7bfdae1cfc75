"""CPU tests of the oracle's building blocks: RNG known answers, storage formats, the texture-unit rules A1-A8,
depth reconstruction. (SURVEY.md App. B / App. D "oracle self-tests".)"""
import numpy as np
import pytest

from althea_b200 import scene


def test_rng_known_answers(oracle):
    # exact uint32 arithmetic of SSAO.glsl:5-11 / PreFilterEnvMap.comp:36-42
    kat = {
        (0, 0): ([0x94727F20, 0x21C1E403, 0x6AAB4CAD, 0xE54A1673], [0.579872072, 0.131864786, 0.416676313, 0.895661712]),
        (1, 2): ([0xE030B4CA, 0x0114F97C, 0x6FC823E2, 0x6227B801], [0.87574321, 0.00422629621, 0.436647654, 0.38341856]),
        (1279, 719): ([0x2D7598A5, 0xF1C5B1CF, 0xF933DE2E, 0x94862436], [0.177575633, 0.944422841, 0.97344768, 0.580171824]),
        (3839, 2159): ([0x3DBA150F, 0x17B7B391, 0xC4898A24, 0x96FC64CC], [0.241120636, 0.0926468149, 0.76772368, 0.589788735]),
    }
    for (sx, sy), (us, fs) in kat.items():
        u, f = oracle.rng(sx, sy, 4)
        assert [int(x) for x in u] == us
        np.testing.assert_allclose(f, np.array(fs, np.float32), rtol=0, atol=1e-7)


def test_rng_can_return_exactly_one(oracle):
    # float(n) * 2^-32 rounds to 1.0 for n >= 2^32 - 128
    u = np.array([0xFFFFFFFF, 0xFFFFFF80, 0xFFFFFF7F], np.uint32)
    f = u.astype(np.float32) * np.float32(2.0 ** -32)
    assert f[0] == 1.0 and f[1] == 1.0 and f[2] < 1.0


def test_half_conversion_matches_ieee(oracle):
    rs = np.random.default_rng(3)
    x = (rs.standard_normal(200000) * 10.0 ** rs.integers(-9, 6, 200000)).astype(np.float32)
    x = np.concatenate([x, np.array([0, -0.0, 65504, 65519.99, 65520, 1e9, -1e9, np.inf, -np.inf, 2 ** -24, 2 ** -25,
                                     2 ** -25 * 1.0001, 2 ** -14, 6.1e-5, 5.96e-8 * 1.5, 5.96e-8 * 2.5], np.float32)])
    h, f = oracle.half_roundtrip(x)
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16)
    assert np.array_equal(h, ref.view(np.uint16))
    assert np.array_equal(f, ref.astype(np.float32))


def test_bilinear_rules(oracle):
    # 4x2 RGBA32F ramp: texel (x, y) = (x, y, x + 10 y, 1)
    w, h = 4, 2
    img = np.zeros((h, w, 4), np.float32)
    for y in range(h):
        for x in range(w):
            img[y, x] = (x, y, x + 10 * y, 1)
    O = oracle
    # A3: pixel centres are exact
    uv = [((x + 0.5) / w, (y + 0.5) / h, 0.0) for y in range(h) for x in range(w)]
    out = O.sample(img, w, h, 1, O.FMT_RGBA32F, O.ADDR_CLAMP, uv)
    assert np.array_equal(out, img.reshape(-1, 4))
    # A1: halfway between texels 1 and 2 on row 0
    out = O.sample(img, w, h, 1, O.FMT_RGBA32F, O.ADDR_CLAMP, [(2.0 / w, 0.5 / h, 0)])
    np.testing.assert_allclose(out[0], [1.5, 0, 1.5, 1])
    # A2 clamp: left of texel 0 stays texel 0; repeat: wraps to the average of texels 3 and 0
    out_c = O.sample(img, w, h, 1, O.FMT_RGBA32F, O.ADDR_CLAMP, [(0.0, 0.25, 0)])
    out_r = O.sample(img, w, h, 1, O.FMT_RGBA32F, O.ADDR_REPEAT, [(0.0, 0.25, 0)])
    np.testing.assert_allclose(out_c[0, 0], 0.0)
    np.testing.assert_allclose(out_r[0, 0], 1.5)
    # A4: corner-aligned compute uv texelPos/size at 2:1 = 4-texel average around the corner (2x-1, 2x)
    out = O.sample(img, w, h, 1, O.FMT_RGBA32F, O.ADDR_REPEAT, [(1 / 2.0, 1 / 1.0 * 0.5, 0)])
    np.testing.assert_allclose(out[0, 0], 1.5)


def test_trilinear_lod_clamp(oracle):
    O = oracle
    w, h = 4, 4
    chain = np.concatenate([np.full(16 * 4, 1.0, np.float32), np.full(4 * 4, 2.0, np.float32), np.full(4, 4.0, np.float32)])
    f = lambda lod: O.sample(chain, w, h, 3, O.FMT_RGBA32F, O.ADDR_CLAMP, [(0.5, 0.5, lod)])[0, 0]  # noqa: E731
    assert f(-3.0) == 1.0 and f(0.0) == 1.0 and f(1.0) == 2.0 and f(2.0) == 4.0 and f(9.0) == 4.0
    np.testing.assert_allclose(f(0.25), 1.25)
    np.testing.assert_allclose(f(1.5), 3.0)


def test_unorm8_and_half_texels(oracle):
    O = oracle
    img8 = np.array([[[0, 128, 255, 51]]], np.uint8)
    out = O.sample(img8, 1, 1, 1, O.FMT_RGBA8, O.ADDR_CLAMP, [(0.5, 0.5, 0)])
    np.testing.assert_allclose(out[0], np.array([0, 128, 255, 51], np.float32) / np.float32(255.0), rtol=0, atol=0)
    img16 = np.array([[[1.5, -2.0, 0.333251953125, 65504]]], np.float16).view(np.uint16)
    out = O.sample(img16, 1, 1, 1, O.FMT_RGBA16F, O.ADDR_CLAMP, [(0.5, 0.5, 0)])
    np.testing.assert_array_equal(out[0], np.array([1.5, -2.0, 0.333251953125, 65504], np.float32))


def test_cube_face_table(oracle):
    # layer value = light*6 + face, constant per face: the lookup must select the Vulkan face (rule A8)
    res = 4
    cubes = np.zeros((2, 6, res, res), np.float32)
    for li in range(2):
        for f in range(6):
            cubes[li, f] = li * 6 + f
    dirs = {0: (1, 0.2, -0.3), 1: (-1, 0.2, 0.3), 2: (0.1, 1, 0.3), 3: (0.1, -1, -0.3), 4: (0.2, -0.1, 1), 5: (0.2, 0.1, -1)}
    for f, q in dirs.items():
        assert oracle.sample_cube(cubes, res, 1, q) == 6 + f
    # (s, t) orientation on +X: s = (-z/|x| + 1)/2, t = (-y/|x| + 1)/2
    ramp = np.zeros((1, 6, res, res), np.float32)
    ramp[0, 0] = np.arange(res)[None, :] + 10 * np.arange(res)[:, None]
    v = oracle.sample_cube(ramp, res, 0, (1.0, -0.25, -0.75))  # s = .875 -> x = 3, t = .625 -> y = 2
    assert v == pytest.approx(3 + 10 * 2)


def test_depth_reconstruction_roundtrip(oracle):
    W, H = 1280, 720
    g = scene.make_uniforms(W, H, pos=(1.0, 2.0, 3.0), yaw=0.4, pitch=-0.2)
    og = oracle.GlobalUniforms.from_buffer_copy(bytes(g))
    P = np.array(list(g.projection), np.float64).reshape(4, 4).T
    V = np.array(list(g.view), np.float64).reshape(4, 4).T
    rs = np.random.default_rng(1)
    for _ in range(50):
        eye = np.array([rs.uniform(-3, 3), rs.uniform(-2, 2), -rs.uniform(1, 60), 1.0])
        world = np.linalg.inv(V) @ eye
        clip = P @ eye
        ndc = clip[:3] / clip[3]
        u, v = 0.5 * ndc[0] + 0.5, 0.5 * ndc[1] + 0.5
        p = oracle.reconstruct_position(og, u, v, ndc[2])
        # fp32 depth near 1.0 has ~6e-8 resolution, which is ~0.03 units at 60 units distance with near=0.01
        np.testing.assert_allclose(p, world[:3], atol=2e-3 * abs(eye[2]) + 1e-3)


def test_deferred_pixels_against_an_independent_float64_evaluation(oracle):
    """Known-answer pin for the shading stage: with constant-colour IBL maps and a constant BRDF LUT every
    texture lookup has a closed-form value, so each pixel's colour follows from the published equations alone
    (Shaders/DeferredPass.vert:10-22, DeferredPass.frag:41-93, PBR/PBRMaterial.glsl:41-162). Evaluated here in float64
    straight from those equations and compared with the oracle's fp32 result."""
    W, H = 12, 10
    rs = np.random.default_rng(7)
    n_lights = 3
    g = scene.make_uniforms(W, H, pos=(0.5, 1.0, 4.0), yaw=0.3, pitch=-0.15, light_count=n_lights)
    og = oracle.GlobalUniforms.from_buffer_copy(bytes(g))
    invP = np.array(list(g.inverseProjection), np.float64).reshape(4, 4).T
    invV = np.array(list(g.inverseView), np.float64).reshape(4, 4).T

    env_c, pre_c, irr_c = np.array([0.3, 0.5, 0.7]), np.array([0.9, 0.4, 0.2]), np.array([0.25, 0.35, 0.15])
    const = lambda c, h, w: np.broadcast_to(np.append(c, 1.0).astype(np.float32), (h, w, 4)).copy()
    env, irr = const(env_c, 8, 16), const(irr_c, 4, 8)
    pre_w, pre_h = 16, 8
    pre = np.broadcast_to(np.append(pre_c, 1.0).astype(np.float32), (oracle.chain_texels(pre_w, pre_h, 5), 4)).copy().ravel()
    lut = np.zeros((4, 4, 4), np.uint8)
    lut[..., 0], lut[..., 1] = 153, 51                      # envBRDF = (0.6, 0.2) everywhere
    lights = np.zeros((n_lights, 8), np.float32)
    lights[:, 0:3] = rs.uniform(-3, 3, (n_lights, 3))
    lights[:, 4:7] = rs.uniform(2, 20, (n_lights, 3))

    position = np.zeros((H, W, 4), np.float32)
    position[..., :3] = rs.uniform(-2, 2, (H, W, 3))
    position[..., 3] = 1.0
    position[0, 0, 3] = position[H - 1, W - 1, 3] = 0.0     # two empty pixels: the sky branch
    nrm = rs.normal(size=(H, W, 3))
    nrm *= rs.uniform(0.5, 2.0, (H, W, 1)) / np.linalg.norm(nrm, axis=-1, keepdims=True)  # the shader normalises
    normal16 = np.zeros((H, W, 4), np.float16)
    normal16[..., :3] = nrm
    albedo = rs.integers(0, 256, (H, W, 4), dtype=np.uint8)
    mro = rs.integers(0, 256, (H, W, 4), dtype=np.uint8)
    ao = rs.integers(0, 25, (H, W), dtype=np.uint8)
    depth = np.full((H, W), 0.5, np.float32)
    chain = np.zeros(oracle.chain_texels(W, H, 5) * 4, np.uint16)  # reflection alpha 0 => prefiltered map is used

    # shadow cubes with one value per light: light 1 stores distance 0 (everything behind it is shadowed), the others
    # store 1.0 = 1000 units (nothing is)
    cubes = np.ones((n_lights * 6, 4, 4), np.float32)
    cubes[6:12] = 0.0
    fr = oracle.Frame(og, W, H, position, depth, normal16.view(np.uint16), albedo, mro, env, pre, (pre_w, pre_h), 5, irr,
                      lut, lights, cubes, 4)
    got = oracle.deferred_shade(fr, chain, 5, oracle.SKIP_TONEMAP, ao)

    def fresnel(c, F0, rough):
        return F0 + (np.maximum(1.0 - rough, F0) - F0) * (1.0 - c) ** 5

    def g1(c, k):
        return c / (c * (1.0 - k) + k)

    worst = 0.0
    for y in range(H):
        for x in range(W):
            u, v = (x + 0.5) / W, (y + 0.5) / H
            if position[y, x, 3] == 0.0:
                want = env_c
            else:
                d = invV[:3, :3] @ (invP @ np.array([2 * u - 1, 2 * v - 1, 0.0, 1.0]))[:3]
                V = d / np.linalg.norm(d)
                N = normal16[y, x, :3].astype(np.float64)
                N /= np.linalg.norm(N)
                base = albedo[y, x, :3] / 255.0
                metallic, rough = mro[y, x, 0] / 255.0, mro[y, x, 1] / 255.0
                occl = 1.0 - ao[y, x] / 24.0
                NdotV = max(N @ -V, 0.0)
                F0 = 0.04 * (1 - metallic) + base * metallic
                a = rough * rough
                a2 = a * a
                k = (a + 1.0) ** 2 / 8.0
                F = fresnel(NdotV, F0, rough)
                want = (irr_c * (1 - F) * base * (1 - metallic) + pre_c * (F * 0.6 + 0.2)) * occl
                P = position[y, x, :3].astype(np.float64)
                for i, li in enumerate(lights.astype(np.float64)):
                    L = li[0:3] - P
                    d2 = L @ L
                    if i == 1 and 0.0 < np.sqrt(d2) - 0.5:
                        continue
                    L = L / np.sqrt(d2)
                    Hh = (V + L) / np.linalg.norm(V + L)    # as published: V points from the eye to the surface
                    NdotL, NdotH = max(N @ L, 0.0), max(N @ Hh, 0.0)
                    Fl = fresnel(NdotH, F0, rough)
                    diff = (1 - Fl) * base * (1 - metallic) / np.pi
                    t = NdotH * NdotH * (a2 - 1.0) + 1.0
                    spec = a2 / (np.pi * t * t) * Fl * g1(NdotV, k) * g1(NdotL, k) / (4 * NdotL * NdotV + 0.0001)
                    want = want + (diff + spec) * (li[4:7] / d2) * NdotL
            err = np.abs(got[y, x, :3] - want) / (np.abs(want) + 1e-3)
            worst = max(worst, err.max())
            assert got[y, x, 3] == 1.0
    assert worst < 1e-5, worst


def _bilinear_clamp64(img, u, v):
    """Vulkan linear filter, CLAMP_TO_EDGE, unnormalised coordinate = uv * size - 0.5; float64. img (h, w, 4); u, v arrays."""
    h, w = img.shape[:2]
    x, y = u * w - 0.5, v * h - 0.5
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = (x - x0)[..., None], (y - y0)[..., None]
    xi0, xi1 = np.clip(x0, 0, w - 1).astype(int), np.clip(x0 + 1, 0, w - 1).astype(int)
    yi0, yi1 = np.clip(y0, 0, h - 1).astype(int), np.clip(y0 + 1, 0, h - 1).astype(int)
    top = img[yi0, xi0] * (1 - fx) + img[yi0, xi1] * fx
    bot = img[yi1, xi0] * (1 - fx) + img[yi1, xi1] * fx
    return top * (1 - fy) + bot * fy


@pytest.mark.parametrize("W,H", [(64, 40), (37, 23)])
def test_glossy_convolve_against_an_independent_float64_evaluation(oracle, W, H):
    """Second, independent evaluation of Shaders/SSRGlossyConvolve.comp:26-55 with the host loop of
    Src/ReflectionBuffer.cpp:224-278 (level sizes, alternating direction, offsets divided by the TARGET width, texel
    coordinate without the half-texel), in numpy float64. Each level is evaluated from the oracle's previous level, so
    differences cannot compound; the oracle's halves must be the float64 value rounded to half, give or take one ulp
    where the fp32 sum lands on the other side of a rounding boundary."""
    rs = np.random.default_rng(3)
    mip0 = rs.uniform(0, 4, (H, W, 4)).astype(np.float16)
    mip0[..., 3] = rs.uniform(0, 1, (H, W)) * (rs.uniform(size=(H, W)) < 0.6)
    chain = oracle.glossy_convolve(mip0.view(np.uint16), 5).view(np.float16)
    offs = [1.411764705882353, 3.2941176470588234, 5.176470588235294]
    wts = [0.2969069646728344, 0.09447039785044732, 0.010381362401148057]
    base, prev = 0, mip0.astype(np.float64)
    for level in range(1, 5):
        base += prev.shape[0] * prev.shape[1] * 4
        w, h = max(W >> level, 1), max(H >> level, 1)
        got = chain[base:base + w * h * 4].reshape(h, w, 4)
        d = np.array([0.0, 1.0]) if level & 1 else np.array([1.0, 0.0])
        yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        u, v = xx / w, yy / h
        want = _bilinear_clamp64(prev, u, v) * 0.1964825501511404
        for o, wt in zip(offs, wts):
            du, dv = o * d[0] / w, o * d[1] / w
            want += (_bilinear_clamp64(prev, u + du, v + dv) + _bilinear_clamp64(prev, u - du, v - dv)) * wt
        want16 = want.astype(np.float16)
        ulp = np.abs(got.view(np.int16).astype(np.int32) - want16.view(np.int16).astype(np.int32))
        assert ulp.max() <= 1, (level, ulp.max())
        assert (ulp == 0).mean() > 0.97, (level, (ulp == 0).mean())
        prev = got.astype(np.float64)


def test_glossy_convolve_keeps_a_constant_image_constant(oracle):
    """The seven weights sum to 1 (to 1e-16), so a constant reflection buffer stays constant through all four levels."""
    mip0 = np.full((48, 80, 4), 0.75, np.float16)
    chain = oracle.glossy_convolve(mip0.view(np.uint16), 5).view(np.float16)
    assert np.all(chain == np.float16(0.75))
