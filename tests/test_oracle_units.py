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
