"""GPU parity tests of the IBL precompute (parity mode = the reference's equirect layout + hash RNG; benchmark mode =
BASELINE config 2's cube layout + Hammersley), through the C ABI, against the pinned CPU oracle."""
import numpy as np
import pytest

from helpers import golden_env

pytestmark = pytest.mark.gpu


def rel_close(a, b, tol):
    return np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))


@pytest.fixture(scope="module")
def env_setup(ctx_fast, oracle):
    import torch

    from althea_b200 import _capi, engine
    env = golden_env()  # 512 x 256
    H, W = env.shape[:2]
    chain_ref, mips = oracle.env_mip_chain(env)
    img = ctx_fast.new_image(_capi.FORMAT_R32G32B32A32_SFLOAT, W, H, mips)
    img.tensor[: W * H * 16].copy_(torch.from_numpy(env.view(np.uint8).reshape(-1)))
    engine.ImageBasedLighting.generateMipMaps(ctx_fast, img)
    torch.cuda.synchronize()
    return dict(env=env, W=W, H=H, mips=mips, chain_ref=chain_ref, img=img)


def test_mip_chain_is_the_box_chain(env_setup):
    got = env_setup["img"].tensor.cpu().numpy().view(np.float32)
    assert env_setup["mips"] == 10
    assert np.array_equal(got, env_setup["chain_ref"])  # 2:1 LINEAR blit == exact 2x2 box: bit-identical


def _probes(w, h, n, seed):
    rs = np.random.default_rng(seed)
    pts = {(0, 0), (w - 1, h - 1), (w // 2, h // 2), (0, h - 1)}
    while len(pts) < n:
        pts.add((int(rs.integers(0, w)), int(rs.integers(0, h))))
    return sorted(pts)


def test_irradiance_and_prefilter_reference_layout(ctx_fast, oracle, env_setup):
    import torch

    from althea_b200 import _capi, engine
    W, H, mips = env_setup["W"], env_setup["H"], env_setup["mips"]
    F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
    ow, oh = 96, 48  # output size is free in the C ABI; the reference uses the env size (too slow for the CPU oracle here)
    irr = ctx_fast.new_image(F32, ow, oh)
    pre = ctx_fast.new_image(F32, W >> 1, H >> 1, 5)
    engine.ImageBasedLighting.precomputeResources(ctx_fast, env_setup["img"], irr, pre)
    torch.cuda.synchronize()
    got_irr = irr.level_numpy(0).view(np.float32).reshape(oh, ow, 4)
    pts = _probes(ow, oh, 24, 1)
    want = oracle.ibl_irradiance(env_setup["chain_ref"], W, H, mips, ow, oh, [(x, y, 0) for x, y in pts])
    got = np.array([got_irr[y, x] for x, y in pts])
    assert rel_close(got, want, 1e-3).all(), np.abs(got - want).max()
    assert np.isfinite(got_irr).all() and (got_irr[..., 3] == 1).all()
    for level in range(5):
        lw, lh = (W >> 1) >> level, (H >> 1) >> level
        got_l = pre.level_numpy(level).view(np.float32).reshape(lh, lw, 4)
        assert np.isfinite(got_l).all()
        pts = _probes(lw, lh, 12, 10 + level)
        want = oracle.ibl_prefilter(env_setup["chain_ref"], W, H, mips, lw, lh, level / 4.0, [(x, y, 0) for x, y in pts])
        got = np.array([got_l[y, x] for x, y in pts])
        assert rel_close(got, want, 1e-3).all(), (level, float(np.max(np.abs(got - want) / np.maximum(1, np.abs(want)))))


def test_cube_layout_hammersley(ctx_fast, oracle, env_setup):
    import torch

    from althea_b200 import _capi, engine
    W, H, mips = env_setup["W"], env_setup["H"], env_setup["mips"]
    F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
    irr = ctx_fast.new_image(F32, 16, 16, 1, 6)
    pre = ctx_fast.new_image(F32, 32, 32, 4, 6)
    engine.ImageBasedLighting.precomputeResources(ctx_fast, env_setup["img"], irr, pre, layout=_capi.IBL_LAYOUT_CUBE,
                                                  sequence=_capi.IBL_SEQ_HAMMERSLEY, prefilter_samples=2048)
    torch.cuda.synchronize()
    tex = [(x, y, f) for f in range(6) for (x, y) in ((0, 0), (7, 9), (15, 15))]
    want = oracle.ibl_irradiance(env_setup["chain_ref"], W, H, mips, 16, 16, tex, layout=oracle.LAYOUT_CUBE)
    got = np.array([irr.level_numpy(0, f).view(np.float32).reshape(16, 16, 4)[y, x] for x, y, f in tex])
    assert rel_close(got, want, 1e-3).all()
    for level in range(4):
        s = 32 >> level
        tex = [(x % s, y % s, f) for f in range(6) for (x, y) in ((0, 0), (5, 3), (31, 31))]
        want = oracle.ibl_prefilter(env_setup["chain_ref"], W, H, mips, s, s, level / 3.0, tex, layout=oracle.LAYOUT_CUBE,
                                    num_samples=2048, sequence=oracle.SEQ_HAMMERSLEY)
        got = np.array([pre.level_numpy(level, f).view(np.float32).reshape(s, s, 4)[y, x] for x, y, f in tex])
        assert rel_close(got, want, 1e-3).all(), level


def test_config1_shape_probe_texels(ctx_fast, oracle):
    """BASELINE configs[1] at its own shape: 4096 x 2048 equirect environment -> 32^2 irradiance cube (300 x 150 samples) and
    512^2 six-mip GGX prefilter cube with 10 000 Hammersley samples per texel; probe texels of levels 0, 3 and 5 (and of the
    irradiance cube) against the oracle."""
    import torch

    from althea_b200 import _capi, engine, scene
    W, H = 4096, 2048
    env = scene.procedural_env(W, H).numpy()
    chain_ref, mips = oracle.env_mip_chain(env)
    F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
    img = ctx_fast.new_image(F32, W, H, mips)
    img.tensor[: W * H * 16].copy_(torch.from_numpy(env.view(np.uint8).reshape(-1)))
    engine.ImageBasedLighting.generateMipMaps(ctx_fast, img)
    irr = ctx_fast.new_image(F32, 32, 32, 1, 6)
    pre = ctx_fast.new_image(F32, 512, 512, 6, 6)
    engine.ImageBasedLighting.precomputeResources(ctx_fast, img, irr, pre, layout=_capi.IBL_LAYOUT_CUBE, sequence=_capi.IBL_SEQ_HAMMERSLEY,
                                                  prefilter_samples=10000)
    torch.cuda.synchronize()
    tex = [(x, y, f) for f in range(6) for (x, y) in ((0, 0), (13, 21), (31, 31))]
    want = oracle.ibl_irradiance(chain_ref, W, H, mips, 32, 32, tex, layout=oracle.LAYOUT_CUBE)
    got = np.array([irr.level_numpy(0, f).view(np.float32).reshape(32, 32, 4)[y, x] for x, y, f in tex])
    assert rel_close(got, want, 1e-3).all(), float(np.abs(got - want).max())
    for level in (0, 3, 5):
        s = 512 >> level
        tex = [(x % s, y % s, f) for f in range(6) for (x, y) in ((0, 0), (101, 377), (511, 511))]
        want = oracle.ibl_prefilter(chain_ref, W, H, mips, s, s, level / 5.0, tex, layout=oracle.LAYOUT_CUBE, num_samples=10000, sequence=oracle.SEQ_HAMMERSLEY)
        got = np.array([pre.level_numpy(level, f).view(np.float32).reshape(s, s, 4)[y, x] for x, y, f in tex])
        assert rel_close(got, want, 1e-3).all(), (level, float(np.max(np.abs(got - want) / np.maximum(1, np.abs(want)))))
        assert np.isfinite(pre.level_numpy(level, 0).view(np.float32)).all()


def test_brdf_lut(ctx_fast, oracle):
    import torch

    from althea_b200 import _capi, engine
    size = 64
    lut32 = ctx_fast.new_image(_capi.FORMAT_R32G32B32A32_SFLOAT, size, size)
    lut8 = ctx_fast.new_image(_capi.FORMAT_R8G8B8A8_UNORM, size, size)
    engine.ImageBasedLighting.generateBrdfLut(ctx_fast, lut32, 512)
    engine.ImageBasedLighting.generateBrdfLut(ctx_fast, lut8, 512)
    torch.cuda.synchronize()
    want = oracle.brdf_lut(size, 512, 0)[::-1]  # the kernel writes the reference asset's orientation: row 0 = roughness 1
    got = lut32.level_numpy(0).view(np.float32).reshape(size, size, 4)
    assert np.abs(got[..., :2] - want).max() < 2e-4
    assert (got[..., 2] == 0).all() and (got[..., 3] == 1).all()
    got8 = lut8.level_numpy(0).reshape(size, size, 4)
    assert np.abs(got8[..., :2].astype(np.float32) / 255.0 - np.clip(want, 0, 1)).max() <= 0.5 / 255 + 2e-4
    assert (got8[..., 3] == 255).all()


def test_create_resources_shapes(ctx_fast):
    from althea_b200 import engine
    env = golden_env()[::4, ::4].copy()  # 128 x 64
    res = engine.ImageBasedLighting.createResources(ctx_fast, env, lut_size=32)
    assert (res.irradianceMap.w, res.irradianceMap.h) == (128, 64)
    assert (res.prefilteredMap.w, res.prefilteredMap.h, res.prefilteredMap.mips) == (64, 32, 5)
    assert (res.brdfLut.w, res.brdfLut.h) == (32, 32)
    # roughness 0 level == corner-aligned bilinear decimation of the env map (what Prefiltered1.hdr holds)
    p0 = res.prefilteredMap.level_numpy(0).view(np.float32).reshape(32, 64, 4)[..., :3]
    e = env[..., :3]
    want = 0.25 * (e[1:-1:2, 1:-1:2] + e[1:-1:2, 2::2] + e[2::2, 1:-1:2] + e[2::2, 2::2])  # texels (2x-1,2x) x (2y-1,2y)
    assert np.abs(p0[1:, 1:] - want).max() <= 2e-3 * max(1.0, float(want.max()))
