"""The on-disk IBL cache in the reference's format: RGBE codec properties on CPU (against the
oracle-side codec and, when mounted, the reference's shipped files), and the GPU round trip precompute -> .hdr -> reload."""
import os

import numpy as np
import pytest

from althea_b200 import hdr_cache
from helpers import GOLDEN, REFERENCE, golden_env
from oracle import hdrio


def test_rgbe_codec_matches_stb_semantics(tmp_path):
    rs = np.random.default_rng(3)
    img = np.exp(rs.uniform(-12, 9, (37, 53, 3))).astype(np.float32)
    img[0, 0] = 0.0
    img[1, 1] = (1e-35, 0.0, 0.0)      # below stb's 1e-32 cut-off -> (0,0,0,0)
    img[2, 2] = (1000.0, 1e-3, 1.0)    # shared exponent: small channels truncate to 0
    p = str(tmp_path / "a" / "t.hdr")
    hdr_cache.write_hdr(p, img)
    back = hdr_cache.read_hdr(p)
    assert back.shape == img.shape
    assert np.array_equal(back, hdrio.read_hdr(p))            # same decode as the oracle-side reader
    assert (back <= img + 1e-30).all()                         # truncation never rounds up
    step = img.max(axis=-1, keepdims=True) / 128.0             # one 8-bit mantissa step of the largest channel
    assert (img - back <= step * 1.0001 + 1e-30).all()
    assert not back[0, 0].any() and not back[1, 1].any()
    # idempotent: re-encoding decoded values reproduces the same bytes
    assert np.array_equal(hdr_cache.float_to_rgbe(back), hdr_cache.float_to_rgbe(hdr_cache.rgbe_to_float(hdr_cache.float_to_rgbe(img))))
    # narrow images are written flat, as stb does
    hdr_cache.write_hdr(str(tmp_path / "n.hdr"), img[:, :5])
    assert np.array_equal(hdr_cache.read_hdr(str(tmp_path / "n.hdr")), hdr_cache.rgbe_to_float(hdr_cache.float_to_rgbe(img[:, :5])))


def test_reads_the_committed_fixture_like_the_oracle_reader():
    p = os.path.join(GOLDEN, "env_512x256.hdr")
    assert np.array_equal(hdr_cache.read_hdr(p), hdrio.read_hdr(p))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference not mounted (GPU box)")
def test_reads_the_reference_shipped_cache_and_rewrites_it_losslessly(tmp_path):
    src = os.path.join(REFERENCE, "Content/PrecomputedMaps/LuxuryRoom/Prefiltered4.hdr")
    a = hdr_cache.read_hdr(src)
    assert a.shape == (128, 256, 3)
    assert np.array_equal(a, hdrio.read_hdr(src))
    out = str(tmp_path / "again.hdr")
    hdr_cache.write_hdr(out, a)
    assert np.array_equal(hdr_cache.read_hdr(out), a)          # decoded RGBE values survive a rewrite exactly
    assert hdr_cache.cache_complete(os.path.join(REFERENCE, "Content"), "LuxuryRoom")
    assert not hdr_cache.cache_complete(os.path.join(REFERENCE, "Content"), "NoSuchEnv")


@pytest.mark.gpu
def test_cache_round_trip_on_gpu(tmp_path, ctx_fast):
    from althea_b200 import _capi, engine
    env = golden_env()[::2, ::2].copy()  # 256 x 128
    res = engine.ImageBasedLighting.createResources(ctx_fast, env, lut_size=32)
    content = str(tmp_path / "Content")
    assert not hdr_cache.cache_complete(content, "TestEnv")
    hdr_cache.save_precomputed_maps(content, "TestEnv", res.irradianceMap, res.prefilteredMap)
    assert hdr_cache.cache_complete(content, "TestEnv")
    irr, pre = hdr_cache.load_precomputed_maps(content, "TestEnv")
    gi = res.irradianceMap.level_numpy(0).view(np.float32).reshape(128, 256, 4)
    assert irr.shape == gi.shape and (irr[..., 3] == 1).all()
    assert (np.abs(irr[..., :3] - gi[..., :3]) <= gi[..., :3].max(axis=-1, keepdims=True) / 128.0 + 1e-12).all()
    assert [p.shape[:2] for p in pre] == [(64, 128), (32, 64), (16, 32), (8, 16), (4, 8)]
    for k, p in enumerate(pre):
        g = res.prefilteredMap.level_numpy(k).view(np.float32).reshape(p.shape)
        assert (np.abs(p[..., :3] - g[..., :3]) <= g[..., :3].max(axis=-1, keepdims=True) / 128.0 + 1e-12).all()
    # what the engine would re-load goes back to the device as the run-time IBL set
    F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
    flat = np.concatenate([p.reshape(-1) for p in pre])
    reloaded = ctx_fast.image_from_numpy(flat, F32, 128, 64, 5)
    assert reloaded.mips == 5
