"""GPU parity tests of the per-frame stages, through the C ABI, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): per-channel max abs error <= 1e-3 and PSNR >= 50 dB on linear-HDR outputs; AO and SSR
hit masks within 0.1 % of pixels. "abs 1e-3" is applied as |a-b| <= 1e-3 * max(1, |b|): the outputs are HDR (values up
to ~150), and RGBA16F targets carry half-precision rounding (1 ulp = 2^-10 relative), so a pure absolute bound is
meaningless above 1.0. The parity build (-fmad=false) is additionally required to reproduce the oracle's threshold
decisions EXACTLY: SSAO counts and SSR hit masks are integer work and the bar for those is bit-exact.
"""
import os

import numpy as np
import pytest

from helpers import GOLDEN, FrameData, GpuFrame, half_to_float, psnr

pytestmark = pytest.mark.gpu

TOL = 1e-3
MASK_BAR = 1e-3  # 0.1 % of pixels


def close(a, b, tol=TOL):
    return np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))


@pytest.fixture(scope="module")
def frames():
    return {"scene": FrameData("scene", 256, 144, n_lights=4, shadow_res=64), "rand": FrameData("rand", 160, 90, n_lights=3, shadow_res=32)}


@pytest.fixture(scope="module")
def oracle_out(frames, oracle):
    out = {}
    for k, fd in frames.items():
        fr = fd.oracle_frame()
        refl, hit, steps = oracle.ssr_capture(fr)
        chain = oracle.glossy_convolve(refl)
        ao = oracle.ssao(fr)
        col = oracle.deferred_shade(fr, chain, 5, oracle.SKIP_TONEMAP, ao)
        out[k] = dict(refl=refl, hit=hit, chain=chain, ao=ao, color=col, frame=fr)
    return out


def _ctx(request, which):
    return request.getfixturevalue("ctx_parity" if which == "parity" else "ctx_fast")


@pytest.mark.parametrize("which", ["parity", "fast"])
@pytest.mark.parametrize("kind", ["scene", "rand"])
def test_ssr_capture(request, frames, oracle_out, which, kind):
    ctx = _ctx(request, which)
    fd, ref = frames[kind], oracle_out[kind]
    gf = GpuFrame(ctx, fd)
    gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
    got = gf.reflection_level(0)
    got_f, ref_f = half_to_float(got), half_to_float(ref["refl"])
    hit = got_f[..., 3] != 0
    mismatch = float(np.mean(hit != (ref["hit"] != 0)))
    assert ref["hit"].mean() > 0.05, "test scene must produce SSR hits"
    if which == "parity":
        assert mismatch == 0.0
    else:
        assert mismatch <= MASK_BAR
    agree = hit == (ref["hit"] != 0)
    ok = close(got_f, ref_f)[agree]
    if which == "parity":
        assert ok.all(), "max rel err %g" % float(np.max(np.abs(got_f - ref_f)[agree] / np.maximum(1, np.abs(ref_f))[agree]))
    else:
        # FFMA contraction can move a march's terminating step by one (same hit flag, neighbouring hit point). Those are
        # threshold flips as well, so they get the same 0.1 % bar as the mask itself (S-rand, with an unrelated depth in
        # every pixel, is the worst case: every step of every march sits next to a discontinuity)
        moved = float(np.mean(~ok.all(axis=-1))) * float(agree.mean())
        assert moved <= MASK_BAR, "moved hit points %g" % moved
    assert psnr(got_f[agree], ref_f[agree]) >= 50.0
    # empty pixels (normal.a == 0) write exactly (0,0,0,0)
    empty = fd.normal[..., 3] == 0
    assert not got[empty].any()


@pytest.mark.parametrize("which", ["parity", "fast"])
@pytest.mark.parametrize("kind", ["scene", "rand"])
def test_glossy_convolve(request, frames, oracle_out, which, kind):
    ctx = _ctx(request, which)
    fd, ref = frames[kind], oracle_out[kind]
    gf = GpuFrame(ctx, fd)
    import torch
    img = gf.ssr.getReflectionBuffer().image
    n0 = fd.W * fd.H * 8
    img.tensor[:n0].copy_(torch.from_numpy(ref["refl"].view(np.uint8).reshape(-1)))  # stage-isolated: oracle's mip 0 in
    gf.ssr.convolveReflectionBuffer()
    got = half_to_float(gf.reflection_chain())
    want = half_to_float(ref["chain"])
    assert np.array_equal(got[: fd.W * fd.H * 4], want[: fd.W * fd.H * 4])
    # both sides round to RGBA16F: allow one half ulp flip (2^-10 relative), nothing more
    assert close(got, want, 2.0 ** -10 + 1e-6).all()
    if which == "parity":
        assert np.mean(got != want) < 1e-3  # essentially bit-identical


@pytest.mark.parametrize("which", ["parity", "fast"])
@pytest.mark.parametrize("kind", ["scene", "rand"])
def test_ssao_counts(request, frames, oracle_out, which, kind):
    ctx = _ctx(request, which)
    fd, ref = frames[kind], oracle_out[kind]
    gf = GpuFrame(ctx, fd)
    from althea_b200 import _capi
    gf.ssr.convolveReflectionBuffer()
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
    got = gf.ao_counts()
    want = ref["ao"]
    assert (want[want < 255] > 0).mean() > 0.02, "test scene must produce occlusion"
    assert np.array_equal(got == 255, want == 255)
    mismatch = float(np.mean(got != want))
    if which == "parity":
        assert mismatch == 0.0
    else:
        assert mismatch <= MASK_BAR
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 2


@pytest.mark.parametrize("which", ["parity", "fast"])
@pytest.mark.parametrize("kind", ["scene", "rand"])
def test_deferred_shade_given_ao_and_reflection(request, frames, oracle_out, which, kind):
    """Shading parity in isolation: the oracle's reflection chain and AO counts go in, so threshold flips upstream cannot
    leak into this comparison."""
    ctx = _ctx(request, which)
    fd, ref = frames[kind], oracle_out[kind]
    gf = GpuFrame(ctx, fd)
    import torch
    from althea_b200 import _capi
    gf.ssr.getReflectionBuffer().image.tensor.copy_(torch.from_numpy(ref["chain"].view(np.uint8).reshape(-1)))
    gf.deferred.aoCounts.tensor.copy_(torch.from_numpy(ref["ao"].reshape(-1)))
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP | _capi.SHADE_AO_FROM_IMAGE)
    got, want = gf.color(), ref["color"]
    assert np.isfinite(got).all()
    assert close(got, want).all(), "max rel err %g" % float(np.max(np.abs(got - want) / np.maximum(1, np.abs(want))))
    assert psnr(got, want) >= 50.0
    assert (got[..., 3] == 1.0).all()


@pytest.mark.parametrize("which", ["parity", "fast"])
def test_full_chain(request, frames, oracle_out, which):
    """SSR -> convolve -> SSAO -> shade, all on the GPU, against the oracle's chain."""
    ctx = _ctx(request, which)
    fd, ref = frames["scene"], oracle_out["scene"]
    gf = GpuFrame(ctx, fd)
    from althea_b200 import _capi
    gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
    gf.ssr.convolveReflectionBuffer()
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
    got, want = gf.color(), ref["color"]
    ok = close(got, want).all(axis=-1)
    if which == "parity":
        assert ok.mean() >= 0.9995  # only half-ulp flips of the RGBA16F reflection mips can differ
    else:
        assert ok.mean() >= 0.995   # a flipped SSR hit is blurred into its neighbours by the mip chain
    assert psnr(got, want) >= 50.0


def test_against_committed_golden(ctx_parity, oracle):
    """The committed vectors (tests/golden/frame_golden.npz, written by make_golden.py) must be reproduced on the GPU."""
    gold = np.load(os.path.join(GOLDEN, "frame_golden.npz"))
    from althea_b200 import _capi
    for tag, kind, W, H in (("scene", "scene", 160, 90), ("rand", "rand", 96, 54)):
        fd = FrameData(kind, W, H, n_lights=4, shadow_res=32)
        gf = GpuFrame(ctx_parity, fd)
        gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
        hit = half_to_float(gf.reflection_level(0))[..., 3] != 0
        assert np.array_equal(hit, gold[tag + "_hit"] != 0)
        gf.ssr.convolveReflectionBuffer()
        gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
        assert np.array_equal(gf.ao_counts(), gold[tag + "_ao"])
        got, want = gf.color(), gold[tag + "_color"]
        assert close(got, want).all(axis=-1).mean() >= 0.9995
        assert psnr(got, want) >= 50.0


# ---- edge cases -------------------------------------------------------------------------------------------------------
def test_odd_sizes_tonemap_and_half_output(ctx_parity, oracle):
    from althea_b200 import _capi
    fd = FrameData("scene", 101, 67, n_lights=2, shadow_res=16)
    fr = fd.oracle_frame()
    refl, hit, _ = oracle.ssr_capture(fr)
    chain = oracle.glossy_convolve(refl)
    ao = oracle.ssao(fr)
    want = oracle.deferred_shade(fr, chain, 5, 0, ao)  # tonemapped
    gf = GpuFrame(ctx_parity, fd, out_format=_capi.FORMAT_R16G16B16A16_SFLOAT)
    gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
    gf.ssr.convolveReflectionBuffer()
    got_chain = half_to_float(gf.reflection_chain())
    assert close(got_chain, half_to_float(chain), 2.0 ** -10 + 1e-6).all()  # mips 50x33, 25x16, 12x8, 6x4
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, 0)
    assert np.array_equal(gf.ao_counts(), ao)
    got = gf.color()
    assert close(got, want.astype(np.float16).astype(np.float32), 2.0 ** -10 + 1e-6).all(axis=-1).mean() >= 0.999
    assert got.max() <= 1.0  # 1 - exp(-c * exposure)


def test_empty_gbuffer_draws_environment(ctx_fast, oracle):
    from althea_b200 import _capi
    fd = FrameData("scene", 64, 36, n_lights=0)
    fd.position[:] = 0
    fd.normal[:] = 0
    fd.depth[:] = 1
    fd.albedo[:] = 0
    fd.mro[:] = 0
    gf = GpuFrame(ctx_fast, fd)
    gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, None)
    assert not gf.reflection_level(0).any()
    gf.ssr.convolveReflectionBuffer()
    assert not gf.reflection_chain().any()
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, None, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
    want = oracle.deferred_shade(fd.oracle_frame(), np.zeros_like(gf.reflection_chain()), 5, oracle.SKIP_TONEMAP, None)
    assert close(gf.color(), want).all()
    assert (gf.ao_counts() == 255).all()


def test_no_lights_and_no_shadow_maps(ctx_parity, oracle):
    from althea_b200 import _capi
    fd = FrameData("scene", 96, 54, n_lights=0)
    fr = fd.oracle_frame()
    refl, hit, _ = oracle.ssr_capture(fr)
    gf = GpuFrame(ctx_parity, fd)
    gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, None)
    assert np.array_equal(half_to_float(gf.reflection_level(0))[..., 3] != 0, hit != 0)
    assert close(half_to_float(gf.reflection_level(0)), half_to_float(refl)).all()


def test_error_behaviour(ctx_fast):
    """The reference throws std::runtime_error; the C ABI returns a status + message, the mirror raises AltheaError."""
    import ctypes as C

    from althea_b200 import _capi, engine
    fd = FrameData("scene", 32, 18, n_lights=1, shadow_res=8)
    gf = GpuFrame(ctx_fast, fd)
    lib, p = ctx_fast._lib, ctx_fast._ptr
    gb, ib = gf.gbuffer.struct(), gf.ibl.struct()
    # bad handle
    rc = lib.althea_cuda_glossy_convolve(p, 987654, None)
    assert rc == -4 and b"not a live image" in lib.althea_cuda_last_error(p)
    # wrong format for the reflection target
    rc = lib.althea_cuda_ssr_capture(p, C.byref(fd.uniforms), C.byref(gb), C.byref(ib), gf.lights.buffer.handle, gf.lights.shadow_handle,
                                     gf.gbuffer.albedo.handle, None)
    assert rc == -1 and b"expected VkFormat 97" in lib.althea_cuda_last_error(p)
    # size mismatch between G-buffer and reflection buffer
    small = engine.ScreenSpaceReflection(ctx_fast, 16, 16)
    with pytest.raises(engine.AltheaError, match="must all be"):
        small.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
    # lightCount larger than the light buffer
    u = _capi.GlobalUniforms.from_buffer_copy(bytes(fd.uniforms))
    u.lightCount = 9
    with pytest.raises(engine.AltheaError, match="lights_buf holds"):
        gf.ssr.captureReflection(u, gf.gbuffer, gf.ibl, gf.lights)
    # release invalidates
    img = ctx_fast.new_image(_capi.FORMAT_R8_UINT, 8, 8)
    h = img.handle
    img.release()
    assert lib.althea_cuda_release(p, h) == -4
    # host pointers are rejected by wrap
    import numpy as np
    host = np.zeros(64, np.uint8)
    out = C.c_uint64()
    rc = lib.althea_cuda_wrap_linear_image(p, host.ctypes.data_as(C.c_void_p), 0, _capi.FORMAT_R8_UINT, 8, 8, 1, 1, C.byref(out))
    assert rc == -1 and b"not device memory" in lib.althea_cuda_last_error(p)


def _ao_both_ways(ctx, gf, fd, parity):
    """SSAO counts from the packed-proxy kernels (position records: the default; ray-depth records: opt-in) and from the
    straight fp32-texel kernel, same context. Returns (position proxy, exact taps); asserts the ray-depth proxy equals them."""
    from althea_b200 import _capi
    base = _capi.CTX_PARITY_MATH if parity else 0
    ctx.set_flags(base | _capi.CTX_SSAO_RAY_DEPTH_PROXY)
    gf.deferred.aoCounts.tensor.zero_()
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
    ray = gf.ao_counts().copy()
    out = []
    for flags in (base, base | _capi.CTX_SSAO_EXACT_TAPS):
        ctx.set_flags(flags)
        gf.deferred.aoCounts.tensor.zero_()
        gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
        out.append(gf.ao_counts().copy())
    ctx.set_flags(base)
    assert np.array_equal(ray, out[1]), "ray-depth proxy differs from the fp32-texel march on %d pixels" % int((ray != out[1]).sum())
    return out


@pytest.mark.parametrize("which", ["parity", "fast"])
def test_ssao_proxy_equals_exact_taps_on_hostile_positions(request, which):
    """The filtered march must take the exact path whenever its error bound cannot decide: positions with NaN / inf / huge
    magnitudes, fp16-overflowing neighbour deltas, denormals and exact zeros must give the same counts as the fp32-texel march."""
    ctx = _ctx(request, which)
    fd = FrameData("scene", 192, 108, n_lights=0)
    rng = np.random.default_rng(7)
    pos = fd.position
    H, W = pos.shape[:2]
    n = 400
    ys, xs = rng.integers(0, H, n), rng.integers(0, W, n)
    vals = [np.nan, np.inf, -np.inf, 1e30, -3e38, 7e4, -7e4, 1e-40, 0.0, 65504.0 * 1.5]
    for k in range(n):
        pos[ys[k], xs[k], rng.integers(0, 3)] = vals[k % len(vals)]
    pos[60:70, 80:120, :3] *= 4000.0          # a far slab: neighbour deltas overflow fp16
    pos[20:24, 10:60, :3] = pos[20, 10, :3]   # a run of identical texels: projections that are exactly equal / zero
    gf = GpuFrame(ctx, fd)
    a, b = _ao_both_ways(ctx, gf, fd, which == "parity")
    assert np.array_equal(a, b)
    assert (b[b < 255] > 0).mean() > 0.02


@pytest.mark.parametrize("parity", [True, False])
def test_ssao_direction_table_is_what_the_hash_yields(lib_built, parity):
    """The SSAO rays' tangent-space directions come from a table built once per frame size (a function of pixel + 3 ray alone,
    params.h); ALTHEA_SSAO_DIR_TABLE=0 hashes them inline as the shader does. Same counts bit for bit in both builds, also after the
    table has been rebuilt for a larger frame."""
    from althea_b200 import _capi, engine
    os.environ["ALTHEA_SSAO_DIR_TABLE"] = "0"
    try:
        ctx_inline = engine.Context(0, parity_math=parity)
    finally:
        del os.environ["ALTHEA_SSAO_DIR_TABLE"]
    ctx_table = engine.Context(0, parity_math=parity)
    try:
        for kind, W, H in (("rand", 83, 47), ("scene", 200, 113)):  # the second frame is larger: the table grows
            fd = FrameData(kind, W, H, n_lights=0)
            counts = []
            for ctx in (ctx_inline, ctx_table):
                gf = GpuFrame(ctx, fd)
                gf.ssr.convolveReflectionBuffer()
                gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
                counts.append(gf.ao_counts().copy())
            assert (counts[0][counts[0] < 255] > 0).mean() > 0.02
            assert np.array_equal(counts[0], counts[1])
    finally:
        ctx_inline.close()
        ctx_table.close()


# ---- size-independent properties at the full 4K size (BASELINE configs C3/C5) -------------------------------------------
def test_properties_at_4k(ctx_fast):
    import torch

    from althea_b200 import _capi, engine, scene
    W, H = 3840, 2160
    dev = "cuda:0"
    g = scene.make_uniforms(W, H, pos=(0.0, 2.0, 6.0), yaw=0.0, pitch=-0.25, light_count=0)
    gbd = scene.s_scene(g, W, H, scene.make_scene(64, device=dev), device=dev)
    gb = engine.GBufferResources(ctx_fast, W, H)
    gb.upload(position=gbd.position, depth=gbd.depth, normal=gbd.normal, albedo=gbd.albedo, mro=gbd.mro)
    fd = FrameData("scene", 32, 18, n_lights=0)
    small = GpuFrame(ctx_fast, fd)  # reuse its IBL set
    ssr = engine.ScreenSpaceReflection(ctx_fast, W, H)
    dp = engine.DeferredPass(ctx_fast, W, H, _capi.FORMAT_R16G16B16A16_SFLOAT)

    def run():
        ssr.captureReflection(g, gb, small.ibl, None)
        ssr.convolveReflectionBuffer()
        dp.draw(g, gb, small.ibl, None, ssr, _capi.SHADE_SKIP_TONEMAP)
        torch.cuda.synchronize()
        return ssr.getReflectionBuffer().image.tensor.clone(), dp.colorTarget.tensor.clone(), dp.aoCounts.tensor.clone()

    r1, c1, a1 = run()
    r2, c2, a2 = run()
    assert torch.equal(r1, r2) and torch.equal(c1, c2) and torch.equal(a1, a2)  # deterministic / idempotent
    # the packed-proxy march against the fp32-texel march. With -fmad=false every decision is the same IEEE
    # operation in both kernels: bit for bit. In the fast build the compiler contracts the two kernels' inlined copies of
    # the tap arithmetic separately, so a projection within an ulp of zero can flip: at most a handful of pixels in 8.3 M.
    # Round 2: the default path classifies taps from per-block plane records first (ssao_cull_kernel); CTX_SSAO_NO_CULL is the
    # round-1 march over the position records. Both against the fp32-texel march.
    NC = _capi.CTX_SSAO_NO_CULL
    for flags, bar in ((_capi.CTX_PARITY_MATH, 0), (0, 8), (_capi.CTX_PARITY_MATH | NC, 0), (NC, 8), (_capi.CTX_PARITY_MATH | _capi.CTX_SSAO_RAY_DEPTH_PROXY, 0),
                       (_capi.CTX_SSAO_RAY_DEPTH_PROXY, 8)):
        ctx_fast.set_flags(flags)
        _, _, ap = run()
        ctx_fast.set_flags(flags | _capi.CTX_SSAO_EXACT_TAPS)
        _, _, ae = run()
        ctx_fast.set_flags(0)
        assert int((ap != ae).sum()) <= bar, "proxy vs fp32-texel SSAO counts differ on %d pixels (flags %d)" % (int((ap != ae).sum()), flags)
        assert int((ap.int() - ae.int()).abs().max()) <= 1
    # SSR with the plane-record sign test (opt-in) against the plain march: the same reflection image bit for bit in the parity
    # build, hit masks within the 0.1 % bar in the fast build (whose folded step algebra differs between the two kernels)
    n0r = W * H * 8
    for flags, bar in ((_capi.CTX_PARITY_MATH, 0.0), (0, MASK_BAR)):
        ctx_fast.set_flags(flags)
        rm, _, _ = run()
        ctx_fast.set_flags(flags | _capi.CTX_SSR_PLANE_SKIP)
        rs, _, _ = run()
        ctx_fast.set_flags(0)
        am = rm[:n0r].view(torch.float16).view(H, W, 4)[..., 3] != 0
        as_ = rs[:n0r].view(torch.float16).view(H, W, 4)[..., 3] != 0
        assert float((am != as_).float().mean()) <= bar
        if bar == 0.0:
            assert torch.equal(rm, rs)
    ao = a1.view(H, W)
    empty = gbd.position[..., 3] == 0
    assert bool((ao[empty] == 255).all()) and bool((ao[~empty] <= 24).all())
    refl0 = r1[: W * H * 8].view(torch.float16).view(H, W, 4).float()
    assert bool((refl0[empty.to(dev)] == 0).all())                # sky writes (0,0,0,0)
    alpha = refl0[..., 3]
    assert bool(((alpha == 0) | (alpha == 1)).all())                # hit alpha is exactly 0 or 1
    assert 0.02 < float(alpha.mean()) < 0.6
    col = c1.view(torch.float16).view(H, W, 4).float()
    assert bool(torch.isfinite(col).all()) and bool((col[..., 3] == 1).all())
    # convolve properties: a constant image stays constant (weights sum to 1 within rounding); linear in its input
    rb = engine.ReflectionBuffer(ctx_fast, W, H)
    n0 = W * H * 4
    t16 = rb.image.tensor.view(torch.float16)
    t16[:n0] = 0.75
    rb.convolveReflectionBuffer()
    torch.cuda.synchronize()
    assert float((t16.float() - 0.75).abs().max()) <= 2.0 ** -10
    t16[:n0] = refl0.reshape(-1).half()
    rb.convolveReflectionBuffer()
    torch.cuda.synchronize()
    base = t16.float().clone()
    assert torch.equal(t16[n0:].view(torch.int16), r1.view(torch.int16)[n0:])  # same input -> same mips as the SSR buffer
    t16[:n0] = (refl0.reshape(-1) * 0.5).half()
    rb.convolveReflectionBuffer()
    torch.cuda.synchronize()
    scaled = t16.float()
    big = base.abs() > 1e-3  # away from fp16 subnormals scaling by 2 is exact
    assert float(((scaled * 2.0 - base).abs()[big] / base.abs()[big]).max()) <= 2.0 ** -9


def test_gather_diagnostics(ctx_fast):
    """bench.py's roofline evidence for SSAO: the counting instantiations return the same AO counts. The round-1 march
    (CTX_SSAO_NO_CULL) gathers at most 24 rays x 11 taps of position records per shaded pixel; with the coarse sign test the
    plane-record lookups obey that bound instead, the taps it could not call are a fraction of them, and the records gathered
    are one per such tap plus two per step that changes sign. The ceiling measurement returns a rate."""
    from althea_b200 import _capi
    fd = FrameData("scene", 192, 108, n_lights=0)
    gf = GpuFrame(ctx_fast, fd)
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
    plain = gf.ao_counts().copy()
    shaded = int((plain < 255).sum())
    try:
        ctx_fast.set_flags(_capi.CTX_SSAO_COUNT_TAPS | _capi.CTX_SSAO_NO_CULL)
        gf.deferred.aoCounts.tensor.zero_()
        gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
        assert np.array_equal(plain, gf.ao_counts())
        gathers = ctx_fast.ssao_gathers()
        assert 0 <= ctx_fast.ssao_exact_fallbacks() < gathers // 4
        assert shaded * 24 < gathers <= shaded * 24 * 11  # more than one tap per ray on average
        ctx_fast.set_flags(_capi.CTX_SSAO_COUNT_TAPS)
        gf.deferred.aoCounts.tensor.zero_()
        gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
        assert np.array_equal(plain, gf.ao_counts())
        c = ctx_fast.ssao_cull_counts()
        assert shaded * 24 < c["plane_lookups"] <= shaded * 24 * 11
        assert 0 < c["exact_steps"] < c["plane_lookups"]           # the taps the plane records could not call (most, at this size:
        # a ray is ~8 texels long here; ~11 % at 4K)
        assert c["exact_steps"] <= c["records"] <= c["exact_steps"] + 2 * shaded * 24 * 10
    finally:
        ctx_fast.set_flags(0)
    rate = ctx_fast.gather_ceiling(512, 512, 32, 16)
    assert rate > 1e9


@pytest.mark.parametrize("which", ["parity", "fast"])
def test_mode_d_positions_from_depth(request, oracle, which):
    """A G-buffer without the legacy position attachment (today's GBufferResources): the lighting pass and SSAO work on
    positions reconstructed from depth, empty where normal.a == 0. Equal to the mode-P path fed with the restatement's
    reconstructPosition of every pixel: bit for bit in the parity build, to the mask / colour bars in the fast build."""
    from althea_b200 import _capi, engine
    ctx = _ctx(request, which)
    fd = FrameData("scene", 160, 90, n_lights=2, shadow_res=32)
    H, W = fd.H, fd.W
    og = oracle.GlobalUniforms.from_buffer_copy(bytes(fd.uniforms))
    nrm_a = half_to_float(fd.normal)[..., 3]
    pos = np.zeros((H, W, 4), np.float32)
    for y in range(H):
        for x in range(W):
            if nrm_a[y, x] != 0:
                pos[y, x, :3] = oracle.reconstruct_position(og, np.float32(x + 0.5) / np.float32(W), np.float32(y + 0.5) / np.float32(H), fd.depth[y, x])
                pos[y, x, 3] = 1.0
    fd.position = pos  # the mode-P twin: the same positions handed over as an attachment
    gp = GpuFrame(ctx, fd)
    gp.deferred.draw(fd.uniforms, gp.gbuffer, gp.ibl, gp.lights, gp.ssr, _capi.SHADE_SKIP_TONEMAP)
    want_ao, want_col = gp.ao_counts().copy(), gp.color().copy()
    gd = GpuFrame(ctx, fd)
    gd.gbuffer = engine.GBufferResources(ctx, W, H, with_position=False)
    gd.gbuffer.upload(depth=fd.depth, normal=fd.normal, albedo=fd.albedo, mro=fd.mro)
    gd.deferred.draw(fd.uniforms, gd.gbuffer, gd.ibl, gd.lights, gd.ssr, _capi.SHADE_SKIP_TONEMAP)
    got_ao, got_col = gd.ao_counts(), gd.color()
    if which == "parity":
        assert np.array_equal(got_ao, want_ao) and np.array_equal(got_col, want_col)
    else:
        assert (got_ao != want_ao).mean() <= 1e-3
        ok = (np.abs(got_col - want_col) <= 1e-3 * np.maximum(1.0, np.abs(want_col))).all(axis=-1)
        assert ok.mean() >= 0.999
    assert (got_ao[nrm_a == 0] == 255).all() and (got_ao[nrm_a != 0] <= 24).all()
    # ... and against the ORACLE directly (mode D in the fast build is what bench.py times): SSR hit mask, AO counts, colour
    gd.ssr.captureReflection(fd.uniforms, gd.gbuffer, gd.ibl, gd.lights)
    gd.ssr.convolveReflectionBuffer()
    gd.deferred.draw(fd.uniforms, gd.gbuffer, gd.ibl, gd.lights, gd.ssr, _capi.SHADE_SKIP_TONEMAP)
    fr = fd.oracle_frame()  # fd.position now holds the restatement's reconstructed positions
    refl, hit, _ = oracle.ssr_capture(fr)
    chain = oracle.glossy_convolve(refl)
    ao = oracle.ssao(fr)
    col = oracle.deferred_shade(fr, chain, 5, oracle.SKIP_TONEMAP, ao)
    got_hit = half_to_float(gd.reflection_level(0))[..., 3] != 0
    got_ao, got_col = gd.ao_counts(), gd.color()
    if which == "parity":
        assert np.array_equal(got_hit, hit != 0) and np.array_equal(got_ao, ao)
    else:
        assert (got_hit != (hit != 0)).mean() <= MASK_BAR and (got_ao != ao).mean() <= MASK_BAR
    same = (got_hit == (hit != 0)) & (got_ao == ao)
    ok = close(got_col, col).all(axis=-1)
    assert ok[same].mean() >= 0.999, "max abs colour error %.3g" % np.abs(got_col - col)[same].max()


def test_4k_frame_against_the_oracle(ctx_parity, oracle):
    """BASELINE configs[2]'s shape against the oracle itself (not against another GPU path): one 3840 x 2160 S-scene frame, four
    lights with 64^2 shadow cubes, parity build, mode D. SSR hit mask and SSAO counts bit for bit over all 8.3 M pixels (the
    coarse sign test, the tap compaction and the position-record filter all sit in front of these decisions at this size);
    colour on the RGBA32F target to the relative bar, with the absolute error reported."""
    import torch

    from althea_b200 import _capi, engine, scene
    from helpers import golden_env, golden_lut, ibl_standins
    W, H, dev, n_lights, res = 3840, 2160, "cuda:0", 4, 64
    ctx = ctx_parity
    g = scene.make_uniforms(W, H, pos=(0.0, 2.0, 6.0), yaw=0.0, pitch=-0.25, light_count=n_lights)
    sc = scene.make_scene(64, device=dev)
    gbd = scene.s_scene(g, W, H, sc, device=dev)
    lights_t = scene.make_lights(n_lights, device=dev)
    cubes = scene.shadow_cubes(sc, lights_t, res)
    gb = engine.GBufferResources(ctx, W, H, with_position=False)
    gb.upload(depth=gbd.depth, normal=gbd.normal, albedo=gbd.albedo, mro=gbd.mro)
    lights = engine.PointLightCollection(ctx, n_lights, res, True)
    lnp = lights_t.cpu().numpy()
    for i in range(n_lights):
        lights.setLight(i, engine.PointLight(lnp[i, 0:3], lnp[i, 4:7]))
    lights.updateResource()
    lights.setShadowMaps(cubes.cpu().numpy())
    env, lut = golden_env(), golden_lut()
    pre, pre_size, irr = ibl_standins(env)
    F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
    ibl = engine.IBLResources(ctx.image_from_numpy(env, F32, env.shape[1], env.shape[0]), ctx.image_from_numpy(pre, F32, pre_size[0], pre_size[1], 5),
                              ctx.image_from_numpy(irr, F32, irr.shape[1], irr.shape[0]), ctx.image_from_numpy(lut, _capi.FORMAT_R8G8B8A8_UNORM, lut.shape[1], lut.shape[0]))
    ssr = engine.ScreenSpaceReflection(ctx, W, H)
    dp = engine.DeferredPass(ctx, W, H, F32)
    ssr.captureReflection(g, gb, ibl, lights)
    ssr.convolveReflectionBuffer()
    dp.draw(g, gb, ibl, lights, ssr, _capi.SHADE_SKIP_TONEMAP)
    torch.cuda.synchronize()
    # the oracle on the same inputs; its position attachment = its own reconstructPosition of every covered pixel, which the
    # parity build reproduces bit for bit (test_mode_d_positions_from_depth): read back from the engine's scratch via mode P
    d = gbd.numpy()
    og = oracle.GlobalUniforms.from_buffer_copy(bytes(g))
    pos = oracle.reconstruct_positions(og, W, H, d["depth"], d["normal"])
    fr = oracle.Frame(og, W, H, pos, d["depth"], d["normal"], d["albedo"], d["mro"], env, pre, pre_size, 5, irr, lut, lnp, cubes.cpu().numpy(), res)
    refl, hit, _ = oracle.ssr_capture(fr)
    chain = oracle.glossy_convolve(refl)
    ao = oracle.ssao(fr)
    col = oracle.deferred_shade(fr, chain, 5, oracle.SKIP_TONEMAP, ao)
    got_refl = ssr.getReflectionBuffer().image.level_numpy(0).view(np.uint16).reshape(H, W, 4)
    assert np.array_equal(half_to_float(got_refl)[..., 3] != 0, hit != 0)
    assert 0.05 < hit.mean() < 0.6
    assert np.array_equal(dp.aoCounts.tensor.view(H, W).cpu().numpy(), ao)
    got_col = dp.colorTarget.tensor.view(torch.float32).view(H, W, 4).cpu().numpy()
    err = np.abs(got_col - col)
    ok = close(got_col, col).all(axis=-1)
    assert ok.mean() >= 0.9995, "colour off the bar on %.4f %% of pixels, max abs error %.3g" % (100 * (1 - ok.mean()), err.max())
    assert float(np.percentile(err, 99.9)) <= 1e-3, "99.9th percentile of the absolute colour error: %.3g" % np.percentile(err, 99.9)
