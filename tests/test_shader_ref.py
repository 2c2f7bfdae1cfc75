"""The restatement and the CUDA path against the REFERENCE'S OWN SHADER TEXT, executed.

oracle/_ref/libshader_ref.so runs SSR.vert/.frag, DeferredPass.vert/.frag (with SSAO.glsl and PBR/PBRMaterial.glsl),
SSRGlossyConvolve.comp, Misc/ReconstructPosition.glsl and the two IBL_Precompute integrators on the CPU: oracle/glsl2cpp.py rewrites the GLSL where it lies under
/root/reference/Shaders (declarations, literals, constructor braces; never an expression) and oracle/glsl_compat.h supplies the
language. What it yields on two seeded frames is committed as tests/golden/shader_ref.npz (make_shader_golden.py), so these tests
run wherever the tree goes; where the reference is mounted the library is rebuilt and checked against the fixture as well.

Bars. Integer work is exact: SSAO counts and the four convolve levels are bit for bit the shader's. Where a threshold sits on
floating-point noise the two evaluations may part: the view direction is interpolated from three vertices by the shader stage
and evaluated in closed form by the restatement (1 ulp apart), and the restatement fuses `dRaw * (far - near) - far` as GPU
compilers do while g++ evaluates the text unfused; SSR hit masks are held to north_star's 0.1 %."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, FrameData, GpuFrame, golden_env, half_to_float

FRAMES = {"scene": ("scene", 128, 72), "rand": ("rand", 96, 54)}
MASK_BAR = 1e-3


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "shader_ref.npz"))


@pytest.fixture(scope="module")
def frames():
    return {n: FrameData(k, W, H, n_lights=4, shadow_res=64) for n, (k, W, H) in FRAMES.items()}


def _close(a, b, tol):
    return np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))


@pytest.mark.parametrize("name", ["scene", "rand"])
def test_restatement_equals_the_executed_shader_text(oracle, golden, frames, name):
    fr = frames[name].oracle_frame()
    # computeSSAO: every count
    ao = oracle.ssao(fr)
    assert (golden[name + "_ao"][golden[name + "_ao"] < 255] > 0).mean() > 0.02, "the frame must produce occlusion"
    assert np.array_equal(ao, golden[name + "_ao"])
    # SSR.frag: hit mask to the bar, colour of the common hits to half storage
    refl, hit, _ = oracle.ssr_capture(fr)
    assert golden[name + "_hit"].mean() > 0.05
    assert np.mean(hit != golden[name + "_hit"]) <= MASK_BAR
    same = hit == golden[name + "_hit"]
    ok = _close(half_to_float(refl), half_to_float(golden[name + "_refl"]), 2e-3).all(-1)
    assert ok[same].mean() >= 0.998
    # ... and bit for bit once the two things GLSL and the rasteriser leave open are taken as the restatement takes them: one fma in
    # ReconstructPosition.glsl:8 (as GPU compilers contract it), the view-direction varying evaluated at the pixel (not interpolated)
    assert np.array_equal(refl, golden[name + "_refl_aligned"])
    # SSRGlossyConvolve.comp x 4 on the shader's own mip 0: every half
    assert np.array_equal(oracle.glossy_convolve(golden[name + "_refl"]), golden[name + "_chain"])
    # DeferredPass.frag main (its own computeSSAO included), linear and tone-mapped
    for key, flags in (("_color_linear", oracle.SKIP_TONEMAP), ("_color_tonemapped", 0)):
        if name + key not in golden:
            continue
        got = oracle.deferred_shade(fr, golden[name + "_chain"], flags=flags)
        want = golden[name + key]
        # (the interpolated direction is an ulp off the closed form; the equirect lookups amplify that near the poles)
        assert _close(got, want, 2e-4).all(), "max abs %.3g" % np.abs(got - want).max()
        assert _close(got, want, 1e-5).mean() >= 0.999
    # with the varying evaluated at the pixel: identical but for the few texels where the restatement's fp32 filter weights at a pixel
    # centre leave ~1e-7 of a neighbour in the reflection fetch (rule A3 is applied to the G-buffer fetches only): two ulp at most
    got = oracle.deferred_shade(fr, golden[name + "_chain"], flags=oracle.SKIP_TONEMAP)
    want = golden[name + "_color_linear_aligned"]
    assert _close(got, want, 1e-6).all() and (got == want).all(-1).mean() >= 0.99
    if name == "scene":
        assert np.array_equal(got, want)


def test_ibl_integrators_equal_the_executed_shader_text(oracle, golden):
    """GenIrradianceMap.comp (300 x 150 samples) and PreFilterEnvMap.comp (10 000 hash-RNG samples, the five roughnesses) at probe
    texels of the reference's own equirect layout: every float of the restatement is the shader's."""
    env = golden_env()
    H, W = env.shape[:2]
    chain, mips = oracle.env_mip_chain(env)
    assert np.array_equal(oracle.ibl_irradiance(chain, W, H, mips, W, H, golden["irr_texels"]), golden["irr"])
    for i, r in enumerate((0.0, 0.25, 0.5, 0.75, 1.0)):
        got = oracle.ibl_prefilter(chain, W, H, mips, W >> (i + 1), H >> (i + 1), r, golden["pre%d_texels" % i])
        assert np.array_equal(got, golden["pre%d" % i]), r
        assert np.isfinite(got).all() and got[:, :3].max() > 0.05


def test_producer_stages_equal_the_executed_shader_text(oracle, golden):
    """G-buffer pass and omni shadow pass: the fixture was drawn by the restatement's rasteriser with its
    vertex and fragment STAGES replaced by the reference's shader text (Gltf/Gltf.vert + .frag with InstanceData.glsl's fetchMaterial,
    ShadowMapBindless.vert + .frag) through althea_oracle_raster.cpp's stage hooks. The restatement's own stages must draw the same
    attachments bit for bit: textures with sRGB / mirror / clamp / nearest-mip samplers, a normal map, alpha cutout, blending."""
    from producer_scene import gbuffer_case, shadow_case
    proj, view, prims, W, H = gbuffer_case()
    got = oracle.draw_gbuffer(proj, view, prims, W, H)
    assert 0.3 < (golden["gbuffer_tri"] != 0xFFFFFFFF).mean() < 1.0
    for k, v in got.items():
        assert np.array_equal(v, golden["gbuffer_" + k]), k
    lights, projection, views, sc, res = shadow_case()
    cubes = oracle.draw_shadow_cubes(lights, projection, views, sc, res)
    assert (golden["shadow_cubes"] < 1).mean() > 0.99
    assert np.array_equal(cubes, golden["shadow_cubes"])


def test_view_direction_of_the_vertex_stage(golden, frames):
    """DeferredPass.vert's varying at the pixel centres == the closed form every kernel evaluates."""
    fd = frames["scene"]
    g = fd.oracle_frame().g
    inv_p = np.array(g.inverseProjection, np.float64).reshape(4, 4).T  # column-major storage
    inv_v = np.array(g.inverseView, np.float64).reshape(4, 4).T
    ys, xs = np.mgrid[0:fd.H, 0:fd.W]
    ndc = np.stack([2 * (xs + 0.5) / fd.W - 1, 2 * (ys + 0.5) / fd.H - 1, np.zeros_like(xs, float), np.ones_like(xs, float)], -1)
    want = (ndc @ inv_p.T)[..., :3] @ inv_v[:3, :3].T
    assert np.abs(golden["scene_dir"] - want).max() <= 1e-6 * np.abs(want).max()


def test_live_library_reproduces_the_fixture(oracle, golden, frames):
    from oracle import shader_ref as S
    if not S.available():
        pytest.skip("oracle/_ref/libshader_ref.so needs /root/reference to be built")
    fr = frames["rand"].oracle_frame()
    assert np.array_equal(S.ssao(fr), golden["rand_ao"])
    refl, hit = S.ssr_capture(fr)
    assert np.array_equal(refl, golden["rand_refl"]) and np.array_equal(hit, golden["rand_hit"])
    assert np.array_equal(S.glossy_convolve(refl), golden["rand_chain"])
    assert np.array_equal(S.deferred_shade(fr, golden["rand_chain"], flags=oracle.SKIP_TONEMAP), golden["rand_color_linear"])
    # a frame the fixture does not hold: other view, odd size
    fr2 = FrameData("rand", 75, 41, n_lights=2, shadow_res=32, view=5).oracle_frame()
    assert np.array_equal(S.ssao(fr2), oracle.ssao(fr2))
    r2, h2, _ = oracle.ssr_capture(fr2)
    assert np.array_equal(S.glossy_convolve(r2), oracle.glossy_convolve(r2))
    assert np.mean(S.ssr_capture(fr2)[1] != h2) <= 2 * MASK_BAR
    env = golden_env()
    chain, mips = oracle.env_mip_chain(env)
    assert np.array_equal(S.ibl_irradiance(chain, 512, 256, mips, 512, 256, golden["irr_texels"][:3]), golden["irr"][:3])
    assert np.array_equal(S.ibl_prefilter(chain, 512, 256, mips, 128, 64, 0.25, golden["pre1_texels"][:3]), golden["pre1"][:3])
    # the producers, drawn again with the stages hooked to the shader text
    from producer_scene import gbuffer_case, shadow_case
    S.set_raster_stage_hooks(True)
    try:
        proj, view, prims, W, H = gbuffer_case()
        for k, v in oracle.draw_gbuffer(proj, view, prims, W, H).items():
            assert np.array_equal(v, golden["gbuffer_" + k]), k
        lights, projection, views, sc, res = shadow_case()
        assert np.array_equal(oracle.draw_shadow_cubes(lights, projection, views, sc, res), golden["shadow_cubes"])
    finally:
        S.set_raster_stage_hooks(False)
    # Misc/ReconstructPosition.glsl: the restatement fuses dRaw (far - near) - far (one FFMA, as GPU compilers emit it); the text
    # run by g++ is unfused, and the cancellation shows: agreement to ~1e-3 relative only, which is why the choice is pinned
    p0, p1 = oracle.reconstruct_position(fr.g, 0.3, 0.6, 0.9991), S.reconstruct_position(fr.g, 0.3, 0.6, 0.9991)
    assert np.abs(p0 - p1).max() <= 5e-3 * np.abs(p0).max()


GLSL_SNIPPET = """#version 450
layout(location=0) in vec2 uv;
layout(location=0) out vec4 outColor;
layout(set=0, binding=1) uniform sampler2D maps[];
layout(push_constant) uniform Push { uint handle; float scale; } push;
uvec2 seed;
float next() { seed += uvec2(1); return float(seed.x); }
void split(in vec3 v, out vec2 a, inout float b) { a = v.xy; b += v.z; }
void main() {
  vec3 order = vec3(next(), next(), next());
  vec4 c = vec4(0.5, 0.25, 0.125, 1.0);
  vec4 d = vec4(9.0);
  d.rgb = c.rgb;
  vec2 a; float b = 1.0;
  split(order, a, b);
  outColor = vec4(order.x * 100.0 + order.y * 10.0 + order.z, d.a, a.y, b + push.scale * 2.0e0);
}
"""


def test_translator_and_language_library_on_a_synthetic_shader():
    """No reference needed: glsl2cpp.py + glsl_compat.h on GLSL written here. Pins the rules the rewrite rests on: arguments are
    evaluated left to right, `a.rgb = b.rgb` writes three components, `out` parameters are references, a GLSL literal is a float."""
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "sh"))
        open(os.path.join(td, "sh", "t.frag"), "w").write(GLSL_SNIPPET)
        inc = os.path.join(td, "t.inc")
        subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "glsl2cpp.py"), "--root", os.path.join(td, "sh"), "t.frag", inc],
                              stdout=subprocess.DEVNULL)
        text = open(inc).read()
        assert "vec3{next(), next(), next()}" in text and "0.5f" in text and "2.0e0f" in text and "vec2& a, float& b" in text
        assert "sampler2D* maps;" in text and "layout" not in text and "struct Push" in text
        src = os.path.join(td, "t.cpp")
        open(src, "w").write('#include "glsl_compat.h"\n#include <cstdio>\nnamespace glsl { struct T : ShaderBase {\n#include "%s"\n}; }\n'
                             'int main() { glsl::T t; t.push.scale = 0.5f; t.main(); std::printf("%%g %%g %%g %%g", t.outColor.x, t.outColor.y, t.outColor.z, t.outColor.w); }\n' % inc)
        exe = os.path.join(td, "t")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-w", "-I", os.path.join(ROOT, "oracle"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
        # next() returns 1, 2, 3 in GLSL's order -> 123; d.a keeps its 9; a.y = order.y = 2; b = 1 + 3 + 0.5 * 2
        assert [float(v) for v in out] == [123.0, 9.0, 2.0, 5.0]


def _ctx(request, which):
    return request.getfixturevalue("ctx_parity" if which == "parity" else "ctx_fast")


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["parity", "fast"])
@pytest.mark.parametrize("name", ["scene", "rand"])
def test_cuda_path_against_the_executed_shader_text(request, golden, frames, which, name):
    """The kernels through the C ABI against what the reference's GLSL yields (not against the restatement): AO counts bit for bit
    in the parity build, masks within 0.1 % otherwise, convolve levels and colour to north_star's bar."""
    from althea_b200 import _capi
    ctx = _ctx(request, which)
    fd = frames[name]
    gf = GpuFrame(ctx, fd)
    gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
    gf.ssr.convolveReflectionBuffer()
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
    ao, want_ao = gf.ao_counts(), golden[name + "_ao"]
    assert np.array_equal(ao == 255, want_ao == 255)
    if which == "parity":
        assert np.array_equal(ao, want_ao)
    else:
        assert np.mean(ao != want_ao) <= MASK_BAR
    hit = half_to_float(gf.reflection_level(0))[..., 3] != 0
    assert np.mean(hit != (golden[name + "_hit"] != 0)) <= MASK_BAR
    # SSRGlossyConvolve.comp on the shader's own mip 0: the four levels it wrote (parity build: every half; fast: half an ulp)
    import torch
    chain = golden[name + "_chain"]
    refl_img = gf.ssr.getReflectionBuffer().image
    refl_img.tensor.copy_(torch.from_numpy(chain.view(np.uint8).reshape(-1)))
    gf.ssr.convolveReflectionBuffer()
    got_chain = gf.reflection_chain().reshape(-1)
    if which == "parity":
        assert np.mean(got_chain != chain) < 1e-3
    assert _close(half_to_float(got_chain), half_to_float(chain), 2.0 ** -10 + 1e-6).all()
    # DeferredPass.frag on the shader's own reflection chain and AO counts (a flipped SSR hit would otherwise be blurred into its
    # neighbourhood by the mips): colour to north_star's bar on every pixel
    refl_img.tensor.copy_(torch.from_numpy(chain.view(np.uint8).reshape(-1)))
    gf.deferred.aoCounts.tensor.copy_(torch.from_numpy(want_ao.reshape(-1)))
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP | _capi.SHADE_AO_FROM_IMAGE)
    got, want = gf.color(), golden[name + "_color_linear"]
    assert _close(got, want, 1e-3).all(), "max rel err %g" % float(np.max(np.abs(got - want) / np.maximum(1, np.abs(want))))


@pytest.mark.gpu
def test_cuda_producers_against_the_executed_shader_text(ctx_fast, golden):
    """The visibility-buffer rasteriser through the C ABI against attachments drawn with the reference's own vertex / fragment text."""
    import torch

    import test_raster as TR
    from althea_b200 import engine, scene
    from producer_scene import gbuffer_case, shadow_case
    proj, view, prims, W, H = gbuffer_case()
    g = scene.make_uniforms(W, H, pos=(0.3, 0.8, 3.5), yaw=0.1, pitch=-0.2)
    got, _ = TR._gpu_gbuffer(ctx_fast, g, prims, W, H)
    want = {k[len("gbuffer_"):]: golden[k] for k in golden.files if k.startswith("gbuffer_")}
    # (the cutout's alpha test compares a filtered value with the cutoff: contraction may move a handful of pixels)
    TR._compare_gbuffer(got, want, exact=False, max_bad=int(0.001 * W * H))
