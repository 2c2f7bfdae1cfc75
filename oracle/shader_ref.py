"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/_ref/libshader_ref.so: the reference's own GLSL text
(SSR.vert/.frag, DeferredPass.vert/.frag with SSAO.glsl and PBR/PBRMaterial.glsl, SSRGlossyConvolve.comp,
Misc/ReconstructPosition.glsl, the two IBL_Precompute integrators, Gltf/Gltf.vert/.frag, ShadowMapBindless.vert/.frag) rewritten by oracle/glsl2cpp.py where it lies and run on the CPU over oracle/glsl_compat.h.

It exists to pin the restatement (oracle/althea_oracle.cpp) against the text it restates. It can only be BUILT where
/root/reference is mounted (`make -C oracle ref`); the prebuilt library travels with the tree, and the vectors it produced are
committed under tests/golden/shader_ref.npz (tests/golden/make_shader_golden.py) for everywhere else.
Only tests/ and tests/golden/make_shader_golden.py import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libshader_ref.so")
_lib = None


def available(build: bool = True) -> bool:
    """True when the library exists (after trying to build it where the reference's shaders are mounted)."""
    if build and os.path.isfile("/root/reference/Shaders/SSR.frag"):
        subprocess.check_call(["make", "-C", _HERE, "_ref/libshader_ref.so"], stdout=subprocess.DEVNULL)
    return os.path.isfile(_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libshader_ref.so is not built (needs /root/reference)")
        _lib = C.CDLL(_PATH)
        _lib.shaderref_reconstruct_position.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
    return _lib


def reconstruct_position(g: O.GlobalUniforms, u, v, d_raw):
    out = np.zeros(3, np.float32)
    lib().shaderref_reconstruct_position(C.byref(g), u, v, d_raw, O._p(out))
    return out


def ssr_capture(fr: O.Frame, fused_reconstruct: bool = False, closed_form_direction: bool = False):
    """SSR.vert + SSR.frag main over the frame -> (RGBA16F reflection mip 0 as uint16, hit mask). The two switches align what GLSL and
    the rasteriser leave open with the restatement's choices: `dRaw * (far - near) - far` (ReconstructPosition.glsl:8) contracted to
    one fma, and the view-direction varying evaluated per pixel from SSR.vert's expression instead of interpolated from 3 vertices."""
    gb, ibl, li = fr._structs()
    refl = np.zeros((fr.H, fr.W, 4), np.uint16)
    hit = np.zeros((fr.H, fr.W), np.uint8)
    flags = (1 if fused_reconstruct else 0) | (2 if closed_form_direction else 0)
    lib().shaderref_ssr_capture_flags(C.byref(fr.g), C.byref(gb), C.byref(ibl), C.byref(li), O._p(refl), O._p(hit), C.c_uint32(flags))
    return refl, hit


def glossy_convolve(mip0_u16, mip_count=5):
    """SSRGlossyConvolve.comp dispatched per level as ReflectionBuffer::convolveReflectionBuffer does -> tight RGBA16F chain."""
    H, W = mip0_u16.shape[:2]
    chain = np.zeros(O.chain_texels(W, H, mip_count) * 4, np.uint16)
    chain[: W * H * 4] = np.ascontiguousarray(mip0_u16, np.uint16).ravel()
    lib().shaderref_glossy_convolve(O._p(chain), W, H, mip_count)
    return chain


def ssao(fr: O.Frame):
    """computeSSAO (SSAO.glsl) per pixel, seeded and called as DeferredPass.frag does -> occluded-ray counts (255 = empty pixel)."""
    gb, _, _ = fr._structs()
    out = np.zeros((fr.H, fr.W), np.uint8)
    lib().shaderref_ssao(C.byref(fr.g), C.byref(gb), O._p(out))
    return out


def deferred_shade(fr: O.Frame, refl_chain_u16, refl_mips=5, flags=O.SKIP_TONEMAP, closed_form_direction: bool = False):
    """DeferredPass.vert + DeferredPass.frag main (computeSSAO included) -> RGBA32F colour. closed_form_direction: the varying
    evaluated per pixel from the vertex stage's own expression (as the restatement and the kernels do) instead of interpolated."""
    flags = int(flags) | (4 if closed_form_direction else 0)
    gb, ibl, li = fr._structs()
    refl_chain_u16 = O._c(refl_chain_u16, np.uint16)
    out = np.zeros((fr.H, fr.W, 4), np.float32)
    lib().shaderref_deferred_shade(C.byref(fr.g), C.byref(gb), C.byref(ibl), C.byref(li), O._p(refl_chain_u16), refl_mips,
                                   C.c_uint32(flags), O._p(out))
    return out


def view_directions(g: O.GlobalUniforms, W, H):
    out = np.zeros((H, W, 3), np.float32)
    lib().shaderref_view_directions(C.byref(g), W, H, O._p(out))
    return out


def ibl_irradiance(chain, W, H, mips, out_w, out_h, texels):
    """IBL_Precompute/GenIrradianceMap.comp main at probe texels (x, y, 0) -> (n, 4) float32."""
    t, n = O._texels(texels)
    out = np.empty((n, 4), np.float32)
    lib().shaderref_ibl_irradiance(O._p(chain), W, H, mips, out_w, out_h, O._p(t), n, O._p(out))
    return out


def ibl_prefilter(chain, W, H, mips, out_w, out_h, roughness, texels):
    """IBL_Precompute/PreFilterEnvMap.comp main (10 000 hash-RNG samples) at probe texels of one level -> (n, 4) float32."""
    t, n = O._texels(texels)
    out = np.empty((n, 4), np.float32)
    lib().shaderref_ibl_prefilter(O._p(chain), W, H, mips, out_w, out_h, C.c_float(roughness), O._p(t), n, O._p(out))
    return out


def set_raster_stage_hooks(on: bool) -> None:
    """Makes liboracle.so's two rasterising passes (oracle.draw_gbuffer / draw_shadow_cubes) take their PROGRAMMABLE stages from the
    reference's shader text (Gltf/Gltf.vert + .frag with InstanceData.glsl's fetchMaterial; ShadowMapBindless.vert + .frag), keeping
    their own fixed-function part (coverage, depth test, interpolation, derivatives, blending). A hooked draw must equal the plain one."""
    L = O.lib()
    if not on:
        L.oracle_set_stage_hooks(None)
        return
    lib().shaderref_raster_hooks.restype = C.c_void_p
    hooks = lib().shaderref_raster_hooks(C.cast(L.oracle_sample_texture, C.c_void_p))
    L.oracle_set_stage_hooks(C.c_void_p(hooks))
