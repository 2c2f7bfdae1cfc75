"""Shared test plumbing: builds one synthetic frame's inputs (numpy), and hands them to the CPU oracle and to the CUDA
engine (through the C ABI) alike."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from althea_b200 import scene  # noqa: E402
from oracle import hdrio  # noqa: E402
from oracle import oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"  # only present in the build container; never required


def golden_env() -> np.ndarray:
    """(256, 512, 4) float32: NeoclassicalInterior.hdr box-downsampled 8x (tests/golden/make_golden.py)."""
    rgb = hdrio.read_hdr(os.path.join(GOLDEN, "env_512x256.hdr"))
    return np.concatenate([rgb, np.ones(rgb.shape[:2] + (1,), np.float32)], -1)


def golden_lut() -> np.ndarray:
    """(450, 450, 4) uint8: the reference's Content/PrecomputedMaps/brdf_lut.png, palette expanded."""
    return np.load(os.path.join(GOLDEN, "brdf_lut.npz"))["lut"]


def ibl_standins(env_rgba: np.ndarray):
    """IBL maps for the per-frame tests. They are INPUTS there, so cheap stand-ins suffice: prefiltered = box mips 1..5 of
    the env map, irradiance = box mip 4. (The real precompute is tested in test_ibl_parity.py.)"""
    H, W = env_rgba.shape[:2]
    chain, mips = O.env_mip_chain(env_rgba)
    l0 = W * H * 4
    pre_w, pre_h = W >> 1, H >> 1
    pre = chain[l0:l0 + O.chain_texels(pre_w, pre_h, 5) * 4].copy()
    off4 = O.chain_texels(W, H, 4) * 4
    irr = chain[off4:off4 + (W >> 4) * (H >> 4) * 4].reshape(H >> 4, W >> 4, 4).copy()
    return pre, (pre_w, pre_h), irr


class FrameData:
    """All inputs of one frame as numpy arrays."""

    def __init__(self, kind: str, W: int, H: int, n_lights: int = 4, shadow_res: int = 64, view: int = 0, env=None, lut=None,
                 cam=None):
        self.kind, self.W, self.H = kind, W, H
        cam = cam or dict(pos=(0.0, 2.0, 6.0), yaw=0.0, pitch=-0.25)
        if kind == "rand":
            cam = dict(pos=(0.0, 0.0, 0.0), yaw=view * math_radians(5.625), pitch=0.0)
        self.uniforms = scene.make_uniforms(W, H, light_count=n_lights, **cam)
        sc = scene.make_scene(64)
        gb = scene.s_scene(self.uniforms, W, H, sc) if kind == "scene" else scene.s_rand(self.uniforms, W, H, view=view)
        d = gb.numpy()
        self.position, self.depth, self.normal, self.albedo, self.mro = d["position"], d["depth"], d["normal"], d["albedo"], d["mro"]
        self.lights = scene.make_lights(n_lights).numpy() if n_lights else None
        self.shadow_res = shadow_res
        self.shadow = scene.shadow_cubes(sc, torch.from_numpy(self.lights), shadow_res).numpy() if n_lights else None
        self.env = golden_env() if env is None else env
        self.lut = golden_lut() if lut is None else lut
        self.pre, self.pre_size, self.irr = ibl_standins(self.env)

    def oracle_frame(self) -> O.Frame:
        og = O.GlobalUniforms.from_buffer_copy(bytes(self.uniforms))
        return O.Frame(og, self.W, self.H, self.position, self.depth, self.normal, self.albedo, self.mro, self.env, self.pre,
                       self.pre_size, 5, self.irr, self.lut, self.lights, self.shadow, self.shadow_res)


def math_radians(deg: float) -> float:
    return deg * np.pi / 180.0


class GpuFrame:
    """The same inputs registered with the CUDA engine through the C ABI."""

    def __init__(self, ctx, fd: FrameData, out_format=None):
        from althea_b200 import _capi, engine
        self.ctx, self.fd = ctx, fd
        self.gbuffer = engine.GBufferResources(ctx, fd.W, fd.H)
        self.gbuffer.upload(position=fd.position, depth=fd.depth, normal=fd.normal, albedo=fd.albedo, mro=fd.mro)
        F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
        env = ctx.image_from_numpy(fd.env, F32, fd.env.shape[1], fd.env.shape[0])
        pre = ctx.image_from_numpy(fd.pre, F32, fd.pre_size[0], fd.pre_size[1], 5)
        irr = ctx.image_from_numpy(fd.irr, F32, fd.irr.shape[1], fd.irr.shape[0])
        lut = ctx.image_from_numpy(fd.lut, _capi.FORMAT_R8G8B8A8_UNORM, fd.lut.shape[1], fd.lut.shape[0])
        self.ibl = engine.IBLResources(env, pre, irr, lut)
        n = 0 if fd.lights is None else fd.lights.shape[0]
        self.lights = None
        if n:
            self.lights = engine.PointLightCollection(ctx, n, fd.shadow_res, True)
            for i in range(n):
                self.lights.setLight(i, engine.PointLight(fd.lights[i, 0:3], fd.lights[i, 4:7]))
            self.lights.updateResource()
            self.lights.setShadowMaps(fd.shadow)
        self.ssr = engine.ScreenSpaceReflection(ctx, fd.W, fd.H)
        self.deferred = engine.DeferredPass(ctx, fd.W, fd.H, out_format or _capi.FORMAT_R32G32B32A32_SFLOAT)

    # raw readbacks
    def reflection_level(self, level: int) -> np.ndarray:
        return self.ssr.getReflectionBuffer().image.level_numpy(level).view(np.uint16).reshape(
            max(1, self.fd.H >> level), max(1, self.fd.W >> level), 4)

    def reflection_chain(self) -> np.ndarray:
        img = self.ssr.getReflectionBuffer().image
        return img.tensor.cpu().numpy().view(np.uint16)

    def color(self) -> np.ndarray:
        img = self.deferred.colorTarget
        raw = img.level_numpy(0)
        from althea_b200 import _capi
        if img.format == _capi.FORMAT_R32G32B32A32_SFLOAT:
            return raw.view(np.float32).reshape(self.fd.H, self.fd.W, 4)
        return raw.view(np.float16).reshape(self.fd.H, self.fd.W, 4).astype(np.float32)

    def ao_counts(self) -> np.ndarray:
        return self.deferred.aoCounts.level_numpy(0).reshape(self.fd.H, self.fd.W)


def half_to_float(u16: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(u16).view(np.float16).astype(np.float32)


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    peak = float(max(np.abs(b).max(), 1e-12))
    return 200.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)
