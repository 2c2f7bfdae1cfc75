"""Host-side mirror (Python flavour) of the reference's interface for the deferred screen-space path.

Class and method names follow the reference's C++ API so a test reads like a use of the engine:
    GBufferResources          Include/Althea/DeferredRendering.h:38-102
    ReflectionBuffer          Include/Althea/ReflectionBuffer.h:27-73      (convolveReflectionBuffer)
    ScreenSpaceReflection     Include/Althea/ScreenSpaceReflection.h:26-64 (captureReflection, convolveReflectionBuffer)
    IBLResources / ImageBasedLighting.createResources   Include/Althea/ImageBasedLighting.h:17-54
    PointLight / PointLightCollection                   Include/Althea/PointLight.h:31-153
    GlobalUniforms            Include/Althea/GlobalUniforms.h:15-31
Every method forwards to the C ABI (include/althea_cuda.h); errors surface as AltheaError (a RuntimeError, the
reference throws std::runtime_error). PyTorch is used for device memory and streams only. The C++ twin of this file is
althea_b200/host/Althea/*.h.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import GlobalUniforms  # noqa: F401  (re-exported)


class AltheaError(RuntimeError):
    pass


def _mip_dim(d: int, k: int) -> int:
    return max(1, d >> k)


def image_bytes(fmt: int, w: int, h: int, mips: int = 1, layers: int = 1) -> int:
    return int(_capi.load().althea_cuda_image_bytes(fmt, w, h, mips, layers))


class Image:
    """A registered image handle plus (optionally) the torch tensor that owns its memory."""

    def __init__(self, ctx: "Context", handle: int, fmt: int, w: int, h: int, mips: int, layers: int, tensor=None):
        self.ctx, self.handle, self.format, self.w, self.h, self.mips, self.layers, self.tensor = ctx, handle, fmt, w, h, mips, layers, tensor

    @property
    def nbytes(self) -> int:
        return image_bytes(self.format, self.w, self.h, self.mips, self.layers)

    def level_offset_bytes(self, level: int, layer: int = 0) -> int:
        bpp = _capi.BYTES_PER_TEXEL[self.format]
        chain = sum(_mip_dim(self.w, k) * _mip_dim(self.h, k) * bpp for k in range(self.mips))
        return layer * chain + sum(_mip_dim(self.w, k) * _mip_dim(self.h, k) * bpp for k in range(level))

    def level_numpy(self, level: int = 0, layer: int = 0) -> np.ndarray:
        """Device -> host copy of one mip level as raw texels (uint8 view reshaped to (h, w, bytes_per_texel))."""
        if self.tensor is None:
            raise AltheaError("image has no backing tensor")
        bpp = _capi.BYTES_PER_TEXEL[self.format]
        off = self.level_offset_bytes(level, layer)
        lw, lh = _mip_dim(self.w, level), _mip_dim(self.h, level)
        raw = self.tensor.view(-1)[off:off + lw * lh * bpp].cpu().numpy()
        return raw.reshape(lh, lw, bpp)

    def release(self):
        if self.handle:
            self.ctx._check(self.ctx._lib.althea_cuda_release(self.ctx._ptr, self.handle))
            self.handle = 0


class Buffer:
    def __init__(self, ctx, handle, tensor=None):
        self.ctx, self.handle, self.tensor = ctx, handle, tensor


class Context:
    """althea_cuda_ctx: one per (process, CUDA device)."""

    def __init__(self, device: int = 0, parity_math: bool = False):
        self._lib = _capi.load()
        p = C.c_void_p()
        rc = self._lib.althea_cuda_create(C.byref(p), device, None)
        if rc != 0:
            raise AltheaError("althea_cuda_create failed (%d): %s" % (rc, (self._lib.althea_cuda_last_error(None) or b"").decode()))
        self._ptr = p
        self.device = device
        self.flags = 0
        if parity_math:
            self.set_flags(_capi.CTX_PARITY_MATH)

    def _check(self, rc: int):
        if rc != 0:
            raise AltheaError("althea_cuda error %d: %s" % (rc, (self._lib.althea_cuda_last_error(self._ptr) or b"").decode()))

    def close(self):
        if getattr(self, "_ptr", None):
            self._lib.althea_cuda_destroy(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_flags(self, flags: int):
        self._check(self._lib.althea_cuda_set_flags(self._ptr, flags))
        self.flags = flags

    def set_scissor_rows(self, y0: int = 0, y1: int = 0):
        """Restricts the per-frame stages to what rows [y0, y1) of the final image need (row-band multi-GPU mode);
        (0, 0) restores the whole frame."""
        self._check(self._lib.althea_cuda_set_scissor_rows(self._ptr, y0, y1))

    # ---- resources ----
    def wrap_tensor(self, tensor, fmt: int, w: int, h: int, mips: int = 1, layers: int = 1) -> Image:
        """Registers a contiguous CUDA torch tensor as a linear image (althea_cuda_wrap_linear_image)."""
        need = image_bytes(fmt, w, h, mips, layers)
        have = tensor.numel() * tensor.element_size()
        if not tensor.is_cuda or not tensor.is_contiguous() or have < need:
            raise AltheaError("wrap_tensor needs a contiguous CUDA tensor of >= %d bytes (got %d)" % (need, have))
        hnd = C.c_uint64()
        self._check(self._lib.althea_cuda_wrap_linear_image(self._ptr, C.c_void_p(tensor.data_ptr()), 0, fmt, w, h, mips, layers, C.byref(hnd)))
        return Image(self, hnd.value, fmt, w, h, mips, layers, tensor.view(-1).view(dtype=_torch().uint8))

    def new_image(self, fmt: int, w: int, h: int, mips: int = 1, layers: int = 1, zero: bool = True) -> Image:
        torch = _torch()
        n = image_bytes(fmt, w, h, mips, layers)
        t = (torch.zeros if zero else torch.empty)(n, dtype=torch.uint8, device="cuda:%d" % self.device)
        return self.wrap_tensor(t, fmt, w, h, mips, layers)

    def image_from_numpy(self, arr: np.ndarray, fmt: int, w: int, h: int, mips: int = 1, layers: int = 1) -> Image:
        torch = _torch()
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        t = torch.from_numpy(raw.copy()).to("cuda:%d" % self.device)
        return self.wrap_tensor(t, fmt, w, h, mips, layers)

    def wrap_buffer(self, tensor) -> Buffer:
        hnd = C.c_uint64()
        self._check(self._lib.althea_cuda_wrap_buffer(self._ptr, C.c_void_p(tensor.data_ptr()), tensor.numel() * tensor.element_size(), C.byref(hnd)))
        return Buffer(self, hnd.value, tensor)

    def create_image(self, fmt, w, h, mips=1, layers=1) -> Image:
        """Context-owned device image (althea_cuda_create_image); use upload/download for host transfers."""
        hnd = C.c_uint64()
        self._check(self._lib.althea_cuda_create_image(self._ptr, fmt, w, h, mips, layers, C.byref(hnd)))
        return Image(self, hnd.value, fmt, w, h, mips, layers, None)

    def create_buffer(self, size: int) -> Buffer:
        hnd = C.c_uint64()
        self._check(self._lib.althea_cuda_create_buffer(self._ptr, size, C.byref(hnd)))
        return Buffer(self, hnd.value, None)

    def upload(self, res, host_ptr: int, nbytes: int, stream: int = 0):
        self._check(self._lib.althea_cuda_upload(self._ptr, res.handle, C.c_void_p(host_ptr), nbytes, C.c_void_p(stream)))

    def download(self, res, host_ptr: int, nbytes: int, stream: int = 0):
        self._check(self._lib.althea_cuda_download(self._ptr, res.handle, C.c_void_p(host_ptr), nbytes, C.c_void_p(stream)))

    def synchronize(self, stream: int = 0):
        """Waits for the work issued on `stream`. stream None/0 => torch's current stream, the one the engine's calls run on by
        default (_sync_ref); the ctx's private stream carries nothing unless a host without torch drives the C ABI."""
        self._check(self._lib.althea_cuda_synchronize(self._ptr, C.c_void_p(stream if stream else current_stream_ptr(self.device))))

    # ---- instrumentation ----
    def enable_timing(self, on: bool = True):
        self._check(self._lib.althea_cuda_enable_timing(self._ptr, int(on)))

    def reset_timings(self):
        self._check(self._lib.althea_cuda_reset_timings(self._ptr))

    def timings(self) -> dict:
        cap = 32
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        cnt = (C.c_uint32 * cap)()
        n = self._lib.althea_cuda_get_timings(self._ptr, names, ms, cnt, cap)
        if n < 0:
            self._check(n)
        return {names[i].decode(): {"total_ms": float(ms[i]), "launches": int(cnt[i])} for i in range(n)}

    def launch_count(self) -> int:
        return int(self._lib.althea_cuda_launch_count(self._ptr))

    def ssao_gathers(self) -> int:
        """Proxy records gathered by the last SSAO launch made with CTX_SSAO_COUNT_TAPS set (diagnostics)."""
        out = C.c_uint64(0)
        self._check(self._lib.althea_cuda_diag_ssao_gathers(self._ptr, C.byref(out)))
        return int(out.value)

    def ssao_exact_fallbacks(self) -> int:
        """Taps of that launch that were re-evaluated from the fp32 texels (diagnostics, ray-depth proxy)."""
        out = C.c_uint64(0)
        self._check(self._lib.althea_cuda_diag_ssao_exact_fallbacks(self._ptr, C.byref(out)))
        return int(out.value)

    def ssao_cull_counts(self) -> dict:
        """All counters of that launch: position records gathered, exact re-evaluations, plane-record lookups of the coarse
        sign test, and the march steps it could not drop (diagnostics)."""
        out = (C.c_uint64 * 4)()
        self._check(self._lib.althea_cuda_diag_ssao_cull(self._ptr, out))
        return {"records": int(out[0]), "exact_taps": int(out[1]), "plane_lookups": int(out[2]), "exact_steps": int(out[3])}

    def gather_ceiling(self, w: int, h: int, radius: int, taps_per_pixel: int = 64) -> float:
        """Measured records/s of divergent 32-byte gathers within +-radius records of each 16x16 tile (diagnostics)."""
        out = C.c_double(0.0)
        self._check(self._lib.althea_cuda_diag_gather_ceiling(self._ptr, w, h, radius, taps_per_pixel, C.byref(out)))
        return float(out.value)


def band_rows(w: int, h: int, mips: int, y0: int, y1: int):
    """[(lo, hi)] per reflection mip: the rows a band [y0, y1) of a w x h frame reads or writes (althea_cuda_band_rows)."""
    lo, hi = (C.c_uint32 * mips)(), (C.c_uint32 * mips)()
    rc = _capi.load().althea_cuda_band_rows(w, h, mips, y0, y1, lo, hi)
    if rc != 0:
        raise AltheaError("althea_cuda_band_rows(%d, %d, %d, %d, %d) failed: %d" % (w, h, mips, y0, y1, rc))
    return [(int(lo[k]), int(hi[k])) for k in range(mips)]


def _torch():
    import torch
    return torch


CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: how the NULL stream is named when NULL itself means "the ctx stream"


def current_stream_ptr(device: int = 0) -> int:
    """torch's current stream as a cudaStream_t value the C ABI accepts (the NULL stream is passed as cudaStreamLegacy)."""
    return int(_torch().cuda.current_stream(device).cuda_stream) or CUDA_STREAM_LEGACY


def _sync_ref(stream, device: int = 0):
    """althea_sync for a launch. stream None/0 => torch's current stream, so engine calls are ordered with the torch copies
    that fill and read the images (the ctx's own stream is for hosts without torch)."""
    s = _capi.Sync()
    s.cuda_stream = stream if stream else current_stream_ptr(device)
    return C.byref(s), s


# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class PointLight:
    """Include/Althea/PointLight.h:31-34 (32 bytes: vec3 position, pad, vec3 emission, pad)."""
    position: Sequence[float]
    emission: Sequence[float]


class PointLightCollection:
    """Consumer side of Src/PointLight.cpp: the light SSBO + the omni shadow cube array (producer is out of scope)."""

    def __init__(self, ctx: Context, light_count: int, shadow_res: int = 256, create_shadow_map: bool = True):
        torch = _torch()
        self.ctx = ctx
        self._lights = np.zeros((light_count, 8), np.float32)
        self._dirty = True
        dev = "cuda:%d" % ctx.device
        self._buf_t = torch.zeros(max(1, light_count) * 8, dtype=torch.float32, device=dev)
        self.buffer = ctx.wrap_buffer(self._buf_t)
        self.shadow_res = shadow_res
        self.shadow_map: Optional[Image] = None
        if create_shadow_map and light_count > 0:
            t = torch.ones(light_count * 6 * shadow_res * shadow_res, dtype=torch.float32, device=dev)
            self.shadow_map = ctx.wrap_tensor(t, _capi.FORMAT_R32_SFLOAT, shadow_res, shadow_res, 1, 6 * light_count)

    def getCount(self) -> int:
        return self._lights.shape[0]

    def setLight(self, light_id: int, light: PointLight):
        self._lights[light_id, 0:3] = light.position
        self._lights[light_id, 4:7] = light.emission
        self._dirty = True

    def getLight(self, light_id: int) -> PointLight:
        return PointLight(self._lights[light_id, 0:3].copy(), self._lights[light_id, 4:7].copy())

    def updateResource(self):  # PointLight.cpp:193-203
        if self._dirty and self._lights.size:
            self._buf_t.copy_(_torch().from_numpy(self._lights.reshape(-1)))
            self._dirty = False

    def setShadowMaps(self, cubes: np.ndarray):
        """cubes: (lights, 6, res, res) float32 holding length(p - light)/1000 (ShadowMapBindless.frag:41)."""
        self.shadow_map.tensor.view(dtype=_torch().float32).copy_(_torch().from_numpy(np.ascontiguousarray(cubes, np.float32).reshape(-1)))

    def drawShadowMaps(self, models, stream: int = 0):
        """PointLightCollection::drawShadowMaps (Src/PointLight.cpp:235-282): renders every light's six cube faces from the
        models' primitives (`models`: althea_b200.model.UploadedModel instances)."""
        from . import model as _model
        if self.shadow_map is None or self.getCount() == 0:
            return
        self.updateResource()
        if not hasattr(self, "_constants"):
            self._constants = _model.point_light_constants()
        ref, keep = _sync_ref(stream, self.ctx.device)
        for k, m in enumerate(models):
            if k > 0:
                raise NotImplementedError("pass one UploadedModel holding all primitives (each call clears the cubes)")
            self.ctx._check(self.ctx._lib.althea_cuda_draw_shadow_cubes(
                self.ctx._ptr, self.buffer.handle, self.getCount(), C.byref(self._constants), m.array, m.count, self.shadow_map.handle, ref))

    @property
    def shadow_handle(self) -> int:
        return self.shadow_map.handle if self.shadow_map is not None else 0


class GBufferResources:
    """Src/DeferredRendering.cpp:37-155: depth (R32F view of D32), normal RGBA16F, albedo RGBA8, MRO RGBA8, plus the legacy
    RGBA32F position target DeferredPass.frag reads (mode P)."""

    def __init__(self, ctx: Context, width: int, height: int, with_position: bool = True):
        """with_position=False is today's GBufferResources (no legacy position attachment): the lighting pass and SSAO then
        work on positions reconstructed from depth (mode D)."""
        self.ctx, self.width, self.height = ctx, width, height
        self.depth = ctx.new_image(_capi.FORMAT_R32_SFLOAT, width, height)
        self.position = ctx.new_image(_capi.FORMAT_R32G32B32A32_SFLOAT, width, height) if with_position else None
        self.normal = ctx.new_image(_capi.FORMAT_R16G16B16A16_SFLOAT, width, height)
        self.albedo = ctx.new_image(_capi.FORMAT_R8G8B8A8_UNORM, width, height)
        self.mro = ctx.new_image(_capi.FORMAT_R8G8B8A8_UNORM, width, height)

    def upload(self, position=None, depth=None, normal=None, albedo=None, mro=None):
        """Fills the attachments from host (numpy) or device (torch) arrays of raw texel data (tests / bench inputs)."""
        torch = _torch()
        for img, src in ((self.position, position), (self.depth, depth), (self.normal, normal), (self.albedo, albedo), (self.mro, mro)):
            if src is None or img is None:
                continue
            if isinstance(src, np.ndarray):
                src = torch.from_numpy(np.ascontiguousarray(src).view(np.uint8).reshape(-1))
            else:
                src = src.contiguous().view(-1).view(dtype=torch.uint8)
            img.tensor.copy_(src, non_blocking=True)

    def struct(self) -> _capi.GBuffer:
        return _capi.GBuffer(self.depth.handle, self.position.handle if self.position is not None else 0, self.normal.handle, self.albedo.handle, self.mro.handle)


class SceneToGBufferPass:
    """SceneToGBufferPass (Src/DeferredRendering.cpp:268-330) with the Gltf.vert/.frag subpass: rasterises a model's
    primitives into the G-buffer attachments."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def draw(self, globalUniforms: GlobalUniforms, model, gBuffer: GBufferResources, stream: int = 0):
        gb = gBuffer.struct()
        ref, keep = _sync_ref(stream, self.ctx.device)
        self.ctx._check(self.ctx._lib.althea_cuda_draw_gbuffer(self.ctx._ptr, C.byref(globalUniforms), model.array, model.count, C.byref(gb), ref))


class IBLResources:
    """Include/Althea/ImageBasedLighting.h:25-43."""

    def __init__(self, environmentMap: Image, prefilteredMap: Image, irradianceMap: Image, brdfLut: Image):
        self.environmentMap, self.prefilteredMap, self.irradianceMap, self.brdfLut = environmentMap, prefilteredMap, irradianceMap, brdfLut

    def struct(self) -> _capi.IBL:
        return _capi.IBL(self.environmentMap.handle, self.prefilteredMap.handle, self.irradianceMap.handle, self.brdfLut.handle)


class ImageBasedLighting:
    """namespace ImageBasedLighting (Src/ImageBasedLighting.cpp)."""

    @staticmethod
    def generateMipMaps(ctx: Context, image: Image, stream: int = 0):
        ref, keep = _sync_ref(stream, ctx.device)
        ctx._check(ctx._lib.althea_cuda_generate_mips(ctx._ptr, image.handle, ref))

    @staticmethod
    def precomputeResources(ctx: Context, env_with_mips: Image, out_irradiance: Optional[Image], out_prefiltered: Optional[Image],
                            layout=_capi.IBL_LAYOUT_EQUIRECT, sequence=_capi.IBL_SEQ_REFERENCE_HASH, prefilter_samples=0,
                            theta_samples=0, stream: int = 0):
        """ImageBasedLighting.cpp:137-412: irradiance + GGX-prefiltered mips from an equirect env map with its mip chain."""
        desc = _capi.IblPrecomputeDesc(layout, sequence, prefilter_samples, theta_samples)
        ref, keep = _sync_ref(stream, ctx.device)
        ctx._check(ctx._lib.althea_cuda_ibl_precompute(ctx._ptr, env_with_mips.handle, C.byref(desc),
                                                       out_irradiance.handle if out_irradiance else 0,
                                                       out_prefiltered.handle if out_prefiltered else 0, ref))

    @staticmethod
    def generateBrdfLut(ctx: Context, out_lut: Image, samples: int = 1024, stream: int = 0):
        ref, keep = _sync_ref(stream, ctx.device)
        ctx._check(ctx._lib.althea_cuda_brdf_lut(ctx._ptr, samples, out_lut.handle, ref))

    @staticmethod
    def createResources(ctx: Context, env_rgba: np.ndarray, brdf_lut_rgba8: Optional[np.ndarray] = None, lut_size: int = 512,
                        stream: int = 0) -> IBLResources:
        """ImageBasedLighting.cpp:415-605 with the reference's shapes: env W x H -> irradiance W x H, prefiltered W/2 x H/2
        with 5 mips, LUT loaded if given (the reference loads brdf_lut.png) else generated."""
        env_rgba = np.ascontiguousarray(env_rgba, np.float32)
        H, W = env_rgba.shape[:2]
        mips = 1 + int(np.floor(np.log2(max(W, H))))  # Utilities.cpp:97-100
        chain = ctx.new_image(_capi.FORMAT_R32G32B32A32_SFLOAT, W, H, mips)
        torch = _torch()
        chain.tensor[: W * H * 16].copy_(torch.from_numpy(env_rgba.view(np.uint8).reshape(-1)))
        if stream and stream != current_stream_ptr(ctx.device):
            torch.cuda.current_stream(ctx.device).synchronize()  # the upload ran on torch's stream, the kernels run on `stream`
        ImageBasedLighting.generateMipMaps(ctx, chain, stream)
        irr = ctx.new_image(_capi.FORMAT_R32G32B32A32_SFLOAT, W, H)
        pre = ctx.new_image(_capi.FORMAT_R32G32B32A32_SFLOAT, W >> 1, H >> 1, 5)
        ImageBasedLighting.precomputeResources(ctx, chain, irr, pre, stream=stream)
        env = ctx.image_from_numpy(env_rgba, _capi.FORMAT_R32G32B32A32_SFLOAT, W, H)
        if brdf_lut_rgba8 is not None:
            lut = ctx.image_from_numpy(brdf_lut_rgba8, _capi.FORMAT_R8G8B8A8_UNORM, brdf_lut_rgba8.shape[1], brdf_lut_rgba8.shape[0])
        else:
            lut = ctx.new_image(_capi.FORMAT_R8G8B8A8_UNORM, lut_size, lut_size)
            ImageBasedLighting.generateBrdfLut(ctx, lut, stream=stream)
        res = IBLResources(env, pre, irr, lut)
        res._chain = chain
        if stream and stream != current_stream_ptr(ctx.device):
            ctx.synchronize(stream)  # callers read the maps on torch's stream (level_numpy, uploads of the next frame)
        return res


    @staticmethod
    def createResourcesFromContent(ctx: Context, content_dir: str, env_name: str, stream: int = 0) -> IBLResources:
        """createResources(app, commandBuffer, envMapName) as the reference runs it (ImageBasedLighting.cpp:415-605): the
        environment map is <content_dir>/HDRI_Skybox/<env_name>.hdr; if any of PrecomputedMaps/<env_name>/IrradianceMap.hdr,
        Prefiltered1..5.hdr is missing the maps are computed (on the GPU here) and written out; then ALL maps are loaded back from
        the .hdr files, so the run-time data always carries the files' RGBE quantisation, cache hit or miss. The BRDF LUT is
        PrecomputedMaps/brdf_lut.png when present (the reference ships it), generated otherwise."""
        from . import hdr_cache
        env_path = hdr_cache.environment_map_path(content_dir, env_name)
        env_rgb = hdr_cache.read_hdr(env_path)
        env_rgba = np.concatenate([env_rgb, np.ones(env_rgb.shape[:2] + (1,), np.float32)], -1)
        if not hdr_cache.cache_complete(content_dir, env_name):
            fresh = ImageBasedLighting.createResources(ctx, env_rgba, lut_size=16, stream=stream)
            ctx.synchronize(stream)  # the stream the precompute was issued on, before the maps are read back and cached on disk
            hdr_cache.save_precomputed_maps(content_dir, env_name, fresh.irradianceMap, fresh.prefilteredMap)
            del fresh
        irr, pre = hdr_cache.load_precomputed_maps(content_dir, env_name)
        F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
        env = ctx.image_from_numpy(env_rgba, F32, env_rgba.shape[1], env_rgba.shape[0])
        irr_img = ctx.image_from_numpy(irr, F32, irr.shape[1], irr.shape[0])
        pre_img = ctx.image_from_numpy(np.concatenate([p.reshape(-1) for p in pre]), F32, pre[0].shape[1], pre[0].shape[0], len(pre))
        lut_path = os.path.join(content_dir, "PrecomputedMaps", "brdf_lut.png")
        if os.path.exists(lut_path):
            from PIL import Image as PILImage
            lut_np = np.ascontiguousarray(np.asarray(PILImage.open(lut_path).convert("RGBA"), np.uint8))
            lut = ctx.image_from_numpy(lut_np, _capi.FORMAT_R8G8B8A8_UNORM, lut_np.shape[1], lut_np.shape[0])
        else:
            lut = ctx.new_image(_capi.FORMAT_R8G8B8A8_UNORM, 512, 512)
            ImageBasedLighting.generateBrdfLut(ctx, lut, stream=stream)
        return IBLResources(env, pre_img, irr_img, lut)


class ReflectionBuffer:
    """Src/ReflectionBuffer.cpp: one RGBA16F image, 5 mips (:28-39)."""

    MIP_COUNT = 5

    def __init__(self, ctx: Context, width: int, height: int, mip_count: int = MIP_COUNT):
        self.ctx = ctx
        self.image = ctx.new_image(_capi.FORMAT_R16G16B16A16_SFLOAT, width, height, mip_count)

    def convolveReflectionBuffer(self, stream: int = 0):
        ref, keep = _sync_ref(stream, self.ctx.device)
        self.ctx._check(self.ctx._lib.althea_cuda_glossy_convolve(self.ctx._ptr, self.image.handle, ref))


class ScreenSpaceReflection:
    """Src/ScreenSpaceReflection.cpp."""

    def __init__(self, ctx: Context, width: int, height: int):
        self.ctx = ctx
        self._reflectionBuffer = ReflectionBuffer(ctx, width, height)

    def getReflectionBuffer(self) -> ReflectionBuffer:
        return self._reflectionBuffer

    def captureReflection(self, globalUniforms: GlobalUniforms, gBuffer: GBufferResources, ibl: IBLResources,
                          lights: Optional[PointLightCollection], stream: int = 0):
        gb, ib = gBuffer.struct(), ibl.struct()
        ref, keep = _sync_ref(stream, self.ctx.device)
        self.ctx._check(self.ctx._lib.althea_cuda_ssr_capture(
            self.ctx._ptr, C.byref(globalUniforms), C.byref(gb), C.byref(ib), lights.buffer.handle if lights else 0,
            lights.shadow_handle if lights else 0, self._reflectionBuffer.image.handle, ref))

    def convolveReflectionBuffer(self, stream: int = 0):
        self._reflectionBuffer.convolveReflectionBuffer(stream)


class DeferredPass:
    """The app-owned deferred lighting pass (Shaders/DeferredPass.vert/.frag) as one call."""

    def __init__(self, ctx: Context, width: int, height: int, out_format: int = _capi.FORMAT_R16G16B16A16_SFLOAT):
        self.ctx = ctx
        self.colorTarget = ctx.new_image(out_format, width, height)
        self.aoCounts = ctx.new_image(_capi.FORMAT_R8_UINT, width, height)

    def draw(self, globalUniforms: GlobalUniforms, gBuffer: GBufferResources, ibl: IBLResources, lights: Optional[PointLightCollection],
             ssr: ScreenSpaceReflection, flags: int = _capi.SHADE_SKIP_TONEMAP, stream: int = 0):
        gb, ib = gBuffer.struct(), ibl.struct()
        ref, keep = _sync_ref(stream, self.ctx.device)
        self.ctx._check(self.ctx._lib.althea_cuda_deferred_shade(
            self.ctx._ptr, C.byref(globalUniforms), C.byref(gb), C.byref(ib), lights.buffer.handle if lights else 0,
            lights.shadow_handle if lights else 0, ssr.getReflectionBuffer().image.handle, self.colorTarget.handle,
            self.aoCounts.handle if self.aoCounts is not None else 0, flags, ref))  # no image: the ctx's per-stream AO scratch
