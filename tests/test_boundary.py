"""The half of the C-ABI boundary that needs no Vulkan device to be exercised: the external-memory / external-semaphore import
entry points on inputs a host can get wrong (status code + message, nothing leaked, the context stays usable), an import of
memory another API exported as a POSIX fd where the driver allows it, and two frames in flight on two CUDA streams
(`althea_sync.cuda_stream`), the reference's MAX_FRAMES_IN_FLIGHT = 2 (Include/Althea/Library.h:3)."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import FrameData, GpuFrame

pytestmark = pytest.mark.gpu

ERR_INVALID_ARGUMENT, ERR_UNSUPPORTED = None, None


def _codes():
    import re
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "althea_cuda.h")).read()
    return {m.group(1): int(m.group(2)) for m in re.finditer(r"(ALTHEA_(?:OK|ERR_[A-Z_]+))\s*=\s*(-?\d+)", text)}


def _err(ctx):
    return (ctx._lib.althea_cuda_last_error(ctx._ptr) or b"").decode()


def test_import_entry_points_reject_what_a_host_can_get_wrong(ctx_fast):
    from althea_b200 import _capi
    ctx, lib, codes = ctx_fast, ctx_fast._lib, _codes()
    import torch
    free0 = torch.cuda.mem_get_info()[0]
    h = C.c_uint64(0)
    F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
    # a file descriptor that is not one
    assert lib.althea_cuda_import_image(ctx._ptr, -1, 1 << 20, 0, F32, 64, 64, 1, 1, 0, 0, C.byref(h)) == codes["ALTHEA_ERR_INVALID_ARGUMENT"]
    assert "fd" in _err(ctx)
    assert lib.althea_cuda_import_buffer(ctx._ptr, -1, 4096, 0, C.byref(h)) == codes["ALTHEA_ERR_INVALID_ARGUMENT"]
    assert lib.althea_cuda_import_semaphore(ctx._ptr, -1, 1, C.byref(h)) == codes["ALTHEA_ERR_INVALID_ARGUMENT"]
    # an empty allocation, an offset outside it, an empty extent
    r, w = os.pipe()  # a real descriptor that is no exported allocation
    try:
        assert lib.althea_cuda_import_image(ctx._ptr, r, 0, 0, F32, 64, 64, 1, 1, 0, 0, C.byref(h)) == codes["ALTHEA_ERR_INVALID_ARGUMENT"]
        assert lib.althea_cuda_import_image(ctx._ptr, r, 4096, 4096, F32, 64, 64, 1, 1, 0, 0, C.byref(h)) == codes["ALTHEA_ERR_INVALID_ARGUMENT"]
        assert lib.althea_cuda_import_image(ctx._ptr, r, 1 << 20, 0, F32, 0, 64, 1, 1, 0, 0, C.byref(h)) == codes["ALTHEA_ERR_INVALID_ARGUMENT"]
        assert lib.althea_cuda_import_buffer(ctx._ptr, r, 0, 0, C.byref(h)) == codes["ALTHEA_ERR_INVALID_ARGUMENT"]
        assert lib.althea_cuda_import_buffer(ctx._ptr, r, 64, 128, C.byref(h)) == codes["ALTHEA_ERR_INVALID_ARGUMENT"]
        # what every image of the reference is (VK_IMAGE_TILING_OPTIMAL, Src/Image.cpp:23): refused with the way out in the message
        rc = lib.althea_cuda_import_image(ctx._ptr, r, 1 << 20, 0, F32, 64, 64, 1, 1, _capi.IMAGE_OPTIMAL_TILING, 0, C.byref(h))
        assert rc == codes["ALTHEA_ERR_UNSUPPORTED"] and "LINEAR" in _err(ctx) and "INTEGRATION.md" in _err(ctx)
        # a descriptor the driver cannot import (a pipe): the CUDA error comes back as a status, nothing is registered
        for call in (lambda: lib.althea_cuda_import_image(ctx._ptr, os.dup(r), 1 << 20, 0, F32, 64, 64, 1, 1, 0, 0, C.byref(h)),
                     lambda: lib.althea_cuda_import_buffer(ctx._ptr, os.dup(r), 1 << 20, 0, C.byref(h)),
                     lambda: lib.althea_cuda_import_semaphore(ctx._ptr, os.dup(r), 1, C.byref(h)),
                     lambda: lib.althea_cuda_import_semaphore(ctx._ptr, os.dup(r), 0, C.byref(h))):
            h.value = 0
            rc = call()
            assert rc == codes["ALTHEA_ERR_CUDA"] and h.value == 0 and "cudaImportExternal" in _err(ctx), (rc, _err(ctx))
    finally:
        os.close(r)
        os.close(w)
    # waiting on / signalling a handle that is not a semaphore
    fd = FrameData("scene", 64, 36, n_lights=0)
    gf = GpuFrame(ctx, fd)
    s = _capi.Sync()
    s.wait_sem = gf.gbuffer.depth.handle  # an image handle
    rc = lib.althea_cuda_glossy_convolve(ctx._ptr, gf.ssr.getReflectionBuffer().image.handle, C.byref(s))
    assert rc == codes["ALTHEA_ERR_BAD_HANDLE"] and "semaphore" in _err(ctx)
    s = _capi.Sync()
    s.signal_sem = 0xDEAD
    rc = lib.althea_cuda_glossy_convolve(ctx._ptr, gf.ssr.getReflectionBuffer().image.handle, C.byref(s))
    assert rc == codes["ALTHEA_ERR_BAD_HANDLE"]
    # the context is still usable and nothing leaked on the device
    gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
    gf.ssr.convolveReflectionBuffer()
    torch.cuda.synchronize()
    del gf
    torch.cuda.empty_cache()
    assert torch.cuda.mem_get_info()[0] >= free0 - (64 << 20)


def test_import_of_memory_exported_as_a_posix_fd(ctx_fast):
    """No Vulkan device here, but the CUDA driver itself can export an allocation as a POSIX file descriptor (cuMemCreate with a
    shareable handle type). cudaImportExternalMemory's opaque-fd type is specified for Vulkan / D3D exports; the driver may or
    may not accept its own export. Either way the boundary must behave: a usable image whose bytes are the allocation's, or a
    clean ALTHEA_ERR_CUDA."""
    try:
        from cuda.bindings import driver as cu
    except Exception:
        try:
            from cuda import cuda as cu
        except Exception:
            pytest.skip("cuda-python is not importable")
    import torch

    from althea_b200 import _capi
    ctx, lib, codes = ctx_fast, ctx_fast._lib, _codes()
    torch.cuda.synchronize()
    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    err, gran = cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM)
    if err != cu.CUresult.CUDA_SUCCESS:
        pytest.skip("cuMemGetAllocationGranularity: %s" % err)
    size = ((64 * 64 * 16 + gran - 1) // gran) * gran
    err, handle = cu.cuMemCreate(size, prop, 0)
    if err != cu.CUresult.CUDA_SUCCESS:
        pytest.skip("cuMemCreate with a POSIX fd handle type: %s" % err)
    try:
        err, fd = cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)
        if err != cu.CUresult.CUDA_SUCCESS:
            pytest.skip("cuMemExportToShareableHandle: %s" % err)
        h = C.c_uint64(0)
        rc = lib.althea_cuda_import_image(ctx._ptr, int(fd), size, 0, _capi.FORMAT_R32G32B32A32_SFLOAT, 64, 64, 1, 1, 0, 0, C.byref(h))
        if rc == 0:
            assert h.value != 0
            assert lib.althea_cuda_release(ctx._ptr, h.value) == 0
        else:
            os.close(int(fd))
            assert rc == codes["ALTHEA_ERR_CUDA"] and _err(ctx)
    finally:
        cu.cuMemRelease(handle)


@pytest.mark.parametrize("which", ["parity", "fast"])
def test_two_frames_in_flight_on_two_streams(request, which):
    """Two different frames submitted back to back on two CUDA streams of one context (althea_sync.cuda_stream), several
    times over, with nothing between them but stream order: each must come out exactly as it does alone. The engine-internal
    scratch (positions reconstructed from depth, SSAO records, plane records, padded depth, AO counts) is kept per stream."""
    import torch

    from althea_b200 import _capi, engine
    ctx = request.getfixturevalue("ctx_parity" if which == "parity" else "ctx_fast")
    frames = []
    for view, (W, H) in enumerate(((384, 216), (320, 200))):
        fd = FrameData("scene", W, H, n_lights=2, shadow_res=32, cam=dict(pos=(0.3 * view, 2.0, 6.0 - view), yaw=0.1 * view, pitch=-0.25))
        gf = GpuFrame(ctx, fd, out_format=_capi.FORMAT_R16G16B16A16_SFLOAT)
        gf.gbuffer_d = engine.GBufferResources(ctx, W, H, with_position=False)  # mode D: the position scratch is in play
        gf.gbuffer_d.upload(depth=fd.depth, normal=fd.normal, albedo=fd.albedo, mro=fd.mro)
        gf.deferred.aoCounts = None if view == 0 else gf.deferred.aoCounts       # frame 0 uses the ctx's AO scratch
        frames.append(gf)

    def submit(gf, stream):
        fd = gf.fd
        gf.ssr.captureReflection(fd.uniforms, gf.gbuffer_d, gf.ibl, gf.lights, stream)
        gf.ssr.convolveReflectionBuffer(stream)
        gf.deferred.draw(fd.uniforms, gf.gbuffer_d, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP, stream)

    torch.cuda.synchronize()
    alone = []
    for gf in frames:
        submit(gf, 0)
        torch.cuda.synchronize()
        alone.append((gf.deferred.colorTarget.tensor.clone(), gf.ssr.getReflectionBuffer().image.tensor.clone()))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(6):
        for gf in frames:
            gf.deferred.colorTarget.tensor.zero_()
            gf.ssr.getReflectionBuffer().image.tensor.zero_()
        torch.cuda.synchronize()
        order = (0, 1) if rep % 2 == 0 else (1, 0)
        for k in order:
            submit(frames[k], (s1, s2)[k].cuda_stream)
        torch.cuda.synchronize()
        for k, gf in enumerate(frames):
            assert torch.equal(gf.deferred.colorTarget.tensor, alone[k][0]), (rep, k)
            assert torch.equal(gf.ssr.getReflectionBuffer().image.tensor, alone[k][1]), (rep, k)
