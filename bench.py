#!/usr/bin/env python
"""bench.py — deferred+SSAO+SSR Mpixel/s at 4K (BASELINE.json metric), plus IBL prefilter ms.

A "step" = the hot path (ssr_capture -> glossy_convolve x4 -> ssao -> deferred_shade) over this rank's batch of
VIEWS_PER_GPU synthetic 3840x2160 camera views (BASELINE configs[2] frame: 16 point lights with 256^2 omni shadow cubes,
SSR + SSAO + glossy mips; sharded by view as in configs[4]: 8 views per GPU, one process per GPU, no data-path
collective => weak scaling). value = pixels all ranks shaded / time. Inputs are resident in HBM for `value`; `e2e`
repeats the measurement through the C ABI with the G-buffers in pinned HOST memory (H2D of every view's attachments and
D2H of its colour target inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160
N_LIGHTS = 16
SHADOW_RES = 256
ENV_W, ENV_H = 4096, 2048
VIEWS_PER_GPU = 8
METRIC = "deferred+SSAO+SSR Mpixel/s at 4K"
UNIT = "Mpixel/s"

# algorithmic bytes per pixel (SURVEY.md 8d / DESIGN.md): each buffer read once, written once
BYTES_PER_PX = {
    "ssr_capture": 28.0,      # depth 4 + normal 8 + albedo 4 + MRO 4 + write RGBA16F 8
    "reconstruct_position": 28.0,  # mode D: depth 4 + normal 8 read, RGBA32F position scratch 16 written
    "ssr_depth_pad": 8.0,     # depth 4 read + padded copy 4 written per pixel (engine scratch)
    "glossy_convolve": 13.28,  # read 8*(1+1/4+1/16+1/64) + write 8*(1/4+...+1/256)
    "ssao": 25.0,             # position 16 + normal 8 + write count 1 (the proxy records are engine scratch, not algorithmic)
    "ssao_cull": 25.0,        # the same stage with the coarse sign test in front (round 2): same algorithmic bytes
    "ssao_quads": 52.0,       # position 16 read + 32-byte proxy record + 4-byte reciprocal depth written per pixel (engine scratch)
    "ssao_planes": 4.5,       # reciprocal depths 4 read, plane records (three levels) ~0.33 written per pixel (engine scratch)
    "deferred_shade": 51.66,  # position 16 + normal 8 + albedo 4 + MRO 4 + AO count 1 + reflection mips (upper bound) 10.66 + write RGBA16F 8
}
BOUND_NAMES = {"L1 data pipe (LSU wavefronts)": "l1_gather", "instruction issue": "issue", "DRAM": "hbm", "L2 (lts throughput)": "l2",
               "FMA pipe": "fp32", "XU pipe": "xu"}


def workload_config(V: int, mode: str):
    """The `config` object of the headline workload: shared by this arm and `--impl reference` (same workload, same words)."""
    px_frame = W4K * H4K
    return {"workload": "C3/C5 stand-in: %d views/GPU of a 3840x2160 S-scene deferred+SSAO+SSR+glossy frame, 16 point lights + 256^2 omni "
                        "shadow cubes, view-sharded (no collective)" % V,
            "views_per_gpu": V, "resolution": [W4K, H4K], "lights": N_LIGHTS, "shadow_res": SHADOW_RES,
            "l2_policy": "inputs (%.1f GB/step/GPU) exceed the 126 MB L2" % (V * px_frame * (36 if mode == "P" else 20) / 1e9),
            "math": "fast build (FFMA); parity build checked in tests",
            "gbuffer": "mode %s: %s" % (mode, "depth + normal + albedo + MRO attachments (20 B/px), positions reconstructed from depth inside the frame"
                                        if mode != "P" else "legacy position attachment as input (36 B/px)")}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_evt, self.proc = gpu_index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self._stop_evt.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank: int):
    """Pins this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers of the e2e leg are
    allocated: with 8 ranks each moving ~3 GB per step over PCIe, buffers on the far socket halve the copy rate. Best effort."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        return None
    return None


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this workload (profiles/ncu_traffic.json,
    written by tools/ncu_summary.py traffic ...); None when no capture is on record."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        rec = json.load(open(p)).get(kernel)
        if rec:
            return rec
    return None


def measure_bands8k(ctx, rank, local_rank, world, device, ibl, lights, steps, warmup, mode_p=False):
    """BASELINE configs[3]: one synthetic 7680x4320 S-rand G-buffer, rendered in row bands over the ranks. Every step = NCCL
    broadcast of the producing rank's G-buffer + band render (reflection halo recomputed locally) + NCCL all-gather of the
    RGBA16F bands. Two measurements: `latency`, the three phases back to back for one frame, and the headline `ms_per_step`
    with two frames in flight (althea_b200.bands.BandPipeline: the collectives of the neighbouring frames run under the band
    render, as the reference engine's MAX_FRAMES_IN_FLIGHT = 2 would have them). Returns the record on rank 0."""
    import torch
    import torch.distributed as dist

    from althea_b200 import _capi, bands, engine, scene
    W8, H8 = 7680, 4320
    g = scene.make_uniforms(W8, H8, pos=(0.0, 0.0, 0.0), yaw=0.0, pitch=0.0, light_count=N_LIGHTS)
    stream = engine.current_stream_ptr(local_rank)
    slots = []
    gbd = scene.s_rand(g, W8, H8, device=device) if rank == 0 else None
    for _ in range(2):
        gb = engine.GBufferResources(ctx, W8, H8, with_position=mode_p)
        if rank == 0:  # the producing rank; the others receive it by broadcast every step
            gb.upload(position=gbd.position, depth=gbd.depth, normal=gbd.normal, albedo=gbd.albedo, mro=gbd.mro)
        slots.append((gb, bands.BandedFrame(ctx, W8, H8)))
    del gbd
    bf0 = slots[0][1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def render(slot):
        slot[1].render(g, slot[0], ibl, lights, _capi.SHADE_SKIP_TONEMAP, stream)

    pipe = bands.BandPipeline(slots, lambda sl: sl[1].broadcast_gbuffer(sl[0], src=0, async_op=True), render, lambda sl: sl[1].gather(async_op=True))
    pipe.run(max(warmup, 1))
    launches0 = ctx.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.run(steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    launches = ctx.launch_count() - launches0
    # ---- one frame, phases back to back (no overlap): latency and the split between collectives and render
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    lat = torch.zeros(3, dtype=torch.float64, device=device)
    reps = 2
    for _ in range(reps):
        barrier()
        ev[0].record()
        bf0.broadcast_gbuffer(slots[0][0], src=0)
        ev[1].record()
        render(slots[0])
        ev[2].record()
        bf0.gather()
        ev[3].record()
        torch.cuda.synchronize()
        lat += torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])], dtype=torch.float64, device=device)
    lat /= reps
    # ---- the same frame on ONE GPU (rank 0, whole frame): the denominator of the strong-scaling efficiency
    single = torch.zeros(1, dtype=torch.float64, device=device)
    if rank == 0 and world > 1:
        whole = bands.BandedFrame(ctx, W8, H8, rank=0, world=1)
        whole.render(g, slots[0][0], ibl, lights, _capi.SHADE_SKIP_TONEMAP, stream)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            whole.render(g, slots[0][0], ibl, lights, _capi.SHADE_SKIP_TONEMAP, stream)
        ev[1].record()
        torch.cuda.synchronize()
        single[0] = ev[0].elapsed_time(ev[1]) / reps
        del whole
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lat, op=dist.ReduceOp.MAX)
        dist.all_reduce(single, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    ms_step = float(t.item()) / steps
    bcast, rend, gath = (float(v) for v in lat.tolist())
    one = float(single.item()) if world > 1 else rend
    gbytes = W8 * H8 * (36 if mode_p else 20)
    lo, hi = bf0.halo_rows()
    exchanged = bands.exchanges_halo(W8, H8, bf0.ssr.getReflectionBuffer().image.mips, world)
    halo_how = ("exchanged point to point with the neighbouring ranks while SSAO runs: %d rows of mip 0 around a band of %d" % ((hi - lo) - (bf0.y1 - bf0.y0), bf0.band)
                if exchanged else "recomputed locally, no collective inside the frame: %d rows around a band of %d" % ((hi - lo) - (bf0.y1 - bf0.y0), bf0.band))
    return {"metric": "deferred+SSAO+SSR Mpixel/s, one 8K frame in row bands", "value": W8 * H8 / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": max(warmup, 1), "ms_per_step": ms_step, "scaling": "strong", "frames_in_flight": 2,
            "workload": "configs[3]: one 7680x4320 S-rand G-buffer (random depth/normal/albedo/MRO, 5 %% empty), 16 lights; rows split over %d rank(s); every "
                        "step = NCCL broadcast of the G-buffer (%.2f GB, mode %s) + band render (reflection halo %s) + NCCL all-gather of "
                        "the RGBA16F bands (%.0f MB)" % (world, gbytes / 1e9, "P" if mode_p else "D", halo_how, W8 * H8 * 8 / 1e6),
            "band_rows": bf0.band, "gpu_launches": int(launches) * world,
            "latency": {"broadcast_ms": bcast, "render_ms": rend, "allgather_ms": gath, "frame_ms": bcast + rend + gath,
                        "note": "one frame, phases back to back, max over ranks, device events"},
            "exposed_comm_frac": max(0.0, ms_step - rend) / ms_step,
            "single_gpu_ms": one, "strong_scaling_efficiency": one / (world * ms_step),
            "roofline_chain": {"bytes_per_px": 92.0, "achieved_GBps": 92.0 * W8 * H8 / (ms_step * 1e-3) / 1e9}}


def main_bands8k(args, ctx, rank, local_rank, world, device):
    import torch.distributed as dist
    ibl, lights, _, _ = build_rank_inputs(ctx, rank, 0, device, quick_ibl=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    rec = measure_bands8k(ctx, rank, local_rank, world, device, ibl, lights, args.steps, args.warmup, mode_p=args.gbuffer_mode == "P")
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        hbm_peak, peak_src, _ = peaks()
        rec.update({"higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "clocks": clocks,
                    "config": {"workload": rec.pop("workload"), "band_rows": rec.pop("band_rows"), "l2_policy": "inputs (0.66 GB) exceed the 126 MB L2"}})
        rec["roofline_chain"]["hbm_frac"] = rec["roofline_chain"]["achieved_GBps"] / (hbm_peak * world)
        rec["roofline_chain"]["peak_source"] = peak_src
        print(json.dumps(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def build_rank_inputs(ctx, rank: int, views: int, device: str, quick_ibl: bool = False, with_position: bool = True):
    """All device-resident inputs of this rank: IBL set (built with our own precompute kernels), lights + shadow cubes, and
    `views` G-buffers. Returns the objects plus the measured IBL precompute timings."""
    import torch

    from althea_b200 import _capi, engine, scene
    F32 = _capi.FORMAT_R32G32B32A32_SFLOAT
    # --- IBL: procedural 4096x2048 HDR env -> mips -> GGX prefilter at the reference's shape (5 mips, 10000 samples, hash RNG)
    env_t = scene.procedural_env(ENV_W, ENV_H, device)
    mips = 13
    chain = ctx.new_image(F32, ENV_W, ENV_H, mips)
    chain.tensor[: ENV_W * ENV_H * 16].copy_(env_t.view(-1).view(torch.uint8))
    engine.ImageBasedLighting.generateMipMaps(ctx, chain)
    pre = ctx.new_image(F32, ENV_W >> 1, ENV_H >> 1, 5)
    irr_small = ctx.new_image(F32, 512, 256)
    ctx.enable_timing(True)
    ctx.reset_timings()
    if quick_ibl:  # tools/stage_bench.py: per-frame kernel tuning does not need converged IBL maps
        engine.ImageBasedLighting.precomputeResources(ctx, chain, irr_small, pre, prefilter_samples=64, theta_samples=16)
    else:
        engine.ImageBasedLighting.precomputeResources(ctx, chain, irr_small, pre)
    ctx.synchronize()
    ibl_t = {"reference_shape": {k: v["total_ms"] for k, v in ctx.timings().items()}}
    if not quick_ibl and views > 0 and rank == 0:
        # BASELINE configs[1]: 32^2 diffuse irradiance cube + 512^2 6-mip GGX prefilter cube (Hammersley, 10000 samples) + 512^2 BRDF LUT
        prec = ctx.new_image(F32, 512, 512, 6, 6)
        irrc = ctx.new_image(F32, 32, 32, 1, 6)
        lutc = ctx.new_image(_capi.FORMAT_R8G8B8A8_UNORM, 512, 512)
        for timed in (False, True):
            ctx.reset_timings()
            engine.ImageBasedLighting.precomputeResources(ctx, chain, irrc, prec, layout=_capi.IBL_LAYOUT_CUBE, sequence=_capi.IBL_SEQ_HAMMERSLEY)
            engine.ImageBasedLighting.generateBrdfLut(ctx, lutc, 1024)
            ctx.synchronize()
        ibl_t["config1_cube"] = {k: v["total_ms"] for k, v in ctx.timings().items()}
        del prec, irrc, lutc
    ctx.enable_timing(False)
    ctx.reset_timings()
    # the reference keeps a full-size irradiance image (ImageBasedLighting.cpp:534-568): same footprint, upsampled content
    irr_s = irr_small.tensor.view(torch.float32).view(1, 256, 512, 4).permute(0, 3, 1, 2)
    irr_full = torch.nn.functional.interpolate(irr_s, size=(ENV_H, ENV_W), mode="bilinear", align_corners=False).permute(0, 2, 3, 1).contiguous()
    irr = ctx.wrap_tensor(irr_full.view(-1), F32, ENV_W, ENV_H)
    env = ctx.wrap_tensor(env_t.view(-1), F32, ENV_W, ENV_H)
    lut = ctx.new_image(_capi.FORMAT_R8G8B8A8_UNORM, 512, 512)
    engine.ImageBasedLighting.generateBrdfLut(ctx, lut, 1024)
    ibl = engine.IBLResources(env, pre, irr, lut)
    ibl._keep = (chain, irr_small)
    # --- scene, lights, shadow cubes (S-scene ring around the camera so every yaw sees similar content)
    sc = scene.make_ring_scene(160, device=device)
    lights_t = scene.make_lights(N_LIGHTS, device=device, ring=True)
    cubes = scene.shadow_cubes(sc, lights_t, SHADOW_RES)
    lights = engine.PointLightCollection(ctx, N_LIGHTS, SHADOW_RES, True)
    lights._buf_t.copy_(lights_t.view(-1))
    lights._lights[:] = lights_t.cpu().numpy()
    lights._dirty = False
    lights.shadow_map.tensor.view(torch.float32).copy_(cubes.view(-1))
    # --- views
    out = []
    for v in range(views):
        gv = rank * views + v  # global view id, yaw = view * 5.625 deg
        g = scene.make_uniforms(W4K, H4K, pos=(0.0, 2.0, 0.0), yaw=gv * 5.625 * 3.141592653589793 / 180.0, pitch=-0.2, light_count=N_LIGHTS)
        gbd = scene.s_scene(g, W4K, H4K, sc, device=device)
        # with_position=False: the four attachments today's GBufferResources holds (Src/DeferredRendering.cpp:42-99); the engine
        # reconstructs positions from depth inside the frame (mode D). True: the legacy RGBA32F position attachment as well (mode P)
        gb = engine.GBufferResources(ctx, W4K, H4K, with_position=with_position)
        gb.upload(position=gbd.position, depth=gbd.depth, normal=gbd.normal, albedo=gbd.albedo, mro=gbd.mro)
        ssr = engine.ScreenSpaceReflection(ctx, W4K, H4K)
        dp = engine.DeferredPass(ctx, W4K, H4K, _capi.FORMAT_R16G16B16A16_SFLOAT)
        out.append((g, gb, ssr, dp))
        del gbd
    torch.cuda.synchronize()
    return ibl, lights, out, ibl_t


def main_mesh4k(args, ctx, rank, local_rank, world, device):
    """BASELINE configs[2]'s shape from geometry: a mesh scene with 16 point lights; every step renders the lights' omni shadow
    cubes and the 4K G-buffer with the rasterising producers, then SSR, glossy mips, SSAO and deferred shading on them. (The
    reference's Sponza asset cannot travel to the GPU box; the scene is tools/raster_bench.py's procedural one, 1.06 M triangles.)
    View-sharded like views4k: each rank renders its own camera, no collective."""
    import numpy as np
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import raster_bench
    from althea_b200 import _capi, engine, scene
    from althea_b200 import model as _model
    ibl, _, _, _ = build_rank_inputs(ctx, rank, 0, device, quick_ibl=True)
    if args.model:  # a glTF asset from disk, e.g. the reference's Content/Models/Sponza/glTF/Sponza.gltf (BASELINE configs[1])
        prims = _model.load_gltf(args.model, max_texture_size=args.model_texture_size)
        cam = [float(v) for v in args.model_camera.split(",")]
    else:
        prims = raster_bench.build_scene(512)
    up = _model.UploadedModel(ctx, prims)
    lights = engine.PointLightCollection(ctx, N_LIGHTS, shadow_res=SHADOW_RES)
    prng = np.random.default_rng(1)
    if args.model:
        box = [float(v) for v in args.model_light_box.split(",")]  # xmin, xmax, ymin, ymax, zmin, zmax
        for i in range(N_LIGHTS):
            lights.setLight(i, engine.PointLight((prng.uniform(box[0], box[1]), prng.uniform(box[2], box[3]), prng.uniform(box[4], box[5])), (10.0, 10.0, 10.0)))
        g = scene.make_uniforms(W4K, H4K, pos=tuple(cam[:3]), yaw=cam[3] + 0.05 * rank, pitch=cam[4], light_count=N_LIGHTS)
    else:
        for i in range(N_LIGHTS):
            lights.setLight(i, engine.PointLight((prng.uniform(-4, 4), prng.uniform(0, 4), prng.uniform(-3, 4)), (10.0, 10.0, 10.0)))
        g = scene.make_uniforms(W4K, H4K, pos=(0.3, 0.8, 3.5), yaw=0.1 + 0.05 * rank, pitch=-0.2, light_count=N_LIGHTS)
    gb = engine.GBufferResources(ctx, W4K, H4K, with_position=False)
    gpass = engine.SceneToGBufferPass(ctx)
    ssr = engine.ScreenSpaceReflection(ctx, W4K, H4K)
    dp = engine.DeferredPass(ctx, W4K, H4K, _capi.FORMAT_R16G16B16A16_SFLOAT)
    stream = engine.current_stream_ptr(local_rank)

    def step():
        lights.drawShadowMaps([up], stream)
        gpass.draw(g, up, gb, stream)             # step 5
        run_frame((g, gb, ssr, dp), ibl, lights, stream)  # steps 6-8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    ctx.enable_timing(True)
    ctx.reset_timings()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count() - launches0
    kt = ctx.timings()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    if rank == 0:
        cov = float((gb.depth.tensor.view(torch.float32) < 1).float().mean())
        print(json.dumps({
            "metric": "shadow cubes + G-buffer + deferred+SSAO+SSR Mpixel/s at 4K, from geometry", "value": world * W4K * H4K / (ms_step * 1e-3) / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2] shape from geometry: %d-triangle %s, 16 point lights; per step: 16 x 6 x 256^2 omni shadow "
                                   "cubes + 3840x2160 G-buffer (rasterising producers) + SSR + glossy mips + SSAO + deferred shading"
                                   % (up.triangle_count, ("glTF scene " + os.path.basename(args.model)) if args.model else "procedural mesh scene"),
                       "coverage": cov, "l2_policy": "one frame's attachments (0.3 GB) exceed the 126 MB L2"},
            "gpu_launches": int(launches) * world, "clocks": clocks,
            "stages": {k: {"ms_per_step": v["total_ms"] / args.steps, "launches": v["launches"]} for k, v in kt.items()}}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_frame(view, ibl, lights, stream):
    from althea_b200 import _capi
    g, gb, ssr, dp = view
    ssr.captureReflection(g, gb, ibl, lights, stream)
    ssr.convolveReflectionBuffer(stream)
    dp.draw(g, gb, ibl, lights, ssr, _capi.SHADE_SKIP_TONEMAP, stream)


_CPU_INPUTS = {}


def cpu_baseline_sample(repeats: int = 1):
    """The reference's algorithm on the host CPU on a bounded sample of the same workload, all host threads: one 1280x720 view of
    the same S-scene with the workload's own 16 lights and 256^2 shadow cubes (per-pixel cost is the workload's). `kind` says what
    ran: "reference" = the reference's own shader text executed by oracle/_ref/libshader_ref.so, "port" = the oracle (when that
    library did not travel); `port_value` is the oracle's rate either way."""
    import numpy as np
    import torch

    from althea_b200 import scene
    from oracle import oracle as O
    # all the host's cores: undo torchrun's OMP_NUM_THREADS=1 and this rank's NUMA pinning (bind_to_gpu_numa_node)
    try:
        os.sched_setaffinity(0, range(os.cpu_count()))
    except OSError:
        pass
    O.set_num_threads(len(os.sched_getaffinity(0)))
    sw, sh = 1280, 720
    if "frame" not in _CPU_INPUTS:  # inputs are built once (the cubes on the GPU when there is one: they are inputs, not the measured work)
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        sc = scene.make_ring_scene(160, device=dev)
        g = scene.make_uniforms(sw, sh, pos=(0.0, 2.0, 0.0), yaw=0.0, pitch=-0.2, light_count=N_LIGHTS)
        gbd = scene.s_scene(g, sw, sh, sc, device=dev).numpy()
        lights_t = scene.make_lights(N_LIGHTS, ring=True, device=dev)
        cubes = scene.shadow_cubes(sc, lights_t, SHADOW_RES).cpu().numpy()
        env = scene.procedural_env(512, 256).numpy()
        chain, mips = O.env_mip_chain(env)
        l0 = 512 * 256 * 4
        pre = chain[l0:l0 + O.chain_texels(256, 128, 5) * 4].copy()
        off4 = O.chain_texels(512, 256, 4) * 4
        irr = chain[off4:off4 + 32 * 16 * 4].reshape(16, 32, 4).copy()
        lut = np.zeros((64, 64, 4), np.uint8)
        lut[..., :2] = (np.clip(O.brdf_lut(64, 64)[::-1], 0, 1) * 255 + 0.5).astype(np.uint8)
        og = O.GlobalUniforms.from_buffer_copy(bytes(g))
        _CPU_INPUTS["frame"] = O.Frame(og, sw, sh, gbd["position"], gbd["depth"], gbd["normal"], gbd["albedo"], gbd["mro"], env, pre, (256, 128), 5, irr, lut,
                                       lights_t.cpu().numpy(), cubes, SHADOW_RES)
    fr = _CPU_INPUTS["frame"]
    # oracle/_ref/libshader_ref.so = the reference's own shader text run on the CPU (oracle/ref_shader_driver.cpp): when it is there (it
    # is built where /root/reference is mounted and travels with the tree) it IS the reference arm; the port is timed beside it
    from oracle import shader_ref as S
    use_text = S.available()
    S_threads = len(os.sched_getaffinity(0))

    def run_port():
        refl, hit, _ = O.ssr_capture(fr)
        ch = O.glossy_convolve(refl)
        O.deferred_shade(fr, ch, 5, O.SKIP_TONEMAP, None)  # SSAO computed inside, as the shader does

    def run_text():
        refl, hit = S.ssr_capture(fr)          # SSR.vert + SSR.frag main
        ch = S.glossy_convolve(refl)           # SSRGlossyConvolve.comp x 4
        S.deferred_shade(fr, ch, 5, O.SKIP_TONEMAP)  # DeferredPass.vert + .frag main (computeSSAO inside)

    def best(fn):
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            fn()
            times.append(time.perf_counter() - t0)
        return min(times)

    t_port = best(run_port)
    if use_text:
        S.lib().shaderref_set_num_threads(S_threads)
        t = best(run_text)
    else:
        t = t_port
    what = "the reference's GLSL text run on the CPU (oracle/_ref/libshader_ref.so)" if use_text else "the CPU port of the oracle"
    return {"value": sw * sh / t / 1e6, "unit": UNIT, "cores": O.num_threads(), "kind": "reference" if use_text else "port",
            "sample": "%s: 1 view of the same S-scene at 1280x720 (1/9 of a 4K view), 16 lights, 256^2 shadow cubes; %d run(s), best %.2f s" % (what, repeats, t),
            "port_value": sw * sh / t_port / 1e6, "seconds": t}


def main_reference(args):
    """--impl reference: the reference's own algorithm on the host CPU. The reference's path is GLSL + Vulkan and cannot run here as
    a Vulkan program (no loader / lavapipe / glslc); what runs is its shader text, compiled as C++ into
    oracle/_ref/libshader_ref.so (kind "reference"), or the oracle port where that library is absent, all host threads. A step is
    a bounded sample of the workload (one 1280x720 view of the 3840x2160 views: the metric is per pixel); exactly --warmup
    untimed and --steps timed samples."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    for _ in range(args.warmup):
        cpu_baseline_sample(1)
    samples = [cpu_baseline_sample(1) for _ in range(max(1, args.steps))]
    secs = sum(s["seconds"] for s in samples) / len(samples)
    value = 1280 * 720 / secs / 1e6
    cb = dict(samples[0])
    cb["value"] = value
    cb["sample"] = cb["sample"].split(";")[0] + "; mean of %d timed samples of %.2f s" % (len(samples), secs)
    cb.pop("seconds")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(samples), "warmup": args.warmup,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.views, args.gbuffer_mode),
            "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--views", type=int, default=VIEWS_PER_GPU)
    ap.add_argument("--workload", default="views4k", choices=["views4k", "bands8k", "mesh4k"],
                    help="views4k (default, the contract benchmark): 4K views sharded by view, weak scaling. bands8k: ONE 7680x4320 "
                         "S-rand frame split in row bands, NCCL broadcast of the G-buffer + all-gather of the bands, strong scaling. mesh4k: a frame from "
                         "geometry (shadow cubes and G-buffer rasterised from a mesh scene, then the deferred chain), one camera per rank")
    ap.add_argument("--model", default=None, help="mesh4k: render this .glb / .gltf instead of the procedural scene")
    ap.add_argument("--model-texture-size", type=int, default=1024, help="mesh4k --model: cap on the texture edge")
    ap.add_argument("--model-camera", default="-8,2,0,-1.5707963,0", help="mesh4k --model: x,y,z,yaw,pitch (default: down Sponza's nave)")
    ap.add_argument("--model-light-box", default="-12,12,0.5,8,-3,3", help="mesh4k --model: xmin,xmax,ymin,ymax,zmin,zmax for the 16 lights")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-producers", action="store_true")
    ap.add_argument("--no-bands", action="store_true", help="with more than one rank: skip the bands8k (configs[3]) object of the line")
    ap.add_argument("--gbuffer-mode", default="D", choices=["D", "P"],
                    help="D (default): depth / normal / albedo / MRO, what the reference's GBufferResources holds; positions are reconstructed from "
                         "depth inside the frame. P: the legacy RGBA32F position attachment is an input too (SURVEY.md 8(c-bis) R5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout, which carries the one JSON line
        # NCCL's kernels on a high-priority stream: the row-band pipeline runs its collectives under the band render, whose kernels fill
        # every SM; at normal priority the collective's CTAs queue behind them and the overlap is lost (measured: 19.7 ms per 8K frame
        # on 8 GPUs against 16.9 ms of render)
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device(device), pg_options=opts)

    from althea_b200 import _capi, engine
    ctx = engine.Context(local_rank)
    if args.workload == "bands8k":
        return main_bands8k(args, ctx, rank, local_rank, world, device)
    if args.workload == "mesh4k":
        return main_mesh4k(args, ctx, rank, local_rank, world, device)
    mode_p = args.gbuffer_mode == "P"
    ibl, lights, views, ibl_t = build_rank_inputs(ctx, rank, args.views, device, with_position=mode_p)
    stream = engine.current_stream_ptr(local_rank)
    V = len(views)
    px_per_step = V * W4K * H4K

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        for vw in views:
            run_frame(vw, ibl, lights, stream)

    for _ in range(args.warmup):
        step()
    # ---- timed region: inputs resident in HBM; per-kernel CUDA events on the launching stream for the roofline ----
    ctx.enable_timing(True)
    ctx.reset_timings()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()  # no-op unless a profiler runs with --profile-from-start off: the ncu launch list is this region
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count() - launches0
    kt = ctx.timings()
    ctx.enable_timing(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * px_per_step / (ms_step * 1e-3) / 1e6

    # ---- what actually binds the SSAO march: divergent 32-byte gathers. Outside the timed region, rank 0: the
    # number of proxy records one frame gathers (counting instantiation of the kernel) and the device's measured rate for the
    # same access pattern (one 256-bit load per lane, random positions in a window around the lane's tile).
    gather = None
    if rank == 0:
        ctx.set_flags(_capi.CTX_SSAO_COUNT_TAPS)
        run_frame(views[0], ibl, lights, stream)
        torch.cuda.synchronize()
        counts = ctx.ssao_cull_counts()
        ctx.set_flags(_capi.CTX_SSAO_COUNT_TAPS | _capi.CTX_SSAO_NO_CULL)
        run_frame(views[0], ibl, lights, stream)
        torch.cuda.synchronize()
        records_march = ctx.ssao_gathers()
        ctx.set_flags(0)
        ceilings = {str(r): ctx.gather_ceiling(W4K, H4K, r, 64) for r in (32, 96)}
        gather = {"records_per_frame": counts["records"], "plane_lookups": counts["plane_lookups"], "exact_steps": counts["exact_steps"],
                  "records_march": records_march, "ceiling_records_per_s": ceilings}

    # ---- the rasterising producers upstream of the path, informational, outside the timed region:
    # G-buffer pass at 4K and the 16 x 6 shadow-cube faces of a 1.06 M triangle mesh scene (tools/raster_bench.py)
    producers, indoor = None, None
    if rank == 0 and not args.no_producers:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import numpy as np
        import raster_bench
        from althea_b200 import model as _model
        up = _model.UploadedModel(ctx, raster_bench.build_scene(512))
        gpass = engine.SceneToGBufferPass(ctx)
        pgb = engine.GBufferResources(ctx, W4K, H4K)
        pl = engine.PointLightCollection(ctx, N_LIGHTS, shadow_res=SHADOW_RES)
        prng = np.random.default_rng(1)
        for i in range(N_LIGHTS):
            pl.setLight(i, engine.PointLight((prng.uniform(-4, 4), prng.uniform(0, 4), prng.uniform(-3, 4)), (10.0, 10.0, 10.0)))
        pg = views[0][0]
        producers = {"scene": "procedural mesh scene, %d triangles (512x1024 textured sphere, a small sphere, six room quads)" % up.triangle_count}
        for name, fn in (("draw_gbuffer_4k_ms", lambda: gpass.draw(pg, up, pgb, stream)), ("draw_shadow_cubes_16x6x256_ms", lambda: pl.drawShadowMaps([up], stream))):
            fn()
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(3):
                fn()
            p1.record()
            torch.cuda.synchronize()
            producers[name] = p0.elapsed_time(p1) / 3
        # ... and the deferred chain on THAT G-buffer: an indoor, sky-free view (every SSAO ray stays on screen, every SSR ray has
        # something to hit), beside the S-scene of the headline whose 27 % sky flatters it
        ssr_i = engine.ScreenSpaceReflection(ctx, W4K, H4K)
        dp_i = engine.DeferredPass(ctx, W4K, H4K, _capi.FORMAT_R16G16B16A16_SFLOAT)
        iview = (pg, pgb, ssr_i, dp_i)
        run_frame(iview, ibl, pl, stream)
        torch.cuda.synchronize()
        ctx.enable_timing(True)
        ctx.reset_timings()
        for _ in range(3):
            run_frame(iview, ibl, pl, stream)
        torch.cuda.synchronize()
        it = {k: v["total_ms"] / 3 for k, v in ctx.timings().items()}
        ctx.enable_timing(False)
        ctx.reset_timings()
        covered = float((pgb.position.tensor.view(torch.float32).view(H4K, W4K, 4)[..., 3] != 0).float().mean())
        indoor = {"scene": producers["scene"] + ", camera inside the room, 16 lights with the shadow cubes rendered from the mesh; G-buffer mode P",
                  "covered_fraction": covered, "ms_per_4k_frame": sum(it.values()), "Mpixel_per_s": W4K * H4K / (sum(it.values()) * 1e-3) / 1e6,
                  "stages_ms": it}
        del up, pgb, pl, ssr_i, dp_i

    # ---- e2e: same metric through the C ABI with HOST buffers (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        F = _capi
        fmts = ([("position", F.FORMAT_R32G32B32A32_SFLOAT)] if mode_p else []) + [
            ("depth", F.FORMAT_R32_SFLOAT), ("normal", F.FORMAT_R16G16B16A16_SFLOAT), ("albedo", F.FORMAT_R8G8B8A8_UNORM), ("mro", F.FORMAT_R8G8B8A8_UNORM)]
        # one PACKED staging buffer per view (the attachments back to back, 256-byte aligned): one H2D copy per view instead of four
        sizes = [_capi.BYTES_PER_TEXEL[f] * W4K * H4K for _, f in fmts]
        offs = [0]
        for sz in sizes:
            offs.append((offs[-1] + sz + 255) & ~255)
        packed = offs[-1]
        host_in = []
        for (g, gb, ssr, dp) in views:
            hp = torch.empty(packed, dtype=torch.uint8).pin_memory()
            for (n, _), o, sz in zip(fmts, offs, sizes):
                hp[o:o + sz].copy_(getattr(gb, n).tensor.view(-1)[:sz])
            host_in.append(hp)
        host_out = [torch.empty(W4K * H4K * 8, dtype=torch.uint8).pin_memory() for _ in views]
        h2d = packed * V
        d2h = host_out[0].numel() * V
        # device side: TWO sets of G-buffer images (views into one packed device buffer each) and colour targets (double buffering,
        # what a host engine with MAX_FRAMES_IN_FLIGHT = 2 keeps, Include/Althea/Library.h:3). Uploads run on a copy-in stream, the
        # frame on the compute stream, the colour read-back on a copy-out stream; CUDA events order them, so the H2D of view i+1 and
        # the D2H of view i-1 overlap the kernels of view i. Every byte still crosses PCIe inside the timed region.
        NBUF = 2
        dgbs, ddps, dpacked = [], [], []
        for _ in range(NBUF):
            dev_t = torch.empty(packed, dtype=torch.uint8, device=device)
            dpacked.append(ctx.wrap_buffer(dev_t))
            dgb = engine.GBufferResources.__new__(engine.GBufferResources)
            dgb.ctx, dgb.width, dgb.height, dgb.position = ctx, W4K, H4K, None
            for (n, f), o, sz in zip(fmts, offs, sizes):
                setattr(dgb, n, ctx.wrap_tensor(dev_t[o:o + sz], f, W4K, H4K))
            dgbs.append(dgb)
            ddps.append(engine.DeferredPass(ctx, W4K, H4K, F.FORMAT_R16G16B16A16_SFLOAT))
        dssr = engine.ScreenSpaceReflection(ctx, W4K, H4K)
        s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
        s_cmp = torch.cuda.current_stream(device)
        ev_in = [torch.cuda.Event() for _ in range(NBUF)]    # H2D of the set finished
        ev_cmp = [torch.cuda.Event() for _ in range(NBUF)]   # frame on the set finished (set and colour target reusable / readable)
        ev_out = [torch.cuda.Event() for _ in range(NBUF)]   # D2H of the colour target finished
        counter = [0]

        def e2e_step():
            for i, (g, gb, ssr, dp) in enumerate(views):
                k = counter[0] % NBUF
                first = counter[0] < NBUF
                counter[0] += 1
                if not first:
                    s_in.wait_event(ev_cmp[k])     # the previous frame on this set has consumed it
                ctx.upload(dpacked[k], host_in[i].data_ptr(), packed, s_in.cuda_stream)
                ev_in[k].record(s_in)
                s_cmp.wait_event(ev_in[k])
                if not first:
                    s_cmp.wait_event(ev_out[k])    # the colour target of this set has been read back
                run_frame((g, dgbs[k], dssr, ddps[k]), ibl, lights, stream)
                ev_cmp[k].record(s_cmp)
                s_out.wait_event(ev_cmp[k])
                ctx.download(ddps[k].colorTarget, host_out[i].data_ptr(), host_out[i].numel(), s_out.cuda_stream)
                ev_out[k].record(s_out)
            s_cmp.wait_stream(s_out)               # the step ends when its last result is on the host

        e2e_step()
        barrier()
        k2 = max(1, min(args.steps, 3))
        e0.record()
        for _ in range(k2):
            e2e_step()
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item()) / k2
        e2e = {"value": world * px_per_step / (ms2 * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms2, "steps": k2, "numa_node": numa_node,
               "pcie_GBps_per_rank": {"h2d": h2d / (ms2 * 1e-3) / 1e9, "d2h": d2h / (ms2 * 1e-3) / 1e9},
               "how": "C ABI with pinned host buffers: ONE packed upload of the %d G-buffer attachments per view, frame, download of the RGBA16F colour " % len(fmts) +
                      "target; double-buffered, copies overlapped with kernels on separate streams"}

    # ---- BASELINE configs[3] beside the headline whenever there is more than one rank: one 8K frame in row bands, the only
    # mode with data-path collectives (NCCL broadcast + all-gather). Outside the timed region of the headline.
    bands8k = None
    if world > 1 and not args.no_bands:
        # (fail fast rather than hang: this is the only part of the bench with data-path collectives; it takes a few seconds)
        bands_done = threading.Event()

        def bands_watchdog():
            if not bands_done.wait(300.0):
                sys.stderr.write("bench.py: the bands8k measurement did not finish within 300 s on rank %d; giving up\n" % rank)
                sys.stderr.flush()
                os._exit(3)
        threading.Thread(target=bands_watchdog, daemon=True).start()
        bands8k = measure_bands8k(ctx, rank, local_rank, world, device, ibl, lights, 16, 2, mode_p=False)  # 16 frames: the pipeline's fill and drain (one exposed broadcast + all-gather per run) weigh 1-2 %, as in a running engine
        bands_done.set()

    if rank == 0:
        hbm_peak, peak_src, sm_max = peaks()
        px_frame = W4K * H4K
        stages = {}
        for name, rec in kt.items():
            n = rec["launches"]
            # glossy_convolve: 4 launches per frame, bytes accounted per frame
            per_frame_ms = rec["total_ms"] / (args.steps * V)
            gbs = BYTES_PER_PX.get(name, 0.0) * px_frame / (per_frame_ms * 1e-3) / 1e9 if per_frame_ms > 0 else 0.0
            stages[name] = {"ms_per_frame": per_frame_ms, "launches": n, "algorithmic_GBps": gbs, "hbm_frac": gbs / hbm_peak}
            tr = ncu_traffic(name)
            if tr:  # what binds the kernel, from the committed ncu --set full capture of this workload (not a flop model)
                stages[name]["ncu"] = {"binding_resource": tr["binding_resource"], "utilisation_pct": tr["utilisation_pct"], "source": tr["source"]}
        dom = max(stages, key=lambda k: stages[k]["ms_per_frame"]) if stages else None
        roofline = None
        if dom:
            s = stages[dom]
            tr = ncu_traffic(dom)
            top = tr["binding_resource"].rsplit(" ", 5)[0] if tr else None
            roofline = {"kernel": dom, "bound": BOUND_NAMES.get(top, "issue") if tr else "issue", "achieved": s["algorithmic_GBps"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": s["hbm_frac"], "traffic": tr["dram_bytes_per_launch"] if tr else None, "traffic_source": tr["source"] if tr else None,
                        "binding_resource": tr.get("binding_resource") if tr else None, "peak_source": peak_src,
                        "note": "achieved / peak / frac are the ALGORITHMIC bytes of the stage (25 B/px: position 16 + normal 8 + count 1) over its measured time "
                                "against the measured HBM copy rate, as the contract defines them; the stage is not HBM-bound: `bound` and `binding` name the "
                                "unit ncu finds busiest (instruction issue and the L1 data pipe, both ~80 %, profiles/)"}
        if roofline and dom in ("ssao", "ssao_cull") and gather:
            peak = max(gather["ceiling_records_per_s"].values())
            ach = gather["records_per_frame"] / (stages[dom]["ms_per_frame"] * 1e-3)
            roofline["binding"] = {
                "resource": "plane-record lookups in shared memory (LDS.128, ~7 wavefronts each) + divergent 32-byte gathers of position records (one L1 wavefront "
                            "per distinct 128-byte line) + instruction issue",
                "plane_lookups_per_launch": gather.get("plane_lookups"), "taps_left_to_exact_path_per_launch": gather.get("exact_steps"),
                "records_per_launch": gather["records_per_frame"], "records_per_launch_round1_march": gather.get("records_march"),
                "gather_rate_Grecords_per_s": ach / 1e9, "gather_ceiling_Grecords_per_s": peak / 1e9,
                "peak_how": "althea_cuda_diag_gather_ceiling measured in this run: one 256-bit load per lane at random positions within "
                            "+-32 / +-96 records of the lane's 16x16 tile over the 3841x2161 record grid (the larger rate is the peak)",
                "ceilings": {k: v / 1e9 for k, v in gather["ceiling_records_per_s"].items()}}
        frame_ms = ms_step / V
        chain = {"bytes_per_px": 92.0, "achieved_GBps": 92.0 * px_frame / (frame_ms * 1e-3) / 1e9, "hbm_frac": 92.0 * px_frame / (frame_ms * 1e-3) / 1e9 / hbm_peak,
                 "ms_per_4k_frame": frame_ms, "frames_per_s": 1e3 / frame_ms * world}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(V, args.gbuffer_mode),
                "gpu_launches": int(launches) * world, "clocks": clocks, "e2e": e2e, "roofline": roofline, "roofline_chain": chain, "stages": stages, "producers": producers, "indoor_view": indoor,
                "ibl_prefilter_ms": ibl_t.get("config1_cube", {}).get("ibl_prefilter"),
                "ibl_precompute_ms": ibl_t,
                "ibl_config": {"config1_cube": "BASELINE configs[1]: %dx%d equirect env -> 32^2 irradiance cube (300x150 samples) + 512^2 6-mip GGX prefilter cube "
                                               "(10000 Hammersley samples/texel) + 512^2 BRDF LUT (1024 samples); second of two runs" % (ENV_W, ENV_H),
                               "reference_shape": "the reference's own layout: 5 equirect GGX mips 2048x1024..128x64, 10000 hash-RNG samples/texel; irradiance "
                                                  "at 512x256 (the reference's 4096x2048 irradiance is 64x the work)"}}
        if bands8k is not None:
            line["bands8k"] = bands8k
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            O.build()
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline_sample(2).items() if k != "seconds"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
