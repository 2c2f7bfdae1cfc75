"""Row-band mode (one frame over several GPUs): host logic on CPU with gloo (world_size 2), band == whole-frame equality on
one GPU through the scissor, and the NCCL path when two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch

from althea_b200 import bands, engine


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_split_rows_covers_the_frame():
    for H in (1, 7, 67, 2160, 4320, 4321):
        for world in (1, 2, 3, 4, 8):
            sp = bands.split_rows(H, world)
            assert len(sp) == world and sp[0][0] == 0 and sp[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(sp, sp[1:]))
            assert all(0 <= y1 - y0 <= bands.band_height(H, world) for y0, y1 in sp)
            assert bands.padded_rows(H, world) >= H


def test_band_rows_contain_what_the_band_reads(lib_built):
    from althea_b200 import engine
    W, H, mips = 7680, 4320, 5
    for y0, y1 in bands.split_rows(H, 8) + [(0, H), (100, 101)]:
        rows = engine.band_rows(W, H, mips, y0, y1)
        for L, (lo, hi) in enumerate(rows):
            hL = max(1, H >> L)
            assert 0 <= lo < hi <= hL
            # the deferred pass reads level L around v * hL - 0.5 for v in the band
            assert lo <= max(0, int(np.floor((y0 + 0.5) / H * hL - 0.5))) and hi >= min(hL, int(np.floor((y1 - 0.5) / H * hL - 0.5)) + 2)
        for L in range(1, mips):  # level L rows come from level L-1 rows around 2r (+- 5.18 * h_src / w_dst on vertical passes)
            (lo, hi), (slo, shi) = rows[L], rows[L - 1]
            hS, hD, wD = max(1, H >> (L - 1)), max(1, H >> L), max(1, W >> L)
            off = 5.176470588235294 * hS / wD if L & 1 else 0.0
            assert slo <= max(0, int(np.floor(lo * hS / hD - 0.5 - off)))
            assert shi >= min(hS, int(np.floor((hi - 1) * hS / hD - 0.5 + off)) + 2)
    full = engine.band_rows(W, H, mips, 0, H)
    assert full == [(0, max(1, H >> L)) for L in range(mips)]
    with pytest.raises(engine.AltheaError):
        engine.band_rows(W, H, mips, 10, 10)


def test_halo_plan_is_symmetric_and_covers_what_a_band_reads(lib_built):
    """Every row of reflection mip 0 a band's glossy mips read lies in the band itself or in exactly one receive of its plan, and
    what rank r receives from q is what q sends to r."""
    from althea_b200 import engine
    for (W, H) in ((7680, 4320), (320, 181)):
        for world in (2, 3, 8):
            plans = [bands.halo_plan(W, H, 5, world, r) for r in range(world)]
            spans = bands.split_rows(H, world)
            for r, (recv, send) in enumerate(plans):
                y0, y1 = spans[r]
                if y1 <= y0:
                    assert not recv and not send
                    continue
                lo, hi = engine.band_rows(W, H, 5, y0, y1)[0]
                got = np.zeros(H, np.int32)
                got[y0:y1] += 1
                for q, a, b in recv:
                    assert spans[q][0] <= a < b <= spans[q][1]
                    got[a:b] += 1
                    assert (r, a, b) in plans[q][1]
                assert (got[lo:hi] == 1).all() and got[:lo].sum() == 0 and got[hi:].sum() == 0
                for q, a, b in send:
                    assert (r, a, b) in plans[q][0]


def test_halo_exchange_decision_is_the_same_on_every_rank():
    """The exchange of the reflection halos is a rendezvous between neighbouring ranks: whether it happens must not depend on a
    rank's own band. The rule this replaced compared each rank's halo with its band; the outermost bands have a halo on one side
    only, so at 4 ranks of the 8K frame of BASELINE configs[3] the inner ranks exchanged and the outer ones did not (a deadlock,
    seen on 4 GPUs). `exchanges_halo` takes the largest halo of ANY band."""
    W, H = 7680, 4320
    assert [bands.exchanges_halo(W, H, 5, w) for w in (1, 2, 4, 8)] == [False, False, False, True]
    spans = bands.split_rows(H, 4)
    own = []
    for y0, y1 in spans:
        lo, hi = engine.band_rows(W, H, 5, y0, y1)[0]
        own.append((hi - lo) - (y1 - y0) >= 0.15 * (y1 - y0))  # the old, per-rank rule
    assert own == [False, True, True, False]
    for size in ((7680, 4320), (3840, 2160), (320, 181), (64, 9)):
        for world in range(1, 10):
            assert bands.exchanges_halo(size[0], size[1], 5, world) in (True, False)
    assert not bands.exchanges_halo(64, 9, 5, 1)


def _gloo_worker(rank, world, port, H, W, out):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1234)
        frame = torch.randint(0, 255, (H, W), dtype=torch.uint8, generator=g)
        mine = frame.clone() if rank == 0 else torch.zeros_like(frame)  # only the producing rank has the G-buffer
        bands.broadcast_tensors([mine], src=0)
        assert torch.equal(mine, frame)
        # "shade" the band: any per-pixel function of the WHOLE input (here: depends on a far-away row, as SSR does)
        def shade(y0, y1):
            return (mine[y0:y1].to(torch.int32) * 3 + mine.flip(0)[y0:y1].to(torch.int32)).to(torch.uint8)
        y0, y1 = bands.split_rows(H, world)[rank]
        bh = bands.band_height(H, world)
        buf = torch.zeros(bands.padded_rows(H, world) * W, dtype=torch.uint8)
        buf[y0 * W:y1 * W] = shade(y0, y1).reshape(-1)
        bands.allgather_rows(buf, W, bh)
        want = shade(0, H)
        assert torch.equal(buf[: H * W].view(H, W), want), "rank %d" % rank
        out.put((rank, True))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H", [64, 67])
def test_broadcast_and_allgather_world2_gloo(H):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, H, 48, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5)[0] for _ in range(2)) == [0, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["parity", "fast"])
def test_bands_equal_whole_frame_on_one_gpu(request, which):
    """Every band, rendered alone with poisoned surroundings, reproduces its rows of the whole-frame result bit for bit."""
    from helpers import FrameData, GpuFrame
    from althea_b200 import _capi
    ctx = request.getfixturevalue("ctx_parity" if which == "parity" else "ctx_fast")
    fd = FrameData("scene", 320, 181, n_lights=2, shadow_res=32)
    gf = GpuFrame(ctx, fd)

    def run():
        gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
        gf.ssr.convolveReflectionBuffer()
        gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
        torch.cuda.synchronize()
        return gf.deferred.colorTarget.tensor.clone().view(fd.H, -1), gf.deferred.aoCounts.tensor.clone().view(fd.H, -1)

    full_c, full_a = run()
    assert float(gf.color()[..., :3].std()) > 0.01
    for world in (2, 3, 5):
        for y0, y1 in bands.split_rows(fd.H, world):
            gf.ssr.getReflectionBuffer().image.tensor.fill_(0xFF)  # NaN halves everywhere the band does not write
            gf.deferred.colorTarget.tensor.fill_(0x7F)
            gf.deferred.aoCounts.tensor.fill_(0x7F)
            ctx.set_scissor_rows(y0, y1)
            c, a = run()
            ctx.set_scissor_rows(0, 0)
            assert torch.equal(c[y0:y1], full_c[y0:y1]), (world, y0, y1)
            assert torch.equal(a[y0:y1], full_a[y0:y1])
            assert bool((c[:y0] == 0x7F).all()) and bool((c[y1:] == 0x7F).all())  # nothing outside the band is touched
    with pytest.raises(Exception):
        ctx.set_scissor_rows(0, fd.H + 1)
        run()
    ctx.set_scissor_rows(0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["parity", "fast"])
def test_band_with_exchanged_halo_equals_whole_frame(request, which):
    """The split the multi-GPU path runs (ALTHEA_CTX_BAND_EXCHANGE_HALO + ALTHEA_SHADE_AO_ONLY / AO_FROM_IMAGE): the band's own
    reflection rows, the halo rows copied in from whoever owns them (here: from the whole-frame result, as a neighbour's band would
    hold them), SSAO on its own, then convolve and shading: the band's rows come out bit-identical to the whole frame's."""
    from helpers import FrameData, GpuFrame
    from althea_b200 import _capi, engine
    ctx = request.getfixturevalue("ctx_parity" if which == "parity" else "ctx_fast")
    fd = FrameData("scene", 320, 181, n_lights=2, shadow_res=32)
    gf = GpuFrame(ctx, fd)
    base = ctx.flags
    gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
    gf.ssr.convolveReflectionBuffer()
    gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
    torch.cuda.synchronize()
    full_c = gf.deferred.colorTarget.tensor.clone().view(fd.H, -1)
    full_r = gf.ssr.getReflectionBuffer().image.tensor.clone()
    rb = fd.W * 8
    try:
        for world in (2, 5):
            for rank, (y0, y1) in enumerate(bands.split_rows(fd.H, world)):
                refl = gf.ssr.getReflectionBuffer().image.tensor
                refl.fill_(0xFF)
                gf.deferred.colorTarget.tensor.fill_(0x7F)
                gf.deferred.aoCounts.tensor.fill_(0x7F)
                ctx.set_scissor_rows(y0, y1)
                ctx.set_flags(base | _capi.CTX_BAND_EXCHANGE_HALO)
                gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
                torch.cuda.synchronize()
                assert torch.equal(refl[y0 * rb:y1 * rb], full_r[y0 * rb:y1 * rb])
                assert bool((refl[:y0 * rb] == 0xFF).all()) and bool((refl[y1 * rb:fd.H * rb] == 0xFF).all())  # own rows only
                recv, _ = bands.halo_plan(fd.W, fd.H, 5, world, rank)
                for _, a, b in recv:
                    refl[a * rb:b * rb] = full_r[a * rb:b * rb]
                gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_AO_ONLY)
                assert bool((gf.deferred.colorTarget.tensor == 0x7F).all())  # AO_ONLY does not shade
                gf.ssr.convolveReflectionBuffer()
                gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP | _capi.SHADE_AO_FROM_IMAGE)
                torch.cuda.synchronize()
                c = gf.deferred.colorTarget.tensor.view(fd.H, -1)
                assert torch.equal(c[y0:y1], full_c[y0:y1]), (world, rank)
                ctx.set_flags(base)
                ctx.set_scissor_rows(0, 0)
    finally:
        ctx.set_flags(base)
        ctx.set_scissor_rows(0, 0)


def _nccl_worker(rank, world, port, q, rank_mode=1):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=torch.device("cuda:%d" % rank))
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from helpers import FrameData, GpuFrame
        from althea_b200 import _capi, engine
        ctx = engine.Context(rank)
        fd = FrameData("scene", 384, 216, n_lights=2, shadow_res=32)
        gf = GpuFrame(ctx, fd)
        gf.ssr.captureReflection(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights)
        gf.ssr.convolveReflectionBuffer()
        gf.deferred.draw(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, gf.ssr, _capi.SHADE_SKIP_TONEMAP)
        torch.cuda.synchronize()
        want = gf.deferred.colorTarget.tensor.clone()
        if rank != 0:  # only rank 0 "produced" the G-buffer
            for img in (gf.gbuffer.position, gf.gbuffer.depth, gf.gbuffer.normal, gf.gbuffer.albedo, gf.gbuffer.mro):
                img.tensor.zero_()
        bf = bands.BandedFrame(ctx, fd.W, fd.H, out_format=_capi.FORMAT_R32G32B32A32_SFLOAT)
        bf.broadcast_gbuffer(gf.gbuffer, src=0)
        bf.render(fd.uniforms, gf.gbuffer, gf.ibl, gf.lights, exchange_halo=(rank_mode == 1))
        bf.gather()
        torch.cuda.synchronize()
        ok = torch.equal(bf.color_rows().reshape(-1), want)
        q.put((rank, bool(ok)))
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", [1, 0])
def test_bands_over_two_gpus_nccl(exchange):
    """exchange = 1: the reflection halo comes from the neighbour (P2P inside the frame); 0: every rank recomputes it."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q, exchange)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}


def _pipeline_worker(rank, world, port, out):
    """BandPipeline over gloo: frames in flight share nothing but the slot they cycle through, collectives are issued in the same
    order on both ranks, and every assembled frame equals the single-process result."""
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        H, W, frames = 37, 16, 5
        bh = bands.band_height(H, world)
        y0, y1 = bands.split_rows(H, world)[rank]
        produced = [torch.full((H, W), 10 * k + 1, dtype=torch.uint8) + torch.arange(H, dtype=torch.uint8)[:, None] for k in range(frames)]
        slots = [dict(gb=torch.zeros(H, W, dtype=torch.uint8), out=torch.zeros(bands.padded_rows(H, world) * W, dtype=torch.uint8), frame=-1) for _ in range(2)]
        state = dict(next_replicate=0, next_render=0, results=[])

        def replicate(slot):
            k = state["next_replicate"]
            state["next_replicate"] += 1
            slot["frame"] = k
            if rank == 0:
                slot["gb"].copy_(produced[k])   # the producing rank wrote frame k's G-buffer into the slot
            return bands.broadcast_tensors([slot["gb"]], src=0, async_op=True)

        def render(slot):
            k = state["next_render"]
            state["next_render"] += 1
            assert slot["frame"] == k and torch.equal(slot["gb"], produced[k]), "frame %d rendered from a slot holding %d" % (k, slot["frame"])
            band = (slot["gb"][y0:y1].to(torch.int32) * 2 + slot["gb"].flip(0)[y0:y1].to(torch.int32)).to(torch.uint8)  # reads far rows, as SSR does
            slot["out"][y0 * W:y1 * W] = band.reshape(-1)

        def assemble(slot):
            w = bands.allgather_rows(slot["out"], W, bh, async_op=True)
            state["results"].append((slot["frame"], slot, w))
            return w

        pipe = bands.BandPipeline(slots, replicate, render, assemble, cuda=False)
        checked = []

        def check_done():
            for k, slot, w in state["results"]:
                if k not in checked:
                    w.wait()
                    want = (produced[k].to(torch.int32) * 2 + produced[k].flip(0).to(torch.int32)).to(torch.uint8)
                    assert torch.equal(slot["out"][: H * W].view(H, W), want), "frame %d on rank %d" % (k, rank)
                    checked.append(k)

        orig_render = pipe.render

        def render_and_check(slot):  # a slot's previous result must be consumed before the slot is rendered into again
            check_done()
            orig_render(slot)
        pipe.render = render_and_check
        pipe.run(frames)
        check_done()
        assert checked == list(range(frames))
        kinds = [k for k, _ in pipe.issued]
        assert pipe.issued[0] == ("replicate", 0) and kinds.count("render") == frames and kinds.count("assemble") == frames
        # frame k + 1 is replicated before frame k is rendered (two frames in flight)
        assert pipe.issued.index(("replicate", 1)) < pipe.issued.index(("render", 0))
        out.put((rank, pipe.issued))
    finally:
        dist.destroy_process_group()


def test_band_pipeline_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res[0] == res[1]  # the same issue order on both ranks


def _halo_worker(rank, world, port, out):
    """BandedFrame.exchange_halo over gloo with CPU tensors: every rank writes its own band of mip 0, the grouped send / recv plan
    brings in exactly the halo rows its glossy mips read, on every rank, without a rendezvous left open."""
    import types

    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        W, H, mips = 64, 360, 5
        y0, y1 = bands.split_rows(H, world)[rank]
        rb = W * 8
        t = torch.zeros(H * rb, dtype=torch.uint8)
        rows = torch.arange(H, dtype=torch.int32)
        pattern = ((rows * 7 + 3) % 251).to(torch.uint8)[:, None].expand(H, rb).contiguous()
        t.view(H, rb)[y0:y1] = pattern[y0:y1]
        image = types.SimpleNamespace(tensor=t, mips=mips)
        stub = types.SimpleNamespace(W=W, H=H, world=world, rank=rank, group=None,
                                     ssr=types.SimpleNamespace(getReflectionBuffer=lambda: types.SimpleNamespace(image=image)))
        stub.halo_plan = lambda: bands.halo_plan(W, H, mips, world, rank)
        for w in bands.BandedFrame.exchange_halo(stub):
            w.wait()
        lo, hi = engine.band_rows(W, H, mips, y0, y1)[0]
        got = t.view(H, rb)
        assert torch.equal(got[lo:hi], pattern[lo:hi]), "rank %d: halo rows [%d, %d)" % (rank, lo, hi)
        assert int(got[:lo].sum()) == 0 and int(got[hi:].sum()) == 0
        out.put((rank, True))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5)[0] for _ in range(world)) == list(range(world))
