"""Row-band split of ONE large frame over the GPUs of a node (BASELINE configs[3]: a 7680x4320 frame over 2/4/8 B200).

SSR rays and SSAO taps may read any screen location, so INPUTS are whole-frame on every rank (the producing rank's G-buffer
is replicated with a broadcast over NCCL/NVLink); OUTPUTS are per pixel, so each rank shades a contiguous band of rows
(`Context.set_scissor_rows`, the C ABI's althea_cuda_set_scissor_rows) and the bands are assembled with an all-gather. The
only stage with cross-band dependence is the glossy convolve: each rank recomputes the halo of reflection rows its band's
mips need (althea_cuda_band_rows) instead of exchanging halos, so a band is bit-identical to the same rows of a single-GPU
frame. One process per GPU; torch.distributed is the plumbing (backend "nccl" on the GPUs, "gloo" in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from . import _capi


def band_height(height: int, world: int) -> int:
    """Rows per band: equal bands of ceil(H / world) rows (the last ranks' bands may be shorter or empty)."""
    return (height + world - 1) // world


def split_rows(height: int, world: int) -> List[Tuple[int, int]]:
    """[(y0, y1)] per rank, contiguous, covering [0, height)."""
    bh = band_height(height, world)
    return [(min(r * bh, height), min((r + 1) * bh, height)) for r in range(world)]


def padded_rows(height: int, world: int) -> int:
    """Rows of the gather buffer: every rank contributes band_height rows so the all-gather is one equal-sized collective."""
    return band_height(height, world) * world


def _dist():
    import torch.distributed as dist
    return dist


def broadcast_tensors(tensors: Sequence, src: int = 0, group=None, async_op: bool = False):
    """Replicates the producing rank's G-buffer attachments (or the IBL tables at init) on every rank."""
    dist = _dist()
    works = [dist.broadcast(t, src=src, group=group, async_op=async_op) for t in tensors]
    return works if async_op else None


def allgather_rows(buf, row_bytes: int, band_rows: int, group=None, async_op: bool = False):
    """In-place all-gather of row bands: `buf` is a flat uint8 tensor of world * band_rows * row_bytes bytes whose rank-th
    chunk has been written by this rank; on return (or once the returned work has been waited for) every chunk is filled in."""
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = band_rows * row_bytes
    flat = buf.view(-1)[: world * n]
    mine = flat[rank * n:(rank + 1) * n]
    if flat.is_cuda:
        return dist.all_gather_into_tensor(flat, mine, group=group, async_op=async_op)  # NCCL in place: input is the rank-th slice of the output
    chunks = [flat[r * n:(r + 1) * n] for r in range(world)]
    return dist.all_gather(chunks, mine.clone(), group=group, async_op=async_op)


HALO_EXCHANGE_MIN_SHARE = 0.2  # of a band's rows


def exchanges_halo(width: int, height: int, mips: int, world: int) -> bool:
    """Whether the ranks exchange their reflection halos (True) or every rank ray-marches its own (False). THE SAME ANSWER ON EVERY
    RANK: the exchange is a rendezvous between neighbours, so it cannot depend on a rank's own band. (An earlier rule compared each
    rank's own halo with its band: the outermost bands have a halo on one side only, and at 4 ranks of an 8K frame the inner ranks
    exchanged while the outer ones did not: a deadlock.) True when the largest halo of any band is at least HALO_EXCHANGE_MIN_SHARE
    of a band: 0.30 at 8 ranks of an 8K frame (35 % more SSR when recomputed), 0.15 at 4, 0.04 at 2."""
    if world <= 1:
        return False
    from . import engine
    worst = 0
    for (y0, y1) in split_rows(height, world):
        if y1 > y0:
            lo, hi = engine.band_rows(width, height, mips, y0, y1)[0]
            worst = max(worst, (hi - lo) - (y1 - y0))
    return worst >= HALO_EXCHANGE_MIN_SHARE * band_height(height, world)


def halo_plan(width: int, height: int, mips: int, world: int, rank: int):
    """Rows of reflection mip 0 that `rank`'s glossy mips read but other ranks own, and the rows it owns that others read:
    ([(peer, y0, y1)] to receive, [(peer, y0, y1)] to send), from althea_cuda_band_rows of every rank's band."""
    from . import engine
    spans = split_rows(height, world)
    need = [engine.band_rows(width, height, mips, y0, y1)[0] if y1 > y0 else (0, 0) for (y0, y1) in spans]
    my0, my1 = spans[rank]
    recv, send = [], []
    for q, (q0, q1) in enumerate(spans):
        if q == rank or q1 <= q0 or my1 <= my0:
            continue
        lo, hi = max(need[rank][0], q0), min(need[rank][1], q1)   # rows of q's band inside my halo
        if hi > lo:
            recv.append((q, lo, hi))
        lo, hi = max(need[q][0], my0), min(need[q][1], my1)       # rows of my band inside q's halo
        if hi > lo:
            send.append((q, lo, hi))
    return recv, send


class BandPipeline:
    """Keeps `frames_in_flight` frames in flight (the reference engine keeps two, Include/Althea/Library.h:3): the replication
    of frame k + 1's G-buffer and the assembly of frame k - 1's bands run on the collective stream while frame k's band is
    being rendered, so in steady state a step costs max(render, collectives) instead of their sum. The collectives are issued
    in the same order on every rank: broadcast(0), [broadcast(k + 1), render(k), all-gather(k)] for k = 0, 1, ...

    `slots` are per-frame resources (each with its own G-buffer tensors, colour target, reflection buffer); `replicate(slot)`
    and `assemble(slot)` issue the collectives and return what must be waited for (a list of works or None), `render(slot)`
    enqueues the band's kernels on the current stream. On CUDA the collectives are issued under a side stream, so that they
    only order after the work they depend on (recorded events), not after whatever the compute stream holds."""

    def __init__(self, slots, replicate, render, assemble, cuda: bool = True):
        self.slots, self.replicate, self.render, self.assemble, self.cuda = list(slots), replicate, render, assemble, cuda
        n = len(self.slots)
        self._replicated = [None] * n   # works of the broadcast into slot i
        self._assembled = [None] * n    # works of the all-gather out of slot i
        self._rendered = [None] * n     # CUDA event: the band of the frame in slot i has been rendered
        self._side = None
        self.issued = []                # ("replicate" | "render" | "assemble", frame): the order the work was enqueued in
        if cuda:
            import torch
            self._side = torch.cuda.Stream(priority=-1)  # collectives ahead of the render kernels that fill the SMs

    @staticmethod
    def _wait(works):
        if works is None:
            return
        for w in works if isinstance(works, (list, tuple)) else [works]:
            if w is not None:
                w.wait()  # CUDA: makes the current stream wait for the collective; CPU: blocks

    def _on_side(self, fn, slot, after=None):
        if not self.cuda:
            return fn(slot)
        import torch
        if after is not None:
            self._side.wait_event(after)
        with torch.cuda.stream(self._side):
            return fn(slot)

    def _replicate(self, k):
        i = k % len(self.slots)
        # the slot's G-buffer may be overwritten once the frame that used it has been rendered
        self._replicated[i] = self._on_side(self.replicate, self.slots[i], self._rendered[i])
        self.issued.append(("replicate", k))

    def run(self, frames: int):
        """Renders `frames` frames; returns when everything has been enqueued (CUDA) / finished (CPU)."""
        n = len(self.slots)
        self._replicate(0)
        for k in range(frames):
            i = k % n
            if k + 1 < frames and n > 1:
                self._replicate(k + 1)
            self._wait(self._replicated[i])   # the G-buffer of frame k has arrived
            self._wait(self._assembled[i])    # the colour target of slot i has been read by the all-gather of frame k - n
            self.render(self.slots[i])
            self.issued.append(("render", k))
            ev = None
            if self.cuda:
                import torch
                ev = torch.cuda.Event()
                ev.record()
                self._rendered[i] = ev
            self._assembled[i] = self._on_side(self.assemble, self.slots[i], ev)
            self.issued.append(("assemble", k))
            if k + 1 < frames and n == 1:
                self._replicate(k + 1)
        for i in range(n):
            self._wait(self._assembled[i])


class BandedFrame:
    """One frame of W x H pixels rendered in row bands by the ranks of `group`.

    Usage (every rank):
        bf = BandedFrame(ctx, W, H)                 # allocates the padded colour target and this rank's reflection buffer
        bf.broadcast_gbuffer(gbuffer, src=0)        # NCCL broadcast of the 5 attachments
        bf.render(uniforms, gbuffer, ibl, lights)   # ssr_capture -> glossy_convolve -> ssao -> deferred_shade on the band
        bf.gather()                                 # NCCL all-gather of the bands; bf.color_rows() is then the full frame
    """

    def __init__(self, ctx, width: int, height: int, out_format: int = _capi.FORMAT_R16G16B16A16_SFLOAT, group=None,
                 rank: Optional[int] = None, world: Optional[int] = None):
        import torch

        from . import engine
        dist = _dist()
        self.ctx, self.W, self.H, self.group, self.format = ctx, width, height, group, out_format
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.band = band_height(height, self.world)
        self.y0, self.y1 = split_rows(height, self.world)[self.rank]
        self.row_bytes = width * _capi.BYTES_PER_TEXEL[out_format]
        self._color_t = torch.zeros(padded_rows(height, self.world) * self.row_bytes, dtype=torch.uint8, device="cuda:%d" % ctx.device)
        self.deferred = engine.DeferredPass.__new__(engine.DeferredPass)
        self.deferred.ctx = ctx
        self.deferred.colorTarget = ctx.wrap_tensor(self._color_t, out_format, width, height)
        self.deferred.aoCounts = ctx.new_image(_capi.FORMAT_R8_UINT, width, height)
        self.ssr = engine.ScreenSpaceReflection(ctx, width, height)
        self._halo_stream = None

    @staticmethod
    def gbuffer_tensors(gbuffer):
        """The attachments a rank needs replicated: depth / normal / albedo / MRO, plus the legacy position attachment when the
        G-buffer has one (mode P; 36 B/px instead of 20)."""
        imgs = ([gbuffer.position] if gbuffer.position is not None else []) + [gbuffer.depth, gbuffer.normal, gbuffer.albedo, gbuffer.mro]
        return [i.tensor for i in imgs]

    def broadcast_gbuffer(self, gbuffer, src: int = 0, async_op: bool = False):
        if self.world > 1:
            return broadcast_tensors(self.gbuffer_tensors(gbuffer), src=src, group=self.group, async_op=async_op)
        return None

    def halo_rows(self):
        """Rows [lo, hi) of reflection mip 0 the band's glossy mips read (the band and its halo)."""
        from . import engine
        return engine.band_rows(self.W, self.H, self.ssr.getReflectionBuffer().image.mips, self.y0, self.y1)[0]

    def halo_plan(self):
        return halo_plan(self.W, self.H, self.ssr.getReflectionBuffer().image.mips, self.world, self.rank)

    def exchange_halo(self):
        """Point-to-point exchange of the reflection rows (mip 0, RGBA16F) between the ranks whose bands border each other:
        a few MB per neighbour over NVLink instead of every rank ray-marching its halo again. Returns the requests to wait for."""
        dist = _dist()
        if self.world == 1:
            return []
        recv, send = self.halo_plan()
        t = self.ssr.getReflectionBuffer().image.tensor
        rb = self.W * 8  # bytes per row of mip 0
        ops = []
        # the same global order on every rank (by the lower rank of the pair, then direction): what NCCL's grouped P2P needs
        for q, lo, hi in send:
            ops.append(((min(self.rank, q), max(self.rank, q), 0 if self.rank < q else 1), dist.P2POp(dist.isend, t[lo * rb:hi * rb], q, group=self.group)))
        for q, lo, hi in recv:
            ops.append(((min(self.rank, q), max(self.rank, q), 0 if q < self.rank else 1), dist.P2POp(dist.irecv, t[lo * rb:hi * rb], q, group=self.group)))
        ops.sort(key=lambda o: o[0])
        return dist.batch_isend_irecv([o[1] for o in ops]) if ops else []

    def render(self, uniforms, gbuffer, ibl, lights, flags: int = _capi.SHADE_SKIP_TONEMAP, stream: int = 0, exchange_halo: Optional[bool] = None):
        """The band's frame. With more than one rank the reflection halo (the mip-0 rows around the band its glossy mips read) comes
        from the ranks that own those rows (exchange_halo, default) while this rank's SSAO runs; with exchange_halo=False every rank
        ray-marches its halo itself (no collective inside the frame, ~35 % more SSR work at 8 ranks of an 8K frame). Either way the
        band is bit-identical to the same rows of a single-GPU frame."""
        if exchange_halo is None:  # worth a synchronisation point between neighbours once the halo is a fair share of a band
            exchange_halo = exchanges_halo(self.W, self.H, self.ssr.getReflectionBuffer().image.mips, self.world)  # rank-independent
        if self.y1 <= self.y0:
            if self.world > 1 and exchange_halo:
                for w in self.exchange_halo():  # an empty band still owes nothing, but the grouped call must match its peers'
                    w.wait()
            return
        self.ctx.set_scissor_rows(self.y0, self.y1)
        base = self.ctx.flags
        try:
            if not exchange_halo:
                self.ssr.captureReflection(uniforms, gbuffer, ibl, lights, stream)
                self.ssr.convolveReflectionBuffer(stream)
                self.deferred.draw(uniforms, gbuffer, ibl, lights, self.ssr, flags, stream)
                return
            import torch
            self.ctx.set_flags(base | _capi.CTX_BAND_EXCHANGE_HALO)
            self.ssr.captureReflection(uniforms, gbuffer, ibl, lights, stream)      # the band's own rows only
            cuda = self.ssr.getReflectionBuffer().image.tensor.is_cuda
            if cuda:
                done = torch.cuda.Event()
                done.record()
                if self._halo_stream is None:
                    self._halo_stream = torch.cuda.Stream()
                self._halo_stream.wait_event(done)
                with torch.cuda.stream(self._halo_stream):
                    works = self.exchange_halo()
            else:
                works = self.exchange_halo()
            self.deferred.draw(uniforms, gbuffer, ibl, lights, self.ssr, _capi.SHADE_AO_ONLY, stream)  # SSAO while the halo travels
            for w in works:
                w.wait()
            self.ssr.convolveReflectionBuffer(stream)
            self.deferred.draw(uniforms, gbuffer, ibl, lights, self.ssr, flags | _capi.SHADE_AO_FROM_IMAGE, stream)
        finally:
            self.ctx.set_flags(base)
            self.ctx.set_scissor_rows(0, 0)

    def gather(self, async_op: bool = False):
        if self.world > 1:
            return allgather_rows(self._color_t, self.row_bytes, self.band, self.group, async_op=async_op)
        return None

    def color_rows(self):
        """(H, W * bytes_per_texel) uint8 view of the assembled colour target."""
        return self._color_t[: self.H * self.row_bytes].view(self.H, self.row_bytes)
