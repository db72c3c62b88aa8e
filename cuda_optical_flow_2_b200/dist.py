"""Multi-GPU host logic for the pyramidal LK path: one process per GPU over torch.distributed.

The reference is single-GPU (no streams, no NCCL, no threads); both modes here are new:

* frame-batch sharding -- pairs are independent, so rank r takes a contiguous block of the batch and
  runs the whole path on its own GPU.  No data-path collective; only timings / statistics are
  gathered at the end.  (`shard_range`, `flow_pairs_sharded`)
* row strips -- one very large pair is cut into horizontal strips.  Strips nest across pyramid
  levels (bounds are fixed on the coarsest level and doubled per finer level), every level needs
  its neighbours' image rows (stencil halo + warp reach) and the next-coarser level's cumulative
  flow rows, exchanged with the rank above and below once per level: `dist.batch_isend_irecv`
  (NCCL over NVLink on GPUs, gloo in the CPU tests).  The result is bit-identical to the 1-GPU
  result as long as no warp sample reaches past the exchanged rows; the kernel raises a flag
  otherwise and `StripRunner` turns it into an error instead of returning different numbers.

torch is used for device memory, streams and the process group only; all arithmetic is in the
C-ABI library.  The strip schedule itself (`StripPlan`) is plain Python and unit-tested on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from ._lib import WARP_AS_WRITTEN, WARP_BILINEAR
from .api import align_up


# ------------------------------------------------------------------------------------------ frame batch
def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n_items for `rank`; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flow_pairs_sharded(ctx, prev, next, w: int, levels: int, win: int, warp_mode: int = WARP_BILINEAR,
                       flow_scale: float = 1.0, world: int = 1, rank: int = 0, stream: int = 0):
    """Run this rank's block of a (n, h, pitch) uint8 CUDA batch; returns (lo, hi, flows) with flows the
    residual-flow tensors of the local pairs.  Every rank holds (or is given) only its own block."""
    lo, hi = shard_range(prev.shape[0], world, rank)
    if hi == lo:
        return lo, hi, []
    flows = ctx.flow_pairs_device(prev[lo:hi].contiguous(), next[lo:hi].contiguous(), w, levels, win, warp_mode=warp_mode,
                                  flow_scale=flow_scale, stream=stream)
    return lo, hi, flows


# ------------------------------------------------------------------------------------------ row strips
@dataclass
class LevelStrip:
    """Geometry of one pyramid level on one rank (all rows are GLOBAL row numbers of that level)."""
    w: int
    h: int          # height of the whole level
    y0: int         # rows this rank owns (and produces flow for): [y0, y1)
    y1: int
    by0: int        # rows its image buffers hold: own rows plus halo, clipped to the image: [by0, by1)
    by1: int
    cy0: int        # rows of the NEXT-COARSER level's cumulative flow its buffer holds: [cy0, cy1)
    cy1: int

    @property
    def own_rows(self) -> int:
        return self.y1 - self.y0

    @property
    def buf_rows(self) -> int:
        return self.by1 - self.by0


class StripPlan:
    """Row-strip schedule of a W x H pair over `world` ranks.

    Bounds are set on the coarsest level (floor(r*H_c/world)) and doubled per finer level so that a
    rank's strip at level k sits exactly on top of its strip at level k+1; the last rank also takes
    the rows that odd heights leave over.  `reach` is the largest vertical warp displacement (in
    rows of the level being warped) the halo must absorb.
    """

    def __init__(self, W: int, H: int, levels: int, win: int, world: int, reach: int = 16):
        if world < 1 or levels < 1 or win % 2 == 0:
            raise ValueError("bad strip plan arguments")
        self.W, self.H, self.levels, self.win, self.world, self.reach = W, H, levels, win, world, reach
        self.r = win // 2
        hc = H >> (levels - 1)
        if hc // world < 2 * (self.r + 2 + reach) // max(1, 2 ** (levels - 1)) and world > 1 and hc < world:
            raise ValueError(f"coarsest level has {hc} rows, cannot cut it into {world} strips")
        self.coarse_bounds = [(rk * hc) // world for rk in range(world)] + [hc]
        # image halo: 3x3 stencil + window radius + one row of even-row alignment + warp reach + bilinear tap
        self.img_halo = self.r + 2 + reach + 2
        # coarser cumulative-flow halo: rows (y >> 1) for every image-buffer row that can be warped or composed
        self.cum_halo = (self.r + 2) // 2 + 1

    def level(self, k: int, rank: int) -> LevelStrip:
        sh = self.levels - 1 - k
        w, h = self.W >> k, self.H >> k
        y0 = self.coarse_bounds[rank] << sh
        y1 = h if rank == self.world - 1 else self.coarse_bounds[rank + 1] << sh
        by0, by1 = max(0, y0 - self.img_halo), min(h, y1 + self.img_halo)
        if k < self.levels - 1:
            hc = self.H >> (k + 1)
            # the kernel looks up cum[(y >> 1)] for the rows it warps (its W tile: own rows -r-2 .. +r+1)
            cy0 = max(0, ((y0 - self.r - 2) >> 1))
            cy1 = min(hc, ((y1 + self.r + 1) >> 1) + 1)
        else:
            cy0 = cy1 = 0
        return LevelStrip(w, h, y0, y1, by0, by1, cy0, cy1)

    def halo_messages(self, k: int, rank: int) -> List[Tuple[int, int, int, int]]:
        """Image rows this rank must SEND at level k: list of (peer, first_row, last_row_exclusive, tag).
        A peer needs the part of its buffer [by0, by1) that this rank owns."""
        me = self.level(k, rank)
        out = []
        for peer in range(self.world):
            if peer == rank:
                continue
            pl = self.level(k, peer)
            lo, hi = max(me.y0, pl.by0), min(me.y1, pl.by1)
            if lo < hi:
                out.append((peer, lo, hi, k))
        return out

    def cum_messages(self, k: int, rank: int) -> List[Tuple[int, int, int, int]]:
        """Rows of cum_{k+1} this rank must SEND before level k is solved (it owns them at level k+1)."""
        if k >= self.levels - 1:
            return []
        mine = self.level(k + 1, rank)
        out = []
        for peer in range(self.world):
            if peer == rank:
                continue
            pl = self.level(k, peer)
            lo, hi = max(mine.y0, pl.cy0), min(mine.y1, pl.cy1)
            if lo < hi:
                out.append((peer, lo, hi, 100 + k))
        return out

    def validate(self) -> None:
        for k in range(self.levels):
            covered = 0
            for rk in range(self.world):
                ls = self.level(k, rk)
                if ls.y0 != covered:
                    raise AssertionError(f"level {k}: strips do not tile the image")
                if ls.own_rows < 1:
                    raise AssertionError(f"level {k}: rank {rk} owns no rows")
                covered = ls.y1
            if covered != (self.H >> k):
                raise AssertionError(f"level {k}: strips do not cover the image")


class Transport:
    """Moves halo rows between ranks.  `exchange(sends, recvs)`: sends = [(peer, tensor)], recvs =
    [(peer, tensor)] filled in place; both lists are ordered by peer on every rank."""

    def exchange(self, sends, recvs) -> None:  # pragma: no cover - interface
        raise NotImplementedError


class DistTransport(Transport):
    """torch.distributed point-to-point (NCCL on GPUs: NVLink/NVSwitch peer copies; gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group

    def exchange(self, sends, recvs) -> None:
        dist = self.dist
        ops = [dist.P2POp(dist.irecv, t, peer, group=self.group) for peer, t in recvs]
        ops += [dist.P2POp(dist.isend, t, peer, group=self.group) for peer, t in sends]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


class GatherTransport(Transport):
    """Halo exchange as ONE all-gather per exchange: every rank contributes two fixed-size slots (its message to
    the rank below, its message to the rank above) and reads its neighbours' slots.  More bytes than point-to-point
    (world x 2 slots, still a few MB over NVSwitch) but a single collective of constant shape per exchange, which
    -- unlike grouped send/recv -- captures cleanly in a CUDA graph, so a whole pair (kernels + exchanges) replays
    without host work.  Only adjacent ranks may talk (halo smaller than a strip).  Slot sizes are agreed on with an
    all-reduce the first time each exchange of a step is seen; `begin_step()` restarts the numbering."""

    def __init__(self, rank: int, world: int, device, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world, self.dev = rank, world, device
        self._idx, self._bufs = 0, []

    def begin_step(self) -> None:
        self._idx = 0

    def exchange(self, sends, recvs) -> None:
        torch, dist = self.torch, self.dist
        i = self._idx
        self._idx += 1
        if i >= len(self._bufs):  # first step only (never under graph capture): agree on the slot size
            n = max([t.numel() * t.element_size() for _, t in list(sends) + list(recvs)] + [16])
            tt = torch.tensor([n], dtype=torch.int64, device=self.dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=self.group)
            n = (int(tt.item()) + 15) // 16 * 16
            self._bufs.append((torch.zeros((2, n), dtype=torch.uint8, device=self.dev),
                               torch.zeros((self.world, 2, n), dtype=torch.uint8, device=self.dev)))
        out, inn = self._bufs[i]
        for peer, t in sends:
            if abs(peer - self.rank) != 1:
                raise ValueError("GatherTransport: only adjacent ranks exchange halos")
            b = t.reshape(-1).view(torch.uint8)
            out[0 if peer < self.rank else 1, : b.numel()].copy_(b)
        dist.all_gather_into_tensor(inn.view(-1), out.view(-1), group=self.group)
        for peer, t in recvs:
            if abs(peer - self.rank) != 1:
                raise ValueError("GatherTransport: only adjacent ranks exchange halos")
            nb = t.numel() * t.element_size()
            t.copy_(inn[peer, 1 if peer < self.rank else 0, :nb].view(t.dtype).view(t.shape))


class StripRunner:
    """One rank's share of a row-strip solve.  All tensors are torch tensors on this rank's device."""

    def __init__(self, ctx, plan: StripPlan, rank: int, transport: Transport, device, warp_mode: int = WARP_BILINEAR,
                 flow_scale: float = 1.0):
        import torch

        if warp_mode == WARP_AS_WRITTEN and plan.levels > 1:
            raise ValueError("OFB_WARP_AS_WRITTEN needs pixel (0,0) of every coarser level: not available on strips")
        self.torch, self.ctx, self.plan, self.rank, self.tp, self.dev = torch, ctx, plan, rank, transport, device
        self.warp_mode, self.flow_scale = warp_mode, flow_scale
        # CUDA stream handle the kernels are enqueued on (0 = the default stream).  `use_current_stream()`
        # re-reads torch's current stream, which is what CUDA-graph capture of `step()` needs.
        self.stream = 0
        self.strips = [plan.level(k, rank) for k in range(plan.levels)]
        self.pitch = [align_up(s.w, 64) for s in self.strips]
        z = lambda rows, cols, dt: torch.zeros((rows, cols), dtype=dt, device=device)
        self.prev = [z(s.buf_rows, p, torch.uint8) for s, p in zip(self.strips, self.pitch)]
        self.next = [z(s.buf_rows, p, torch.uint8) for s, p in zip(self.strips, self.pitch)]
        # flow / cumulative flow: rows of the image buffer (same origin by0), only own rows are written
        self.flow = [torch.zeros((s.buf_rows, s.w, 2), dtype=torch.float32, device=device) for s in self.strips]
        self.cum = [torch.zeros((s.buf_rows, s.w, 2), dtype=torch.float32, device=device) for s in self.strips]
        # coarser cumulative flow as the level-k kernel wants it: rows [cy0, cy1) of level k+1
        self.cum_in = [torch.zeros((max(s.cy1 - s.cy0, 1), s.w >> 1, 2), dtype=torch.float32, device=device)
                       for s in self.strips]
        self.overflow = torch.zeros(1, dtype=torch.int32, device=device)

    def use_current_stream(self) -> None:
        self.stream = self.torch.cuda.current_stream(self.dev).cuda_stream

    def step(self, prev_rows, next_rows) -> None:
        """One pair, everything asynchronous on the stream (no host synchronisation): capturable in a CUDA
        graph together with its NCCL halo exchanges.  Call check_overflow() afterwards (it synchronises)."""
        if hasattr(self.tp, "begin_step"):
            self.tp.begin_step()
        self.load_level0(prev_rows, next_rows)
        self.build_pyramid()
        self.solve(check=False)

    # -- data in ---------------------------------------------------------------------------------
    def load_level0(self, prev_rows, next_rows) -> None:
        """prev_rows/next_rows: (own_rows, w) uint8 tensors with this rank's OWN rows of the two frames."""
        s = self.strips[0]
        a = s.y0 - s.by0
        self.prev[0][a:a + s.own_rows, :s.w] = prev_rows
        self.next[0][a:a + s.own_rows, :s.w] = next_rows

    # -- exchanges -------------------------------------------------------------------------------
    # Each exchange is described as (sends, recvs, finish): tensors to send per peer, buffers to
    # receive into per peer, and a closure that files the received rows.  `_run` hands them to the
    # transport; the single-process emulation matches sends to recvs itself.
    def image_exchange(self, k: int):
        plan, s, torch = self.plan, self.strips[k], self.torch
        sends, recvs, stage = [], [], []
        for peer, lo, hi, _ in plan.halo_messages(k, self.rank):
            sends.append((peer, torch.stack([self.prev[k][lo - s.by0:hi - s.by0],
                                             self.next[k][lo - s.by0:hi - s.by0]]).contiguous()))
        for peer in range(plan.world):
            if peer == self.rank:
                continue
            pl = plan.level(k, peer)
            lo, hi = max(pl.y0, s.by0), min(pl.y1, s.by1)
            if lo < hi:
                buf = torch.empty((2, hi - lo, self.pitch[k]), dtype=torch.uint8, device=self.dev)
                recvs.append((peer, buf))
                stage.append((lo, hi, buf))

        def finish():
            for lo, hi, buf in stage:
                self.prev[k][lo - s.by0:hi - s.by0] = buf[0]
                self.next[k][lo - s.by0:hi - s.by0] = buf[1]

        return sends, recvs, finish

    def cum_exchange(self, k: int):
        """Rows [cy0, cy1) of cum_{k+1} into cum_in[k]: own rows copied here, the rest from neighbours."""
        plan, s, torch = self.plan, self.strips[k], self.torch
        up = self.strips[k + 1]
        src = self.cum[k + 1] if k + 1 < plan.levels - 1 else self.flow[k + 1]  # cum of the coarsest level is its flow
        lo, hi = max(up.y0, s.cy0), min(up.y1, s.cy1)
        if lo < hi:
            self.cum_in[k][lo - s.cy0:hi - s.cy0] = src[lo - up.by0:hi - up.by0]
        sends, recvs, stage = [], [], []
        for peer, lo, hi, _ in plan.cum_messages(k, self.rank):
            sends.append((peer, src[lo - up.by0:hi - up.by0].contiguous()))
        for peer in range(plan.world):
            if peer == self.rank:
                continue
            pu = plan.level(k + 1, peer)
            lo, hi = max(pu.y0, s.cy0), min(pu.y1, s.cy1)
            if lo < hi:
                buf = torch.empty((hi - lo, up.w, 2), dtype=torch.float32, device=self.dev)
                recvs.append((peer, buf))
                stage.append((lo, hi, buf))

        def finish():
            for lo, hi, buf in stage:
                self.cum_in[k][lo - s.cy0:hi - s.cy0] = buf

        return sends, recvs, finish

    def _run(self, ex) -> None:
        sends, recvs, finish = ex
        self.tp.exchange(sends, recvs)
        finish()

    # -- compute ---------------------------------------------------------------------------------
    def build_pyramid(self) -> None:
        """Level k+1 own rows from level k own rows +-1 (after the level-k halo exchange), for both frames."""
        plan = self.plan
        for k in range(plan.levels):
            self._run(self.image_exchange(k))
            if k + 1 < plan.levels:
                self.pyr_down_own(k)

    def pyr_down_own(self, k: int) -> None:
        s, d = self.strips[k], self.strips[k + 1]
        for buf in (self.prev, self.next):
            self.ctx.pyr_down_strip_device(buf[k], s.w, s.by0, buf[k + 1][d.y0 - d.by0:d.y1 - d.by0], d.y0, d.y1,
                                           stream=self.stream)

    def lk_level_own(self, k: int) -> None:
        plan, s = self.plan, self.strips[k]
        want_cum = (0 < k < plan.levels - 1) or (k == 0 and plan.levels > 1)
        self.ctx.lk_level_strip_device(self.prev[k], self.next[k], s.w, s.by0, s.h, s.y0 - s.by0, s.y1 - s.by0, plan.win,
                                       self.warp_mode, self.flow_scale, self.cum_in[k] if k < plan.levels - 1 else None,
                                       s.cy0, self.flow[k], cum_out=self.cum[k] if want_cum else None,
                                       overflow_flag=self.overflow, stream=self.stream)

    def check_overflow(self) -> None:
        if int(self.overflow.item()) != 0:
            raise RuntimeError(f"rank {self.rank}: a warp sample reached past the exchanged halo rows: raise "
                               f"StripPlan.reach (currently {self.plan.reach} rows)")

    def solve(self, check: bool = True) -> None:
        """Coarse to fine: exchange cum_{k+1} halo rows, run the fused strip kernel on own rows."""
        plan = self.plan
        self.overflow.zero_()
        for k in range(plan.levels - 1, -1, -1):
            if k < plan.levels - 1:
                self._run(self.cum_exchange(k))
            self.lk_level_own(k)
        if check:
            self.check_overflow()

    def own_flow(self, k: int):
        s = self.strips[k]
        return self.flow[k][s.y0 - s.by0:s.y1 - s.by0]

    def own_total_flow(self):
        s = self.strips[0]
        src = self.cum[0] if self.plan.levels > 1 else self.flow[0]
        return src[s.y0 - s.by0:s.y1 - s.by0]


class NativeStrips:
    """The native row-strip runner (csrc/strips.cu): the same schedule as StripRunner, but a whole pair is enqueued
    from C++ on one stream.  transport "nccl": the halo rows move by NCCL send/recv between the ranks' buffers;
    "peer": by the sender's copy kernel straight into the receiver's memory over NVLink (CUDA IPC), epoch flags instead
    of a collective; "local": peer memory between ranks living in this process (connect_local, tests).
    `torch.distributed` is used once, to hand round the NCCL unique id or the IPC handles."""

    def __init__(self, ctx, w: int, h: int, levels: int, win: int, world: int, rank: int, device, warp_mode: int = WARP_BILINEAR,
                 flow_scale: float = 1.0, reach: int = 16, transport: str = "nccl"):
        import ctypes as C

        import torch

        from . import _lib as L

        if transport not in ("nccl", "peer", "local"):
            raise ValueError(f"unknown strip transport {transport!r}")
        self.torch, self.C, self.L, self.lib, self.ctx = torch, C, L, L.load(), ctx
        self.w, self.h, self.levels, self.dev, self.world, self.rank = w, h, levels, device, world, rank
        self._transport = transport
        idbuf = torch.zeros(128, dtype=torch.uint8)
        use_nccl = world > 1 and transport == "nccl"
        if use_nccl:
            import torch.distributed as dist

            if rank == 0:
                raw = (C.c_ubyte * 128)()
                L.check(self.lib.ofb_strips_nccl_unique_id(C.cast(raw, C.c_void_p)))
                idbuf = torch.tensor(list(raw), dtype=torch.uint8)
            idbuf = idbuf.to(device)
            dist.broadcast(idbuf, 0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        hnd = C.c_void_p()
        L.check(self.lib.ofb_strips_create(ctx._h, w, h, levels, win, warp_mode, C.c_float(flow_scale), world, rank, reach,
                                           C.cast(raw, C.c_void_p) if use_nccl else None, C.byref(hnd)))
        self._h = hnd
        if world > 1 and transport == "peer":
            import torch.distributed as dist

            blob = (C.c_ubyte * 128)()
            L.check(self.lib.ofb_strips_peer_handle(self._h, C.cast(blob, C.c_void_p)))
            mine = torch.tensor(list(blob), dtype=torch.uint8, device=device)
            every = torch.empty(world * 128, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(every, mine)
            blobs = (C.c_ubyte * (world * 128))(*every.cpu().tolist())
            L.check(self.lib.ofb_strips_peer_connect(self._h, C.cast(blobs, C.c_void_p)))
            dist.barrier()  # nobody pushes before everybody has mapped its neighbours

    def input_rows(self):
        """(prev, next) uint8 torch views (own_rows, pitch) of where the own rows of level 0 live inside the runner:
        fill them in place and pass them to run() to save the upload copy."""
        C = self.C
        pp, pn, pitch = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self.L.check(self.lib.ofb_strips_input(self._h, C.byref(pp), C.byref(pn), C.byref(pitch)))
        y0, y1 = self.own_rows(0)

        def view(ptr):
            class _H:
                __cuda_array_interface__ = {"shape": ((y1 - y0) * pitch.value,), "typestr": "|u1", "data": (ptr, False), "version": 3}

            return self.torch.as_tensor(_H(), device=self.dev).view(y1 - y0, pitch.value)

        return view(pp.value), view(pn.value)

    def set_total(self, on: bool) -> None:
        self.L.check(self.lib.ofb_strips_set_total(self._h, 1 if on else 0))

    def set_fused(self, on: bool) -> None:
        """Peer-memory transport: cumulative-flow halo rows pushed by the level kernels themselves (default) or by
        separate copy / wait kernels."""
        self.L.check(self.lib.ofb_strips_set_fused(self._h, 1 if on else 0))

    def arena(self) -> int:
        p = self.C.c_void_p()
        self.L.check(self.lib.ofb_strips_peer_arena(self._h, self.C.byref(p)))
        return p.value

    @staticmethod
    def connect_local(ranks) -> None:
        """Peer-memory transport between NativeStrips objects of ONE process (ranks[r] is rank r)."""
        import ctypes as C

        arenas = (C.c_void_p * len(ranks))(*[r.arena() for r in ranks])
        for r in ranks:
            r.L.check(r.lib.ofb_strips_peer_connect_local(r._h, arenas))

    def own_rows(self, level: int = 0):
        y0, y1 = self.C.c_int(), self.C.c_int()
        self.L.check(self.lib.ofb_strips_own_rows(self._h, level, self.C.byref(y0), self.C.byref(y1)))
        return y0.value, y1.value

    def run(self, prev_own, next_own, stream: int = 0) -> None:
        """prev_own / next_own: (own_rows, pitch) uint8 CUDA tensors (or row-sliced views) of level 0."""
        self.L.check(self.lib.ofb_strips_run_device(self._h, prev_own.data_ptr(), next_own.data_ptr(), prev_own.stride(0),
                                                    self.C.c_void_p(stream)))

    def run_phase(self, prev_own, next_own, phase: int, stream: int = 0) -> None:
        """One of the 2 * levels phases of a pair (peer-memory transport); see run_local_sequenced."""
        self.L.check(self.lib.ofb_strips_run_phase_device(self._h, prev_own.data_ptr(), next_own.data_ptr(), prev_own.stride(0),
                                                          int(phase), self.C.c_void_p(stream)))

    @staticmethod
    def run_local_sequenced(ranks, inputs, stream: int = 0) -> None:
        """One pair on ranks that live in ONE process and share ONE device (connect_local): phase p of every rank is
        enqueued on `stream` before phase p + 1 of any, so every flag a kernel waits for has been raised by a kernel
        AHEAD of it in the stream.  (Ranks that spin on each other from separate streams of one GPU are not
        guaranteed to be co-scheduled; across GPUs each rank simply calls run().)"""
        for phase in range(2 * ranks[0].levels):
            for ns, (pin, nin) in zip(ranks, inputs):
                ns.run_phase(pin, nin, phase, stream)

    def check(self, stream: int = 0) -> None:
        """Synchronises; raises if any pair since the last check went wrong (the flag is sticky), and clears the flag."""
        v = self.C.c_int()
        self.L.check(self.lib.ofb_strips_check(self._h, self.C.c_void_p(stream), self.C.byref(v)))
        if v.value & 2:
            raise RuntimeError("a neighbour's halo rows did not arrive (peer-memory transport timed out)")
        if v.value & 1:
            raise RuntimeError("a warp sample reached past the exchanged halo rows: raise `reach`")

    def own_flow(self, level: int, total: bool = False):
        """A torch view (no copy) of the own rows of the residual (or cumulative) flow of `level`."""
        f, t = self.C.c_void_p(), self.C.c_void_p()
        self.L.check(self.lib.ofb_strips_result(self._h, level, self.C.byref(f), self.C.byref(t)))
        y0, y1 = self.own_rows(level)
        n = (y1 - y0) * (self.w >> level) * 2
        ptr = t.value if total else f.value
        return _device_view(self.torch, ptr, (y1 - y0, self.w >> level, 2), self.dev)

    def close(self) -> None:
        """Collective across the ranks of a multi-process run: nobody frees the memory its neighbours write into before
        everybody has finished (the library additionally waits for the neighbours' last acknowledgements)."""
        if getattr(self, "_h", None):
            if self.world > 1 and self._transport != "local":
                import torch.distributed as dist

                if dist.is_available() and dist.is_initialized():
                    self.torch.cuda.synchronize(self.dev)
                    dist.barrier()
            self.lib.ofb_strips_destroy(self._h)
            self._h = None

    def __del__(self):
        # (no collective from a finaliser: only the library-side bounded wait)
        try:
            if getattr(self, "_h", None):
                self.lib.ofb_strips_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _device_view(torch, ptr: int, shape, device):
    """float32 torch tensor over device memory owned by the library (valid while its owner lives)."""
    n = 1
    for d in shape:
        n *= d

    class _Holder:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Holder(), device=device).view(*shape)


def run_strips_local(ctx, prev0, next0, levels: int, win: int, world: int, warp_mode: int = WARP_BILINEAR,
                     flow_scale: float = 1.0, reach: int = 16):
    """Emulate `world` ranks on ONE device: the same StripRunner code, with every halo exchange done
    as local copies between the rank objects.  prev0/next0: (H, >=W) uint8 tensors holding the whole
    frames.  Returns (plan, runners).  Used to check that strips + halos reproduce the whole-frame
    result bit for bit without needing several GPUs."""
    return _run_strips_local(ctx, prev0, next0, levels, win, world, warp_mode, flow_scale, reach, None)


def run_strips_local_w(ctx, prev0, next0, w: int, levels: int, win: int, world: int, warp_mode: int = WARP_BILINEAR,
                       flow_scale: float = 1.0, reach: int = 16):
    """As run_strips_local for pitched frames: only the first w columns of each row are image."""
    return _run_strips_local(ctx, prev0, next0, levels, win, world, warp_mode, flow_scale, reach, w)


def _run_strips_local(ctx, prev0, next0, levels, win, world, warp_mode, flow_scale, reach, width: Optional[int]):
    H = prev0.shape[0]
    W = width if width is not None else prev0.shape[1]
    plan = StripPlan(W, H, levels, win, world, reach)
    plan.validate()
    runners = [StripRunner(ctx, plan, rk, Transport(), prev0.device, warp_mode, flow_scale) for rk in range(world)]
    for rn in runners:
        s = rn.strips[0]
        rn.load_level0(prev0[s.y0:s.y1, :W], next0[s.y0:s.y1, :W])

    def lockstep(make):
        posts = [(rn, make(rn)) for rn in runners]
        outbox = {(rn.rank, dst): t for rn, (sends, _, _) in posts for dst, t in sends}
        for rn, (_, recvs, finish) in posts:
            for src, t in recvs:
                t.copy_(outbox[(src, rn.rank)])
            finish()

    for k in range(levels):
        lockstep(lambda rn: rn.image_exchange(k))
        if k + 1 < levels:
            for rn in runners:
                rn.pyr_down_own(k)
    for rn in runners:
        rn.overflow.zero_()
    for k in range(levels - 1, -1, -1):
        if k < levels - 1:
            lockstep(lambda rn: rn.cum_exchange(k))
        for rn in runners:
            rn.lk_level_own(k)
    for rn in runners:
        rn.check_overflow()
    return plan, runners
