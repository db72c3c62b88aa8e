#!/usr/bin/env python
"""bench.py -- megapixel frame-pairs/s of the dense pyramidal LK path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs B] [--impl b200|reference]

Workload (BASELINE.json configs[1]): 1920x1080 frame pairs, 3-level Gaussian pyramid, 9x9 window.
A step is one pass of the whole path (both pyramids + 3 fused LK levels) over a batch of B synthetic
pairs per GPU.  `value` is timed with CUDA events with every input already resident in HBM; the
batch's inputs (B x 4.1 MB) are larger than the 126 MB L2, so no step re-reads cached frames.
`e2e` is the same metric through the host-pointer C-ABI call (ofb_flow_pairs_host) with pinned host
buffers in the reference's 3-channel layout, H2D and D2H inside the timed region.
N > 1 (torchrun): pairs are sharded per GPU, no data-path collective, weak scaling.

`--impl reference` times the reference's own CPU implementation (cpu::gauss_pyramid +
cpu::calc_optical_flow, compiled unmodified into oracle/_ref/libofref.so) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, LEVELS, WIN = 1920, 1080, 3, 9
METRIC = "megapixel frame-pairs/sec (1080p, 3-level, win 9)"
UNIT = "Mpx-pairs/s"


# ----------------------------------------------------------------------------------------------
def synth_pairs_torch(n, w, h, device, seed):
    """Value-noise frames on the GPU: random 8-px grid, bilinear, next = prev shifted sub-pixel."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    cell = 8
    gw, gh = w // cell + 6, h // cell + 6
    pitch = (w + 63) // 64 * 64
    prev = torch.zeros((n, h, pitch), dtype=torch.uint8, device=device)
    nxt = torch.zeros((n, h, pitch), dtype=torch.uint8, device=device)
    xs = torch.arange(w, device=device, dtype=torch.float32)
    ys = torch.arange(h, device=device, dtype=torch.float32)
    for i in range(n):
        grid = torch.randint(0, 256, (gh, gw), generator=g).to(device=device, dtype=torch.float32)
        dxy = (torch.rand(2, generator=g) * 8.0 - 4.0).tolist()  # up to +-4 px: the pyramid matters
        for dst, (dx, dy) in ((prev, (0.0, 0.0)), (nxt, dxy)):
            fx = (xs - dx) / cell + 2.5
            fy = (ys - dy) / cell + 2.5
            ix = fx.floor().clamp(0, gw - 2).long()
            iy = fy.floor().clamp(0, gh - 2).long()
            ax = (fx - ix).view(1, -1)
            ay = (fy - iy).view(-1, 1)
            g00 = grid[iy][:, ix]
            g01 = grid[iy][:, ix + 1]
            g10 = grid[iy + 1][:, ix]
            g11 = grid[iy + 1][:, ix + 1]
            v = (1 - ay) * ((1 - ax) * g00 + ax * g01) + ay * ((1 - ax) * g10 + ax * g11)
            dst[i, :, :w] = v.clamp(0, 255).to(torch.uint8)
    return prev, nxt, pitch


def synth_pairs_numpy(n, w, h, seed):
    """Same construction on the host (3-channel reference layout) for the CPU arms and e2e."""
    rng = np.random.default_rng(seed)
    cell = 8
    gw, gh = w // cell + 6, h // cell + 6
    prev = np.empty((n, h, w, 3), np.uint8)
    nxt = np.empty((n, h, w, 3), np.uint8)
    xs = np.arange(w, dtype=np.float32)
    ys = np.arange(h, dtype=np.float32)
    for i in range(n):
        grid = rng.integers(0, 256, (gh, gw)).astype(np.float32)
        dxy = rng.random(2) * 8.0 - 4.0
        for dst, (dx, dy) in ((prev, (0.0, 0.0)), (nxt, dxy)):
            fx = (xs - np.float32(dx)) / cell + 2.5
            fy = (ys - np.float32(dy)) / cell + 2.5
            ix = np.clip(np.floor(fx), 0, gw - 2).astype(np.int64)
            iy = np.clip(np.floor(fy), 0, gh - 2).astype(np.int64)
            ax = (fx - ix)[None, :]
            ay = (fy - iy)[:, None]
            g00, g01 = grid[iy][:, ix], grid[iy][:, ix + 1]
            g10, g11 = grid[iy + 1][:, ix], grid[iy + 1][:, ix + 1]
            v = (1 - ay) * ((1 - ax) * g00 + ax * g01) + ay * ((1 - ax) * g10 + ax * g11)
            dst[i] = np.clip(v, 0, 255).astype(np.uint8)[:, :, None]
    return prev, nxt


class ClockSampler:
    """SM clock, power and throttle reasons sampled through NVML every 5 ms while the timed region runs."""

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self.err = index, [], threading.Event(), None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as ex:  # no NVML: report it instead of inventing numbers
            self.err = repr(ex)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((clk, rs, pw))
            except Exception as ex:
                self.err = repr(ex)
                return
            time.sleep(0.005)

    def stop(self):
        if self.err and not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + self.err], "samples": 0}
        self._stop.set()
        self.t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        # samples under load = the upper half by power draw (the region also covers launch gaps)
        s = sorted(self.samples, key=lambda x: x[2])
        hot = s[len(s) // 2:] if s else []
        reasons = sorted({n for n, bit in names.items() for (_, rs, _) in self.samples if rs & bit})
        return {"sm_mhz": float(np.median([c for c, _, _ in hot])) if hot else None, "sm_max_mhz": float(self.max_sm),
                "reasons": reasons, "samples": len(self.samples),
                "power_w_max": max((pw for _, _, pw in self.samples), default=None)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the level-0 kernel from the committed ncu --set full summary."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------- reference arm
def ref_lib():
    from oracle import oracle as O  # the one place bench.py executes oracle/: the CPU baseline

    if not O.have_ref():
        return None, O
    return O.ref(), O


def run_reference_cpu(n_pairs: int, threads: int, reps: int = 1):
    """Reference CPU path on `threads` host threads, one pair per task.  Returns (seconds per rep, cores)."""
    from concurrent.futures import ThreadPoolExecutor

    lib, O = ref_lib()
    prev, nxt = synth_pairs_numpy(min(n_pairs, 4), W, H, 7)
    kind = "reference"
    if lib is None:
        kind = "port"

    def one(i):
        p, q = prev[i % len(prev)], nxt[i % len(nxt)]
        if lib is not None:
            flows = [np.zeros((H >> k, W >> k, 2), np.float32) for k in range(LEVELS)]
            ptr = (O._f32p * LEVELS)(*[f.ctypes.data_as(O._f32p) for f in flows])
            lib.ref_cpu_flow_pair(p.ctypes.data_as(O._u8p), q.ctypes.data_as(O._u8p), W, H, LEVELS, ptr)
        else:
            O.flow_pair(np.ascontiguousarray(p[:, :, 0]), np.ascontiguousarray(q[:, :, 0]), LEVELS, WIN,
                        O.WARP_AS_WRITTEN, O.SUMS_F32_SEQUENTIAL)

    times = []
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(reps):
            t0 = time.perf_counter()
            list(ex.map(one, range(n_pairs)))
            times.append(time.perf_counter() - t0)
    return times, kind


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_pairs = cores  # one pair per host thread per step: a bounded sample of the workload
    times, kind = run_reference_cpu(n_pairs, cores, reps=args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    value = (W * H / 1e6) * n_pairs * len(timed) / total
    sample = f"{n_pairs} pairs per step on {cores} threads, {len(timed)} steps"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8->int->f64 (CPU)",
            "data": "synthetic", "config": {"workload": "1920x1080 pair, 3-level pyramid, 9x9 window",
                                            "pairs_per_step": n_pairs, "where": "host CPU"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------- B200 arm
def main_b200(args):
    import torch
    import torch.distributed as dist

    from cuda_optical_flow_2_b200 import WARP_BILINEAR, Context, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the B200 arm has no CPU fallback"}), flush=True)
        return 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.pairs
    ctx = Context(local)
    prev, nxt, pitch = synth_pairs_torch(B, W, H, dev, 1000 + rank)
    flows = [torch.empty((B, H >> k, W >> k, 2), dtype=torch.float32, device=dev) for k in range(LEVELS)]
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ctx.flow_pairs_device(prev, nxt, W, LEVELS, WIN, warp_mode=WARP_BILINEAR, flows=flows, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.profile_enable(True)
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    lvl_ms = [ctx.profile_read(k) for k in range(LEVELS)]
    pyr_ms = ctx.profile_read(_lib.PROFILE_PYRAMID)
    ctx.profile_enable(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    if world > 1:  # kernels launched inside the timed region, all ranks
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = (W * H / 1e6) * B * world / (ms_step / 1e3)

    # ---- latency of ONE pair (BASELINE configs[1] as the reference's frame loop meets it): device-resident,
    #      back-to-back calls on the stream; reported beside the batched throughput, not instead of it
    one_ms = one_graph_ms = None
    if rank == 0:
        f1 = [f[:1] for f in flows]
        for _ in range(5):
            ctx.flow_pairs_device(prev[:1], nxt[:1], W, LEVELS, WIN, warp_mode=WARP_BILINEAR, flows=f1, stream=stream)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            ctx.flow_pairs_device(prev[:1], nxt[:1], W, LEVELS, WIN, warp_mode=WARP_BILINEAR, flows=f1, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        one_ms = e0.elapsed_time(e1) / 50
        # the same call captured once as a CUDA graph and replayed (no allocation, no host state per call)
        try:
            if world > 1:  # the single-GPU line carries it; no capture next to a live NCCL communicator
                raise RuntimeError("skipped at world > 1")
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                ctx.flow_pairs_device(prev[:1], nxt[:1], W, LEVELS, WIN, warp_mode=WARP_BILINEAR, flows=f1,
                                      stream=torch.cuda.current_stream(dev).cuda_stream)
            for _ in range(5):
                g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(50):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            one_graph_ms = e0.elapsed_time(e1) / 50
        except Exception as ex:  # reported, never fatal
            one_graph_ms = None
            if world == 1:
                print(f"single-pair graph capture failed: {ex!r}", file=sys.stderr)

    # ---- end to end through the host-pointer C-ABI call (rank-local, all ranks run it concurrently)
    Be = args.e2e_pairs
    hp, hn = synth_pairs_numpy(min(Be, 2), W, H, 2000 + rank)
    lib = _lib.load()

    def pinned(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _lib.check(lib.ofb_host_alloc(C.byref(p), n))
        return np.frombuffer((C.c_uint8 * n).from_address(p.value), dtype=dtype).reshape(shape), p

    hprev, p1 = pinned((Be, H, W, 3), np.uint8)
    hnext, p2 = pinned((Be, H, W, 3), np.uint8)
    for i in range(Be):
        hprev[i], hnext[i] = hp[i % len(hp)], hn[i % len(hn)]
    houts, pouts = [], []
    for k in range(LEVELS):
        a, p = pinned((Be, H >> k, W >> k, 2), np.float32)
        houts.append(a)
        pouts.append(p)
    for _ in range(2):
        ctx.flow_pairs_host(hprev, hnext, LEVELS, WIN, warp_mode=WARP_BILINEAR, out=houts)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.flow_pairs_host(hprev, hnext, LEVELS, WIN, warp_mode=WARP_BILINEAR, out=houts)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = (W * H / 1e6) * Be * world * e2e_steps / te.item()
    h2d = int(hprev.nbytes + hnext.nbytes)
    d2h = int(sum(a.nbytes for a in houts))
    # what bounds e2e: the device-to-host link.  One plain pinned copy of the level-0 flow buffer, timed the same way.
    pcie_d2h = None
    if rank == 0:
        src = flows[0][:Be].contiguous() if B >= Be else flows[0]
        dst = torch.from_numpy(houts[0].reshape(-1)[:src.numel()])
        dst.copy_(src.reshape(-1))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src.reshape(-1), non_blocking=True)
        torch.cuda.synchronize()
        pcie_d2h = 3 * src.numel() * 4 / (time.perf_counter() - t0) / 1e9
    for p in [p1, p2] + pouts:
        lib.ofb_host_free(p)

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel: the fused LK kernel at level 0.  Algorithmic bytes per launch:
        # 1 B prev + 1 B next + 8 B flow out per pixel + 8 B per coarser pixel of cumulative flow in
        n0 = W * H
        alg_bytes = B * (n0 * (1 + 1 + 8) + (W >> 1) * (H >> 1) * 8)
        k_ms, k_n = lvl_ms[0]
        avg_ms = k_ms / max(k_n, 1)
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        traffic = ncu_traffic()
        roofline = {"bound": "hbm", "kernel": "lk_level_kernel<9,bilinear> level 0", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms,
                    "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                    "traffic_note": (traffic or {}).get("note"),
                    "step_share": {"pyramid": pyr_ms[0] / (ms_step * args.steps) if ms_step else None,
                                   **{f"lk_level_{k}": lvl_ms[k][0] / (ms_step * args.steps) for k in range(LEVELS)}}}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                times, kind = run_reference_cpu(cores, cores, reps=1)
                cpu = {"value": (W * H / 1e6) * cores / times[0], "unit": UNIT, "cores": cores, "kind": kind,
                       "sample": f"{cores} pairs of the same 1080p/3-level/win-9 workload, one per host thread, "
                                 f"{times[0]:.1f} s wall"}
            except Exception as ex:  # the baseline is a report, never a reason to lose the GPU number
                cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "unavailable", "sample": repr(ex)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8 -> int32 sums -> f64 solve -> f32 flow", "data": "synthetic",
                "config": {"workload": "1920x1080 pair, 3-level Gaussian pyramid, 9x9 window, bilinear warp",
                           "pairs_per_gpu_per_step": B, "single_pair_latency_ms": one_ms, "single_pair_latency_ms_cuda_graph": one_graph_ms, "l2": f"inputs exceed L2 ({B * 2 * pitch * H / 1e6:.0f} MB of "
                           "frames per step, fresh outputs each level)", "parallelism": f"frame-batch x{world}"},
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "pairs_per_step": Be, "layout": "reference 3-channel u8 in, float2 flow of every level out",
                        "d2h_gbs_achieved": d2h * world * e2e_steps / te.item() / 1e9 / world,
                        "d2h_gbs_plain_copy": pcie_d2h,
                        "bound": "PCIe device-to-host: 10.5 B of flow per pixel leave the GPU, 6 B of frames enter"},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=256, help="frame pairs per GPU per step (device-resident)")
    ap.add_argument("--e2e-pairs", type=int, default=32, help="frame pairs per GPU per end-to-end step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_b200(args)


if __name__ == "__main__":
    sys.exit(main())
