#!/usr/bin/env python
"""bench.py -- NT-Xent forward+backward throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (normalise -> [all-gather] -> fused forward -> backward) over
one synthetic batch.  Workload at every N: BASELINE.json's target configuration, global batch 32768
pairs x d=128, tau=0.5, both inputs requiring grad, row-sharded over the N ranks (strong scaling);
it fits one GPU because the 2Bx2B logit matrix is never materialised.  configs[1] (4096 pairs) is
reported as a secondary number inside ``config`` at N=1.

Prints ONE JSON line on rank 0 (see the keys below).  ``--impl reference`` times the reference's
own CPU implementation of the path (the torch port in oracle/, all host threads) on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ntxent_fwd_bwd_pairs_per_sec"
UNIT = "pairs/s"
PAIRS = 32768
DIM = 128
TAU = 0.5


def workload_name(pairs, dim, tau):
    return f"NT-Xent fwd+bwd, global batch {pairs} pairs, d={dim}, tau={tau}, both inputs require grad"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(bf16=float(p["bf16_tflops"]), bf16_sustained=float(p.get("bf16_tflops_sustained", 0)),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples SM clock, power and throttle reasons (NVML, ~10 ms period) while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.err = None

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, pw, rs))
                time.sleep(0.01)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[f"no samples ({self.err})"])
        sm = [x[0] for x in self.samples]
        mask = 0
        for x in self.samples:
            mask |= x[2]
        reasons = sorted(n for bit, n in self.REASONS.items() if mask & bit)
        return dict(sm_mhz=statistics.median(sm), sm_min_mhz=min(sm), sm_max_mhz=self.sm_max,
                    power_w_max=max(x[1] for x in self.samples), samples=len(sm), reasons=reasons)


def run_reference(args, rank):
    """CPU arm: the reference's own implementation (torch port) on all host threads."""
    if rank != 0:
        return
    from oracle.cpu_baseline import time_port_stripe
    import torch
    b_sample = 256
    r = time_port_stripe(PAIRS, DIM, TAU, b_sample, steps=args.steps, warmup=max(args.warmup, 2))
    sample = (f"stripe of {r['b_sample']} anchor pairs x all {PAIRS} global keys per step, fwd+bwd "
              f"(keys constant as in the reference's world_size>1 branch), torch {torch.__version__} fp32")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 2),
        "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(PAIRS, DIM, TAU), "pairs_global": PAIRS, "dim": DIM,
                   "temperature": TAU},
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["pairs_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


_REAL_STDOUT = sys.stdout


def main():
    # Exactly one line may reach stdout (the JSON): libraries such as NCCL print banners there, so
    # fd 1 is pointed at stderr for the whole run and the JSON goes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS, help="global batch (pairs)")
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--tau", type=float, default=TAU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import maai_b200
    from maai_b200.Objective import _Profiler

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = maai_b200._lib.load()
    peaks = load_peaks()

    from maai_b200.Objective import _peer_state

    def peer_mode_used():  # decided (and voted on by the ranks) at the first call with a given shape
        return world > 1 and any(v for k, v in _peer_state.items() if isinstance(k, tuple) and k[0] == "usable")
    B, d, tau = args.pairs, args.dim, args.tau
    assert B % world == 0
    b = B // world

    def make_inputs(bb):
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        return (torch.randn(bb, d, generator=g, device=dev), torch.randn(bb, d, generator=g, device=dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(x, y):
        x.grad = None
        y.grad = None
        loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, local_rank=rank, world_size=world,
                                                device=dev)
        loss.backward()
        return loss

    def timed_run(bb, steps, warmup, profile, grad1=True):
        h1, h2 = make_inputs(bb)
        x = h1.requires_grad_(grad1)
        y = h2.requires_grad_(True)
        for _ in range(warmup):
            step(x, y)
        barrier()
        _Profiler.reset()
        _Profiler.enabled = profile
        launches0 = lib.maai_launch_count()
        evs = []
        for _ in range(steps):
            flush_buf.fill_(1)  # flush L2 between timed iterations (outside the event bracket)
            a = torch.cuda.Event(enable_timing=True)
            e = torch.cuda.Event(enable_timing=True)
            a.record()
            loss = step(x, y)
            e.record()
            evs.append((a, e))
        barrier()
        _Profiler.enabled = False
        launches = lib.maai_launch_count() - launches0
        ms = [a.elapsed_time(e) for a, e in evs]
        total_ms = sum(ms)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t)
        spans = {k: [a.elapsed_time(e) for a, e in v] for k, v in _Profiler.events.items()}
        return dict(ms_per_step=total_ms / steps, ms_median=statistics.median(ms), launches=launches,
                    spans=spans, loss=float(loss.detach()))

    def timed_graphed(bb, steps, warmup):
        """same step through maai_b200.GraphedNTXentLoss (forward and backward as one CUDA graph each)"""
        h1, h2 = make_inputs(bb)
        x = h1.requires_grad_(True)
        y = h2.requires_grad_(True)
        fn = maai_b200.GraphedNTXentLoss(bb, d, tau, dtype=x.dtype, device=dev, hidden1_requires_grad=True)

        def gstep():
            x.grad = y.grad = None
            loss = fn(x, y)
            loss.backward()
            return loss
        for _ in range(warmup):
            gstep()
        torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            flush_buf.fill_(1)
            a = torch.cuda.Event(enable_timing=True)
            e = torch.cuda.Event(enable_timing=True)
            a.record()
            loss = gstep()
            e.record()
            evs.append((a, e))
        torch.cuda.synchronize()
        ms = [a.elapsed_time(e) for a, e in evs]
        return dict(ms_per_step=sum(ms) / steps, loss=float(loss.detach()))

    # ---------------- main timed region (device-resident inputs) ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_r = timed_run(b, args.steps, args.warmup, profile=True)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- end to end through the public API with HOST buffers ----------------
    # Every step copies ITS inputs pinned host -> device and ITS results (loss, dh1, dh2) device ->
    # pinned host.  "pipelined": the copies run on a copy stream, double-buffered, so the H2D of step
    # i+1 and the D2H of step i-1 overlap the kernels of step i (what an input pipeline with prefetch
    # does); "serial": one stream, copy -> compute -> copy -> host sync per step.
    g = torch.Generator().manual_seed(1234 + rank)
    hp1 = torch.randn(b, d, generator=g).pin_memory()
    hp2 = torch.randn(b, d, generator=g).pin_memory()
    out_g1 = [torch.empty(b, d).pin_memory() for _ in range(2)]
    out_g2 = [torch.empty(b, d).pin_memory() for _ in range(2)]
    out_loss = [torch.empty(()).pin_memory() for _ in range(2)]

    def e2e_serial_step():
        x = hp1.to(dev, non_blocking=True).requires_grad_(True)
        y = hp2.to(dev, non_blocking=True).requires_grad_(True)
        loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, local_rank=rank, world_size=world,
                                                device=dev)
        loss.backward()
        out_loss[0].copy_(loss.detach(), non_blocking=True)
        out_g1[0].copy_(x.grad, non_blocking=True)
        out_g2[0].copy_(y.grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the host results every step

    copy_s = torch.cuda.Stream(device=dev)
    main_s = torch.cuda.current_stream()

    def h2d():
        """inputs of one step on the copy stream; returns (x, y, event)"""
        with torch.cuda.stream(copy_s):
            x = hp1.to(dev, non_blocking=True)
            y = hp2.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_s)
        return x, y, ev

    def e2e_pipelined(steps):
        nxt = h2d()
        pending = []  # (event, slot) of result copies in flight
        for i in range(steps):
            x, y, ev = nxt
            main_s.wait_event(ev)
            x.record_stream(main_s)
            y.record_stream(main_s)
            if i + 1 < steps:
                nxt = h2d()  # prefetch the next step's inputs while this step computes
            x.requires_grad_(True)
            y.requires_grad_(True)
            loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, local_rank=rank, world_size=world,
                                                    device=dev)
            loss.backward()
            done = torch.cuda.Event()
            done.record(main_s)
            slot = i & 1
            if len(pending) == 2:  # the host buffers of this slot must have been read back
                pending.pop(0)[0].synchronize()
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(done)
                lg, g1, g2 = loss.detach(), x.grad, y.grad
                for tns in (lg, g1, g2):
                    tns.record_stream(copy_s)
                out_loss[slot].copy_(lg, non_blocking=True)
                out_g1[slot].copy_(g1, non_blocking=True)
                out_g2[slot].copy_(g2, non_blocking=True)
                fin = torch.cuda.Event()
                fin.record(copy_s)
            pending.append((fin, slot))
        for fin, _ in pending:
            fin.synchronize()

    def timed_e2e(fn_steps):
        barrier()
        t0 = time.perf_counter()
        fn_steps(args.steps)
        barrier()
        sec = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([sec], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = float(tt)
        return sec

    for _ in range(args.warmup):
        e2e_serial_step()
    e2e_pipelined(args.warmup)
    e2e_serial_s = timed_e2e(lambda k: [e2e_serial_step() for _ in range(k)])
    e2e_pipe_s = timed_e2e(e2e_pipelined)
    # headline = the faster schedule (with several ranks on one host the second copy stream can lose:
    # N=2 measured 3.4 ms pipelined vs 2.0 ms serial)
    e2e_s = min(e2e_pipe_s, e2e_serial_s)
    e2e_mode = "pipelined" if e2e_pipe_s <= e2e_serial_s else "serial"
    e2e_pairs = B * args.steps / e2e_s
    h2d_bytes = 2 * b * d * 4
    d2h = 2 * b * d * 4 + 4

    # ---------------- secondary: configs[1] (4096 pairs, 1 GPU) ----------------
    secondary = None
    if world == 1 and not args.no_secondary and B != 4096:
        s = timed_run(4096, max(args.steps, 50), args.warmup, profile=False)
        secondary = {"workload": workload_name(4096, d, tau), "ms_per_step": s["ms_per_step"],
                     "pairs_per_s": 4096 / (s["ms_per_step"] * 1e-3),
                     "frac_bf16_peak": 24.0 * 4096 ** 2 * d / (s["ms_per_step"] * 1e-3) / (peaks["bf16"] * 1e12)}
        sg = timed_graphed(4096, max(args.steps, 50), args.warmup)
        secondary["cuda_graph_ms_per_step"] = sg["ms_per_step"]
        secondary["cuda_graph_pairs_per_s"] = 4096 / (sg["ms_per_step"] * 1e-3)

    # ---------------- secondary: the reference's training call, hidden1 detached ----------------
    # (Contrastive_Learning.py:685-690 passes hidden1=outputs1.data: only dh2 is needed, half of the backward)
    detached = None
    if not args.no_secondary:
        s2 = timed_run(b, max(args.steps // 2, 20), args.warmup, profile=False, grad1=False)
        detached = {"workload": workload_name(B, d, tau).replace("both inputs require grad", "hidden1 detached (training call)"),
                    "ms_per_step": s2["ms_per_step"], "pairs_per_s": B / (s2["ms_per_step"] * 1e-3)}

    # ---------------- secondary: configs[0] (256 pairs, the reference's own CPU-runnable case) ----------------
    configs0 = None
    if rank == 0 and world == 1 and not args.no_secondary and not args.no_cpu_baseline:
        from oracle.cpu_baseline import time_port_full
        s0 = timed_run(256, max(args.steps, 100), args.warmup, profile=False)
        c0 = time_port_full(256, d, tau, steps=30, warmup=5)
        g0 = timed_graphed(256, max(args.steps, 100), args.warmup)
        configs0 = {"workload": workload_name(256, d, tau), "ms_per_step": s0["ms_per_step"],
                    "pairs_per_s": 256 / (s0["ms_per_step"] * 1e-3), "loss": s0["loss"],
                    "cuda_graph_ms_per_step": g0["ms_per_step"], "cuda_graph_pairs_per_s": 256 / (g0["ms_per_step"] * 1e-3),
                    "cpu_reference_port": {"ms_per_step": c0["s_per_step"] * 1e3, "pairs_per_s": c0["pairs_per_s"],
                                           "cores": c0["threads"], "loss": c0["loss"]},
                    "note": "host-launch-bound on the GPU (9 launches from Python per step); different random "
                            "draws on the two devices, parity at this shape is tests/test_gpu_parity.py"}

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_baseline import time_port_stripe
        r = time_port_stripe(B, d, tau, 256, steps=10, warmup=2)
        cpu = {"value": r["pairs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"10 steps of a stripe of {r['b_sample']} anchor pairs x all {B} global keys, fwd+bwd, "
                         f"torch {torch.__version__} fp32 CPU ({r['s_per_step']:.3f} s/step)"}

    if rank == 0:
        ms = main_r["ms_per_step"]
        value = B / (ms * 1e-3)
        # dominant kernel: ntxent_tile_kernel<D, BWD> (recompute S + P.Z): algorithmic 16*b*B*d flops per
        # launch (16 B^2 d of the 24 B^2 d per step are backward, split over the ranks)
        bwd_ms = statistics.mean(main_r["spans"]["bwd"]) if main_r["spans"]["bwd"] else None
        fwd_ms = statistics.mean(main_r["spans"]["fwd"]) if main_r["spans"]["fwd"] else None
        flops_bwd = 16.0 * b * B * d
        achieved = flops_bwd / (bwd_ms * 1e-3) / 1e12 if bwd_ms else None
        step_tflops = 24.0 * B * B * d / (ms * 1e-3) / 1e12 / world
        traffic = None
        try:  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tj = json.load(f).get(f"pairs={B},dim={d},n_gpus={world}", {})
            traffic = next((v for k, v in tj.items() if "BWD" in k), None)
        except Exception:
            traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(B, d, tau), "pairs_global": B, "pairs_per_gpu": b,
                       "dim": d, "temperature": tau,
                       "parallelism": f"dp{world}: anchor rows sharded, bf16 all-gather of z, fp32 all-gather of row factors"
                                      + ((" -- both fused into the producing kernels as NVLink peer stores + symmetric-memory barrier"
                                          if peer_mode_used() else " -- NCCL all_gather_into_tensor") if world > 1 else ""),
                       "l2": "flushed between timed steps (256 MiB write outside the event bracket)",
                       "step_tflops_per_gpu_algorithmic": step_tflops,
                       "step_frac_bf16_peak": step_tflops / peaks["bf16"],
                       # executed flops: a single rank runs the symmetric forward (half the forward's MMAs)
                       "step_tflops_per_gpu_executed": step_tflops * ((20.0 / 24.0) if (world == 1 and maai_b200.padded_dim(d) <= 128) else 1.0),
                       "ms_median": main_r["ms_median"], "loss": main_r["loss"],
                       "fwd_call_ms": fwd_ms, "bwd_call_ms": bwd_ms,
                       "span_ms_mean": {k: (statistics.mean(v) if v else None) for k, v in main_r["spans"].items()}},
            "clocks": clocks,
            "e2e": {"value": e2e_pairs, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "schedule": e2e_mode,
                    "pipelined_ms_per_step": e2e_pipe_s / args.steps * 1e3,
                    "serial_value": B * args.steps / e2e_serial_s,
                    "serial_ms_per_step": e2e_serial_s / args.steps * 1e3,
                    "what": "every step: pinned host h1,h2 -> device, contrastive_loss + backward, loss + dh1 + dh2 -> "
                            "pinned host; wall clock over all steps.  Two schedules are timed and value is the faster one "
                            "(schedule): pipelined = copies on a second stream, double-buffered (H2D of step i+1 / D2H of "
                            "step i-1 overlap the kernels of step i); serial = one stream, host sync after every step"},
            "gpu_launches": int(main_r["launches"]),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["bf16"]) if achieved else None, "traffic": traffic,
                         "traffic_note": "dram read+write bytes per launch, ncu --set full (profiles/r1_ncu_summary_v10.md): one pass "
                                         "over z (16.8 MB) + the fp32 accumulator lines the atomics touch (33.5 MB); the "
                                         "2Bx2B logits never reach HBM (the reference moves ~100*b*B bytes)",
                         "kernel": f"ntxent_tile_kernel<D={maai_b200.padded_dim(d)},BWD,NQ=1>",
                         "how": "16*b*B*d algorithmic flops / mean CUDA-event time of the maai_ntxent_bwd call "
                                "(memset + tile kernel + dh kernel) over the timed steps",
                         "peak_source": peaks["source"] + ", burst cuBLAS bf16",
                         "peak_sustained": peaks["bf16_sustained"]},
            "cpu_baseline": cpu,
        }
        if secondary:
            line["config"]["configs1_4096_pairs"] = secondary
        if detached:
            line["config"]["hidden1_detached"] = detached
        if configs0:
            line["config"]["configs0_256_pairs"] = configs0
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
