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

Before anything is timed, a PARITY GATE runs at every N on the bench shape itself: one step of the
product path (the same gather mode / forward schedule the timed region uses), its per-rank loss and
the dH rows of sampled pairs of every rank against the fp64 oracle (oracle/ntxent_oracle.py
ntxent_rows_oracle; the all-row softmax denominators the key-side term needs come from plain torch
fp32 ops, validated against fp64 on the sampled rows).  Errors are all-reduced (MAX) and reported
under "parity"; above the north-star tolerances (loss 1e-3, dH 1e-2) the run exits non-zero.  The
oracle is used there only as the checker.

Prints ONE JSON line on rank 0 (see the keys below).  ``--impl reference`` times the reference's
own CPU implementation of the path -- the UNMODIFIED reference file (oracle/_ref/Objective.py, a
build-time copy; the torch port of oracle/ if that copy is absent), all host threads -- on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import gc
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
    """SM clock, power and throttle reasons while the timed region runs: the recipe's own
    ``nvidia-smi --query-gpu=... -lms`` as a SEPARATE process (in-process NVML queries from a second thread
    take the driver lock of the launching process: measured multi-millisecond stalls of the launch path at
    N > 1).  A reader thread stamps every line on arrival; ``stop()`` reports the samples that arrived inside
    the timed region.  Runs shorter than the sampler's period (~25 ms per query: 20 steps at 8 GPUs take 8 ms)
    get their under-load samples from an untimed continuation of the same loop (``need_more`` / bench.py's
    timed_run), and say so."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=20):
        self.index = index
        self.period_ms = period_ms
        self.samples = []   # (sm MHz, W, [reasons], host time of arrival)
        self.proc = None
        self.thread = None
        self.err = None
        self.sm_max = None
        self.t0 = self.t1 = None

    def _reader(self):
        try:
            for ln in self.proc.stdout:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                try:
                    rs = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7])
                          if v.lower().startswith("active")]
                    self.samples.append((float(f[0]), float(f[2]), rs, time.time()))
                    self.sm_max = float(f[1])
                except ValueError:
                    continue
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        import shutil
        exe = shutil.which("nvidia-smi")
        if not exe:
            self.err = "nvidia-smi not found"
            return
        try:
            self.proc = subprocess.Popen([exe, f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._reader, daemon=True)
        self.thread.start()
        t = time.time()
        while not self.samples and time.time() - t < 3.0:  # the process is up and reporting
            time.sleep(0.01)

    def mark(self):
        """the timed region starts now"""
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def inside(self):
        return [x for x in self.samples if self.t0 is not None and self.t0 <= x[3] <= (self.t1 or time.time())]

    def need_more(self, t_since):
        """no sample has arrived since t_since yet (and the sampler works at all)"""
        return self.proc is not None and self.samples and not any(x[3] >= t_since for x in self.samples)

    def stop(self, continuation_from=None):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:  # noqa: BLE001
                self.proc.kill()
            if self.thread:
                self.thread.join(timeout=2)
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[f"no samples ({self.err})"])
        inside = self.inside()
        where = "inside the timed region"
        use = inside
        if not use and continuation_from is not None:
            use = [x for x in self.samples if x[3] >= continuation_from]
            where = ("the timed region is shorter than the sampler's period: sampled during an untimed continuation of "
                     "the same loop, right after it")
        if not use:
            use, where = self.samples, "around the timed region"
        sm = [x[0] for x in use]
        reasons = sorted({r for x in use for r in x[2]})
        return dict(sm_mhz=statistics.median(sm), sm_min_mhz=min(sm), sm_max_mhz=self.sm_max,
                    power_w_max=max(x[1] for x in use), samples=len(use), sampled=where, reasons=reasons,
                    how="nvidia-smi -lms %d (separate process)" % self.period_ms)


def run_reference(args, rank):
    """CPU arm: the reference's own implementation on all host threads (oracle/ref_runner.py)."""
    if rank != 0:
        return
    from oracle.ref_runner import time_reference_stripe
    import torch
    b_sample = 256
    r = time_reference_stripe(PAIRS, DIM, TAU, b_sample, steps=args.steps, warmup=max(args.warmup, 2))
    sample = reference_sample_text(r, PAIRS, torch.__version__)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 2),
        "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(PAIRS, DIM, TAU), "pairs_global": PAIRS, "dim": DIM,
                   "temperature": TAU},
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": r["kind"],
                         "sample": sample},
        "e2e": {"value": r["pairs_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def reference_sample_text(r, pairs, torch_version):
    if r["kind"] == "reference":
        return (f"UNMODIFIED reference SimCLR/Objective.py::contrastive_loss, one rank's share of the {pairs}-pair "
                f"workload per step: {r['b_sample']} anchor pairs x all {pairs} gathered keys through its own "
                f"world_size={r['world_emulated']} branch (Objective.py:51-58), fwd+bwd; dist.all_gather replaced by a "
                f"local fill from constant keys; torch {torch_version} fp32 CPU, {r['s_per_step']:.3f} s/step")
    return (f"torch port of the reference (oracle/_ref absent): stripe of {r['b_sample']} anchor pairs x all {pairs} "
            f"global keys per step, fwd+bwd, torch {torch_version} fp32 CPU, {r['s_per_step']:.3f} s/step")


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
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate (profiling runs only)")
    ap.add_argument("--parity-pairs", type=int, default=256, help="sampled pairs over all ranks in the parity gate")
    ap.add_argument("--require-peer", action="store_true",
                    help="N > 1: fail unless the fused NVLink peer-store gathers are the mode that ran")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
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

    from maai_b200 import Objective as P

    B, d, tau = args.pairs, args.dim, args.tau
    assert B % world == 0
    b = B // world
    dp = maai_b200.padded_dim(d)

    def path_used(bb):
        """Which dataflow the public API takes for bb pairs per rank (decided at the first call per shape)."""
        if world == 1:
            return dict(gather_mode="single rank", sym_forward=bool(lib.maai_ntxent_fwd_is_symmetric(bb, 1, dp)))
        # the same (cached, collectively voted) decision contrastive_loss takes at its first call with this shape
        peer = bool(P.peer_gather_available() and P._peer_usable(bb, dp, world, rank, dev, None))
        if not peer:
            return dict(gather_mode="nccl all_gather_into_tensor", sym_forward=False)
        ws = P.PeerWorkspace.get(bb, dp, world, rank, dev, None)
        return dict(gather_mode="peer stores, NVSwitch multicast" if ws.mc_z[0] else "peer stores, unicast",
                    sym_forward=P._sym_forward_mode(bb, dp, world, ws.use_flags) != "off",
                    sym_forward_mode=P._sym_forward_mode(bb, dp, world, ws.use_flags), peer_order="in-kernel flags" if ws.use_flags else "barrier launches")

    def make_inputs(bb, r=None):
        g = torch.Generator(device=dev).manual_seed(1234 + (rank if r is None else r))
        return (torch.randn(bb, d, generator=g, device=dev), torch.randn(bb, d, generator=g, device=dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(x, y):
        x.grad = None
        y.grad = None
        loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, local_rank=rank, world_size=world,
                                                device=dev, key_grad=True)
        loss.backward()
        return loss

    # ---------------- parity gate: the bench shape, every N, before anything is timed ----------------
    def parity_gate():
        from oracle import ntxent_oracle as O          # the checker, never the thing measured
        from oracle.large_batch import den_all_torch
        parts = [make_inputs(b, r) for r in range(world)]   # every rank regenerates the global batch
        H1 = torch.cat([p[0] for p in parts])
        H2 = torch.cat([p[1] for p in parts])
        den_all = den_all_torch(H1, H2, tau).cpu().numpy()
        H1c, H2c = H1.cpu().numpy(), H2.cpu().numpy()
        n = max(16, args.parity_pairs // world)
        rng = np.random.default_rng(99 + rank)
        pairs = np.unique(np.concatenate([rng.choice(b, min(n, b), replace=False), [0, b - 1]]))
        ref = O.ntxent_rows_oracle(H1c, H2c, pairs, tau, rank=rank, world=world, den_all=den_all)
        own = rank * b + pairs
        den_err = max(np.abs(den_all[own] / ref["den"][0] - 1).max(), np.abs(den_all[B + own] / ref["den"][1] - 1).max())
        loss_ref = O.loss_from_denominators(H1c, H2c, den_all, tau, rank, world)
        x = parts[rank][0].clone().requires_grad_(True)
        y = parts[rank][1].clone().requires_grad_(True)
        loss = step(x, y)
        torch.cuda.synchronize()
        g1, g2 = x.grad[pairs].double().cpu().numpy(), y.grad[pairs].double().cpu().numpy()
        fro = lambda a, r_: float(np.linalg.norm(a - r_) / np.linalg.norm(r_))
        mx = lambda a, r_: float(np.abs(a - r_).max() / np.abs(r_).max())
        e = allmax([abs(float(loss) - loss_ref) / abs(loss_ref), max(fro(g1, ref["dh1"]), fro(g2, ref["dh2"])),
                    max(mx(g1, ref["dh1"]), mx(g2, ref["dh2"])), den_err,
                    0.0 if bool(torch.isfinite(x.grad).all() and torch.isfinite(y.grad).all()) else 1.0])
        out = {"loss_rel": e[0], "dh_rel_fro": e[1], "dh_rel_max": e[2], "rows": int(2 * len(pairs) * world),
               "pairs_per_rank": int(len(pairs)), "ranks": world, "reduction": "max over ranks",
               "torch_fp32_denominators_vs_fp64": e[3], "all_finite": e[4] == 0.0,
               "oracle": "oracle/ntxent_oracle.py::ntxent_rows_oracle (fp64, sampled pairs of every rank x all keys) "
                         "+ loss_from_denominators", "tolerance": {"loss_rel": 1e-3, "dh_rel_fro": 1e-2},
               "loss": float(loss), "loss_oracle_rank0": loss_ref if rank == 0 else None}
        out.update(path_used(b))
        out["ok"] = bool(e[0] <= 1e-3 and e[1] <= 1e-2 and e[2] <= 2e-2 and e[3] <= 1e-4 and e[4] == 0.0)
        return out

    parity = None
    if not args.no_parity:
        parity = parity_gate()
        if not parity["ok"]:
            if rank == 0:
                sys.stderr.write("PARITY GATE FAILED: " + json.dumps(parity) + "\n")
                emit({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "parity": parity,
                      "error": "parity gate failed; nothing was timed"})
            raise SystemExit(3)
    if args.require_peer and world > 1 and not path_used(b)["gather_mode"].startswith("peer"):
        raise SystemExit("--require-peer: the fused peer-store gathers are not the mode that ran: "
                         + json.dumps(path_used(b)))

    def warm_up(fn, warmup):
        """2 W untimed steps, then more of them until the GPU has been busy for ~10 ms in total: after the parity
        gate's CPU work the SM clock sits at idle (120 MHz) and W = 3..5 steps of 0.4 ms (8 GPUs) end before it
        has ramped up -- the first timed steps then run slow (step maxima 0.57 vs 0.40 ms median).  Not longer:
        warm-up spends the box's power budget and the "short" run then measures the capped clock (2 GPUs after
        0.15 s of warm-up: 1.28 -> 1.42 ms at 1680 MHz, sw_power_cap; 1 GPU after 35 ms: 2.35 -> 2.47 ms) --
        that regime is what the 200-step run reports.  At 1 GPU (2.3 ms per step) nothing is added.  Every rank
        runs the same number of steps (rank 0 decides): they depend on each other."""
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        per = max(1e-5, (time.perf_counter() - t0) / warmup)
        n = torch.tensor([max(0, min(500, int(0.010 / per) - 2 * warmup))], device=dev, dtype=torch.int64)
        if world > 1:
            dist.broadcast(n, 0)
        for _ in range(int(n)):
            fn()
        barrier()

    def timed_run(bb, steps, warmup, profile, grad1=True, sample_clocks=False):
        h1, h2 = make_inputs(bb)
        x = h1.requires_grad_(grad1)
        y = h2.requires_grad_(True)
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()  # a separate process: its start-up overlaps the warm-up steps
        warm_up(lambda: step(x, y), warmup)
        _Profiler.reset()
        _Profiler.enabled = profile
        if sampler:
            sampler.mark()
        launches0 = lib.maai_launch_count()
        evs = []
        gc.disable()  # a collection pause on one rank is a stall of every rank
        t_host = time.perf_counter()
        for _ in range(steps):
            flush_buf.fill_(1)  # flush L2 between timed iterations (outside the event bracket)
            a = torch.cuda.Event(enable_timing=True)
            e = torch.cuda.Event(enable_timing=True)
            a.record()
            loss = step(x, y)
            e.record()
            evs.append((a, e))
        host_issue_ms = (time.perf_counter() - t_host) * 1e3 / steps  # how fast this rank's host enqueues a step
        barrier()
        gc.enable()
        _Profiler.enabled = False
        clocks = None
        launches_timed = lib.maai_launch_count() - launches0
        if sample_clocks:
            # A run shorter than the sampler's period (20 steps at 8 GPUs take 8 ms) has no sample inside it:
            # every rank then keeps the same load going, untimed, for ~80 ms (rank 0 decides, all ranks follow:
            # the steps depend on each other across ranks) and the clocks are read there.
            n_extra = 0
            if sampler:
                sampler.mark_end()
                if not sampler.inside():
                    per_step = max(1e-5, (sampler.t1 - sampler.t0) / steps)
                    n_extra = min(2000, int(0.08 / per_step) + 1)
            if world > 1:
                t = torch.tensor([n_extra], device=dev, dtype=torch.int64)
                dist.broadcast(t, 0)
                n_extra = int(t)
            cont = time.time() if n_extra else None
            for _ in range(n_extra):
                step(x, y)
            barrier()
            if sampler:
                clocks = sampler.stop(cont)
        _Profiler.enabled = False
        launches = launches_timed if sample_clocks else lib.maai_launch_count() - launches0
        ms = [a.elapsed_time(e) for a, e in evs]
        total_ms = sum(ms)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t)
        spans = _Profiler.collect_ms()
        if profile and world > 1:  # every rank's own view (stderr): a rank that waits for a peer shows it here
            sys.stderr.write(f"[rank {rank}] step ms mean {sum(ms) / steps:.4f} median {statistics.median(ms):.4f} "
                             f"max {max(ms):.4f}; spans " +
                             json.dumps({k: round(statistics.mean(v), 4) for k, v in spans.items() if v}) + "\n")
        return dict(ms_per_step=total_ms / steps, ms_median=statistics.median(ms), launches=launches,
                    spans=spans, loss=float(loss.detach()), clocks=clocks, ms_max=max(ms), host_issue_ms=host_issue_ms)

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
        warm_up(gstep, warmup)
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
    # Two run lengths, the short one first: short runs see boost clocks (~1.9 GHz), 200 back-to-back steps run
    # into the power cap (MEASURED_PEAKS: burst vs sustained).  The line's value is the one --steps names; both
    # are reported with their clocks.  e2e is measured right after the run --steps names, i.e. in the same
    # clock regime as `value` (the other run length follows the e2e block when it is the long one).
    other_steps = 20 if args.steps >= 100 else 200
    other_r = None
    if args.steps >= 100:
        other_r = timed_run(b, other_steps, args.warmup, profile=False, sample_clocks=True)
    main_r = timed_run(b, args.steps, args.warmup, profile=True, sample_clocks=True)
    clocks = main_r["clocks"]

    # ---------------- end to end through the public API with HOST buffers ----------------
    # Every step copies ITS inputs pinned host -> device and ITS results (loss, dh1, dh2) device -> pinned
    # host.  One packed pinned input [h1|h2] -> ONE H2D copy per step on its own stream; the results go into
    # one packed pinned buffer [dh1|dh2|loss] on a third stream, so both DMA directions overlap the kernels
    # of the neighbouring steps (what an input pipeline with prefetch does).  "serial": one stream, copy ->
    # compute -> copy -> host sync per step.
    g = torch.Generator().manual_seed(1234 + rank)
    n_in = b * d
    h_in = torch.empty(2, b, d).pin_memory()
    h_in[0].copy_(torch.randn(b, d, generator=g))
    h_in[1].copy_(torch.randn(b, d, generator=g))
    NSLOT = 4  # depth of the host<->device pipeline (the host consumes slot i's results before reusing it at i+NSLOT)
    h_out = [torch.empty(2 * n_in + 1).pin_memory() for _ in range(NSLOT)]
    d_in = [torch.empty(2, b, d, device=dev) for _ in range(NSLOT)]

    def e2e_serial_step():
        d_in[0].copy_(h_in, non_blocking=True)
        x = d_in[0][0].detach().requires_grad_(True)
        y = d_in[0][1].detach().requires_grad_(True)
        loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, local_rank=rank, world_size=world,
                                                device=dev, key_grad=True)
        loss.backward()
        h_out[0][:n_in].copy_(x.grad.view(-1), non_blocking=True)
        h_out[0][n_in:2 * n_in].copy_(y.grad.view(-1), non_blocking=True)
        h_out[0][2 * n_in:].copy_(loss.detach().view(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the host results every step

    in_s = torch.cuda.Stream(device=dev)
    out_s = torch.cuda.Stream(device=dev)
    main_s = torch.cuda.current_stream()
    ev_in = [torch.cuda.Event() for _ in range(NSLOT)]      # inputs of the slot are on the device
    ev_done = [torch.cuda.Event() for _ in range(NSLOT)]    # results of the slot are computed
    ev_out = [torch.cuda.Event() for _ in range(NSLOT)]     # results of the slot are on the host

    e2e_host = {}

    def e2e_pipelined(steps):
        keep = [None] * NSLOT  # the slot's device results stay referenced until their D2H copy has been consumed
        e2e_host["t0"] = time.perf_counter()
        e2e_host["spans"] = []
        e2e_host["blocked_s"] = 0.0  # host time spent WAITING for the device (slot reuse), not issuing

        def h2d(i):
            s_ = i % NSLOT
            with torch.cuda.stream(in_s):
                if i >= NSLOT:
                    in_s.wait_event(ev_done[s_])  # the step that read this slot's inputs has finished
                d_in[s_].copy_(h_in, non_blocking=True)
                ev_in[s_].record(in_s)
        h2d(0)
        for i in range(steps):
            s_ = i % NSLOT
            if i + 1 < steps:
                h2d(i + 1)  # prefetch the next step's inputs while this step computes
            main_s.wait_event(ev_in[s_])
            x = d_in[s_][0].detach().requires_grad_(True)
            y = d_in[s_][1].detach().requires_grad_(True)
            c0 = torch.cuda.Event(enable_timing=True)
            c1 = torch.cuda.Event(enable_timing=True)
            c0.record(main_s)
            loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, local_rank=rank, world_size=world,
                                                    device=dev, key_grad=True)
            loss.backward()
            c1.record(main_s)
            e2e_host.setdefault("spans", []).append((c0, c1))
            ev_done[s_].record(main_s)
            if i >= NSLOT:
                tb = time.perf_counter()
                ev_out[s_].synchronize()  # the host has consumed this slot's previous results
                e2e_host["blocked_s"] += time.perf_counter() - tb
            keep[s_] = (loss, x, y)
            with torch.cuda.stream(out_s):
                out_s.wait_event(ev_done[s_])
                h_out[s_][:n_in].copy_(x.grad.view(-1), non_blocking=True)
                h_out[s_][n_in:2 * n_in].copy_(y.grad.view(-1), non_blocking=True)
                h_out[s_][2 * n_in:].copy_(loss.detach().view(1), non_blocking=True)
                ev_out[s_].record(out_s)
        e2e_host["issue_s"] = time.perf_counter() - e2e_host["t0"]
        for s_ in range(min(NSLOT, steps)):
            ev_out[s_].synchronize()

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

    def copies_only(steps):
        """the H2D + D2H traffic of the e2e schedule without any compute: the PCIe / host-memory floor of e2e"""
        dummy = torch.empty(2 * n_in + 1, device=dev)
        for i in range(steps):
            s_ = i % NSLOT
            with torch.cuda.stream(in_s):
                d_in[s_].copy_(h_in, non_blocking=True)
            with torch.cuda.stream(out_s):
                h_out[s_].copy_(dummy, non_blocking=True)
        in_s.synchronize()
        out_s.synchronize()

    e2e_pipelined(args.warmup)
    e2e_pipe_s = timed_e2e(e2e_pipelined)  # first: same clock regime as the device-resident run just before
    e2e_pipe_blocked_s = e2e_host.get("blocked_s", 0.0)
    e2e_compute_ms = statistics.mean(a.elapsed_time(e) for a, e in e2e_host["spans"]) if e2e_host.get("spans") else None
    for _ in range(args.warmup):
        e2e_serial_step()
    copies_only(args.warmup)
    copy_floor_s = timed_e2e(copies_only)
    e2e_serial_s = timed_e2e(lambda k: [e2e_serial_step() for _ in range(k)])
    e2e_loss = float(h_out[(args.steps - 1) % NSLOT][2 * n_in])
    e2e_s = min(e2e_pipe_s, e2e_serial_s)
    e2e_mode = "pipelined" if e2e_pipe_s <= e2e_serial_s else "serial"
    e2e_pairs = B * args.steps / e2e_s
    h2d_bytes = 2 * b * d * 4
    d2h = 2 * b * d * 4 + 4

    if other_r is None:
        other_r = timed_run(b, other_steps, args.warmup, profile=False, sample_clocks=True)

    # ---------------- secondary: configs[1] (4096 pairs, 1 GPU) ----------------
    secondary = None
    if world == 1 and not args.no_secondary and B != 4096:
        s = timed_run(4096, max(args.steps, 50), args.warmup, profile=False)
        secondary = {"workload": workload_name(4096, d, tau), "ms_per_step": s["ms_per_step"],
                     "pairs_per_s": 4096 / (s["ms_per_step"] * 1e-3), "gpu_launches_per_step": s["launches"] / max(args.steps, 50),
                     "frac_bf16_peak": 24.0 * 4096 ** 2 * d / (s["ms_per_step"] * 1e-3) / (peaks["bf16"] * 1e12)}
        sg = timed_graphed(4096, max(args.steps, 50), args.warmup)
        secondary["cuda_graph_ms_per_step"] = sg["ms_per_step"]
        secondary["cuda_graph_pairs_per_s"] = 4096 / (sg["ms_per_step"] * 1e-3)
        secondary["cuda_graph_frac_bf16_peak"] = 24.0 * 4096 ** 2 * d / (sg["ms_per_step"] * 1e-3) / (peaks["bf16"] * 1e12)
        # "vs reference PyTorch loss" (BASELINE.json configs[1]): the unmodified reference on this GPU
        try:
            from oracle.ref_runner import time_reference_full
            for nref in (4096, 16384):
                tr = time_reference_full(nref, d, tau, steps=5, warmup=2, device=dev)
                if tr is None:
                    break
                secondary[f"torch_gpu_reference_{nref}_pairs"] = {
                    "what": "UNMODIFIED reference contrastive_loss(device='cuda') fwd+bwd on this B200, fp32 (TF32 off), "
                            "host wall clock with a device sync per step, median of 5",
                    "ms_per_step": tr["s_per_step"] * 1e3, "pairs_per_s": tr["pairs_per_s"], "loss": tr["loss"]}
                torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001  (out of memory on a shared box must not kill the bench line)
            secondary["torch_gpu_reference_error"] = repr(ex)[:200]
        if secondary.get("torch_gpu_reference_4096_pairs"):
            secondary["speedup_vs_torch_gpu_reference"] = (secondary["torch_gpu_reference_4096_pairs"]["ms_per_step"]
                                                           / s["ms_per_step"])

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
        from oracle.ref_runner import time_reference_full
        s0 = timed_run(256, max(args.steps, 100), args.warmup, profile=False)
        c0 = time_reference_full(256, d, tau, steps=30, warmup=5)
        g0 = timed_graphed(256, max(args.steps, 100), args.warmup)
        configs0 = {"workload": workload_name(256, d, tau), "ms_per_step": s0["ms_per_step"],
                    "pairs_per_s": 256 / (s0["ms_per_step"] * 1e-3), "loss": s0["loss"],
                    "cuda_graph_ms_per_step": g0["ms_per_step"], "cuda_graph_pairs_per_s": 256 / (g0["ms_per_step"] * 1e-3),
                    "cpu_reference": {"kind": c0["kind"], "ms_per_step": c0["s_per_step"] * 1e3,
                                      "pairs_per_s": c0["pairs_per_s"], "cores": c0["threads"], "loss": c0["loss"]},
                    "note": "host-launch-bound on the GPU (4 launches from Python per step); different random "
                            "draws on the two devices, parity at this shape is tests/test_gpu_parity.py"}

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.ref_runner import time_reference_stripe
        r = time_reference_stripe(B, d, tau, 256, steps=10, warmup=2)
        cpu = {"value": r["pairs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": r["kind"],
               "sample": "10 steps; " + reference_sample_text(r, B, torch.__version__)}

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
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                tj = json.load(f).get(f"pairs={B},dim={d},n_gpus={world}", {})
            traffic = next((v for k, v in tj.items() if "BWD" in k), None)
        except Exception:
            traffic = None
        used = path_used(b)
        short, long_ = (other_r, main_r) if args.steps >= 100 else (main_r, other_r)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(B, d, tau), "pairs_global": B, "pairs_per_gpu": b,
                       "dim": d, "temperature": tau,
                       "parallelism": f"dp{world}: anchor rows sharded, bf16 gather of z, fp32 gather of row factors"
                                      + (f" -- {used['gather_mode']}" if world > 1 else ""),
                       "gather_mode": used["gather_mode"], "sym_forward": used["sym_forward"],
                       "l2": "flushed between timed steps (256 MiB write outside the event bracket)",
                       "step_tflops_per_gpu_algorithmic": step_tflops,
                       "step_frac_bf16_peak": step_tflops / peaks["bf16"],
                       # executed flops: the symmetric forward runs half the forward's MMAs
                       "step_tflops_per_gpu_executed": step_tflops * ((20.0 / 24.0) if used["sym_forward"] else 1.0),
                       "ms_median": main_r["ms_median"], "ms_max": main_r["ms_max"], "loss": main_r["loss"],
                       "host_issue_ms_per_step": main_r["host_issue_ms"],
                       "fwd_call_ms": fwd_ms, "bwd_call_ms": bwd_ms,
                       "span_ms_mean": {k: (statistics.mean(v) if v else None) for k, v in main_r["spans"].items()},
                       "run_lengths": {
                           "short": {"steps": min(args.steps, other_steps), "ms_per_step": short["ms_per_step"],
                                     "pairs_per_s": B / (short["ms_per_step"] * 1e-3), "clocks": short["clocks"]},
                           "sustained": {"steps": max(args.steps, other_steps), "ms_per_step": long_["ms_per_step"],
                                         "pairs_per_s": B / (long_["ms_per_step"] * 1e-3), "clocks": long_["clocks"],
                                         "frac_bf16_peak_sustained": (24.0 * B * B * d / (long_["ms_per_step"] * 1e-3) / 1e12 / world
                                                                      / peaks["bf16_sustained"]) if peaks["bf16_sustained"] else None}}},
            "clocks": clocks,
            "parity": parity,
            "e2e": {"value": e2e_pairs, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "schedule": e2e_mode, "loss_read_on_host": e2e_loss,
                    "pipelined_ms_per_step": e2e_pipe_s / args.steps * 1e3,
                    "pipelined_host_issue_ms_per_step": e2e_host.get("issue_s", 0.0) / args.steps * 1e3,
                    "pipelined_host_blocked_ms_per_step": e2e_pipe_blocked_s / args.steps * 1e3,
                    "pipelined_compute_span_ms": e2e_compute_ms,
                    "copies_only_ms_per_step": copy_floor_s / args.steps * 1e3,
                    "copies_only_note": "the same H2D + D2H bytes per step on the two copy streams with NO compute, all ranks "
                                        "at once: the PCIe / host-memory floor of e2e on this box (e2e >= max(this, ms_per_step))",
                    "serial_value": B * args.steps / e2e_serial_s,
                    "serial_ms_per_step": e2e_serial_s / args.steps * 1e3,
                    "what": "every step: ONE pinned host buffer [h1|h2] -> device, contrastive_loss + backward, "
                            "[dh1|dh2|loss] -> ONE pinned host buffer; wall clock over all steps, max over ranks.  Two "
                            "schedules are timed and value is the faster one (schedule): pipelined = H2D and D2H on their "
                            "own streams, four slots deep (H2D of step i+1 / D2H of step i-1 overlap the kernels of step "
                            "i); serial = one stream, host sync after every step"},
            "gpu_launches": int(main_r["launches"]),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["bf16"]) if achieved else None, "traffic": traffic,
                         "traffic_note": "dram read+write bytes per launch, ncu --set full (profiles/r2_ncu_summary_*.md): one pass "
                                         "over z + the fp32 accumulator lines the atomics touch; the "
                                         "2Bx2B logits never reach HBM (the reference moves ~100*b*B bytes)",
                         "kernel": f"ntxent_tile_kernel<D={dp},BWD,NQ=1>",
                         "how": "16*b*B*d algorithmic flops / mean CUDA-event time of the maai_ntxent_bwd call "
                                "(tile kernel + dh kernel) over the timed steps",
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
