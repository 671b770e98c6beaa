"""CPU baseline timing of the reference's NT-Xent path -- BENCH INFRASTRUCTURE, NOT PRODUCT.

bench.py's ``cpu_baseline`` leg and ``--impl reference`` call ``time_port_stripe``: the torch port
of the reference (oracle/ntxent_torch_port.py) is timed on the host cores on a BOUNDED sample of
the benchmark workload: a stripe of ``b_sample`` anchor pairs against all B global keys, forward +
backward.  Every anchor still meets all 2B-1 keys, so work per pair equals the full workload's
(the full B = 32768 problem needs ~100*B^2 bytes = 107 GB in the reference's formulation and
cannot run at all).  In the stripe the keys are constants, exactly like the reference's own
world_size > 1 branch (Objective.py:112-114), so the backward is the query-side gradient."""
from __future__ import annotations

import os
import time

import torch

from .ntxent_torch_port import ntxent_port


def time_port_stripe(pairs_global: int, dim: int, tau: float, b_sample: int, steps: int, warmup: int,
                     threads: int | None = None, seed: int = 1234):
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    H1 = torch.randn(pairs_global, dim, generator=g)
    H2 = torch.randn(pairs_global, dim, generator=g)
    K1 = torch.nn.functional.normalize(H1, dim=1)
    K2 = torch.nn.functional.normalize(H2, dim=1)
    b_sample = min(b_sample, pairs_global)
    x = H1[:b_sample].clone().requires_grad_(True)
    y = H2[:b_sample].clone().requires_grad_(True)
    times = []
    loss = None
    for i in range(warmup + steps):
        x.grad = None
        y.grad = None
        t0 = time.perf_counter()
        loss = ntxent_port(x, y, tau, keys1=K1, keys2=K2, rank=0)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return dict(pairs_per_s=b_sample / mean, s_per_step=mean, b_sample=b_sample, threads=threads,
                loss=float(loss.detach()), steps=steps, warmup=warmup)


def time_port_full(pairs: int, dim: int, tau: float, steps: int, warmup: int, threads: int | None = None,
                   seed: int = 1234):
    """The reference's single-process call on a WHOLE (small) batch, forward + backward, both inputs
    requiring grad (BASELINE.json configs[0]: 256 pairs, d=128, tau=0.5 on CPU)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(pairs, dim, generator=g).requires_grad_(True)
    y = torch.randn(pairs, dim, generator=g).requires_grad_(True)
    times = []
    for i in range(warmup + steps):
        x.grad = None
        y.grad = None
        t0 = time.perf_counter()
        loss = ntxent_port(x, y, tau)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return dict(pairs_per_s=pairs / med, s_per_step=med, threads=threads, loss=float(loss.detach()))
