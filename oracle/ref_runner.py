"""Runs the UNMODIFIED reference file (a build-time copy of /root/reference/SimCLR/Objective.py under
oracle/_ref/, git-ignored, made by __graft_entry__.build() in the container where the reference exists;
it travels to the GPU box like a built .so) -- BENCH / TEST INFRASTRUCTURE, NOT PRODUCT.

Only bench.py (cpu_baseline leg, --impl reference, the torch-GPU-reference secondary) and tests/ import
this module.  When oracle/_ref/Objective.py is absent every entry point falls back to the torch port
(oracle/ntxent_torch_port.py, bit-identical to the reference on the golden cases) and says so in `kind`.

Bounded sample of the 32768-pair workload.  The reference's single-process formulation needs ~100 B^2
bytes (107 GB at 32768 pairs) and cannot run that size at all; what does run is one RANK's share of it:
``contrastive_loss(hidden1[:b], hidden2[:b], world_size=B/b, local_rank=0)`` -- the reference's own
world_size > 1 branch (Objective.py:51-58), b anchor pairs against all B gathered keys, forward + backward.
Its ``dist.all_gather`` (Objective.py:113) is the one thing replaced: a local stand-in fills the tensor
list from pre-normalised constant keys (there are no peer processes on the host); the reference's code is
executed as is, including the per-step one_hot label/mask construction and their ``.to(device)``."""
from __future__ import annotations

import importlib.util
import os
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = os.path.join(HERE, "_ref", "Objective.py")


def load_reference():
    """The reference module, or None when the build-time copy does not exist."""
    if not os.path.exists(REF_FILE):
        return None
    spec = importlib.util.spec_from_file_location("maai_reference_objective", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_REF = None


def reference_loss(hidden1, hidden2, temperature, rank=0, world=1):
    """The reference's loss value (autograd-connected) for harnesses that compare whole training steps
    (tools/simclr_step.py, tools/convergence_parity.py): the unmodified reference file when the build-time
    copy exists -- its own ``dist.all_gather`` branch for world > 1 -- else the torch port of this package
    (keys gathered with ``dist.all_gather`` like the reference does).  The ONE place outside tests where the
    reference's formulation is spelled out is oracle/ntxent_torch_port.py."""
    global _REF
    if _REF is None:
        _REF = load_reference() or False
    h1, h2 = hidden1.float(), hidden2.float()
    if _REF:
        return _REF.contrastive_loss(h1, h2, temperature=temperature, local_rank=rank, world_size=world,
                                     device=h1.device)[0]
    from .ntxent_torch_port import ntxent_port
    if world == 1:
        return ntxent_port(h1, h2, temperature)
    import torch.distributed as dist

    def gather(t):
        outs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return torch.cat(outs, 0)
    k1 = gather(torch.nn.functional.normalize(h1.detach(), dim=1))
    k2 = gather(torch.nn.functional.normalize(h2.detach(), dim=1))
    return ntxent_port(h1, h2, temperature, keys1=k1, keys2=k2, rank=rank)


class _LocalGather:
    """Stand-in for torch.distributed inside the reference module: all_gather(list, tensor) fills the list
    with the chunks of pre-normalised keys (alternating view a / view b, the order of Objective.py:52-53)."""

    def __init__(self, keys1, keys2, b):
        self.keys = (keys1, keys2)
        self.b = b
        self.calls = 0

    def all_gather(self, tensor_list, tensor):
        k = self.keys[self.calls % 2]
        self.calls += 1
        for i, t in enumerate(tensor_list):
            t.copy_(k[i * self.b:(i + 1) * self.b])


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()


def time_reference_stripe(pairs_global, dim, tau, b_sample, steps, warmup, threads=None, seed=1234, device="cpu"):
    """One rank's share (b_sample pairs x all keys) of the pairs_global workload, fwd + bwd, through the
    reference's own world_size > 1 branch.  Returns dict(pairs_per_s, s_per_step, kind, ...)."""
    ref = load_reference()
    threads = threads or os.cpu_count() or 1
    if torch.device(device).type == "cpu":
        torch.set_num_threads(threads)
    if ref is None:
        from .cpu_baseline import time_port_stripe
        r = time_port_stripe(pairs_global, dim, tau, b_sample, steps, warmup, threads, seed)
        r["kind"] = "port"
        return r
    g = torch.Generator().manual_seed(seed)
    H1 = torch.randn(pairs_global, dim, generator=g)
    H2 = torch.randn(pairs_global, dim, generator=g)
    b = min(b_sample, pairs_global)
    while pairs_global % b:
        b -= 1
    world = pairs_global // b
    K1 = torch.nn.functional.normalize(H1, dim=1).to(device)
    K2 = torch.nn.functional.normalize(H2, dim=1).to(device)
    x = H1[:b].clone().to(device).requires_grad_(True)
    y = H2[:b].clone().to(device).requires_grad_(True)
    saved = ref.dist
    ref.dist = _LocalGather(K1, K2, b)
    times, loss = [], None
    try:
        for i in range(warmup + steps):
            x.grad = None
            y.grad = None
            _sync(device)
            t0 = time.perf_counter()
            loss, _, _ = ref.contrastive_loss(x, y, temperature=tau, local_rank=0, world_size=world, device=device)
            loss.backward()
            _sync(device)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    finally:
        ref.dist = saved
    mean = sum(times) / len(times)
    return dict(pairs_per_s=b / mean, s_per_step=mean, b_sample=b, threads=threads, loss=float(loss.detach()),
                steps=steps, warmup=warmup, kind="reference", world_emulated=world)


def time_reference_full(pairs, dim, tau, steps, warmup, threads=None, seed=1234, device="cpu", grad1=True):
    """The reference's single-process call on a WHOLE batch, fwd + bwd (BASELINE.json configs[0] on CPU,
    configs[1] "vs reference PyTorch loss" with device='cuda')."""
    ref = load_reference()
    threads = threads or os.cpu_count() or 1
    if torch.device(device).type == "cpu":
        torch.set_num_threads(threads)
    if ref is None:
        if torch.device(device).type != "cpu":
            return None
        from .cpu_baseline import time_port_full
        r = time_port_full(pairs, dim, tau, steps, warmup, threads, seed)
        r["kind"] = "port"
        return r
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(pairs, dim, generator=g).to(device).requires_grad_(grad1)
    y = torch.randn(pairs, dim, generator=g).to(device).requires_grad_(True)
    times, loss = [], None
    for i in range(warmup + steps):
        x.grad = None
        y.grad = None
        _sync(device)
        t0 = time.perf_counter()
        loss, _, _ = ref.contrastive_loss(x, y, temperature=tau, device=device)
        loss.backward()
        _sync(device)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return dict(pairs_per_s=pairs / med, s_per_step=med, threads=threads, loss=float(loss.detach()), kind="reference",
                dx_norm=float(x.grad.norm()) if grad1 else None, dy_norm=float(y.grad.norm()))
