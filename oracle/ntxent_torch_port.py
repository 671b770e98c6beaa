"""Torch-CPU port of the reference NT-Xent path -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

Used only by ``bench.py`` (``cpu_baseline`` and ``--impl reference``) and ``tests/``.  The
reference is a Python file that cannot travel to the GPU box (/root/reference does not exist
there), so this module restates its op sequence in fp32 torch so that the CPU timing is
representative of what the reference executes on host cores:

  normalise (Objective.py:41-43) -> int64 one-hot labels and masks built from a Python range
  (Objective.py:62-65) -> four matmuls divided by the temperature (Objective.py:67-74) ->
  ``- mask * 1e9`` on the two same-view blocks (Objective.py:68,71) -> concat + log_softmax +
  ``-(targets * logprobs).sum() / rows`` twice (Objective.py:76-77, 123-125) -> sum (Objective.py:79);
  backward is plain autograd, as in the reference (Contrastive_Learning.py:698).

It is validated against the imported reference by ``tests/golden/make_golden.py`` (bitwise-close
fp32 agreement recorded in the golden file) and against the fp64 oracle in ``tests/test_oracle.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_BIG = 1e9  # Objective.py:6


def _soft_ce(onehot: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
    # Objective.py:123-125
    lp = F.log_softmax(logits, dim=1)
    return -(onehot * lp).sum() / logits.shape[0]


def ntxent_port(h1: torch.Tensor, h2: torch.Tensor, temperature: float = 1.0,
                keys1: torch.Tensor | None = None, keys2: torch.Tensor | None = None,
                rank: int = 0) -> torch.Tensor:
    """Scalar loss with the reference's op sequence.  ``keys1/keys2`` (already normalised,
    (B,d)) emulate the gathered tensors of the world_size>1 branch; by default keys alias the
    local queries exactly like Objective.py:60-61."""
    if h1.shape != h2.shape:
        raise AssertionError("hidden1.shape != hidden2.shape")  # Objective.py:45
    a = F.normalize(h1, dim=1, p=2)
    c = F.normalize(h2, dim=1, p=2)
    n = a.shape[0]
    ka = a if keys1 is None else keys1
    kc = c if keys2 is None else keys2
    big = ka.shape[0]
    idx = torch.tensor(range(n)) + rank * n
    target = F.one_hot(idx, 2 * big).to(a.device)
    selfmask = F.one_hot(idx, big).to(a.device)
    t = temperature
    s_aa = torch.matmul(a, ka.t()) / t - selfmask * _BIG
    s_cc = torch.matmul(c, kc.t()) / t - selfmask * _BIG
    s_ac = torch.matmul(a, kc.t()) / t
    s_ca = torch.matmul(c, ka.t()) / t
    return _soft_ce(target, torch.cat([s_ac, s_aa], 1)) + _soft_ce(target, torch.cat([s_ca, s_cc], 1))


def ntxent_port_fwd_bwd(h1: torch.Tensor, h2: torch.Tensor, temperature: float):
    """One forward + backward with both inputs requiring grad.  Returns (loss, dh1, dh2)."""
    x = h1.detach().clone().requires_grad_(True)
    y = h2.detach().clone().requires_grad_(True)
    loss = ntxent_port(x, y, temperature)
    loss.backward()
    return loss.detach(), x.grad, y.grad
