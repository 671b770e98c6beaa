"""All-row softmax denominators at BASELINE.json's full sizes -- TEST INFRASTRUCTURE, NOT PRODUCT.

The sampled-anchor oracle (``ntxent_oracle.ntxent_rows_oracle``) needs, for the key-side gradient,
the softmax normaliser of EVERY global row (2B values, O(B^2 d) work: 1.1 TFLOP at 32768 pairs x
d=128, 8.8 TFLOP at 65536 x 256 -- minutes of fp64 numpy).  This module computes them with plain
PyTorch ops (fp32 ``torch.matmul`` with TF32 off + ``torch.exp``, fp64 row sums) on whatever device
the inputs live on: the "plain torch fp32 reference" of the same quantity.  It shares no code with
the CUDA library.  Callers validate its output against ``ntxent_oracle.row_denominators`` (fp64
numpy) on the sampled rows before using it (tests/test_gpu_large_batch.py, bench.py's parity gate).
Only ``tests/`` and ``bench.py``'s parity gate import it."""
from __future__ import annotations

import torch


def den_all_torch(H1: torch.Tensor, H2: torch.Tensor, temperature: float, block: int = 8192) -> torch.Tensor:
    """den_i = sum_{j != i} exp((z_i.z_j - 1)/tau) for the stacked rows [F.normalize(H1); F.normalize(H2)]
    (Objective.py:41-43, 67-77).  Returns (2B,) float64 on H1.device."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            Z = torch.cat([torch.nn.functional.normalize(H1.float(), dim=1),
                           torch.nn.functional.normalize(H2.float(), dim=1)], 0)
            M = Z.shape[0]
            out = torch.empty(M, dtype=torch.float64, device=Z.device)
            inv_tau = 1.0 / float(temperature)
            for r0 in range(0, M, block):
                r1 = min(M, r0 + block)
                s = torch.matmul(Z[r0:r1], Z.t())
                e = torch.exp((s - 1.0) * inv_tau)
                idx = torch.arange(r0, r1, device=Z.device)
                e[idx - r0, idx] = 0.0
                out[r0:r1] = e.sum(1, dtype=torch.float64)
            return out
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
