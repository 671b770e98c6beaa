"""Parity at BASELINE.json's own sizes (configs[2], configs[4]): 16384 / 32768 / 65536 pairs x
d in {64, 128, 256} x tau in {0.1, 0.5} on one GPU, symmetric forward forced on and off.

At these sizes the full fp64 oracle is minutes of CPU, so SAMPLED anchors are compared:
``oracle.ntxent_rows_oracle`` evaluates >= 256 pairs (both views: >= 512 anchor rows) against all 2B
keys in fp64 -- the row sums l', the loss terms and the dH rows.  The key-side half of the gradient
needs every row's softmax denominator; those come from a plain-torch fp32 computation on the GPU
(oracle/large_batch.py, shares nothing with the CUDA library) that is first checked against the
fp64 values on the sampled rows.  The scalar loss is compared with the loss rebuilt from those
denominators.  Tolerances are the north-star's: loss 1e-3, dH 1e-2 (relative Frobenius over the
sampled rows, and max-abs / max|ref|)."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_fro, rel_max

pytestmark = pytest.mark.gpu

NPAIRS = 256


def _check(B, d, tau, aligned, sym_modes):
    import maai_b200
    from oracle import ntxent_oracle as O
    from oracle.large_batch import den_all_torch
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(B + 7 * d + int(100 * tau))
    h1 = torch.randn(B, d, generator=g, device=dev)
    h2 = h1 + 0.5 * torch.randn(B, d, generator=g, device=dev) if aligned else torch.randn(B, d, generator=g, device=dev)
    den_all = den_all_torch(h1, h2, tau).cpu().numpy()
    H1, H2 = h1.cpu().numpy(), h2.cpu().numpy()
    rng = np.random.default_rng(B + d)
    # sampled pairs: random ones plus the corners of the triangular / folded tile schedule
    pairs = np.unique(np.concatenate([rng.choice(B, NPAIRS, replace=False),
                                      [0, 1, 127, 128, 255, 256, B // 2 - 1, B // 2, B - 257, B - 129, B - 128, B - 1]]))
    ref = O.ntxent_rows_oracle(H1, H2, pairs, tau, den_all=den_all)
    # the torch fp32 all-row denominators against fp64 on the sampled rows
    assert np.abs(den_all[pairs] / ref["den"][0] - 1).max() < 2e-5
    assert np.abs(den_all[B + pairs] / ref["den"][1] - 1).max() < 2e-5
    loss_ref = O.loss_from_denominators(H1, H2, den_all, tau)
    # mean over the sampled pairs of (term_a + term_b) estimates the loss (Objective.py:79, 125)
    assert abs(ref["terms"].sum() / len(pairs) - loss_ref) < 0.05 * abs(loss_ref)
    old = os.environ.get("MAAI_FWD_SYM")
    out = {}
    try:
        for mode in sym_modes:
            os.environ["MAAI_FWD_SYM"] = mode
            x = h1.clone().requires_grad_(True)
            y = h2.clone().requires_grad_(True)
            stash = {}
            loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, device=dev, return_logits=False, _stash=stash)
            loss.backward()
            torch.cuda.synchronize()
            lneg = stash["rowsum"].cpu().numpy().astype(np.float64)
            got1, got2 = x.grad[pairs].cpu().numpy(), y.grad[pairs].cpu().numpy()
            res = dict(loss=abs(float(loss) - loss_ref) / abs(loss_ref),
                       lneg=max(np.abs(lneg[pairs] / ref["lneg"][0] - 1).max(), np.abs(lneg[B + pairs] / ref["lneg"][1] - 1).max()),
                       dh1=rel_fro(got1, ref["dh1"]), dh2=rel_fro(got2, ref["dh2"]),
                       dh1_max=rel_max(got1, ref["dh1"]), dh2_max=rel_max(got2, ref["dh2"]),
                       finite=bool(torch.isfinite(x.grad).all() and torch.isfinite(y.grad).all()))
            out[mode] = res
            assert res["finite"], (mode, res)
            assert res["loss"] <= 1e-3, (mode, res)
            assert res["lneg"] <= 3e-3, (mode, res)       # row sums of bf16 products, fp32 accumulate
            assert res["dh1"] <= 1e-2 and res["dh2"] <= 1e-2, (mode, res)
            assert res["dh1_max"] <= 2e-2 and res["dh2_max"] <= 2e-2, (mode, res)
    finally:
        if old is None:
            os.environ.pop("MAAI_FWD_SYM", None)
        else:
            os.environ["MAAI_FWD_SYM"] = old
    return out


@pytest.mark.parametrize("tau", [0.1, 0.5])
@pytest.mark.parametrize("d", [64, 128, 256])
@pytest.mark.parametrize("B", [16384, 32768])
def test_large_batch_sampled_rows(B, d, tau):
    _check(B, d, tau, aligned=False, sym_modes=("1", "0"))


@pytest.mark.parametrize("d,tau", [(64, 0.5), (128, 0.1), (128, 0.5), (256, 0.1)])
def test_65536_pairs_sampled_rows(d, tau):
    """The upper end of configs[4]: 131072 rows, 1024 x 1024 key tiles."""
    _check(65536, d, tau, aligned=False, sym_modes=("1", "0") if d == 128 else ("1",))


def test_large_batch_aligned_pairs_low_temperature():
    """Trained-like inputs (peaked softmax: the positive dominates the row) at the headline size."""
    _check(32768, 128, 0.1, aligned=True, sym_modes=("1", "0"))
