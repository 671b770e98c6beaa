"""CPU tests of the Model_Util mirrors against golden values the imported reference produced
(tests/golden/make_model_util_golden.py): learning_rate_schedule (Model_Util.py:9-39), top_k_accuracy
(Model_Util.py:104-113); LARC against its definition."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


@pytest.fixture(scope="module")
def mu_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "model_util_golden.npz"))


def test_learning_rate_schedule_matches_reference(mu_golden):
    import maai_b200
    from make_model_util_golden import LR_CASES, lr_curve
    for name, kw in LR_CASES.items():
        ref = mu_golden[f"lr.{name}"]
        got = lr_curve(maai_b200.Model_Util.learning_rate_schedule, kw, len(ref))
        # the reference keeps Adam's float32 step tensor through the cosine (fp32 rounding, ~1e-8); the
        # mirror converts it to a Python float first
        assert np.allclose(got, ref, rtol=1e-6, atol=1e-7), name  # atol: the end of the cosine is ~0
    with pytest.raises(ValueError):
        lr_curve(maai_b200.Model_Util.learning_rate_schedule, dict(LR_CASES["linear_warm"], learning_rate_scaling="cubic"), 1)


def test_top_k_accuracy_matches_reference(mu_golden):
    import maai_b200
    preds = torch.from_numpy(mu_golden["topk.preds"])
    tgt = torch.from_numpy(mu_golden["topk.target"])
    for k in (1, 5, 10):
        assert abs(float(maai_b200.top_k_accuracy(preds, tgt, k)) - float(mu_golden[f"topk.idx.k{k}"])) < 1e-7
        oh = torch.nn.functional.one_hot(tgt, 50)
        assert abs(float(maai_b200.top_k_accuracy(preds, oh, k)) - float(mu_golden[f"topk.onehot.k{k}"])) < 1e-7
    # fused form: the int rank vector
    rank = (preds > preds.gather(1, tgt.view(-1, 1))).sum(1).to(torch.int32)
    assert abs(float(maai_b200.top_k_accuracy(rank, None, 5)) - float(mu_golden["topk.idx.k5"])) < 1e-7
    with pytest.raises(TypeError):
        maai_b200.top_k_accuracy(preds, None, 5)


def test_larc_scales_gradients_by_the_trust_ratio():
    import maai_b200
    from types import SimpleNamespace
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 4)
    opt = maai_b200.Model_Util.get_optimizer(lin, SimpleNamespace(optimizer="lars", lr=0.5))
    assert isinstance(opt, maai_b200.Model_Util.LARC)
    lin(torch.randn(3, 6)).pow(2).sum().backward()
    w0 = lin.weight.detach().clone(); g0 = lin.weight.grad.detach().clone()
    expect = min(float(0.02 * w0.norm() / (g0.norm() + 1e-8)) / 0.5, 1.0)
    inner = torch.optim.Adam([torch.nn.Parameter(w0.clone())], 0.5)
    inner.param_groups[0]["params"][0].grad = g0 * expect
    inner.step()
    opt.step()
    assert torch.allclose(lin.weight.detach(), inner.param_groups[0]["params"][0].detach(), atol=1e-7)
    for name in ("sgd", "adam"):
        maai_b200.Model_Util.get_optimizer(lin, SimpleNamespace(optimizer=name, lr=0.1, momentum=0.9, weight_decay=1e-4))
    with pytest.raises(ValueError):
        maai_b200.Model_Util.get_optimizer(lin, SimpleNamespace(optimizer="rmsprop", lr=0.1))
