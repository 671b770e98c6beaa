"""CPU tests: the fp64 oracle and the torch port against the golden vectors that the imported
reference produced (tests/golden/make_golden.py).  These pin the oracle; the GPU parity tests
then compare the CUDA path with the oracle and with the same golden vectors."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, rel_fro
from oracle import ntxent_oracle as O
from oracle.ntxent_torch_port import ntxent_port_fwd_bwd


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_fp64(golden, name):
    h1, h2, tau = golden[f"{name}.h1"], golden[f"{name}.h2"], float(golden[f"{name}.tau"])
    loss, dh1, dh2 = O.contrastive_loss_oracle(h1, h2, tau)
    ref = float(golden[f"{name}.loss64"])
    # fp64 restatement vs the reference code itself run in fp64: only summation order differs
    assert abs(loss - ref) <= 1e-9 * max(1.0, abs(ref))
    assert rel_fro(dh1, golden[f"{name}.dh1_64"]) < 1e-8 or np.linalg.norm(golden[f"{name}.dh1_64"]) < 1e-12
    assert rel_fro(dh2, golden[f"{name}.dh2_64"]) < 1e-8 or np.linalg.norm(golden[f"{name}.dh2_64"]) < 1e-12


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if "t005" not in c and "b1_" not in c])
def test_oracle_matches_reference_fp32(golden, name):
    """fp32 reference noise floor: <= 3e-7 loss, <= 2e-6 grad for tau in {0.1, 0.5, 1}."""
    h1, h2, tau = golden[f"{name}.h1"], golden[f"{name}.h2"], float(golden[f"{name}.tau"])
    loss, dh1, dh2 = O.contrastive_loss_oracle(h1, h2, tau)
    ref = float(golden[f"{name}.loss"])
    assert abs(loss - ref) <= 2e-6 * abs(ref)
    assert rel_fro(dh1, golden[f"{name}.dh1"]) < 2e-5
    assert rel_fro(dh2, golden[f"{name}.dh2"]) < 2e-5
    # hidden1 detached (Contrastive_Learning.py:685): dh2 is unchanged, no path into h2 is cut
    assert rel_fro(dh2, golden[f"{name}.dh2_h1detached"]) < 2e-5


def test_survey_pin(golden):
    """SURVEY 8c / BASELINE.md 2: manual_seed(0), randn(256,128) x2, tau=0.5 -> 12.521350."""
    assert abs(float(golden["c1_b256_d128_t05.loss"]) - 12.521350) < 2e-6
    assert abs(np.linalg.norm(golden["c1_b256_d128_t05.dh1"]) - 0.0221135) < 1e-6
    assert abs(np.linalg.norm(golden["c1_b256_d128_t05.dh2"]) - 0.0222122) < 1e-6


def test_hidden1_detached_gradient_is_query_plus_own_keys(golden):
    """With hidden1 = outputs1.data (the training loop) dh2 keeps its query-side (ba, bb) and
    key-side (ab, bb) terms: the full-gradient oracle must reproduce ``dh2_h1detached``."""
    name = "c1_b256_d128_t05"
    h1, h2, tau = golden[f"{name}.h1"], golden[f"{name}.h2"], float(golden[f"{name}.tau"])
    # dh2 is the same whether or not h1 requires grad: detaching h1 removes no path into h2.
    _, _, dh2 = O.contrastive_loss_oracle(h1, h2, tau)
    assert rel_fro(dh2, golden[f"{name}.dh2_h1detached"]) < 2e-5


def test_port_matches_reference(golden):
    for name in GOLDEN_CASES:
        assert float(golden[f"{name}.port_err"]) <= 1e-6
    name = "ragged_b37_d20_t05"
    l, g1, g2 = ntxent_port_fwd_bwd(torch.from_numpy(golden[f"{name}.h1"]),
                                    torch.from_numpy(golden[f"{name}.h2"]), float(golden[f"{name}.tau"]))
    assert abs(float(l) - float(golden[f"{name}.loss"])) < 1e-5
    assert rel_fro(g1.numpy(), golden[f"{name}.dh1"]) < 1e-5
    assert rel_fro(g2.numpy(), golden[f"{name}.dh2"]) < 1e-5


def test_legacy_loop_cross_check(golden):
    """SimCLR.compute_loss == contrastive_loss * N^2/2 (the reference's only internal redundancy)."""
    h1, h2 = golden["legacy.h1"], golden["legacy.h2"]
    legacy = O.legacy_compute_loss(h1, h2, 0.5)
    assert abs(legacy - float(golden["legacy.compute_loss"])) < 1e-3
    modern, _, _ = O.contrastive_loss_oracle(h1, h2, 0.5)
    assert abs(modern - float(golden["legacy.contrastive_loss"])) < 1e-5
    N = h1.shape[0]
    assert abs(legacy - modern * N * N / 2) < 1e-6 * legacy


def test_distributed_oracle_matches_reference_world2(golden):
    """The reference's own world_size=2 gloo run: per-rank loss and (query-side only) gradients."""
    b = int(golden["dist2.b"]); tau = float(golden["dist2.tau"])
    h1, h2 = golden["dist2.h1"], golden["dist2.h2"]
    h1r = [h1[r * b:(r + 1) * b] for r in range(2)]
    h2r = [h2[r * b:(r + 1) * b] for r in range(2)]
    losses, d1, d2 = O.contrastive_loss_oracle_distributed(h1r, h2r, tau, key_grad=False)
    for r in range(2):
        assert abs(losses[r] - float(golden[f"dist2.loss.{r}"])) < 2e-6 * abs(losses[r])
        assert rel_fro(d1[r], golden[f"dist2.dh1.{r}"]) < 2e-5
        assert rel_fro(d2[r], golden[f"dist2.dh2.{r}"]) < 2e-5
    # mean of the rank losses == single-process loss on the concatenated batch
    assert abs(np.mean(losses) - float(golden["dist2.global_loss"])) < 2e-6 * abs(np.mean(losses))
    # full gradient (key side reduce-scattered) / W == single-process reference gradient
    _, f1, f2 = O.contrastive_loss_oracle_distributed(h1r, h2r, tau, key_grad=True)
    assert rel_fro(np.concatenate(f1) / 2, golden["dist2.global_dh1"]) < 2e-5
    assert rel_fro(np.concatenate(f2) / 2, golden["dist2.global_dh2"]) < 2e-5
    # and the reference's distributed gradient is NOT the full one (SURVEY 8e parity caveat)
    assert rel_fro(np.concatenate(d1) / 2, golden["dist2.global_dh1"]) > 0.2


def test_properties():
    rng = np.random.default_rng(0)
    h1 = rng.standard_normal((33, 24)); h2 = rng.standard_normal((33, 24))
    l0, g1, g2 = O.contrastive_loss_oracle(h1, h2, 0.5)
    # invariance to positive row scaling (normalisation)
    s = rng.uniform(0.1, 10, (33, 1))
    l1, _, _ = O.contrastive_loss_oracle(h1 * s, h2 / s, 0.5)
    assert abs(l0 - l1) < 1e-10
    # invariance to a joint permutation of the pairs
    p = rng.permutation(33)
    l2, _, _ = O.contrastive_loss_oracle(h1[p], h2[p], 0.5)
    assert abs(l0 - l2) < 1e-10
    # gradient orthogonal to h (row-wise)
    assert np.abs((g1 * h1).sum(1)).max() < 1e-12
    assert np.abs((g2 * h2).sum(1)).max() < 1e-12
    # finite-difference check of one coordinate
    e = 1e-6
    hp = h1.copy(); hp[3, 5] += e
    hm = h1.copy(); hm[3, 5] -= e
    fd = (O.contrastive_loss_oracle(hp, h2, 0.5)[0] - O.contrastive_loss_oracle(hm, h2, 0.5)[0]) / (2 * e)
    assert abs(fd - g1[3, 5]) < 1e-6
    # row blocking does not change results
    l3, g3, _ = O.contrastive_loss_oracle(h1, h2, 0.5, block=7)
    assert abs(l3 - l0) < 1e-12 and np.abs(g3 - g1).max() < 1e-14


def test_topk_oracle(golden):
    for name in ("c1_b256_d128_t05", "aligned_b100_d64_t01"):
        h1, h2 = golden[f"{name}.h1"], golden[f"{name}.h2"]
        assert abs(O.contrastive_topk_oracle(h1, h2, 1) - float(golden[f"{name}.top1"])) < 1e-6
        assert abs(O.contrastive_topk_oracle(h1, h2, 5) - float(golden[f"{name}.top5"])) < 1e-6


def test_positive_rank_oracle_matches_topk(golden):
    """rank-of-positive restatement == the reference's top-k numbers (golden top1/top5 were produced
    by torch.topk on the imported reference's logits_ab)."""
    for name in GOLDEN_CASES:
        h1, h2 = golden[f"{name}.h1"], golden[f"{name}.h2"]
        r = O.positive_rank_oracle([h1], [h2])[0]
        for k, key in ((1, "top1"), (5, "top5")):
            if f"{name}.{key}" in golden:
                kk = min(k, h1.shape[0])
                assert abs(float((r < kk).mean()) - float(golden[f"{name}.{key}"])) < 1e-6


def test_top_k_accuracy_mirror_both_forms():
    """Model_Util.top_k_accuracy mirror: reference form (scores + one-hot / index targets,
    Model_Util.py:104-113) and fused form (pos_rank, None) agree."""
    import torch
    import maai_b200
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(50, 120, generator=g)
    tgt = torch.randint(0, 120, (50,), generator=g)
    onehot = torch.nn.functional.one_hot(tgt, 120)
    ranks = (logits > logits[torch.arange(50), tgt][:, None]).sum(1).to(torch.int32)
    for k in (1, 5, 10):
        a = maai_b200.top_k_accuracy(logits, tgt, k)
        b = maai_b200.top_k_accuracy(logits, onehot, k)
        c = maai_b200.top_k_accuracy(ranks, None, k)
        assert float(a) == float(b) == float(c)
    with pytest.raises(TypeError):
        maai_b200.top_k_accuracy(logits, None, 1)


@pytest.mark.parametrize("world,b,d,tau", [(1, 300, 32, 0.5), (3, 100, 16, 0.1), (4, 64, 128, 0.2)])
def test_sampled_rows_oracle_matches_full_oracle(world, b, d, tau):
    """The sampled-anchor oracle used at BASELINE's full sizes (ntxent_rows_oracle: sampled pairs of one
    rank x all 2B keys) is the full oracle restricted to those rows -- loss terms, both gradient
    semantics -- and the plain-torch all-row denominators (oracle/large_batch.py) agree with fp64."""
    from oracle.large_batch import den_all_torch
    rng = np.random.default_rng(world * 10 + b)
    B = world * b
    H1 = rng.standard_normal((B, d))
    H2 = H1 + 0.5 * rng.standard_normal((B, d))
    h1r = [H1[p * b:(p + 1) * b] for p in range(world)]
    h2r = [H2[p * b:(p + 1) * b] for p in range(world)]
    den = O.row_denominators(H1, H2, tau)
    den_t = den_all_torch(torch.from_numpy(H1), torch.from_numpy(H2), tau, block=77).numpy()
    assert np.abs(den_t / den - 1).max() < 1e-5  # fp32 matmul + exp
    for kg in (True, False):
        ol, o1, o2 = O.contrastive_loss_oracle_distributed(h1r, h2r, tau, key_grad=kg)
        for p in range(world):
            pairs = rng.choice(b, 17, replace=False)
            r = O.ntxent_rows_oracle(H1, H2, pairs, tau, rank=p, world=world, den_all=den if kg else None,
                                     key_grad=kg, block=5)
            assert np.abs(r["dh1"] - o1[p][pairs]).max() <= 1e-10 * np.abs(o1[p]).max()
            assert np.abs(r["dh2"] - o2[p][pairs]).max() <= 1e-10 * np.abs(o2[p]).max()
            assert abs(O.loss_from_denominators(H1, H2, den, tau, p, world) - ol[p]) <= 1e-10 * abs(ol[p])
            own = p * b + pairs
            assert np.abs(r["den"][0] / den[own] - 1).max() < 1e-12 and np.abs(r["den"][1] / den[B + own] - 1).max() < 1e-12
    if world == 1:
        loss, g1, g2 = O.contrastive_loss_oracle(H1, H2, tau)
        r = O.ntxent_rows_oracle(H1, H2, np.arange(b), tau)
        assert abs(r["terms"].sum() / b - loss) <= 1e-12 * abs(loss)
        assert rel_fro(r["dh1"], g1) < 1e-12 and rel_fro(r["dh2"], g2) < 1e-12
