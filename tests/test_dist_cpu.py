"""CPU (gloo, world_size 2) test of the N>1 host logic: the product's gather helpers
(gather_rows / gather_row_factors / positive_index) move each rank's block into the rank-major
layout, and the rank-local symmetry identity the CUDA kernels evaluate,
    dZ_i = (1/tau) [ sum_{j != i, pos} E_ij (r_i + r_j) z_j + cpos_i z_pos(i) ],
restated here in numpy on the gathered data, reproduces the oracle's full distributed gradient
(sum over ranks of the per-rank losses; after DDP's 1/W it is the single-process reference)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, rel_fro


def _worker(rank, world, port, b, d, tau, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import maai_b200
    from maai_b200 import Objective as P
    g = torch.Generator().manual_seed(100 + rank)
    h1 = torch.randn(b, d, generator=g, dtype=torch.float64)
    h2 = h1 + 0.5 * torch.randn(b, d, generator=g, dtype=torch.float64)
    z = torch.cat([torch.nn.functional.normalize(h1, dim=1), torch.nn.functional.normalize(h2, dim=1)])
    z_all = torch.zeros(world, 2 * b, d, dtype=torch.float64)
    z_all[rank] = z
    P.gather_rows(z_all, rank)                                   # product helper, gloo
    Z = z_all.reshape(world * 2 * b, d).numpy()
    pos = P.positive_index(b, world).numpy()
    loc = np.arange(rank * 2 * b, (rank + 1) * 2 * b)
    # what K2 computes for this rank's anchors
    E = np.exp((Z[loc] @ Z.T - 1.0) / tau)
    e_pos = E[np.arange(2 * b), pos[loc]].copy()
    E[np.arange(2 * b), loc] = 0.0
    E[np.arange(2 * b), pos[loc]] = 0.0
    lneg = E.sum(1)
    loss = np.log1p(lneg / e_pos).sum() / b
    r_col = torch.zeros(world * 2 * b + 7, dtype=torch.float64)
    r_col[loc] = torch.from_numpy(1.0 / (b * (e_pos + lneg)))
    P.gather_row_factors(r_col, rank, b, world)                  # product helper, gloo
    r = r_col[:world * 2 * b].numpy()
    # what K3 + K4 compute
    A = (E * (r[loc][:, None] + r[None, :])) @ Z
    lp = lneg[np.concatenate([np.arange(b, 2 * b), np.arange(0, b)])]
    cpos = -(lneg / (e_pos + lneg) + lp / (e_pos + lp)) / b
    dz = (A + cpos[:, None] * Z[pos[loc]]) / tau
    torch.save(dict(h1=h1, h2=h2, loss=torch.tensor(float(loss), dtype=torch.float64), dz=torch.from_numpy(dz)), f"{out}.{rank}")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gloo_world2_identity_matches_oracle(tmp_path, world):
    from oracle import ntxent_oracle as O
    b, d, tau = 24, 16, 0.3
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(world, 29611, b, d, tau, out), nprocs=world, join=True)
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    h1r = [x["h1"].numpy() for x in res]
    h2r = [x["h2"].numpy() for x in res]
    losses, d1, d2 = O.contrastive_loss_oracle_distributed(h1r, h2r, tau, key_grad=True)
    for r in range(world):
        assert abs(float(res[r]["loss"]) - losses[r]) < 1e-10 * abs(losses[r])
        # oracle gradients are w.r.t. h; push the identity's dz through the normalisation Jacobian
        for view, (h, dref) in enumerate(((h1r[r], d1[r]), (h2r[r], d2[r]))):
            z, n = O.l2_normalise(h)
            dz = res[r]["dz"].numpy()[view * b:(view + 1) * b]
            dh = O._normalise_backward(h, z, n, dz)
            assert rel_fro(dh, dref) < 1e-10


def _worker_dataflows(rank, world, port, b, d, tau, out):
    """(1) key-side reduce-scatter backward, (2) cross-rank symmetric forward driven by the library's own
    group plan -- both restated in numpy on gloo-exchanged data."""
    import ctypes
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import maai_b200  # noqa: F401
    from maai_b200 import Objective as P, _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(300 + rank)
    h1 = torch.randn(b, d, generator=g, dtype=torch.float64)
    h2 = h1 + 0.5 * torch.randn(b, d, generator=g, dtype=torch.float64)
    z = torch.cat([torch.nn.functional.normalize(h1, dim=1), torch.nn.functional.normalize(h2, dim=1)])
    z_all = torch.zeros(world, 2 * b, d, dtype=torch.float64)
    z_all[rank] = z
    P.gather_rows(z_all, rank)
    m, M = 2 * b, 2 * b * world
    Z = z_all.reshape(M, d).numpy()
    pos = P.positive_index(b, world).numpy()
    loc = np.arange(rank * m, (rank + 1) * m)

    def masked_E(rows, cols):
        E = np.exp((Z[rows] @ Z[cols].T - 1.0) / tau)
        E[rows[:, None] == cols[None, :]] = 0.0
        E[pos[rows][:, None] == cols[None, :]] = 0.0
        return E

    # ---- (2) symmetric forward: the tiles of this rank's group plan ----
    ng = ctypes.c_int()
    q0 = (ctypes.c_int * 9)(); rows = (ctypes.c_int * 9)(); nkt = (ctypes.c_int * 9)(); items = (ctypes.c_longlong * 9)()
    _lib.check(lib.maai_debug_group_plan(b, world, rank, 128, ctypes.byref(ng), q0, rows, nkt, items), "plan")
    rowsum = masked_E(loc, loc).sum(1)                     # group 0: own block (the kernel walks its upper triangle)
    stage = torch.zeros(world, m, dtype=torch.float64)     # this rank's staging vectors
    for gi in range(1, ng.value):
        a = np.arange(q0[gi], q0[gi] + rows[gi])
        k = loc[:min(nkt[gi] * 128, m)]
        E = masked_E(a, k)
        stage.view(-1)[torch.from_numpy(a)] += torch.from_numpy(E.sum(1))   # row sums of the remote anchors
        rowsum[:len(k)] += E.sum(0)                                          # column sums = own row sums
    stages = [torch.zeros_like(stage) for _ in range(world)]
    dist.all_gather(stages, stage)                         # stands in for the peer-mapped staging vectors
    lneg_sym = rowsum + sum(stages[p][rank].numpy() for p in range(world) if p != rank)
    lneg = masked_E(loc, np.arange(M)).sum(1)              # what the full forward computes
    # ---- (1) reduce-scatter backward ----
    e_pos = np.exp(((Z[loc] * Z[pos[loc]]).sum(1) - 1.0) / tau)
    r_loc = 1.0 / (b * (e_pos + lneg))
    allrows = np.arange(M)
    Ek = masked_E(allrows, loc)                            # anchors: everyone, keys: mine
    keyside = torch.from_numpy((Ek * r_loc[None, :]) @ Z[loc])
    dist.all_reduce(keyside)                               # reduce_scatter(sum) = all_reduce + own slice
    A = (masked_E(loc, allrows) * r_loc[:, None]) @ Z + keyside.numpy()[loc]
    lp = lneg[np.concatenate([np.arange(b, m), np.arange(0, b)])]
    cpos = -(lneg / (e_pos + lneg) + lp / (e_pos + lp)) / b
    dz = (A + cpos[:, None] * Z[pos[loc]]) / tau
    torch.save(dict(h1=h1, h2=h2, lneg=torch.from_numpy(lneg), lneg_sym=torch.from_numpy(lneg_sym),
                    dz=torch.from_numpy(dz)), f"{out}.{rank}")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,b", [(2, 96), (3, 70), (4, 130)])
def test_gloo_reduce_scatter_and_symmetric_forward_dataflows(tmp_path, world, b):
    from oracle import ntxent_oracle as O
    d, tau = 12, 0.3
    out = str(tmp_path / "r")
    mp.spawn(_worker_dataflows, args=(world, 29613 + world, b, d, tau, out), nprocs=world, join=True)
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    h1r = [x["h1"].numpy() for x in res]
    h2r = [x["h2"].numpy() for x in res]
    _, d1, d2 = O.contrastive_loss_oracle_distributed(h1r, h2r, tau, key_grad=True)
    for r in range(world):
        assert np.allclose(res[r]["lneg_sym"].numpy(), res[r]["lneg"].numpy(), rtol=1e-12, atol=0)
        for view, (h, dref) in enumerate(((h1r[r], d1[r]), (h2r[r], d2[r]))):
            z, n = O.l2_normalise(h)
            dz = res[r]["dz"].numpy()[view * b:(view + 1) * b]
            assert rel_fro(O._normalise_backward(h, z, n, dz), dref) < 1e-10


def test_positive_index_layout():
    from maai_b200 import Objective as P
    pos = P.positive_index(3, 2).tolist()
    assert pos == [3, 4, 5, 0, 1, 2, 9, 10, 11, 6, 7, 8]
