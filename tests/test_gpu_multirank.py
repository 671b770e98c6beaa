"""Multi-rank parity on the GPU.

* ``test_emulated_ranks_one_gpu``: every rank's K1 -> (gather = the shared buffer) -> K2 -> K3/K4
  is driven through the C ABI on ONE GPU, rank after rank, and compared with the fp64 oracle of the
  reference's world_size>1 branch: per-rank loss (Objective.py:51-79), the reference's query-side-
  only gradient (key_grad=0) and the full gradient (key_grad=1).
* ``test_two_gpu_torchrun``: the public API under torch.distributed.run with 2 real ranks (NCCL);
  skipped when fewer than 2 GPUs are visible.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_fro

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,b,d,tau", [(2, 96, 128, 0.5), (4, 200, 64, 0.1), (8, 64, 128, 0.5), (3, 130, 256, 0.2)])
def test_emulated_ranks_one_gpu(world, b, d, tau):
    import maai_b200
    from maai_b200 import _lib
    from oracle import ntxent_oracle as O
    lib = _lib.load()
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(world * 1000 + b)
    H1 = torch.randn(world * b, d, generator=g)
    H2 = H1 + 0.5 * torch.randn(world * b, d, generator=g)
    h1r = [H1[p * b:(p + 1) * b].contiguous() for p in range(world)]
    h2r = [H2[p * b:(p + 1) * b].contiguous() for p in range(world)]
    dp = lib.maai_padded_dim(d)
    z_all = torch.zeros(world, 2 * b, dp, dtype=torch.bfloat16, device=dev)
    inv = torch.zeros(world, 2 * b, device=dev)
    cos = torch.zeros(world, b, device=dev)
    dh = [(h1r[p].to(dev), h2r[p].to(dev)) for p in range(world)]
    for p in range(world):
        _lib.check(lib.maai_ntxent_normalize(dh[p][0].data_ptr(), dh[p][1].data_ptr(), b, d, 0,
                                             z_all[p].data_ptr(), inv[p].data_ptr(), cos[p].data_ptr(), None, 0, s), "k1")
    rowsum = torch.zeros(world, 2 * b, device=dev)
    r_col = torch.zeros(lib.maai_ntxent_r_len(b, world), device=dev)
    losses = torch.zeros(world, device=dev)
    for p in range(world):
        _lib.check(lib.maai_ntxent_fwd(z_all.data_ptr(), b, world, p, dp, 1.0 / tau, cos[p].data_ptr(),
                                       rowsum[p].data_ptr(), r_col[p * 2 * b:].data_ptr(),
                                       losses[p:].data_ptr(), 0, None, s), "k2")
    ol, o1q, o2q = O.contrastive_loss_oracle_distributed([h.numpy() for h in h1r], [h.numpy() for h in h2r], tau, key_grad=False)
    _, o1f, o2f = O.contrastive_loss_oracle_distributed([h.numpy() for h in h1r], [h.numpy() for h in h2r], tau, key_grad=True)
    got = losses.cpu().numpy()
    for p in range(world):
        assert abs(got[p] - ol[p]) <= 1e-3 * abs(ol[p])
    one = torch.ones((), device=dev)
    zeros = torch.zeros_like(r_col)
    acc = torch.empty(2 * b, dp, device=dev)
    for p in range(world):
        r_row = r_col[p * 2 * b:(p + 1) * 2 * b].clone()
        for key_grad, (r1, r2) in ((1, (o1f, o2f)), (0, (o1q, o2q))):
            g1 = torch.zeros(b, d, device=dev); g2 = torch.zeros(b, d, device=dev)
            _lib.check(lib.maai_ntxent_bwd(z_all.data_ptr(), r_row.data_ptr(),
                                           (r_col if key_grad else zeros).data_ptr(), key_grad,
                                           rowsum[p].data_ptr(), cos[p].data_ptr(), dh[p][0].data_ptr(),
                                           dh[p][1].data_ptr(), 0, inv[p].data_ptr(), one.data_ptr(), b, world, p,
                                           d, dp, 1.0 / tau, 3, g1.data_ptr(), g2.data_ptr(), acc.data_ptr(), 0, None, s), "k3")
            torch.cuda.synchronize()
            assert rel_fro(g1.cpu().numpy(), r1[p]) <= 1e-2, (p, key_grad)
            assert rel_fro(g2.cpu().numpy(), r2[p]) <= 1e-2, (p, key_grad)


@pytest.mark.parametrize("world,b,d", [(2, 96, 128), (4, 200, 64), (3, 130, 256), (8, 64, 128)])
def test_emulated_ranks_fused_topk(world, b, d):
    """maai_ntxent_fwd_eval with world > 1: rank p's pos_rank counts the view-b keys of ALL ranks
    (column rank*b + k of logits_ab, Objective.py:55,73) -- vs the fp64 oracle."""
    from maai_b200 import _lib
    from oracle import ntxent_oracle as O
    lib = _lib.load()
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(7 * world + b)
    H1 = torch.randn(world * b, d, generator=g)
    H2 = H1 + 0.9 * torch.randn(world * b, d, generator=g)
    h1r = [H1[p * b:(p + 1) * b].contiguous() for p in range(world)]
    h2r = [H2[p * b:(p + 1) * b].contiguous() for p in range(world)]
    dp = lib.maai_padded_dim(d)
    z_all = torch.zeros(world, 2 * b, dp, dtype=torch.bfloat16, device=dev)
    inv = torch.zeros(world, 2 * b, device=dev)
    cos = torch.zeros(world, b, device=dev)
    for p in range(world):
        a, c = h1r[p].to(dev), h2r[p].to(dev)
        _lib.check(lib.maai_ntxent_normalize(a.data_ptr(), c.data_ptr(), b, d, 0, z_all[p].data_ptr(),
                                             inv[p].data_ptr(), cos[p].data_ptr(), None, 0, s), "k1")
    ref = O.positive_rank_oracle([h.numpy() for h in h1r], [h.numpy() for h in h2r])
    ol, _, _ = O.contrastive_loss_oracle_distributed([h.numpy() for h in h1r], [h.numpy() for h in h2r], 0.5, key_grad=False)
    for p in range(world):
        rowsum = torch.zeros(2 * b, device=dev)
        loss = torch.zeros((), device=dev)
        ranks = torch.full((b,), -1, dtype=torch.int32, device=dev)
        _lib.check(lib.maai_ntxent_fwd_eval(z_all.data_ptr(), b, world, p, dp, 2.0, cos[p].data_ptr(),
                                            rowsum.data_ptr(), loss.data_ptr(), ranks.data_ptr(), s), "eval")
        torch.cuda.synchronize()
        got = ranks.cpu().numpy()
        assert abs(float(loss) - ol[p]) <= 1e-3 * abs(ol[p])
        assert np.mean(np.abs(got - ref[p]) <= 1) > 0.9 and np.abs(got - ref[p]).max() <= max(3, 0.02 * b * world), p


@pytest.mark.parametrize("world,b,d,tau", [(2, 96, 128, 0.5), (4, 200, 64, 0.1), (3, 130, 256, 0.2), (8, 64, 128, 0.5)])
def test_emulated_ranks_reduce_scatter_dataflow(world, b, d, tau):
    """Key-side reduce-scatter backward: every rank's maai_ntxent_bwd_keyside over
    its own keys, the reduce_scatter emulated by summing the partial results, then query-side tiles
    + maai_ntxent_bwd_dh -- must equal the full gradient of the oracle (and of the identity form)."""
    from maai_b200 import _lib
    from oracle import ntxent_oracle as O
    lib = _lib.load()
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(world * 977 + b)
    H1 = torch.randn(world * b, d, generator=g)
    H2 = H1 + 0.5 * torch.randn(world * b, d, generator=g)
    h1r = [H1[p * b:(p + 1) * b].contiguous() for p in range(world)]
    h2r = [H2[p * b:(p + 1) * b].contiguous() for p in range(world)]
    dp = lib.maai_padded_dim(d)
    z_all = torch.zeros(world, 2 * b, dp, dtype=torch.bfloat16, device=dev)
    inv = torch.zeros(world, 2 * b, device=dev)
    cos = torch.zeros(world, b, device=dev)
    dh = [(h1r[p].to(dev), h2r[p].to(dev)) for p in range(world)]
    for p in range(world):
        _lib.check(lib.maai_ntxent_normalize(dh[p][0].data_ptr(), dh[p][1].data_ptr(), b, d, 0,
                                             z_all[p].data_ptr(), inv[p].data_ptr(), cos[p].data_ptr(), None, 0, s), "k1")
    rowsum = torch.zeros(world, 2 * b, device=dev)
    r_loc = torch.zeros(world, 2 * b, device=dev)
    losses = torch.zeros(world, device=dev)
    for p in range(world):
        _lib.check(lib.maai_ntxent_fwd(z_all.data_ptr(), b, world, p, dp, 1.0 / tau, cos[p].data_ptr(),
                                       rowsum[p].data_ptr(), r_loc[p].data_ptr(), losses[p:].data_ptr(), 0, None, s), "k2")
    # every rank's key-side partial sums for ALL anchors, then the reduce-scatter (sum over ranks)
    total = torch.zeros(world * 2 * b, dp, device=dev)
    for p in range(world):
        r_pad = torch.zeros(lib.maai_ntxent_r_len(b, 1), device=dev)
        r_pad[:2 * b] = r_loc[p]
        part = torch.full((world * 2 * b, dp), float("nan"), device=dev)  # zeroed inside
        _lib.check(lib.maai_ntxent_bwd_keyside(z_all.data_ptr(), r_pad.data_ptr(), b, world, p, dp, 1.0 / tau,
                                               part.data_ptr(), s), "keyside")
        total += part
    _, o1f, o2f = O.contrastive_loss_oracle_distributed([h.numpy() for h in h1r], [h.numpy() for h in h2r], tau, key_grad=True)
    one = torch.ones((), device=dev)
    zeros = torch.zeros(lib.maai_ntxent_r_len(b, world), device=dev)
    for p in range(world):
        acc = torch.empty(2 * b, dp, device=dev)
        _lib.check(lib.maai_ntxent_bwd_tiles(z_all.data_ptr(), r_loc[p].data_ptr(), zeros.data_ptr(), b, world, p, dp,
                                             1.0 / tau, 3, acc.data_ptr(), s), "tiles")
        extra = total[p * 2 * b:(p + 1) * 2 * b].contiguous()
        g1 = torch.zeros(b, d, device=dev); g2 = torch.zeros(b, d, device=dev)
        _lib.check(lib.maai_ntxent_bwd_dh(acc.data_ptr(), extra.data_ptr(), rowsum[p].data_ptr(), cos[p].data_ptr(),
                                          dh[p][0].data_ptr(), dh[p][1].data_ptr(), 0, inv[p].data_ptr(),
                                          one.data_ptr(), b, d, dp, 1.0 / tau, 1, 3, g1.data_ptr(), g2.data_ptr(), s),
                   "dh")
        torch.cuda.synchronize()
        assert rel_fro(g1.cpu().numpy(), o1f[p]) <= 1e-2, p
        assert rel_fro(g2.cpu().numpy(), o2f[p]) <= 1e-2, p


@pytest.mark.parametrize("world,b,d,tau", [(2, 96, 128, 0.5), (2, 700, 128, 0.2), (4, 200, 64, 0.1), (3, 130, 256, 0.2),
                                           (8, 64, 128, 0.5), (8, 333, 96, 0.5), (5, 1024, 128, 0.3), (16, 40, 32, 0.5)])
def test_emulated_ranks_symmetric_forward(world, b, d, tau):
    """Cross-rank symmetric forward (maai_ntxent_fwd_sym_tiles + barrier + maai_ntxent_fwd_sym_finalize):
    every pair of rank slots is computed once; row sums, row factors and losses must equal those of the
    full forward (maai_ntxent_fwd) and the fp64 oracle.  Ranks are run one after the other on one GPU,
    the 'peer' addresses are the ranks' own staging buffers."""
    from maai_b200 import _lib
    from oracle import ntxent_oracle as O
    lib = _lib.load()
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(world * 31 + b)
    H1 = torch.randn(world * b, d, generator=g)
    H2 = H1 + 0.5 * torch.randn(world * b, d, generator=g)
    h1r = [H1[p * b:(p + 1) * b].contiguous() for p in range(world)]
    h2r = [H2[p * b:(p + 1) * b].contiguous() for p in range(world)]
    dp = lib.maai_padded_dim(d)
    z_all = torch.zeros(world, 2 * b, dp, dtype=torch.bfloat16, device=dev)
    inv = torch.zeros(world, 2 * b, device=dev)
    cos = torch.zeros(world, b, device=dev)
    for p in range(world):
        a, c = h1r[p].to(dev), h2r[p].to(dev)
        _lib.check(lib.maai_ntxent_normalize(a.data_ptr(), c.data_ptr(), b, d, 0, z_all[p].data_ptr(),
                                             inv[p].data_ptr(), cos[p].data_ptr(), None, 0, s), "k1")
    # reference: the full forward of every rank
    rowsum_full = torch.zeros(world, 2 * b, device=dev)
    r_full = torch.zeros(world, 2 * b, device=dev)
    loss_full = torch.zeros(world, device=dev)
    for p in range(world):
        _lib.check(lib.maai_ntxent_fwd(z_all.data_ptr(), b, world, p, dp, 1.0 / tau, cos[p].data_ptr(),
                                       rowsum_full[p].data_ptr(), r_full[p].data_ptr(), loss_full[p:].data_ptr(), 0, None, s), "k2")
    # symmetric: tile part of every rank, "barrier", then the finalize of every rank
    rowsum = torch.full((world, 2 * b), float("nan"), device=dev)
    stage = torch.full((world, world, 2 * b), float("nan"), device=dev)  # stage[p] = rank p's staging vectors
    for p in range(world):
        _lib.check(lib.maai_ntxent_fwd_sym_tiles(z_all.data_ptr(), b, world, p, dp, 1.0 / tau, rowsum[p].data_ptr(),
                                                 stage[p].data_ptr(), 0, None, s), "sym tiles")
    torch.cuda.synchronize()
    tab = torch.tensor([stage[p].data_ptr() for p in range(world)], dtype=torch.int64, device=dev)
    r_sym = torch.zeros(world, 2 * b, device=dev)
    loss_sym = torch.zeros(world, device=dev)
    for p in range(world):
        _lib.check(lib.maai_ntxent_fwd_sym_finalize(rowsum[p].data_ptr(), tab.data_ptr(), b, world, p, 1.0 / tau,
                                                    cos[p].data_ptr(), r_sym[p].data_ptr(), None, None,
                                                    loss_sym[p:].data_ptr(), None, s), "sym finalize")
    torch.cuda.synchronize()
    # (polynomial / MUFU exp2 assignment differs between the two tile walks: agreement to ~1e-4, not bitwise)
    assert torch.allclose(rowsum, rowsum_full, rtol=2e-3, atol=1e-6), (rowsum - rowsum_full).abs().max()
    assert torch.allclose(r_sym, r_full, rtol=2e-3, atol=0)
    ol, _, _ = O.contrastive_loss_oracle_distributed([h.numpy() for h in h1r], [h.numpy() for h in h2r], tau, key_grad=False)
    got = loss_sym.cpu().numpy()
    for p in range(world):
        assert abs(got[p] - ol[p]) <= 1e-3 * abs(ol[p]), (p, got[p], ol[p])


def test_two_gpu_torchrun(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    out = tmp_path / "dist.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29655",
           os.path.join(ROOT, "tests", "_dist_gpu_worker.py"), str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_OK" in res.stdout


@pytest.mark.parametrize("world,b,d,tau,sym", [(2, 96, 128, 0.5, False), (4, 200, 64, 0.1, False), (8, 64, 128, 0.5, False),
                                               (3, 130, 256, 0.2, False), (2, 700, 128, 0.2, True), (4, 200, 64, 0.1, True),
                                               (5, 300, 128, 0.3, True)])
def test_emulated_ranks_inkernel_flags(world, b, d, tau, sym):
    """The in-kernel peer synchronisation (maai_peer_sync): every emulated rank has its own key buffer, row-factor
    array, staging vectors and FLAG BLOCK on the one GPU; K1 stores into all buffers and signals, the forward's
    TMA producer waits per rank slot, the finalize stores the row factors everywhere and signals, the
    backward waits per slot before reading r_col.  Two consecutive steps (sequence numbers 1, 2) with
    different inputs; losses and full gradients against the fp64 oracle.  (Ranks run one after the other
    here, so every flag is already set when it is waited for; real concurrency is tests/_dist_gpu_worker.py.)"""
    import ctypes
    from maai_b200 import _lib
    from oracle import ntxent_oracle as O
    lib = _lib.load()
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    dp = lib.maai_padded_dim(d)
    r_len = lib.maai_ntxent_r_len(b, world)
    zbuf = [torch.zeros(world, 2 * b, dp, dtype=torch.bfloat16, device=dev) for _ in range(world)]
    rbuf = [torch.zeros(r_len, device=dev) for _ in range(world)]
    stage = [torch.zeros(world, 2 * b, device=dev) for _ in range(world)]
    flags = [torch.zeros(_lib.FLAG_WORDS, dtype=torch.int32, device=dev) for _ in range(world)]
    ctr = [torch.zeros(4, dtype=torch.int32, device=dev) for _ in range(world)]
    tab = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)
    z_tab, r_tab, s_tab, f_tab = tab(zbuf), tab(rbuf), tab(stage), tab(flags)
    one = torch.ones((), device=dev)
    for seq in (1, 2):
        g = torch.Generator().manual_seed(world * 131 + b + seq)
        H1 = torch.randn(world * b, d, generator=g)
        H2 = H1 + 0.5 * torch.randn(world * b, d, generator=g)
        h = [(H1[p * b:(p + 1) * b].contiguous().to(dev), H2[p * b:(p + 1) * b].contiguous().to(dev)) for p in range(world)]
        sync = [_lib.PeerSync(f_tab.data_ptr(), flags[p].data_ptr(), ctr[p].data_ptr(), seq, 5) for p in range(world)]
        inv = torch.zeros(world, 2 * b, device=dev); cos = torch.zeros(world, b, device=dev)
        wsb = lib.maai_ntxent_workspace_bytes(b, dp, 1) // 4
        head = (2 * b + _lib.WS_CTL_WORDS + 127) // 128 * 128
        ws = [torch.full((wsb,), float("nan"), device=dev) for _ in range(world)]
        for p in range(world):
            _lib.check(lib.maai_ntxent_normalize_peer(h[p][0].data_ptr(), h[p][1].data_ptr(), b, d, 0, z_tab.data_ptr(), None,
                                                      world, p, inv[p].data_ptr(), cos[p].data_ptr(), ws[p].data_ptr(),
                                                      wsb * 4, ctypes.byref(sync[p]), s), "k1 peer")
        torch.cuda.synchronize()
        for p in range(world):
            f = flags[p].cpu().numpy()
            assert (f[:world] == seq).all(), f[:world]          # kind 0 (rows) from every rank
            assert int(ctr[p][0]) == 0                            # the CTA counter reset itself
        losses = torch.zeros(world, device=dev)
        for p in range(world):
            if sym:
                _lib.check(lib.maai_ntxent_fwd_sym_tiles(zbuf[p].data_ptr(), b, world, p, dp, 1.0 / tau, ws[p].data_ptr(),
                                                         stage[p].data_ptr(), _lib.F_PREZEROED, ctypes.byref(sync[p]), s), "sym tiles")
            else:
                _lib.check(lib.maai_ntxent_fwd_peer(zbuf[p].data_ptr(), b, world, p, dp, 1.0 / tau, cos[p].data_ptr(),
                                                    ws[p].data_ptr(), r_tab.data_ptr(), None, losses[p:].data_ptr(),
                                                    _lib.F_PREZEROED, ctypes.byref(sync[p]), s), "fwd peer")
        if sym:
            torch.cuda.synchronize()
            for p in range(world):
                assert (flags[p].cpu().numpy()[64:64 + world] == seq).all()   # kind 2 (staged sums) from every rank
            for p in range(world):
                _lib.check(lib.maai_ntxent_fwd_sym_finalize(ws[p].data_ptr(), s_tab.data_ptr(), b, world, p, 1.0 / tau,
                                                            cos[p].data_ptr(), None, r_tab.data_ptr(), None,
                                                            losses[p:].data_ptr(), ctypes.byref(sync[p]), s), "sym finalize")
        torch.cuda.synchronize()
        for p in range(world):
            assert (flags[p].cpu().numpy()[32:32 + world] == seq).all()       # kind 1 (row factors) from every rank
        hr1 = [H1[p * b:(p + 1) * b].numpy() for p in range(world)]
        hr2 = [H2[p * b:(p + 1) * b].numpy() for p in range(world)]
        ol, o1, o2 = O.contrastive_loss_oracle_distributed(hr1, hr2, tau, key_grad=True)
        got = losses.cpu().numpy()
        for p in range(world):
            assert abs(got[p] - ol[p]) <= 1e-3 * abs(ol[p]), (seq, p)
            assert torch.equal(rbuf[p], rbuf[0])                  # every rank holds the same gathered row factors
            g1 = torch.zeros(b, d, device=dev); g2 = torch.zeros(b, d, device=dev)
            r_row = rbuf[p][p * 2 * b:(p + 1) * 2 * b]
            _lib.check(lib.maai_ntxent_bwd(zbuf[p].data_ptr(), r_row.data_ptr(), rbuf[p].data_ptr(), 1, ws[p].data_ptr(),
                                           cos[p].data_ptr(), h[p][0].data_ptr(), h[p][1].data_ptr(), 0, inv[p].data_ptr(),
                                           one.data_ptr(), b, world, p, d, dp, 1.0 / tau, 3, g1.data_ptr(), g2.data_ptr(),
                                           ws[p][head:].data_ptr(), _lib.F_PREZEROED, ctypes.byref(sync[p]), s), "bwd")
            torch.cuda.synchronize()
            assert rel_fro(g1.cpu().numpy(), o1[p]) <= 1e-2, (seq, p)
            assert rel_fro(g2.cpu().numpy(), o2[p]) <= 1e-2, (seq, p)


def test_emulated_ranks_direct_symmetric_forward():
    """maai_ntxent_fwd_sym_direct (cross-rank symmetric forward in one launch: remote red.add of the partial row
    sums, kind-2 flags, per-row tail in the kernel): emulated ranks on concurrent streams, in a subprocess
    (tests/_emulated_direct_worker.py explains why), losses and full gradients against the fp64 oracle."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_emulated_direct_worker.py")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "EMU_DIRECT_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]


def test_tensors_on_a_non_current_device():
    """one process, two GPUs -- the tensors live on cuda:1 while cuda:0 is the current device.  The host
    mirror makes the tensors' device current for the C-ABI calls (and takes ITS stream), the library keeps its
    per-device caches (SM count, dynamic shared-memory attribute) by device ordinal."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import maai_b200
    from oracle import ntxent_oracle as O
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(3)
    h1 = torch.randn(300, 128, generator=g); h2 = h1 + 0.5 * torch.randn(300, 128, generator=g)
    ol, o1, o2 = O.contrastive_loss_oracle(h1.numpy(), h2.numpy(), 0.5)
    for dev in ("cuda:1", "cuda:0", "cuda:1"):
        x = h1.to(dev).requires_grad_(True); y = h2.to(dev).requires_grad_(True)
        loss = maai_b200.contrastive_loss(x, y, temperature=0.5)[0]
        loss.backward()
        torch.cuda.synchronize(dev)
        assert torch.cuda.current_device() == 0
        assert abs(float(loss.detach()) - ol) <= 1e-3 * abs(ol), dev
        assert rel_fro(x.grad.cpu().numpy(), o1) <= 1e-2 and rel_fro(y.grad.cpu().numpy(), o2) <= 1e-2, dev
