"""Subprocess worker of tests/test_gpu_multirank.py::test_emulated_ranks_direct_symmetric_forward.

maai_ntxent_fwd_sym_direct waits, in its last CTA, for the OTHER ranks' kernels (their partial row sums arrive
by NVLink red.add, then their kind-2 flag), so emulated ranks cannot run one after the other: here every rank's
kernel is launched on its own stream of the one GPU and the shapes are small enough (a handful of CTAs per rank)
that all of them are resident at once.  Runs in its own process: a protocol bug ends in a device trap, which
must not take the pytest process down with it."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maai_b200 import _lib  # noqa: E402
from oracle import ntxent_oracle as O  # noqa: E402


def run(world, b, d, tau):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    dp = lib.maai_padded_dim(d)
    r_len = lib.maai_ntxent_r_len(b, world)
    wsb = lib.maai_ntxent_workspace_bytes(b, dp, 1) // 4
    head = (2 * b + _lib.WS_CTL_WORDS + 127) // 128 * 128
    zbuf = [torch.zeros(world, 2 * b, dp, dtype=torch.bfloat16, device=dev) for _ in range(world)]
    rbuf = [torch.zeros(r_len, device=dev) for _ in range(world)]
    flags = [torch.zeros(_lib.FLAG_WORDS, dtype=torch.int32, device=dev) for _ in range(world)]
    ctr = [torch.zeros(4, dtype=torch.int32, device=dev) for _ in range(world)]
    ws = [torch.full((wsb,), float("nan"), device=dev) for _ in range(world)]
    tab = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)
    z_tab, r_tab, f_tab = tab(zbuf), tab(rbuf), tab(flags)
    w_host = (ctypes.c_void_p * world)(*[w.data_ptr() for w in ws])
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    one = torch.ones((), device=dev)
    worst = 0.0
    for seq in (1, 2, 3):
        g = torch.Generator().manual_seed(world * 17 + b + seq)
        H1 = torch.randn(world * b, d, generator=g)
        H2 = H1 + 0.5 * torch.randn(world * b, d, generator=g)
        h = [(H1[p * b:(p + 1) * b].contiguous().to(dev), H2[p * b:(p + 1) * b].contiguous().to(dev)) for p in range(world)]
        sync = [_lib.PeerSync(f_tab.data_ptr(), flags[p].data_ptr(), ctr[p].data_ptr(), seq, 10) for p in range(world)]
        inv = torch.zeros(world, 2 * b, device=dev)
        cos = torch.zeros(world, b, device=dev)
        losses = torch.zeros(world, device=dev)
        torch.cuda.synchronize()
        for p in range(world):   # K1 of every rank: rows into every rank's buffer, zero fill of its own workspace, signal
            with torch.cuda.stream(streams[p]):
                _lib.check(lib.maai_ntxent_normalize_peer(h[p][0].data_ptr(), h[p][1].data_ptr(), b, d, 0, z_tab.data_ptr(), None,
                                                          world, p, inv[p].data_ptr(), cos[p].data_ptr(), ws[p].data_ptr(),
                                                          wsb * 4, ctypes.byref(sync[p]), streams[p].cuda_stream), "k1")
        for p in range(world):   # one launch per rank, all concurrently resident
            with torch.cuda.stream(streams[p]):
                _lib.check(lib.maai_ntxent_fwd_sym_direct(zbuf[p].data_ptr(), b, world, p, dp, 1.0 / tau, cos[p].data_ptr(),
                                                          ws[p].data_ptr(), w_host, None, r_tab.data_ptr(), None,
                                                          losses[p:].data_ptr(), ctypes.byref(sync[p]),
                                                          streams[p].cuda_stream), "direct")
        torch.cuda.synchronize()
        hr1 = [H1[p * b:(p + 1) * b].numpy() for p in range(world)]
        hr2 = [H2[p * b:(p + 1) * b].numpy() for p in range(world)]
        ol, o1, o2 = O.contrastive_loss_oracle_distributed(hr1, hr2, tau, key_grad=True)
        got = losses.cpu().numpy()
        for p in range(world):
            worst = max(worst, abs(got[p] - ol[p]) / abs(ol[p]))
            assert abs(got[p] - ol[p]) <= 1e-3 * abs(ol[p]), (seq, p, got[p], ol[p])
            g1 = torch.zeros(b, d, device=dev); g2 = torch.zeros(b, d, device=dev)
            r_row = rbuf[p][p * 2 * b:(p + 1) * 2 * b]
            _lib.check(lib.maai_ntxent_bwd(zbuf[p].data_ptr(), r_row.data_ptr(), rbuf[p].data_ptr(), 1, ws[p].data_ptr(),
                                           cos[p].data_ptr(), h[p][0].data_ptr(), h[p][1].data_ptr(), 0, inv[p].data_ptr(),
                                           one.data_ptr(), b, world, p, d, dp, 1.0 / tau, 3, g1.data_ptr(), g2.data_ptr(),
                                           ws[p][head:].data_ptr(), _lib.F_PREZEROED, ctypes.byref(sync[p]),
                                           torch.cuda.current_stream().cuda_stream), "bwd")
            torch.cuda.synchronize()
            e1 = np.linalg.norm(g1.cpu().numpy() - o1[p]) / np.linalg.norm(o1[p])
            e2 = np.linalg.norm(g2.cpu().numpy() - o2[p]) / np.linalg.norm(o2[p])
            assert e1 <= 1e-2 and e2 <= 1e-2, (seq, p, e1, e2)
    return worst


if __name__ == "__main__":
    for cfg in ((2, 96, 128, 0.5), (4, 100, 64, 0.2), (3, 130, 256, 0.3), (5, 60, 128, 0.1)):
        w = run(*cfg)
        print(f"world={cfg[0]} b={cfg[1]} d={cfg[2]} tau={cfg[3]}: direct symmetric forward ok, worst loss rel err {w:.2e}", flush=True)
    print("EMU_DIRECT_OK")
