"""Stage-level GPU diagnostic: drives the C ABI directly and checks every intermediate
(z, inv_norm, pos_cos, row sums l, loss, dz accumulators, dh) against numpy fp64 computed from the
SAME bf16 rows.  Run on the B200 box:  python tools/gpu_check.py [b d tau [world]]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maai_b200  # noqa: E402
from maai_b200 import _lib  # noqa: E402


def run_case(b, d, tau, world=1, rank=0, seed=0, aligned=False, verbose=True):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    B = b * world
    H1 = torch.randn(B, d, generator=g)
    H2 = H1 + 0.3 * torch.randn(B, d, generator=g) if aligned else torch.randn(B, d, generator=g)
    dp = lib.maai_padded_dim(d)
    s = torch.cuda.current_stream().cuda_stream
    z_all = torch.zeros(world, 2 * b, dp, dtype=torch.bfloat16, device=dev)
    inv_all = torch.zeros(world, 2 * b, device=dev)
    cos_all = torch.zeros(world, b, device=dev)
    hs = []
    for p in range(world):  # emulate every rank's K1 + the all-gather on one GPU
        h1 = H1[p * b:(p + 1) * b].contiguous().to(dev)
        h2 = H2[p * b:(p + 1) * b].contiguous().to(dev)
        hs.append((h1, h2))
        _lib.check(lib.maai_ntxent_normalize(h1.data_ptr(), h2.data_ptr(), b, d, 0, z_all[p].data_ptr(),
                                             inv_all[p].data_ptr(), cos_all[p].data_ptr(), None, 0, s), "normalize")
    torch.cuda.synchronize()
    out = {}
    # ---- stage 1: K1
    zf = z_all.float().cpu().double().numpy().reshape(world * 2 * b, dp)
    h = np.concatenate([np.concatenate([H1[p * b:(p + 1) * b].numpy(), H2[p * b:(p + 1) * b].numpy()]) for p in range(world)]).astype(np.float64)
    n = np.maximum(np.linalg.norm(h, axis=1, keepdims=True), 1e-12)
    out["z_maxabs"] = float(np.abs(zf[:, :d] - h / n).max())
    out["z_pad_zero"] = float(np.abs(zf[:, d:]).max()) if dp > d else 0.0
    out["inv_rel"] = float(np.abs(inv_all.cpu().numpy().reshape(-1) * n[:, 0] - 1).max())
    # ---- stage 2: forward of `rank`
    M = world * 2 * b
    S = zf @ zf.T
    E = np.exp((S - 1.0) / tau)
    np.fill_diagonal(E, 0.0)
    l_ref = E.sum(1)
    loc = slice(rank * 2 * b, (rank + 1) * 2 * b)
    pos = np.concatenate([np.arange(b, 2 * b), np.arange(0, b)]) + rank * 2 * b
    rows = np.arange(rank * 2 * b, (rank + 1) * 2 * b)
    loss_ref = float((np.log(l_ref[loc]) + (1 - S[rows, pos]) / tau).sum() / b)
    rowsum = torch.zeros(2 * b, device=dev)
    r_len = lib.maai_ntxent_r_len(b, world)
    r_col = torch.zeros(r_len, device=dev)
    r_row = torch.zeros(2 * b, device=dev)
    loss = torch.zeros((), device=dev)
    t0 = time.time()
    _lib.check(lib.maai_ntxent_fwd(z_all.data_ptr(), b, world, rank, dp, 1.0 / tau, cos_all[rank].data_ptr(),
                                   rowsum.data_ptr(), r_row.data_ptr(), loss.data_ptr(), 0, None, s), "fwd")
    torch.cuda.synchronize()
    out["fwd_ms_first"] = (time.time() - t0) * 1e3
    out["cos_maxabs"] = float(np.abs(cos_all[rank].cpu().numpy() - S[rows[:b], pos[:b]]).max())
    lneg_ref = l_ref[loc] - E[rows, pos]
    out["l_relmax"] = float(np.abs(rowsum.cpu().numpy() / np.maximum(lneg_ref, 1e-300) - 1).max())
    out["loss"] = float(loss)
    out["loss_ref_bf16z"] = loss_ref
    out["loss_rel"] = abs(float(loss) - loss_ref) / abs(loss_ref)
    # ---- stage 3: backward accumulators (full gradient: r_col = every rank's r)
    r_full = 1.0 / (b * l_ref)
    r_col[:M] = torch.from_numpy(r_full).float().to(dev)
    r_row.copy_(r_col[loc])
    Eloc = E[loc].copy()
    Epos = Eloc[np.arange(2 * b), pos].copy()
    Eloc[np.arange(2 * b), pos] = 0.0  # the positive column is handled in fp32 by dh_kernel
    A_ref = (Eloc * (r_full[loc][:, None] + r_full[None, :])) @ zf
    h1, h2 = hs[rank]
    dh1 = torch.zeros_like(h1); dh2 = torch.zeros_like(h2)
    dz_acc = torch.full((2 * b, dp), float("nan"), device=dev)
    gl = torch.ones((), device=dev)
    _lib.check(lib.maai_ntxent_bwd(z_all.data_ptr(), r_row.data_ptr(), r_col.data_ptr(), 1,
                                   rowsum.data_ptr(), cos_all[rank].data_ptr(), h1.data_ptr(), h2.data_ptr(), 0, inv_all[rank].data_ptr(), gl.data_ptr(),
                                   b, world, rank, d, dp, 1.0 / tau, 3, dh1.data_ptr(), dh2.data_ptr(),
                                   dz_acc.data_ptr(), 0, None, s), "bwd")
    torch.cuda.synchronize()
    A = dz_acc.cpu().double().numpy()
    out["A_rel_fro"] = float(np.linalg.norm(A - A_ref) / np.linalg.norm(A_ref))
    out["A_nan"] = int(np.isnan(A).sum())
    # dh vs fp64 formula from the same A_ref
    zi = h[loc] / n[loc]
    zp = h[pos] / n[pos]
    cpos = Epos * (r_full[loc] + r_full[pos]) - 2.0 / b
    dz = (A_ref[:, :d] + cpos[:, None] * zp) / tau
    dh_ref = (dz - zi * (zi * dz).sum(1, keepdims=True)) / n[loc]
    dh = np.concatenate([dh1.cpu().numpy(), dh2.cpu().numpy()]).astype(np.float64)
    out["dh_rel_fro_vs_bf16z"] = float(np.linalg.norm(dh - dh_ref) / np.linalg.norm(dh_ref))
    if verbose:
        print(f"case b={b} d={d} tau={tau} world={world} rank={rank} aligned={aligned}:")
        for k, v in out.items():
            print(f"   {k:24s} {v:.6g}" if isinstance(v, float) else f"   {k:24s} {v}")
        if out["A_rel_fro"] > 0.02 and b <= 256:
            err = np.abs(A - A_ref)
            print("   A err by 32-col block:", [float(err[:, c:c + 32].max()) for c in range(0, dp, 32)])
            print("   A err by 32-row block:", [float(err[r:r + 32].max()) for r in range(0, min(2 * b, 256), 32)])
            print("   A[0,:8]    ", A[0, :8]); print("   Aref[0,:8] ", A_ref[0, :8])
    return out


if __name__ == "__main__":
    if len(sys.argv) >= 4:
        run_case(int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 1)
    else:
        print(torch.cuda.get_device_name(0))
        run_case(64, 128, 0.5)             # one row block, one key tile
        run_case(256, 128, 0.5)            # C1: 2 row blocks x 4 key tiles
        run_case(100, 64, 0.1, aligned=True)
        run_case(37, 20, 0.5)
        run_case(192, 256, 0.1)
        run_case(1000, 128, 0.5)           # ragged, multi-CTA
        run_case(96, 128, 0.5, world=2, rank=1)
        run_case(4096, 128, 0.5, verbose=True)
