"""Generate golden vectors for the NT-Xent path from the *imported reference*.

Run ONLY in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

It imports /root/reference/SimCLR/Objective.py (and SimCLR.py for the legacy loop) unmodified,
runs them on seeded inputs and writes tests/golden/ntxent_golden.npz (+ a W=2 gloo fixture
produced by the reference's own ``world_size > 1`` branch).  Nothing in tests/, smoke() or
bench.py reads /root/reference at run time -- they read the committed .npz files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/SimCLR"

# (name, b, d, tau, distribution, seed)
CASES = [
    ("c1_b256_d128_t05", 256, 128, 0.5, "randn", 0),          # BASELINE configs[0]; survey pin 12.521350
    ("aligned_b100_d64_t01", 100, 64, 0.1, "aligned", 1),      # trained-like, peaked softmax, ragged b
    ("ragged_b37_d20_t05", 37, 20, 0.5, "randn", 2),           # d not a multiple of 8, b not of 32
    ("b192_d256_t01", 192, 256, 0.1, "randn", 3),
    ("b130_d128_t005_aligned", 130, 128, 0.05, "aligned", 4),  # CLI default temperature (Contrastive_Learning.py:130)
    ("b64_d128_t1_scaled", 64, 128, 1.0, "scaled", 5),         # function default temperature, row norms 1e-3..1e3
    ("b1_d16_t05", 1, 16, 0.5, "randn", 6),                    # minimum batch
]


def make_inputs(b, d, dist_name, seed):
    g = torch.Generator().manual_seed(seed)
    h1 = torch.randn(b, d, generator=g)
    if dist_name == "randn":
        h2 = torch.randn(b, d, generator=g)
    elif dist_name == "aligned":
        h2 = h1 + 0.3 * torch.randn(b, d, generator=g)
    elif dist_name == "scaled":
        h2 = torch.randn(b, d, generator=g)
        s = torch.logspace(-3, 3, b).unsqueeze(1)
        h1 = h1 * s
        h2 = h2 * s.flip(0)
    else:
        raise ValueError(dist_name)
    return h1, h2


def _dist_worker(rank, world, port, b, d, tau, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, REF)
    import Objective  # the unmodified reference
    h1, h2 = make_inputs(world * b, d, "aligned", seed)
    x = h1[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    y = h2[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    loss, logits_ab, labels = Objective.contrastive_loss(x, y, temperature=tau, local_rank=rank,
                                                         world_size=world, device="cpu")
    loss.backward()
    torch.save(dict(loss=loss.detach(), dh1=x.grad, dh2=y.grad, logits_ab=logits_ab.detach(),
                    labels=labels), f"{out}.{rank}")
    dist.destroy_process_group()


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    import Objective  # noqa: the unmodified reference
    import SimCLR as RefSimCLR
    from oracle.ntxent_torch_port import ntxent_port_fwd_bwd

    torch.set_num_threads(4)
    blob = {}
    for name, b, d, tau, dist_name, seed in CASES:
        h1, h2 = make_inputs(b, d, dist_name, seed)
        x = h1.clone().requires_grad_(True)
        y = h2.clone().requires_grad_(True)
        loss, logits_ab, labels = Objective.contrastive_loss(x, y, temperature=tau)
        loss.backward()
        # the same reference code run in float64 = noise floor of the fp32 reference
        xd = h1.double().requires_grad_(True)
        yd = h2.double().requires_grad_(True)
        loss64, _, _ = Objective.contrastive_loss(xd, yd, temperature=tau)
        loss64.backward()
        # hidden1 detached, as the training loop calls it (Contrastive_Learning.py:685)
        y2 = h2.clone().requires_grad_(True)
        loss_det, _, _ = Objective.contrastive_loss(h1.clone(), y2, temperature=tau)
        loss_det.backward()
        # port cross-check recorded at generation time
        pl, pg1, pg2 = ntxent_port_fwd_bwd(h1, h2, tau)
        port_err = max(float((pl - loss.detach()).abs() / loss.detach().abs().clamp_min(1e-30)),
                       float((pg1 - x.grad).norm() / x.grad.norm().clamp_min(1e-30)),
                       float((pg2 - y.grad).norm() / y.grad.norm().clamp_min(1e-30)))
        blob[f"{name}.h1"] = h1.numpy()
        blob[f"{name}.h2"] = h2.numpy()
        blob[f"{name}.tau"] = np.float64(tau)
        blob[f"{name}.loss"] = loss.detach().numpy()
        blob[f"{name}.dh1"] = x.grad.numpy()
        blob[f"{name}.dh2"] = y.grad.numpy()
        blob[f"{name}.loss64"] = loss64.detach().numpy()
        blob[f"{name}.dh1_64"] = xd.grad.numpy()
        blob[f"{name}.dh2_64"] = yd.grad.numpy()
        blob[f"{name}.dh2_h1detached"] = y2.grad.numpy()
        blob[f"{name}.port_err"] = np.float64(port_err)
        k5 = min(5, b)
        topk = torch.topk(logits_ab.detach(), k=k5, dim=1)[1]
        tgt = torch.argmax(labels, dim=1)
        blob[f"{name}.top1"] = np.float64((topk[:, :1] == tgt[:, None]).any(1).float().mean())
        blob[f"{name}.top5"] = np.float64((topk == tgt[:, None]).any(1).float().mean())
        if b <= 64:
            blob[f"{name}.logits_ab"] = logits_ab.detach().numpy()
            blob[f"{name}.labels"] = labels.numpy()
        print(f"{name}: loss={float(loss):.7f} |dh1|={float(x.grad.norm()):.7f} "
              f"|dh2|={float(y.grad.norm()):.7f} port_err={port_err:.2e}")

    # legacy Algorithm-1 loop: the reference's only internal redundancy (SimCLR.py:132-144)
    h1, h2 = make_inputs(8, 16, "randn", 7)
    legacy = RefSimCLR.compute_loss(h1, h2, 0.5)
    modern, _, _ = Objective.contrastive_loss(h1, h2, temperature=0.5)
    blob["legacy.h1"] = h1.numpy(); blob["legacy.h2"] = h2.numpy()
    blob["legacy.compute_loss"] = np.float64(legacy)
    blob["legacy.contrastive_loss"] = np.float64(modern)
    print(f"legacy compute_loss={float(legacy):.6f}  contrastive_loss={float(modern):.7f} "
          f"ratio={float(legacy) / float(modern):.4f} (N^2/2 = 32)")

    # the reference's own world_size=2 branch under gloo
    W, b, d, tau, seed = 2, 48, 32, 0.5, 11
    out = "/tmp/maai_golden_dist.pt"
    mp.spawn(_dist_worker, args=(W, 29533, b, d, tau, seed, out), nprocs=W, join=True)
    h1, h2 = make_inputs(W * b, d, "aligned", seed)
    blob["dist2.h1"] = h1.numpy(); blob["dist2.h2"] = h2.numpy()
    blob["dist2.tau"] = np.float64(tau); blob["dist2.b"] = np.int64(b)
    for r in range(W):
        t = torch.load(f"{out}.{r}")
        blob[f"dist2.loss.{r}"] = t["loss"].numpy()
        blob[f"dist2.dh1.{r}"] = t["dh1"].numpy()
        blob[f"dist2.dh2.{r}"] = t["dh2"].numpy()
        os.remove(f"{out}.{r}")
    x = h1.clone().requires_grad_(True); y = h2.clone().requires_grad_(True)
    gl, _, _ = Objective.contrastive_loss(x, y, temperature=tau)
    gl.backward()
    blob["dist2.global_loss"] = gl.detach().numpy()
    blob["dist2.global_dh1"] = x.grad.numpy(); blob["dist2.global_dh2"] = y.grad.numpy()
    print(f"dist2: rank losses {[float(blob[f'dist2.loss.{r}']) for r in range(W)]} "
          f"global {float(gl):.7f}")

    blob["meta.torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "ntxent_golden.npz"), **blob)
    print("wrote", os.path.join(HERE, "ntxent_golden.npz"))


if __name__ == "__main__":
    main()
