"""multi-step drop-in check.  Two copies of one small encoder + MLP
projection head (same initial weights, same synthetic view stream) are trained for N steps in the
reference's loop structure (Contrastive_Learning.py:638-700: hidden1 = the previous outputs, DETACHED;
hidden2 = model(new view); loss.backward(); optimizer.step(); warm-up + cosine learning rate in the
manner of Model_Util.py:9-54), one with the reference's formulation of the loss in PyTorch fp32 ops
on the GPU, one with maai_b200.contrastive_loss.  Prints both loss curves and their largest relative
difference.

    python tools/convergence_parity.py [--steps 100] [--batch 512] [--out gpurun_out/convergence.json]
"""
import argparse
import copy
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_runner import reference_loss  # noqa: E402  (the unmodified reference file, or the one torch port)


def reference_formulation(hidden1, hidden2, temperature):
    """Objective.py:41-79, world_size == 1 branch, as the reference executes it."""
    return reference_loss(hidden1, hidden2, temperature)


def lr_at(step, base_lr, warmup, total):
    if step < warmup:
        return base_lr * (step + 1) / warmup
    return base_lr * 0.5 * (1 + math.cos(math.pi * min(step - warmup, total - warmup) / (total - warmup)))


def run(steps=100, batch=512, in_dim=96, out_dim=128, temperature=0.5, base_lr=0.3, seed=0, device="cuda:0"):
    import maai_b200
    dev = torch.device(device)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(seed)
        proto = torch.nn.Sequential(torch.nn.Linear(in_dim, 256), torch.nn.ReLU(), torch.nn.Linear(256, 512),
                                    torch.nn.ReLU(), torch.nn.Linear(512, out_dim)).to(dev)
        g = torch.Generator(device=dev).manual_seed(seed + 1)
        # "images": a fixed set of latent points; a view = the point through a random smooth distortion + noise
        latents = torch.randn(batch, in_dim, generator=g, device=dev)
        noises = [0.3 * torch.randn(batch, in_dim, generator=g, device=dev) for _ in range(steps + 1)]
        curves = {}
        for arm in ("reference", "fused"):
            model = copy.deepcopy(proto)
            opt = torch.optim.SGD(model.parameters(), lr=base_lr, momentum=0.9, weight_decay=1e-6)
            out1 = model(latents + noises[0])
            curve = []
            for t in range(steps):
                for pg in opt.param_groups:
                    pg["lr"] = lr_at(t, base_lr, max(1, steps // 10), steps)
                out2 = model(latents + noises[t + 1])
                if arm == "fused":
                    loss, _, _ = maai_b200.contrastive_loss(hidden1=out1.data, hidden2=out2, temperature=temperature,
                                                            local_rank=0, world_size=1, device=dev)
                else:
                    loss = reference_formulation(out1.data, out2, temperature)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                out1 = out2  # chained views (Contrastive_Learning.py:700)
                curve.append(float(loss.detach()))
            curves[arm] = curve
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    rel = [abs(a - b) / abs(a) for a, b in zip(curves["reference"], curves["fused"])]
    return dict(steps=steps, batch=batch, temperature=temperature, curves=curves, max_rel_diff=max(rel),
                final_rel_diff=rel[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--temperature", type=float, default=0.5)
    ap.add_argument("--base-lr", type=float, default=0.3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = run(a.steps, a.batch, temperature=a.temperature, base_lr=a.base_lr)
    c = res["curves"]
    for t in range(0, a.steps, max(1, a.steps // 10)):
        print(f"step {t:4d}: reference {c['reference'][t]:.5f}  fused {c['fused'][t]:.5f}")
    print(f"step {a.steps - 1:4d}: reference {c['reference'][-1]:.5f}  fused {c['fused'][-1]:.5f}")
    print(json.dumps({k: v for k, v in res.items() if k != "curves"}))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
