"""Full-gradient backward at world_size > 1: symmetry-identity form (default, no gradient collective)
vs key-side reduce-scatter form (key_grad="reduce_scatter"), same inputs, CUDA-event timed, max over ranks.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/rs_compare.py \
        --pairs 32768 --dim 128 --out gpurun_out/rs_compare_n8.json
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maai_b200  # noqa: E402
from maai_b200.Objective import peer_gather_available  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=32768, help="global batch (pairs)")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--temperature", type=float, default=0.5)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    b = a.pairs // world
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    h1 = torch.randn(b, a.dim, generator=g, device=dev).requires_grad_(True)
    h2 = (h1.detach() + 0.5 * torch.randn(b, a.dim, generator=g, device=dev)).requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peer = peer_gather_available()
    arms = [("identity", True, peer), ("identity_nccl", True, False), ("reduce_scatter", "reduce_scatter", peer),
            ("query_side_only", False, peer)]
    res, grads = {}, {}
    for name, kg, pg in arms:
        def step():
            h1.grad = h2.grad = None
            loss, _, _ = maai_b200.contrastive_loss(h1, h2, temperature=a.temperature, local_rank=rank,
                                                    world_size=world, device=dev, key_grad=kg, peer_gather=pg)
            loss.backward()
            return loss
        for _ in range(a.warmup):
            step()
        ts = []
        for _ in range(a.steps):
            flush.zero_()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([sum(ts) / len(ts), sorted(ts)[len(ts) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = dict(ms_mean=float(t[0]), ms_median=float(t[1]))
        grads[name] = (h1.grad.clone(), h2.grad.clone())
    d1 = (grads["identity"][0] - grads["reduce_scatter"][0]).norm() / grads["identity"][0].norm()
    d2 = (grads["identity"][1] - grads["reduce_scatter"][1]).norm() / grads["identity"][1].norm()
    if rank == 0:
        out = dict(config=dict(pairs=a.pairs, dim=a.dim, temperature=a.temperature, n_gpus=world, steps=a.steps,
                               timing="one fwd+bwd call per step, barrier + L2 flush between steps, CUDA events, "
                                      "max over ranks", peer_gather=peer),
                   arms=res, identity_vs_reduce_scatter_grad_rel=[float(d1), float(d2)])
        print(json.dumps(out, indent=1))
        if a.out:
            json.dump(out, open(a.out, "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
