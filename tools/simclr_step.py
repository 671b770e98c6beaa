"""BASELINE.json configs[3]: SimCLR ResNet-50 + MLP projection head (2048 -> 128) training step on
synthetic 224x224 views, with the fused NT-Xent as a drop-in for the reference loss.

    python tools/simclr_step.py [--batch-per-gpu 256] [--steps 10]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/simclr_step.py --batch-per-gpu 512

Mirrors the reference's training step (Contrastive_Learning.py:638-700): both views go through
model = g(f(x)) (SimCLR_Module, SimCLR.py:23-31; f = ResNet-50 backbone, g = MLP head,
multilayerPerceptron.py:9-22), loss = contrastive_loss(hidden1=outputs1.data, hidden2=outputs2, ...)
with hidden1 DETACHED (:685-690), loss.backward(), optimizer.step().  The backbone is torchvision's
ResNet-50 (library cuDNN kernels: out of scope of the rebuilt path), bf16 autocast, channels_last, DDP.

Two arms, same model and inputs, timed with CUDA events:
  * "fused":     maai_b200.contrastive_loss (this repository)
  * "reference": the reference's formulation of the loss in plain PyTorch ops on the GPU (normalise,
    all_gather, one_hot labels/masks, 4 matmuls, mask, cat, log_softmax -- Objective.py:41-79), written
    out below only so that the two step times and the loss values can be compared on the GPU box.
Reports ms/step for each arm, the time spent in the loss call + its backward share, and the loss values.
"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maai_b200  # noqa: E402

from oracle.ref_runner import reference_loss  # noqa: E402  (the unmodified reference file, or the one torch port)


def reference_formulation(hidden1, hidden2, temperature, rank, world):
    """Objective.py:41-79 as the reference executes it (fp32, int64 one-hots built on the host then moved, 4 matmuls,
    its own dist.all_gather for world > 1)."""
    return reference_loss(hidden1, hidden2, temperature, rank, world)


class MLP(torch.nn.Module):
    """multilayerPerceptron.py:9-22: flatten -> Linear -> ReLU -> Linear"""

    def __init__(self, i, h, o):
        super().__init__()
        self.fc1 = torch.nn.Linear(i, h)
        self.fc2 = torch.nn.Linear(h, o)

    def forward(self, x):
        return self.fc2(F.relu(self.fc1(x.flatten(1))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch-per-gpu", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--temperature", type=float, default=0.5)
    ap.add_argument("--hidden", type=int, default=4096)  # Contrastive_Learning.py:269 MLP(2048, 4096, 128)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torchvision
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    f = torchvision.models.resnet50(weights=None)
    f.fc = torch.nn.Identity()
    model = torch.nn.Sequential(f, MLP(2048, args.hidden, 128)).to(dev).to(memory_format=torch.channels_last)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    b = args.batch_per_gpu
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    v1 = torch.randn(b, 3, 224, 224, generator=g, device=dev).contiguous(memory_format=torch.channels_last)
    v2 = (v1 + 0.1 * torch.randn(b, 3, 224, 224, generator=g, device=dev)).contiguous(memory_format=torch.channels_last)

    def step(arm, timers):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out1 = model(v1)
            out2 = model(v2)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        if arm == "fused":
            loss, _, _ = maai_b200.contrastive_loss(hidden1=out1.detach(), hidden2=out2.float(),
                                                    temperature=args.temperature, local_rank=rank,
                                                    world_size=world, device=dev)
        else:
            loss = reference_formulation(out1.detach(), out2, args.temperature, rank, world)
        e1.record()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        e2.record()
        opt.step()
        timers.append((e0, e1, e2))
        return loss

    res = {}
    # same embeddings through both formulations (drop-in check on real model outputs)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        o1, o2 = model(v1), model(v2)
    lf = maai_b200.contrastive_loss(o1.float(), o2.float(), temperature=args.temperature, local_rank=rank,
                                    world_size=world, device=dev, return_logits=False)[0]
    lr = reference_formulation(o1, o2, args.temperature, rank, world)
    res["same_inputs_loss"] = dict(fused=float(lf), reference=float(lr),
                                   rel_diff=abs(float(lf) - float(lr)) / abs(float(lr)))
    for arm in ("fused", "reference"):
        try:
            for _ in range(args.warmup):
                step(arm, [])
            torch.cuda.synchronize()
            timers, steps_ms = [], []
            for _ in range(args.steps):
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                loss = step(arm, timers)
                e.record()
                torch.cuda.synchronize()
                steps_ms.append(a.elapsed_time(e))
            res[arm] = dict(ms_per_step=statistics.median(steps_ms),
                            loss_fwd_ms=statistics.median(t[0].elapsed_time(t[1]) for t in timers),
                            model_and_loss_bwd_ms=statistics.median(t[1].elapsed_time(t[2]) for t in timers),
                            loss=float(loss.detach()), peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
        except torch.cuda.OutOfMemoryError as e:  # the reference materialises ~100*b*B bytes
            res[arm] = dict(error="CUDA out of memory: " + str(e)[:120])
        torch.cuda.reset_peak_memory_stats()
    if rank == 0:
        out = dict(config=dict(model="torchvision resnet50 + MLP(2048,%d,128)" % args.hidden, batch_per_gpu=b,
                               global_batch=b * world, n_gpus=world, views="synthetic 3x224x224, bf16 autocast, channels_last",
                               temperature=args.temperature, hidden1_detached=True), **res)
        print(json.dumps(out, indent=1))
        if args.out:
            json.dump(out, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
