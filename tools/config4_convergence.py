"""BASELINE.json configs[3] as a convergence check: the reference's training step
(Contrastive_Learning.py:685-700) on torchvision ResNet-50 + MLP(2048, hidden, 128), bf16 autocast, with the
reference's own learning-rate schedule and 'lars' optimiser through the mirrors in maai_b200.Model_Util, run twice
from identical weights and data -- once with the reference's loss formulation (oracle/ref_runner.reference_loss:
the unmodified file when present) and once with the fused loss as a maintainer would use it
(NTXentLoss(chain_views=True, key_grad=False): reference gradient semantics, chained views) -- and compared step
by step.  Single GPU or torchrun (DDP, both arms).  The two arms diverge slowly by construction (bf16 autocast
backbone, atomics): the check is that the curves track each other, not bit equality.

    python tools/config4_convergence.py --steps 100 --batch-per-gpu 128 [--out profiles/r2_config4_convergence_n1.json]
"""
import argparse
import copy
import json
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maai_b200  # noqa: E402
from maai_b200 import Model_Util  # noqa: E402
from oracle.ref_runner import reference_loss  # noqa: E402


class MLP(torch.nn.Module):
    """multilayerPerceptron.py:9-22: flatten -> Linear -> ReLU -> Linear"""

    def __init__(self, i, h, o):
        super().__init__()
        self.fc1 = torch.nn.Linear(i, h)
        self.fc2 = torch.nn.Linear(h, o)

    def forward(self, x):
        return self.fc2(torch.nn.functional.relu(self.fc1(x.flatten(1))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch-per-gpu", type=int, default=128)
    ap.add_argument("--temperature", type=float, default=0.5)
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--image", type=int, default=96, help="synthetic view size (224 = configs[3]; smaller is faster)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torchvision
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    f = torchvision.models.resnet50(weights=None)
    f.fc = torch.nn.Identity()
    proto = torch.nn.Sequential(f, MLP(2048, a.hidden, 128)).to(dev).to(memory_format=torch.channels_last)
    b = a.batch_per_gpu
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    base = torch.randn(b, 3, a.image, a.image, generator=g, device=dev)
    views = [(base + 0.2 * torch.randn(b, 3, a.image, a.image, generator=g, device=dev)).contiguous(memory_format=torch.channels_last)
             for _ in range(8)]   # a small pool of augmentations of the same images, cycled
    curves = {}
    # Two more arms put the fused arm's deviation in proportion: the reference a second time (is the reference itself
    # reproducible here?) and the reference fed with embeddings rounded to bf16 (straight-through gradient) -- a
    # perturbation of the size of the bf16 tensor-core operands of the fused path; training through the plateau and
    # the sharp drop of this loss curve amplifies any such perturbation.
    for arm in ("reference", "fused", "reference_again", "reference_bf16_inputs"):
        model = copy.deepcopy(proto)
        if world > 1:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
        opt = Model_Util.get_optimizer(model, SimpleNamespace(optimizer="lars", lr=1e-3))       # Model_Util.py:80-83
        sched = dict(optimizer=opt, warmup_epochs=1, num_examples=10 * b * world, batch_size=b, world_size=world,
                     learning_rate_scaling="linear", base_learning_rate=0.02, train_epochs=max(2, a.steps // 10 + 1))
        loss_fn = maai_b200.NTXentLoss(temperature=a.temperature, local_rank=rank, world_size=world, key_grad=False,
                                       chain_views=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outputs1 = model(views[0]).float()
        curve = []
        for t in range(a.steps):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outputs2 = model(views[(t + 1) % len(views)]).float()
            if arm == "fused":
                loss = loss_fn(outputs1.data, outputs2)                                # Contrastive_Learning.py:685-690
            elif arm == "reference_bf16_inputs":
                rb = lambda h: h + (h.bfloat16().float() - h).detach()
                loss = reference_loss(rb(outputs1.data), rb(outputs2), a.temperature, rank, world)
            else:
                loss = reference_loss(outputs1.data, outputs2, a.temperature, rank, world)
            Model_Util.learning_rate_schedule(sched)                                   # :693
            opt.zero_grad()
            loss.backward()                                                            # :698
            opt.step()
            outputs1 = outputs2                                                        # :700
            lv = loss.detach().clone()
            if world > 1:
                dist.all_reduce(lv); lv /= world
            curve.append(float(lv))
        curves[arm] = curve
        if arm == "fused":
            chained = loss_fn.chained_steps
    rel = [abs(x - y) / abs(x) for x, y in zip(curves["reference"], curves["fused"])]
    rel_floor = [abs(x - y) / abs(x) for x, y in zip(curves["reference"], curves["reference_again"])]
    rel_bf16 = [abs(x - y) / abs(x) for x, y in zip(curves["reference"], curves["reference_bf16_inputs"])]
    res = dict(config=dict(model="torchvision resnet50 + MLP(2048,%d,128), bf16 autocast" % a.hidden, batch_per_gpu=b, n_gpus=world,
                           image=a.image, steps=a.steps, temperature=a.temperature, optimizer="lars (Adam inside LARC mirror)",
                           lr_schedule="Model_Util.learning_rate_schedule mirror: linear scaling, 1 warm-up epoch, cosine",
                           fused_arm="NTXentLoss(chain_views=True, key_grad=False)", chained_steps=chained),
               loss_first=(curves["reference"][0], curves["fused"][0]), loss_last=(curves["reference"][-1], curves["fused"][-1]),
               max_rel_diff=max(rel), mean_rel_diff=sum(rel) / len(rel),
               noise_floor_reference_vs_reference=dict(max_rel_diff=max(rel_floor), mean_rel_diff=sum(rel_floor) / len(rel_floor),
                                                       loss_last=curves["reference_again"][-1]),
               sensitivity_reference_with_bf16_rounded_inputs=dict(max_rel_diff=max(rel_bf16), mean_rel_diff=sum(rel_bf16) / len(rel_bf16),
                                                                   loss_last=curves["reference_bf16_inputs"][-1]),
               curves=curves)
    if rank == 0:
        for t in range(0, a.steps, max(1, a.steps // 10)):
            print(f"step {t:4d}: reference {curves['reference'][t]:.5f}  fused {curves['fused'][t]:.5f}  "
                  f"reference again {curves['reference_again'][t]:.5f}  reference on bf16-rounded inputs {curves['reference_bf16_inputs'][t]:.5f}")
        print(json.dumps({k: v for k, v in res.items() if k != "curves"}))
        if a.out:
            json.dump(res, open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
