"""K1 span with and without chained views at the bench shape (torchrun, N ranks): NTXentLoss(chain_views=...) over the
training loop's call shape (hidden1 = last step's hidden2, detached), CUDA-event spans of the normalize call."""
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maai_b200  # noqa: E402
from maai_b200.Objective import _Profiler  # noqa: E402

rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d, tau, iters = 128, 0.5, 40
b = B // world
g = torch.Generator(device=dev).manual_seed(5 + rank)
outs = [torch.randn(b, d, generator=g, device=dev).requires_grad_(True) for _ in range(4)]
for chained in (False, True):
    mod = maai_b200.NTXentLoss(temperature=tau, local_rank=rank, world_size=world, key_grad=True, chain_views=chained)
    o1 = outs[0]
    spans = []
    for t in range(iters + 5):
        if t == 5:
            torch.cuda.synchronize(); _Profiler.reset(); _Profiler.enabled = True
        o2 = outs[(t + 1) % 4]
        o2.grad = None
        a = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        a.record()
        loss = mod(o1.data, o2)
        loss.backward()
        e.record()
        spans.append((a, e))
        o1 = o2
    torch.cuda.synchronize()
    _Profiler.enabled = False
    k1 = statistics.median(_Profiler.collect_ms()["normalize"])
    step = statistics.median(x.elapsed_time(y) for x, y in spans[5:])
    if rank == 0:
        print(f"W={world} b={b} chain_views={chained}: K1 span {k1 * 1e3:.1f} us, step (hidden1 detached) {step * 1e3:.1f} us, "
              f"chained steps {mod.chained_steps}, loss {float(loss.detach()):.5f}", flush=True)
if world > 1:
    dist.destroy_process_group()
