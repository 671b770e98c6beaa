"""Per-step time (CUDA events over the whole step, L2 not flushed: the small-batch regime lives in L2) of
eager and graphed fwd+bwd at the batch sizes the reference trains with (configs[0]/[1]).
    [MAAI_PDL=1 MAAI_DEBUG_LIB=<variant.so>] python tools/small_batch_time.py [pairs ...]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maai_b200  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [256, 4096]
tag = f"lib={os.path.basename(os.environ.get('MAAI_DEBUG_LIB', 'default'))} PDL={os.environ.get('MAAI_PDL', '0')}"
for b in sizes:
    d, tau = 128, 0.5
    x = torch.randn(b, d, device="cuda", requires_grad=True)
    y = torch.randn(b, d, device="cuda", requires_grad=True)
    fn = maai_b200.GraphedNTXentLoss(b, d, tau, device="cuda", hidden1_requires_grad=True)

    def eager():
        x.grad = y.grad = None
        loss = maai_b200.contrastive_loss(x, y, temperature=tau)[0]
        loss.backward()
        return loss

    def graphed():
        x.grad = y.grad = None
        loss = fn(x, y)
        loss.backward()
        return loss

    for name, step in (("eager", eager), ("graphed", graphed)):
        for _ in range(20):
            step()
        torch.cuda.synchronize()
        ms = []
        for _ in range(200):
            a = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            a.record(); loss = step(); e.record()
            ms.append((a, e))
        torch.cuda.synchronize()
        t = [a.elapsed_time(e) for a, e in ms]
        print(f"{tag} pairs={b} {name}: median {statistics.median(t) * 1e3:.1f} us  min {min(t) * 1e3:.1f} us  "
              f"loss {float(loss):.5f} |dx| {float(x.grad.norm()):.5e}", flush=True)
