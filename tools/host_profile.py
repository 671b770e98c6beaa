"""cProfile of the Python host path at a tiny batch (tools/host_overhead.py measures the total)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maai_b200  # noqa: E402

b, d, n = 256, 128, 2000
x = torch.randn(b, d, device="cuda", requires_grad=True)
y = torch.randn(b, d, device="cuda", requires_grad=True)


def run(k):
    for _ in range(k):
        x.grad = None
        y.grad = None
        loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=0.5)
        loss.backward()
    torch.cuda.synchronize()


run(50)
pr = cProfile.Profile()
pr.enable()
run(n)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
