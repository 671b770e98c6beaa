"""Small fwd+bwd + evaluation forward for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maai_b200  # noqa: E402

for (b, d, tau) in ((200, 128, 0.5), (37, 20, 0.1), (130, 256, 0.2), (96, 64, 0.5)):
    g = torch.Generator().manual_seed(b)
    x = torch.randn(b, d, generator=g).cuda().requires_grad_(True)
    y = torch.randn(b, d, generator=g).cuda().requires_grad_(True)
    loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau)
    loss.backward()
    with torch.no_grad():
        l2, ranks, _ = maai_b200.contrastive_loss(x, y, temperature=tau, fused_topk=True)
    torch.cuda.synchronize()
    print(b, d, tau, float(loss.detach()), float(l2), int(ranks.sum()), float(x.grad.norm()))
print("SANITIZE_RUN_OK")
