"""Short fwd+bwd loop at the bench workload for ncu (launch list / --set full)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maai_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randn(B, d, generator=g, device="cuda").requires_grad_(True)
y = torch.randn(B, d, generator=g, device="cuda").requires_grad_(True)
for _ in range(steps):
    x.grad = None; y.grad = None
    loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=0.5, device="cuda")
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()))
