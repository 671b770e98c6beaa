"""Host-side cost per fwd+bwd step at a tiny batch (the GPU work is ~60 us): wall time per step with a
sync only at the end, i.e. how fast Python + ctypes + torch can enqueue the 9 launches."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maai_b200  # noqa: E402

b, d, n = 256, 128, 2000
x = torch.randn(b, d, device="cuda", requires_grad=True)
y = torch.randn(b, d, device="cuda", requires_grad=True)
for _ in range(50):
    loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=0.5)
    loss.backward()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    x.grad = None
    y.grad = None
    loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=0.5)
    loss.backward()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
t0 = time.perf_counter()
for _ in range(n):
    loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=0.5)
torch.cuda.synchronize()
df = (time.perf_counter() - t0) / n
print(f"b={b} d={d}: fwd+bwd {dt * 1e6:.1f} us/step, fwd only (graph built, no backward) {df * 1e6:.1f} us/step")
