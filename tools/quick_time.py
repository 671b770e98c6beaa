"""Quick CUDA-event timing of the fwd / bwd C-ABI calls through the public API (for A/B builds:
MAAI_DEBUG_LIB=<variant.so> python tools/quick_time.py [B d iters tau])."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maai_b200  # noqa: E402
from maai_b200.Objective import _Profiler  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tau = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randn(B, d, generator=g, device="cuda").requires_grad_(True)
y = torch.randn(B, d, generator=g, device="cuda").requires_grad_(True)
for i in range(iters + 5):
    if i == 5:
        torch.cuda.synchronize()
        _Profiler.reset()
        _Profiler.enabled = True
    x.grad = None
    y.grad = None
    loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, device="cuda")
    loss.backward()
torch.cuda.synchronize()
ms = _Profiler.collect_ms()
f, w = ms["fwd"], ms["bwd"]
fm, wm = statistics.median(f), statistics.median(w)
fl = 24.0 * B * B * d
print(f"{os.path.basename(os.environ.get('MAAI_DEBUG_LIB', 'default')):24s} B={B} d={d} fwd {fm:.4f} (min {min(f):.4f}) "
      f"bwd {wm:.4f} (min {min(w):.4f}) sum {fm + wm:.4f} ms -> {fl / ((fm + wm) * 1e-3) / 1e12:.1f} TFLOP/s alg  "
      f"loss {float(loss.detach()):.6f} |dx| {float(x.grad.norm()):.6e}")
