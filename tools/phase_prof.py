"""Per-warp phase cycle breakdown of the tile kernels (needs a MAAI_PROF=1 build:
python tools/ab_variants.py build prof=MAAI_PROF=1).  Run on the B200:
    MAAI_DEBUG_LIB=multimodal-active-ai_b200/variants/prof.so python tools/phase_prof.py [B d]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maai_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
lib = maai_b200._lib.load()
lib.maai_debug_prof_read.restype = ctypes.c_int
lib.maai_debug_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
NB, NW, NP = 148, 20, 8


def read():
    buf = np.zeros(160 * NW * NP, dtype=np.int64)
    rc = lib.maai_debug_prof_read(buf.ctypes.data, buf.size)
    assert rc == 0, rc
    return buf.reshape(160, NW, NP)[:NB]


def show(name, a, labels):
    print(f"== {name}: mean cycles per CTA over {NB} CTAs (and share of the warp's total)")
    for role, warps in (("producer", [0]), ("mma", [1]), ("softmax", list(range(4, 20)))):
        x = a[:, warps, :].mean(axis=(0, 1))
        tot = x.sum()
        lab = labels[role]
        print(f"  {role:9s} total {tot:10.0f}: " + "  ".join(f"{lab[i]}={x[i]:.0f} ({100 * x[i] / max(tot, 1):.0f}%)" for i in range(NP) if lab[i]))


g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randn(B, d, generator=g, device="cuda").requires_grad_(True)
y = torch.randn(B, d, generator=g, device="cuda").requires_grad_(True)
for it in range(3):
    x.grad = None
    y.grad = None
    loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=0.5, device="cuda")
    torch.cuda.synchronize()
    fwd = read()
    loss.backward()
    torch.cuda.synchronize()
    bwd = read()
tiles = (2 * B / 128) ** 2 / NB
print(f"B={B} d={d}: {tiles:.0f} S tiles per CTA")
lab_f = {"producer": ["wait_q_empty", "wait_k_empty", "issue", "", "", "", "", ""],
         "mma": ["wait_q/k_full", "wait_sm_done", "issue_S", "issue_PV", "wait_dz_free", "seg_end", "", ""],
         "softmax": ["wait_s_full", "tmem_ld", "compute", "st+arrive", "seg_epilogue", "tile_prologue", "", ""]}
show("FWD", fwd, lab_f)
show("BWD", bwd, lab_f)
for nm, a in (("FWD", fwd), ("BWD", bwd)):
    sm = a[:, 4:20, :].mean(axis=(0, 1))
    print(f"{nm} softmax per S tile handled by a warp: " + " ".join(f"{v / tiles:.0f}" for v in sm[:6]),
          f"| total/tile {sm.sum() / tiles:.0f} clk (every tile in BWD NQ=1, every other tile in FWD NQ=2)")
