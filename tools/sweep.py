"""BASELINE.json configs[4]: NT-Xent sweep d in {64,128,256} x global batch 1K..64K pairs x tau in
{0.1, 0.5}, forward+backward through the public API, CUDA-event timing, vs the tensor-core roofline.

    python tools/sweep.py [--out gpurun_out/sweep_n1.json]                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py ...

Per point: median ms per step (L2 flushed between steps), pairs/s of the whole job, algorithmic
TFLOP/s per GPU (24 B^2 d / W) and its fraction of the measured bf16 peak.  (Parity at these shapes
is covered by tests/test_gpu_parity.py::test_sweep_shapes_parity; this tool only measures.)
"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maai_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--dims", default="64,128,256")
    ap.add_argument("--pairs", default="1024,2048,4096,8192,16384,32768,65536")
    ap.add_argument("--taus", default="0.1,0.5")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:
        peak = 1590.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for d in [int(x) for x in args.dims.split(",")]:
        for B in [int(x) for x in args.pairs.split(",")]:
            if B % world:
                continue
            b = B // world
            for tau in [float(x) for x in args.taus.split(",")]:
                g = torch.Generator(device=dev).manual_seed(1234 + rank)
                x = torch.randn(b, d, generator=g, device=dev).requires_grad_(True)
                y = torch.randn(b, d, generator=g, device=dev).requires_grad_(True)
                ms = []
                for i in range(args.iters + 3):
                    x.grad = None
                    y.grad = None
                    flush.fill_(1)
                    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, local_rank=rank,
                                                            world_size=world, device=dev)
                    loss.backward()
                    e.record()
                    torch.cuda.synchronize()
                    if i >= 3:
                        ms.append(a.elapsed_time(e))
                med = statistics.median(ms)
                if world > 1:
                    t = torch.tensor([med], device=dev, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    med = float(t)
                tf = 24.0 * B * B * d / (med * 1e-3) / 1e12 / world
                row = dict(d=d, pairs_global=B, tau=tau, n_gpus=world, ms_per_step=med, pairs_per_s=B / (med * 1e-3),
                           tflops_per_gpu_algorithmic=tf, frac_bf16_peak=tf / peak, loss=float(loss.detach()))
                rows.append(row)
                if rank == 0:
                    print(f"d={d:3d} B={B:6d} tau={tau:.1f} W={world}: {med:8.4f} ms  {row['pairs_per_s']:.3e} pairs/s  "
                          f"{tf:7.1f} TFLOP/s/GPU ({100 * tf / peak:5.1f}% of {peak:.0f})", flush=True)
    if rank == 0 and args.out:
        json.dump(dict(peak_bf16_tflops=peak, rows=rows), open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
