"""torchrun worker for tests/test_gpu_multirank.py::test_two_gpu_torchrun and tools: runs the public
API on W real ranks (NCCL) and checks per-rank loss, full gradient and the reference-semantics
(key_grad=False) gradient against the fp64 oracle on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maai_b200  # noqa: E402
from oracle import ntxent_oracle as O  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    from maai_b200.Objective import peer_gather_available
    # (peer gather?, cross-rank symmetric forward forced on / off)
    modes = [(False, "0")] + ([(True, "0"), (True, "1")] if peer_gather_available() else [])
    if rank == 0:
        print("modes under test (peer: False = NCCL all_gather, True = fused NVLink peer stores; sym forward):", modes)
    for (b, d, tau), (peer, sym) in [(c, m) for c in ((192, 128, 0.5), (1000, 64, 0.1), (512, 256, 0.2)) for m in modes]:
        os.environ["MAAI_FWD_SYM_MULTI"] = sym
        g = torch.Generator().manual_seed(77)
        H1 = torch.randn(world * b, d, generator=g)
        H2 = H1 + 0.5 * torch.randn(world * b, d, generator=g)
        res = {}
        for kg in (True, False, "reduce_scatter"):
            x = H1[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
            y = H2[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
            loss, _, _ = maai_b200.contrastive_loss(x, y, temperature=tau, local_rank=rank, world_size=world,
                                                    device=dev, key_grad=kg, peer_gather=peer)
            loss.backward()
            pack = torch.cat([loss.detach().reshape(1), x.grad.reshape(-1), y.grad.reshape(-1)])
            allp = [torch.empty_like(pack) for _ in range(world)]
            dist.all_gather(allp, pack)
            res[kg] = [p.cpu().numpy() for p in allp]
        if rank == 0:
            h1r = [H1[p * b:(p + 1) * b].numpy() for p in range(world)]
            h2r = [H2[p * b:(p + 1) * b].numpy() for p in range(world)]
            for kg in (True, False, "reduce_scatter"):  # the reduce-scatter dataflow yields the full gradient
                ol, o1, o2 = O.contrastive_loss_oracle_distributed(h1r, h2r, tau, key_grad=bool(kg))
                for p in range(world):
                    l = res[kg][p][0]; g1 = res[kg][p][1:1 + b * d].reshape(b, d); g2 = res[kg][p][1 + b * d:].reshape(b, d)
                    e = (abs(l - ol[p]) / abs(ol[p]), np.linalg.norm(g1 - o1[p]) / np.linalg.norm(o1[p]),
                         np.linalg.norm(g2 - o2[p]) / np.linalg.norm(o2[p]))
                    print(f"b={b} d={d} tau={tau} peer={peer} sym={sym} key_grad={kg} rank={p}: loss rel {e[0]:.2e} dh1 {e[1]:.2e} dh2 {e[2]:.2e}")
                    ok = ok and e[0] <= 1e-3 and e[1] <= 1e-2 and e[2] <= 1e-2
            # full gradient / W == single-process reference on the concatenated batch
            gl, s1, s2 = O.contrastive_loss_oracle(H1.numpy(), H2.numpy(), tau)
            full1 = np.concatenate([res[True][p][1:1 + b * d].reshape(b, d) for p in range(world)]) / world
            e = np.linalg.norm(full1 - s1) / np.linalg.norm(s1)
            ml = np.mean([res[True][p][0] for p in range(world)])
            print(f"   vs single-process global batch: loss rel {abs(ml - gl) / gl:.2e} dh1 rel {e:.2e}")
            ok = ok and e <= 1e-2 and abs(ml - gl) / gl <= 1e-3
    if rank == 0:
        print("DIST_OK" if ok else "DIST_FAIL")
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
