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
    modes = [(False, "0")] + ([(True, "0"), (True, "1"), (True, "direct")] if peer_gather_available() else [])
    if rank == 0:
        print("modes under test (peer: False = NCCL all_gather, True = fused NVLink peer stores; cross-rank symmetric "
              "forward: 0 = off, 1 = staged, direct = one launch with remote adds):", modes,
              "| ordering of the peer stores:", "barrier launches" if os.environ.get("MAAI_PEER_FLAGS") == "0" else "in-kernel flags")
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
    # ---- peer workspace reuse (PeerWorkspace docstring): several forwards in flight before their backwards.
    # Pattern f0 f1 b1 b0 f2 f3 b3 b2 f4 b4 with different inputs per step; every step's gradient must match
    # the oracle (a rank racing ahead and overwriting a buffer another rank's backward still reads shows up
    # as a wrong gradient); a third forward in flight must raise instead of overwriting.
    if peer_gather_available():
        os.environ["MAAI_FWD_SYM_MULTI"] = "0"
        b, d, tau = 384, 128, 0.3

        def make(step):
            g = torch.Generator().manual_seed(1000 + step)
            H1 = torch.randn(world * b, d, generator=g)
            H2 = H1 + 0.5 * torch.randn(world * b, d, generator=g)
            x = H1[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
            y = H2[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
            return H1, H2, x, y

        def fwd(step):
            H1, H2, x, y = make(step)
            loss = maai_b200.contrastive_loss(x, y, temperature=tau, local_rank=rank, world_size=world, device=dev,
                                              key_grad=True, peer_gather=True)[0]
            return dict(H1=H1, H2=H2, x=x, y=y, loss=loss)

        def check(st, tag):
            nonlocal ok
            _, o1, o2 = O.contrastive_loss_oracle_distributed(
                [st["H1"][p * b:(p + 1) * b].numpy() for p in range(world)],
                [st["H2"][p * b:(p + 1) * b].numpy() for p in range(world)], tau, key_grad=True)
            e1 = np.linalg.norm(st["x"].grad.cpu().numpy() - o1[rank]) / np.linalg.norm(o1[rank])
            e2 = np.linalg.norm(st["y"].grad.cpu().numpy() - o2[rank]) / np.linalg.norm(o2[rank])
            flag = torch.tensor([max(e1, e2)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"workspace reuse {tag}: max dh rel err over ranks {float(flag):.2e}")
            ok = ok and float(flag) <= 1e-2

        for rep in range(3):
            s0, s1 = fwd(10 * rep), fwd(10 * rep + 1)
            if rank == world - 1:
                torch.cuda._sleep(20_000_000)   # this rank's backwards run late: the others race ahead
            s1["loss"].backward(); s0["loss"].backward()
            s2, s3 = fwd(10 * rep + 2), fwd(10 * rep + 3)
            s3["loss"].backward(); s2["loss"].backward()
            for k, st in enumerate((s0, s1, s2, s3)):
                check(st, f"rep {rep} two-in-flight step {k}")
        # three in flight: the third forward reuses ... no, takes the third set; the fourth must raise
        a, b_, c = fwd(100), fwd(101), fwd(102)
        raised = False
        try:
            fwd(103)
        except RuntimeError as e:
            raised = "in flight" in str(e)
        c["loss"].backward(); b_["loss"].backward(); a["loss"].backward()
        nxt = fwd(104)   # set whose backward was issued after the previous forward: extra barrier path
        nxt["loss"].backward()
        for k, st in enumerate((a, b_, c, nxt)):
            check(st, f"three-in-flight step {k}")
        if rank == 0:
            print("fourth forward in flight raised:", raised)
        ok = ok and raised
    # ---- world > 1 convergence parity of the DEFAULT gradient semantics: a small encoder trained for
    # 40 SGD steps (a) on W ranks with the fused loss, key_grad=True, parameter gradients averaged over the ranks
    # as DDP does, and (b) in a single process on the concatenated batch with the reference's own formulation
    # (oracle/ref_runner.reference_loss: the unmodified file when present).  Same initial weights, same data:
    # the mean of the rank losses must follow the reference's loss curve and the weights must end up together.
    from oracle.ref_runner import reference_loss
    b, din, dout, tau, steps, lr = 256, 48, 32, 0.2, 40, 0.5

    def make_net():
        torch.manual_seed(7)
        return torch.nn.Sequential(torch.nn.Linear(din, 96), torch.nn.ReLU(), torch.nn.Linear(96, dout)).to(dev)

    g = torch.Generator().manual_seed(11)
    X1 = torch.randn(steps, world * b, din, generator=g)
    X2 = X1 + 0.4 * torch.randn(steps, world * b, din, generator=g)
    net_f, net_r = make_net(), make_net()
    opt_f = torch.optim.SGD(net_f.parameters(), lr=lr)
    opt_r = torch.optim.SGD(net_r.parameters(), lr=lr)
    curve_f, curve_r = [], []
    for t in range(steps):
        xa = X1[t, rank * b:(rank + 1) * b].to(dev); xb = X2[t, rank * b:(rank + 1) * b].to(dev)
        loss = maai_b200.contrastive_loss(net_f(xa), net_f(xb), temperature=tau, local_rank=rank, world_size=world,
                                          device=dev, key_grad=True)[0]
        opt_f.zero_grad(); loss.backward()
        for prm in net_f.parameters():
            dist.all_reduce(prm.grad); prm.grad /= world        # DDP's gradient averaging
        opt_f.step()
        lm = loss.detach().clone(); dist.all_reduce(lm); curve_f.append(float(lm) / world)
        lr_ = reference_loss(net_r(X1[t].to(dev)), net_r(X2[t].to(dev)), tau)   # single process, global batch
        opt_r.zero_grad(); lr_.backward(); opt_r.step()
        curve_r.append(float(lr_.detach()))
    curve_f, curve_r = np.array(curve_f), np.array(curve_r)
    dcurve = float(np.abs(curve_f - curve_r).max() / np.abs(curve_r).max())
    dw = max(float((pf - pr).norm() / pr.norm()) for pf, pr in zip(net_f.parameters(), net_r.parameters()))
    if rank == 0:
        print(f"convergence parity, default full gradient + DDP averaging vs single-process reference on the global batch: "
              f"loss {curve_r[0]:.4f} -> {curve_r[-1]:.4f}, max curve diff {dcurve:.2e}, max weight diff {dw:.2e}")
    ok = ok and dcurve <= 5e-3 and dw <= 2e-2
    # ---- chained views across ranks (NTXentLoss(chain_views=True)): half the gather payload, same results
    for peer in ([False, True] if peer_gather_available() else [False]):
        b, d, tau = 320, 128, 0.4
        outs = []
        for t in range(4):
            g = torch.Generator().manual_seed(500 + t)
            outs.append(torch.randn(world * b, d, generator=g))
        got = {}
        for chained in (True, False):
            mod = maai_b200.NTXentLoss(temperature=tau, local_rank=rank, world_size=world, key_grad=True,
                                       chain_views=chained, peer_gather=peer)
            o1 = outs[0][rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
            rows = []
            for t in range(1, 4):
                o2 = outs[t][rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
                loss = mod(o1.data, o2)
                loss.backward()
                rows.append((float(loss.detach()), o2.grad.clone()))
                o1 = o2
            got[chained] = rows
            assert mod.chained_steps == (2 if chained else 0), mod.chained_steps
        worst = 0.0
        for (lc, gc), (lu, gu) in zip(got[True], got[False]):
            worst = max(worst, abs(lc - lu) / abs(lu), float((gc - gu).norm() / gu.norm()))
        # and against the oracle: hidden1 detached, full gradient w.r.t. hidden2
        _, _, o2f = O.contrastive_loss_oracle_distributed(
            [outs[2][p * b:(p + 1) * b].numpy() for p in range(world)],
            [outs[3][p * b:(p + 1) * b].numpy() for p in range(world)], tau, key_grad=True)
        eo = float(np.linalg.norm(got[True][2][1].cpu().numpy() - o2f[rank]) / np.linalg.norm(o2f[rank]))
        flag = torch.tensor([worst, eo], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"chained views peer={peer}: chained vs unchained {float(flag[0]):.2e}, vs oracle dh2 {float(flag[1]):.2e}")
        ok = ok and float(flag[0]) <= 1e-5 and float(flag[1]) <= 1e-2
    if rank == 0:
        print("DIST_OK" if ok else "DIST_FAIL")
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
