"""Golden values for the Model_Util mirrors from the *imported reference* (build container only):

    python tests/golden/make_model_util_golden.py

/root/reference/SimCLR/Model_Util.py imports ``apex.parallel.LARC`` at module level (not installable
here); an empty stand-in module is registered for that one import so that the file itself is executed
unmodified.  Written: tests/golden/model_util_golden.npz -- learning-rate curves of
``learning_rate_schedule`` (Model_Util.py:9-39) driven by an Adam optimiser over a whole (short) run,
both scalings, with and without warm-up, and ``top_k_accuracy`` (Model_Util.py:104-113) on seeded scores
with index and one-hot targets."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/SimCLR"

LR_CASES = {  # name -> arguments (without the optimiser)
    "linear_warm": dict(warmup_epochs=2, num_examples=1000, batch_size=50, world_size=4, learning_rate_scaling="linear",
                        base_learning_rate=0.3, train_epochs=6),
    "sqrt_warm": dict(warmup_epochs=1, num_examples=640, batch_size=64, world_size=8, learning_rate_scaling="sqrt",
                      base_learning_rate=0.075, train_epochs=5),
    "linear_nowarm": dict(warmup_epochs=0, num_examples=500, batch_size=100, world_size=1, learning_rate_scaling="linear",
                          base_learning_rate=0.1, train_epochs=4),
}


def lr_curve(fn, kw, steps):
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    opt = torch.optim.Adam(lin.parameters(), 1e-3)
    args = dict(kw, optimizer=opt)
    out = []
    for _ in range(steps):
        fn(args)
        out.append(opt.param_groups[0]["lr"])
        opt.zero_grad()
        lin(torch.ones(2, 4)).sum().backward()
        opt.step()
    return np.array(out)


def main():
    apex = types.ModuleType("apex")
    apex.parallel = types.ModuleType("apex.parallel")
    apex.parallel.LARC = types.ModuleType("apex.parallel.LARC")
    sys.modules.update({"apex": apex, "apex.parallel": apex.parallel, "apex.parallel.LARC": apex.parallel.LARC})
    sys.path.insert(0, REF)
    import Model_Util as R  # the unmodified reference file
    out = {}
    for name, kw in LR_CASES.items():
        steps = kw["num_examples"] * kw["train_epochs"] // kw["batch_size"] + 5
        out[f"lr.{name}"] = lr_curve(R.learning_rate_schedule, kw, steps)
    g = torch.Generator().manual_seed(3)
    preds = torch.randn(200, 50, generator=g)
    tgt = torch.randint(0, 50, (200,), generator=g)
    out["topk.preds"] = preds.numpy()
    out["topk.target"] = tgt.numpy()
    for k in (1, 5, 10):
        out[f"topk.idx.k{k}"] = np.array(float(R.top_k_accuracy(preds, tgt, k)))
        out[f"topk.onehot.k{k}"] = np.array(float(R.top_k_accuracy(preds, torch.nn.functional.one_hot(tgt, 50), k)))
    np.savez_compressed(os.path.join(HERE, "model_util_golden.npz"), **out)
    print({k: (v.shape if v.ndim else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
