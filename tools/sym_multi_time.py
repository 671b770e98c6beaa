"""Times the tile part of the cross-rank symmetric forward (maai_ntxent_fwd_sym_tiles) for every
emulated rank on ONE GPU against the full forward (maai_ntxent_fwd) of the same rank.
    python tools/sym_multi_time.py [pairs_global] [dim] [worlds...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maai_b200  # noqa: E402,F401
from maai_b200 import _lib  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    worlds = [int(x) for x in sys.argv[3:]] or [2, 8]
    lib = _lib.load()
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    dp = lib.maai_padded_dim(d)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, reps=15):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        return sorted(ts)[len(ts) // 2]

    for world in worlds:
        b = B // world
        z = torch.randn(world, 2 * b, dp, device=dev)
        z = (z / z.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
        cos = torch.zeros(b, device=dev)
        rowsum = torch.zeros(2 * b, device=dev)
        r = torch.zeros(2 * b, device=dev)
        loss = torch.zeros((), device=dev)
        stage = torch.zeros(world, 2 * b, device=dev)
        for p in sorted({0, world // 2, world - 1}):
            t_full = timeit(lambda: _lib.check(lib.maai_ntxent_fwd(z.data_ptr(), b, world, p, dp, 2.0, cos.data_ptr(),
                                                                   rowsum.data_ptr(), r.data_ptr(), loss.data_ptr(), 0, None, s), "f"))
            t_sym = timeit(lambda: _lib.check(lib.maai_ntxent_fwd_sym_tiles(z.data_ptr(), b, world, p, dp, 2.0,
                                                                            rowsum.data_ptr(), stage.data_ptr(), 0, None, s), "s"))
            print(f"world={world} rank={p} b={b} d={d}: full fwd (+finalize) {t_full:.4f} ms, sym tiles {t_sym:.4f} ms", flush=True)
        if world == worlds[0]:
            z1 = z.reshape(1, -1, dp)
            bb = B
            cos1 = torch.zeros(bb, device=dev); rs1 = torch.zeros(2 * bb, device=dev); r1 = torch.zeros(2 * bb, device=dev)
            t1 = timeit(lambda: _lib.check(lib.maai_ntxent_fwd(z1.data_ptr(), bb, 1, 0, dp, 2.0, cos1.data_ptr(),
                                                               rs1.data_ptr(), r1.data_ptr(), loss.data_ptr(), 0, None, s), "f1"))
            print(f"world=1 b={bb}: single-rank symmetric forward (+finalize) {t1:.4f} ms", flush=True)


if __name__ == "__main__":
    main()
