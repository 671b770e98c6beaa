"""Host-side logic of the forward's tile schedule, checked without a GPU through the library's host-only
entry points (the same functions the kernels / launcher use):

* ``maai_debug_tri_locate``: the folded triangular item list visits every (row block, key tile) on or
  above the diagonal exactly once, segment by segment;
* ``maai_debug_group_plan``: over all ranks, the anchor groups of the cross-rank symmetric forward cover
  every unordered pair of global rows exactly once (so every row sum is complete and nothing is counted
  twice), and the per-group item counts add up.
"""
import ctypes

import numpy as np
import pytest


def _lib():
    from maai_b200 import _lib
    return _lib, _lib.load()


@pytest.mark.parametrize("T,nq", [(1, 1), (1, 2), (2, 2), (3, 2), (7, 1), (8, 2), (9, 2), (64, 2), (65, 2), (33, 1), (512, 2)])
def test_folded_triangle_visits_every_tile_once(T, nq):
    L, lib = _lib()
    nrb = (T + nq - 1) // nq
    total = nrb * T - nq * nrb * (nrb - 1) // 2
    rb, off, cnt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    seen = set()
    it = 0
    segments = []
    while it < total:
        L.check(lib.maai_debug_tri_locate(it, T, nrb, nq, ctypes.byref(rb), ctypes.byref(off), ctypes.byref(cnt)), "tri")
        assert 0 <= rb.value < nrb and cnt.value == T - nq * rb.value and 0 <= off.value < cnt.value
        # the kernel consumes the rest of the segment in one go
        for o in range(off.value, cnt.value):
            key = (rb.value, nq * rb.value + o)
            assert key not in seen
            seen.add(key)
        segments.append((rb.value, cnt.value - off.value))
        it += cnt.value - off.value
    assert it == total
    assert seen == {(r, k) for r in range(nrb) for k in range(nq * r, T)}
    # folded order: consecutive pairs of segments hold the same number of items
    if nrb >= 4:
        pair = [segments[i][1] + segments[i + 1][1] for i in range(0, 2 * (nrb // 2), 2)]
        assert len(set(pair)) == 1
    # any item index inside a segment resolves to the same row block
    for probe in np.linspace(0, total - 1, 17).astype(int):
        L.check(lib.maai_debug_tri_locate(int(probe), T, nrb, nq, ctypes.byref(rb), ctypes.byref(off), ctypes.byref(cnt)), "tri")
        assert (rb.value, nq * rb.value + off.value) in seen


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8, 16])
@pytest.mark.parametrize("b,d_pad", [(64, 128), (96, 128), (200, 64), (130, 256), (512, 128)])
def test_group_plan_covers_every_pair_once(world, b, d_pad):
    L, lib = _lib()
    m = 2 * b
    M = m * world
    nq = 1 if d_pad == 256 else 2
    rb_rows = 128 * nq
    T = (m + 127) // 128
    count = np.zeros((M, M), dtype=np.int32)
    for rank in range(world):
        ng = ctypes.c_int()
        q0 = (ctypes.c_int * 9)(); rows = (ctypes.c_int * 9)(); nkt = (ctypes.c_int * 9)()
        items = (ctypes.c_longlong * 9)()
        L.check(lib.maai_debug_group_plan(b, world, rank, d_pad, ctypes.byref(ng), q0, rows, nkt, items), "plan")
        assert 1 <= ng.value <= 9
        # group 0: own block, triangular -> every pair inside the slot once (both orders by symmetry)
        assert (q0[0], rows[0], nkt[0]) == (rank * m, m, T)
        nrb = (m + rb_rows - 1) // rb_rows
        assert items[0] == nrb * T - nq * nrb * (nrb - 1) // 2
        lo = rank * m
        count[lo:lo + m, lo:lo + m] += 1
        for g in range(1, ng.value):
            a0, ar, kt = q0[g], rows[g], nkt[g]
            assert ar > 0 and 0 < kt <= T and a0 // m != rank and (a0 % m) + ar <= m  # inside ONE other slot
            assert items[g] == ((ar + rb_rows - 1) // rb_rows) * kt
            k1 = min(kt * 128, m)  # local keys [0, k1)
            # E_ak serves anchor a's row sum (staged for its owner) and key k's row sum (local column sum)
            count[a0:a0 + ar, lo:lo + k1] += 1
            count[lo:lo + k1, a0:a0 + ar] += 1
    assert (count == 1).all(), np.argwhere(count != 1)[:5]


def test_group_plan_rejects_bad_arguments():
    L, lib = _lib()
    ng = ctypes.c_int()
    q0 = (ctypes.c_int * 9)(); rows = (ctypes.c_int * 9)(); nkt = (ctypes.c_int * 9)(); items = (ctypes.c_longlong * 9)()
    assert lib.maai_debug_group_plan(64, 17, 0, 128, ctypes.byref(ng), q0, rows, nkt, items) == -2
    assert lib.maai_debug_group_plan(64, 4, 4, 128, ctypes.byref(ng), q0, rows, nkt, items) == -1
    assert lib.maai_debug_group_plan(64, 4, 0, 100, ctypes.byref(ng), q0, rows, nkt, items) == -2


def test_cross_rank_symmetric_forward_switch(monkeypatch):
    """Every rank must take the same decision from (b, d_pad, world) and the environment alone."""
    from maai_b200.Objective import _sym_forward_enabled
    monkeypatch.delenv("MAAI_FWD_SYM_MULTI", raising=False)
    assert _sym_forward_enabled(16384, 128, 2) and _sym_forward_enabled(8192, 128, 4)
    assert not _sym_forward_enabled(4096, 128, 8)
    assert not _sym_forward_enabled(1 << 20, 128, 17)  # group table of the kernel: world <= 16
    monkeypatch.setenv("MAAI_FWD_SYM_MULTI", "1")
    assert _sym_forward_enabled(64, 128, 8) and not _sym_forward_enabled(64, 128, 32)
    monkeypatch.setenv("MAAI_FWD_SYM_MULTI", "0")
    assert not _sym_forward_enabled(16384, 128, 2)


def test_peer_buffer_set_reuse_policy():
    """PeerWorkspace's reuse rule as pure host logic (SetReusePolicy): which issue orders of forwards / backwards
    are free, which need the extra barrier, which must raise (a rank racing ahead must never overwrite
    a set another rank's backward still reads)."""
    import pytest
    from maai_b200.Objective import SetReusePolicy
    # f b f b ...: never an extra barrier, sets in turn
    p = SetReusePolicy(3)
    for t in range(9):
        i, extra = p.next_set(True)
        assert (i, extra) == (t % 3, False)
        p.backward_issued(i)
    # two in flight: f0 f1 b1 b0 f2 f3 b3 b2 f4 f5 b5 b4 -- what three sets buy: still no extra barrier
    p = SetReusePolicy(3)
    for rep in range(4):
        (a, ea), (b, eb) = p.next_set(True), p.next_set(True)
        assert not ea and not eb
        p.backward_issued(b); p.backward_issued(a)
    # three in flight: the fourth forward must raise; after the backwards the set is reusable, but its backward
    # came after the previous forward was issued -> extra barrier
    p = SetReusePolicy(3)
    sets = [p.next_set(True)[0] for _ in range(3)]
    with pytest.raises(RuntimeError):
        p.next_set(True)
    assert p.step == 3  # the refused forward was not counted
    for i in reversed(sets):
        p.backward_issued(i)
    i, extra = p.next_set(True)
    assert i == 0 and extra
    p.backward_issued(i)
    i, extra = p.next_set(True)   # set 1: its backward was issued before the previous forward (step 3) -> free
    assert i == 1 and not extra
    # forward-only calls (no_grad) never block a set
    p = SetReusePolicy(3)
    for t in range(10):
        assert p.next_set(False) == (t % 3, False)
    # with two sets (round 1) the advisor's pattern f0 f1 b1 b0 f2 needs the extra barrier
    p = SetReusePolicy(2)
    a, b = p.next_set(True)[0], p.next_set(True)[0]
    p.backward_issued(b); p.backward_issued(a)
    assert p.next_set(True) == (0, True)


def test_peer_workspace_eviction_spares_pending_backwards():
    """A cached peer workspace is evicted oldest-first, but never while one of its sets waits for a backward (the C++
    binding's backward holds raw addresses into it)."""
    from types import SimpleNamespace
    from maai_b200.Objective import PeerWorkspace, SetReusePolicy
    mk = lambda: SimpleNamespace(policy=SetReusePolicy(3))
    cache = {k: mk() for k in "abcd"}
    i, _ = cache["a"].policy.next_set(True)      # a: forward issued, backward outstanding
    PeerWorkspace._evict(cache, 4)
    assert list(cache) == ["a", "c", "d"]        # b (oldest idle) went, a stayed
    cache["a"].policy.backward_issued(i)
    cache["e"] = mk()
    PeerWorkspace._evict(cache, 4)
    assert list(cache) == ["c", "d", "e"]
    for w in cache.values():                     # everything pending: nothing is evicted, the cache grows by one
        w.policy.next_set(True)
    cache["f"] = mk(); cache["f"].policy.next_set(True)
    PeerWorkspace._evict(cache, 4)
    assert list(cache) == ["c", "d", "e", "f"]
    # the C++ backward's direct writes (state[1+set] = 0, state[1+nbuf+set] = state[0]) are backward_issued
    p, q = SetReusePolicy(3), SetReusePolicy(3)
    for pol in (p, q):
        pol.next_set(True); pol.next_set(True)
    p.backward_issued(1)
    q.state[1 + 1] = 0; q.state[1 + 3 + 1] = q.state[0]
    assert (p.state == q.state).all() and p.any_pending() and q.state.dtype.itemsize == 8
