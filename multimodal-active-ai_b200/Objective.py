"""B200-native drop-in for /root/reference/SimCLR/Objective.py.

``contrastive_loss`` keeps the reference signature and return tuple (Objective.py:17-22, :81):

    loss, logits_ab, labels = contrastive_loss(hidden1, hidden2, hidden_norm=True, temperature=1.0,
                                               local_rank=0, world_size=1, device='cpu')

but the work is done by hand-written sm_100a kernels reached through the C ABI in
include/maai_ntxent.h: one normalise+cast pass, one fused tcgen05 stripe kernel that never writes
a logit to HBM, and a recompute-based backward.  PyTorch is only plumbing here (device memory,
streams, autograd hook, torch.distributed for the embedding all-gather).

Differences from the reference, all deliberate:
  * ``logits_ab`` / ``labels`` (used only by validate(), Contrastive_Learning.py:860-868) are
    produced only when autograd is disabled or ``return_logits=True``; the training path returns
    ``None`` for them instead of materialising (b, B) fp32 + (b, 2B) int64 tensors every step.
  * ``world_size > 1``: the reference's ``dist.all_gather`` is not differentiable, so it silently
    drops the key-side gradient (Objective.py:112-114).  Here the gradient is the full one (query
    side + key side) unless ``key_grad=False`` is passed.
  * ``hidden_norm=False`` is not supported (no caller uses it; the fixed-maximum logsumexp needs
    unit-norm rows) and raises ``NotImplementedError``.
  * CPU tensors raise: there is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib

LARGE_NUM = 1e9  # kept for API parity with Objective.py:6 (the fused kernels mask by predicate)
MIN_TEMPERATURE = 0.025  # fixed-maximum logsumexp: exp((cos-1)/tau) must stay a normal fp32 for cos >= -1

_DTYPES = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16, torch.float16: _lib.DT_F16}


class _Profiler:
    """Optional CUDA-event brackets around the C-ABI calls (bench.py's live per-kernel timing)."""
    enabled = False
    events = {"normalize": [], "gather_z": [], "fwd": [], "gather_l": [], "finalize": [], "gather_r": [], "bwd": [],
              "bwd_keyside": [], "reduce_scatter_wait": [], "bwd_dh": []}

    _ext = None      # the C++ binding keeps its own brackets (same names) once a call has gone through it
    _ext_on = False

    @classmethod
    def reset(cls):
        for v in cls.events.values():
            v.clear()
        if cls._ext is not None:
            cls._ext.span_reset()

    @classmethod
    def sync_ext(cls, ext):
        if cls._ext is None:
            cls._ext = ext
        if cls._ext_on != cls.enabled:
            ext.span_timing(bool(cls.enabled))
            cls._ext_on = bool(cls.enabled)

    @classmethod
    def collect_ms(cls):
        """{span name: [ms per bracketed call]} over the Python path's and the C++ binding's brackets (synchronises)."""
        out = {}
        for k, v in cls.events.items():
            if v:
                v[-1][1].synchronize()
            out[k] = [a.elapsed_time(e) for a, e in v]
        if cls._ext is not None:
            for k, ms in zip(("normalize", "fwd", "bwd"), cls._ext.span_read()):
                out[k] = out[k] + list(ms)
        return out

    class span:
        def __init__(self, name):
            self.name = name

        def __enter__(self):
            if _Profiler.enabled:
                self.a = torch.cuda.Event(enable_timing=True)
                self.b = torch.cuda.Event(enable_timing=True)
                self.a.record()
            return self

        def __exit__(self, *exc):
            if _Profiler.enabled:
                self.b.record()
                _Profiler.events[self.name].append((self.a, self.b))
            return False


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(dev: Optional[torch.device] = None) -> int:
    # raw handle of the current stream of the tensors' device (torch.cuda.current_stream() builds a
    # Python Stream object and costs ~10 us per call; three calls per step were 1/4 of the host time
    # at small batches)
    idx = torch.cuda.current_device() if dev is None or dev.index is None else dev.index
    try:
        return torch._C._cuda_getCurrentRawStream(idx)
    except AttributeError:  # pragma: no cover - older / newer torch without the private hook
        return torch.cuda.current_stream(idx).cuda_stream


class _on_device:
    """The C ABI launches on the CURRENT CUDA device: make that the tensors' device for the duration of
    the calls (a no-op, and no context-manager cost, in the one-process-per-GPU case)."""

    def __init__(self, dev: torch.device):
        self.guard = None
        if dev.index is not None and dev.index != torch.cuda.current_device():
            self.guard = torch.cuda.device(dev)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()
        return self

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


def _step_workspace(lib, b: int, dp: int, need_bwd: bool, dev):
    """The step's fp32 scratch (include/maai_ntxent.h "step workspace"): row sums, control words and -- if a
    backward will follow -- the dz accumulator, ONE allocation that K1 zero-fills on the side, so the step
    has no memset / zero kernels and the forward finalizes inside its tile kernel."""
    nbytes = lib.maai_ntxent_workspace_bytes(b, dp, 1 if need_bwd else 0)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    head = (2 * b + _lib.WS_CTL_WORDS + 127) // 128 * 128
    dz_acc = ws[head:head + 2 * b * dp].view(2 * b, dp) if need_bwd else None
    return ws, ws[:2 * b], dz_acc


def padded_dim(d: int) -> int:
    dp = _lib.load().maai_padded_dim(int(d))
    if dp < 0:
        raise ValueError(f"embedding dim {d} unsupported: the sm_100a tile kernels take 1 <= d <= 256")
    return dp


def gather_rows(z_all: torch.Tensor, rank: int, group=None) -> torch.Tensor:
    """All-gather of each rank's stacked (2b, d_pad) block of normalised rows, in place into the
    (world, 2b, d_pad) buffer whose slot ``rank`` the caller has filled.  Replaces the two fp32 list
    all_gathers + cat of _cross_replica_concat (Objective.py:52-53, 102-114) by ONE collective on
    bf16 rows.  Global row order is rank-major: (p, view, k) -> p*2b + view*b + k."""
    world = z_all.shape[0]
    if world > 1:
        dist.all_gather_into_tensor(z_all.view(-1), z_all[rank].reshape(-1), group=group)
    return z_all


def gather_row_factors(r_col: torch.Tensor, rank: int, b: int, world: int, group=None) -> torch.Tensor:
    """All-gather of the per-anchor factors r = 1/(b (e_pos + l')) (2b fp32 per rank), in place:
    slot ``rank`` of ``r_col`` (length >= world*2b) already holds this rank's values.  This is the
    only backward-side exchange of the full-gradient path (symmetry identity)."""
    if world > 1:
        dist.all_gather_into_tensor(r_col[:world * 2 * b], r_col[rank * 2 * b:(rank + 1) * 2 * b],
                                    group=group)
    return r_col


class SetReusePolicy:
    """Which of NBUF peer-written buffer sets a forward may take, decided from the local issue order of forwards
    and backwards (pure host logic; PeerWorkspace's docstring has the argument).  Every rank runs the same program,
    so every rank reaches the same decisions.  The state lives in one int64 array -- [0] forwards issued so far,
    [1 .. nbuf] "the set's last forward still waits for its backward", [1 + nbuf ..] forwards issued at the time
    that backward was issued (-1: none yet) -- so that the C++ binding's backward (csrc/maai_torch_ext.cpp, which
    runs on the autograd engine's thread) can mark its set through the array's address."""

    def __init__(self, nbuf: int):
        import numpy as np
        self.nbuf = nbuf
        self.state = np.zeros(1 + 2 * nbuf, dtype=np.int64)
        self.state[1 + nbuf:] = -1

    @property
    def step(self) -> int:
        return int(self.state[0])

    def next_set(self, needs_bwd: bool):
        """Returns (set index, extra_barrier) for the forward being issued; raises when the set's previous
        backward is still outstanding.  extra_barrier: that backward was issued after the previous forward, so
        the peers' synchronisation of the previous step does not cover it: a barrier must precede the first store."""
        st, n = self.state, self.nbuf
        t = int(st[0])
        i = t % n
        if st[1 + i]:
            raise RuntimeError(
                f"maai NT-Xent: {n} forward passes through the peer-gather workspace are waiting for "
                f"their backward; at most {n - 1} may be in flight (the next forward would overwrite "
                "buffers a backward still reads, on this or another rank). Call backward first, or pass "
                "peer_gather=False to use the NCCL all-gather path")
        extra = bool(st[1 + n + i] >= t)
        st[0] = t + 1
        st[1 + i] = 1 if needs_bwd else 0
        st[1 + n + i] = -1
        return i, extra

    def any_pending(self) -> bool:
        return bool(self.state[1:1 + self.nbuf].any())

    def backward_issued(self, i: int):
        self.state[1 + i] = 0
        self.state[1 + self.nbuf + i] = self.state[0]


class PeerWorkspace:
    """Symmetric (peer-mapped over NVLink) buffers for the fused gathers of the multi-GPU path.

    Replaces the two collectives of the path -- the bf16 all-gather of the normalised rows
    (Objective.py:52-53, 102-114) and the fp32 all-gather of the row factors -- by stores issued
    from the producing kernels themselves (maai_ntxent_normalize_peer, maai_ntxent_fwd_peer) into
    every rank's buffer, followed by a symmetric-memory barrier (~7 us).  Memory comes from
    torch.distributed._symmetric_memory (plumbing: allocation, handle exchange, barrier).

    Buffer reuse across ranks.  NBUF = 3 sets are used in turn.  A peer may write set i (its K1 /
    finalize stores of step t) as soon as IT has passed the barriers of step t-1; the readers of the set's
    previous contents (step u = t-3) on THIS rank are the forward of u -- stream-ordered before this
    rank's barriers of step u+1 -- and the backward of u.  So the set may be reused iff every rank issued
    its backward of u before its last barrier of step t-1.  Every rank runs the same program (they meet in
    the barriers, or deadlock), so each rank decides from its own issue order, identically:
      * backward of u issued before the forward of t-1 (f b f b ..., or two steps in flight
        f0 f1 b1 b0 f2 f3 b3 b2: what three sets buy): nothing to do;
      * backward of u issued later, but before this forward: one extra barrier in front of K1 orders it;
      * backward of u still outstanding (more than NBUF-1 forwards in flight): RuntimeError at the forward
        -- never a silent overwrite; ``peer_gather=False`` (NCCL path, private buffers) has no limit.
    Forward-only calls (no_grad) never have a backward outstanding.
    """
    _cache = {}
    NBUF = 3

    def __init__(self, b, dp, world, rank, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        lib = _lib.load()
        self.b, self.dp, self.world, self.rank = b, dp, world, rank
        self.r_len = lib.maai_ntxent_r_len(b, world)
        zb = world * 2 * b * dp * 2                      # bytes of one key buffer
        rb = self.r_len * 4                              # bytes of one r array
        self.zb, self.rb = zb, rb
        align = lambda x: (x + 255) // 256 * 256
        self.off_z = [i * align(zb) for i in range(self.NBUF)]
        self.off_r = [self.NBUF * align(zb) + i * align(rb) for i in range(self.NBUF)]
        # staging vectors of the cross-rank symmetric forward: (world, 2b) fp32 partial row sums that
        # this rank computed for the other ranks' anchors (maai_ntxent_fwd_sym_tiles)
        sb = world * 2 * b * 4
        self.off_s = [self.NBUF * (align(zb) + align(rb)) + i * align(sb) for i in range(self.NBUF)]
        # flag block of the in-kernel peer synchronisation (maai_peer_sync) + one counter word
        self.off_f = self.NBUF * (align(zb) + align(rb) + align(sb))
        # step workspaces (row sums | control words | dz accumulator) in peer-mapped memory: in the one-launch
        # cross-rank symmetric forward the peers add their partial row sums straight into the owner's
        wb = lib.maai_ntxent_workspace_bytes(b, dp, 1)
        self.wb = wb
        self.off_w = [self.off_f + align(_lib.FLAG_WORDS * 4) + 256 + i * align(wb) for i in range(self.NBUF)]
        total = self.off_w[-1] + align(wb)
        self.raw = symm_mem.empty((total,), dtype=torch.uint8, device=device)
        self.raw.zero_()                                 # r padding must read as zero
        self.hdl = symm_mem.rendezvous(self.raw, group if group is not None else dist.group.WORLD)
        ptrs = list(self.hdl.buffer_ptrs)
        tab = lambda off: torch.tensor([p + off for p in ptrs], dtype=torch.int64, device=device)
        self.z_tab = [tab(o) for o in self.off_z]        # device arrays of peer base addresses
        self.r_tab = [tab(o) for o in self.off_r]
        self.s_tab = [tab(o) for o in self.off_s]
        self.stage = [self.raw[o:o + sb].view(torch.float32) for o in self.off_s]
        self.f_tab = tab(self.off_f)
        self.wsbuf = [self.raw[o:o + wb].view(torch.float32) for o in self.off_w]
        self.w_host = [(ctypes.c_void_p * world)(*[p + o for p in ptrs]) for o in self.off_w]  # HOST arrays of peer addresses
        self.head = (2 * b + _lib.WS_CTL_WORDS + 127) // 128 * 128
        self.flags = self.raw[self.off_f:self.off_f + _lib.FLAG_WORDS * 4].view(torch.int32)
        self.counter = self.raw[self.off_f + align(_lib.FLAG_WORDS * 4):][:4].view(torch.int32)
        # MAAI_PEER_FLAGS=0: order the fused gathers with symmetric-memory barrier launches instead (A/B runs)
        self.use_flags = os.environ.get("MAAI_PEER_FLAGS", "1") != "0"
        # how long a kernel waits for a late peer before it traps (a dead rank must not hang the others forever)
        self.timeout_s = int(os.environ.get("MAAI_PEER_TIMEOUT_S", "300"))
        self.seq = 0
        self.z = [self.raw[o:o + zb].view(torch.bfloat16).view(world, 2 * b, dp) for o in self.off_z]
        self.r = [self.raw[o:o + rb].view(torch.float32) for o in self.off_r]
        # NVSwitch multicast mapping of the same allocation (one store reaches every rank), if the
        # fabric offers it; MAAI_PEER_MULTICAST=0 forces per-peer unicast stores
        mc = 0
        try:
            if os.environ.get("MAAI_PEER_MULTICAST", "1") != "0":
                mc = int(self.hdl.multicast_ptr)
        except Exception:  # noqa: BLE001
            mc = 0
        self.mc_z = [mc + o if mc else None for o in self.off_z]
        self.mc_r = [mc + o if mc else None for o in self.off_r]
        self.policy = SetReusePolicy(self.NBUF)
        self._fast = None
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=0)                      # every rank's zero fill is done before first use

    MAX_CACHED = 4  # shapes kept alive (symmetric memory is not returned to the caching allocator)

    @classmethod
    def get(cls, b, dp, world, rank, device, group):
        key = (b, dp, world, rank, str(device), id(group))
        ws = cls._cache.get(key)
        if ws is None:
            cls._evict(cls._cache, cls.MAX_CACHED)
            ws = cls._cache[key] = cls(b, dp, world, rank, device, group)
        return ws

    @staticmethod
    def _evict(cache: dict, max_cached: int) -> None:
        """Drop the oldest workspaces down to max_cached - 1 entries, but never one whose backward is outstanding:
        the C++ binding's backward holds raw addresses of the set and of the policy state, not a reference.  Every
        rank sees the same sequence of shapes and of forwards / backwards, so all ranks evict the same entries."""
        while len(cache) >= max_cached:
            victim = next((k for k, w in cache.items() if not w.policy.any_pending()), None)
            if victim is None:
                return
            cache.pop(victim)

    def next_set(self, needs_bwd: bool):
        """Set for the forward being issued (see the class docstring): (set index, extra_barrier)."""
        return self.policy.next_set(needs_bwd)

    def fast_args(self, i: int, seq: int):
        """Argument vector of the C++ binding's multi-rank step (csrc/maai_torch_ext.cpp, enum PeerArg) for set i."""
        if self._fast is None:
            self._fast = [[self.rank, self.world, self.z[k].data_ptr(), self.z_tab[k].data_ptr(), self.mc_z[k] or 0,
                           self.r[k].data_ptr(), self.r_tab[k].data_ptr(), self.mc_r[k] or 0, self.f_tab.data_ptr(),
                           self.flags.data_ptr(), self.counter.data_ptr(), 0, self.timeout_s, k, self.NBUF,
                           int(self.policy.state.ctypes.data)] for k in range(self.NBUF)]
        a = self._fast[i]
        a[11] = seq
        return a

    def sync_for(self, seq: int):
        """maai_peer_sync for step `seq` (None when barrier launches are used instead)."""
        if not self.use_flags:
            return None
        return _lib.PeerSync(self.f_tab.data_ptr(), self.flags.data_ptr(), self.counter.data_ptr(), seq, self.timeout_s)

    def backward_issued(self, i: int):
        self.policy.backward_issued(i)


_peer_state = {"ok": None}


def _sym_forward_mode(b, dp, world, flags: bool) -> str:
    """Cross-rank symmetric forward (peer mode only): "off", "staged" (partial row sums staged locally, pulled by
    a finalize kernel: maai_ntxent_fwd_sym_tiles + _finalize) or "direct" (added straight into the owners' row
    sums over NVLink, per-row tail inside the one launch: maai_ntxent_fwd_sym_direct; needs the in-kernel flags).
    A tile that serves a row AND a column costs 1.5x a plain one, so halving the tile count buys 25 % of the tile
    time, against which stand the exchange of the partial sums and a per-row tail that waits for the peers.
    Measured at 32768 pairs, d=128 (profiles/r2_tuning_log.md): 2 ranks x 16384 pairs staged 1.273 ms/step, direct
    1.288, off 1.36 -> staged from 16384 rows per rank up; 8 ranks x 4096 pairs: see the tuning log.
    MAAI_FWD_SYM_MULTI = 0 / staged / direct forces a mode (1 = staged); at most 16 ranks (group table)."""
    v = os.environ.get("MAAI_FWD_SYM_MULTI", "")
    if v == "0" or world > 16:
        return "off"
    if v == "direct":
        return "direct" if flags else "staged"
    if v in ("1", "staged"):
        return "staged"
    return "staged" if 2 * b >= 16384 else "off"


def _sym_forward_enabled(b, dp, world) -> bool:
    return _sym_forward_mode(b, dp, world, True) != "off"


def peer_gather_available() -> bool:
    """True when torch symmetric memory can be used for the fused gathers (probed once)."""
    if _peer_state["ok"] is None:
        ok = False
        if os.environ.get("MAAI_PEER_GATHER", "1") != "0" and dist.is_available() and dist.is_initialized() \
                and dist.get_backend() == "nccl":
            try:
                import torch.distributed._symmetric_memory  # noqa: F401
                ok = True
            except Exception:  # noqa: BLE001
                ok = False
        _peer_state["ok"] = ok
    return _peer_state["ok"]


def _peer_usable(b, dp, world, rank, device, group) -> bool:
    """Automatic mode: build the workspace for this shape once; every rank then votes (all_reduce MIN)
    so that either all of them take the peer path or all take the NCCL path.  A box on which symmetric
    memory cannot be set up (no P2P mapping, container restrictions) silently keeps the NCCL gathers."""
    key = ("usable", b, dp, world, rank, str(device), id(group))
    ok = _peer_state.get(key)
    if ok is None:
        good = 1
        try:
            PeerWorkspace.get(b, dp, world, rank, device, group)
        except Exception as e:  # noqa: BLE001
            good = 0
            if rank == 0:
                import warnings
                warnings.warn(f"maai NT-Xent: symmetric memory unavailable ({e!r}); using NCCL all_gather")
        flag = torch.tensor([good], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        ok = _peer_state[key] = bool(int(flag.item()))
    return ok


def positive_index(b: int, world: int) -> torch.Tensor:
    """Global row index of every global row's positive under the rank-major layout (labels_idx +
    rank*b of Objective.py:55 restated for stacked [view a; view b] blocks)."""
    i = torch.arange(2 * b * world)
    k = i % (2 * b)
    return torch.where(k < b, i + b, i - b)


class _NTXentFunction(torch.autograd.Function):
    """loss = NT-Xent(hidden1, hidden2) for this rank (Objective.py:79), autograd-compatible."""

    @staticmethod
    def forward(ctx, hidden1, hidden2, temperature, rank, world, group, key_grad, stash, peer=False, chain=None,
                carry=None):
        # chain: {"z_all", "inv_norm"} of the previous step when hidden1 IS that step's hidden2 (chained views,
        # maai_ntxent_normalize_chain); carry: dict that receives this step's {"z_all", "inv_norm"}
        lib = _lib.load()
        b, d = hidden1.shape
        dev = hidden1.device
        dp = padded_dim(d)
        h1 = hidden1.contiguous()
        h2 = hidden2.contiguous()
        dt = _DTYPES[h1.dtype]
        inv_tau = 1.0 / float(temperature)

        needs_grad = any(ctx.needs_input_grad[:2])
        # key_grad="reduce_scatter": full gradient through the key-side reduce-scatter dataflow
        # (backward below); the forward then needs no gather of the row factors at all
        rs = key_grad == "reduce_scatter" and world > 1
        full = (bool(key_grad) and not rs) or world == 1
        ctx.rs_group = group if rs else None
        ctx.rs = rs
        with _on_device(dev):
            if peer and world > 1:
                return _NTXentFunction._forward_peer(ctx, h1, h2, dt, inv_tau, rank, world, group, needs_grad,
                                                     full, stash, chain, carry)
            st = _stream(dev)
            z_all = torch.empty((world, 2 * b, dp), dtype=torch.bfloat16, device=dev)
            ws, rowsum, dz_acc = _step_workspace(lib, b, dp, needs_grad, dev)
            # one allocation for the small fp32 outputs (each torch.empty is ~3 us of host time)
            small = torch.empty(3 * b + 4, dtype=torch.float32, device=dev)
            inv_norm = small[:2 * b]
            pos_cos = small[2 * b:3 * b]
            loss = small[3 * b:3 * b + 1].view(())
            r_len = lib.maai_ntxent_r_len(b, world)
            r_col = torch.zeros(r_len, dtype=torch.float32, device=dev) if needs_grad else None
            if needs_grad and full:
                r_row = r_col[rank * 2 * b:(rank + 1) * 2 * b]  # this rank's slot of the gathered factors
            elif needs_grad:
                r_row = torch.empty(2 * b, dtype=torch.float32, device=dev)  # keys detached: r_col stays 0
            else:
                r_row = None

            with _Profiler.span("normalize"):
                if chain is not None:
                    _lib.check(lib.maai_ntxent_normalize_chain(_ptr(h2), b, d, dt, _ptr(chain["z_all"]),
                                                               _ptr(chain["inv_norm"]), _ptr(z_all), None, None, world,
                                                               rank, _ptr(inv_norm), _ptr(pos_cos), _ptr(ws),
                                                               ws.numel() * 4, None, st), "maai_ntxent_normalize_chain")
                else:
                    _lib.check(lib.maai_ntxent_normalize(_ptr(h1), _ptr(h2), b, d, dt, _ptr(z_all[rank]),
                                                         _ptr(inv_norm), _ptr(pos_cos), _ptr(ws), ws.numel() * 4, st),
                               "maai_ntxent_normalize")
            if carry is not None:
                carry["z_all"], carry["inv_norm"] = z_all, inv_norm
            with _Profiler.span("gather_z"):
                gather_rows(z_all, rank, group)
            with _Profiler.span("fwd"):
                _lib.check(lib.maai_ntxent_fwd(_ptr(z_all), b, world, rank, dp, inv_tau, _ptr(pos_cos),
                                               _ptr(rowsum), _ptr(r_row), _ptr(loss), _lib.F_PREZEROED, None, st),
                           "maai_ntxent_fwd")
            if needs_grad:
                if full:
                    with _Profiler.span("gather_r"):
                        gather_row_factors(r_col, rank, b, world, group)
                ctx.save_for_backward(h1, h2, z_all, inv_norm, r_row, r_col, pos_cos, rowsum)
                ctx.cfg = (b, d, dp, dt, inv_tau, rank, world, full)
                ctx.peer = None
                ctx.dz_acc = dz_acc
                ctx.acc_clean = True   # zero-filled by K1; a second backward (retain_graph) zeroes it itself
        if stash is not None:
            stash["z_all"] = z_all
            stash["rowsum"] = rowsum
            stash["pos_cos"] = pos_cos
        return loss

    @staticmethod
    def _forward_peer(ctx, h1, h2, dt, inv_tau, rank, world, group, needs_grad, full, stash, chain=None, carry=None):
        """world > 1 with the gathers fused into the producing kernels (PeerWorkspace).  Ordering across ranks:
        in-kernel flags (maai_peer_sync: the producers signal, the consumers' TMA producers wait, no barrier
        launch) or, with MAAI_PEER_FLAGS=0, symmetric-memory barrier launches between the calls."""
        lib = _lib.load()
        b, d = h1.shape
        dev = h1.device
        dp = padded_dim(d)
        st = _stream(dev)
        ws = PeerWorkspace.get(b, dp, world, rank, dev, group)
        i, extra_barrier = ws.next_set(needs_grad)
        ws.seq += 1
        sync = ws.sync_for(ws.seq)   # ctypes struct (kept alive until the calls below have returned) or None
        psync = ctypes.byref(sync) if sync is not None else None
        z_all = ws.z[i]
        sym_mode = _sym_forward_mode(b, dp, world, sync is not None)
        direct = sym_mode == "direct"
        if direct:  # the peers add into this rank's row sums: the step workspace lives in the symmetric allocation
            wsp = ws.wsbuf[i]
            rowsum = wsp[:2 * b]
            dz_acc = wsp[ws.head:ws.head + 2 * b * dp].view(2 * b, dp) if needs_grad else None
        else:
            wsp, rowsum, dz_acc = _step_workspace(lib, b, dp, needs_grad, dev)
        small = torch.empty(3 * b + 4, dtype=torch.float32, device=dev)
        inv_norm = small[:2 * b]
        pos_cos = small[2 * b:3 * b]
        loss = small[3 * b:3 * b + 1].view(())
        if extra_barrier:  # the previous user of this set ran its backward late: see PeerWorkspace
            ws.hdl.barrier(channel=0)
        with _Profiler.span("normalize"):
            if chain is not None:  # view-a halves: local copy from the previous set; only view b crosses NVLink
                _lib.check(lib.maai_ntxent_normalize_chain(_ptr(h2), b, d, dt, _ptr(chain["z_all"]),
                                                           _ptr(chain["inv_norm"]), _ptr(z_all), _ptr(ws.z_tab[i]),
                                                           ws.mc_z[i], world, rank, _ptr(inv_norm), _ptr(pos_cos),
                                                           _ptr(wsp), wsp.numel() * 4, psync, st),
                           "maai_ntxent_normalize_chain")
            else:
                _lib.check(lib.maai_ntxent_normalize_peer(_ptr(h1), _ptr(h2), b, d, dt, _ptr(ws.z_tab[i]), ws.mc_z[i],
                                                          world, rank, _ptr(inv_norm), _ptr(pos_cos), _ptr(wsp),
                                                          wsp.numel() * 4, psync, st),
                           "maai_ntxent_normalize_peer")
        if carry is not None:
            carry["z_all"], carry["inv_norm"] = z_all, inv_norm
        if sync is None:
            with _Profiler.span("gather_z"):
                ws.hdl.barrier(channel=0)  # every rank's rows have landed in every buffer
        r_row = r_col = None
        peer_r = needs_grad and full
        if not peer_r and needs_grad:  # keys detached (reference semantics): local row factors only, r_col = 0
            r_row = torch.empty(2 * b, dtype=torch.float32, device=dev)
            r_col = torch.zeros(ws.r_len, dtype=torch.float32, device=dev)
        if direct:
            # every pair of rank slots is computed once, in ONE launch: own block + the anchors of the ranks ahead
            # on the ring against the local keys; the partial row sums of those anchors are added straight into
            # their owners' row sums over NVLink, and the kernel's last CTA waits for the peers' signals and runs
            # the per-row tail (maai_ntxent_fwd_sym_direct)
            with _Profiler.span("fwd"):
                _lib.check(lib.maai_ntxent_fwd_sym_direct(_ptr(z_all), b, world, rank, dp, inv_tau, _ptr(pos_cos),
                                                          _ptr(rowsum), ws.w_host[i], _ptr(r_row),
                                                          _ptr(ws.r_tab[i]) if peer_r else None,
                                                          ws.mc_r[i] if peer_r else None, _ptr(loss), psync, st),
                           "maai_ntxent_fwd_sym_direct")
        elif sym_mode == "staged":
            # staged form: the other half of each row sum arrives through the peers' staging vectors (after a
            # barrier, or once their flags say so) and a separate finalize kernel pulls them
            with _Profiler.span("fwd"):
                _lib.check(lib.maai_ntxent_fwd_sym_tiles(_ptr(z_all), b, world, rank, dp, inv_tau, _ptr(rowsum),
                                                         _ptr(ws.stage[i]), _lib.F_PREZEROED, psync, st),
                           "maai_ntxent_fwd_sym_tiles")
            if sync is None:
                with _Profiler.span("gather_l"):
                    ws.hdl.barrier(channel=2)
            with _Profiler.span("finalize"):
                _lib.check(lib.maai_ntxent_fwd_sym_finalize(_ptr(rowsum), _ptr(ws.s_tab[i]), b, world, rank, inv_tau,
                                                            _ptr(pos_cos), _ptr(r_row),
                                                            _ptr(ws.r_tab[i]) if peer_r else None,
                                                            ws.mc_r[i] if peer_r else None, _ptr(loss), psync, st),
                           "maai_ntxent_fwd_sym_finalize")
        else:
            with _Profiler.span("fwd"):
                if peer_r:
                    _lib.check(lib.maai_ntxent_fwd_peer(_ptr(z_all), b, world, rank, dp, inv_tau, _ptr(pos_cos),
                                                        _ptr(rowsum), _ptr(ws.r_tab[i]), ws.mc_r[i], _ptr(loss),
                                                        _lib.F_PREZEROED, psync, st),
                               "maai_ntxent_fwd_peer")
                else:
                    _lib.check(lib.maai_ntxent_fwd(_ptr(z_all), b, world, rank, dp, inv_tau, _ptr(pos_cos),
                                                   _ptr(rowsum), _ptr(r_row), _ptr(loss), _lib.F_PREZEROED, psync, st),
                               "maai_ntxent_fwd")
        if needs_grad:
            if full:
                if sync is None:
                    with _Profiler.span("gather_r"):
                        ws.hdl.barrier(channel=1)  # every rank's row factors have landed everywhere
                r_col = ws.r[i]
                r_row = r_col[rank * 2 * b:(rank + 1) * 2 * b]
            ctx.save_for_backward(h1, h2, z_all, inv_norm, r_row, r_col, pos_cos, rowsum)
            ctx.cfg = (b, d, dp, dt, inv_tau, rank, world, full)
            bwd_flags = sync is not None and full
            if bwd_flags and os.environ.get("MAAI_PEER_FLAGS_BWD", "1") == "0":  # A/B: barrier launch instead
                ws.hdl.barrier(channel=1)
                bwd_flags = False
            ctx.peer = (ws, i, ws.seq if bwd_flags else 0)
            ctx.dz_acc = dz_acc
            ctx.acc_clean = True
        if stash is not None:
            stash["z_all"] = z_all
            stash["rowsum"] = rowsum
            stash["pos_cos"] = pos_cos
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _lib.load()
        h1, h2, z_all, inv_norm, r_row, r_col, pos_cos, rowsum = ctx.saved_tensors
        b, d, dp, dt, inv_tau, rank, world, full = ctx.cfg
        need = (1 if ctx.needs_input_grad[0] else 0) | (2 if ctx.needs_input_grad[1] else 0)
        dev = h1.device
        g = grad_loss.to(device=dev, dtype=torch.float32).contiguous()
        dh1 = torch.empty_like(h1) if need & 1 else None
        dh2 = torch.empty_like(h2) if need & 2 else None
        dz_acc = ctx.dz_acc
        flags = _lib.F_PREZEROED if ctx.acc_clean else 0
        ctx.acc_clean = False
        psync = None
        if ctx.peer is not None and ctx.peer[2]:  # the peers' row factors: waited for inside the kernel
            sync = ctx.peer[0].sync_for(ctx.peer[2])
            psync = ctypes.byref(sync)
        with _on_device(dev):
            if ctx.rs:
                out = _NTXentFunction._backward_reduce_scatter(ctx, g, need, dh1, dh2, dz_acc)
            else:
                with _Profiler.span("bwd"):
                    _lib.check(lib.maai_ntxent_bwd(_ptr(z_all), _ptr(r_row), _ptr(r_col), 1 if full else 0,
                                                   _ptr(rowsum), _ptr(pos_cos), _ptr(h1), _ptr(h2), dt, _ptr(inv_norm),
                                                   _ptr(g), b, world, rank, d, dp, inv_tau, need, _ptr(dh1),
                                                   _ptr(dh2), _ptr(dz_acc), flags, psync, _stream(dev)),
                               "maai_ntxent_bwd")
                out = (dh1, dh2) + (None,) * 9
        if ctx.peer is not None:
            ctx.peer[0].backward_issued(ctx.peer[1])  # the set's readers are all on the stream now (PeerWorkspace reuse rule)
        return out


def _backward_reduce_scatter(ctx, g, need, dh1, dh2, dz_acc):
    """Full gradient by the dataflow SURVEY.md section 7 names: key-side partial sums of every rank's
    anchors over the local keys -> reduce_scatter(sum) -> local rows; the query-side tile pass runs
    while the collective is in flight.  Every rank must call backward with the same need mask."""
    lib = _lib.load()
    h1, h2, z_all, inv_norm, r_row, r_col, pos_cos, rowsum = ctx.saved_tensors
    b, d, dp, dt, inv_tau, rank, world, _ = ctx.cfg
    dev = h1.device
    st = _stream(dev)
    r_pad = torch.zeros(lib.maai_ntxent_r_len(b, 1), dtype=torch.float32, device=dev)
    r_pad[:2 * b].copy_(r_row)
    dz_keys = torch.empty((world * 2 * b, dp), dtype=torch.float32, device=dev)
    dz_mine = torch.empty((2 * b, dp), dtype=torch.float32, device=dev)
    with _Profiler.span("bwd_keyside"):
        _lib.check(lib.maai_ntxent_bwd_keyside(_ptr(z_all), _ptr(r_pad), b, world, rank, dp, inv_tau,
                                               _ptr(dz_keys), st), "maai_ntxent_bwd_keyside")
    work = dist.reduce_scatter_tensor(dz_mine, dz_keys, group=ctx.rs_group, async_op=True)
    with _Profiler.span("bwd"):
        _lib.check(lib.maai_ntxent_bwd_tiles(_ptr(z_all), _ptr(r_row), _ptr(r_col), b, world, rank, dp, inv_tau,
                                             need, _ptr(dz_acc), st), "maai_ntxent_bwd_tiles")
    with _Profiler.span("reduce_scatter_wait"):
        work.wait()
    with _Profiler.span("bwd_dh"):
        _lib.check(lib.maai_ntxent_bwd_dh(_ptr(dz_acc), _ptr(dz_mine), _ptr(rowsum), _ptr(pos_cos), _ptr(h1),
                                          _ptr(h2), dt, _ptr(inv_norm), _ptr(g), b, d, dp, inv_tau, 1, need,
                                          _ptr(dh1), _ptr(dh2), st), "maai_ntxent_bwd_dh")
    return (dh1, dh2) + (None,) * 9


_NTXentFunction._backward_reduce_scatter = staticmethod(_backward_reduce_scatter)


def _forward_eval(hidden1, hidden2, temperature, rank, world, group):
    """validate() path without logits: loss + rank of every positive among the view-b keys
    (maai_ntxent_fwd_eval).  No autograd graph."""
    lib = _lib.load()
    b, d = hidden1.shape
    dev = hidden1.device
    dp = padded_dim(d)
    h1 = hidden1.detach().contiguous()
    h2 = hidden2.detach().contiguous()
    z_all = torch.empty((world, 2 * b, dp), dtype=torch.bfloat16, device=dev)
    inv_norm = torch.empty(2 * b, dtype=torch.float32, device=dev)
    pos_cos = torch.empty(b, dtype=torch.float32, device=dev)
    rowsum = torch.empty(2 * b, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    ranks = torch.empty(b, dtype=torch.int32, device=dev)
    with _on_device(dev):
        st = _stream(dev)
        _lib.check(lib.maai_ntxent_normalize(_ptr(h1), _ptr(h2), b, d, _DTYPES[h1.dtype], _ptr(z_all[rank]),
                                             _ptr(inv_norm), _ptr(pos_cos), None, 0, st), "maai_ntxent_normalize")
        gather_rows(z_all, rank, group)
        _lib.check(lib.maai_ntxent_fwd_eval(_ptr(z_all), b, world, rank, dp, 1.0 / float(temperature),
                                            _ptr(pos_cos), _ptr(rowsum), _ptr(loss), _ptr(ranks), st),
                   "maai_ntxent_fwd_eval")
    return loss, ranks


def _validate(hidden1, hidden2, hidden_norm, temperature, world_size, local_rank):
    if hidden1.shape != hidden2.shape:
        raise AssertionError(f"hidden1.shape {tuple(hidden1.shape)} != hidden2.shape "
                             f"{tuple(hidden2.shape)}")  # Objective.py:45
    if hidden1.dim() != 2:
        raise ValueError("hidden1/hidden2 must be (bsz, dim)")
    if not hidden_norm:
        raise NotImplementedError("hidden_norm=False is not supported by the fused sm_100a path "
                                  "(fixed-maximum logsumexp needs unit-norm rows); no reference "
                                  "caller uses it")
    if torch.are_deterministic_algorithms_enabled():
        # same contract as torch's own non-deterministic CUDA ops (at::globalContext().alertNotDeterministic)
        msg = ("maai NT-Xent does not have a deterministic implementation: row sums and the gradient accumulator are "
               "combined with fp32 atomics across CTAs (run-to-run differences <= 1e-6 relative, DESIGN.md section 3). "
               "Use torch.use_deterministic_algorithms(True, warn_only=True) to run it anyway")
        if not torch.is_deterministic_algorithms_warn_only_enabled():
            raise RuntimeError(msg)
        _warn_once("deterministic", msg)
    if not hidden1.is_cuda or not hidden2.is_cuda:
        raise RuntimeError("maai NT-Xent runs on CUDA (sm_100a) tensors only: there is no CPU fallback")
    if hidden1.dtype not in _DTYPES or hidden2.dtype not in _DTYPES:
        raise TypeError("hidden1/hidden2 must be float32, bfloat16 or float16")
    t = float(temperature)
    if not (t >= MIN_TEMPERATURE) or not math.isfinite(t):
        raise ValueError(f"temperature must be >= {MIN_TEMPERATURE} (got {temperature})")
    if world_size < 1 or not (0 <= local_rank < world_size):
        raise ValueError("need 0 <= local_rank < world_size")
    if hidden1.shape[0] < 1:
        raise ValueError("empty batch")


_warned = set()


def _warn_once(key: str, msg: str) -> None:
    if key not in _warned:
        _warned.add(key)
        import warnings
        warnings.warn(msg, stacklevel=3)


def contrastive_loss(hidden1, hidden2, hidden_norm=True, temperature=1.0, local_rank=0, world_size=1,
                     device="cpu", *, group=None, key_grad=None, return_logits=None, fused_topk=False,
                     peer_gather=None, _stash=None, _chain=None, _carry=None):
    """Drop-in for Objective.contrastive_loss (Objective.py:17-81).

    Args (reference): hidden1, hidden2 (bsz, dim); hidden_norm; temperature; local_rank (really the
      global rank, Contrastive_Learning.py:688); world_size; device (ignored: the tensors' device).
    Extra keyword-only args: ``group`` process group for the gathers; ``key_grad`` see module doc
      (True: full gradient via the symmetry of E, no gradient collective; False: the reference's
      query-side-only gradient; "reduce_scatter": the same full gradient computed as key-side partial
      sums + ``dist.reduce_scatter_tensor`` -- kept for comparison, 1.5x the tensor-core work).
      Default None = True; with ``world_size > 1`` that is NOT what the reference computes -- its
      ``dist.all_gather`` is non-differentiable (Objective.py:112-114), so an unmodified multi-GPU run
      trains on the query-side half only (about half the gradient norm at the same learning rate) -- so the
      first such call warns once; pass ``key_grad=True`` / ``False`` explicitly to choose;
      ``return_logits`` force (True) / suppress (False) the (logits_ab, labels) outputs, default:
      only when autograd is disabled (the validate() path).
      ``peer_gather`` (world_size > 1): True = the two gathers of the path are fused into the producing
      kernels as NVLink peer stores + a symmetric-memory barrier (PeerWorkspace); False = NCCL
      all_gather_into_tensor; None (default) = peer stores when torch symmetric memory is usable.

      ``fused_topk=True`` (evaluation only: autograd disabled or no input requires grad): the second
      return value is the int32 vector ``pos_rank`` (bsz,) -- how many view-b keys of all ranks are
      more similar to each view-a anchor than its positive -- instead of the (bsz, B) logits, the
      third is None; ``Model_Util.top_k_accuracy(pos_rank, None, k)`` of this package turns it into
      the reference's contrastive top-k accuracy (Contrastive_Learning.py:867-868) without the
      (bsz, B) fp32 logits and (bsz, 2B) int64 one-hot labels ever existing.

    Returns (loss, logits_ab, labels) like the reference.
    """
    _validate(hidden1, hidden2, hidden_norm, temperature, world_size, local_rank)
    if key_grad is None:
        key_grad = True
        if int(world_size) > 1 and torch.is_grad_enabled() and (hidden1.requires_grad or hidden2.requires_grad):
            _warn_once("key_grad", "maai NT-Xent: world_size > 1 with the default key_grad=True yields the FULL "
                       "gradient (query + key side; after DDP's 1/W it equals the single-process reference on the "
                       "global batch). The reference's own multi-GPU branch drops the key-side half "
                       "(non-differentiable dist.all_gather, Objective.py:112-114): pass key_grad=False to "
                       "reproduce it, or key_grad=True to silence this warning (see INTEGRATION.md on the "
                       "learning-rate implication)")
    if hidden1.dtype != hidden2.dtype:
        # e.g. hidden1 = outputs1.data kept from an autocast forward, hidden2 fp32: compute in the wider
        # type like the reference's F.normalize / matmul type promotion would
        wide = torch.promote_types(hidden1.dtype, hidden2.dtype)
        hidden1, hidden2 = hidden1.to(wide), hidden2.to(wide)
    if fused_topk:
        if torch.is_grad_enabled() and (hidden1.requires_grad or hidden2.requires_grad):
            raise ValueError("fused_topk=True is the validate() path: call it under torch.no_grad() "
                             "or with inputs that do not require grad")
        loss, ranks = _forward_eval(hidden1, hidden2, float(temperature), int(local_rank),
                                    int(world_size), group)
        return loss, ranks, None
    stash = _stash
    want_logits = (not torch.is_grad_enabled()) if return_logits is None else bool(return_logits)
    if want_logits and stash is None:
        stash = {}
    peer = False
    if int(world_size) > 1:
        if peer_gather is None:
            peer = peer_gather_available() and _peer_usable(hidden1.shape[0], padded_dim(hidden1.shape[1]),
                                                            int(world_size), int(local_rank), hidden1.device, group)
        else:
            peer = bool(peer_gather)
    ext = None
    loss = None
    if stash is None and _carry is None and _chain is None:
        ext = _lib.fast_ext()  # C++ autograd binding of the same C-ABI calls (host cost only)
        if ext is not None and (_Profiler.enabled or _Profiler._ext_on):
            _Profiler.sync_ext(ext)
    if ext is not None and int(world_size) == 1:
        loss = ext.ntxent_loss(hidden1, hidden2, float(temperature))
    elif ext is not None and peer and key_grad is True:
        loss = _peer_fast(ext, hidden1, hidden2, float(temperature), int(local_rank), int(world_size), group)
    if loss is None:
        loss = _NTXentFunction.apply(hidden1, hidden2, float(temperature), int(local_rank),
                                     int(world_size), group, key_grad, stash, peer, _chain, _carry)
    logits_ab = labels = None
    if want_logits:
        logits_ab, labels = _logits_and_labels(stash["z_all"], hidden1.shape[0], int(local_rank),
                                               int(world_size), float(temperature))
    return loss, logits_ab, labels


def _peer_fast(ext, hidden1, hidden2, temperature, rank, world, group):
    """The default multi-rank training call (peer gathers ordered by in-kernel flags, full gradient, no cross-rank
    symmetric forward) through the C++ binding: Python only picks the buffer set.  None = not this case."""
    b, d = hidden1.shape
    dev = hidden1.device
    dp = padded_dim(d)
    ws = PeerWorkspace.get(b, dp, world, rank, dev, group)
    if not ws.use_flags or _sym_forward_mode(b, dp, world, True) != "off":
        return None
    needs_grad = torch.is_grad_enabled() and (hidden1.requires_grad or hidden2.requires_grad)
    i, extra_barrier = ws.next_set(needs_grad)
    ws.seq += 1
    if extra_barrier:  # the previous user of this set ran its backward late: see PeerWorkspace
        with _on_device(dev):
            ws.hdl.barrier(channel=0)
    return ext.ntxent_loss_peer(hidden1, hidden2, temperature, ws.fast_args(i, ws.seq))


def _logits_and_labels(z_all, b, rank, world, temperature):
    """validate()-only outputs (Contrastive_Learning.py:860-868): logits_ab = z1 . Z2^T / tau
    (Objective.py:73) from the normalised bf16 rows, labels = one_hot(rank*b + k, 2B) (Objective.py:57)."""
    z1 = z_all[rank, :b].float()
    z2_all = z_all[:, b:].reshape(world * b, -1).float()
    logits_ab = torch.matmul(z1, z2_all.t()) / temperature
    idx = torch.arange(b, device=z_all.device) + rank * b
    labels = torch.nn.functional.one_hot(idx, 2 * world * b)
    return logits_ab, labels


class GraphedNTXentLoss(torch.nn.Module):
    """Single-rank :func:`contrastive_loss` whose forward and backward are each replayed as ONE CUDA
    graph (``torch.cuda.make_graphed_callables`` around the C-ABI launches, which are all issued on
    the current stream and allocate nothing outside torch's capturing allocator).

    At the batch sizes the reference trains with (256 - 4096 pairs per GPU, Contrastive_Learning.py:92)
    a step of this path is a few tens of microseconds of kernels behind ~150 us of Python, ctypes and
    autograd-engine time; the graph removes the host side.  Shapes, dtype, temperature and which inputs
    require grad are fixed at construction.  The returned loss is the graph's static output tensor: it
    is overwritten by the next call (read or clone it before), as with any graphed callable.

        loss_fn = GraphedNTXentLoss(bsz, dim, temperature, hidden1_requires_grad=False, device=dev)
        loss = loss_fn(outputs1.data, outputs2)        # Contrastive_Learning.py:685-690
        loss.backward()
    """

    def __init__(self, bsz, dim, temperature=1.0, dtype=torch.float32, device=None, hidden1_requires_grad=False,
                 hidden2_requires_grad=True):
        super().__init__()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if not (hidden1_requires_grad or hidden2_requires_grad):
            raise ValueError("GraphedNTXentLoss is the training call: at least one input must require grad")
        t = float(temperature)
        a = torch.randn(bsz, dim, device=dev).to(dtype).requires_grad_(bool(hidden1_requires_grad))
        b = torch.randn(bsz, dim, device=dev).to(dtype).requires_grad_(bool(hidden2_requires_grad))
        _validate(a, b, True, t, 1, 0)
        _lib.load()
        self.shape, self.dtype, self.temperature = (int(bsz), int(dim)), dtype, t
        self.requires = (bool(hidden1_requires_grad), bool(hidden2_requires_grad))

        ext = _lib.fast_ext()

        def fn(h1, h2):
            if ext is not None:
                return ext.ntxent_loss(h1, h2, t)
            return _NTXentFunction.apply(h1, h2, t, 0, 1, None, True, None, False)
        self._graphed = torch.cuda.make_graphed_callables(fn, (a, b))

    def forward(self, hidden1, hidden2):
        if tuple(hidden1.shape) != self.shape or tuple(hidden2.shape) != self.shape:
            raise ValueError(f"GraphedNTXentLoss was captured for shape {self.shape}")
        if hidden1.dtype != self.dtype or hidden2.dtype != self.dtype:
            raise TypeError(f"GraphedNTXentLoss was captured for dtype {self.dtype}")
        if torch.is_grad_enabled() and (hidden1.requires_grad, hidden2.requires_grad) != self.requires:
            raise ValueError(f"GraphedNTXentLoss was captured for requires_grad = {self.requires}")
        return self._graphed(hidden1, hidden2)


class NTXentLoss(torch.nn.Module):
    """Module form of :func:`contrastive_loss` holding temperature / rank / world / group.

    ``chain_views=True``: the reference's training loop passes every step's
    ``outputs2`` to the next step as ``hidden1`` (``outputs1 = outputs2``, Contrastive_Learning.py:700; consumed
    as ``outputs1.data``, :685).  The module then keeps the previous step's gathered bf16 rows and 1/norms and,
    when it recognises the chain -- ``hidden1`` is the very tensor memory passed as ``hidden2`` last time,
    unmodified since, same shape / dtype, not requiring grad -- skips normalising ``hidden1`` and, across
    ranks, HALF of the embedding gather: the view-a rows of all ranks are copied locally from the previous
    step's buffer (maai_ntxent_normalize_chain).  Results are identical to the unchained call (the same
    bf16 rows enter the tile kernels).  Any mismatch (first step, new tensor, in-place update, eval call)
    silently takes the normal path."""

    def __init__(self, temperature=1.0, local_rank=0, world_size=1, group=None, key_grad=None, chain_views=False,
                 peer_gather=None):
        super().__init__()
        self.temperature = float(temperature)
        self.local_rank = int(local_rank)
        self.world_size = int(world_size)
        self.group = group
        self.key_grad = key_grad
        self.chain_views = bool(chain_views)
        self.peer_gather = peer_gather
        self._carry = None
        self.chained_steps = 0  # how many calls took the chained K1 (tests / reporting)

    def _chain_for(self, hidden1):
        c = self._carry
        if (c is None or not torch.is_grad_enabled() or hidden1.requires_grad or not hidden1.is_contiguous()
                or hidden1.data_ptr() != c["h2"].data_ptr() or hidden1.shape != c["h2"].shape
                or hidden1.dtype != c["h2"].dtype or c["h2"]._version != c["version"]):
            return None
        return c

    def forward(self, hidden1, hidden2):
        chain = self._chain_for(hidden1) if self.chain_views else None
        carry = {} if (self.chain_views and torch.is_grad_enabled()) else None
        loss = contrastive_loss(hidden1, hidden2, True, self.temperature, self.local_rank,
                                self.world_size, hidden1.device, group=self.group,
                                key_grad=self.key_grad, peer_gather=self.peer_gather, return_logits=False,
                                _chain=chain, _carry=carry)[0]
        if carry:
            # keeps hidden2's storage alive, so an equal data_ptr next time means the same memory
            carry["h2"] = hidden2.detach()
            carry["version"] = hidden2._version
            self._carry = carry
            self.chained_steps += chain is not None
        return loss
