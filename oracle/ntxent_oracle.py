"""CPU oracle for the NT-Xent hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``multimodal-active-ai_b200``) never imports anything under ``oracle/``.

What it restates (reference = /root/reference/SimCLR/Objective.py):

* ``Objective.py:41-43``   L2 normalisation, eps 1e-12            -> ``l2_normalise``
* ``Objective.py:51-58``   cross-replica concat + label offset     -> ``rank`` / ``world`` arguments
* ``Objective.py:67-74``   four logit blocks aa, bb, ab, ba, / tau -> ``_view_loss_and_grad``
* ``Objective.py:68,71``   self-similarity mask (-1e9 on aa / bb)  -> ``LARGE_NUM``
* ``Objective.py:76-77``   CE over cat([cross, self]) / local b    -> ``_view_loss_and_grad``
* ``Objective.py:79``      loss_a + loss_b                         -> ``ntxent_rank``
* autograd through the above (no explicit code in the reference; trigger is
  ``Contrastive_Learning.py:698``)                                 -> analytic gradient, fp64

Everything is float64 numpy, row-blocked so that B = 32768 pairs fits in RAM.

Pinning: the reference ships no tests and no golden vectors ("parity unpinned" by the
reference's own tests).  This oracle is pinned instead against outputs
of the imported reference itself: ``tests/golden/make_golden.py`` (run in the build container,
where /root/reference exists) wrote ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks the
oracle against every one of them on CPU.
"""
from __future__ import annotations

import numpy as np

LARGE_NUM = 1e9  # Objective.py:6
NORM_EPS = 1e-12  # torch.nn.functional.normalize default, used at Objective.py:42-43


def l2_normalise(h: np.ndarray):
    """z = h / max(||h||, eps) row-wise (Objective.py:41-43).  Returns (z, norm_clamped)."""
    h = np.asarray(h, dtype=np.float64)
    n = np.sqrt((h * h).sum(axis=1, keepdims=True))
    n = np.maximum(n, NORM_EPS)
    return h / n, n


def _normalise_backward(h, z, n, dz):
    """Backward of ``l2_normalise``: dh = (dz - z (z.dz)) / n  (rows with ||h|| < eps: dz / eps)."""
    raw = np.sqrt((np.asarray(h, np.float64) ** 2).sum(axis=1, keepdims=True))
    proj = (z * dz).sum(axis=1, keepdims=True)
    dh = (dz - z * proj) / n
    clamped = raw < NORM_EPS
    if clamped.any():
        dh = np.where(clamped, dz / n, dh)
    return dh


def _view_loss_and_grad(q, k_cross, k_self, label_off, tau, b_div, block):
    """One of the two soft-label cross entropies of Objective.py:76-77.

    q        (b, d)  local anchors of this view (already normalised)
    k_cross  (B, d)  all keys of the other view      -> logits_ab / logits_ba  (Objective.py:73-74)
    k_self   (B, d)  all keys of the same view       -> logits_aa / logits_bb  (Objective.py:67,70)
    label_off        rank * b: column of anchor r's positive inside k_cross, and of its own copy
                     inside k_self (Objective.py:55-58)

    Returns loss (float), dq (b,d) query-side gradient, dk_cross (B,d), dk_self (B,d) key-side
    gradients, and lse (b,) -- all for the *un-detached* graph, i.e. what single-process
    autograd produces when keys alias queries.
    """
    b, _ = q.shape
    B = k_cross.shape[0]
    loss = 0.0
    dq = np.zeros_like(q)
    dkc = np.zeros_like(k_cross)
    dks = np.zeros_like(k_self)
    lse_all = np.empty(b)
    for r0 in range(0, b, block):
        r1 = min(b, r0 + block)
        rows = np.arange(r0, r1)
        cross = q[r0:r1] @ k_cross.T / tau
        own = q[r0:r1] @ k_self.T / tau
        own[rows - r0, rows + label_off] -= LARGE_NUM
        logits = np.concatenate([cross, own], axis=1)  # Objective.py:76: cat([ab, aa], 1)
        m = logits.max(axis=1, keepdims=True)
        ex = np.exp(logits - m)
        den = ex.sum(axis=1, keepdims=True)
        lse = (m + np.log(den))[:, 0]
        lse_all[r0:r1] = lse
        pos = logits[rows - r0, rows + label_off]  # one-hot target column (Objective.py:57)
        loss += float((lse - pos).sum())
        g = ex / den
        g[rows - r0, rows + label_off] -= 1.0
        g /= b_div  # "/ inputs.shape[0]" (Objective.py:125)
        gc, gs = g[:, :B], g[:, B:]
        dq[r0:r1] = (gc @ k_cross + gs @ k_self) / tau
        dkc += gc.T @ q[r0:r1] / tau
        dks += gs.T @ q[r0:r1] / tau
    return loss / b_div, dq, dkc, dks, lse_all


def ntxent_rank(z1_loc, z2_loc, z1_all, z2_all, rank, tau, block=2048):
    """Loss of one rank and gradients w.r.t. *normalised* embeddings.

    Returns dict with
      loss        scalar, exactly Objective.py:79 for this rank
      dq1, dq2    (b,d)   query-side gradients on the local anchors
      dk1, dk2    (B,d)   key-side gradients on every global key (what the reference's
                          non-differentiable all_gather drops when world_size > 1)
      lse1, lse2  (b,)    log-sum-exp per anchor (diagnostic)
    """
    b = z1_loc.shape[0]
    off = rank * b
    la, dqa, dk2_a, dk1_a, lse1 = _view_loss_and_grad(z1_loc, z2_all, z1_all, off, tau, b, block)
    lb, dqb, dk1_b, dk2_b, lse2 = _view_loss_and_grad(z2_loc, z1_all, z2_all, off, tau, b, block)
    return dict(loss=la + lb, dq1=dqa, dq2=dqb, dk1=dk1_a + dk1_b, dk2=dk2_a + dk2_b,
                lse1=lse1, lse2=lse2)


def contrastive_loss_oracle(hidden1, hidden2, temperature=1.0, hidden_norm=True, block=2048):
    """Single-process reference semantics (world_size == 1): loss and full gradients.

    Mirrors ``contrastive_loss(hidden1, hidden2, hidden_norm, temperature)`` of
    Objective.py:17-81 followed by ``loss.backward()`` with both inputs requiring grad.
    Returns (loss, dh1, dh2) in float64.
    """
    h1 = np.asarray(hidden1, np.float64)
    h2 = np.asarray(hidden2, np.float64)
    assert h1.shape == h2.shape  # Objective.py:45
    if hidden_norm:
        z1, n1 = l2_normalise(h1)
        z2, n2 = l2_normalise(h2)
    else:
        z1, z2 = h1, h2
    r = ntxent_rank(z1, z2, z1, z2, 0, float(temperature), block)
    dz1 = r["dq1"] + r["dk1"]
    dz2 = r["dq2"] + r["dk2"]
    if hidden_norm:
        return r["loss"], _normalise_backward(h1, z1, n1, dz1), _normalise_backward(h2, z2, n2, dz2)
    return r["loss"], dz1, dz2


def contrastive_loss_oracle_distributed(h1_ranks, h2_ranks, temperature=1.0, key_grad=True, block=2048):
    """World-size W semantics.  ``h1_ranks`` / ``h2_ranks``: lists of (b,d) arrays, one per rank.

    Per-rank loss is exactly the reference's ``world_size > 1`` branch (Objective.py:51-58, 67-79).

    key_grad=False  gradients as the reference produces them: all_gather output carries no
                    grad_fn (Objective.py:112-114), so only the query-side term survives.
    key_grad=True   gradient of sum_r loss_r w.r.t. every rank's inputs (what a differentiable
                    gather + reduce-scatter gives; after DDP's 1/W averaging this equals the
                    single-process reference on the concatenated batch).

    Returns (losses[W], dh1[W], dh2[W]).
    """
    W = len(h1_ranks)
    zs1, ns1, zs2, ns2 = [], [], [], []
    for p in range(W):
        z, n = l2_normalise(h1_ranks[p]); zs1.append(z); ns1.append(n)
        z, n = l2_normalise(h2_ranks[p]); zs2.append(z); ns2.append(n)
    Z1 = np.concatenate(zs1, 0)  # Objective.py:114 torch.cat(tensor_list, 0): rank order
    Z2 = np.concatenate(zs2, 0)
    b = zs1[0].shape[0]
    losses = []
    dz1 = [np.zeros_like(z) for z in zs1]
    dz2 = [np.zeros_like(z) for z in zs2]
    for p in range(W):
        r = ntxent_rank(zs1[p], zs2[p], Z1, Z2, p, float(temperature), block)
        losses.append(r["loss"])
        dz1[p] += r["dq1"]
        dz2[p] += r["dq2"]
        if key_grad:
            for q in range(W):
                dz1[q] += r["dk1"][q * b:(q + 1) * b]
                dz2[q] += r["dk2"][q * b:(q + 1) * b]
    dh1 = [_normalise_backward(h1_ranks[p], zs1[p], ns1[p], dz1[p]) for p in range(W)]
    dh2 = [_normalise_backward(h2_ranks[p], zs2[p], ns2[p], dz2[p]) for p in range(W)]
    return losses, dh1, dh2


def legacy_compute_loss(z1, z2, temperature):
    """Restatement of the legacy Algorithm-1 loop, SimCLR/SimCLR.py:132-144, *including* its
    operator-precedence quirk ``Sum / 2*N`` (= Sum*N/2).  Views are interleaved as in
    SimCLR.py:63-66 (row 2k = z2[k], row 2k+1 = z1[k]); similarity is cosine (SimCLR.py:53-125);
    ``_compute_l`` (SimCLR.py:36-48) has no max subtraction."""
    z1 = np.asarray(z1, np.float64); z2 = np.asarray(z2, np.float64)
    N = z1.shape[0]
    z = np.empty((2 * N, z1.shape[1]))
    z[0::2] = z2
    z[1::2] = z1
    zn = z / np.maximum(np.sqrt((z * z).sum(1, keepdims=True)), 1e-8)  # nn.CosineSimilarity eps
    s = zn @ zn.T
    total = 0.0
    for k in range(N):
        for i, j in ((2 * k + 1, 2 * k), (2 * k, 2 * k + 1)):
            e = np.exp(s[i] / temperature)
            den = e.sum() - e[i]
            total += -np.log(np.exp(s[i, j] / temperature) / den)
    return total / 2 * N


def contrastive_topk_oracle(hidden1, hidden2, k, hidden_norm=True):
    """top_k_accuracy(logits_ab, labels, k) of Model_Util.py:104-113 on the ab block
    (Contrastive_Learning.py:867-868), single process: fraction of anchors whose positive
    z2[r] ranks within the k largest of row r of z1 @ z2.T (ties broken like a strict count)."""
    h1 = np.asarray(hidden1, np.float64); h2 = np.asarray(hidden2, np.float64)
    if hidden_norm:
        h1, _ = l2_normalise(h1); h2, _ = l2_normalise(h2)
    ab = h1 @ h2.T
    pos = np.diag(ab)
    rank_of_pos = (ab > pos[:, None]).sum(axis=1)
    return float((rank_of_pos < k).mean())


def positive_rank_oracle(h1_ranks, h2_ranks, hidden_norm=True):
    """Per rank p: for every local view-a anchor r, how many of the GLOBAL view-b keys are strictly
    more similar than its positive -- the 0-based rank of column p*b + r inside row r of logits_ab
    (Objective.py:73 with the cross-replica keys of :52-53; labels_idx + rank*b of :55).
    top_k_accuracy(logits_ab, labels, k) (Model_Util.py:104-113) == mean(rank < k) barring exact ties.
    Returns a list of int64 arrays (b,), one per rank."""
    h1 = [np.asarray(h, np.float64) for h in h1_ranks]
    h2 = [np.asarray(h, np.float64) for h in h2_ranks]
    if hidden_norm:
        h1 = [l2_normalise(h)[0] for h in h1]
        h2 = [l2_normalise(h)[0] for h in h2]
    z2_all = np.concatenate(h2, axis=0)
    out = []
    off = 0
    for z1 in h1:
        ab = z1 @ z2_all.T
        b = z1.shape[0]
        pos = ab[np.arange(b), off + np.arange(b)]
        out.append((ab > pos[:, None]).sum(axis=1).astype(np.int64))
        off += b
    return out


# ------------------------------------------------------------------------------------------------
# Sampled-anchor oracle for BASELINE.json's full sizes (32768 / 65536 pairs): the full fp64 oracle
# above is O(B^2 d) in time AND needs every row; here only `pairs` sampled pairs of one rank are
# evaluated against ALL 2B keys, in fp64, row-blocked.
# ------------------------------------------------------------------------------------------------
def row_denominators(H1, H2, temperature, rows=None, block=1024):
    """den_i = sum_{j != i} exp((z_i.z_j - 1)/tau) over ALL 2B keys (the self-similarity entry is the
    one Objective.py:68,71 masks with -1e9; the positive is included), for the stacked global rows
    [view a of every pair; view b of every pair] -- i.e. exp(-1/tau) times the softmax denominator of
    Objective.py:76-77, 123-125 (logits cat([cross, self], 1)).  ``rows``: indices into the stacked
    2B rows (default: all, O(B^2 d) -- for small B and for validating faster all-row references).
    fp64 numpy.  Returns (len(rows),)."""
    z1, _ = l2_normalise(H1)
    z2, _ = l2_normalise(H2)
    Z = np.concatenate([z1, z2], 0)
    rows = np.arange(Z.shape[0]) if rows is None else np.asarray(rows, np.int64)
    out = np.empty(len(rows))
    for r0 in range(0, len(rows), block):
        idx = rows[r0:r0 + block]
        e = np.exp((Z[idx] @ Z.T - 1.0) / float(temperature))
        e[np.arange(len(idx)), idx] = 0.0
        out[r0:r0 + block] = e.sum(1)
    return out


def ntxent_rows_oracle(H1, H2, pairs, temperature, rank=0, world=1, den_all=None, key_grad=True, block=256):
    """Loss terms and input gradients of SAMPLED anchors of one rank against all 2B global keys.

    H1, H2   (B, d) the GLOBAL batch in rank order (rank p owns pairs [p*b, (p+1)*b), b = B/world:
             Objective.py:52-53 torch.cat(tensor_list, 0)).
    pairs    indices k in [0, b) of the sampled pairs of rank ``rank``; both anchors of a pair are
             evaluated (view a = hidden1[k], view b = hidden2[k]).
    den_all  (2B,) ``row_denominators`` of every stacked global row.  Only the key-side gradient
             needs it (row j's softmax normaliser for every key j); if None it is computed here in
             fp64 (O(B^2 d)).  Tests at B >= 16384 pass a faster all-row reference that they first
             validate against this module's fp64 values on the sampled rows.
    key_grad True: gradient of sum_ranks loss_rank (single-process reference on the concatenated batch
             after DDP's 1/W; Objective.py:60-61 aliasing keys and queries); False: the reference's own
             world_size > 1 gradient, keys detached by the non-differentiable all_gather
             (Objective.py:112-114).

    Returns dict:
      terms   (2, n)  lse_i - s_i,pos per sampled anchor (view a, view b): the summands of
                      Objective.py:123-125 before the division by b
      lneg    (2, n)  sum over the NEGATIVES of exp((z_i.z_j - 1)/tau) (no self, no positive)
      den     (2, n)  the same plus the positive (the sampled entries of ``row_denominators``)
      pos_cos (n,)    cosine of the pair
      dh1, dh2 (n, d) gradient rows of hidden1[pairs], hidden2[pairs] (rank-local indexing)
    """
    H1 = np.asarray(H1, np.float64)
    H2 = np.asarray(H2, np.float64)
    tau = float(temperature)
    B, d = H1.shape
    assert B % world == 0
    b = B // world
    pairs = np.asarray(pairs, np.int64)
    n = len(pairs)
    gk = rank * b + pairs                       # global pair index (labels_idx + rank*b, Objective.py:55)
    z1, n1 = l2_normalise(H1)
    z2, n2 = l2_normalise(H2)
    Z = np.concatenate([z1, z2], 0)             # stacked global rows: view a, then view b
    if den_all is None and key_grad:
        den_all = row_denominators(H1, H2, tau)
    out = dict(terms=np.empty((2, n)), lneg=np.empty((2, n)), den=np.empty((2, n)),
               pos_cos=(z1[gk] * z2[gk]).sum(1))
    dz = [np.empty((n, d)), np.empty((n, d))]
    inv_den_all = None if den_all is None else 1.0 / np.asarray(den_all, np.float64)
    for v in (0, 1):
        own = gk + v * B                        # stacked row index of the anchors
        pos = gk + (1 - v) * B                  # ... of their positives
        for r0 in range(0, n, block):
            sl = slice(r0, min(n, r0 + block))
            m = sl.stop - sl.start
            ar = np.arange(m)
            q = Z[own[sl]]
            e = np.exp((q @ Z.T - 1.0) / tau)   # (m, 2B)
            e[ar, own[sl]] = 0.0                # self-similarity mask (Objective.py:68,71)
            den = e.sum(1)
            e_pos = e[ar, pos[sl]].copy()
            out["den"][v, sl] = den
            out["lneg"][v, sl] = den - e_pos
            out["terms"][v, sl] = np.log(den) - np.log(e_pos)
            # G_ij = softmax_ij - onehot_ij (divided by the local b below, Objective.py:125)
            w = e / den[:, None]
            if key_grad:                        # + G_ji = E_ij / den_j - [i == pos(j)]
                w += e * inv_den_all[None, :]
                w[ar, pos[sl]] -= 2.0
            else:
                w[ar, pos[sl]] -= 1.0
            dz[v][sl] = (w @ Z) / (b * tau)
    out["dh1"] = _normalise_backward(H1[gk], z1[gk], n1[gk], dz[0])
    out["dh2"] = _normalise_backward(H2[gk], z2[gk], n2[gk], dz[1])
    return out


def loss_from_denominators(H1, H2, den_all, temperature, rank=0, world=1):
    """Per-rank loss of Objective.py:76-79 from the all-row denominators:
    (1/b) sum over the rank's 2b anchors of [ln den_i + 1/tau - cos_i,pos / tau]."""
    H1 = np.asarray(H1, np.float64)
    H2 = np.asarray(H2, np.float64)
    B = H1.shape[0]
    b = B // world
    z1, _ = l2_normalise(H1[rank * b:(rank + 1) * b])
    z2, _ = l2_normalise(H2[rank * b:(rank + 1) * b])
    cos = (z1 * z2).sum(1)
    den = np.asarray(den_all, np.float64)
    idx = np.arange(rank * b, (rank + 1) * b)
    t = float(temperature)
    return float((np.log(den[idx]) + np.log(den[B + idx]) + 2.0 * (1.0 - cos) / t).sum() / b)
