/* C ABI of the B200-native NT-Xent path (libmaai_ntxent.so).
 *
 * Drop-in boundary for /root/reference/SimCLR/Objective.py::contrastive_loss (Objective.py:17-81,
 * helpers :102-125) and the autograd replay it triggers (Contrastive_Learning.py:698).  The
 * Python host (multimodal-active-ai_b200/Objective.py) binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says host; nothing is allocated, freed or
 *     synchronised inside the library; every call only enqueues work on `stream`
 *     (a cudaStream_t passed as void*), so all calls are CUDA-graph capturable;
 *   - return value: 0 ok, MAAI_E_ARG bad argument, MAAI_E_SHAPE unsupported shape,
 *     MAAI_E_CUDA CUDA failure (text in maai_last_error());
 *   - b = pairs owned by this rank, world = ranks, B = world*b, rows are laid out rank-major:
 *     global row (p, v, k) = p*2b + v*b + k  (v = 0: hidden1 / view a, v = 1: hidden2 / view b),
 *     i.e. exactly the order of an all-gather of each rank's stacked (2b, d_pad) block.  The
 *     positive of a row is the other view of the same (p, k)  (labels_idx + rank*b, Objective.py:55).
 *   - d_pad = maai_padded_dim(d) in {64, 128, 256}; z rows are zero-padded to d_pad.
 */
#ifndef MAAI_NTXENT_H_
#define MAAI_NTXENT_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAAI_ABI_VERSION 8

#define MAAI_OK 0
#define MAAI_E_ARG (-1)
#define MAAI_E_SHAPE (-2)
#define MAAI_E_CUDA (-3)

/* element type of hidden1 / hidden2 / dh1 / dh2 */
#define MAAI_DT_F32 0
#define MAAI_DT_BF16 1
#define MAAI_DT_F16 2

int maai_abi_version(void);
const char* maai_last_error(void);

/* Padded embedding width used by the tile kernels: 64, 128 or 256; MAAI_E_SHAPE if d < 1 or d > 256. */
int maai_padded_dim(int d);

/* Step workspace (optional, caller-owned; SURVEY.md section 8b "no allocation inside the C ABI"): one
 * 16-byte aligned fp32 region that K1 zero-fills (its zero_fill argument), so that the step needs no
 * memset / zero kernel and the forward can finish its per-row tail inside the tile kernel:
 *   words [0, 2b)                       rowsum_l
 *   words [2b, 2b + MAAI_WS_CTL_WORDS)  control words (word 0: CTA-done counter of the forward)
 *   words [head, head + 2b*d_pad)       dz_acc, only with need_bwd; head = 2b + 32 rounded up to 128
 * Pass its base as rowsum_l to the forward and base + head as dz_acc to the backward, both with
 * MAAI_F_PREZEROED.  Returns 0 for an unsupported shape. */
#define MAAI_WS_CTL_WORDS 32
#define MAAI_F_PREZEROED 1
size_t maai_ntxent_workspace_bytes(int b, int d_pad, int need_bwd);

/* 1 if maai_ntxent_fwd will run the symmetric (upper-triangle) tile schedule for this shape: single rank
 * only, by size, MAAI_FWD_SYM=0/1 forces.  Reporting / tests; the result is the same either way. */
int maai_ntxent_fwd_is_symmetric(int b, int world, int d_pad);

/* In-kernel peer synchronisation (world > 1, optional; NULL = the caller orders the fused gathers with its own
 * barrier between the calls).  Every rank owns a FLAG BLOCK of MAAI_FLAG_WORDS 32-bit words in peer-mapped
 * memory, zero before the first step: word [kind*32 + p] = the last step whose stores of that kind rank p has
 * completed into THIS rank's buffers (kind 0: bf16 rows, 1: row factors, 2: staged partial row sums).  A
 * producing kernel fences its peer / multicast stores at system scope and its last CTA release-stores `seq`
 * into its word of every rank's block; a consuming kernel spins (bounded: it traps after ~3 s, it never
 * hangs) on its LOCAL block with acquire loads right before its TMA producer first touches rows of the
 * rank in question.  No barrier kernel, no host involvement: K1's stores and the first tiles of the forward
 * overlap, and the whole multi-rank step is a fixed sequence of this library's kernels.
 *   peer_flag_bases  DEVICE array of `world` device pointers: rank p's flag block as mapped into this process
 *   local_flags      this rank's own flag block
 *   counter          one zero-initialised 32-bit device word of scratch (CTA-done counter, self-resetting)
 *   seq              step number: > 0, the same on every rank, larger at every step
 *   timeout_s        wall-clock limit of a wait for a peer, seconds (0 = 300)
 * Buffer reuse stays the caller's business exactly as with barriers: a rank that has observed seq = t from a
 * peer knows that peer has enqueued, and its GPU executed, everything before its K1 of step t. */
#define MAAI_FLAG_WORDS 96
typedef struct maai_peer_sync {
  const void* const* peer_flag_bases;
  unsigned int* local_flags;
  unsigned int* counter;
  unsigned int seq;
  unsigned int timeout_s;
} maai_peer_sync;

/* Number of floats the `r_glob` array of maai_ntxent_bwd must hold: world*2b rounded up to 128. */
size_t maai_ntxent_r_len(int b, int world);

/* K1 -- replaces F.normalize x2 (Objective.py:41-43) and the fp32 send buffers of
 * _cross_replica_concat (Objective.py:112).
 *   h1, h2     (b, d) row-major contiguous, dtype in_dtype
 *   z_out      (2b, d_pad) bf16: rows [0,b) = normalised h1, rows [b,2b) = normalised h2; pass the
 *              address of this rank's slot of the (world, 2b, d_pad) gather buffer
 *   inv_norm   (2b) fp32   1 / max(||h||, 1e-12)
 *   pos_cos    (b)  fp32   cosine of each positive pair computed from the bf16 rows
 *   zero_fill  optional (NULL / 0): 16-byte aligned region of zero_bytes bytes (multiple of 4) that the
 *              kernel zero-fills on the side: the step workspace above */
int maai_ntxent_normalize(const void* h1, const void* h2, int b, int d, int in_dtype, void* z_out,
                          float* inv_norm, float* pos_cos, void* zero_fill, size_t zero_bytes, void* stream);

/* K2 -- replaces the four matmuls, /temperature, the LARGE_NUM self-mask, cat + log_softmax and the
 * masked sum (Objective.py:67-79, 123-125) for this rank's 2b anchor rows against all world*2b keys.
 *   z_glob     (world*2b, d_pad) bf16, gathered rows (for world == 1 the K1 output itself)
 *   pos_cos    (b) from K1
 *   rowsum_l   (2b) out: l'_i = sum over the NEGATIVES j (j != i, j != pos(i)) of exp((z_i.z_j - 1)/tau);
 *              the positive's term e_pos = exp((pos_cos - 1)/tau) is added in fp32 where needed
 *   r_out      (2b) out, may be NULL: 1 / (b * (e_pos + l'_i)), the row factor the backward needs
 *   loss_out   (1)  out: this rank's loss, exactly Objective.py:79
 *   flags      0: rowsum_l is zeroed inside (one more launch) and the per-row tail (loss, r) is a separate
 *              finalize launch.  MAAI_F_PREZEROED: rowsum_l is the base of a step workspace that K1 has
 *              zero-filled; no zero launch, and up to 32768 rows the LAST CTA of the tile kernel to finish
 *              runs the per-row tail itself (one launch for the whole forward). */
int maai_ntxent_fwd(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                    const float* pos_cos, float* rowsum_l, float* r_out, float* loss_out, int flags,
                    const maai_peer_sync* sync /* world > 1: wait for the peers' rows in the kernel, or NULL */,
                    void* stream);

/* K1 fused with the cross-replica embedding gather (replaces Objective.py:41-43 AND :52-53, 102-114):
 * the normalised bf16 rows go straight into slot `rank` of every rank's (world, 2b, d_pad) key buffer
 * through peer-mapped pointers (NVLink / NVSwitch stores issued by the kernel itself, no collective
 * launch).  The caller orders the stores before any rank reads with a symmetric-memory barrier.
 *   peer_z_bases  DEVICE array of `world` device pointers: entry p = base address, as mapped into this
 *                 process, of rank p's key buffer (entry `rank` is the local buffer)
 *   mc_z_base     NVLink multicast (multimem) address of the same buffers, or NULL: when given, each
 *                 row is stored ONCE and the NVSwitch replicates it into every rank's buffer */
int maai_ntxent_normalize_peer(const void* h1, const void* h2, int b, int d, int in_dtype,
                               const void* const* peer_z_bases, void* mc_z_base, int world, int rank,
                               float* inv_norm, float* pos_cos, void* zero_fill, size_t zero_bytes,
                               const maai_peer_sync* sync /* signal kind 0 when the rows have landed, or NULL */,
                               void* stream);

/* K1 for CHAINED VIEWS: the reference's training loop passes this step's
 * outputs2 to the next step as hidden1 ("outputs1 = outputs2", Contrastive_Learning.py:700; consumed
 * detached, :685), so the view-a rows of every rank at step t are the view-b rows of step t-1 that every
 * rank already holds, normalised and in bf16, in its gathered key buffer of step t-1.  Reads only h2:
 * normalises it into the view-b rows (local store, peer stores or multicast exactly as above: HALF the
 * gather payload), copies the view-b rows of all `world` slots of z_prev into the view-a rows of z_new
 * (local copy) and carries the view-a 1/norm over from inv_norm_prev.  hidden1 must not require grad.
 *   z_prev         (world, 2b, d_pad) bf16: the previous step's gathered buffer (this rank's copy)
 *   inv_norm_prev  (2b) the previous step's inv_norm
 *   z_new          (world, 2b, d_pad) bf16: this rank's buffer of this step (!= z_prev)
 *   peer_z_bases / mc_z_base   as in maai_ntxent_normalize_peer, or NULL / NULL (view b stored into z_new) */
int maai_ntxent_normalize_chain(const void* h2, int b, int d, int in_dtype, const void* z_prev,
                                const float* inv_norm_prev, void* z_new, const void* const* peer_z_bases,
                                void* mc_z_base, int world, int rank, float* inv_norm, float* pos_cos,
                                void* zero_fill, size_t zero_bytes, const maai_peer_sync* sync, void* stream);

/* K2 fused with the all-gather of the row factors: like maai_ntxent_fwd, but r_i is stored into slot
 * `rank` of every rank's gathered r array (maai_ntxent_r_len floats each, zero padded by the owner).
 *   peer_r_bases  DEVICE array of `world` peer-mapped base addresses of the r arrays
 *   mc_r_base     multicast address of the r arrays, or NULL */
int maai_ntxent_fwd_peer(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                         const float* pos_cos, float* rowsum_l, const void* const* peer_r_bases,
                         void* mc_r_base, float* loss_out, int flags,
                         const maai_peer_sync* sync /* wait kind 0 per key slot, signal kind 1 after the r stores */,
                         void* stream);

/* K2 across ranks with the symmetry of E (world > 1): E_ij = E_ji, so every (anchor slot, key slot)
 * pair of ranks needs its tiles computed ONCE.  Rank p computes its own block (the tiles on / above the
 * diagonal) and, for the ranks q = p+1 .. p+world/2 (mod world; the pair at distance world/2 is split
 * between the two), the tiles (anchors of q) x (keys of p): their column sums are row sums of p's own
 * anchors (-> rowsum_l), their row sums belong to q's anchors and go to slot q of `stage`.  Half the
 * MMAs / exp2s of maai_ntxent_fwd.  After a barrier across the ranks, maai_ntxent_fwd_sym_finalize adds
 * slot `rank` of every other rank's staging vectors (read through peer-mapped addresses) to rowsum_l
 * (in place), then produces loss and r like maai_ntxent_fwd / maai_ntxent_fwd_peer.
 *   stage        (world, 2b) fp32, this rank's staging vectors (zeroed inside)
 *   stage_bases  device array of `world` addresses: every rank's `stage`, as mapped into this process
 *   r_out        (2b) or NULL;  peer_r_bases / mc_r_base as in maai_ntxent_fwd_peer, or NULL */
int maai_ntxent_fwd_sym_tiles(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                              float* rowsum_l, float* stage, int flags /* MAAI_F_PREZEROED: rowsum_l */,
                              const maai_peer_sync* sync /* wait kind 0 per anchor slot, signal kind 2 at the end */,
                              void* stream);
int maai_ntxent_fwd_sym_finalize(float* rowsum_l, const void* const* stage_bases, int b, int world, int rank,
                                 float inv_tau, const float* pos_cos, float* r_out,
                                 const void* const* peer_r_bases, void* mc_r_base, float* loss_out,
                                 const maai_peer_sync* sync /* wait kind 2 of every peer, signal kind 1 */,
                                 void* stream);

/* The cross-rank symmetric forward in ONE launch (needs the in-kernel synchronisation, sync != NULL): instead of
 * staging the partial row sums it computed for the other ranks' anchors, the tile kernel adds them straight into
 * the OWNERS' row sums over NVLink (red.add through peer-mapped addresses), its last CTA signals kind 2, waits
 * for every peer's kind-2 signal (then this rank's row sums are complete) and runs the per-row tail itself:
 * no staging vectors, no zero launch for them, no finalize launch, no pull.
 *   rowsum_l          this rank's row sums: base of a step workspace IN PEER-MAPPED memory, zero-filled by K1
 *                     (the peers add into it once they have observed this rank's kind-0 signal, i.e. after K1)
 *   peer_rowsum_host  HOST array of `world` device pointers: rank p's rowsum_l as mapped into this process
 *   r_out / peer_r_bases / mc_r_base / loss_out   as in maai_ntxent_fwd_peer */
int maai_ntxent_fwd_sym_direct(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                               const float* pos_cos, float* rowsum_l, const void* const* peer_rowsum_host,
                               float* r_out, const void* const* peer_r_bases, void* mc_r_base, float* loss_out,
                               const maai_peer_sync* sync, void* stream);

/* K2 for validate() (Contrastive_Learning.py:860-868): the same forward plus, for every view-a anchor
 * k of this rank, pos_rank[k] = number of view-b keys of ALL ranks whose similarity to the anchor is
 * strictly greater than its positive's, i.e. the 0-based rank of the positive inside the row of
 * logits_ab (Objective.py:73).  top-k accuracy (Model_Util.py:104-113) = mean(pos_rank < k): neither
 * the (b, B) logits nor the (b, 2B) one-hot labels are materialised.  Exact ties count for the
 * positive (torch.topk breaks ties arbitrarily).
 *   pos_rank   (b) int32 out (zeroed inside) */
int maai_ntxent_fwd_eval(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                         const float* pos_cos, float* rowsum_l, float* loss_out, int* pos_rank,
                         void* stream);

/* K3 + K4 -- replaces loss.backward() through Objective.py:41-79 (Contrastive_Learning.py:698).
 *   dZ_i = (1/tau) [ sum_{j != pos(i)} E_ij (r_row_i + r_col_j) z_j + cpos_i z_pos(i) ]   (i: this rank's anchors)
 * cpos_i, the positive's softmax-minus-target coefficient, is formed in fp32 from rowsum_l and
 * pos_cos, outside the bf16 MMA (it is a small residual when the softmax is peaked).
 * key_grad = 1: full gradient (query side + key side) of sum_ranks loss_rank w.r.t. this rank's
 * inputs, computed rank-locally from the symmetry of E: r_row = this rank's r_out, r_col =
 * all-gathered r_out of every rank.  key_grad = 0: the reference's world_size > 1 semantics (keys
 * detached by the non-differentiable all_gather, Objective.py:112-114): r_col = zeros.
 *   r_row      (2b) floats
 *   r_col      maai_ntxent_r_len(b, world) floats, zero padded
 *   rowsum_l   (2b) from K2;  pos_cos (b) from K1
 *   grad_loss  (1) fp32 device scalar: upstream gradient of the loss
 *   need_mask  bit 0: dh1 wanted, bit 1: dh2 wanted (hidden1 is detached in the reference's
 *              training loop, Contrastive_Learning.py:685)
 *   dh1, dh2   (b, d) dtype in_dtype, written only when the matching bit is set (may be NULL otherwise)
 *   dz_acc     (2b, d_pad) fp32 scratch
 *   flags      0: dz_acc is zeroed inside (one more launch); MAAI_F_PREZEROED: the caller (K1's zero_fill)
 *              has zeroed it */
int maai_ntxent_bwd(const void* z_glob, const float* r_row, const float* r_col, int key_grad,
                    const float* rowsum_l, const float* pos_cos, const void* h1, const void* h2, int in_dtype,
                    const float* inv_norm, const float* grad_loss, int b, int world, int rank, int d, int d_pad,
                    float inv_tau, int need_mask, void* dh1, void* dh2, float* dz_acc, int flags,
                    const maai_peer_sync* sync /* wait kind 1 per key slot before its r_col is read, or NULL */,
                    void* stream);

/* The backward in its three pieces, for the key-side REDUCE-SCATTER dataflow (SURVEY.md section 7 /
 * BASELINE.json north_star): instead of using the symmetry of E (maai_ntxent_bwd, key_grad = 1), every
 * rank computes, for the anchors of ALL ranks, the key-side sums over ITS OWN keys, and a
 * reduce_scatter(sum) of the (world*2b, d_pad) fp32 result delivers each rank its rows:
 *   maai_ntxent_bwd_keyside : dz_keys[i] = sum_{j in this rank's slot, j != i, pos(i)} E_ij r_j z_j   (all i)
 *   <reduce_scatter of dz_keys over the ranks, e.g. torch.distributed.reduce_scatter_tensor>
 *   maai_ntxent_bwd_tiles   : dz_acc[i]  = sum_{j != i, pos(i)} E_ij (r_row_i + r_col_j) z_j  (this rank's i;
 *                             the first half of maai_ntxent_bwd; pass r_col = zeros for the query side only)
 *   maai_ntxent_bwd_dh      : dh from dz_acc + dz_extra (the reduce-scattered rows, or NULL) -- the second
 *                             half of maai_ntxent_bwd; key_grad = 1 keeps the key-side term of the positive.
 * Both dataflows give the same gradient (tests/test_gpu_multirank.py); the identity form needs no
 * gradient collective and 2/3 of the tensor-core work, and is the default.
 *   r_col_loc  maai_ntxent_r_len(b, 1) floats: this rank's r_out, zero padded
 *   dz_keys    (world*2b, d_pad) fp32 out, zeroed inside
 *   dz_extra   (2b, d_pad) fp32 or NULL */
int maai_ntxent_bwd_tiles(const void* z_glob, const float* r_row, const float* r_col, int b, int world, int rank,
                          int d_pad, float inv_tau, int need_mask, float* dz_acc, void* stream);
int maai_ntxent_bwd_keyside(const void* z_glob, const float* r_col_loc, int b, int world, int rank, int d_pad,
                            float inv_tau, float* dz_keys, void* stream);
int maai_ntxent_bwd_dh(const float* dz_acc, const float* dz_extra, const float* rowsum_l, const float* pos_cos,
                       const void* h1, const void* h2, int in_dtype, const float* inv_norm,
                       const float* grad_loss, int b, int d, int d_pad, float inv_tau, int key_grad, int need_mask,
                       void* dh1, void* dh2, void* stream);

/* Host-only views of the forward's tile schedule, for tests without a GPU (no CUDA call inside):
 *   maai_debug_group_plan : the anchor groups rank `rank` computes in maai_ntxent_fwd_sym_tiles (arrays of
 *                           9 entries): first anchor row (global), rows, key tiles, items of every group;
 *                           group 0 is the rank's own triangular block
 *   maai_debug_tri_locate : item `idx` of the folded triangular list of n_row_blocks row blocks over
 *                           n_key_tiles key tiles (nq Q tiles per row block) -> row block, offset inside its
 *                           segment, length of the segment */
int maai_debug_group_plan(int b, int world, int rank, int d_pad, int* ngroups, int* qrow0, int* rows, int* nkt,
                          long long* items);
int maai_debug_tri_locate(long long idx, int n_key_tiles, int n_row_blocks, int nq, int* rb, int* off, int* cnt);

/* Kernel launches enqueued by this library since load (bench.py's gpu_launches claim). */
unsigned long long maai_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MAAI_NTXENT_H_ */
