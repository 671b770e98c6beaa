// Fused NT-Xent stripe kernel for sm_100a (tcgen05 + TMEM + TMA), forward and backward.
//
// Replaces, without ever writing a logit to HBM, the four matmuls, the self-mask, the
// concat + log_softmax + masked sum of /root/reference/SimCLR/Objective.py:67-79 (forward) and
// the autograd replay of those ops triggered at Contrastive_Learning.py:698 (backward).
//
// Formulation.  Z = all normalised rows (bf16), M_glob = 2B of them,
// this rank owns M_loc = 2b "anchor" rows.  With c1 = log2(e)/tau and the fixed maximum 1/tau
// (rows are unit-norm, so every logit is <= 1/tau):
//     E_ij = exp2(c1 * z_i.z_j - c1)         (j != i, j != pos(i), j < M_glob)
//   FWD:  l'_i = sum_j E_ij   (negatives only)                  -> atomically added to l[]
//   BWD:  A_i  = sum_j E_ij (rr_i + rc_j) z_j                   -> atomically added to dz_acc[]
//         (full gradient: rr = rc = r = 1/(b*l); query-side only: rc = 0; key-side only: rr = 0).
//   The positive column is left out of both on purpose.  When the softmax is peaked the positive
//   dominates the row sum and its gradient coefficient E_i,pos (rr_i + rc_pos) - 2/b cancels to a
//   small residual; folding it in fp32 afterwards (finalize_loss_kernel: log1p(l'/e_pos),
//   dh_kernel: -(1/b) l'/(e_pos + l')) keeps loss and gradient exact at small temperatures.
// (the positive-pair term and the normalisation Jacobian are O(M d) and live in ntxent_aux.cuh).
//
// One persistent CTA per SM walks a contiguous range of (row block, key tile) items
// ("stream-K"); a maximal run of items inside one row block is a segment.  Every item yields NQ
// "S tiles" (128 anchors x 128 keys); the CTA-wide running S-tile index sigma fixes, identically in
// every role, which softmax team (sigma & 1) and which TMEM buffer a tile uses.  Roles:
//   warp 0 lane 0 : TMA producer   (Q tiles once per segment, K tiles + r_j through a ring)
//   warp 1        : tcgen05.mma issuer, warp-uniform control flow, one elected lane per instruction
//                   (S = Q K^T into TMEM; BWD also A += P Z_J with P read from TMEM and Z_J read
//                   from the *same* smem tile as an MN-major operand)
//   warp 2        : TMEM allocator
//   warps 4-19    : 16 softmax warps, one TMEM lane (= anchor row) per thread:
//                   tcgen05.ld S -> exp2 -> row sum (FWD) / P = E (r_i + r_j) -> bf16 -> TMEM over
//                   the S columns already consumed (BWD).
// NQ == 2 (two 128-row Q tiles per row block, each K tile feeds both): halves the L2 -> smem
//   traffic per flop; used by the MUFU-bound forward.  The softmax warps form two teams of 8;
//   team t owns Q tile t and its two warpgroups split the 128 key columns (64 per thread).
// NQ == 1 (one Q tile): leaves TMEM room for three S buffers next to the accumulator, so the
//   tensor pipe always has another tile's MMAs to run while a softmax is in flight; used by the
//   backward (profiles/r1_ncu_summary_v1.md shows the exposed S -> softmax -> P -> PV chain of the
//   one-buffer-per-Q-tile layout).  All 16 softmax warps work on every S tile (32 columns per
//   thread), which halves the latency of the softmax stage of each tile.
#pragma once
#include <type_traits>

#include "ntxent_aux.cuh"
#include "ptx_sm100.cuh"

// Of every 8 column pairs a softmax thread handles, this many take the polynomial exp2 on the FMA
// pipe instead of MUFU.EX2 (build-time knobs so the split can be A/B-measured).
#ifndef MAAI_POLY_FWD
#define MAAI_POLY_FWD 3
#endif
#ifndef MAAI_POLY_BWD
#define MAAI_POLY_BWD 0
#endif
// Optional second split for the odd softmax warpgroups (two of the four warps of every SM
// sub-partition): lets MUFU-heavy and FMA-heavy warps share a scheduler.  Default: same split.
#ifndef MAAI_POLY_FWD_B
#define MAAI_POLY_FWD_B MAAI_POLY_FWD
#endif
#ifndef MAAI_POLY_BWD_B
#define MAAI_POLY_BWD_B MAAI_POLY_BWD
#endif
#ifndef MAAI_POLY_DEG_FWD
#define MAAI_POLY_DEG_FWD 3
#endif
#ifndef MAAI_POLY_DEG_BWD
#define MAAI_POLY_DEG_BWD 3
#endif
// Timing-only ablations (results are WRONG when non-zero; tools/ablate.py): bit 0 no exp,
// bit 1 no P.Z MMAs, bit 2 no S MMAs, bit 3 no r_j loads, bit 4 no K-tile TMA traffic.
#ifndef MAAI_ABL
#define MAAI_ABL 0
#endif
// Columns per tcgen05.ld / register chunk of a softmax thread that owns 64 columns of an S tile:
// 32 = one register buffer, load -> wait -> process (more independent work per chunk; measured
// 3-5 % faster in both kernels), 16 = two buffers with the next load in flight and a cross-tile
// prefetch behind a non-blocking barrier probe.
#ifndef MAAI_FWD_CW
#define MAAI_FWD_CW 32
#endif
#ifndef MAAI_BWD_CW
#define MAAI_BWD_CW 32
#endif
// NQ == 1: 2 = the softmax warps form two teams of 8 that take alternate S tiles (64 columns per
// thread), 1 = all 16 warps work on every S tile (32 columns per thread).
// Register redistribution between the control warpgroup and the softmax warpgroups (96 = off).
#ifndef MAAI_REGS_SM
#define MAAI_REGS_SM 96
#endif
#ifndef MAAI_REGS_WG0
#define MAAI_REGS_WG0 64
#endif
// Symmetric forward, tiles above the diagonal: of the 4 column blocks (8 columns) of each 32-column
// chunk, this many take the polynomial exp2 (chunk 0 / chunk 1).  Measured best: 1 + 1 (these tiles
// already carry the column-sum work on the FMA / ALU pipes; 0+0 is 2.5 % slower, 2+1 3 % slower).
#ifndef MAAI_POLY_SYM0
#define MAAI_POLY_SYM0 1
#endif
#ifndef MAAI_POLY_SYM1
#define MAAI_POLY_SYM1 1
#endif
#ifndef MAAI_NQ1_TEAMS
#define MAAI_NQ1_TEAMS 2
#endif
// Backward softmax: exponentials and the P = E (r_i + r_j) -> bf16 post-processing in source-level software
// pipeline over groups of this many column pairs (0 = all exponentials of a chunk first, then all the
// post-processing: ptxas then issues the chunk's 32 MUFU.EX2 as one burst, and 39.7 % of the backward's warp
// samples sit in mio_throttle on them, profiles/r2_ncu_summary_v1.md).
#ifndef MAAI_BWD_ILV
#define MAAI_BWD_ILV 0
#endif

// Per-warp phase timers (clock64) for tools/phase_prof.py; compiled out unless MAAI_PROF=1.
#ifndef MAAI_PROF
#define MAAI_PROF 0
#endif

namespace maai {

#if MAAI_PROF
__device__ long long g_prof[160 * 20 * 8];
#define PROF_INIT()                                  \
  long long prof_t = clock64();                      \
  long long prof_a[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF_MARK(i)                                 \
  do {                                               \
    const long long _n = clock64();                  \
    prof_a[i] += _n - prof_t;                        \
    prof_t = _n;                                     \
  } while (0)
#define PROF_FLUSH()                                                          \
  do {                                                                        \
    if ((threadIdx.x & 31) == 0)                                              \
      for (int _i = 0; _i < 8; ++_i)                                          \
        g_prof[((size_t)blockIdx.x * 20 + (threadIdx.x >> 5)) * 8 + _i] = prof_a[_i]; \
  } while (0)
#else
#define PROF_INIT() do { } while (0)
#define PROF_MARK(i) do { } while (0)
#define PROF_FLUSH() do { } while (0)
#endif

constexpr int kMaxGroups = 9;  // own block + up to 8 remote anchor groups (world <= 16)

struct TileParams {
  int m_loc;             // anchor rows covered by tmap_q
  int m_glob;            // key rows covered by tmap_k
  int row_global_base;   // global (key-space) index of anchor row 0
  int pos_split;         // anchor rows < pos_split have their positive at +pos_delta, others at -pos_delta
  int pos_delta;         //      (= b; pos_split = b - first anchor row of this launch)
  int pos_period;        // != 0: anchors span several rank slots (key-side pass of the reduce-scatter
  int pos_phase;         //      dataflow): view a iff (row + pos_phase) % pos_period < pos_delta
  int nrb;               // row blocks
  int nkt;               // key tiles (128 keys each)
  float c1;              // log2(e) / tau
  const float* r_row;    // BWD: row factor per anchor row (m_loc floats); null = zeros
  const float* r_col;    // BWD: column factor per global key, padded to a multiple of 128 floats
  float* l_out;          // FWD: row sums, m_loc floats, pre-zeroed
  float* dz_acc;         // BWD: m_loc x D fp32, pre-zeroed
  uint32_t pv_lbo;       // BWD: leading / stride byte offsets of the MN-major Z_J operand
  uint32_t pv_sbo;
  // RANK (evaluation forward): for every view-a anchor, how many view-b keys are strictly more
  // similar than its positive (contrastive top-k without logits, Model_Util.py:104-113)
  const float* pos_cos;  // positive cosine per local pair (pos_split floats)
  int* rank_out;         // pos_split ints, pre-zeroed
  // GRP (symmetric forward across ranks, keys = this rank's slot): group 0 = this rank's own anchors
  // (triangular, as SYM); groups 1.. = anchors of another rank's slot (rows of tmap_q from g_qrow0)
  // against this rank's key tiles [0, g_nkt): every tile adds its row sums to the group's g_out
  // (a staging vector the owner of those anchors collects later) and its column sums to l_out.
  long long total_items;
  int ngroups;
  int g_qrow0[kMaxGroups];
  int g_rows[kMaxGroups];
  int g_nkt[kMaxGroups];
  long long g_items[kMaxGroups];  // items of the group (row blocks x key tiles; group 0 triangular)
  float* g_out[kMaxGroups];
  // FWD, optional: the last CTA to finish runs the per-row tail of the forward (loss, row factors and
  // their peer stores: finalize_rows of ntxent_aux.cuh) instead of a separate finalize launch.
  // done_ctr: zero-initialised 32-bit counter (K1's zero fill); null = no in-kernel finalize.
  unsigned int* done_ctr;
  FinalizeArgs fin;
  // Multi-rank, optional (null = the caller ordered the gathers with a barrier launch): right after its
  // prologue the CTA waits until word [wait_kind][s] of this rank's flag block has reached wait_seq for every
  // rank slot s != wait_my_slot -- FWD: the peers' bf16 rows (FLAG_Z), BWD: the peers' row factors (FLAG_R).
  // The wait sits BEFORE the mbarrier pipeline starts on purpose: a peer may legitimately be seconds late
  // (host stall on that rank), and the pipeline's own watchdog (mbar_wait, ~2 s) must only ever see waits
  // that this CTA's warps resolve among themselves.
  const unsigned int* wait_flags;
  unsigned int wait_seq;
  int wait_kind;
  int wait_my_slot;
  int wait_nslots;     // rank slots (world)
  unsigned int wait_timeout_s;
  // GRP: the last CTA signals FLAG_L (this rank's staged partial sums are complete) through grp_sync
  PeerSync grp_sync;
};

template <int D, bool BWD, int NQ>
struct TileCfg {
  static_assert(D == 64 || D == 128 || D == 256, "padded embedding dim must be 64, 128 or 256");
  static_assert(NQ == 1 || NQ == 2, "one or two Q tiles per row block");
  static_assert(NQ == 1 || D <= 128, "two Q tiles need D <= 128 (smem, TMEM)");
  static constexpr int RB_ROWS = 128 * NQ;
  static constexpr int KT = 128;                       // keys per tile
  static constexpr int CHUNKS = D / 64;                // 128-byte swizzle chunks per row
  static constexpr int CHUNK_BYTES = 128 * 128;        // 128 rows x 128 B
  static constexpr int TILE_BYTES = CHUNKS * CHUNK_BYTES;
  // K ring depth.  A stage lives from its S MMA until its P.Z MMA has completed, i.e. about NB + 1
  // tile periods in the backward, so the ring must be deeper than that for the TMA prefetch to run
  // ahead (with 4 stages and NB = 3 the load latency was exposed on every tile).
  static constexpr int NST = (D == 256) ? 2 : (D == 64 ? 8 : ((NQ == 1 || !BWD) ? 5 : 4));
  // S buffers (128 TMEM columns each).  FWD: all of TMEM.  BWD: what the accumulators leave.
  static constexpr int NB = !BWD ? 4 : (NQ == 2 ? 2 : (D <= 128 ? 3 : 2));
  static constexpr int TMEM_DZ0 = NB * 128;            // BWD accumulators: DZ0 + q*D
  static_assert(!BWD || NB * 128 + NQ * D <= 512, "TMEM budget");
  // 20 warps: warpgroup 0 = TMA producer, MMA issuer, TMEM allocator, one idle warp; warpgroups
  // 1-4 = 16 softmax warps.  Launched at 96 registers per thread (65536 / 640); warpgroup 0 then
  // gives registers back (setmaxnreg.dec) and the softmax warpgroups take them (setmaxnreg.inc).
  static constexpr int NTHREADS = 640;
  static constexpr int SM_WARP0 = 4;                   // first softmax warp
  static constexpr int REGS_WG0 = MAAI_REGS_WG0;       // 4 warps x 32 x (96 - REGS_WG0) registers freed
  static constexpr int REGS_SM = MAAI_REGS_SM;         // 16 warps x 32 x (REGS_SM - 96) taken
  static_assert(4 * (96 - REGS_WG0) >= 16 * (REGS_SM - 96), "register pool");
  // shared memory carve-up (offsets from a 1024-B aligned base)
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NQ * TILE_BYTES;
  static constexpr int OFF_RK = OFF_K + NST * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_RK + (BWD ? NST * KT * 4 : 0);  // r_j ring: backward only
  static constexpr int NBAR = 2 + 2 * NST + 2 * NB + 2;
  static constexpr int OFF_TMEMPTR = OFF_BAR + NBAR * 8;
  static constexpr int SMEM_BYTES = OFF_TMEMPTR + 16 + 1024;  // + alignment slack
  static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB per CTA)");
};

// TMEM buffer and mbarrier phase of the S tile with CTA-wide running index `sg`.
template <int NQ, int NB>
__device__ __forceinline__ void s_tile_slot(uint32_t sg, int& buf, uint32_t& phase) {
  if (NQ == 2) {  // team = sg & 1 owns buffers [team*NB/2, (team+1)*NB/2)
    const uint32_t team = sg & 1, u = sg >> 1;
    buf = int(team * (NB / 2) + u % (NB / 2));
    phase = (u / (NB / 2)) & 1;
  } else {
    buf = int(sg % NB);
    phase = (sg / NB) & 1;
  }
}

__device__ __forceinline__ void tmem_ld_cw(uint32_t a, uint32_t (&r)[16]) { tmem_ld_x16(a, r); }
__device__ __forceinline__ void tmem_ld_cw(uint32_t a, uint32_t (&r)[32]) { tmem_ld_x32(a, r); }
__device__ __forceinline__ void tmem_st_pk(uint32_t a, const uint32_t (&r)[8]) { tmem_st_x8(a, r); }
__device__ __forceinline__ void tmem_st_pk(uint32_t a, const uint32_t (&r)[16]) { tmem_st_x16(a, r); }

// One CW-column chunk of one TMEM lane (= anchor row) of an S tile, already in registers: E -> row
// sum (FWD) or P = E (r_i + r_j) -> packed bf16 in pk (BWD; the caller stores it to TMEM).  Of the 8 column pairs, POLY take the
// polynomial exp2 on the FMA pipe, the rest MUFU.EX2.  The polynomial path yields E * 2^c1 (see
// exp2_dot_poly2); FWD keeps those in their own accumulators, BWD folds 2^-c1 into (r_i + r_j).
struct ChunkCtx {
  float c1;        // log2(e) / tau
  float kscale;    // 2^-c1
  float r_i;       // BWD: row factor of this anchor
  float r_ik;      // BWD: r_i * 2^-c1
  int grow, gpos;  // key indices of this anchor's diagonal / positive entry
  int m_glob;
  float pos_s;     // RANK: similarity of the positive (+inf for rows that are not counted)
  int cnt;         // RANK: view-b keys more similar than the positive so far
  int pairs;       // RANK: b (keys k with (k mod 2b) >= b belong to view b)
#if MAAI_PROF
  long long prof_t;
  long long prof_a[8];
#endif
};
#if MAAI_PROF
#define CX_MARK(i)                                   \
  do {                                               \
    const long long _n = clock64();                  \
    cx.prof_a[i] += _n - cx.prof_t;                  \
    cx.prof_t = _n;                                  \
  } while (0)
#else
#define CX_MARK(i) do { } while (0)
#endif
template <bool BWD, int POLY, int DEG, int CW, bool RANK = false>
__device__ __forceinline__ void softmax_chunk(uint32_t (&v)[CW], uint32_t (&pk)[CW / 2], const float* rk,
                                              ChunkCtx& cx, bool special, int kc0,
                                              float2 (&acc_m)[2], float2 (&acc_p)[2], int cmode = 0) {
  if (RANK && cmode != 0) {
    // cmode 1: every key of this tile is a view-b key and none needs a predicate; 2: test each key
    int c = 0;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      bool hit = __uint_as_float(v[i]) > cx.pos_s;
      if (cmode == 2) {
        const int kc = kc0 + i;
        hit = hit && kc != cx.gpos && kc < cx.m_glob &&
              (unsigned(kc) % unsigned(2 * cx.pairs)) >= unsigned(cx.pairs);
      }
      c += hit ? 1 : 0;
    }
    cx.cnt += c;
  }
  // packed f32x2 math throughout (FFMA2 / FADD2 / FMUL2): half the FMA-pipe issue slots
  const float2 c1p = make_float2(cx.c1, cx.c1), c1n = make_float2(-cx.c1, -cx.c1);
  constexpr int NP = CW / 2;  // column pairs
  if (BWD && MAAI_BWD_ILV > 0 && POLY == 0 && !(MAAI_ABL & 1)) {
    // software pipeline: exponentials of group g + 1 are issued before the post-processing of group g
    constexpr int G = MAAI_BWD_ILV > 0 ? MAAI_BWD_ILV : 2, NG = NP / G;
    static_assert(NP % G == 0 && G % 2 == 0, "group size");
    const float2 ri2 = make_float2(cx.r_i, cx.r_i);
    float2 ec[G], en[G];
    auto expo = [&](int g, float2 (&o)[G]) {
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int jj = g * G + j;
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[2 * jj]), __uint_as_float(v[2 * jj + 1])), c1p, c1n);
        o[j].x = ex2_approx_v(x.x);
        o[j].y = ex2_approx_v(x.y);
        if (special) {
          const int kc = kc0 + 2 * jj;
          if (kc == cx.grow || kc == cx.gpos || kc >= cx.m_glob) o[j].x = 0.f;
          if (kc + 1 == cx.grow || kc + 1 == cx.gpos || kc + 1 >= cx.m_glob) o[j].y = 0.f;
        }
      }
    };
    expo(0, ec);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (g + 1 < NG) expo(g + 1, en);
#pragma unroll
      for (int j = 0; j < G; j += 2) {
        const int jj = g * G + j;
        const float4 rj = *reinterpret_cast<const float4*>(rk + 2 * jj);
        const float2 p0 = __fmul2_rn(ec[j], __fadd2_rn(ri2, make_float2(rj.x, rj.y)));
        const float2 p1 = __fmul2_rn(ec[j + 1], __fadd2_rn(ri2, make_float2(rj.z, rj.w)));
        pk[jj] = pack_bf16x2_v(p0.x, p0.y);
        pk[jj + 1] = pack_bf16x2_v(p1.x, p1.y);
      }
#pragma unroll
      for (int j = 0; j < G; ++j) ec[j] = en[j];
    }
    CX_MARK(2);
    return;
  }
  float2 e[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const float2 sv = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
    if (MAAI_ABL & 1) {
      e[j] = __ffma2_rn(sv, c1p, c1n);
    } else if ((j & 7) < POLY) {
      e[j] = exp2_dot_poly2<DEG>(sv, c1p);
    } else {
      const float2 x = __ffma2_rn(sv, c1p, c1n);
      e[j].x = ex2_approx(x.x);
      e[j].y = ex2_approx(x.y);
    }
  }
  if (special) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int kc = kc0 + 2 * j;
      if (kc == cx.grow || kc == cx.gpos || kc >= cx.m_glob) e[j].x = 0.f;
      if (kc + 1 == cx.grow || kc + 1 == cx.gpos || kc + 1 >= cx.m_glob) e[j].y = 0.f;
    }
  }
  if (!BWD) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if ((j & 7) < POLY) acc_p[j & 1] = __fadd2_rn(acc_p[j & 1], e[j]);
      else acc_m[j & 1] = __fadd2_rn(acc_m[j & 1], e[j]);
    }
  } else {
    const float2 ri2 = make_float2(cx.r_i, cx.r_i), rik2 = make_float2(cx.r_ik, cx.r_ik);
    const float2 ks2 = make_float2(cx.kscale, cx.kscale);
#pragma unroll
    for (int j = 0; j < NP; j += 2) {
      float4 rj = (MAAI_ABL & 8) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                 : *reinterpret_cast<const float4*>(rk + 2 * j);
      const float2 rj0 = make_float2(rj.x, rj.y), rj1 = make_float2(rj.z, rj.w);
      const float2 w0 = ((j & 7) < POLY) ? __ffma2_rn(rj0, ks2, rik2) : __fadd2_rn(ri2, rj0);
      const float2 w1 = (((j + 1) & 7) < POLY) ? __ffma2_rn(rj1, ks2, rik2) : __fadd2_rn(ri2, rj1);
      const float2 p0 = __fmul2_rn(e[j], w0);
      const float2 p1 = __fmul2_rn(e[j + 1], w1);
      pk[j] = pack_bf16x2(p0.x, p0.y);
      pk[j + 1] = pack_bf16x2(p1.x, p1.y);
    }
  }
  CX_MARK(2);
}

// Triangular item list (row block r holds T - NQ r key tiles), FOLDED: the row blocks are visited in
// the order 0, nrb-1, 1, nrb-2, ... so that every pair of consecutive segments holds the same
// S = 2T - NQ (nrb - 1) items.  A CTA's contiguous range of N items then spans about 2N/S segments
// wherever it lies; in row-block order the last CTA collected all the short segments (21 of them at
// 32768 pairs: each segment costs a Q-tile load + a pipeline refill, ~3-5 us).
template <int NQ>
__host__ __device__ __forceinline__ void tri_locate(long long idx, int T, int nrb, int& rb, int& off, int& cnt) {
  const int S = 2 * T - NQ * (nrb - 1);
  const int pr = int(idx / S);
  const int r = int(idx - (long long)pr * S);
  const int a = T - NQ * pr;
  if (r < a) {
    rb = pr;
    off = r;
    cnt = a;
  } else {
    rb = nrb - 1 - pr;
    off = r - a;
    cnt = T - NQ * rb;
  }
}

// Walks the CTA's item range segment by segment (a segment = a maximal run of key tiles inside one
// row block).  SYM (symmetric forward, anchors == keys): row block rb only visits key tiles
// kt >= 2*rb, because E_ij = E_ji lets every tile above the diagonal contribute its row sums to the
// anchors AND its column sums to the keys; the tiles below the diagonal are never computed.
template <bool SYM, int NQ>
struct SegWalk {
  int nkt, nrb;
  static constexpr int g = 0;
  __device__ __forceinline__ explicit SegWalk(const TileParams& p) : nkt(p.nkt), nrb(p.nrb) {}
  // segment that starts at item `it`: row block, first key tile, number of key tiles
  __device__ __forceinline__ void locate(long long it, long long it_end, int& rb_out, int& j0, int& n) {
    if (SYM) {
      int off, cnt;
      tri_locate<NQ>(it, nkt, nrb, rb_out, off, cnt);
      j0 = NQ * rb_out + off;
      n = int(min((long long)(cnt - off), it_end - it));
    } else {
      rb_out = int(it / nkt);
      j0 = int(it % nkt);
      n = int(min((long long)(nkt - j0), it_end - it));
    }
  }
};

// GRP: the item list is group after group; group 0 triangular (like SYM), the others rectangular.
template <int NQ>
struct GrpWalk {
  const TileParams& p;
  int g = 0;
  long long gbase = 0;  // items before group g
  __device__ __forceinline__ explicit GrpWalk(const TileParams& p_) : p(p_) {}
  __device__ __forceinline__ void locate(long long it, long long it_end, int& rb_out, int& j0, int& n) {
    while (it >= gbase + p.g_items[g]) gbase += p.g_items[g++];
    const long long idx = it - gbase;
    int cnt, off;
    if (g == 0) {
      tri_locate<NQ>(idx, p.g_nkt[0], p.nrb, rb_out, off, cnt);
      j0 = NQ * rb_out + off;
    } else {
      const int nk = p.g_nkt[g];
      rb_out = int(idx / nk);
      off = int(idx - (long long)rb_out * nk);
      cnt = nk;
      j0 = off;
    }
    n = int(min((long long)(cnt - off), it_end - it));
  }
};

// Per-row tail of the forward inside the tile kernel, run by ONE CTA once every CTA's sums are complete
// (finalize_rows over all 2b rows, fixed-order fp64 sum, loss, row factors + their peer stores + the FLAG_R
// signal).  Out of line on purpose: it runs once per launch and must not cost the hot loops a register.
// part_smem: shared-window address of a dead region (the Q tiles).  wait_l: cross-rank symmetric forward in
// "direct" mode -- the peers add their partial row sums straight into this rank's row sums (NVLink
// red.add); they are complete once every peer has signalled FLAG_L.
__device__ __noinline__ void tail_finalize_body(const TileParams& p, uint32_t part_smem, bool wait_l) {
  if (wait_l) {
    if (threadIdx.x < 32)
      wait_all_peers(p.grp_sync.local_flags, FLAG_L, p.grp_sync.seq, p.grp_sync.world, p.grp_sync.rank, p.grp_sync.timeout_s);
    __syncthreads();
  }
  __threadfence();
  double* part = static_cast<double*>(__cvta_shared_to_generic(part_smem));
  const double acc = finalize_rows(p.fin, threadIdx.x, blockDim.x);
  if (p.fin.sync.peer_flags) __threadfence_system();  // this thread's row-factor stores to the peers
  const double v = finalize_block_sum(acc, part);     // (contains a CTA barrier)
  if (threadIdx.x == 0) *p.fin.loss_out = float(v / double(p.fin.b));
  if (p.fin.sync.peer_flags) {
    __syncthreads();
    __threadfence_system();
    signal_peers(p.fin.sync, FLAG_R);
  }
}
// Forward, single launch: the last CTA to finish (ticket on done_ctr; every other CTA's threads fenced their
// sums before their CTA took its ticket) runs the tail.  flag_smem: a dead shared word (the mbarriers).
__device__ __noinline__ void fwd_tail_finalize(const TileParams& p, uint32_t flag_smem, uint32_t part_smem) {
  uint32_t* flag = static_cast<uint32_t*>(__cvta_shared_to_generic(flag_smem));
  if (threadIdx.x == 0) *flag = (atomicAdd(p.done_ctr, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (*flag) tail_finalize_body(p, part_smem, false);
}

template <int D, bool BWD, int NQ, bool RANK = false, bool SYM = false, bool GRP = false>
__global__ void __launch_bounds__(640, 1)
ntxent_tile_kernel(const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ TileParams p) {
  using C = TileCfg<D, BWD, NQ>;
  constexpr int NST = C::NST, NB = C::NB;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t sQ = smem_base + C::OFF_Q;
  const uint32_t sK = smem_base + C::OFF_K;
  const uint32_t sRK = smem_base + C::OFF_RK;
  const float* rk_gen = reinterpret_cast<const float*>(smem_gen + C::OFF_RK);
  const uint32_t bar0 = smem_base + C::OFF_BAR;
  // barrier map
  const uint32_t bar_q_full = bar0 + 0 * 8;
  const uint32_t bar_q_empty = bar0 + 1 * 8;
  auto bar_k_full = [&](int st) { return bar0 + (2 + st) * 8; };
  auto bar_k_empty = [&](int st) { return bar0 + (2 + NST + st) * 8; };
  auto bar_s_full = [&](int buf) { return bar0 + (2 + 2 * NST + buf) * 8; };
  auto bar_sm_done = [&](int buf) { return bar0 + (2 + 2 * NST + NB + buf) * 8; };
  const uint32_t bar_dz_full = bar0 + (2 + 2 * NST + 2 * NB) * 8;
  const uint32_t bar_dz_free = bar0 + (2 + 2 * NST + 2 * NB + 1) * 8;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + C::OFF_TMEMPTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents<4>();

  // ---- this CTA's contiguous range of (row block, key tile) items ----
  static_assert(!SYM || (!BWD && !RANK && MAAI_FWD_CW == 32 && (NQ == 2 || MAAI_NQ1_TEAMS == 2)),
                "symmetric mode: plain forward, 64 columns per softmax thread in 32-column chunks");
  static_assert(!GRP || SYM, "anchor groups are a variant of the symmetric forward");
  using Walk = typename std::conditional<GRP, GrpWalk<NQ>, SegWalk<SYM, NQ>>::type;
  // sum over rb of (nkt - NQ rb)
  const long long total = GRP ? p.total_items
                          : SYM ? (long long)p.nrb * p.nkt - (long long)NQ * p.nrb * (p.nrb - 1) / 2
                                : (long long)p.nrb * p.nkt;
  const long long it_begin = total * blockIdx.x / gridDim.x;
  const long long it_end = total * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(bar_q_full, 1);
    mbar_init(bar_q_empty, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_k_full(s), 1);
      mbar_init(bar_k_empty(s), 1);
    }
    for (int u = 0; u < NB; ++u) {
      mbar_init(bar_s_full(u), 1);
      mbar_init(bar_sm_done(u), (NQ == 2 || MAAI_NQ1_TEAMS == 2) ? 8 : 16);  // one arrive per warp on the tile
    }
    mbar_init(bar_dz_full, 1);
    mbar_init(bar_dz_free, 16);      // one arrive per softmax warp
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_base + C::OFF_TMEMPTR, 512);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch) touches no global data and may overlap the tail of the previous kernel; from here on
  // every role reads what earlier kernels wrote (z rows, r factors, zeroed accumulators) or adds to
  // it, so EVERY thread waits for the previous grid -- also the threads that never touch global
  // memory: a grid whose threads skip the wait can complete before its predecessor has, and the
  // kernel after it would then see that predecessor's writes unordered (round 1's PDL failure).
  pdl_wait();
  if (p.wait_flags) {  // in-kernel replacement of the barrier launch behind the fused gather (see TileParams)
    if (warp == 0) wait_all_peers(p.wait_flags, p.wait_kind, p.wait_seq, p.wait_nslots, p.wait_my_slot, p.wait_timeout_s);
    __syncthreads();
    if (warp == 0 && lane == 0) fence_proxy_async_all();  // the TMA (async proxy) reads what the peers' stores wrote
  }
  if (C::REGS_SM > 96) {
    if (warp < C::SM_WARP0) reg_dealloc<C::REGS_WG0>();
    else reg_alloc<C::REGS_SM>();
  }

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      PROF_INIT();
      uint32_t t = 0, useg = 0;
      Walk walk(p);
      for (long long it = it_begin; it < it_end;) {
        int rb, j0, n;
        walk.locate(it, it_end, rb, j0, n);
        mbar_wait(bar_q_empty, (useg & 1) ^ 1);
        PROF_MARK(0);
        mbar_arrive_expect_tx(bar_q_full, NQ * C::TILE_BYTES);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
          for (int c = 0; c < C::CHUNKS; ++c)
            tma_load_2d(sQ + q * C::TILE_BYTES + c * C::CHUNK_BYTES, &tmap_q, c * 64,
                        (GRP ? p.g_qrow0[walk.g] : 0) + rb * C::RB_ROWS + q * 128, bar_q_full);
        for (int jj = 0; jj < n; ++jj, ++t) {
          const int st = t % NST;
          mbar_wait(bar_k_empty(st), ((t / NST) & 1) ^ 1);
          PROF_MARK(1);
          if ((MAAI_ABL & 16) && t >= NST) {  // timing ablation: no K / r_j traffic after the first ring fill
            mbar_arrive(bar_k_full(st));
            continue;
          }
          mbar_arrive_expect_tx(bar_k_full(st), C::TILE_BYTES + (BWD ? C::KT * 4 : 0));
#pragma unroll
          for (int c = 0; c < C::CHUNKS; ++c)
            tma_load_2d(sK + st * C::TILE_BYTES + c * C::CHUNK_BYTES, &tmap_k, c * 64,
                        (j0 + jj) * C::KT, bar_k_full(st));
          if (BWD)
            bulk_load_1d(sRK + st * C::KT * 4, p.r_col + (size_t)(j0 + jj) * C::KT, C::KT * 4,
                         bar_k_full(st));
          PROF_MARK(2);
        }
        it += n;
        ++useg;
      }
      PROF_FLUSH();
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // The whole warp walks the schedule (warp-uniform control flow and descriptor arithmetic);
    // one elected lane executes each tcgen05.mma / tcgen05.commit.
    constexpr uint32_t IDESC_S = make_idesc_bf16(128, 128, 0);
    constexpr uint32_t IDESC_PV = make_idesc_bf16(128, D, 1);
    // descriptor of tile base address 0; real addresses are added to the low word (>> 4)
    const uint64_t sdesc_k = make_sdesc_sw128(0, 16, 1024);             // K-major: SBO = 8 rows
    const uint64_t sdesc_mn = make_sdesc_sw128(0, p.pv_lbo, p.pv_sbo);  // MN-major Z_J for P.Z
    uint32_t t = 0;     // key tiles consumed so far (K ring position); S-tile index = t*NQ + q
    uint32_t useg = 0;
    PROF_INIT();

    auto commit = [&](uint32_t bar) {
      if (elect_one()) umma_commit(bar);
      __syncwarp();
    };
    // S tile with running index sg: Q tile q x K stage st -> its TMEM buffer
    auto issue_s = [&](uint32_t sg, int q, int st) {
      int buf;
      uint32_t ph;
      s_tile_slot<NQ, NB>(sg, buf, ph);
      if (!BWD) {  // FWD: wait until the softmax team has drained this buffer's previous tile
        mbar_wait(bar_sm_done(buf), ph ^ 1);
        tc_fence_after();
        PROF_MARK(1);
      }
      const uint32_t d_tmem = tmem_base + buf * 128;
      const uint64_t ad = sdesc_k + ((sQ + q * C::TILE_BYTES) >> 4);
      const uint64_t bd = sdesc_k + ((sK + st * C::TILE_BYTES) >> 4);
      if (elect_one()) {
        if (!(MAAI_ABL & 4)) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t off = ((k >> 2) * C::CHUNK_BYTES + (k & 3) * 32) >> 4;
            umma_ss(d_tmem, ad + off, bd + off, IDESC_S, k > 0);
          }
        }
        umma_commit(bar_s_full(buf));
      }
      __syncwarp();
      PROF_MARK(2);
    };
    // BWD: A_q += P(sg) * Z_J(stage st); P was written over S by the softmax team
    auto issue_pv = [&](uint32_t sg, int q, int st, bool first) {
      int buf;
      uint32_t ph;
      s_tile_slot<NQ, NB>(sg, buf, ph);
      mbar_wait(bar_sm_done(buf), ph);
      tc_fence_after();
      PROF_MARK(1);
      const uint32_t d_tmem = tmem_base + C::TMEM_DZ0 + q * D;
      const uint32_t a_tmem = tmem_base + buf * 128;
      const uint64_t bd = sdesc_mn + ((sK + st * C::TILE_BYTES) >> 4);
      if (!(MAAI_ABL & 2) && elect_one()) {
#pragma unroll
        for (int k = 0; k < C::KT / 16; ++k) {
          // 16 keys = 16 smem rows of 128 B (2048 B).  P (2 keys per 32-bit column) of each
          // thread's key range sits at the start of that range's S columns:
          // 64 keys per thread (two teams) -> key half h at columns [h*64, h*64+32)
          // 32 keys per thread (one team)  -> key quarter c at columns [c*32, c*32+16)
          const uint32_t pcol = (NQ == 2 || MAAI_NQ1_TEAMS == 2) ? (k >> 2) * 64 + (k & 3) * 8
                                                                 : (k >> 1) * 32 + (k & 1) * 8;
          umma_ts(d_tmem, a_tmem + pcol, bd + k * (2048 >> 4), IDESC_PV,
                  (first && k == 0) ? 0u : 1u);
        }
      }
      __syncwarp();
      PROF_MARK(3);
    };

    Walk walk(p);
    uint32_t symu[2] = {0, 0};  // SYM: S tiles issued so far per team (a team skips tiles below the diagonal)
#pragma unroll 1
    for (long long it = it_begin; it < it_end;) {
      int rb, j0, n;
      walk.locate(it, it_end, rb, j0, n);
      mbar_wait(bar_q_full, useg & 1);
      tc_fence_after();
      PROF_MARK(0);
      const int ns = n * NQ;            // S tiles of this segment, local index s = jj*NQ + q
      const uint32_t sg0 = t * NQ;      // running index of the first one
      // S tile s of the segment: waits for its K stage if it is the first to touch it
      auto issue_s_local = [&](int s) {
        const int jj = s / NQ, q = s % NQ;
        const uint32_t tk = t + jj;
        const int st = tk % NST;
        if (q == 0) {
          mbar_wait(bar_k_full(st), (tk / NST) & 1);
          tc_fence_after();
          PROF_MARK(0);
        }
        issue_s(sg0 + s, q, st);
      };
      if (!BWD && SYM && NQ == 2) {
#pragma unroll 1
        for (int jj = 0; jj < n; ++jj) {
          const uint32_t tk = t + jj;
          const int st = tk % NST;
          mbar_wait(bar_k_full(st), (tk / NST) & 1);
          tc_fence_after();
          PROF_MARK(0);
          // Q tile 1 of the row block sits one key tile further down the diagonal: its tile with the
          // first key tile of the block (kt == 2 rb) lies below the diagonal and is not computed
          const int nq_here = (!(GRP && walk.g > 0) && j0 + jj < 2 * rb + 1) ? 1 : 2;
          for (int q = 0; q < nq_here; ++q) {
            issue_s(symu[q] * 2 + q, q, st);
            ++symu[q];
          }
          commit(bar_k_empty(st));
        }
        commit(bar_q_empty);
      } else if (!BWD) {
#pragma unroll 1
        for (int s = 0; s < ns; ++s) {
          issue_s_local(s);
          if (s % NQ == NQ - 1) commit(bar_k_empty((t + s / NQ) % NST));
        }
        commit(bar_q_empty);
      } else {
        // NB S tiles are kept in flight ahead of the P.Z products (P aliases its S buffer, and
        // tcgen05 ops execute in issue order, so S tile s+NB may be issued right after P.Z of s)
        const int la = ns < NB ? ns : NB;
        for (int s = 0; s < la; ++s) issue_s_local(s);
        if (ns <= NB) commit(bar_q_empty);
        // accumulators of the previous segment must have been flushed
        mbar_wait(bar_dz_free, (useg & 1) ^ 1);
        tc_fence_after();
        PROF_MARK(4);
#pragma unroll 1
        for (int s = 0; s < ns; ++s) {
          const int jj = s / NQ, q = s % NQ;
          const int st = (t + jj) % NST;
          issue_pv(sg0 + s, q, st, jj == 0);
          if (q == NQ - 1) commit(bar_k_empty(st));  // last reader of the stage
          if (s + NB < ns) {
            issue_s_local(s + NB);
            if (s + NB == ns - 1) commit(bar_q_empty);  // that was the last read of the Q tiles
          }
        }
        commit(bar_dz_full);
      }
      t += n;
      it += n;
      ++useg;
      PROF_MARK(5);
    }
    PROF_FLUSH();
  } else if (warp >= C::SM_WARP0) {
    // =========================== softmax warps ===========================
    const int sw = warp - C::SM_WARP0;
    const int wgi = sw >> 2;                          // softmax warpgroup 0..3
    constexpr int TEAMS = (NQ == 2) ? 2 : MAAI_NQ1_TEAMS;
    // NQ == 2: team t owns Q tile t.  NQ == 1, two teams: team t takes the S tiles of parity t.
    const int team = (TEAMS == 2) ? (wgi >> 1) : 0;
    constexpr int STEP = (NQ == 1 && TEAMS == 2) ? 2 : 1;  // ring steps between tiles of this warp
    constexpr int TCOLS = (TEAMS == 2) ? 64 : 32;     // key columns per thread per S tile
    // columns per tcgen05.ld / register chunk
    constexpr int CW = (TCOLS == 64) ? (BWD ? MAAI_BWD_CW : MAAI_FWD_CW) : 16;
    constexpr int NCH = TCOLS / CW;                   // register chunks per S tile (even)
    constexpr int MODB = (NQ == 2) ? NB / 2 : NB;     // S buffers this warp cycles through
    const int col_off = (TEAMS == 2) ? (wgi & 1) * 64 : wgi * 32;  // first key column of this thread
    const int w4 = warp & 3;          // TMEM lane quarter this warp may touch
    const int row_in_tile = w4 * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(w4 * 32) << 16) + col_off;
    const float c1 = p.c1;
    const float kscale = ex2_approx(-c1);  // 2^-c1: what the polynomial path leaves out
    // backward at d_pad = 64: the MMAs are half as long, MUFU is the top pipe -> 2 of 8 pairs on the
    // polynomial (1.345 -> 1.301 ms at 32768 pairs); at 128 / 256 the split makes no difference
    constexpr int POLY_BWD_D = (D == 64 && MAAI_POLY_BWD == 0) ? 2 : MAAI_POLY_BWD;
    constexpr int POLY_A = BWD ? POLY_BWD_D : MAAI_POLY_FWD;
    constexpr int POLY_B = BWD ? ((MAAI_POLY_BWD_B == MAAI_POLY_BWD) ? POLY_BWD_D : MAAI_POLY_BWD_B) : MAAI_POLY_FWD_B;
    constexpr int DEG = BWD ? MAAI_POLY_DEG_BWD : MAAI_POLY_DEG_FWD;
    // Running ring positions of the next S tile of this warp (kept incrementally: no div / mod
    // in the per-tile path): S buffer slot + phase, K stage + phase.
    int sb = 0, kst = 0;
    uint32_t sph = 0, kph = 0;
    uint32_t tpar = 0;  // parity of the CTA-wide S-tile counter (NQ == 1, two teams)
    const int kt_ragged = (p.m_glob & (C::KT - 1)) ? p.nkt - 1 : -8;
    uint32_t useg = 0;
    // v0 holds (or is receiving) the first chunk of the next tile when `have` is set
    MBAR_PROBE_DECL();
    uint32_t v0[CW], v1[CW];
    bool have = false;
    ChunkCtx cx;
    cx.c1 = c1;
    cx.kscale = kscale;
    cx.m_glob = p.m_glob;
#if MAAI_PROF
    cx.prof_t = clock64();
    for (int i = 0; i < 8; ++i) cx.prof_a[i] = 0;
#endif

    Walk walk(p);
    // SYM: partial row sums of the tiles above the diagonal, in the 16x256b fragment layout: slot k of
    // a thread is tile row 32 w4 + 8 k + lane / 4 (k = 0..3), summed over the thread's columns
    float2 racc_m[4], racc_p[4];
#pragma unroll 1
    for (long long it = it_begin; it < it_end;) {
      int rb, j0, n;
      walk.locate(it, it_end, rb, j0, n);
      const int q = (NQ == 2) ? team : 0;
      const int grp = GRP ? walk.g : 0;
      const bool remote = GRP && grp > 0;                       // anchors of another rank's slot
      const int rows_here = GRP ? p.g_rows[grp] : p.m_loc;
      const int row = rb * C::RB_ROWS + q * 128 + row_in_tile;  // anchor row (local / inside the group)
      const bool valid = row < rows_here;
      const int grow = p.row_global_base + row;                 // same row in key space
      const int g0 = p.row_global_base + rb * C::RB_ROWS + q * 128;
      // key index of this row's positive (masked like the diagonal)
      const bool view_a =
          p.pos_period ? (row + p.pos_phase) % p.pos_period < p.pos_delta : row < p.pos_split;
      const int gpos = view_a ? grow + p.pos_delta : grow - p.pos_delta;
      // Key tiles that may hold a diagonal entry or a positive of this Q tile (a Q tile may
      // straddle the view boundary, so both placements) or keys past the end take the
      // per-element predicates: tiles kt_x and kt_x + 1 of each of the three 128-key windows.
      const int kt_d = g0 >> 7, kt_p1 = (g0 - p.pos_delta) >> 7, kt_p2 = (g0 + p.pos_delta) >> 7;
      // remote anchors meet neither themselves nor their positives among the local keys; only rows
      // past the end of the group (they belong to the next slot) have to be kept out of the column sums
      const bool ragged_rows = GRP && rb * C::RB_ROWS + q * 128 + 128 > rows_here;
      cx.r_i = (BWD && valid && p.r_row) ? __ldg(p.r_row + row) : 0.f;
      cx.r_ik = cx.r_i * kscale;
      cx.grow = grow;
      cx.gpos = gpos;
      if (RANK) {
        const bool counted = !BWD && valid && row < p.pos_split;  // view-a anchors only
        cx.pos_s = counted ? __ldg(p.pos_cos + row) : __int_as_float(0x7f800000);
        cx.cnt = 0;
        cx.pairs = p.pos_delta;
      }
      float2 acc_m[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      float2 acc_p[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      bool any_upper = false;
      if (SYM) {
#pragma unroll
        for (int k = 0; k < 4; ++k) racc_m[k] = racc_p[k] = make_float2(0.f, 0.f);
      }

#pragma unroll 1
      for (int jj = 0; jj < n; ++jj) {
        const int kt = j0 + jj;
        // SYM: key tile kt against row tile 2 rb + q: below the diagonal -> not computed at all,
        // on it -> ordinary tile (row sums), above it -> row sums AND column sums
        if (SYM && !remote && kt < NQ * rb + q) continue;
        const bool upper = SYM && (remote || kt > NQ * rb + q);
        const int buf = (NQ == 2) ? team * MODB + sb : sb;
        const bool special = remote ? (ragged_rows || kt == kt_ragged)
                                    : (unsigned(kt - kt_d) <= 1u || unsigned(kt - kt_p1) <= 1u ||
                                       unsigned(kt - kt_p2) <= 1u || kt == kt_ragged || ragged_rows);
        int cmode = 0;
        if (RANK) {  // which keys of this tile are view-b keys: none / all / mixed
          const int seg_lo = (kt * C::KT) / p.pos_delta, seg_hi = (kt * C::KT + C::KT - 1) / p.pos_delta;
          cmode = (seg_lo != seg_hi || special) ? 2 : ((seg_lo & 1) ? 1 : 0);
        }
        const uint32_t s_addr = lane_base + buf * 128;
        const float* rk = rk_gen + kst * C::KT + col_off;
        const int kbase = kt * C::KT + col_off;
        if (STEP == 2 && int(tpar) != team) {  // the other team's tile: just step the rings
          if (++sb == MODB) { sb = 0; sph ^= 1; }
          if (++kst == NST) { kst = 0; kph ^= 1; }
          tpar ^= 1;
          continue;
        }
        CX_MARK(5);
        if (!have) {
          if (BWD) mbar_wait(bar_k_full(kst), kph);  // r_j of this stage has landed
          mbar_wait(bar_s_full(buf), sph);
          tc_fence_after();
          if (!upper) tmem_ld_cw(s_addr, v0);
        }
        CX_MARK(0);
        if (SYM && upper) {
          // ---------------- tile above the diagonal: E_ij serves row i and column j ----------------
          any_upper = true;
          const uint32_t t_base = tmem_base + buf * 128 + col_off;
          const int r_lo = w4 * 32 + (lane >> 2);          // tile row of slot 0; slot k adds 8 k
          const int c_lo = 2 * (lane & 3);                 // first of the thread's two columns per 8-column block
          constexpr int PG[2] = {MAAI_POLY_SYM0, MAAI_POLY_SYM1};  // polynomial column blocks (of 4) per chunk
          const float2 c1p = make_float2(c1, c1), c1n = make_float2(-c1, -c1);
          float2 colp[4];  // column partial sums of the current chunk over this thread's 4 rows
          uint32_t v[16];
          auto load_half = [&](int c, int h) {
            tmem_ld_16x256b_x4(t_base + (uint32_t(w4 * 32 + h * 16) << 16) + c * 32, v);
          };
          // one 16-lane half of the warp's 32 rows at a time (16 registers in flight; loading both
          // halves first was measured 15 % slower: more live registers, spills)
          auto half = [&](int c, int h, float2 (&cp)[4]) {
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(v[i]));
            float2 e[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
              for (int hi = 0; hi < 2; ++hi) {
                const float2 sv = make_float2(__uint_as_float(v[4 * g + 2 * hi]), __uint_as_float(v[4 * g + 2 * hi + 1]));
                if (g < PG[c]) {
                  e[2 * g + hi] = exp2_dot_poly2<DEG>(sv, c1p);
                } else {
                  const float2 x = __ffma2_rn(sv, c1p, c1n);
                  e[2 * g + hi].x = ex2_approx(x.x);
                  e[2 * g + hi].y = ex2_approx(x.y);
                }
              }
            }
            // the registers are free again: start the next load before the sums (and, for the last
            // half, hand the S buffer back: every column of the tile is in registers)
            if (!(c == 1 && h == 1)) {
              load_half(h == 1 ? c + 1 : c, h ^ 1);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_sm_done(buf));
            }
            if (special) {  // one branch per half: the positive of a view-a row, keys past the end
#pragma unroll
              for (int g = 0; g < 4; ++g) {
#pragma unroll
                for (int hi = 0; hi < 2; ++hi) {
                  const int lr = rb * C::RB_ROWS + q * 128 + r_lo + 8 * (2 * h + hi);
                  const int gr = remote ? -1 : p.row_global_base + lr;
                  const int gp = remote ? -1 : (lr < p.pos_split ? gr + p.pos_delta : gr - p.pos_delta);
                  const int kc = kt * C::KT + col_off + c * 32 + 8 * g + c_lo;
                  const bool rbad = GRP && lr >= rows_here;
                  if (rbad || kc == gp || kc == gr || kc >= p.m_glob) e[2 * g + hi].x = 0.f;
                  if (rbad || kc + 1 == gp || kc + 1 == gr || kc + 1 >= p.m_glob) e[2 * g + hi].y = 0.f;
                }
              }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
              for (int hi = 0; hi < 2; ++hi) {
                const float2 ev = e[2 * g + hi];
                const int k = 2 * h + hi;  // row slot
                if (g < PG[c]) racc_p[k] = __fadd2_rn(racc_p[k], ev);
                else racc_m[k] = __fadd2_rn(racc_m[k], ev);
                cp[g] = (h == 0 && hi == 0) ? ev : __fadd2_rn(cp[g], ev);
              }
            }
          };
          // column sums over the warp's 32 rows: the 8 lanes that share lane % 4 hold the same 8
          // columns; three exchange stages leave one column per lane
          auto col_flush = [&](int c, const float2 (&cp)[4]) {
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
            const float a[8] = {cp[0].x, cp[0].y, cp[1].x, cp[1].y, cp[2].x, cp[2].y, cp[3].x, cp[3].y};
            float s4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float send = b4 ? a[i] : a[i + 4], keep = b4 ? a[i + 4] : a[i];
              s4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
            float s2[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float send = b3 ? s4[i] : s4[i + 2], keep = b3 ? s4[i + 2] : s4[i];
              s2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            const float send = b2 ? s2[0] : s2[1], keep = b2 ? s2[1] : s2[0];
            float cs = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            // this lane now owns value index 4 b4 + 2 b3 + b2 = 2 g + parity of the 8
            const int gsel = (b4 ? 2 : 0) + (b3 ? 1 : 0);
            if (gsel < PG[c]) cs *= kscale;
            const int key = kt * C::KT + col_off + c * 32 + 8 * gsel + c_lo + (b2 ? 1 : 0);
            if (key < p.m_loc) atomicAdd(p.l_out + key, cs);
          };
          load_half(0, 0);
          half(0, 0, colp);
          half(0, 1, colp);
          col_flush(0, colp);  // the first load of chunk 1 is in flight
          half(1, 0, colp);
          half(1, 1, colp);
          col_flush(1, colp);
          if (++sb == MODB) { sb = 0; sph ^= 1; }
          if (++kst == NST) { kst = 0; kph ^= 1; }
          tpar ^= 1;
          CX_MARK(3);
          continue;
        }
        // ring positions of the tile after this one; probe its barriers now (non-blocking), use
        // the answer when the last chunk of this tile has been loaded
        int sb_n = sb + 1, kst_n = kst + 1;
        uint32_t sph_n = sph, kph_n = kph;
        if (sb_n == MODB) { sb_n = 0; sph_n ^= 1; }
        if (kst_n == NST) { kst_n = 0; kph_n ^= 1; }
        // S buffer + phase of the next tile of THIS warp (two ring steps ahead with alternating teams)
        int sb_p = sb_n;
        uint32_t sph_p = sph_n;
        if (STEP == 2 && ++sb_p == MODB) { sb_p = 0; sph_p ^= 1; }
        const int buf_n = (NQ == 2) ? team * MODB + sb_p : sb_p;
        // (no probe of the next K stage: its k_full phase completed before the MMA warp issued the
        // S tile this probe is about, so the r_j values are in shared memory once s_full is)
        const bool probe = CW == 16 && jj + STEP < n;
        if (probe) mbar_probe_issue(bar_s_full(buf_n), sph_p);
        have = false;

        // Chunks are double-buffered in registers: the tcgen05.ld of chunk c + 1 (or of the next
        // tile's first chunk, if its S tile is already complete) is in flight while chunk c is
        // processed.  keys [col_off + c*16, +16) -> P columns col_off + c*8 .. +8.
        auto process = [&](uint32_t (&v)[CW], int c) {
          uint32_t pk[CW / 2];
          if (POLY_A == POLY_B || !(wgi & 1))
            softmax_chunk<BWD, POLY_A, DEG, CW, RANK>(v, pk, rk + c * CW, cx, special, kbase + c * CW,
                                                      acc_m, acc_p, cmode);
          else
            softmax_chunk<BWD, POLY_B, DEG, CW, RANK>(v, pk, rk + c * CW, cx, special, kbase + c * CW,
                                                      acc_m, acc_p, cmode);
          // P (bf16, 2 keys per column) overwrites S columns this thread has already read
          if (BWD) tmem_st_pk(s_addr + c * (CW / 2), pk);
        };
        if (CW == 32) {
          // 32-column chunks, one register buffer (two would spill under the 96-register cap):
          // load -> wait -> process, no cross-tile prefetch
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (c > 0) tmem_ld_cw(s_addr + c * CW, v0);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < CW; ++i) asm volatile("" : "+r"(v0[i]));
            if (!BWD && c == NCH - 1) {  // every column of this S tile is in registers
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_sm_done(buf));
            }
            CX_MARK(1);
            process(v0, c);
          }
        } else {
#pragma unroll
        for (int c = 0; c < NCH; c += 2) {
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < CW; ++i) asm volatile("" : "+r"(v0[i]));
          tmem_ld_cw(s_addr + (c + 1) * CW, v1);
          CX_MARK(1);
          process(v0, c);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < CW; ++i) asm volatile("" : "+r"(v1[i]));
          if (c + 2 < NCH) {
            tmem_ld_cw(s_addr + (c + 2) * CW, v0);
          } else {
            if (!BWD) {  // every column of this S tile is in registers: hand the buffer back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_sm_done(buf));
            }
            if (probe && mbar_probe_result()) {  // start on the next tile: its S is already there
              tc_fence_after();
              tmem_ld_cw(lane_base + buf_n * 128, v0);
              have = true;
            }
          }
          CX_MARK(1);
          process(v1, c + 1);
        }
        }
        if (BWD) {  // hand P over right away: the P.Z MMA and the next S tile queue behind it
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sm_done(buf));
        }
        sb = sb_n;
        sph = sph_n;
        kst = kst_n;
        kph = kph_n;
        tpar ^= 1;
        CX_MARK(3);
      }

      if (!BWD) {
        if (valid && !remote) {
          const float2 sm = __fadd2_rn(acc_m[0], acc_m[1]), sp = __fadd2_rn(acc_p[0], acc_p[1]);
          atomicAdd(p.l_out + row, (sm.x + sm.y) + kscale * (sp.x + sp.y));
          if (RANK && row < p.pos_split && cx.cnt) atomicAdd(p.rank_out + row, cx.cnt);
        }
        if (SYM && any_upper) {  // fragment-layout partial row sums: 4 lanes per row
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float v = (racc_m[k].x + racc_m[k].y) + kscale * (racc_p[k].x + racc_p[k].y);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            const int lr = rb * C::RB_ROWS + q * 128 + w4 * 32 + 8 * k + (lane >> 2);
            if ((lane & 3) == 0 && lr < rows_here) atomicAdd((GRP ? p.g_out[grp] : p.l_out) + lr, v);
          }
        }
      } else {
        mbar_wait(bar_dz_full, useg & 1);
        tc_fence_after();
        // NQ == 2: team t owns accumulator t (D columns), its two warpgroups flush D/2 each;
        // NQ == 1: the four warpgroups share the single accumulator, max(D/4, 32) columns each
        constexpr int NCOL = (NQ == 2) ? D / 2 : (D / 4 >= 32 ? D / 4 : 32);
        const int col0 = (NQ == 2) ? (wgi & 1) * NCOL : wgi * NCOL;
        const bool has_cols = col0 < D;
        const uint32_t a_addr = tmem_base + (uint32_t(w4 * 32) << 16) + C::TMEM_DZ0 +
                                ((NQ == 2) ? team * D : 0) + col0;
        if (has_cols) {
#pragma unroll
          for (int c = 0; c < NCOL / 32; ++c) {
            uint32_t v[32];
            tmem_ld_x32(a_addr + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
            if (valid) {
              float* dst = p.dz_acc + (size_t)row * D + col0 + c * 32;
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                atomicAdd(reinterpret_cast<float4*>(dst + i),
                          make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                      __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dz_free);
      }
      CX_MARK(4);
      it += n;
      ++useg;
    }
#if MAAI_PROF
    if (lane == 0)
      for (int i = 0; i < 8; ++i) g_prof[((size_t)blockIdx.x * 20 + warp) * 8 + i] = cx.prof_a[i];
#endif
  }

  if (!BWD && !GRP && p.done_ctr) __threadfence();  // this thread's row / column sums before the CTA's ticket
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
  if (!BWD && !GRP && p.done_ctr) fwd_tail_finalize(p, smem_base + C::OFF_BAR, smem_base + C::OFF_Q);
  // cross-rank symmetric forward: this rank's partial sums for the peers (staged here, or added straight into
  // the owners' row sums) are complete once every CTA is here; in direct mode (done_ctr set) the CTA that
  // signalled then waits for the peers' signals and runs the per-row tail itself
  if (GRP) {
    const bool last = signal_when_grid_done(p.grp_sync, FLAG_L);
    if (last && p.done_ctr) tail_finalize_body(p, smem_base + C::OFF_Q, true);
  }
}

}  // namespace maai
