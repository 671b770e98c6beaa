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
// ("stream-K"); a maximal run of items inside one row block is a segment.  Roles:
//   warp 0 lane 0 : TMA producer   (Q tiles once per segment, K tiles + r_j through a ring)
//   warp 1 lane 0 : tcgen05.mma issuer (S = Q K^T into TMEM; BWD also A += P Z_J with P read
//                   from TMEM and Z_J read from the *same* smem tile as an MN-major operand)
//   warp 2        : TMEM allocator
//   warps 4-11 / 12-19 : two "slots" of 256 threads.  A slot owns one S tile at a time; its two
//                   warpgroups split the 128 key columns (64 each), one TMEM lane (= anchor row)
//                   per thread: tcgen05.ld S -> exp2 -> row sum (FWD) / P = E (r_i + r_j) -> bf16
//                   -> TMEM (BWD).  Four softmax warps per SM sub-partition hide tcgen05.ld / MUFU
//                   latency (ncu of the 1-warpgroup-per-slot version: profiles/r1_ncu_summary_v1.md).
// D <= 128: a row block is two 128-row Q tiles, slot s owns Q tile s (each K tile feeds both).
// D == 256: a row block is one Q tile, the slots take alternate key tiles.
#pragma once
#include "ptx_sm100.cuh"

namespace maai {

struct TileParams {
  int m_loc;             // anchor rows covered by tmap_q
  int m_glob;            // key rows covered by tmap_k
  int row_global_base;   // global (key-space) index of anchor row 0
  int pos_split;         // anchor rows < pos_split have their positive at +pos_delta, others at -pos_delta
  int pos_delta;         //      (= b; pos_split = b - first anchor row of this launch)
  int nrb;               // row blocks
  int nkt;               // key tiles (128 keys each)
  float c1;              // log2(e) / tau
  const float* r_row;    // BWD: row factor per anchor row (m_loc floats)
  const float* r_col;    // BWD: column factor per global key, padded to a multiple of 128 floats
  float* l_out;          // FWD: row sums, m_loc floats, pre-zeroed
  float* dz_acc;         // BWD: m_loc x D fp32, pre-zeroed
  uint32_t pv_lbo;       // BWD: leading / stride byte offsets of the MN-major Z_J operand
  uint32_t pv_sbo;
};

template <int D, bool BWD>
struct TileCfg {
  static_assert(D == 64 || D == 128 || D == 256, "padded embedding dim must be 64, 128 or 256");
  static constexpr int NQ = (D <= 128) ? 2 : 1;        // Q tiles per row block
  static constexpr int RB_ROWS = 128 * NQ;
  static constexpr int KT = 128;                       // keys per tile
  static constexpr int CHUNKS = D / 64;                // 128-byte swizzle chunks per row
  static constexpr int CHUNK_BYTES = 128 * 128;        // 128 rows x 128 B
  static constexpr int TILE_BYTES = CHUNKS * CHUNK_BYTES;
  static constexpr int NST = (D == 256) ? 2 : 4;       // K ring depth
  static constexpr int SBUF = BWD ? 1 : 2;             // S buffers per slot
  static constexpr int TMEM_S0 = 0;                    // S buffers: slot s, buffer u -> (s*SBUF+u)*128
  static constexpr int TMEM_DZ0 = 256;                 // BWD accumulators: 256 + q*D
  static constexpr int NTHREADS = 640;
  // shared memory carve-up (offsets from a 1024-B aligned base)
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NQ * TILE_BYTES;
  static constexpr int OFF_RK = OFF_K + NST * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_RK + NST * KT * 4;
  static constexpr int NBAR = 2 + 2 * NST + 2 * 2 * SBUF + 2;
  static constexpr int OFF_TMEMPTR = OFF_BAR + NBAR * 8;
  static constexpr int SMEM_BYTES = OFF_TMEMPTR + 16 + 1024;  // + alignment slack
};

template <int D, bool BWD>
__global__ void __launch_bounds__(640, 1)
ntxent_tile_kernel(const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_k, const TileParams p) {
  using C = TileCfg<D, BWD>;
  constexpr int NQ = C::NQ, NST = C::NST, SBUF = C::SBUF;

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
  auto bar_s_full = [&](int slot, int u) { return bar0 + (2 + 2 * NST + slot * SBUF + u) * 8; };
  auto bar_sm_done = [&](int slot, int u) {
    return bar0 + (2 + 2 * NST + 2 * SBUF + slot * SBUF + u) * 8;
  };
  const uint32_t bar_dz_full = bar0 + (2 + 2 * NST + 4 * SBUF) * 8;
  const uint32_t bar_dz_free = bar0 + (2 + 2 * NST + 4 * SBUF + 1) * 8;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + C::OFF_TMEMPTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- this CTA's contiguous range of (row block, key tile) items ----
  const long long total = (long long)p.nrb * p.nkt;
  const long long it_begin = total * blockIdx.x / gridDim.x;
  const long long it_end = total * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(bar_q_full, 1);
    mbar_init(bar_q_empty, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_k_full(s), 1);
      mbar_init(bar_k_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s)
      for (int u = 0; u < SBUF; ++u) {
        mbar_init(bar_s_full(s, u), 1);
        mbar_init(bar_sm_done(s, u), 256);
      }
    mbar_init(bar_dz_full, 1);
    mbar_init(bar_dz_free, 512);
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

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t t = 0, useg = 0;
      for (long long it = it_begin; it < it_end;) {
        const int rb = int(it / p.nkt), j0 = int(it % p.nkt);
        const int n = int(min((long long)(p.nkt - j0), it_end - it));
        mbar_wait(bar_q_empty, (useg & 1) ^ 1);
        mbar_arrive_expect_tx(bar_q_full, NQ * C::TILE_BYTES);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
          for (int c = 0; c < C::CHUNKS; ++c)
            tma_load_2d(sQ + q * C::TILE_BYTES + c * C::CHUNK_BYTES, &tmap_q, c * 64,
                        rb * C::RB_ROWS + q * 128, bar_q_full);
        for (int jj = 0; jj < n; ++jj, ++t) {
          const int st = t % NST;
          mbar_wait(bar_k_empty(st), ((t / NST) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_k_full(st), C::TILE_BYTES + (BWD ? C::KT * 4 : 0));
#pragma unroll
          for (int c = 0; c < C::CHUNKS; ++c)
            tma_load_2d(sK + st * C::TILE_BYTES + c * C::CHUNK_BYTES, &tmap_k, c * 64,
                        (j0 + jj) * C::KT, bar_k_full(st));
          if (BWD)
            bulk_load_1d(sRK + st * C::KT * 4, p.r_col + (size_t)(j0 + jj) * C::KT, C::KT * 4,
                         bar_k_full(st));
        }
        it += n;
        ++useg;
      }
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
    uint32_t t = 0;          // key tiles consumed so far (ring position)
    uint32_t useg = 0;
    uint32_t u0 = 0, u1 = 0;  // S tiles issued per slot
    uint32_t sig = 0;         // S tiles issued in total (NQ == 1: slot = sig & 1)

    // S tile: slot `slot` <- Q tile `q` x K stage `st`
    auto issue_s = [&](int slot, int q, int st) {
      const uint32_t us = slot ? u1 : u0;
      const int ub = us % SBUF;
      if (!BWD) {  // FWD: wait until the softmax slot has drained this buffer
        mbar_wait(bar_sm_done(slot, ub), ((us / SBUF) & 1) ^ 1);
        tc_fence_after();
      }
      const uint32_t d_tmem = tmem_base + C::TMEM_S0 + (slot * SBUF + ub) * 128;
      const uint64_t ad = sdesc_k + ((sQ + q * C::TILE_BYTES) >> 4);
      const uint64_t bd = sdesc_k + ((sK + st * C::TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = ((k >> 2) * C::CHUNK_BYTES + (k & 3) * 32) >> 4;
          umma_ss(d_tmem, ad + off, bd + off, IDESC_S, k > 0);
        }
        umma_commit(bar_s_full(slot, ub));
      }
      __syncwarp();
      if (slot) ++u1; else ++u0;
      ++sig;
    };
    // BWD: A_q += P(slot) * Z_J(stage st); P was written over S by the softmax slot
    auto issue_pv = [&](int slot, int q, int st, uint32_t uidx, bool first) {
      mbar_wait(bar_sm_done(slot, 0), uidx & 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + C::TMEM_DZ0 + q * D;
      const uint32_t a_tmem = tmem_base + C::TMEM_S0 + slot * 128;
      const uint64_t bd = sdesc_mn + ((sK + st * C::TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < C::KT / 16; ++k) {
          // 16 keys = 16 smem rows of 128 B (2048 B); P of key half h (64 keys, 32 packed
          // columns) sits at S columns [h*64, h*64+32)
          umma_ts(d_tmem, a_tmem + (k >> 2) * 64 + (k & 3) * 8, bd + k * (2048 >> 4), IDESC_PV,
                  (first && k == 0) ? 0u : 1u);
        }
      }
      __syncwarp();
    };
    auto commit = [&](uint32_t bar) {
      if (elect_one()) umma_commit(bar);
      __syncwarp();
    };

#pragma unroll 1
    for (long long it = it_begin; it < it_end;) {
      const int j0 = int(it % p.nkt);
      const int n = int(min((long long)(p.nkt - j0), it_end - it));
      mbar_wait(bar_q_full, useg & 1);
      tc_fence_after();
      if (!BWD) {
#pragma unroll 1
        for (int jj = 0; jj < n; ++jj, ++t) {
          const int st = t % NST;
          mbar_wait(bar_k_full(st), (t / NST) & 1);
          tc_fence_after();
          if (NQ == 2) {
            issue_s(0, 0, st);
            issue_s(1, 1, st);
          } else {
            issue_s(sig & 1, 0, st);
          }
          commit(bar_k_empty(st));
        }
        commit(bar_q_empty);
      } else {
        // S tiles of this segment in issue order: sigma = 0 .. ns-1;
        // NQ == 2: sigma -> (key tile sigma >> 1, slot = Q tile = sigma & 1)
        // NQ == 1: sigma -> (key tile sigma, slot alternates with the running count)
        const int ns = n * NQ;
        const uint32_t t0 = t;
        const uint32_t slot_base = (NQ == 1) ? (sig & 1) : 0;
        auto slot_of = [&](int sg) { return (NQ == 2) ? (sg & 1) : int((slot_base + sg) & 1); };
        auto q_of = [&](int sg) { return (NQ == 2) ? (sg & 1) : 0; };
        auto kt_of = [&](int sg) { return (NQ == 2) ? (sg >> 1) : sg; };
        uint32_t upv0 = u0, upv1 = u1;  // per-slot index of the next P to consume
        auto issue_s_sigma = [&](int sg) {
          const uint32_t tk = t0 + kt_of(sg);
          const int st = tk % NST;
          if (NQ == 1 || (sg & 1) == 0) {  // first S tile that touches this key stage
            mbar_wait(bar_k_full(st), (tk / NST) & 1);
            tc_fence_after();
          }
          issue_s(slot_of(sg), q_of(sg), st);
        };
        issue_s_sigma(0);
        if (ns > 1) issue_s_sigma(1);
        if (ns <= 2) commit(bar_q_empty);
        // accumulators of the previous segment must have been flushed
        mbar_wait(bar_dz_free, (useg & 1) ^ 1);
        tc_fence_after();
#pragma unroll 1
        for (int sg = 0; sg < ns; ++sg) {
          const int slot = slot_of(sg), q = q_of(sg);
          const uint32_t tk = t0 + kt_of(sg);
          const int st = tk % NST;
          const bool first = (NQ == 2) ? (sg < 2) : (sg == 0);
          issue_pv(slot, q, st, slot ? upv1 : upv0, first);
          if (slot) ++upv1; else ++upv0;
          if (NQ == 1 || (sg & 1) == 1) commit(bar_k_empty(st));  // last reader of the stage
          if (sg + 2 < ns) {
            issue_s_sigma(sg + 2);
            if (sg + 3 >= ns) commit(bar_q_empty);  // that was the last read of the Q tiles
          }
        }
        commit(bar_dz_full);
        t = t0 + n;
      }
      it += n;
      ++useg;
    }
  } else if (warp >= 4) {
    // =========================== softmax slots ===========================
    const int sw = warp - 4;
    const int slot = sw >> 3;         // which S stream
    const int half = (sw >> 2) & 1;   // which 64 key columns of each S tile
    const int w4 = warp & 3;          // TMEM lane quarter this warp may touch
    const int row_in_tile = w4 * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(w4 * 32) << 16);
    const float c1 = p.c1;
    uint32_t t = 0, useg = 0, sig = 0, uu = 0;  // uu: S tiles consumed by this slot

    for (long long it = it_begin; it < it_end;) {
      const int rb = int(it / p.nkt), j0 = int(it % p.nkt);
      const int n = int(min((long long)(p.nkt - j0), it_end - it));
      const int q = (NQ == 2) ? slot : 0;
      const int row = rb * C::RB_ROWS + q * 128 + row_in_tile;  // anchor row (local)
      const bool valid = row < p.m_loc;
      const int grow = p.row_global_base + row;                 // same row in key space
      const int g0 = p.row_global_base + rb * C::RB_ROWS + q * 128;
      // key index of this row's positive (masked like the diagonal)
      const int gpos = row < p.pos_split ? grow + p.pos_delta : grow - p.pos_delta;
      float r_i = 0.f;
      if (BWD && valid) r_i = __ldg(p.r_row + row);
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      const uint32_t slot_base = (NQ == 1) ? (sig & 1) : 0;

      for (int jj = 0; jj < n; ++jj) {
        if (NQ == 1 && int((slot_base + jj) & 1) != slot) continue;
        const uint32_t tk = t + jj;
        const int st = tk % NST;
        const int ub = uu % SBUF;
        const int k0 = (j0 + jj) * C::KT;
        // tiles that hold a diagonal entry, a positive (a Q tile may straddle the view
        // boundary, so test both placements) or keys past the end need per-element predicates
        bool special = (k0 < g0 + 128 && g0 < k0 + C::KT) || (k0 + C::KT > p.m_glob);
        special = special || (k0 < g0 - p.pos_delta + 128 && g0 - p.pos_delta < k0 + C::KT) ||
                  (k0 < g0 + p.pos_delta + 128 && g0 + p.pos_delta < k0 + C::KT);
        if (BWD) mbar_wait(bar_k_full(st), (tk / NST) & 1);  // r_j of this stage has landed
        mbar_wait(bar_s_full(slot, ub), (uu / SBUF) & 1);
        tc_fence_after();
        const uint32_t s_addr = lane_base + C::TMEM_S0 + (slot * SBUF + ub) * 128 + half * 64;
        const float* rk = rk_gen + st * C::KT + half * 64;
        const int kbase = k0 + half * 64;

#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // single-buffered on purpose: four softmax warps per SM sub-partition cover the
          // tcgen05.ld latency, and 64 live S registers would spill under the 640-thread cap
          uint32_t v[32];
          tmem_ld_x32(s_addr + c * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(v[i]), c1, -c1)));
          if (special) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int kc = kbase + c * 32 + i;
              if (kc == grow || kc == gpos || kc >= p.m_glob) v[i] = 0u;
            }
          }
          if (!BWD) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              acc0 += __uint_as_float(v[i]);
              acc1 += __uint_as_float(v[i + 1]);
              acc2 += __uint_as_float(v[i + 2]);
              acc3 += __uint_as_float(v[i + 3]);
            }
          } else {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 rj = *reinterpret_cast<const float4*>(rk + c * 32 + i);
              pk[i / 2] = pack_bf16x2(__uint_as_float(v[i]) * (r_i + rj.x),
                                      __uint_as_float(v[i + 1]) * (r_i + rj.y));
              pk[i / 2 + 1] = pack_bf16x2(__uint_as_float(v[i + 2]) * (r_i + rj.z),
                                          __uint_as_float(v[i + 3]) * (r_i + rj.w));
            }
            // P (bf16, 2 keys per column) overwrites S columns this warpgroup has already read:
            // keys [half*64 + c*32, +32) -> columns half*64 + c*16 .. +16
            tmem_st_x16(s_addr + c * 16, pk);
          }
        }

        if (BWD) tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_sm_done(slot, ub));
        ++uu;
      }

      if (!BWD) {
        if (valid) atomicAdd(p.l_out + row, (acc0 + acc1) + (acc2 + acc3));
      } else {
        mbar_wait(bar_dz_full, useg & 1);
        tc_fence_after();
        // NQ == 2: slot s owns accumulator s (D columns), its two warpgroups flush D/2 each;
        // NQ == 1: the four warpgroups flush 64 columns each of the single 256-column accumulator
        constexpr int NCOL = (NQ == 2) ? D / 2 : 64;
        const int col0 = (NQ == 2) ? half * NCOL : (slot * 2 + half) * 64;
        const uint32_t a_addr = lane_base + C::TMEM_DZ0 + ((NQ == 2) ? slot * D : 0) + col0;
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
        tc_fence_before();
        mbar_arrive(bar_dz_free);
      }
      t += n;
      sig += n * NQ;
      it += n;
      ++useg;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace maai
