// O(M d) side kernels of the NT-Xent path (HBM-bound, coalesced, one warp per row):
//   normalize_cast_kernel  Objective.py:41-43  F.normalize -> bf16 z rows (+ 1/norm, positive cosine)
//   finalize_loss_kernel   Objective.py:76-79  loss = (1/b) sum_i [ln l_i + (1 - cos_i,pos)/tau];  r = 1/(b l)
//   dh_kernel              autograd tail: positive-pair term, 1/tau, grad_loss, F.normalize backward
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include <type_traits>
#include "ptx_sm100.cuh"

namespace maai {

enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- in-kernel peer synchronisation (include/maai_ntxent.h: maai_peer_sync) ----
// Flag block of a rank: kFlagKinds x kFlagStride words; word [kind][p] = last step whose stores of that
// kind rank p has completed into this rank's buffers.
constexpr int kFlagStride = 32;  // words per kind (world <= 16 today; one 128-byte line per kind)
enum : int { FLAG_Z = 0, FLAG_R = 1, FLAG_L = 2, kFlagKinds = 3 };
struct PeerSync {
  const unsigned long long* peer_flags;  // device array of `world` peer-mapped flag-block addresses; null = off
  const unsigned int* local_flags;       // this rank's own flag block
  unsigned int* counter;                 // CTA-done counter of the signalling kernel (zero between uses)
  unsigned int seq;
  int world;
  int rank;
  unsigned int timeout_s;                // wall-clock limit of a wait for a peer (0 = 300 s)
};
// One lane per peer waits (in parallel) until that peer's word of kind `kind` has reached s.seq.  Call from a
// full warp; follow with a CTA barrier before the other threads touch the peer-written data.
__device__ __forceinline__ void wait_all_peers(const unsigned int* local_flags, int kind, unsigned int seq, int world,
                                               int my_rank, unsigned int timeout_s) {
  const int lane = threadIdx.x & 31;
  if (lane < world && lane != my_rank) wait_flag_ge(local_flags + kind * kFlagStride + lane, seq, timeout_s);
  __syncwarp();
}
// Called by threads 0 .. world-1 of ONE CTA after every store of this kernel has been fenced at system
// scope and ordered before this point (per-thread __threadfence_system + CTA barrier + ticket counter).
__device__ __forceinline__ void signal_peers(const PeerSync& s, int kind) {
  if (int(threadIdx.x) < s.world)
    st_release_sys_u32(reinterpret_cast<unsigned int*>(s.peer_flags[threadIdx.x]) + kind * kFlagStride + s.rank, s.seq);
}
// Tail of a multi-CTA producer kernel: every thread has issued its peer stores; the last CTA to get here
// signals.  Must be reached by all threads of all CTAs (no early return before it).  Returns true in every
// thread of that last CTA (false everywhere when signalling is off).
__device__ __forceinline__ bool signal_when_grid_done(const PeerSync& s, int kind) {
  if (!s.peer_flags) return false;
  __shared__ unsigned int is_last;
  __syncthreads();  // every thread of the CTA has issued its peer / multicast stores ...
  if (threadIdx.x == 0) {
    __threadfence_system();  // ... and this fence is cumulative over what the barrier ordered before it (one
                             // fence per CTA: 256 of them cost K1 10 us at 16384 pairs, profiles/r2_tuning_log.md)
    is_last = (atomicAdd(s.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (is_last) {
    __threadfence_system();  // cumulativity: the other CTAs' fenced stores, observed through the counter
    signal_peers(s, kind);
    if (threadIdx.x == 0) *s.counter = 0u;  // ready for the next launch
  }
  return is_last != 0u;
}

constexpr float kNormEps = 1e-12f;  // F.normalize default eps (Objective.py:42-43)
constexpr int kMaxPerLane = 8;      // d <= 256 -> at most 8 elements per lane

// VEC consecutive elements of a row as ONE aligned vector load (4/8/16/32 bytes) when the caller has
// checked alignment (d % VEC == 0 and an aligned base), else element by element with a bound check.
template <typename T, int VEC>
__device__ __forceinline__ void load_row_chunk(const T* __restrict__ row, int e0, int d, bool vec_ok, float (&out)[VEC]) {
  struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };
  if (vec_ok) {  // d % VEC == 0: a chunk lies entirely inside or entirely outside the row
    if (e0 < d) {
      const Pack pk = *reinterpret_cast<const Pack*>(row + e0);
#pragma unroll
      for (int i = 0; i < VEC; ++i) out[i] = to_f32<T>(pk.v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) out[i] = 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) out[i] = (e0 + i < d) ? to_f32<T>(row[e0 + i]) : 0.f;
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void store_row_chunk(T* __restrict__ row, int e0, int d, bool vec_ok, const float (&v)[VEC]) {
  struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };
  if (vec_ok) {
    if (e0 < d) {
      Pack pk;
#pragma unroll
      for (int i = 0; i < VEC; ++i) pk.v[i] = from_f32<T>(v[i]);
      *reinterpret_cast<Pack*>(row + e0) = pk;
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      if (e0 + i < d) row[e0 + i] = from_f32<T>(v[i]);
  }
}

// store through an NVLink multicast (multimem) address: every device of the multicast object gets it
__device__ __forceinline__ void multimem_st(void* a, uint32_t v) {
  asm volatile("multimem.st.weak.global.bf16x2 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void multimem_st(void* a, uint2 v) {
  asm volatile("multimem.st.weak.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void multimem_st(void* a, uint4 v) {
  asm volatile("multimem.st.weak.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(a), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st_f32(float* a, float v) {
  asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(a), "f"(v) : "memory");
}

// body of K1 for one pair (see normalize_cast_kernel below)
template <typename T, int VEC>
__device__ __forceinline__ void normalize_pair(const T* __restrict__ h1, const T* __restrict__ h2, int b, int d, bool vec_ok,
                                               __nv_bfloat16* __restrict__ z_local,
                                               const unsigned long long* __restrict__ peer_base, unsigned long long mc_base,
                                               int world, int rank, float* __restrict__ inv_norm,
                                               float* __restrict__ pos_cos, int k, int lane) {
  constexpr int DP = VEC * 32;
  float a[VEC], c[VEC];
  load_row_chunk<T, VEC>(h1 + (size_t)k * d, lane * VEC, d, vec_ok, a);
  load_row_chunk<T, VEC>(h2 + (size_t)k * d, lane * VEC, d, vec_ok, c);
  float sa = 0.f, sc = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    sa += a[i] * a[i];
    sc += c[i] * c[i];
  }
  sa = warp_sum(sa);
  sc = warp_sum(sc);
  const float ia = 1.f / fmaxf(sqrtf(sa), kNormEps);
  const float ic = 1.f / fmaxf(sqrtf(sc), kNormEps);
  float dot = 0.f;
  __align__(16) __nv_bfloat16 za[VEC], zc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    za[i] = __float2bfloat16_rn(a[i] * ia);
    zc[i] = __float2bfloat16_rn(c[i] * ic);
    dot += __bfloat162float(za[i]) * __bfloat162float(zc[i]);  // same bf16 values the MMA multiplies
  }
  dot = warp_sum(dot);
  using Vec = typename std::conditional<VEC == 2, uint32_t, typename std::conditional<VEC == 4, uint2, uint4>::type>::type;
  const Vec va = *reinterpret_cast<const Vec*>(za), vc = *reinterpret_cast<const Vec*>(zc);
  if (z_local) {
    *reinterpret_cast<Vec*>(z_local + (size_t)k * DP + lane * VEC) = va;
    *reinterpret_cast<Vec*>(z_local + (size_t)(b + k) * DP + lane * VEC) = vc;
  }
  const size_t off_a = (((size_t)rank * 2 * b + k) * DP + lane * VEC) * sizeof(__nv_bfloat16);
  const size_t off_c = (((size_t)rank * 2 * b + b + k) * DP + lane * VEC) * sizeof(__nv_bfloat16);
  if (mc_base) {
    multimem_st(reinterpret_cast<char*>(mc_base) + off_a, va);
    multimem_st(reinterpret_cast<char*>(mc_base) + off_c, vc);
  } else if (peer_base) {
    for (int p = 0; p < world; ++p) {
      char* base = reinterpret_cast<char*>(peer_base[p]);
      *reinterpret_cast<Vec*>(base + off_a) = va;
      *reinterpret_cast<Vec*>(base + off_c) = vc;
    }
  }
  if (lane == 0) {
    inv_norm[k] = ia;
    inv_norm[b + k] = ic;
    pos_cos[k] = dot;
  }
}

// K1: F.normalize of both views (Objective.py:41-43) -> bf16 rows, 1/norm, positive cosine; one warp per
// pair k (rows k and b + k of the rank's stacked block); a lane owns DP/32 = VEC consecutive elements of
// both rows, so the input loads and the bf16 stores are 4..32-byte vectors and a warp moves whole rows.
// Destinations of the bf16 rows (any combination):
//   z_local    this rank's (2b, DP) block (single rank, or the NCCL all-gather's send slot)
//   mc_base    NVSwitch multicast mapping of every rank's (world, 2b, DP) key buffer: ONE multimem.st per
//              chunk, replicated by the switch                         } the cross-replica gather of
//   peer_base  table of `world` peer-mapped key buffers: unicast stores } Objective.py:52-53, 102-114 fused
// Columns [d, DP) are zero so the padded tile MMA sees exact zeros.
// The kernel also zero-fills `zero_words` 32-bit words at zero_fill (the step's accumulators: row sums,
// CTA-done counter, dz accumulator), which removes the separate memset / zero kernels of the step.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
normalize_cast_kernel(const T* __restrict__ h1, const T* __restrict__ h2, int b, int d, bool vec_ok,
                      __nv_bfloat16* __restrict__ z_local, const unsigned long long* __restrict__ peer_base,
                      unsigned long long mc_base, int world, int rank, float* __restrict__ inv_norm,
                      float* __restrict__ pos_cos, uint32_t* __restrict__ zero_fill, size_t zero_words,
                      const __grid_constant__ PeerSync sync) {
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents<1>();
  pdl_wait();
  if (zero_words) {  // grid-strided 16-byte stores (zero_fill is 16-byte aligned, checked by the host)
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    uint4* p4 = reinterpret_cast<uint4*>(zero_fill);
    for (size_t i = tid; i < zero_words / 4; i += nth) p4[i] = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (zero_words / 4) * 4 + tid; i < zero_words; i += nth) zero_fill[i] = 0u;
  }
  // pairs are strided over the grid's warps: the host caps the grid at one resident wave, so that with peer
  // signalling every CTA pays the system-scope fence (an NVLink round trip) once, not once per 8 pairs
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < b; k += nwarps)
    normalize_pair<T, VEC>(h1, h2, b, d, vec_ok, z_local, peer_base, mc_base, world, rank, inv_norm, pos_cos, k, lane);
  signal_when_grid_done(sync, FLAG_Z);  // in-kernel replacement of the barrier launch behind the gather
}

// K1 for chained views.  The reference's training loop feeds this step's
// outputs2 to the next step as hidden1 (Contrastive_Learning.py:700 "outputs1 = outputs2", consumed
// detached at :685), so the view-a rows of EVERY rank at step t are the view-b rows of step t-1, which
// every rank already holds normalised, in bf16, in its gathered key buffer of step t-1.  This variant
// therefore reads only hidden2: per pair k it normalises h2[k] -> view-b row (local / multicast / peer
// stores as in normalize_cast_kernel: half the gather payload), copies the view-b rows k of all `world`
// slots of the previous buffer into the view-a rows of the new one (a local, L2-resident copy instead of
// an NVLink transfer) and carries the view-a 1/norm over from the previous step's view b.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
normalize_chain_kernel(const T* __restrict__ h2, int b, int d, bool vec_ok,
                       const __nv_bfloat16* __restrict__ z_prev, const float* __restrict__ inv_prev,
                       __nv_bfloat16* __restrict__ z_new, const unsigned long long* __restrict__ peer_base,
                       unsigned long long mc_base, int world, int rank, float* __restrict__ inv_norm,
                       float* __restrict__ pos_cos, uint32_t* __restrict__ zero_fill, size_t zero_words,
                       const __grid_constant__ PeerSync sync) {
  constexpr int DP = VEC * 32;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents<1>();
  pdl_wait();
  if (zero_words) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    uint4* p4 = reinterpret_cast<uint4*>(zero_fill);
    for (size_t i = tid; i < zero_words / 4; i += nth) p4[i] = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (zero_words / 4) * 4 + tid; i < zero_words; i += nth) zero_fill[i] = 0u;
  }
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < b; k += nwarps) {
  using Vec = typename std::conditional<VEC == 2, uint32_t, typename std::conditional<VEC == 4, uint2, uint4>::type>::type;
  float c[VEC];
  load_row_chunk<T, VEC>(h2 + (size_t)k * d, lane * VEC, d, vec_ok, c);
  float sc = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) sc += c[i] * c[i];
  sc = warp_sum(sc);
  const float ic = 1.f / fmaxf(sqrtf(sc), kNormEps);
  // view-a rows of every slot <- view-b rows of the previous step's buffer
  Vec va_own = Vec();
  for (int q = 0; q < world; ++q) {
    const Vec v = *reinterpret_cast<const Vec*>(z_prev + ((size_t)q * 2 * b + b + k) * DP + lane * VEC);
    *reinterpret_cast<Vec*>(z_new + ((size_t)q * 2 * b + k) * DP + lane * VEC) = v;
    if (q == rank) va_own = v;
  }
  __align__(16) __nv_bfloat16 za[VEC], zc[VEC];
  *reinterpret_cast<Vec*>(za) = va_own;
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    zc[i] = __float2bfloat16_rn(c[i] * ic);
    dot += __bfloat162float(za[i]) * __bfloat162float(zc[i]);
  }
  dot = warp_sum(dot);
  const Vec vc = *reinterpret_cast<const Vec*>(zc);
  const size_t off_c = (((size_t)rank * 2 * b + b + k) * DP + lane * VEC) * sizeof(__nv_bfloat16);
  if (mc_base) {
    multimem_st(reinterpret_cast<char*>(mc_base) + off_c, vc);
  } else if (peer_base) {
    for (int p = 0; p < world; ++p) *reinterpret_cast<Vec*>(reinterpret_cast<char*>(peer_base[p]) + off_c) = vc;
  } else {
    *reinterpret_cast<Vec*>(reinterpret_cast<char*>(z_new) + off_c) = vc;
  }
  if (lane == 0) {
    inv_norm[k] = inv_prev[b + k];
    inv_norm[b + k] = ic;
    pos_cos[k] = dot;
  }
  }
  signal_when_grid_done(sync, FLAG_Z);
}

// Per-row tail of the forward (Objective.py:76-79): with e_pos = exp((cos_pos - 1)/tau) and l' = sum over
// the negatives,
//   lse_i - s_i,pos = ln(e_pos + l'_i) - ln(e_pos) = log1p(l'_i / e_pos)      (Objective.py:76-77)
//   loss = (1/b) * sum_{i < 2b} log1p(l'_i / e_pos(i))                          (Objective.py:79)
// and r_i = 1 / (b * (e_pos + l'_i)) for the backward.  Thread `tid` of `nth` takes rows tid, tid + nth, ...
// and returns its partial sum (fp64, fixed order: the reductions built on it are deterministic given l).
// r_out may be null.  peer_r (optional): device array of `world` peer-mapped base addresses of every
// rank's gathered r array; mc_r: multicast address of the same.  stage_tab (optional): `world` peer-mapped
// base addresses of every rank's (world, 2b) staging vectors (maai_ntxent_fwd_sym).
struct FinalizeArgs {
  float* l;
  const float* pos_cos;
  int b;
  float inv_tau;
  float* r_out;
  float* loss_out;
  const unsigned long long* peer_r;
  int world;
  int my_rank;
  float* mc_r;
  const unsigned long long* stage_tab;
  PeerSync sync;  // peer_flags != null: signal FLAG_R once the row factors have been stored to every rank
                  // (and, with stage_tab, wait for every peer's FLAG_L before pulling its staged sums)
};
__device__ __forceinline__ double finalize_rows(const FinalizeArgs& f, int tid, int nth) {
  double acc = 0.0;
  const float inv_b = 1.f / float(f.b);
  const float c1 = f.inv_tau * 1.4426950408889634f;
  for (int i = tid; i < 2 * f.b; i += nth) {
    float ln = __ldcg(f.l + i);  // written by other CTAs' atomics: read at L2
    if (f.stage_tab) {
      // symmetric forward across ranks: the other ranks hold partial row sums of this rank's anchors
      // (the tiles they computed for their own column sums) in slot my_rank of their staging vectors;
      // read them over NVLink (uncached: written by the peers' kernels before the barrier)
      for (int p = 0; p < f.world; ++p)
        if (p != f.my_rank)
          ln += __ldcv(reinterpret_cast<const float*>(f.stage_tab[p]) + (size_t)f.my_rank * 2 * f.b + i);
      f.l[i] = ln;  // the backward's positive-pair term needs the complete row sum
    }
    const float ep = ex2_approx(fmaf(__ldg(f.pos_cos + (i < f.b ? i : i - f.b)), c1, -c1));
    acc += double(log1pf(ln / ep));
    const float r = inv_b / (ep + ln);
    if (f.r_out) f.r_out[i] = r;
    // fused all-gather of the row factors: slot `rank` of every rank's r array (NVLink stores)
    if (f.mc_r) {
      multimem_st_f32(f.mc_r + (size_t)f.my_rank * 2 * f.b + i, r);
    } else if (f.peer_r) {
      for (int p = 0; p < f.world; ++p) reinterpret_cast<float*>(f.peer_r[p])[(size_t)f.my_rank * 2 * f.b + i] = r;
    }
  }
  return acc;
}
// Sum of the per-thread partials of one CTA, in a fixed order (warp shuffle tree, then warp 0 over the
// per-warp values).  `part` = shared scratch of >= 32 doubles.  Valid in thread 0.
__device__ __forceinline__ double finalize_block_sum(double acc, double* part) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  double v = 0.0;
  if (threadIdx.x < 32) {
    v = (threadIdx.x < ((blockDim.x + 31) >> 5)) ? part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

// Stand-alone finalize: one thread-block cluster of 8 CTAs (8192 threads), one value per CTA written into
// CTA 0's shared memory over DSMEM, summed in rank order.  (A single 1024-thread block took 45 us at 65536
// rows.)  Used for large batches and after the barrier of the cross-rank symmetric forward; small and
// medium batches finalize in the tail of the tile kernel itself (last CTA done, ntxent_tile.cuh).
constexpr int kFinalizeCluster = 8;
__global__ void __cluster_dims__(kFinalizeCluster, 1, 1) __launch_bounds__(1024)
finalize_loss_kernel(const __grid_constant__ FinalizeArgs f) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double part[32];
  __shared__ double cta_part[kFinalizeCluster];  // CTA 0's copy collects one value per CTA
  const unsigned rank = cluster.block_rank();
  pdl_launch_dependents<8>();
  pdl_wait();
  if (f.stage_tab && f.sync.peer_flags) {
    // cross-rank symmetric forward: the peers' staged partial sums (maai_ntxent_fwd_sym_tiles on THEIR GPUs)
    // must be complete before they are pulled: one waiting thread per peer, then the CTA barrier
    if (threadIdx.x < 32) wait_all_peers(f.sync.local_flags, FLAG_L, f.sync.seq, f.world, f.my_rank, f.sync.timeout_s);
    __syncthreads();
  }
  const double acc = finalize_rows(f, rank * blockDim.x + threadIdx.x, kFinalizeCluster * blockDim.x);
  if (f.sync.peer_flags) __threadfence_system();  // this thread's row-factor stores, before the cluster barrier
  const double v = finalize_block_sum(acc, part);
  if (threadIdx.x == 0) cluster.map_shared_rank(cta_part, 0)[rank] = v;
  cluster.sync();
  if (rank == 0) {
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int r = 0; r < kFinalizeCluster; ++r) t += cta_part[r];
      *f.loss_out = float(t / double(f.b));
    }
    if (f.sync.peer_flags) {  // all 8 CTAs have fenced their stores and passed the cluster barrier
      __threadfence_system();
      signal_peers(f.sync, FLAG_R);
    }
  }
}

// Zero fill of an accumulator (replaces cudaMemsetAsync so that the launch chain stays programmatic).
__global__ void __launch_bounds__(256) zero_kernel(uint32_t* __restrict__ p, size_t n_words) {
  pdl_launch_dependents<2>();
  pdl_wait();
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
    uint4* p4 = reinterpret_cast<uint4*>(p);
    for (size_t i = tid; i < n_words / 4; i += nth) p4[i] = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (n_words / 4) * 4 + tid; i < n_words; i += nth) p[i] = 0u;
  } else {
    for (size_t i = tid; i < n_words; i += nth) p[i] = 0u;
  }
}

// One warp per anchor row i of the views that need a gradient; a lane owns VEC = dp/32 consecutive
// elements (vector loads of h and of the fp32 accumulator, vector stores of dh).
//   dz_i = (g / tau) * (A_i + cpos_i z_pos(i)),  A = dz_acc (+ dz_extra) (fp32, stride dp, positive column excluded)
//   cpos_i = [e_pos/(e_pos + l'_i) - 1]/b  (+ the same with l'_pos when the key side is kept)
//          = -(1/b) [ l'_i/(e_pos + l'_i) + key_grad * l'_pos/(e_pos + l'_pos) ]
// i.e. the positive pair's softmax-minus-target coefficient without any cancellation.
//   dh_i = inv_i * (dz_i - z_i (z_i . dz_i))       (rows with ||h|| < eps: dh = dz * inv)
// z_i, z_pos are recomputed in fp32 from h (not the bf16 copies).
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
dh_kernel(const float* __restrict__ dz_acc, const float* __restrict__ dz_extra, const T* __restrict__ h1,
          const T* __restrict__ h2,
          const float* __restrict__ inv_norm, const float* __restrict__ grad_loss,
          const float* __restrict__ lneg, const float* __restrict__ pos_cos, int b, int d, bool vec_ok,
          float inv_tau, int key_grad, int need_mask, T* __restrict__ dh1, T* __restrict__ dh2) {
  constexpr int DP = VEC * 32;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents<16>();
  pdl_wait();
  // rows are enumerated over the views that need a gradient
  const int both = (need_mask == 3);
  const int nrows = both ? 2 * b : b;
  if (w >= nrows) return;
  const int i = both ? w : ((need_mask & 1) ? w : b + w);
  const int view = i >= b;
  const int k = view ? i - b : i;
  const T* hi = (view ? h2 : h1) + (size_t)k * d;
  const T* hp = (view ? h1 : h2) + (size_t)k * d;
  const float inv_i = inv_norm[i];
  const float inv_p = inv_norm[view ? k : b + k];
  const float gs = grad_loss[0] * inv_tau;
  const int ip = view ? k : b + k;  // local row of the positive
  const float c1 = inv_tau * 1.4426950408889634f;
  const float e_pos = ex2_approx(fmaf(pos_cos[k], c1, -c1));
  const float li = __ldcg(lneg + i), lp = __ldcg(lneg + ip);
  const float cpos = -(li / (e_pos + li) + (key_grad ? lp / (e_pos + lp) : 0.f)) / float(b);
  const int e0 = lane * VEC;
  float z[VEC], zp[VEC], a[VEC], dz[VEC];
  load_row_chunk<T, VEC>(hi, e0, d, vec_ok, z);
  load_row_chunk<T, VEC>(hp, e0, d, vec_ok, zp);
  // the accumulator rows are DP wide and 16-byte aligned: always vector loads (L2: written by atomics)
  {
    const float* src = dz_acc + (size_t)i * DP + e0;
#pragma unroll
    for (int j = 0; j < VEC; j += 2) {
      const float2 t = __ldcg(reinterpret_cast<const float2*>(src + j));
      a[j] = t.x;
      a[j + 1] = t.y;
    }
    if (dz_extra) {  // key-side sums that arrived by reduce-scatter
      const float* ex = dz_extra + (size_t)i * DP + e0;
#pragma unroll
      for (int j = 0; j < VEC; j += 2) {
        const float2 t = __ldcg(reinterpret_cast<const float2*>(ex + j));
        a[j] += t.x;
        a[j + 1] += t.y;
      }
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    z[j] *= inv_i;
    dz[j] = (e0 + j < d) ? gs * (a[j] + cpos * (zp[j] * inv_p)) : 0.f;
    dot += z[j] * dz[j];
  }
  dot = warp_sum(dot);
  const bool clamped = inv_i >= 1.f / kNormEps;  // ||h|| < eps: z = h/eps, plain scaling
  float out[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) out[j] = inv_i * (clamped ? dz[j] : (dz[j] - z[j] * dot));
  store_row_chunk<T, VEC>((view ? dh2 : dh1) + (size_t)k * d, e0, d, vec_ok, out);
}

}  // namespace maai
