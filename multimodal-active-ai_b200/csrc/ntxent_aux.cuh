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

constexpr float kNormEps = 1e-12f;  // F.normalize default eps (Objective.py:42-43)
constexpr int kMaxPerLane = 8;      // d <= 256 -> at most 8 elements per lane

// One warp per pair k: rows k (view a) and b + k (view b) of the rank-local z block.
// z_out: (2b, DP) bf16, columns [d, DP) zero-filled so the padded tile MMA sees exact zeros.
template <typename T>
__global__ void __launch_bounds__(256)
normalize_cast_kernel(const T* __restrict__ h1, const T* __restrict__ h2, int b, int d, int dp,
                      __nv_bfloat16* __restrict__ z_out, float* __restrict__ inv_norm,
                      float* __restrict__ pos_cos) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents<1>();
  pdl_wait();
  if (k >= b) return;
  float a[kMaxPerLane], c[kMaxPerLane];
  float sa = 0.f, sc = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int e = lane + 32 * i;
    a[i] = (e < d) ? to_f32<T>(h1[(size_t)k * d + e]) : 0.f;
    c[i] = (e < d) ? to_f32<T>(h2[(size_t)k * d + e]) : 0.f;
    sa += a[i] * a[i];
    sc += c[i] * c[i];
  }
  sa = warp_sum(sa);
  sc = warp_sum(sc);
  const float ia = 1.f / fmaxf(sqrtf(sa), kNormEps);
  const float ic = 1.f / fmaxf(sqrtf(sc), kNormEps);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int e = lane + 32 * i;
    if (e < dp) {
      const __nv_bfloat16 za = __float2bfloat16_rn(a[i] * ia);
      const __nv_bfloat16 zc = __float2bfloat16_rn(c[i] * ic);
      z_out[(size_t)k * dp + e] = za;
      z_out[(size_t)(b + k) * dp + e] = zc;
      dot += __bfloat162float(za) * __bfloat162float(zc);  // same bf16 values the MMA multiplies
    }
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    inv_norm[k] = ia;
    inv_norm[b + k] = ic;
    pos_cos[k] = dot;
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

// K1 fused with the embedding all-gather (Objective.py:41-43 + :52-53, 102-114): the normalised bf16
// rows are stored straight into slot `rank` of EVERY rank's (world, 2b, DP) key buffer through
// peer-mapped (NVLink) pointers, so no collective kernel runs and the payload crosses the switch
// while the rows are being produced; a symmetric-memory barrier then orders the stores before any
// rank's tile kernel reads its buffer.  One warp per pair; a lane owns DP/32 consecutive elements
// of both rows, so every store is a 4/8/16-byte vector and a warp writes whole 128..512-B rows.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
normalize_cast_peer_kernel(const T* __restrict__ h1, const T* __restrict__ h2, int b, int d,
                           const unsigned long long* __restrict__ peer_base, unsigned long long mc_base,
                           int world, int rank, float* __restrict__ inv_norm, float* __restrict__ pos_cos) {
  constexpr int DP = VEC * 32;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents<1>();
  pdl_wait();
  if (k >= b) return;
  float a[VEC], c[VEC];
  float sa = 0.f, sc = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int e = lane * VEC + i;
    a[i] = (e < d) ? to_f32<T>(h1[(size_t)k * d + e]) : 0.f;
    c[i] = (e < d) ? to_f32<T>(h2[(size_t)k * d + e]) : 0.f;
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
  const size_t off_a = (((size_t)rank * 2 * b + k) * DP + lane * VEC) * sizeof(__nv_bfloat16);
  const size_t off_c = (((size_t)rank * 2 * b + b + k) * DP + lane * VEC) * sizeof(__nv_bfloat16);
  if (mc_base) {
    // NVSwitch multicast mapping of the same buffers: ONE store, replicated into every rank's copy by
    // the switch (multimem.st), instead of `world` unicast stores
    multimem_st(reinterpret_cast<char*>(mc_base) + off_a, va);
    multimem_st(reinterpret_cast<char*>(mc_base) + off_c, vc);
  } else {
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

// One thread-block cluster of 8 CTAs (8192 threads), deterministic: fixed thread-strided partial
// sums in fp64, one value per CTA written into CTA 0's shared memory over DSMEM, summed in rank
// order.  (A single 1024-thread block took 45 us at 65536 rows, 1.7 % of the step.)
// With e_pos = exp((cos_pos - 1)/tau) and l' = sum over negatives:
//   lse_i - s_i,pos = ln(e_pos + l'_i) - ln(e_pos) = log1p(l'_i / e_pos)      (Objective.py:76-77)
//   loss = (1/b) * sum_{i < 2b} log1p(l'_i / e_pos(i))                          (Objective.py:79)
// and r_i = 1 / (b * (e_pos + l'_i)) for the backward.  r_out may be null.  peer_r (optional): device
// array of `world` peer-mapped base addresses of every rank's gathered r array.  stage_tab (optional):
// `world` peer-mapped base addresses of every rank's (world, 2b) staging vectors (maai_ntxent_fwd_sym).
constexpr int kFinalizeCluster = 8;
__global__ void __cluster_dims__(kFinalizeCluster, 1, 1) __launch_bounds__(1024)
finalize_loss_kernel(float* __restrict__ l, const float* __restrict__ pos_cos, int b,
                     float inv_tau, float* __restrict__ r_out, float* __restrict__ loss_out,
                     const unsigned long long* __restrict__ peer_r = nullptr, int world = 1, int my_rank = 0,
                     float* mc_r = nullptr, const unsigned long long* __restrict__ stage_tab = nullptr) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double part[32];
  __shared__ double cta_part[kFinalizeCluster];  // CTA 0's copy collects one value per CTA
  const unsigned rank = cluster.block_rank();
  pdl_launch_dependents<8>();
  pdl_wait();
  double acc = 0.0;
  const float inv_b = 1.f / float(b);
  const float c1 = inv_tau * 1.4426950408889634f;
  for (int i = rank * blockDim.x + threadIdx.x; i < 2 * b; i += kFinalizeCluster * blockDim.x) {
    float ln = l[i];
    if (stage_tab) {
      // symmetric forward across ranks: the other ranks hold partial row sums of this rank's anchors
      // (the tiles they computed for their own column sums) in slot my_rank of their staging vectors;
      // read them over NVLink (uncached: written by the peers' kernels before the barrier)
      for (int p = 0; p < world; ++p)
        if (p != my_rank) ln += __ldcv(reinterpret_cast<const float*>(stage_tab[p]) + (size_t)my_rank * 2 * b + i);
      l[i] = ln;  // the backward's positive-pair term needs the complete row sum
    }
    const float ep = ex2_approx(fmaf(pos_cos[i < b ? i : i - b], c1, -c1));
    acc += double(log1pf(ln / ep));
    const float r = inv_b / (ep + ln);
    if (r_out) r_out[i] = r;
    // fused all-gather of the row factors: slot `rank` of every rank's r array (NVLink stores)
    if (mc_r) {
      multimem_st_f32(mc_r + (size_t)my_rank * 2 * b + i, r);
    } else if (peer_r) {
      for (int p = 0; p < world; ++p) reinterpret_cast<float*>(peer_r[p])[(size_t)my_rank * 2 * b + i] = r;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) cluster.map_shared_rank(cta_part, 0)[rank] = v;
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x == 0) {
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < kFinalizeCluster; ++r) v += cta_part[r];
    *loss_out = float(v * double(inv_b));
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

// One warp per anchor row i of the views that need a gradient.
//   dz_i = (g / tau) * (A_i + cpos_i z_pos(i)),  A = dz_acc (+ dz_extra) (fp32, stride dp, positive column excluded)
//   cpos_i = [e_pos/(e_pos + l'_i) - 1]/b  (+ the same with l'_pos when the key side is kept)
//          = -(1/b) [ l'_i/(e_pos + l'_i) + key_grad * l'_pos/(e_pos + l'_pos) ]
// i.e. the positive pair's softmax-minus-target coefficient without any cancellation.
//   dh_i = inv_i * (dz_i - z_i (z_i . dz_i))       (rows with ||h|| < eps: dh = dz * inv)
// z_i, z_pos are recomputed in fp32 from h (not the bf16 copies).
template <typename T>
__global__ void __launch_bounds__(256)
dh_kernel(const float* __restrict__ dz_acc, const float* __restrict__ dz_extra, const T* __restrict__ h1,
          const T* __restrict__ h2,
          const float* __restrict__ inv_norm, const float* __restrict__ grad_loss,
          const float* __restrict__ lneg, const float* __restrict__ pos_cos, int b, int d, int dp,
          float inv_tau, int key_grad, int need_mask, T* __restrict__ dh1, T* __restrict__ dh2) {
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
  const float li = lneg[i], lp = lneg[ip];
  const float cpos = -(li / (e_pos + li) + (key_grad ? lp / (e_pos + lp) : 0.f)) / float(b);
  float z[kMaxPerLane], dz[kMaxPerLane];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j) {
    const int e = lane + 32 * j;
    if (e < d) {
      z[j] = to_f32<T>(hi[e]) * inv_i;
      const float zp = to_f32<T>(hp[e]) * inv_p;
      float a = dz_acc[(size_t)i * dp + e];
      if (dz_extra) a += dz_extra[(size_t)i * dp + e];  // key-side sums that arrived by reduce-scatter
      dz[j] = gs * (a + cpos * zp);
      dot += z[j] * dz[j];
    } else {
      z[j] = 0.f;
      dz[j] = 0.f;
    }
  }
  dot = warp_sum(dot);
  const bool clamped = inv_i >= 1.f / kNormEps;  // ||h|| < eps: z = h/eps, plain scaling
  T* out = (view ? dh2 : dh1) + (size_t)k * d;
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j) {
    const int e = lane + 32 * j;
    if (e < d) out[e] = from_f32<T>(inv_i * (clamped ? dz[j] : (dz[j] - z[j] * dot)));
  }
}

}  // namespace maai
