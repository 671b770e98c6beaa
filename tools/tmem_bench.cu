// Microbenchmark (sm_100a): throughput of tcgen05.ld / tcgen05.st from 4..16 warps, of
// tcgen05.mma (SS and TS, M=128 N=128 K=16 bf16) and of both running concurrently on one SM.
// Answers: does reading an S tile out of TMEM compete with the MMAs that fill the next one?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bench tools/tmem_bench.cu
//   ./tools/tmem_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../multimodal-active-ai_b200/csrc/ptx_sm100.cuh"

using namespace maai;

struct Args {
  int ld_warps;    // 0, 4, 8, 16 softmax-like warps (warps 4..)
  int ld_mode;     // 0 none, 1 ld x32 + wait each, 2 ld x32 x4 then one wait, 3 st x16 (P write), 4 ld+st
  int ld_iters;    // tiles (128 lanes x 128 cols fp32) each warpgroup reads
  int mma_mode;    // 0 none, 1 SS, 2 TS (A from TMEM), 3 SS then TS alternating
  int mma_iters;   // tiles (8 MMAs each)
  int same_cols;   // 1: the loads read the columns the MMAs write
  long long* out;  // per block: [0] ld cycles, [1] mma cycles
};

template <int MM>
__global__ void __launch_bounds__(640, 1) bench(Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32768, bar = base + 65536, tptr = base + 65536 + 64;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 2) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 65536 + 64);

  if (warp == 1 && MM) {
    constexpr uint32_t IDESC = make_idesc_bf16(128, 128, 0);
    constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 128, 1);
    const uint64_t dk = make_sdesc_sw128(0, 16, 1024);
    const uint64_t dmn = make_sdesc_sw128(0, 16384, 1024);
    const uint64_t dmn256 = make_sdesc_sw128(0, 16384, 1024);
    constexpr uint32_t IDESC_256 = make_idesc_bf16(128, 256, 0);
    constexpr uint32_t IDESC_PV256 = make_idesc_bf16(128, 256, 1);
    constexpr uint32_t IDESC_64 = make_idesc_bf16(128, 64, 0);
    const long long t0 = clock64();
    // groups of 8 tiles; one group stays in flight while the previous one is waited for
    for (int it = 0; it < a.mma_iters; ++it) {
      constexpr int m = MM;
      if (m >= 9) {  // straight-line pairs of tiles: it even -> first kind (8 MMAs), it odd -> second kind
        if (it & 1) continue;
        if (elect_one()) {
          const uint32_t d0 = tmem + 256, d1 = tmem + 384;
          auto first = [&](uint32_t d) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
              if (m == 11) umma_ts(d, tmem + 128 + k * 8, dk + (sA >> 4) + off, IDESC, k > 0);
              else umma_ss(d, dk + (sA >> 4) + off, dk + (sB >> 4) + off, IDESC, k > 0);
            }
          };
          auto second = [&](uint32_t d) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
              if (m == 12) umma_ss(d, dk + (sA >> 4) + off, dmn + (sB >> 4) + k * 128, IDESC_PV, k > 0);
              else if (m == 13) umma_ss(d, dk + (sA >> 4) + off, dk + (sB >> 4) + off, IDESC, k > 0);
              else umma_ts(d, tmem + 128 + k * 8, dmn + (sB >> 4) + k * 128, IDESC_PV, k > 0);
            }
          };
          if (m == 10) {  // SS, SS, TS, TS
            if (it & 2) { second(d1); second(d1); } else { first(d0); first(d0); }
          } else {
            first(d0);
            second(d1);
          }
          if ((it & 7) == 6) umma_commit(bar + 8 * ((it >> 3) & 1));
        }
        __syncwarp();
        if ((it & 7) == 6 && it >= 14) {
          const int g = (it >> 3) - 1;
          mbar_wait(bar + 8 * (g & 1), (g >> 1) & 1);
        }
        continue;
      }
      const bool wide = (m == 4 || m == 6 || m == 7);
      const uint32_t d = wide ? tmem + 256 : tmem + 256 + (it & 1) * 128;
      const bool ts = m == 2 || (m == 3 && (it & 1));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
          if (m == 4) umma_ss(d, dk + (sA >> 4) + off, dk + (sA >> 4) + off, IDESC_256, k > 0);  // B: 256 rows from sA
          else if (m == 5) umma_ts(d, tmem + 128 + k * 8, dk + (sB >> 4) + off, IDESC, k > 0);
          else if (m == 6) umma_ts(d, tmem + 128 + k * 8, dk + (sA >> 4) + off, IDESC_256, k > 0);
          else if (m == 7) umma_ts(d, tmem + 128 + k * 8, dmn256 + (sA >> 4) + k * 128, IDESC_PV256, k > 0);
          else if (m == 8) umma_ss(d, dk + (sA >> 4) + off, dk + (sB >> 4) + off, IDESC_64, k > 0);
          else if (!ts) umma_ss(d, dk + (sA >> 4) + off, dk + (sB >> 4) + off, IDESC, k > 0);
          else umma_ts(d, tmem + 128 + k * 8, dmn + (sB >> 4) + k * 128, IDESC_PV, k > 0);
        }
        if ((it & 7) == 7) umma_commit(bar + 8 * ((it >> 3) & 1));
      }
      __syncwarp();
      if ((it & 7) == 7 && it >= 15) {
        const int g = (it >> 3) - 1;
        mbar_wait(bar + 8 * (g & 1), (g >> 1) & 1);
      }
    }
    {
      const int g = (a.mma_iters >> 3) - 1;  // mma_iters is a multiple of 8
      mbar_wait(bar + 8 * (g & 1), (g >> 1) & 1);
      tc_fence_after();
    }
    const long long t1 = clock64();
    if (lane == 0) a.out[blockIdx.x * 2 + 1] = t1 - t0;
  } else if (warp >= 4 && warp < 4 + a.ld_warps && a.ld_mode) {
    const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16) + (a.same_cols ? 256 : 0);
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < a.ld_iters; ++it) {
      if (a.ld_mode == 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_x32(lane_base + c * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= v[i];
        }
      } else if (a.ld_mode == 2) {
        uint32_t v0[16], v1[16];
        tmem_ld_x16(lane_base, v0);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          tmem_ld_x16(lane_base + (c + 1) * 16, v1);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc ^= v0[i];
          tc_wait_ld();
          if (c + 2 < 8) tmem_ld_x16(lane_base + (c + 2) * 16, v0);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc ^= v1[i];
          tc_wait_ld();
        }
      } else if (a.ld_mode == 3) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = acc + i;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st_x16(lane_base + c * 16, pk);
        tc_wait_st();
        acc += 1;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_x32(lane_base + c * 32, v);
          tc_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = v[2 * i] ^ v[2 * i + 1];
          tmem_st_x16(lane_base + c * 16, pk);
        }
        tc_wait_st();
      }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) printf("x");
    if (warp == 4 && lane == 0) a.out[blockIdx.x * 2 + 0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 2 * sms);
  const int smem = 65536 + 2048;
#define FOR_MODES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)
#define SETATTR(M) cudaFuncSetAttribute(bench<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  FOR_MODES(SETATTR)
  auto run = [&](const char* name, int ld_warps, int ld_mode, int ld_iters, int mma_mode, int mma_iters, int same) {
    cudaMemset(out, 0, sizeof(long long) * 2 * sms);
    Args a{ld_warps, ld_mode, ld_iters, mma_mode, mma_iters, same, out};
#define LAUNCH(M) if (mma_mode == M) bench<M><<<sms, 640, smem>>>(a);
    FOR_MODES(LAUNCH)
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
      exit(1);
    }
    std::vector<long long> h(2 * sms);
    cudaMemcpy(h.data(), out, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
    double ld = 0, mm = 0;
    for (int i = 0; i < sms; ++i) {
      ld += h[2 * i];
      mm += h[2 * i + 1];
    }
    ld /= sms;
    mm /= sms;
    const double wgs = ld_warps / 4.0;
    printf("%-44s ld: %9.0f clk = %7.1f clk per warpgroup tile (%6.1f B/clk/SM)   mma: %9.0f clk = %6.1f clk/tile\n",
           name, ld, ld_iters ? ld / ld_iters : 0.0, ld > 0 ? wgs * ld_iters * 65536.0 / ld : 0.0, mm,
           mma_iters ? mm / mma_iters : 0.0);
  };
  const int N = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    run("ld x32+wait, 4 warps", 4, 1, N, 0, 0, 0);
    run("ld x32+wait, 8 warps", 8, 1, N, 0, 0, 0);
    run("ld x32+wait, 16 warps", 16, 1, N, 0, 0, 0);
    run("ld x16 pipelined, 16 warps", 16, 2, N, 0, 0, 0);
    run("st x16 (P write, 32 KB/tile), 16 warps", 16, 3, N, 0, 0, 0);
    run("ld x32 + st x16, 16 warps", 16, 4, N, 0, 0, 0);
    run("mma SS only", 0, 0, 0, 1, N, 0);
    run("mma TS only", 0, 0, 0, 2, N, 0);
    run("mma SS/TS alternating", 0, 0, 0, 3, N, 0);
    run("mma SS N=256 (per 2 tiles)", 0, 0, 0, 4, N, 0);
    run("mma TS K-major B N=128", 0, 0, 0, 5, N, 0);
    run("mma TS K-major B N=256 (per 2 tiles)", 0, 0, 0, 6, N, 0);
    run("mma TS MN-major B N=256 (per 2 tiles)", 0, 0, 0, 7, N, 0);
    run("mma SS N=64 (per half tile)", 0, 0, 0, 8, N, 0);
    run("mma alt SS / TS(MN) (branch, unpredicated)", 0, 0, 0, 9, N, 0);
    run("mma alt SS,SS / TS,TS", 0, 0, 0, 10, N, 0);
    run("mma alt TS(K-major) / TS(MN-major)", 0, 0, 0, 11, N, 0);
    run("mma alt SS(K) / SS(MN-major B)", 0, 0, 0, 12, N, 0);
    run("mma alt SS(D0) / SS(D1) control", 0, 0, 0, 13, N, 0);
    run("alt SS/TS + ld x32 + st x16 16 warps (N/2)", 16, 4, N / 2, 9, N, 0);
    run("mma SS + ld 4 warps (N tiles each)", 4, 1, N, 1, N, 0);
    run("mma SS + ld 16 warps (N/4 tiles per wg)", 16, 1, N / 4, 1, N, 0);
    run("mma SS + ld 16 warps (N/2 tiles per wg)", 16, 1, N / 2, 1, N, 0);
    run("mma SS + ld 16 warps same columns", 16, 1, N / 4, 1, N, 1);
    run("mma SS/TS + ld x32 + st x16 16 warps (N/4)", 16, 4, N / 4, 3, N, 0);
    run("mma SS/TS + ld x32 + st x16 16 warps (N/2)", 16, 4, N / 2, 3, N, 0);
    run("mma SS + ld x16 pipelined 16 warps (N/4)", 16, 2, N / 4, 1, N, 0);
  }
  return 0;
}
