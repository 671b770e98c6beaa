// Probe of the tcgen05.ld.16x256b register layout (sm_100a): every TMEM lane r of a 32-column block
// is filled with r*1000 + c through the 32x32b shape, then read back with 16x256b.x4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_layout tools/tmem_layout.cu
#include <cstdio>
#include "../multimodal-active-ai_b200/csrc/ptx_sm100.cuh"
using namespace maai;

__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
      "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

__global__ void probe(int* out) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tptr), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  const int row = warp * 32 + lane;
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = row * 1000 + c;
  tmem_st_x32(tmem + (uint32_t(warp * 32) << 16), v);
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
    tmem_ld_16x256b_x4(tmem + (uint32_t(warp * 32 + h * 16) << 16), r);
    tc_wait_ld();
    for (int i = 0; i < 16; ++i) out[((warp * 2 + h) * 32 + lane) * 16 + i] = int(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  int* d;
  cudaMalloc(&d, sizeof(int) * 4 * 2 * 32 * 16);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static int h[4 * 2 * 32 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int hh = 0; hh < 2; ++hh)
      for (int t = 0; t < 32; ++t)
        for (int i = 0; i < 16; ++i) {
          const int val = h[((w * 2 + hh) * 32 + t) * 16 + i];
          const int g = i >> 2, k = i & 3;
          const int erow = w * 32 + hh * 16 + (t >> 2) + 8 * (k >> 1), ecol = 8 * g + 2 * (t & 3) + (k & 1);
          if (val != erow * 1000 + ecol) {
            if (bad < 20) printf("w%d h%d t%d reg%d: got row %d col %d, expected row %d col %d\n", w, hh, t, i, val / 1000, val % 1000, erow, ecol);
            ++bad;
          }
        }
  printf("mismatches vs hypothesis (reg 4g+k: row 16h + t/4 + 8(k>>1), col 8g + 2(t%%4) + (k&1)): %d\n", bad);
  for (int i = 0; i < 16; ++i) printf("warp0 h0 t5 reg%d = row %d col %d\n", i, h[(0 * 32 + 5) * 16 + i] / 1000, h[(0 * 32 + 5) * 16 + i] % 1000);
  return 0;
}
