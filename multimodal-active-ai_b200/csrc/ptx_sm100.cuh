// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05
// (TMEM alloc / mma / commit / ld / st / fences).  No CUTLASS, no libcu++: these are the raw
// instructions the NT-Xent kernels in ntxent_tile.cuh are written against.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace maai {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (no-ops when the kernel was launched without the attribute)
// ----------------------------------------------------------------------------------------------
// Every kernel of the library is launched with programmatic stream serialization (maai_ntxent.cu,
// MAAI_PDL=0 turns the attribute off) and, at its top, lets the next kernel of the stream start launching
// (griddepcontrol.launch_dependents): the dependent's launch latency and prologue (barrier init, TMEM
// allocation, descriptor prefetch) overlap this kernel's work.  The dependent then blocks in
// griddepcontrol.wait until this grid has completed and its writes are visible.
// RULE: every thread of every kernel executes pdl_wait() before its first global access AND before it
// exits.  A grid whose threads skip the wait can complete before its predecessor has, and the kernel
// behind it would then see that predecessor's writes unordered: round 1's tile kernel had the trigger
// but no wait, which is why a CUDA-graph replay of two back-to-back steps returned wrong results with
// the attribute on (profiles/r1_tuning_log.md); with the wait in place the replay test passes with all
// five triggers (tests/test_gpu_parity.py::test_c_abi_is_cuda_graph_capturable, profiles/r2_tuning_log.md).
// Which kernels trigger early (bit mask, build-time for A/B runs): 1 normalize, 2 zero, 4 tile, 8 finalize,
// 16 dh.
#ifndef MAAI_PDL_TRIG
#define MAAI_PDL_TRIG 31
#endif
template <int WHO>
__device__ __forceinline__ void pdl_launch_dependents() {
  if (MAAI_PDL_TRIG & WHO) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// ... and, in that next kernel, wait until the previous kernel has completed and its writes are
// visible.  Must precede every global-memory access that depends on (or could clobber) earlier work.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// Cross-GPU flags (system scope): the in-kernel replacement of the symmetric-memory barrier launches
// ----------------------------------------------------------------------------------------------
// A producer rank signals "my stores of step `seq` have landed in your buffers" by a release store of
// `seq` into ITS word of every peer's flag block (peer-mapped NVLink address); a consumer spins on its
// LOCAL flag block with acquire loads until the word has reached `seq` (sequence numbers only grow, so
// there is nothing to reset).
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// generic-proxy observations (the acquire above) before async-proxy reads (TMA loads of what the peer wrote)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Waiting for a PEER is unbounded in principle (a rank may sit in a checkpoint or a data stall for many
// seconds, exactly like a rank arriving late at an NCCL collective), so the timeout is a wall-clock one, long
// (default 300 s, the caller's choice) and only a last resort against a dead rank: after it the kernel traps.
// Slow path out of line: the callers' hot code must not pay registers / stack for the spin and its printf.
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __noinline__ void wait_flag_spin(const unsigned int* flag, unsigned int seq, unsigned int timeout_s) {
  const unsigned long long t0 = global_timer_ns();
  const unsigned long long limit = (unsigned long long)(timeout_s ? timeout_s : 300u) * 1000000000ull;
  while (int(ld_acquire_sys_u32(flag) - seq) < 0) {
    if (global_timer_ns() - t0 > limit) {
      printf("maai: peer flag wait timeout (%u s) block %d thread %d flag %p want %u have %u\n", timeout_s, blockIdx.x,
             threadIdx.x, (const void*)flag, seq, ld_acquire_sys_u32(flag));
      __trap();
    }
    __nanosleep(128);
  }
}
__device__ __forceinline__ void wait_flag_ge(const unsigned int* flag, unsigned int seq, unsigned int timeout_s) {
  if (int(ld_acquire_sys_u32(flag) - seq) < 0) wait_flag_spin(flag, seq, timeout_s);
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Non-blocking probe (try_wait may suspend the thread for a while; test_wait never does)
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Split probe: MBAR_PROBE_DECL() once at function scope declares a PTX predicate; mbar_probe_issue()
// starts a non-blocking test_wait into it and mbar_probe_result() reads it later.  Unlike the
// one-statement form above, nothing consumes the predicate right behind the SYNCS instruction, so
// its ~150 clk latency overlaps whatever the caller puts in between (in-order issue would
// otherwise stall on the selp).  One probe may be outstanding per thread.
#define MBAR_PROBE_DECL() asm volatile(".reg .pred maai_probe_p;")
__device__ __forceinline__ void mbar_probe_issue(uint32_t bar, uint32_t parity) {
  asm volatile("mbarrier.test_wait.parity.shared::cta.b64 maai_probe_p, [%0], %1;" ::"r"(bar),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_probe_result() {
  uint32_t ok;
  asm volatile("selp.u32 %0, 1, 0, maai_probe_p;" : "=r"(ok)::"memory");
  return ok;
}
// Bounded wait: a pipeline bug must trap (-> CUDA error surfaced through the C ABI), never hang
// the GPU.  ~4e9 SM cycles is > 2 s at any clock; a healthy wait is microseconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("maai: mbarrier wait timeout block %d thread %d bar 0x%x parity %u\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// One lane of the (fully converged) warp; the same lane every time.  The callers keep all their
// control flow and descriptor arithmetic warp-uniform and elect only around the instruction, so the
// compiler can leave operands in uniform registers (a `lane == 0` region forces an R2UR waterfall
// loop of ~20 instructions in front of every UTCHMMA -- measured, profiles/r1_ncu_summary_v1.md).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates are {innermost (column, elements), outer (row)}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c_inner,
                                            int c_outer, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(c_inner), "r"(c_outer), "r"(bar)
      : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both addresses 16-B aligned)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM management
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05.mma (kind::f16: bf16 x bf16 -> fp32 in TMEM), single CTA
// ----------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every tcgen05 op previously issued by this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// Instruction descriptor for kind::f16 (bit layout: cute/arch/mma_sm100_desc.hpp InstrDescriptor):
//  [4,6) c_format (1 = F32)  [7,10) a_format (1 = BF16)  [10,13) b_format (1 = BF16)
//  [15] a_major (0 = K)      [16] b_major (0 = K, 1 = MN)
//  [17,23) N >> 3            [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

// Shared-memory matrix descriptor, 128-byte swizzle (bit layout: SmemDescriptor, same header):
//  [0,14) start >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//  [46,48) version = 1 (Blackwell)   [49,52) base offset = 0   [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                     uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// ----------------------------------------------------------------------------------------------
// tcgen05.ld / st: 32 lanes x 32-bit, N consecutive columns per thread
// (warp w of a warpgroup may only touch TMEM lanes [32*(w%4), 32*(w%4)+32))
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 256 bit (8 columns), x4 = 32 columns: thread t of the warp gets, in register 4g + k,
// lane (t / 4) + 8 (k >> 1) of the 16 addressed lanes and column 8 g + 2 (t % 4) + (k & 1)
// (verified on hardware by tools/tmem_layout.cu).  A thread therefore holds two rows x eight columns:
// the register layout of an mma accumulator fragment, which makes column sums cheap (three shuffle
// stages over the 8 lanes that share t % 4) -- used by the symmetric forward.
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
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
      "r"(r[15])
      : "memory");
}

// Register re-distribution between warpgroups (all 4 warps of a warpgroup must execute it)
template <int N> __device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// volatile twins: the compiler keeps volatile asm statements in source order relative to each other, which is
// how the backward's software pipeline (MAAI_BWD_ILV) keeps its exponentials and packs interleaved
__device__ __forceinline__ float ex2_approx_v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2_v(float lo, float hi) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// 2^(s*c) for a pair of dot products on the FMA/ALU pipes (no MUFU), |s*c| <= 126.
// One FFMA2 forms t = s*c + 1.5*2^23, which rounds y = s*c to the nearest integer n in the low
// mantissa bits; f = y - n in [-0.5, 0.5] comes from a second FFMA2 (single rounding of s*c - n);
// a minimax polynomial gives 2^f (max rel. error 7.5e-5 for degree 3, 2.7e-6 for degree 4, fitted
// offline) and one LEA per element adds n << 23 to the exponent field.  6 (degree 3) packed
// FMA-pipe instructions + 2 ALU instructions per PAIR, against 1 FFMA2 + 2 MUFU.EX2 (8 clk each per
// warp) on the MUFU path.  The caller folds the constant 2^(-c) of exp2(s*c - c) in afterwards.
__device__ __forceinline__ float exp_insert(float p, float t) {
  uint32_t r;
  asm("{\n\t"
      ".reg .b32 s;\n\t"
      "shl.b32 s, %2, 23;\n\t"
      "add.s32 %0, s, %1;\n\t"
      "}"
      : "=r"(r)
      : "r"(__float_as_uint(p)), "r"(__float_as_uint(t)));
  return __uint_as_float(r);
}
template <int DEG>
__device__ __forceinline__ float2 exp2_dot_poly2(float2 s, float2 c) {
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 t = __ffma2_rn(s, c, magic);
  const float2 nn = __fadd2_rn(make_float2(-t.x, -t.y), magic);  // -n, exact
  const float2 f = __ffma2_rn(s, c, nn);
  float2 pl;
  if (DEG == 3) {
    pl = __ffma2_rn(f, make_float2(5.517165389e-02f, 5.517165389e-02f),
                    make_float2(2.426111210e-01f, 2.426111210e-01f));
    pl = __ffma2_rn(pl, f, make_float2(6.932609883e-01f, 6.932609883e-01f));
    pl = __ffma2_rn(pl, f, make_float2(9.999280737e-01f, 9.999280737e-01f));
  } else {
    pl = __ffma2_rn(f, make_float2(9.570100157e-03f, 9.570100157e-03f),
                    make_float2(5.591786018e-02f, 5.591786018e-02f));
    pl = __ffma2_rn(pl, f, make_float2(2.402474487e-01f, 2.402474487e-01f));
    pl = __ffma2_rn(pl, f, make_float2(6.931218148e-01f, 6.931218148e-01f));
    pl = __ffma2_rn(pl, f, make_float2(9.999992614e-01f, 9.999992614e-01f));
  }
  return make_float2(exp_insert(pl.x, t.x), exp_insert(pl.y, t.y));
}

// two fp32 -> packed bf16x2 (lo in bits [0,16), hi in bits [16,32)), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

}  // namespace maai
