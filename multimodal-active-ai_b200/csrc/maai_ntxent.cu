// C ABI (include/maai_ntxent.h) over the sm_100a NT-Xent kernels.  Host side only builds TMA
// descriptors and enqueues kernels on the caller's stream: no allocation, no synchronisation.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include <cuda.h>
#include <cuda_runtime.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include "../../include/maai_ntxent.h"
#include "ntxent_aux.cuh"
#include "ntxent_tile.cuh"

namespace {

thread_local std::string g_err;
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(const char* what, cudaError_t e) {
  return fail(MAAI_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define MAAI_CUDA(call)                                  \
  do {                                                   \
    cudaError_t _e = (call);                             \
    if (_e != cudaSuccess) return cuda_fail(#call, _e);  \
  } while (0)

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// bf16 (rows, d_pad) row-major -> boxes of 128 rows x 64 columns (128 B), 128-byte swizzle
int make_rows_tmap(CUtensorMap* m, const void* base, int rows, int d_pad) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(MAAI_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // cuTensorMapEncodeTiled is a DRIVER call: it needs the device's primary context current on THIS thread.
  // A thread that has only used cached allocations so far (the autograd engine's device thread running the
  // C++ binding's backward) may not have one bound yet -> CUDA_ERROR_INVALID_CONTEXT (201).  One runtime
  // call per (thread, device) binds it.
  static thread_local int ctx_bound_dev = -1;
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != ctx_bound_dev) {
    cudaFree(nullptr);
    ctx_bound_dev = dev;
  }
  cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d_pad * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[96];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (CUresult %d), rows=%d d_pad=%d", int(r),
             rows, d_pad);
    return fail(MAAI_E_CUDA, buf);
  }
  return MAAI_OK;
}

// Programmatic stream serialization (PDL) for every kernel of the library: the next kernel's launch
// latency and prologue overlap the tail of the previous one; each kernel calls griddepcontrol.wait before it
// touches global memory (see the rule in ptx_sm100.cuh).  On by default; MAAI_PDL=0 launches without the
// attribute (plain stream order), for A/B measurements.
bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("MAAI_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  ++g_launches;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int zero_words(void* p, size_t n_words, cudaStream_t s) {
  if (n_words == 0) return MAAI_OK;
  size_t blocks = (n_words / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 592) blocks = 592;
  cudaError_t e = launch_k(maai::zero_kernel, dim3((unsigned)blocks), dim3(256), 0, s, static_cast<uint32_t*>(p), n_words);
  if (e != cudaSuccess) return cuda_fail("zero_kernel", e);
  return MAAI_OK;
}

constexpr int kMaxDevices = 64;
int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices ? dev : -1;
}
// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs)
int sm_count() {
  static std::atomic<int> cache[kMaxDevices];
  const int dev = current_device();
  if (dev < 0) return 0;
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel, device)
template <typename K>
cudaError_t ensure_smem_attr(K kernel, int bytes, std::atomic<unsigned long long>& done_mask) {
  const int dev = current_device();
  if (dev < 0) return cudaErrorInvalidDevice;
  const unsigned long long bit = 1ull << dev;
  if (done_mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done_mask.fetch_or(bit, std::memory_order_release);
  return e;
}

// include/maai_ntxent.h maai_peer_sync (host struct, may be null) -> the device-side view
maai::PeerSync to_sync(const maai_peer_sync* sy, int world, int rank) {
  maai::PeerSync s{};
  if (sy && sy->peer_flag_bases) {
    s.peer_flags = reinterpret_cast<const unsigned long long*>(sy->peer_flag_bases);
    s.local_flags = sy->local_flags;
    s.counter = sy->counter;
    s.seq = sy->seq;
    s.world = world;
    s.rank = rank;
    s.timeout_s = sy->timeout_s;
  }
  return s;
}
int check_sync(const maai_peer_sync* sy, int world) {
  if (!sy) return MAAI_OK;
  if (!sy->peer_flag_bases || !sy->local_flags || !sy->counter) return fail(MAAI_E_ARG, "maai_peer_sync: null pointer");
  if (sy->seq == 0) return fail(MAAI_E_ARG, "maai_peer_sync: seq must be > 0");
  if (world > maai::kFlagStride) return fail(MAAI_E_SHAPE, "maai_peer_sync: world must be <= 32");
  return MAAI_OK;
}

struct RankArgs {
  const float* pos_cos = nullptr;
  int* rank_out = nullptr;
  int pos_period = 0;  // anchors spanning several rank slots (maai_ntxent_bwd_tiles)
  int pos_phase = 0;
  // forward: in-kernel finalize by the last CTA (null done_ctr = separate finalize launch)
  unsigned int* done_ctr = nullptr;
  maai::FinalizeArgs fin = {};
  // multi-rank: producer-side waits on the peers' flags (null = ordered by the caller's barrier)
  const unsigned int* wait_flags = nullptr;
  unsigned int wait_seq = 0;
  int wait_kind = 0, wait_my_slot = 0, wait_nslots = 1;
  unsigned int wait_timeout_s = 0;
};

template <int D, bool BWD, int NQ, bool RANK = false, bool SYM = false>
int launch_tile(const void* q_base, int m_loc, const void* k_base, int m_glob, int row_global_base,
                float inv_tau, const float* r_row, const float* r_col, float* l_out, float* dz_acc,
                int pos_split, int pos_delta, cudaStream_t s, RankArgs ra = RankArgs()) {
  using C = maai::TileCfg<D, BWD, NQ>;
  static std::atomic<unsigned long long> attr_done{0};
  cudaError_t attr_err = ensure_smem_attr(maai::ntxent_tile_kernel<D, BWD, NQ, RANK, SYM>, C::SMEM_BYTES, attr_done);
  if (attr_err != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(smem)", attr_err);
  CUtensorMap tq, tk;
  int rc;
  if ((rc = make_rows_tmap(&tq, q_base, m_loc, D)) != MAAI_OK) return rc;
  if ((rc = make_rows_tmap(&tk, k_base, m_glob, D)) != MAAI_OK) return rc;
  maai::TileParams p{};
  p.done_ctr = ra.done_ctr;
  p.fin = ra.fin;
  p.wait_flags = ra.wait_flags;
  p.wait_seq = ra.wait_seq;
  p.wait_kind = ra.wait_kind;
  p.wait_my_slot = ra.wait_my_slot;
  p.wait_nslots = ra.wait_nslots;
  p.wait_timeout_s = ra.wait_timeout_s;
  p.m_loc = m_loc;
  p.m_glob = m_glob;
  p.row_global_base = row_global_base;
  p.pos_split = pos_split;
  p.pos_delta = pos_delta;
  p.pos_period = ra.pos_period;
  p.pos_phase = ra.pos_phase;
  p.nrb = (m_loc + C::RB_ROWS - 1) / C::RB_ROWS;
  p.nkt = (m_glob + C::KT - 1) / C::KT;
  p.c1 = inv_tau * 1.4426950408889634f;
  p.r_row = r_row;
  p.r_col = r_col;
  p.l_out = l_out;
  p.dz_acc = dz_acc;
  p.pos_cos = ra.pos_cos;
  p.rank_out = ra.rank_out;
  // MN-major SW128 operand: LBO = distance between 64-column chunks, SBO = between 8-row groups
  static const uint32_t pv_lbo = getenv("MAAI_DEBUG_PV_LBO") ? atoi(getenv("MAAI_DEBUG_PV_LBO")) : C::CHUNK_BYTES;
  static const uint32_t pv_sbo = getenv("MAAI_DEBUG_PV_SBO") ? atoi(getenv("MAAI_DEBUG_PV_SBO")) : 1024;
  p.pv_lbo = pv_lbo;
  p.pv_sbo = pv_sbo;
  const long long items = SYM ? (long long)p.nrb * p.nkt - (long long)NQ * p.nrb * (p.nrb - 1) / 2
                              : (long long)p.nrb * p.nkt;
  int sms = sm_count();
  if (sms <= 0) return fail(MAAI_E_CUDA, "no CUDA device");
  const int grid = (int)(items < sms ? items : sms);
  MAAI_CUDA(launch_k(maai::ntxent_tile_kernel<D, BWD, NQ, RANK, SYM>, dim3(grid), dim3(C::NTHREADS),
                     C::SMEM_BYTES, s, tq, tk, p));
  return MAAI_OK;
}

// Q tiles per row block.  Forward (MUFU-bound): two, to halve the L2 -> smem traffic per flop.
// Backward: one, which leaves TMEM room for three S buffers (see ntxent_tile.cuh).
// MAAI_DEBUG_BWD_NQ=2 / MAAI_DEBUG_FWD_NQ=1 select the other layout for A/B measurements.
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <bool BWD>
int dispatch_tile(int d_pad, const void* q_base, int m_loc, const void* k_base, int m_glob,
                  int row_global_base, float inv_tau, const float* r_row, const float* r_col,
                  float* l_out, float* dz_acc, int pos_split, int pos_delta, cudaStream_t s,
                  RankArgs ra = RankArgs()) {
  static const int nq = BWD ? env_int("MAAI_DEBUG_BWD_NQ", 1) : env_int("MAAI_DEBUG_FWD_NQ", 2);
#define MAAI_LAUNCH(DD, NQQ)                                                                       \
  return launch_tile<DD, BWD, NQQ>(q_base, m_loc, k_base, m_glob, row_global_base, inv_tau, r_row, \
                                   r_col, l_out, dz_acc, pos_split, pos_delta, s, ra)
  switch (d_pad) {
    case 64:
      if (nq == 2) MAAI_LAUNCH(64, 2);
      MAAI_LAUNCH(64, 1);
    case 128:
      if (nq == 2) MAAI_LAUNCH(128, 2);
      MAAI_LAUNCH(128, 1);
    case 256:
      MAAI_LAUNCH(256, 1);
    default:
      return fail(MAAI_E_SHAPE, "d_pad must be 64, 128 or 256");
  }
#undef MAAI_LAUNCH
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_common(int b, int world, int rank) {
  if (b < 1) return fail(MAAI_E_ARG, "b must be >= 1");
  if (world < 1 || rank < 0 || rank >= world) return fail(MAAI_E_ARG, "need 0 <= rank < world");
  if ((long long)b * 2 * world > (1ll << 30)) return fail(MAAI_E_SHAPE, "world*2b exceeds 2^30 rows");
  return MAAI_OK;
}

}  // namespace

extern "C" {

// MAAI_DEBUG_SEGV=1: print a native backtrace on SIGSEGV (the Python fault handler shows Python frames only)
static void segv_handler(int sig) {
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "maai: SIGSEGV, native backtrace:\n";
  (void)!write(2, msg, sizeof msg - 1);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}
int maai_abi_version(void) {
  static const bool hooked = [] {
    const char* v = getenv("MAAI_DEBUG_SEGV");
    if (v && v[0] == '1') signal(SIGSEGV, segv_handler);
    return true;
  }();
  (void)hooked;
  return MAAI_ABI_VERSION;
}
const char* maai_last_error(void) { return g_err.c_str(); }
unsigned long long maai_launch_count(void) { return g_launches.load(); }

int maai_padded_dim(int d) {
  if (d < 1 || d > 256) return MAAI_E_SHAPE;
  return d <= 64 ? 64 : (d <= 128 ? 128 : 256);
}

size_t maai_ntxent_r_len(int b, int world) {
  const size_t m = (size_t)2 * b * world;
  return (m + 127) / 128 * 128;
}

size_t maai_ntxent_workspace_bytes(int b, int d_pad, int need_bwd) {
  if (b < 1 || (d_pad != 64 && d_pad != 128 && d_pad != 256)) return 0;
  const size_t head = ((size_t)2 * b + MAAI_WS_CTL_WORDS + 127) / 128 * 128;  // row sums + control words
  return (head + (need_bwd ? (size_t)2 * b * d_pad : 0)) * sizeof(float);
}

int maai_ntxent_fwd_is_symmetric(int b, int world, int d_pad) {
  // Single rank: anchors == keys and E_ij = E_ji, so only the tiles on and above the diagonal are
  // computed (half the MMAs and exp2s; measured forward 0.955 -> 0.724 ms at 32768 pairs, d=128).
  // d_pad = 256 below 8192 pairs: the triangular item list cuts a CTA's range into several short
  // segments, each with its own 64 KB Q-tile load (measured 0.17 -> 0.27 ms per step at 4096 pairs).
  // MAAI_FWD_SYM=1 / 0 force it on / off (read per call: the parity tests run both).
  if (world != 1 || env_int("MAAI_DEBUG_FWD_NQ", 2) != 2) return 0;
  const int v = env_int("MAAI_FWD_SYM", -1);
  if (v == 0 || v == 1) return v;
  return (d_pad <= 128 || 2 * b >= 16384) ? 1 : 0;
}

// K1 in all its forms: local slot and / or peer / multicast stores, optional zero fill
static int normalize_impl(const void* h1, const void* h2, int b, int d, int in_dtype, void* z_local,
                          const void* const* peer_z_bases, void* mc_z_base, int world, int rank, float* inv_norm,
                          float* pos_cos, void* zero_fill, size_t zero_bytes, cudaStream_t s,
                          const maai_peer_sync* sync = nullptr) {
  if (!h1 || !h2 || !inv_norm || !pos_cos) return fail(MAAI_E_ARG, "null pointer");
  if (check_sync(sync, world) != MAAI_OK) return MAAI_E_ARG;
  const maai::PeerSync dsync = to_sync(sync, world, rank);
  if (b < 1) return fail(MAAI_E_ARG, "b must be >= 1");
  const int dp = maai_padded_dim(d);
  if (dp < 0) return fail(MAAI_E_SHAPE, "embedding dim must be in [1, 256]");
  if (z_local && !aligned16(z_local)) return fail(MAAI_E_ARG, "z_out must be 16-byte aligned");
  if (zero_bytes && (!zero_fill || !aligned16(zero_fill) || (zero_bytes & 3)))
    return fail(MAAI_E_ARG, "zero_fill must be 16-byte aligned and zero_bytes a multiple of 4");
  const int wpb = 8;
  int grid = (b + wpb - 1) / wpb;
  if (const int cap = sm_count() * 8; cap > 0 && grid > cap) grid = cap;  // one resident wave (8 CTAs of 256 per SM)
  const auto* pb = reinterpret_cast<const unsigned long long*>(peer_z_bases);
  const size_t esz = in_dtype == MAAI_DT_F32 ? 4 : 2;
  const int vec = dp / 32;
  // vector loads need every row chunk aligned: d a multiple of the chunk and aligned base addresses
  const bool vec_ok = (d % vec) == 0 && (reinterpret_cast<uintptr_t>(h1) % (vec * esz)) == 0 &&
                      (reinterpret_cast<uintptr_t>(h2) % (vec * esz)) == 0;
  cudaError_t e = cudaSuccess;
#define MAAI_K1(T, V)                                                                                          \
  e = launch_k(maai::normalize_cast_kernel<T, V>, dim3(grid), dim3(wpb * 32), 0, s, static_cast<const T*>(h1), \
               static_cast<const T*>(h2), b, d, vec_ok, static_cast<__nv_bfloat16*>(z_local), pb,              \
               reinterpret_cast<unsigned long long>(mc_z_base), world, rank, inv_norm, pos_cos,                 \
               static_cast<uint32_t*>(zero_fill), zero_bytes / 4, dsync)
#define MAAI_K1_DP(T)               \
  switch (dp) {                     \
    case 64: MAAI_K1(T, 2); break;  \
    case 128: MAAI_K1(T, 4); break; \
    default: MAAI_K1(T, 8); break;  \
  }
  switch (in_dtype) {
    case MAAI_DT_F32: MAAI_K1_DP(float); break;
    case MAAI_DT_BF16: MAAI_K1_DP(__nv_bfloat16); break;
    case MAAI_DT_F16: MAAI_K1_DP(__half); break;
    default: return fail(MAAI_E_ARG, "in_dtype must be MAAI_DT_F32, MAAI_DT_BF16 or MAAI_DT_F16");
  }
#undef MAAI_K1_DP
#undef MAAI_K1
  MAAI_CUDA(e);
  return MAAI_OK;
}

int maai_ntxent_normalize(const void* h1, const void* h2, int b, int d, int in_dtype, void* z_out,
                          float* inv_norm, float* pos_cos, void* zero_fill, size_t zero_bytes, void* stream) {
  if (!z_out) return fail(MAAI_E_ARG, "null pointer");
  return normalize_impl(h1, h2, b, d, in_dtype, z_out, nullptr, nullptr, 1, 0, inv_norm, pos_cos, zero_fill,
                        zero_bytes, static_cast<cudaStream_t>(stream));
}

static maai::FinalizeArgs make_fin(float* rowsum_l, const float* pos_cos, int b, float inv_tau, float* r_out,
                                   float* loss_out, const void* const* peer_r, int world, int rank, void* mc_r,
                                   const void* const* stage_bases, const maai_peer_sync* sync = nullptr) {
  maai::FinalizeArgs f{};
  // flags only matter when something crosses ranks here: row factors out (peer_r) or staged sums in
  if (sync && (peer_r || mc_r || stage_bases)) f.sync = to_sync(sync, world, rank);
  f.l = rowsum_l;
  f.pos_cos = pos_cos;
  f.b = b;
  f.inv_tau = inv_tau;
  f.r_out = r_out;
  f.loss_out = loss_out;
  f.peer_r = reinterpret_cast<const unsigned long long*>(peer_r);
  f.world = world;
  f.my_rank = rank;
  f.mc_r = static_cast<float*>(mc_r);
  f.stage_tab = reinterpret_cast<const unsigned long long*>(stage_bases);
  return f;
}

// Rows up to which the forward's per-row tail runs inside the tile kernel (last CTA done) instead of
// as a separate 8-CTA cluster launch: one CTA needs ~0.15 us per 1000 rows more than the cluster, a
// launch costs 3-5 us of gap + prologue.  MAAI_TAIL_FINALIZE_ROWS overrides (0 = never).
static int tail_finalize_rows() {
  static const int v = env_int("MAAI_TAIL_FINALIZE_ROWS", 32768);
  return v;
}

static int fwd_impl(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                    const float* pos_cos, float* rowsum_l, float* r_out, float* loss_out,
                    int* pos_rank, int flags, void* stream, const void* const* peer_r = nullptr,
                    void* mc_r = nullptr, const maai_peer_sync* sync = nullptr) {
  if (!z_glob || !pos_cos || !rowsum_l || !loss_out) return fail(MAAI_E_ARG, "null pointer");
  if (check_sync(sync, world) != MAAI_OK) return MAAI_E_ARG;
  int rc = check_common(b, world, rank);
  if (rc != MAAI_OK) return rc;
  if (!(inv_tau > 0.f)) return fail(MAAI_E_ARG, "temperature must be positive");
  if (!aligned16(z_glob)) return fail(MAAI_E_ARG, "z_glob must be 16-byte aligned");
  if (flags & ~MAAI_F_PREZEROED) return fail(MAAI_E_ARG, "unknown flag");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int m_loc = 2 * b, m_glob = 2 * b * world;
  const bool prezeroed = flags & MAAI_F_PREZEROED;  // rowsum_l = head of a step workspace K1 has zero-filled
  if (!prezeroed && (rc = zero_words(rowsum_l, (size_t)m_loc, s)) != MAAI_OK) return rc;
  const char* q_base = static_cast<const char*>(z_glob) + (size_t)rank * m_loc * d_pad * 2;
  const maai::FinalizeArgs fin =
      make_fin(rowsum_l, pos_cos, b, inv_tau, r_out, loss_out, peer_r, world, rank, mc_r, nullptr, sync);
  RankArgs ra;
  if (sync && world > 1) {  // key tiles of the other ranks' slots: wait for their rows in the kernel
    ra.wait_flags = sync->local_flags;
    ra.wait_seq = sync->seq;
    ra.wait_kind = maai::FLAG_Z;
    ra.wait_my_slot = rank;
    ra.wait_nslots = world;
    ra.wait_timeout_s = sync->timeout_s;
  }
  const bool tail = prezeroed && !pos_rank && m_loc <= tail_finalize_rows();
  if (tail) {
    ra.done_ctr = reinterpret_cast<unsigned int*>(rowsum_l + m_loc);  // control word 0 of the workspace
    ra.fin = fin;
  }
  if (pos_rank) {
    // evaluation forward: same kernel + the rank of every view-a anchor's positive among the
    // view-b keys (one instantiation per padded width, the forward's default tile layout)
    if ((rc = zero_words(pos_rank, (size_t)b, s)) != MAAI_OK) return rc;
    ra.pos_cos = pos_cos;
    ra.rank_out = pos_rank;
#define MAAI_LAUNCH_RANK(DD, NQQ)                                                                   \
  rc = launch_tile<DD, false, NQQ, true>(q_base, m_loc, z_glob, m_glob, rank * m_loc, inv_tau, nullptr, \
                                         nullptr, rowsum_l, nullptr, b, b, s, ra)
    switch (d_pad) {
      case 64: MAAI_LAUNCH_RANK(64, 2); break;
      case 128: MAAI_LAUNCH_RANK(128, 2); break;
      case 256: MAAI_LAUNCH_RANK(256, 1); break;
      default: return fail(MAAI_E_SHAPE, "d_pad must be 64, 128 or 256");
    }
#undef MAAI_LAUNCH_RANK
  } else if (maai_ntxent_fwd_is_symmetric(b, world, d_pad)) {
    // a tile above the diagonal adds its row sums to the anchors and its column sums to the keys
    if (d_pad == 64)
      rc = launch_tile<64, false, 2, false, true>(q_base, m_loc, z_glob, m_glob, 0, inv_tau, nullptr, nullptr,
                                                  rowsum_l, nullptr, b, b, s, ra);
    else if (d_pad == 128)
      rc = launch_tile<128, false, 2, false, true>(q_base, m_loc, z_glob, m_glob, 0, inv_tau, nullptr, nullptr,
                                                   rowsum_l, nullptr, b, b, s, ra);
    else if (d_pad == 256)  // one Q tile per row block: row tile rb visits key tiles kt >= rb
      rc = launch_tile<256, false, 1, false, true>(q_base, m_loc, z_glob, m_glob, 0, inv_tau, nullptr, nullptr,
                                                   rowsum_l, nullptr, b, b, s, ra);
    else
      return fail(MAAI_E_SHAPE, "d_pad must be 64, 128 or 256");
  } else {
    rc = dispatch_tile<false>(d_pad, q_base, m_loc, z_glob, m_glob, rank * m_loc, inv_tau, nullptr,
                              nullptr, rowsum_l, nullptr, b, b, s, ra);
  }
  if (rc != MAAI_OK) return rc;
  if (!tail) MAAI_CUDA(launch_k(maai::finalize_loss_kernel, dim3(maai::kFinalizeCluster), dim3(1024), 0, s, fin));
  return MAAI_OK;
}

// Symmetric forward across ranks, tile part: this rank's own block (triangular) + the anchor groups
// of the ranks "ahead" of it on the ring against its own keys (see maai_ntxent.h).
extern "C++" {
// Which tiles rank `rank` computes.  Group 0: its own block, triangular.  For dist = 1 .. world/2 the
// anchors of q = rank + dist (mod world) against the local keys; the pair at distance world/2 (even
// world) is split: the lower rank takes all of q's anchors against its first ceil(T/2) key tiles, the
// higher rank takes the lower rank's anchors from row ceil(T/2)*128 on against all its keys.
struct GroupPlan {
  int ng = 0;
  int qrow0[maai::kMaxGroups];   // first anchor row of the group (global row index)
  int rows[maai::kMaxGroups];
  int nkt[maai::kMaxGroups];
  long long items[maai::kMaxGroups];
  long long total = 0;
};
static GroupPlan plan_groups(int b, int world, int rank, int rb_rows, int nq) {
  GroupPlan gp;
  const int m_loc = 2 * b;
  const int T = (m_loc + 127) / 128, Th = (T + 1) / 2;
  const int nrb = (m_loc + rb_rows - 1) / rb_rows;
  auto add = [&](int qrow0, int rows, int nkt) {
    const int g = gp.ng++;
    gp.qrow0[g] = qrow0;
    gp.rows[g] = rows;
    gp.nkt[g] = nkt;
    gp.items[g] = g == 0 ? (long long)nrb * T - (long long)nq * nrb * (nrb - 1) / 2
                         : (long long)((rows + rb_rows - 1) / rb_rows) * nkt;
    gp.total += gp.items[g];
  };
  add(rank * m_loc, m_loc, T);
  for (int dist = 1; 2 * dist <= world; ++dist) {
    const int q = (rank + dist) % world;  // owner of the anchors
    if (2 * dist < world) add(q * m_loc, m_loc, T);
    else if (rank < q) add(q * m_loc, m_loc, Th);
    else if (m_loc - Th * 128 > 0) add(q * m_loc + Th * 128, m_loc - Th * 128, T);
  }
  return gp;
}

template <int D, int NQ>
static int launch_tile_groups(const void* z_glob, int b, int world, int rank, float inv_tau, float* rowsum_l,
                              float* stage, cudaStream_t s, const maai_peer_sync* sync,
                              const void* const* peer_rowsum_host = nullptr, const maai::FinalizeArgs* fin = nullptr) {
  using C = maai::TileCfg<D, false, NQ>;
  static std::atomic<unsigned long long> attr_done{0};
  cudaError_t attr_err =
      ensure_smem_attr(maai::ntxent_tile_kernel<D, false, NQ, false, true, true>, C::SMEM_BYTES, attr_done);
  if (attr_err != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(smem)", attr_err);
  const int m_loc = 2 * b, m_glob = 2 * b * world;
  const char* k_base = static_cast<const char*>(z_glob) + (size_t)rank * m_loc * D * 2;
  CUtensorMap tq, tk;
  int rc;
  if ((rc = make_rows_tmap(&tq, z_glob, m_glob, D)) != MAAI_OK) return rc;
  if ((rc = make_rows_tmap(&tk, k_base, m_loc, D)) != MAAI_OK) return rc;
  maai::TileParams p{};
  p.m_loc = m_loc;
  p.m_glob = m_loc;  // key space = this rank's slot
  p.row_global_base = 0;
  p.pos_split = b;
  p.pos_delta = b;
  p.nrb = (m_loc + C::RB_ROWS - 1) / C::RB_ROWS;
  p.nkt = (m_loc + C::KT - 1) / C::KT;
  p.c1 = inv_tau * 1.4426950408889634f;
  p.l_out = rowsum_l;
  p.pv_lbo = C::CHUNK_BYTES;
  p.pv_sbo = 1024;
  const GroupPlan gp = plan_groups(b, world, rank, C::RB_ROWS, NQ);
  for (int g = 0; g < gp.ng; ++g) {
    p.g_qrow0[g] = gp.qrow0[g];
    p.g_rows[g] = gp.rows[g];
    p.g_nkt[g] = gp.nkt[g];
    p.g_items[g] = gp.items[g];
    // group 0: this rank's own row sums; others: the anchors' owner -- its slot of this rank's staging vectors
    // ((world, 2b), indexed by global row), or, in direct mode, the owner's own row sums over NVLink
    if (g == 0) {
      p.g_out[g] = rowsum_l;
    } else if (peer_rowsum_host) {
      const int q = gp.qrow0[g] / m_loc;
      p.g_out[g] = static_cast<float*>(const_cast<void*>(peer_rowsum_host[q])) + (gp.qrow0[g] - q * m_loc);
    } else {
      p.g_out[g] = stage + gp.qrow0[g];
    }
  }
  if (fin) {  // direct mode: the per-row tail runs in this kernel (TileParams::done_ctr doubles as the marker)
    p.done_ctr = reinterpret_cast<unsigned int*>(rowsum_l + m_loc);
    p.fin = *fin;
  }
  p.ngroups = gp.ng;
  p.total_items = gp.total;
  if (sync) {  // anchor tiles of the groups come from other ranks' slots; the staged sums are announced (FLAG_L)
    p.wait_flags = sync->local_flags;
    p.wait_seq = sync->seq;
    p.wait_kind = maai::FLAG_Z;
    p.wait_my_slot = rank;
    p.wait_nslots = world;
    p.wait_timeout_s = sync->timeout_s;
    p.grp_sync = to_sync(sync, world, rank);
  }
  const long long total = gp.total;
  int sms = sm_count();
  if (sms <= 0) return fail(MAAI_E_CUDA, "no CUDA device");
  const int grid = (int)(total < sms ? total : sms);
  MAAI_CUDA(launch_k(maai::ntxent_tile_kernel<D, false, NQ, false, true, true>, dim3(grid), dim3(C::NTHREADS),
                     C::SMEM_BYTES, s, tq, tk, p));
  return MAAI_OK;
}
}  // extern "C++"

int maai_ntxent_fwd(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                    const float* pos_cos, float* rowsum_l, float* r_out, float* loss_out, int flags,
                    const maai_peer_sync* sync, void* stream) {
  return fwd_impl(z_glob, b, world, rank, d_pad, inv_tau, pos_cos, rowsum_l, r_out, loss_out, nullptr, flags,
                  stream, nullptr, nullptr, sync);
}

int maai_ntxent_normalize_peer(const void* h1, const void* h2, int b, int d, int in_dtype,
                               const void* const* peer_z_bases, void* mc_z_base, int world, int rank,
                               float* inv_norm, float* pos_cos, void* zero_fill, size_t zero_bytes,
                               const maai_peer_sync* sync, void* stream) {
  if (!peer_z_bases) return fail(MAAI_E_ARG, "null pointer");
  int rc = check_common(b, world, rank);
  if (rc != MAAI_OK) return rc;
  return normalize_impl(h1, h2, b, d, in_dtype, nullptr, peer_z_bases, mc_z_base, world, rank, inv_norm, pos_cos,
                        zero_fill, zero_bytes, static_cast<cudaStream_t>(stream), sync);
}

int maai_ntxent_normalize_chain(const void* h2, int b, int d, int in_dtype, const void* z_prev,
                                const float* inv_norm_prev, void* z_new, const void* const* peer_z_bases,
                                void* mc_z_base, int world, int rank, float* inv_norm, float* pos_cos,
                                void* zero_fill, size_t zero_bytes, const maai_peer_sync* sync, void* stream) {
  if (check_sync(sync, world) != MAAI_OK) return MAAI_E_ARG;
  const maai::PeerSync dsync = to_sync(sync, world, rank);
  if (!h2 || !z_prev || !inv_norm_prev || !z_new || !inv_norm || !pos_cos) return fail(MAAI_E_ARG, "null pointer");
  int rc = check_common(b, world, rank);
  if (rc != MAAI_OK) return rc;
  const int dp = maai_padded_dim(d);
  if (dp < 0) return fail(MAAI_E_SHAPE, "embedding dim must be in [1, 256]");
  if (!aligned16(z_prev) || !aligned16(z_new)) return fail(MAAI_E_ARG, "z_prev and z_new must be 16-byte aligned");
  if (z_prev == z_new) return fail(MAAI_E_ARG, "z_prev and z_new must be different buffers");
  if (zero_bytes && (!zero_fill || !aligned16(zero_fill) || (zero_bytes & 3)))
    return fail(MAAI_E_ARG, "zero_fill must be 16-byte aligned and zero_bytes a multiple of 4");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int wpb = 8;
  int grid = (b + wpb - 1) / wpb;
  if (const int cap = sm_count() * 8; cap > 0 && grid > cap) grid = cap;
  const size_t esz = in_dtype == MAAI_DT_F32 ? 4 : 2;
  const int vec = dp / 32;
  const bool vec_ok = (d % vec) == 0 && (reinterpret_cast<uintptr_t>(h2) % (vec * esz)) == 0;
  cudaError_t e = cudaSuccess;
#define MAAI_K1C(T, V)                                                                                        \
  e = launch_k(maai::normalize_chain_kernel<T, V>, dim3(grid), dim3(wpb * 32), 0, s, static_cast<const T*>(h2), \
               b, d, vec_ok, static_cast<const __nv_bfloat16*>(z_prev), inv_norm_prev,                          \
               static_cast<__nv_bfloat16*>(z_new), reinterpret_cast<const unsigned long long*>(peer_z_bases),   \
               reinterpret_cast<unsigned long long>(mc_z_base), world, rank, inv_norm, pos_cos,                 \
               static_cast<uint32_t*>(zero_fill), zero_bytes / 4, dsync)
#define MAAI_K1C_DP(T)               \
  switch (dp) {                      \
    case 64: MAAI_K1C(T, 2); break;  \
    case 128: MAAI_K1C(T, 4); break; \
    default: MAAI_K1C(T, 8); break;  \
  }
  switch (in_dtype) {
    case MAAI_DT_F32: MAAI_K1C_DP(float); break;
    case MAAI_DT_BF16: MAAI_K1C_DP(__nv_bfloat16); break;
    case MAAI_DT_F16: MAAI_K1C_DP(__half); break;
    default: return fail(MAAI_E_ARG, "in_dtype must be MAAI_DT_F32, MAAI_DT_BF16 or MAAI_DT_F16");
  }
#undef MAAI_K1C_DP
#undef MAAI_K1C
  MAAI_CUDA(e);
  return MAAI_OK;
}

int maai_ntxent_fwd_peer(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                         const float* pos_cos, float* rowsum_l, const void* const* peer_r_bases,
                         void* mc_r_base, float* loss_out, int flags, const maai_peer_sync* sync, void* stream) {
  if (!peer_r_bases) return fail(MAAI_E_ARG, "null pointer");
  return fwd_impl(z_glob, b, world, rank, d_pad, inv_tau, pos_cos, rowsum_l, nullptr, loss_out, nullptr, flags,
                  stream, peer_r_bases, mc_r_base, sync);
}

int maai_ntxent_fwd_sym_tiles(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                              float* rowsum_l, float* stage, int flags, const maai_peer_sync* sync, void* stream) {
  if (flags & ~MAAI_F_PREZEROED) return fail(MAAI_E_ARG, "unknown flag");
  if (check_sync(sync, world) != MAAI_OK) return MAAI_E_ARG;
  if (!z_glob || !rowsum_l || !stage) return fail(MAAI_E_ARG, "null pointer");
  int rc = check_common(b, world, rank);
  if (rc != MAAI_OK) return rc;
  if (world > 2 * (maai::kMaxGroups - 1)) return fail(MAAI_E_SHAPE, "symmetric forward: world must be <= 16");
  if (!(inv_tau > 0.f)) return fail(MAAI_E_ARG, "temperature must be positive");
  if (!aligned16(z_glob)) return fail(MAAI_E_ARG, "z_glob must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!(flags & MAAI_F_PREZEROED) && (rc = zero_words(rowsum_l, (size_t)2 * b, s)) != MAAI_OK) return rc;
  if ((rc = zero_words(stage, (size_t)2 * b * world, s)) != MAAI_OK) return rc;
  switch (d_pad) {
    case 64: return launch_tile_groups<64, 2>(z_glob, b, world, rank, inv_tau, rowsum_l, stage, s, sync);
    case 128: return launch_tile_groups<128, 2>(z_glob, b, world, rank, inv_tau, rowsum_l, stage, s, sync);
    case 256: return launch_tile_groups<256, 1>(z_glob, b, world, rank, inv_tau, rowsum_l, stage, s, sync);
    default: return fail(MAAI_E_SHAPE, "d_pad must be 64, 128 or 256");
  }
}

int maai_ntxent_fwd_sym_direct(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                               const float* pos_cos, float* rowsum_l, const void* const* peer_rowsum_host,
                               float* r_out, const void* const* peer_r_bases, void* mc_r_base, float* loss_out,
                               const maai_peer_sync* sync, void* stream) {
  if (!z_glob || !pos_cos || !rowsum_l || !peer_rowsum_host || !loss_out || !sync) return fail(MAAI_E_ARG, "null pointer");
  int rc = check_common(b, world, rank);
  if (rc != MAAI_OK) return rc;
  if (world > 2 * (maai::kMaxGroups - 1)) return fail(MAAI_E_SHAPE, "symmetric forward: world must be <= 16");
  if (!(inv_tau > 0.f)) return fail(MAAI_E_ARG, "temperature must be positive");
  if (!aligned16(z_glob)) return fail(MAAI_E_ARG, "z_glob must be 16-byte aligned");
  if (check_sync(sync, world) != MAAI_OK) return MAAI_E_ARG;
  for (int q = 0; q < world; ++q)
    if (!peer_rowsum_host[q]) return fail(MAAI_E_ARG, "peer_rowsum_host: null entry");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const maai::FinalizeArgs fin =
      make_fin(rowsum_l, pos_cos, b, inv_tau, r_out, loss_out, peer_r_bases, world, rank, mc_r_base, nullptr, sync);
  switch (d_pad) {
    case 64: return launch_tile_groups<64, 2>(z_glob, b, world, rank, inv_tau, rowsum_l, nullptr, s, sync, peer_rowsum_host, &fin);
    case 128: return launch_tile_groups<128, 2>(z_glob, b, world, rank, inv_tau, rowsum_l, nullptr, s, sync, peer_rowsum_host, &fin);
    case 256: return launch_tile_groups<256, 1>(z_glob, b, world, rank, inv_tau, rowsum_l, nullptr, s, sync, peer_rowsum_host, &fin);
    default: return fail(MAAI_E_SHAPE, "d_pad must be 64, 128 or 256");
  }
}

int maai_debug_group_plan(int b, int world, int rank, int d_pad, int* ngroups, int* qrow0, int* rows, int* nkt,
                          long long* items) {
  if (!ngroups || !qrow0 || !rows || !nkt || !items) return fail(MAAI_E_ARG, "null pointer");
  int rc = check_common(b, world, rank);
  if (rc != MAAI_OK) return rc;
  if (world > 2 * (maai::kMaxGroups - 1)) return fail(MAAI_E_SHAPE, "symmetric forward: world must be <= 16");
  if (d_pad != 64 && d_pad != 128 && d_pad != 256) return fail(MAAI_E_SHAPE, "d_pad must be 64, 128 or 256");
  const int nq = d_pad == 256 ? 1 : 2;
  const GroupPlan gp = plan_groups(b, world, rank, 128 * nq, nq);
  *ngroups = gp.ng;
  for (int g = 0; g < gp.ng; ++g) {
    qrow0[g] = gp.qrow0[g];
    rows[g] = gp.rows[g];
    nkt[g] = gp.nkt[g];
    items[g] = gp.items[g];
  }
  return MAAI_OK;
}

int maai_debug_tri_locate(long long idx, int n_key_tiles, int n_row_blocks, int nq, int* rb, int* off, int* cnt) {
  if (!rb || !off || !cnt || (nq != 1 && nq != 2)) return fail(MAAI_E_ARG, "bad argument");
  if (nq == 1) maai::tri_locate<1>(idx, n_key_tiles, n_row_blocks, *rb, *off, *cnt);
  else maai::tri_locate<2>(idx, n_key_tiles, n_row_blocks, *rb, *off, *cnt);
  return MAAI_OK;
}

int maai_ntxent_fwd_sym_finalize(float* rowsum_l, const void* const* stage_bases, int b, int world, int rank,
                                 float inv_tau, const float* pos_cos, float* r_out,
                                 const void* const* peer_r_bases, void* mc_r_base, float* loss_out,
                                 const maai_peer_sync* sync, void* stream) {
  if (!rowsum_l || !stage_bases || !pos_cos || !loss_out) return fail(MAAI_E_ARG, "null pointer");
  if (check_sync(sync, world) != MAAI_OK) return MAAI_E_ARG;
  int rc = check_common(b, world, rank);
  if (rc != MAAI_OK) return rc;
  if (!(inv_tau > 0.f)) return fail(MAAI_E_ARG, "temperature must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MAAI_CUDA(launch_k(maai::finalize_loss_kernel, dim3(maai::kFinalizeCluster), dim3(1024), 0, s,
                     make_fin(rowsum_l, pos_cos, b, inv_tau, r_out, loss_out, peer_r_bases, world, rank, mc_r_base,
                              stage_bases, sync)));
  return MAAI_OK;
}

int maai_ntxent_fwd_eval(const void* z_glob, int b, int world, int rank, int d_pad, float inv_tau,
                         const float* pos_cos, float* rowsum_l, float* loss_out, int* pos_rank,
                         void* stream) {
  if (!pos_rank) return fail(MAAI_E_ARG, "null pointer");
  return fwd_impl(z_glob, b, world, rank, d_pad, inv_tau, pos_cos, rowsum_l, nullptr, loss_out, pos_rank, 0,
                  stream);
}

// tile pass of the backward over this rank's anchors: zeroes the accumulator rows, then K3
static int bwd_tiles_impl(const void* z_glob, const float* r_row, const float* r_col, int b, int world, int rank,
                          int d_pad, float inv_tau, int need_mask, float* dz_acc, cudaStream_t s,
                          bool prezeroed = false, const maai_peer_sync* sync = nullptr) {
  const int m_loc = 2 * b, m_glob = 2 * b * world;
  // anchor rows that need a gradient: both views, or one contiguous view
  const int row_begin = (need_mask == 2) ? b : 0;
  const int rows = (need_mask == 3) ? m_loc : b;
  float* acc = dz_acc + (size_t)row_begin * d_pad;
  int rc;
  if (!prezeroed && (rc = zero_words(acc, (size_t)rows * d_pad, s)) != MAAI_OK) return rc;
  const char* q_base =
      static_cast<const char*>(z_glob) + ((size_t)rank * m_loc + row_begin) * d_pad * 2;
  RankArgs ra;
  if (sync && world > 1) {  // r_col of the other ranks' slots: wait for their row factors in the kernel
    ra.wait_flags = sync->local_flags;
    ra.wait_seq = sync->seq;
    ra.wait_kind = maai::FLAG_R;
    ra.wait_my_slot = rank;
    ra.wait_nslots = world;
    ra.wait_timeout_s = sync->timeout_s;
  }
  return dispatch_tile<true>(d_pad, q_base, rows, z_glob, m_glob, rank * m_loc + row_begin, inv_tau,
                             r_row + row_begin, r_col, nullptr, acc, b - row_begin, b, s, ra);
}

static int bwd_dh_impl(const float* dz_acc, const float* dz_extra, const float* rowsum_l, const float* pos_cos,
                       const void* h1, const void* h2, int in_dtype, const float* inv_norm,
                       const float* grad_loss, int b, int d, int d_pad, float inv_tau, int key_grad,
                       int need_mask, void* dh1, void* dh2, cudaStream_t s) {
  const int rows = (need_mask == 3) ? 2 * b : b;
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  const size_t esz = in_dtype == MAAI_DT_F32 ? 4 : 2;
  const int vec = d_pad / 32;
  auto al = [&](const void* p) { return !p || (reinterpret_cast<uintptr_t>(p) % (vec * esz)) == 0; };
  const bool vec_ok = (d % vec) == 0 && al(h1) && al(h2) && al(dh1) && al(dh2);
  cudaError_t e = cudaSuccess;
#define MAAI_K4(T, V)                                                                                       \
  e = launch_k(maai::dh_kernel<T, V>, dim3(grid), dim3(wpb * 32), 0, s, dz_acc, dz_extra,                   \
               static_cast<const T*>(h1), static_cast<const T*>(h2), inv_norm, grad_loss, rowsum_l, pos_cos, \
               b, d, vec_ok, inv_tau, key_grad, need_mask, static_cast<T*>(dh1), static_cast<T*>(dh2))
#define MAAI_K4_DP(T)               \
  switch (d_pad) {                  \
    case 64: MAAI_K4(T, 2); break;  \
    case 128: MAAI_K4(T, 4); break; \
    default: MAAI_K4(T, 8); break;  \
  }
  switch (in_dtype) {
    case MAAI_DT_F32: MAAI_K4_DP(float); break;
    case MAAI_DT_BF16: MAAI_K4_DP(__nv_bfloat16); break;
    case MAAI_DT_F16: MAAI_K4_DP(__half); break;
    default: return fail(MAAI_E_ARG, "in_dtype must be MAAI_DT_F32, MAAI_DT_BF16 or MAAI_DT_F16");
  }
#undef MAAI_K4_DP
#undef MAAI_K4
  MAAI_CUDA(e);
  return MAAI_OK;
}

static int bwd_check(int b, int world, int rank, int d, int d_pad, int need_mask) {
  int rc = check_common(b, world, rank);
  if (rc != MAAI_OK) return rc;
  if (need_mask < 0 || need_mask > 3) return fail(MAAI_E_ARG, "need_mask must be in [0, 3]");
  if (d > 0 && maai_padded_dim(d) != d_pad) return fail(MAAI_E_SHAPE, "d_pad does not match maai_padded_dim(d)");
  if (d_pad != 64 && d_pad != 128 && d_pad != 256) return fail(MAAI_E_SHAPE, "d_pad must be 64, 128 or 256");
  return MAAI_OK;
}

int maai_ntxent_bwd(const void* z_glob, const float* r_row, const float* r_col, int key_grad,
                    const float* rowsum_l, const float* pos_cos, const void* h1, const void* h2, int in_dtype,
                    const float* inv_norm, const float* grad_loss, int b, int world, int rank, int d, int d_pad,
                    float inv_tau, int need_mask, void* dh1, void* dh2, float* dz_acc, int flags,
                    const maai_peer_sync* sync, void* stream) {
  if (flags & ~MAAI_F_PREZEROED) return fail(MAAI_E_ARG, "unknown flag");
  if (check_sync(sync, world) != MAAI_OK) return MAAI_E_ARG;
  if (!z_glob || !r_row || !r_col || !rowsum_l || !pos_cos || !h1 || !h2 || !inv_norm || !grad_loss || !dz_acc)
    return fail(MAAI_E_ARG, "null pointer");
  int rc = bwd_check(b, world, rank, d, d_pad, need_mask);
  if (rc != MAAI_OK) return rc;
  if (need_mask == 0) return MAAI_OK;
  if (((need_mask & 1) && !dh1) || ((need_mask & 2) && !dh2)) return fail(MAAI_E_ARG, "null dh");
  if (!aligned16(z_glob) || !aligned16(r_col) || !aligned16(dz_acc))
    return fail(MAAI_E_ARG, "z_glob, r_col and dz_acc must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((rc = bwd_tiles_impl(z_glob, r_row, r_col, b, world, rank, d_pad, inv_tau, need_mask, dz_acc, s,
                           flags & MAAI_F_PREZEROED, sync)) != MAAI_OK)
    return rc;
  return bwd_dh_impl(dz_acc, nullptr, rowsum_l, pos_cos, h1, h2, in_dtype, inv_norm, grad_loss, b, d, d_pad,
                     inv_tau, key_grad, need_mask, dh1, dh2, s);
}

int maai_ntxent_bwd_tiles(const void* z_glob, const float* r_row, const float* r_col, int b, int world, int rank,
                          int d_pad, float inv_tau, int need_mask, float* dz_acc, void* stream) {
  if (!z_glob || !r_row || !r_col || !dz_acc) return fail(MAAI_E_ARG, "null pointer");
  int rc = bwd_check(b, world, rank, 0, d_pad, need_mask);
  if (rc != MAAI_OK) return rc;
  if (need_mask == 0) return MAAI_OK;
  if (!aligned16(z_glob) || !aligned16(r_col) || !aligned16(dz_acc))
    return fail(MAAI_E_ARG, "z_glob, r_col and dz_acc must be 16-byte aligned");
  return bwd_tiles_impl(z_glob, r_row, r_col, b, world, rank, d_pad, inv_tau, need_mask, dz_acc,
                        static_cast<cudaStream_t>(stream));
}

int maai_ntxent_bwd_keyside(const void* z_glob, const float* r_col_loc, int b, int world, int rank, int d_pad,
                            float inv_tau, float* dz_keys, void* stream) {
  if (!z_glob || !r_col_loc || !dz_keys) return fail(MAAI_E_ARG, "null pointer");
  int rc = bwd_check(b, world, rank, 0, d_pad, 3);
  if (rc != MAAI_OK) return rc;
  if (!aligned16(z_glob) || !aligned16(r_col_loc) || !aligned16(dz_keys))
    return fail(MAAI_E_ARG, "z_glob, r_col_loc and dz_keys must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int m_loc = 2 * b, m_glob = 2 * b * world;
  if ((rc = zero_words(dz_keys, (size_t)m_glob * d_pad, s)) != MAAI_OK) return rc;
  // anchors: every row of the gathered buffer; keys: this rank's slot.  In key space (0 = first key
  // of the slot) anchor row 0 sits at -rank*2b; the view of an anchor follows from its slot offset.
  const char* k_base = static_cast<const char*>(z_glob) + (size_t)rank * m_loc * d_pad * 2;
  RankArgs ra;
  ra.pos_period = m_loc;
  ra.pos_phase = 0;
  return dispatch_tile<true>(d_pad, z_glob, m_glob, k_base, m_loc, -rank * m_loc, inv_tau, nullptr, r_col_loc,
                             nullptr, dz_keys, 0, b, s, ra);
}

int maai_ntxent_bwd_dh(const float* dz_acc, const float* dz_extra, const float* rowsum_l, const float* pos_cos,
                       const void* h1, const void* h2, int in_dtype, const float* inv_norm,
                       const float* grad_loss, int b, int d, int d_pad, float inv_tau, int key_grad, int need_mask,
                       void* dh1, void* dh2, void* stream) {
  if (!dz_acc || !rowsum_l || !pos_cos || !h1 || !h2 || !inv_norm || !grad_loss)
    return fail(MAAI_E_ARG, "null pointer");
  int rc = bwd_check(b, 1, 0, d, d_pad, need_mask);
  if (rc != MAAI_OK) return rc;
  if (need_mask == 0) return MAAI_OK;
  if (((need_mask & 1) && !dh1) || ((need_mask & 2) && !dh2)) return fail(MAAI_E_ARG, "null dh");
  return bwd_dh_impl(dz_acc, dz_extra, rowsum_l, pos_cos, h1, h2, in_dtype, inv_norm, grad_loss, b, d, d_pad,
                     inv_tau, key_grad, need_mask, dh1, dh2, static_cast<cudaStream_t>(stream));
}

#if MAAI_PROF
// tools/phase_prof.py: per-CTA, per-warp phase cycle counters of the last tile-kernel launch
int maai_debug_prof_read(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, maai::g_prof, sizeof(long long) * n);
}
#endif

}  // extern "C"
