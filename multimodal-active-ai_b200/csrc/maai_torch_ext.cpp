// C++ autograd binding of the NT-Xent step, single-rank and multi-rank with peer gathers (plumbing, not the product: every kernel it
// enqueues lives in libmaai_ntxent.so behind include/maai_ntxent.h).
//
// Why it exists: at the batch sizes the reference trains with (256 - 4096 pairs per GPU,
// Contrastive_Learning.py:92) the step's kernels take 35 - 85 us while the Python autograd.Function +
// ctypes path spends ~170 us of host time per step (profiles/r2_tuning_log.md).  The same sequence of C-ABI
// calls issued from a torch::autograd::Function costs a fraction of that.  multimodal-active-ai_b200/
// Objective.py routes world_size == 1 training calls here when this module has been built
// (build.py::build_torch_ext); everything else (multi-rank gathers, evaluation outputs, the reduce-scatter
// dataflow) stays in Python.  Semantics are identical to _NTXentFunction in Objective.py.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <mutex>
#include <utility>
#include <vector>

#include "../../include/maai_ntxent.h"

namespace {

int dtype_code(const torch::Tensor& t) {
  switch (t.scalar_type()) {
    case torch::kFloat32: return MAAI_DT_F32;
    case torch::kBFloat16: return MAAI_DT_BF16;
    case torch::kFloat16: return MAAI_DT_F16;
    default: TORCH_CHECK_TYPE(false, "hidden1/hidden2 must be float32, bfloat16 or float16");
  }
  return -1;
}

void check_rc(int rc, const char* what) {
  if (rc == MAAI_OK) return;
  const std::string msg = std::string(what) + ": " + maai_last_error();
  TORCH_CHECK_VALUE(rc != MAAI_E_ARG && rc != MAAI_E_SHAPE, msg);
  TORCH_CHECK(false, msg);
}

// Optional CUDA-event brackets around the three C-ABI calls of a step (bench.py's live per-kernel timing; the
// Python path has the same brackets in Objective.py::_Profiler).  Off by default: two cudaEventRecord per call.
namespace spans {
enum Kind { NORMALIZE, FWD, BWD, KINDS };
std::mutex mu;  // forward runs on the caller's thread, backward on the autograd engine's
bool on = false;
std::vector<cudaEvent_t> pool;
size_t used = 0;
std::vector<std::pair<size_t, size_t>> rec[KINDS];

long mark(void* stream) {
  std::lock_guard<std::mutex> g(mu);
  if (!on) return -1;
  if (used == pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return -1;
    pool.push_back(e);
  }
  cudaEventRecord(pool[used], static_cast<cudaStream_t>(stream));
  return (long)used++;
}
void add(Kind k, long a, long b) {
  if (a < 0 || b < 0) return;
  std::lock_guard<std::mutex> g(mu);
  rec[k].emplace_back((size_t)a, (size_t)b);
}
void enable(bool v) {
  std::lock_guard<std::mutex> g(mu);
  on = v;
}
void reset() {
  std::lock_guard<std::mutex> g(mu);
  used = 0;
  for (auto& r : rec) r.clear();
}
std::vector<std::vector<double>> read() {  // ms per recorded call: [normalize, fwd, bwd]; waits for the events
  std::lock_guard<std::mutex> g(mu);
  std::vector<std::vector<double>> out(KINDS);
  for (int k = 0; k < KINDS; ++k)
    for (const auto& ab : rec[k]) {
      float ms = 0.f;
      cudaEventSynchronize(pool[ab.second]);
      if (cudaEventElapsedTime(&ms, pool[ab.first], pool[ab.second]) == cudaSuccess) out[k].push_back(ms);
    }
  return out;
}
}  // namespace spans

struct Layout {  // one fp32 allocation: [step workspace | r_col] (zero-filled by K1) | inv_norm | pos_cos
  int64_t ws_words, head, r_len, off_r, off_inv, off_cos, total;
};
Layout make_layout(int b, int dp, bool need_bwd) {
  Layout l;
  l.ws_words = (int64_t)(maai_ntxent_workspace_bytes(b, dp, need_bwd ? 1 : 0) / 4);
  l.head = ((int64_t)2 * b + MAAI_WS_CTL_WORDS + 127) / 128 * 128;
  l.r_len = need_bwd ? (int64_t)maai_ntxent_r_len(b, 1) : 0;
  l.off_r = l.ws_words;
  l.off_inv = (l.off_r + l.r_len + 3) / 4 * 4;
  l.off_cos = l.off_inv + 2 * b;
  l.total = l.off_cos + b;
  return l;
}

class NTXentFn : public torch::autograd::Function<NTXentFn> {
 public:
  // need_bwd is decided by the caller (ntxent_loss below): grad mode is switched off inside forward()
  static torch::Tensor forward(torch::autograd::AutogradContext* ctx, torch::Tensor hidden1, torch::Tensor hidden2,
                               double temperature, bool need_bwd) {
    const auto h1 = hidden1.contiguous();
    const auto h2 = hidden2.contiguous();
    const int b = (int)h1.size(0), d = (int)h1.size(1);
    const int dp = maai_padded_dim(d);
    TORCH_CHECK_VALUE(dp > 0, "embedding dim ", d, " unsupported: the sm_100a tile kernels take 1 <= d <= 256");
    const int dt = dtype_code(h1);
    const float inv_tau = (float)(1.0 / temperature);
    c10::cuda::CUDAGuard guard(h1.device());
    void* st = at::cuda::getCurrentCUDAStream(h1.device().index()).stream();
    const Layout l = make_layout(b, dp, need_bwd);
    auto z = torch::empty({(int64_t)2 * b, (int64_t)dp}, h1.options().dtype(torch::kBFloat16));
    auto buf = torch::empty({l.total}, h1.options().dtype(torch::kFloat32));
    // the loss is its own tensor: an output that is a view of a buffer created in here gets no grad_fn
    auto loss = torch::empty({}, h1.options().dtype(torch::kFloat32));
    float* base = buf.data_ptr<float>();
    const long e0 = spans::mark(st);
    check_rc(maai_ntxent_normalize(h1.data_ptr(), h2.data_ptr(), b, d, dt, z.data_ptr(), base + l.off_inv,
                                   base + l.off_cos, base, (size_t)(l.ws_words + l.r_len) * 4, st),
             "maai_ntxent_normalize");
    const long e1 = spans::mark(st);
    check_rc(maai_ntxent_fwd(z.data_ptr(), b, 1, 0, dp, inv_tau, base + l.off_cos, base,
                             need_bwd ? base + l.off_r : nullptr, loss.data_ptr<float>(), MAAI_F_PREZEROED, nullptr, st),
             "maai_ntxent_fwd");
    if (e0 >= 0) {
      spans::add(spans::NORMALIZE, e0, e1);
      spans::add(spans::FWD, e1, spans::mark(st));
    }
    if (need_bwd) {
      ctx->save_for_backward({h1, h2});
      ctx->saved_data["z"] = z;
      ctx->saved_data["buf"] = buf;
      ctx->saved_data["b"] = (int64_t)b;
      ctx->saved_data["d"] = (int64_t)d;
      ctx->saved_data["inv_tau"] = (double)inv_tau;
      ctx->saved_data["clean"] = true;
    }
    return loss;
  }

  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::variable_list grad_out) {
    const auto saved = ctx->get_saved_variables();
    const auto& h1 = saved[0];
    const auto& h2 = saved[1];
    const auto z = ctx->saved_data["z"].toTensor();
    const auto buf = ctx->saved_data["buf"].toTensor();
    const int b = (int)ctx->saved_data["b"].toInt(), d = (int)ctx->saved_data["d"].toInt();
    const float inv_tau = (float)ctx->saved_data["inv_tau"].toDouble();
    const bool clean = ctx->saved_data["clean"].toBool();
    ctx->saved_data["clean"] = false;  // a second backward (retain_graph) zeroes the accumulator itself
    const int dp = maai_padded_dim(d);
    const int need = (ctx->needs_input_grad(0) ? 1 : 0) | (ctx->needs_input_grad(1) ? 2 : 0);
    c10::cuda::CUDAGuard guard(h1.device());
    void* st = at::cuda::getCurrentCUDAStream(h1.device().index()).stream();
    const Layout l = make_layout(b, dp, true);
    float* base = buf.data_ptr<float>();
    auto g = grad_out[0].to(h1.device(), torch::kFloat32).contiguous();
    torch::Tensor dh1, dh2;
    if (need & 1) dh1 = torch::empty_like(h1);
    if (need & 2) dh2 = torch::empty_like(h2);
    const long e0 = spans::mark(st);
    check_rc(maai_ntxent_bwd(z.data_ptr(), base + l.off_r, base + l.off_r, 1, base, base + l.off_cos, h1.data_ptr(),
                             h2.data_ptr(), dtype_code(h1), base + l.off_inv, g.data_ptr<float>(), b, 1, 0, d, dp,
                             inv_tau, need, (need & 1) ? dh1.data_ptr() : nullptr, (need & 2) ? dh2.data_ptr() : nullptr,
                             base + l.head, clean ? MAAI_F_PREZEROED : 0, nullptr, st),
             "maai_ntxent_bwd");
    if (e0 >= 0) spans::add(spans::BWD, e0, spans::mark(st));
    return {dh1, dh2, torch::Tensor(), torch::Tensor()};
  }
};

// ---------------------------------------------------------------------------------------------------------
// Multi-rank step with the fused peer gathers ordered by in-kernel flags (full gradient, no cross-rank symmetric
// forward, no chained views): the same three C-ABI calls as Objective.py::_forward_peer / backward.  The buffer
// set, its peer tables and the flag block are chosen and owned by Python (PeerWorkspace); `a` carries their
// addresses.  `state` is the host int64 array of PeerWorkspace's SetReusePolicy ([0] = forwards issued,
// [1 .. nbuf] = backward pending per set, [1 + nbuf .. ] = stamp per set): the backward marks its set as issued
// there, exactly like SetReusePolicy.backward_issued.
enum PeerArg : int { A_RANK, A_WORLD, A_Z_ALL, A_Z_TAB, A_MC_Z, A_R_COL, A_R_TAB, A_MC_R, A_F_TAB, A_FLAGS, A_COUNTER,
                     A_SEQ, A_TIMEOUT, A_SET, A_NBUF, A_STATE, A_COUNT };

maai_peer_sync make_sync(const std::vector<int64_t>& a) {
  maai_peer_sync s;
  s.peer_flag_bases = reinterpret_cast<const void* const*>(a[A_F_TAB]);
  s.local_flags = reinterpret_cast<unsigned int*>(a[A_FLAGS]);
  s.counter = reinterpret_cast<unsigned int*>(a[A_COUNTER]);
  s.seq = (unsigned int)a[A_SEQ];
  s.timeout_s = (unsigned int)a[A_TIMEOUT];
  return s;
}

class NTXentPeerFn : public torch::autograd::Function<NTXentPeerFn> {
 public:
  static torch::Tensor forward(torch::autograd::AutogradContext* ctx, torch::Tensor hidden1, torch::Tensor hidden2,
                               double temperature, bool need_bwd, std::vector<int64_t> a) {
    TORCH_CHECK((int)a.size() == A_COUNT, "ntxent_loss_peer: bad argument vector");
    const auto h1 = hidden1.contiguous();
    const auto h2 = hidden2.contiguous();
    const int b = (int)h1.size(0), d = (int)h1.size(1);
    const int rank = (int)a[A_RANK], world = (int)a[A_WORLD];
    const int dp = maai_padded_dim(d);
    TORCH_CHECK_VALUE(dp > 0, "embedding dim ", d, " unsupported: the sm_100a tile kernels take 1 <= d <= 256");
    const int dt = dtype_code(h1);
    const float inv_tau = (float)(1.0 / temperature);
    c10::cuda::CUDAGuard guard(h1.device());
    void* st = at::cuda::getCurrentCUDAStream(h1.device().index()).stream();
    // one fp32 allocation: [step workspace] (zero-filled by K1) | inv_norm | pos_cos
    Layout l = make_layout(b, dp, need_bwd);
    l.r_len = 0;  // the gathered row factors live in the peer-mapped set, not here
    l.off_inv = (l.ws_words + 3) / 4 * 4;
    l.off_cos = l.off_inv + 2 * b;
    l.total = l.off_cos + b;
    auto buf = torch::empty({l.total}, h1.options().dtype(torch::kFloat32));
    auto loss = torch::empty({}, h1.options().dtype(torch::kFloat32));
    float* base = buf.data_ptr<float>();
    const maai_peer_sync sync = make_sync(a);
    const long e0 = spans::mark(st);
    check_rc(maai_ntxent_normalize_peer(h1.data_ptr(), h2.data_ptr(), b, d, dt,
                                        reinterpret_cast<const void* const*>(a[A_Z_TAB]), reinterpret_cast<void*>(a[A_MC_Z]),
                                        world, rank, base + l.off_inv, base + l.off_cos, base, (size_t)l.ws_words * 4, &sync, st),
             "maai_ntxent_normalize_peer");
    const long e1 = spans::mark(st);
    check_rc(maai_ntxent_fwd_peer(reinterpret_cast<const void*>(a[A_Z_ALL]), b, world, rank, dp, inv_tau, base + l.off_cos, base,
                                  reinterpret_cast<const void* const*>(a[A_R_TAB]), reinterpret_cast<void*>(a[A_MC_R]),
                                  loss.data_ptr<float>(), MAAI_F_PREZEROED, &sync, st),
             "maai_ntxent_fwd_peer");
    if (e0 >= 0) {
      spans::add(spans::NORMALIZE, e0, e1);
      spans::add(spans::FWD, e1, spans::mark(st));
    }
    if (need_bwd) {
      ctx->save_for_backward({h1, h2});
      ctx->saved_data["buf"] = buf;
      ctx->saved_data["a"] = a;
      ctx->saved_data["inv_tau"] = (double)inv_tau;
      ctx->saved_data["clean"] = true;
    }
    return loss;
  }

  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::variable_list grad_out) {
    const auto saved = ctx->get_saved_variables();
    const auto& h1 = saved[0];
    const auto& h2 = saved[1];
    const auto buf = ctx->saved_data["buf"].toTensor();
    const std::vector<int64_t> a = ctx->saved_data["a"].toIntVector();
    const float inv_tau = (float)ctx->saved_data["inv_tau"].toDouble();
    const bool clean = ctx->saved_data["clean"].toBool();
    ctx->saved_data["clean"] = false;
    const int b = (int)h1.size(0), d = (int)h1.size(1);
    const int rank = (int)a[A_RANK], world = (int)a[A_WORLD];
    const int dp = maai_padded_dim(d);
    const int need = (ctx->needs_input_grad(0) ? 1 : 0) | (ctx->needs_input_grad(1) ? 2 : 0);
    c10::cuda::CUDAGuard guard(h1.device());
    void* st = at::cuda::getCurrentCUDAStream(h1.device().index()).stream();
    Layout l = make_layout(b, dp, true);
    l.off_inv = (l.ws_words + 3) / 4 * 4;
    l.off_cos = l.off_inv + 2 * b;
    float* base = buf.data_ptr<float>();
    auto g = grad_out[0].to(h1.device(), torch::kFloat32).contiguous();
    torch::Tensor dh1, dh2;
    if (need & 1) dh1 = torch::empty_like(h1);
    if (need & 2) dh2 = torch::empty_like(h2);
    const maai_peer_sync sync = make_sync(a);
    const float* r_col = reinterpret_cast<const float*>(a[A_R_COL]);
    const long e0 = spans::mark(st);
    check_rc(maai_ntxent_bwd(reinterpret_cast<const void*>(a[A_Z_ALL]), r_col + (size_t)rank * 2 * b, r_col, 1, base,
                             base + l.off_cos, h1.data_ptr(), h2.data_ptr(), dtype_code(h1), base + l.off_inv,
                             g.data_ptr<float>(), b, world, rank, d, dp, inv_tau, need,
                             (need & 1) ? dh1.data_ptr() : nullptr, (need & 2) ? dh2.data_ptr() : nullptr, base + l.head,
                             clean ? MAAI_F_PREZEROED : 0, &sync, st),
             "maai_ntxent_bwd");
    if (e0 >= 0) spans::add(spans::BWD, e0, spans::mark(st));
    // SetReusePolicy.backward_issued(set): every reader of the set is on the stream now
    int64_t* state = reinterpret_cast<int64_t*>(a[A_STATE]);
    const int64_t nbuf = a[A_NBUF], set = a[A_SET];
    state[1 + set] = 0;
    state[1 + nbuf + set] = state[0];
    return {dh1, dh2, torch::Tensor(), torch::Tensor(), torch::Tensor()};
  }
};

torch::Tensor ntxent_loss_peer(torch::Tensor hidden1, torch::Tensor hidden2, double temperature, std::vector<int64_t> a) {
  const bool need_bwd = torch::GradMode::is_enabled() && (hidden1.requires_grad() || hidden2.requires_grad());
  return NTXentPeerFn::apply(hidden1, hidden2, temperature, need_bwd, a);
}

torch::Tensor ntxent_loss(torch::Tensor hidden1, torch::Tensor hidden2, double temperature) {
  const bool need_bwd = torch::GradMode::is_enabled() && (hidden1.requires_grad() || hidden2.requires_grad());
  return NTXentFn::apply(hidden1, hidden2, temperature, need_bwd);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "C++ autograd binding of the single-rank NT-Xent step over libmaai_ntxent.so";
  m.def("ntxent_loss", &ntxent_loss, "loss = NT-Xent(hidden1, hidden2) on one rank (Objective.py:17-81), autograd-aware");
  m.def("ntxent_loss_peer", &ntxent_loss_peer,
        "multi-rank loss with the fused peer gathers ordered by in-kernel flags (see NTXentPeerFn); the buffer set is the caller's");
  m.def("abi_version", [] { return maai_abi_version(); });
  m.def("span_timing", &spans::enable, "switch the CUDA-event brackets around normalize / fwd / bwd on or off");
  m.def("span_reset", &spans::reset);
  m.def("span_read", &spans::read, "ms per bracketed call since the last reset: [normalize, fwd, bwd]");
}
