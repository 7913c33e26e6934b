// C ABI of libgscan_b200.so (see include/gscan_b200.h): workspace layout and the launch
// sequences for forward / backward / encode / decode-step / greedy decode.
#include "../../include/gscan_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "cnn.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "misc.cuh"
#include "recurrent.cuh"
#include "encoder_res.cuh"
#include "decoder_cluster.cuh"
#include "decoder_v3.cuh"
#include "decoder_v3_bwd.cuh"

using namespace gscan;

#define TRY(expr)                \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)
#define TRYCUDA(expr)                            \
  do {                                           \
    cudaError_t e__ = (expr);                    \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

namespace {

constexpr size_t kMaxSmemBytes = 227 * 1024;

int num_sms() {
  static int per_device[16] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
  int& n = per_device[dev];
  if (n == 0 && (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)) n = 148;
  return n;
}

// ---- optional stage profiling (CUDA events recorded on the caller's stream) ------------------
constexpr int kNumStages = GSCAN_NUM_STAGES;
bool g_profile = false;
cudaEvent_t g_events[kNumStages + 1];
bool g_events_created = false;
bool g_stage_seen[kNumStages + 1];

void prof_mark(int idx, cudaStream_t st) {
  if (!g_profile) return;
  if (!g_events_created) {
    for (auto& e : g_events) cudaEventCreate(&e);
    g_events_created = true;
  }
  cudaEventRecord(g_events[idx], st);
  g_stage_seen[idx] = true;
}

// ---- chain timeline (debug): GSCAN_CHAIN_TIMES=1 records an event at named points of ANY stream and prints, at
// the end of the call, when each point was reached relative to the first one (what the three concurrent chains of a
// call really do to each other cannot be read off a serialised ncu launch list) ---------------------------------
struct ChainMark { const char* name; cudaEvent_t ev; };
std::vector<ChainMark>& chain_marks() { static std::vector<ChainMark> v; return v; }
bool chain_times_on() { static const bool on = getenv("GSCAN_CHAIN_TIMES") != nullptr; return on; }
void chain_mark(const char* name, cudaStream_t st) {
  if (!chain_times_on()) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  chain_marks().push_back({name, e});
}
// GSCAN_CHAIN_TIMES=2: deferred - the marks of a call are printed by the NEXT call with the same title, one step later,
// so that the host stays ahead of the GPU (a report that synchronises exposes the host's launch cadence at the start of
// every pass, which the pipelined loop does not have)
void chain_print(const char* title, std::vector<ChainMark>& marks) {
  fprintf(stderr, "[chain] %s:", title);
  for (auto& m : marks) {
    float ms = 0.f;
    cudaEventSynchronize(m.ev);
    cudaEventElapsedTime(&ms, marks[0].ev, m.ev);
    fprintf(stderr, " %s=%.0f", m.name, ms * 1000.f);
  }
  fprintf(stderr, "\n");
  for (auto& m : marks) cudaEventDestroy(m.ev);
  marks.clear();
}
void chain_report(const char* title) {
  if (!chain_times_on() || chain_marks().empty()) return;
  static const bool deferred = atoi(getenv("GSCAN_CHAIN_TIMES")) == 2;
  if (deferred) {
    static std::vector<std::pair<std::string, std::vector<ChainMark>>> pending;
    for (auto& pr : pending)
      if (pr.first == title) {
        if (!pr.second.empty()) chain_print(title, pr.second);
        pr.second.swap(chain_marks());
        return;
      }
    pending.emplace_back(title, std::vector<ChainMark>());
    pending.back().second.swap(chain_marks());
    return;
  }
  cudaDeviceSynchronize();
  chain_print(title, chain_marks());
}

// ---- internal fork / join ------------------------------------------------------------------
// Independent chains of small kernels (CNN | command encoder | decoder prelude; after the reverse sweep:
// decoder weight gradients | visual keys -> CNN | textual keys -> encoder) run on two helper streams that fork
// from and join back into the caller's stream through events, so the caller still sees plain stream semantics:
// everything a call enqueues completes before anything the caller enqueues on `stream` afterwards.
struct SideStreams {
  cudaStream_t s[3] = {nullptr, nullptr, nullptr};   // [0], [1]: high priority chains; [2]: shadow work, lowest priority
  cudaEvent_t fork_ev[3] = {nullptr, nullptr, nullptr}, join_ev[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t aux_ev[3] = {nullptr, nullptr, nullptr};   // forward pass: [0] command encoder done, [1] Wcomb; backward: [1] gradient destinations zeroed
  bool ok = false;
};
SideStreams* side_streams() {
  static SideStreams per_device[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStreams& S = per_device[dev];
  if (!S.ok) {
    for (int i = 0; i < 3; ++i) {
      // highest priority: the helper chains are long sequences of small kernels, and their CTAs must not queue
      // behind the persistent GEMMs of the caller's stream whenever an SM has room for them
      int prio_lo = 0, prio_hi = 0;
      if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) prio_lo = prio_hi = 0;
      if (cudaStreamCreateWithPriority(&S.s[i], cudaStreamNonBlocking, i < 2 ? prio_hi : prio_lo) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&S.fork_ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&S.join_ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    for (int i = 0; i < 3; ++i)
      if (cudaEventCreateWithFlags(&S.aux_ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    S.ok = true;
  }
  return &S;
}
// side stream i starts after everything enqueued on `from` so far
int fork_side(SideStreams* S, int i, cudaStream_t from) {
  TRYCUDA(cudaEventRecord(S->fork_ev[i], from));
  TRYCUDA(cudaStreamWaitEvent(S->s[i], S->fork_ev[i], 0));
  return 0;
}
// `into` continues after everything enqueued on side stream i so far
int join_side(SideStreams* S, int i, cudaStream_t into) {
  TRYCUDA(cudaEventRecord(S->join_ev[i], S->s[i]));
  TRYCUDA(cudaStreamWaitEvent(into, S->join_ev[i], 0));
  return 0;
}

// ---- workspace layout --------------------------------------------------------------------
struct Layout {
  size_t total = 0;
  size_t take(size_t n) {
    size_t o = total;
    total += (n + 3) & ~size_t(3);
    return o;
  }
  // forward (saved)
  size_t Wt_cnn, feat, KV, enc_x, xg[2], enc_h[2], enc_c[2], enc_g[2], enc_out, h_enc, KT, h0;
  size_t WA_t, WB_t, WC_t, WD_t, WhhE_t[2];
  size_t WA2, WC2, WD2, PT;   // cluster-resident decoder sweep (decoder_cluster.cuh)
  size_t Wcomb;               // [RB][H] = [W_c[:, H:2H] ; W_ih[:, H:2H]] . W_kT: P straight from the encoder outputs
  size_t U, Xe, Cs, gates, alpha, beta, Qp, qT, qV, beta_sum, aux_logp, pre, logp;
  size_t tag;                  // which decoder sweep the forward call ran (checked by the v3 backward kernel)
  size_t progress_f;           // progress words of the forward sweep (output head in its shadow)
  size_t dWh2o;                // hidden_to_output weight gradient formed inside the forward call (gscan_forward_train)
  // backward scratch
  size_t dlogits, dpre, dU, dgates, dd, dqV, dqT, dKT, dKV, dh0, dbeta_aux, dfeat, dconv, dWt_cnn;
  size_t denc_out, dh_enc, dpre0, dga[2], hprev[2], denc_x, dvec;
  size_t progress;             // sweep progress word (shadow scheduling of the weight-gradient GEMM)
  size_t ZV, ZT, WstV, WstT;   // reordered value path of the attentions (v3::attn_value_z_kernel)
  int RA, RB;
};

Layout make_layout(const gscan_dims& d, bool with_backward) {
  Layout L;
  const size_t B = d.B, Ti = d.Ti, Tt = d.Tt, M = (size_t)d.G * d.G, D = 3 * (size_t)d.F, H = d.H, E = d.E, V = d.V;
  CnnShape cs{d.B, d.G, d.C, d.F, d.K3};
  L.RA = (int)(H + (d.conditional_attention ? H : 0) + 4 * H);
  L.RB = (int)((d.conditional_attention ? H : 0) + 4 * H);
  L.Wt_cnn = L.take(cs.wtotal());
  L.feat = L.take(B * M * D);
  L.KV = L.take(B * M * H);
  L.enc_x = L.take(B * Ti * E);
  for (int i = 0; i < 2; ++i) L.xg[i] = L.take(B * Ti * 4 * H);
  for (int i = 0; i < 2; ++i) L.enc_h[i] = L.take(Ti * B * H);
  for (int i = 0; i < 2; ++i) L.enc_c[i] = L.take(Ti * B * H);
  for (int i = 0; i < 2; ++i) L.enc_g[i] = L.take(Ti * B * 4 * H);
  L.enc_out = L.take(Ti * B * H);
  L.h_enc = L.take(B * H);
  L.KT = L.take(Ti * B * H);
  L.h0 = L.take(B * H);
  L.WA_t = L.take(H * L.RA);
  L.WB_t = L.take(H * L.RB);
  L.WC_t = L.take(H * H);
  L.WD_t = L.take(H * 4 * H);
  for (int i = 0; i < 2; ++i) L.WhhE_t[i] = L.take(H * 4 * H);
  L.WA2 = L.take(H * L.RA);
  L.WC2 = L.take(H * H);
  L.WD2 = L.take(H * 4 * H);
  L.PT = L.take(Ti * B * (size_t)L.RB);
  L.Wcomb = L.take((size_t)L.RB * H);
  L.U = L.take((Tt + 1) * B * 4 * H);
  L.Xe = L.take(Tt * B * 4 * H);
  L.Cs = L.take((Tt + 1) * B * H);
  L.gates = L.take(Tt * B * 4 * H);
  L.alpha = L.take(Tt * B * Ti);
  L.beta = L.take(Tt * B * M);
  L.Qp = L.take(Tt * B * H);
  L.qT = L.take(Tt * B * H);
  L.qV = L.take(Tt * B * H);
  L.beta_sum = L.take(B * M);
  L.aux_logp = L.take(B * M);
  L.pre = L.take(Tt * B * H);
  L.logp = L.take(B * Tt * V);
  L.tag = L.take(4);
  L.progress_f = L.take(4);
  L.dWh2o = L.take(V * H);
  if (with_backward) {
    L.dlogits = L.take(Tt * B * V);
    L.dpre = L.take(Tt * B * H);
    L.dU = L.take(Tt * B * 4 * H);
    L.dgates = L.take(Tt * B * 4 * H);
    L.dd = L.take(Tt * B * H);
    L.dqV = L.take(Tt * B * H);
    L.dqT = L.take(Tt * B * H);
    L.dKT = L.take(Ti * B * H);
    L.dKV = L.take(B * M * H);
    L.dh0 = L.take(B * H);
    L.dbeta_aux = L.take(B * M);
    L.dfeat = L.take(B * M * D);
    L.dconv = L.take(B * M * D);
    L.dWt_cnn = L.take(cs.wtotal());
    L.progress = L.take(4);
    L.ZV = L.take(B * M * 5 * H);
    L.ZT = L.take(Ti * B * 6 * H);
    L.WstV = L.take(5 * H * H);
    L.WstT = L.take(6 * H * H);
    L.denc_out = L.take(Ti * B * H);
    L.dh_enc = L.take(B * H);
    L.dpre0 = L.take(B * H);
    for (int i = 0; i < 2; ++i) L.dga[i] = L.take(B * Ti * 4 * H);
    for (int i = 0; i < 2; ++i) L.hprev[i] = L.take(B * Ti * H);
    L.denc_x = L.take(B * Ti * E);
    L.dvec = L.take(2 * H);
  }
  return L;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- examples per CTA for the recurrent sweeps ---------------------------------------------
int env_nb() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("GSCAN_NB");
    v = s ? atoi(s) : 0;
  }
  return v;
}

template <int NB>
size_t dec_smem_bytes(const gscan_dims& d, int RA, int RB, bool bwd, bool greedy) {
  const int M = d.G * d.G;
  size_t f = bwd ? dec_bwd_smem_floats<NB>(d.Ti, M, d.H, kRecThreads)
                 : dec_fwd_smem_floats<NB>(d.Ti, M, d.H, RA, RB, pad4(d.V), greedy, kRecThreads);
  return f * sizeof(float);
}

// returns 4, 2, 1 or 0 (nothing fits)
int pick_nb(const gscan_dims& d, int RA, int RB, bool bwd, bool greedy) {
  int want = env_nb();
  if (want != 1 && want != 2 && want != 4) want = (d.B > 2 * num_sms()) ? 4 : 2;
  for (int nb = want; nb >= 1; nb >>= 1) {
    if (nb * d.H > kRecThreads) continue;
    size_t bytes = nb == 4 ? dec_smem_bytes<4>(d, RA, RB, bwd, greedy)
                 : nb == 2 ? dec_smem_bytes<2>(d, RA, RB, bwd, greedy)
                           : dec_smem_bytes<1>(d, RA, RB, bwd, greedy);
    if (bytes <= kMaxSmemBytes) return nb;
  }
  return 0;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  TRYCUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int NB, bool GREEDY>
int launch_dec_fwd_t(const DecFwdP& p, size_t bytes, cudaStream_t st) {
  TRY(set_smem(decoder_fwd_kernel<NB, GREEDY>, bytes));
  decoder_fwd_kernel<NB, GREEDY><<<ceil_div(p.B, NB), kRecThreads, bytes, st>>>(p);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

int launch_dec_fwd(const gscan_dims& d, const DecFwdP& p, bool greedy, cudaStream_t st) {
  int nb = pick_nb(d, p.RA, p.RB, false, greedy);
  if (nb == 0) return GSCAN_E_UNSUPPORTED;
  if (nb == 4) {
    size_t by = dec_smem_bytes<4>(d, p.RA, p.RB, false, greedy);
    return greedy ? launch_dec_fwd_t<4, true>(p, by, st) : launch_dec_fwd_t<4, false>(p, by, st);
  } else if (nb == 2) {
    size_t by = dec_smem_bytes<2>(d, p.RA, p.RB, false, greedy);
    return greedy ? launch_dec_fwd_t<2, true>(p, by, st) : launch_dec_fwd_t<2, false>(p, by, st);
  }
  size_t by = dec_smem_bytes<1>(d, p.RA, p.RB, false, greedy);
  return greedy ? launch_dec_fwd_t<1, true>(p, by, st) : launch_dec_fwd_t<1, false>(p, by, st);
}

template <int NB>
int launch_dec_bwd_t(const DecBwdP& p, size_t bytes, cudaStream_t st) {
  TRY(set_smem(decoder_bwd_kernel<NB>, bytes));
  decoder_bwd_kernel<NB><<<ceil_div(p.B, NB), kRecThreads, bytes, st>>>(p);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

int launch_dec_bwd(const gscan_dims& d, const DecBwdP& p, cudaStream_t st) {
  int nb = pick_nb(d, 0, 0, true, false);
  if (nb == 0) return GSCAN_E_UNSUPPORTED;
  if (nb == 4) return launch_dec_bwd_t<4>(p, dec_smem_bytes<4>(d, 0, 0, true, false), st);
  if (nb == 2) return launch_dec_bwd_t<2>(p, dec_smem_bytes<2>(d, 0, 0, true, false), st);
  return launch_dec_bwd_t<1>(p, dec_smem_bytes<1>(d, 0, 0, true, false), st);
}


int enc_nb(const gscan_dims& d) {
  int nb = 2;
  while (nb > 1 && nb * d.H > kRecThreads) nb >>= 1;
  return nb;
}

// Resident-weight encoder sweeps (encoder_res.cuh) when W_hh of one direction fits in shared memory beside the
// partial-sum scratch; GSCAN_ENC_STREAMING=1 forces the L2-streaming kernels (A/B measurements, tests).
template <int NB>
int launch_enc_res(const gscan_dims& d, const EncP& p, bool bwd, cudaStream_t st, bool* done) {
  *done = false;
  static const bool force_streaming = getenv("GSCAN_ENC_STREAMING") != nullptr;
  if (force_streaming || NB * d.H > kRecThreads || !enc_res_slices_ok(d.H, kRecThreads)) return 0;
  const size_t by = enc_res_smem_floats<NB>(d.H, kRecThreads, bwd) * sizeof(float);
  if (by > kMaxSmemBytes) return 0;
  dim3 grid(ceil_div(d.B, NB), 2);
  if (bwd) { TRY(set_smem(encoder_bwd_res_kernel<NB>, by)); encoder_bwd_res_kernel<NB><<<grid, kRecThreads, by, st>>>(p); }
  else { TRY(set_smem(encoder_fwd_res_kernel<NB>, by)); encoder_fwd_res_kernel<NB><<<grid, kRecThreads, by, st>>>(p); }
  GSCAN_CHECK_LAUNCH();
  *done = true;
  return 0;
}

int launch_enc(const gscan_dims& d, const EncP& p, bool bwd, cudaStream_t st) {
  bool done = false;
  TRY(launch_enc_res<4>(d, p, bwd, st, &done));
  if (done) return 0;
  int nb = enc_nb(d);
  dim3 grid(ceil_div(d.B, nb), 2);
  if (nb == 2) {
    size_t by = enc_smem_floats<2>(d.H, kRecThreads, bwd) * sizeof(float);
    if (by > kMaxSmemBytes) return GSCAN_E_UNSUPPORTED;
    if (bwd) { TRY(set_smem(encoder_bwd_kernel<2>, by)); encoder_bwd_kernel<2><<<grid, kRecThreads, by, st>>>(p); }
    else { TRY(set_smem(encoder_fwd_kernel<2>, by)); encoder_fwd_kernel<2><<<grid, kRecThreads, by, st>>>(p); }
  } else {
    size_t by = enc_smem_floats<1>(d.H, kRecThreads, bwd) * sizeof(float);
    if (by > kMaxSmemBytes) return GSCAN_E_UNSUPPORTED;
    if (bwd) { TRY(set_smem(encoder_bwd_kernel<1>, by)); encoder_bwd_kernel<1><<<grid, kRecThreads, by, st>>>(p); }
    else { TRY(set_smem(encoder_fwd_kernel<1>, by)); encoder_fwd_kernel<1><<<grid, kRecThreads, by, st>>>(p); }
  }
  GSCAN_CHECK_LAUNCH();
  return 0;
}

// cuStreamWaitValue32 through the runtime's driver entry point table (no libcuda link dependency)
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitValue32Fn stream_wait_value_fn() {
  static StreamWaitValue32Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StreamWaitValue32Fn>(ptr);
  }
  return fn;
}

// SM budgets (tc::ScopedSmCap) of the persistent GEMMs that run beside latency-critical helper chains; 0 = no cap.
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
int cap_prelude() { static const int c = env_int("GSCAN_CAP_PRELUDE", 96); return c; }
int cap_post() { static const int c = env_int("GSCAN_CAP_POST", 80); return c; }

int check_common(const gscan_dims* d, const float* const* params) {
  if (!d || !params) return GSCAN_E_BADARG;
  TRY(gscan_check_dims(d));
  for (int i = 0; i < GSCAN_NUM_PARAMS; ++i) {
    bool optional = (i == GSCAN_P_COND_W || i == GSCAN_P_COND_B) && !d->conditional_attention;
    if (!optional && !params[i]) return GSCAN_E_BADARG;
    if (params[i] && !aligned16(params[i])) return GSCAN_E_UNSUPPORTED;
  }
  return 0;
}

// NT product: C[M,N] = act(A[M,K] (lda) . W[N,K]^T (ldw) + bias + bias2)
int linear(const float* A, long lda, const float* W, long ldw, float* C, long ldc, int M, int N, int K,
           const float* bias, const float* bias2, int act, cudaStream_t st) {
  return launch_gemm(A, lda, 1, W, 1, ldw, C, ldc, M, N, K, bias, bias2, act, 0, 1, st);
}
// NN product: C[M,N] (+)= A[M,K] (lda) . W[K,N] (ldw)
int matmul_nn(const float* A, long lda, const float* W, long ldw, float* C, long ldc, int M, int N, int K,
              int accumulate, cudaStream_t st) {
  return launch_gemm(A, lda, 1, W, ldw, 1, C, ldc, M, N, K, nullptr, nullptr, 0, accumulate, 1, st);
}

int run_cnn_forward(const gscan_dims& d, const float* const* P, const float* situations, DropSrc drop_cnn,
                    float* Wt, float* feat, cudaStream_t st) {
  CnnShape cs{d.B, d.G, d.C, d.F, d.K3};
  cnn_relayout_kernel<<<ceil_div(cs.wtotal(), 256), 256, 0, st>>>(
      cs, const_cast<float*>(P[GSCAN_P_CONV1_W]), const_cast<float*>(P[GSCAN_P_CONV2_W]),
      const_cast<float*>(P[GSCAN_P_CONV3_W]), Wt, 1);
  GSCAN_CHECK_LAUNCH();
  size_t smem = (size_t)cs.M() * cs.C * 8;
  if (smem > 48 * 1024) TRY(set_smem(cnn_forward_kernel, smem));
  const int ysplit = max(1, min(8, ceil_div(cs.M() * cs.D(), 3 * 256)));   // ~3 outputs per thread
  cnn_forward_kernel<<<dim3(d.B, ysplit), 256, smem, st>>>(cs, situations, Wt, P[GSCAN_P_CONV1_B], P[GSCAN_P_CONV2_B],
                                             P[GSCAN_P_CONV3_B], drop_cnn, feat);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

// CNN + visual keys + encoder + textual keys + initial decoder state, shared by forward / encode / greedy.
// The encoder chain runs on `st`.  The CNN chain runs on `cnn_stream` when the caller passes one (and then orders it
// itself), else on helper stream 0, forked from and joined back into `st` here.
int run_encoder_side(const gscan_dims& d, const float* const* P, const long long* commands, const int* cmd_len,
                     const float* situations, DropSrc drop_cnn, DropSrc drop_enc, float* ws,
                     const Layout& L, bool need_keys, cudaStream_t st, cudaStream_t cnn_stream = nullptr,
                     bool cnn_stream_given = false, cudaEvent_t enc_done = nullptr) {
  const int B = d.B, Ti = d.Ti, M = d.G * d.G, D = 3 * d.F, H = d.H, E = d.E;
  // the situation CNN (+ visual keys) is independent of the command encoder
  SideStreams* S = cnn_stream_given ? nullptr : side_streams();
  cudaStream_t sc = cnn_stream_given ? cnn_stream : (S ? S->s[0] : st);
  if (S) TRY(fork_side(S, 0, st));
  TRY(run_cnn_forward(d, P, situations, drop_cnn, ws + L.Wt_cnn, ws + L.feat, sc));
  chain_mark("cnn", sc);
  if (need_keys) TRY(linear(ws + L.feat, D, P[GSCAN_P_VIS_KEY_W], D, ws + L.KV, H, B * M, H, D, nullptr, nullptr, 0, sc));
  chain_mark("KV", sc);
  // weight packing and the zero fills first: the chip is still empty then; behind the input GEMMs they queued for
  // SM slots behind the CNN's CTAs (a 2 us fill took 15 us: tools/step_trace.py)
  {
    PackTable tab;
    tab.n = 2;
    tab.d[0] = PackDesc{P[GSCAN_P_ENC_WHH], H, 0, ws + L.WhhE_t[0], 4 * H, 0, 4 * H, H};
    tab.d[1] = PackDesc{P[GSCAN_P_ENC_WHH_R], H, 0, ws + L.WhhE_t[1], 4 * H, 0, 4 * H, H};
    TRY(launch_pack(tab, st));
  }
  TRYCUDA(cudaMemsetAsync(ws + L.enc_out, 0, sizeof(float) * (size_t)Ti * B * H, st));
  TRYCUDA(cudaMemsetAsync(ws + L.h_enc, 0, sizeof(float) * (size_t)B * H, st));
  // command embeddings and their input-gate pre-activations for both directions
  {
    long n = (long)B * Ti * E;
    embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(commands, d.Ti_stride, P[GSCAN_P_ENC_EMB], E, drop_enc,
                                                              ws + L.enc_x, E, B, Ti, 0);
    GSCAN_CHECK_LAUNCH();
  }
  TRY(linear(ws + L.enc_x, E, P[GSCAN_P_ENC_WIH], E, ws + L.xg[0], 4 * H, B * Ti, 4 * H, E, P[GSCAN_P_ENC_BIH],
             P[GSCAN_P_ENC_BHH], 0, st));
  TRY(linear(ws + L.enc_x, E, P[GSCAN_P_ENC_WIH_R], E, ws + L.xg[1], 4 * H, B * Ti, 4 * H, E, P[GSCAN_P_ENC_BIH_R],
             P[GSCAN_P_ENC_BHH_R], 0, st));
  EncP ep{};
  ep.B = B; ep.Ti = Ti; ep.H = H;
  for (int i = 0; i < 2; ++i) {
    ep.Whh_t[i] = ws + L.WhhE_t[i];
    ep.xg[i] = ws + L.xg[i];
    ep.enc_h[i] = ws + L.enc_h[i];
    ep.enc_c[i] = ws + L.enc_c[i];
    ep.enc_g[i] = ws + L.enc_g[i];
  }
  ep.len = cmd_len;
  ep.enc_out = ws + L.enc_out;
  ep.h_enc = ws + L.h_enc;
  TRY(launch_enc(d, ep, false, st));
  chain_mark("enc_fwd", st);
  // with `enc_done` the caller runs what else hangs off the encoder outputs (initial decoder state, P table) on other
  // streams, beside the textual keys
  if (enc_done) TRYCUDA(cudaEventRecord(enc_done, st));
  if (need_keys) {
    TRY(linear(ws + L.enc_out, H, P[GSCAN_P_TXT_KEY_W], H, ws + L.KT, H, Ti * B, H, H, nullptr, nullptr, 0, st));
    if (!enc_done) TRY(linear(ws + L.h_enc, H, P[GSCAN_P_E2D_W], H, ws + L.h0, H, B, H, H, P[GSCAN_P_E2D_B], nullptr, 1, st));
  }
  if (S) TRY(join_side(S, 0, st));
  return 0;
}

int pack_decoder_weights(const gscan_dims& d, const float* const* P, float* ws, const Layout& L, cudaStream_t st) {
  const int H = d.H;
  PackTable tab;
  int n = 0;
  int r = 0;
  tab.d[n++] = PackDesc{P[GSCAN_P_TXT_QUERY_W], H, 0, ws + L.WA_t, L.RA, r, H, H}; r += H;
  if (d.conditional_attention) { tab.d[n++] = PackDesc{P[GSCAN_P_COND_W], 2 * H, 0, ws + L.WA_t, L.RA, r, H, H}; r += H; }
  tab.d[n++] = PackDesc{P[GSCAN_P_DEC_WHH], H, 0, ws + L.WA_t, L.RA, r, 4 * H, H};
  r = 0;
  if (d.conditional_attention) { tab.d[n++] = PackDesc{P[GSCAN_P_COND_W], 2 * H, H, ws + L.WB_t, L.RB, r, H, H}; r += H; }
  tab.d[n++] = PackDesc{P[GSCAN_P_DEC_WIH], 3 * H, H, ws + L.WB_t, L.RB, r, 4 * H, H};
  tab.d[n++] = PackDesc{P[GSCAN_P_VIS_QUERY_W], H, 0, ws + L.WC_t, H, 0, H, H};
  tab.d[n++] = PackDesc{P[GSCAN_P_DEC_WIH], 3 * H, 2 * H, ws + L.WD_t, 4 * H, 0, 4 * H, H};
  tab.n = n;
  return launch_pack(tab, st);
}

void fill_dec_fwd_common(const gscan_dims& d, const float* const* P, float* ws, const Layout& L, DecFwdP& p) {
  p.B = d.B; p.Ti = d.Ti; p.M = d.G * d.G; p.H = d.H; p.V = d.V; p.cond = d.conditional_attention;
  p.WA_t = ws + L.WA_t; p.RA = L.RA;
  p.WB_t = ws + L.WB_t; p.RB = L.RB;
  p.WC_t = ws + L.WC_t;
  p.WD_t = ws + L.WD_t;
  p.vT = P[GSCAN_P_TXT_ENERGY_W];
  p.vV = P[GSCAN_P_VIS_ENERGY_W];
  p.bc = P[GSCAN_P_COND_B];
}

// ---- cluster-resident decoder sweep (v2) ----------------------------------------------------------
struct ClusterCfg { int C = 0, NB = 0; ClFwdSmem smem; };

int env_dec_version() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("GSCAN_DEC_VERSION");
    v = s ? atoi(s) : 3;
  }
  return v;
}

template <int NB>
cudaError_t cl_fwd_launch(const DecFwd2P& p, const ClFwdSmem& sm, int nclusters, cudaStream_t st, int* max_clusters) {
  auto kern = decoder_fwd_cluster_kernel<NB>;
  size_t bytes = sm.total * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * p.C);
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters) return cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg);
  return cudaLaunchKernelEx(&cfg, kern, p, sm);
}

cudaError_t cl_fwd_dispatch(int NB, const DecFwd2P& p, const ClFwdSmem& sm, int nclusters, cudaStream_t st,
                            int* max_clusters) {
  switch (NB) {
    case 1: return cl_fwd_launch<1>(p, sm, nclusters, st, max_clusters);
    case 2: return cl_fwd_launch<2>(p, sm, nclusters, st, max_clusters);
    case 3: return cl_fwd_launch<3>(p, sm, nclusters, st, max_clusters);
    case 4: return cl_fwd_launch<4>(p, sm, nclusters, st, max_clusters);
    case 5: return cl_fwd_launch<5>(p, sm, nclusters, st, max_clusters);
    case 6: return cl_fwd_launch<6>(p, sm, nclusters, st, max_clusters);
    case 7: return cl_fwd_launch<7>(p, sm, nclusters, st, max_clusters);
    case 8: return cl_fwd_launch<8>(p, sm, nclusters, st, max_clusters);
  }
  return cudaErrorInvalidValue;
}

// Cluster size and examples per cluster for this shape; C == 0 means "use the v1 kernels".
// Cost model: per-step time ~ (slice width) x (examples per cluster); clusters must all be co-resident.
ClusterCfg pick_cluster_cfg(const gscan_dims& d) {
  thread_local gscan_dims cached_d{};
  thread_local ClusterCfg cached{};
  thread_local bool have = false;
  if (have && memcmp(&cached_d, &d, sizeof(d)) == 0) return cached;
  ClusterCfg best{};
  if (env_dec_version() >= 2) {
    const int M = d.G * d.G;
    long best_cost = -1;
    for (int C = 8; C >= 2; --C) {
      if (d.H % (4 * C) != 0) continue;
      const int hs = d.H / C;
      if (hs * (5 + d.conditional_attention) > 32 * kClWarps) continue;
      int NB = min(kClMaxNB, max(1, ceil_div(d.B, max(1, num_sms() / C))));
      for (; NB <= kClMaxNB; ++NB) {
        if (NB * 4 * hs > 2 * kClThreads || NB > kClWarps) break;
        ClFwdSmem sm = cl_fwd_smem(NB, C, d.H, d.Ti, M, d.conditional_attention);
        if (sm.total * sizeof(float) > kMaxSmemBytes) break;
        DecFwd2P p{};
        p.C = C;
        int max_clusters = 0;
        if (cl_fwd_dispatch(NB, p, sm, ceil_div(d.B, NB), nullptr, &max_clusters) != cudaSuccess) {
          cudaGetLastError();
          break;
        }
        const int ncl = ceil_div(d.B, NB);
        if (max_clusters < 1) break;
        const int waves = ceil_div(ncl, max_clusters);
        if (waves > 1 && NB < kClMaxNB) continue;   // try more examples per cluster first
        long cost = (long)hs * NB * waves;
        if (best_cost < 0 || cost < best_cost) {
          best_cost = cost;
          best.C = C;
          best.NB = NB;
          best.smem = sm;
        }
        break;
      }
    }
  }
  if (getenv("GSCAN_DEBUG"))
    fprintf(stderr, "[gscan] decoder sweep config: B=%d H=%d -> C=%d NB=%d clusters=%d smem=%zu B\n", d.B, d.H, best.C, best.NB,
            best.NB ? ceil_div(d.B, best.NB) : 0, best.smem.total * sizeof(float));
  cached_d = d;
  cached = best;
  have = true;
  return best;
}

int launch_dec_fwd_cluster(const gscan_dims& d, const float* const* P, float* ws, const Layout& L, const ClusterCfg& cc,
                           DecFwd2P p, cudaStream_t st) {
  const int H = d.H, hs = H / cc.C;
  ClusterPackP pk{P[GSCAN_P_TXT_QUERY_W], P[GSCAN_P_COND_W], P[GSCAN_P_DEC_WHH], P[GSCAN_P_VIS_QUERY_W],
                  P[GSCAN_P_DEC_WIH], ws + L.WA2, ws + L.WC2, ws + L.WD2, H, cc.C, hs, d.conditional_attention};
  pack_cluster_kernel<<<min(2 * num_sms(), ceil_div(H * (L.RA + 5 * H), 256)), 256, 0, st>>>(pk);
  GSCAN_CHECK_LAUNCH();
  // P = K^T . [W_c[:, H:2H] ; W_ih[:, H:2H]]^T for every command position
  float* PT = ws + L.PT;
  const int cH = d.conditional_attention ? H : 0;
  if (cH) TRY(linear(p.KT, H, P[GSCAN_P_COND_W] + H, 2 * H, PT, L.RB, d.Ti * d.B, H, H, nullptr, nullptr, 0, st));
  TRY(linear(p.KT, H, P[GSCAN_P_DEC_WIH] + H, 3 * H, PT + cH, L.RB, d.Ti * d.B, 4 * H, H, nullptr, nullptr, 0, st));
  p.C = cc.C; p.hs = hs;
  p.WA = ws + L.WA2; p.WC = ws + L.WC2; p.WD = ws + L.WD2; p.PT = PT;
  static const bool want_timeline = getenv("GSCAN_TIMELINE") != nullptr;   // debug only: allocates and synchronises
  long long* tl = nullptr;
  if (want_timeline) {
    cudaMalloc(&tl, sizeof(long long) * 16 * p.T);
    p.timeline = tl;
  }
  TRYCUDA(cl_fwd_dispatch(cc.NB, p, cc.smem, ceil_div(d.B, cc.NB), st, nullptr));
  ++launch_counter();
  if (tl) {
    std::vector<long long> h(16 * (size_t)p.T);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), tl, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
    cudaFree(tl);
    double acc[16] = {0};
    int n = 0;
    for (int t = 2; t + 1 < p.T; ++t, ++n) {
      for (int k = 0; k < 15; ++k) acc[k] += (double)(h[t * 16 + k + 1] - h[t * 16 + k]);
      acc[15] += (double)(h[(t + 1) * 16] - h[t * 16 + 15]);
    }
    fprintf(stderr, "[gscan] fwd cluster timeline (avg cycles per phase over %d steps):", n);
    double tot = 0;
    for (int k = 0; k < 16; ++k) { fprintf(stderr, " %d:%.0f", k, acc[k] / n); tot += acc[k] / n; }
    fprintf(stderr, " total %.0f\n", tot);
  }
  return 0;
}


// ---- register-resident cluster sweep (v3, decoder_v3.cuh): H = 100, 6x6 grid only ---------------------
template <bool COND, bool GREEDY, bool TL = false>
int v3_fwd_prepare(size_t bytes) {
  // per device: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting
  static bool done_d[16] = {}, ok_d[16] = {};
  static size_t done_bytes_d[16] = {};
  int dev_ = 0;
  if (cudaGetDevice(&dev_) != cudaSuccess || dev_ < 0 || dev_ >= 16) return GSCAN_E_UNSUPPORTED;
  bool& done = done_d[dev_];
  bool& ok = ok_d[dev_];
  size_t& done_bytes = done_bytes_d[dev_];
  if (done && bytes <= done_bytes) return ok ? 0 : GSCAN_E_UNSUPPORTED;
  auto kern = v3::dec_fwd_v3_kernel<COND, GREEDY, TL>;
  ok = false;
  done = true;
  done_bytes = bytes;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    cudaGetLastError();
    return GSCAN_E_UNSUPPORTED;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(v3::kC);
  cfg.blockDim = dim3(v3::kThreads);
  cfg.dynamicSmemBytes = bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = v3::kC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
    cudaGetLastError();
    return GSCAN_E_UNSUPPORTED;
  }
  if (getenv("GSCAN_DEBUG")) fprintf(stderr, "[gscan] v3 fwd sweep: smem %zu B, max co-resident clusters %d\n", bytes, max_clusters);
  ok = true;
  return 0;
}

bool v3_shape_ok(const gscan_dims& d) {
  return env_dec_version() >= 3 && d.H == v3::kH && d.G * d.G == v3::kM && d.Ti <= v3::kMaxTi;
}

// P = K^T . [W_c[:, H:2H] ; W_ih[:, H:2H]]^T for every command position
int compute_PT(const gscan_dims& d, const float* const* P, const float* KT, float* PT, int RB, cudaStream_t st) {
  const int H = d.H;
  const int cH = d.conditional_attention ? H : 0;
  if (cH) TRY(linear(KT, H, P[GSCAN_P_COND_W] + H, 2 * H, PT, RB, d.Ti * d.B, H, H, nullptr, nullptr, 0, st));
  TRY(linear(KT, H, P[GSCAN_P_DEC_WIH] + H, 3 * H, PT + cH, RB, d.Ti * d.B, 4 * H, H, nullptr, nullptr, 0, st));
  return 0;
}

// The same table without waiting for the keys: K^T = enc_out . W_kT^T, so P = enc_out . Wcomb^T with
// Wcomb = [W_c[:, H:2H] ; W_ih[:, H:2H]] . W_kT  ([RB x H], weights only: computed off the critical chain)
int compute_Wcomb(const gscan_dims& d, const float* const* P, float* Wcomb, cudaStream_t st) {
  const int H = d.H;
  const int cH = d.conditional_attention ? H : 0;
  if (cH) TRY(matmul_nn(P[GSCAN_P_COND_W] + H, 2 * H, P[GSCAN_P_TXT_KEY_W], H, Wcomb, H, H, H, H, 0, st));
  TRY(matmul_nn(P[GSCAN_P_DEC_WIH] + H, 3 * H, P[GSCAN_P_TXT_KEY_W], H, Wcomb + (size_t)cH * H, H, 4 * H, H, H, 0, st));
  return 0;
}
int compute_PT_direct(const gscan_dims& d, const float* enc_out, const float* Wcomb, float* PT, int RB, cudaStream_t st) {
  return linear(enc_out, d.H, Wcomb, d.H, PT, RB, d.Ti * d.B, RB, d.H, nullptr, nullptr, 0, st);
}

// stamps: [T][16 stamps][16 warps] of CTA 0.  Prints, per stamp interval, the average over the steps of the time between
// consecutive stamps of warp 0 (the round-1 figure), and - what the chain of dependences really looks like - for every
// stamp the average time at which the EARLIEST and the LATEST warp reached it, relative to the step's start.
// GSCAN_TIMELINE=<file> also dumps the raw stamps.
void print_timeline(const char* what, long long* tl, int T, cudaStream_t st) {
  constexpr int NS = v3::kTlStamps;
  std::vector<long long> h((size_t)NS * 16 * T);
  cudaStreamSynchronize(st);
  cudaMemcpy(h.data(), tl, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
  cudaFree(tl);
  auto at = [&](int t, int k, int w) { return h[((size_t)t * NS + k) * 16 + w]; };
  double acc[16] = {0}, lo[NS] = {0}, hi[NS] = {0};
  int n = 0;
  const bool rev = at(2, 0, 0) > at(T - 3, 0, 0);   // the backward sweep walks t downwards
  for (int t = 2; t + 2 < T; ++t, ++n) {
    const int tn = rev ? t - 1 : t + 1;
    for (int k = 0; k < 15; ++k) acc[k] += (double)(at(t, k + 1, 0) - at(t, k, 0));
    acc[15] += (double)(at(tn, 0, 0) - at(t, 15, 0));
    long long t0 = at(t, 0, 0);
    for (int w = 1; w < 16; ++w) if (at(t, 0, w) && at(t, 0, w) < t0) t0 = at(t, 0, w);
    for (int k = 0; k < NS; ++k) {
      long long mn = 0, mx = 0;
      for (int w = 0; w < 16; ++w) {
        const long long v = at(t, k, w);
        if (!v) continue;
        if (!mn || v < mn) mn = v;
        if (v > mx) mx = v;
      }
      if (mx) { lo[k] += (double)(mn - t0); hi[k] += (double)(mx - t0); }
    }
  }
  if (n == 0) return;
  fprintf(stderr, "[gscan] %s timeline (avg cycles per phase over %d steps):", what, n);
  double tot = 0;
  for (int k = 0; k < 16; ++k) { fprintf(stderr, " %d:%.0f", k, acc[k] / n); tot += acc[k] / n; }
  fprintf(stderr, " total %.0f\n", tot);
  fprintf(stderr, "[gscan] %s stamps reached at (first warp / last warp, cycles after the step's first stamp):", what);
  for (int k = 0; k < NS; ++k) fprintf(stderr, " %d:%.0f/%.0f", k, lo[k] / n, hi[k] / n);
  fprintf(stderr, "\n");
  const char* path = getenv("GSCAN_TIMELINE");
  if (path && path[0] && strcmp(path, "1") != 0) {
    std::string f = std::string(path) + (rev ? ".bwd.bin" : ".fwd.bin");
    if (FILE* fp = fopen(f.c_str(), "wb")) { fwrite(h.data(), sizeof(long long), h.size(), fp); fclose(fp); }
  }
}

int launch_dec_fwd_v3(const gscan_dims& d, const float* const* P, float* ws, const Layout& L, v3::DecFwd3P p,
                      bool greedy, cudaStream_t st, bool pt_ready = false) {
  const int cond = d.conditional_attention ? 1 : 0;
  const v3::FwdSmem sm = v3::fwd_smem(d.Ti, cond, greedy ? d.V : 0);
  const size_t bytes = (size_t)sm.total * sizeof(float);
  if (bytes > kMaxSmemBytes) return GSCAN_E_UNSUPPORTED;
  if (greedy) TRY(cond ? (v3_fwd_prepare<true, true>(bytes)) : (v3_fwd_prepare<false, true>(bytes)));
  else TRY(cond ? (v3_fwd_prepare<true, false>(bytes)) : (v3_fwd_prepare<false, false>(bytes)));
  if (!pt_ready) TRY(compute_PT(d, P, p.KT, ws + L.PT, L.RB, st));
  p.PT = ws + L.PT;
  p.W_qT = P[GSCAN_P_TXT_QUERY_W]; p.W_c = P[GSCAN_P_COND_W]; p.W_hh = P[GSCAN_P_DEC_WHH];
  p.W_qV = P[GSCAN_P_VIS_QUERY_W]; p.W_ih = P[GSCAN_P_DEC_WIH];
  static const bool want_timeline = getenv("GSCAN_TIMELINE") != nullptr;   // debug only: allocates and synchronises
  long long* tl = nullptr;
  if (want_timeline && !greedy && cond) {   // the instrumented instantiation exists for the paper configuration only
    TRY((v3_fwd_prepare<true, false, true>(bytes)));
    cudaMalloc(&tl, sizeof(long long) * v3::kTlStamps * 16 * p.T);
    cudaMemsetAsync(tl, 0, sizeof(long long) * v3::kTlStamps * 16 * p.T, st);
    p.timeline = tl;
  }
  const int grid = ceil_div(d.B, v3::kNB) * v3::kC;
  if (tl) {
    v3::dec_fwd_v3_kernel<true, false, true><<<grid, v3::kThreads, bytes, st>>>(p);
  } else if (greedy) {
    if (cond) v3::dec_fwd_v3_kernel<true, true><<<grid, v3::kThreads, bytes, st>>>(p);
    else v3::dec_fwd_v3_kernel<false, true><<<grid, v3::kThreads, bytes, st>>>(p);
  } else {
    if (cond) v3::dec_fwd_v3_kernel<true, false><<<grid, v3::kThreads, bytes, st>>>(p);
    else v3::dec_fwd_v3_kernel<false, false><<<grid, v3::kThreads, bytes, st>>>(p);
  }
  GSCAN_CHECK_LAUNCH();
  if (tl) print_timeline("v3 fwd", tl, p.T, st);
  return 0;
}

template <bool COND, bool TL = false>
int v3_bwd_prepare(size_t bytes) {
  // per device: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting
  static bool done_d[16] = {}, ok_d[16] = {};
  static size_t done_bytes_d[16] = {};
  int dev_ = 0;
  if (cudaGetDevice(&dev_) != cudaSuccess || dev_ < 0 || dev_ >= 16) return GSCAN_E_UNSUPPORTED;
  bool& done = done_d[dev_];
  bool& ok = ok_d[dev_];
  size_t& done_bytes = done_bytes_d[dev_];
  if (done && bytes <= done_bytes) return ok ? 0 : GSCAN_E_UNSUPPORTED;
  auto kern = v3::dec_bwd_v3_kernel<COND, TL>;
  ok = false;
  done = true;
  done_bytes = bytes;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    cudaGetLastError();
    return GSCAN_E_UNSUPPORTED;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(v3::kC);
  cfg.blockDim = dim3(v3::kThreads);
  cfg.dynamicSmemBytes = bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = v3::kC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
    cudaGetLastError();
    return GSCAN_E_UNSUPPORTED;
  }
  if (getenv("GSCAN_DEBUG")) fprintf(stderr, "[gscan] v3 bwd sweep: smem %zu B, max co-resident clusters %d\n", bytes, max_clusters);
  ok = true;
  return 0;
}

bool v3_bwd_shape_ok(const gscan_dims& d) {
  static const bool off = getenv("GSCAN_BWD_V1") != nullptr;
  if (off || !v3_shape_ok(d) || d.Ti > v3::kMaxTiB) return false;
  const size_t bytes = (size_t)v3::bwd_smem(d.Ti, d.conditional_attention ? 1 : 0).total * sizeof(float);
  return bytes <= kMaxSmemBytes;
}

int launch_dec_bwd_v3(const gscan_dims& d, v3::DecBwd3P p, cudaStream_t st) {
  const int cond = d.conditional_attention ? 1 : 0;
  const size_t bytes = (size_t)v3::bwd_smem(d.Ti, cond).total * sizeof(float);
  TRY(cond ? v3_bwd_prepare<true>(bytes) : v3_bwd_prepare<false>(bytes));
  static const bool want_timeline = getenv("GSCAN_TIMELINE") != nullptr;
  long long* tl = nullptr;
  if (want_timeline && cond) {
    TRY((v3_bwd_prepare<true, true>(bytes)));
    cudaMalloc(&tl, sizeof(long long) * v3::kTlStamps * 16 * p.T);
    cudaMemsetAsync(tl, 0, sizeof(long long) * v3::kTlStamps * 16 * p.T, st);
    p.timeline = tl;
  }
  const int grid = ceil_div(d.B, v3::kNB) * v3::kC;
  if (tl) v3::dec_bwd_v3_kernel<true, true><<<grid, v3::kThreads, bytes, st>>>(p);
  else if (cond) v3::dec_bwd_v3_kernel<true><<<grid, v3::kThreads, bytes, st>>>(p);
  else v3::dec_bwd_v3_kernel<false><<<grid, v3::kThreads, bytes, st>>>(p);
  GSCAN_CHECK_LAUNCH();
  if (tl) print_timeline("v3 bwd", tl, p.T, st);
  return 0;
}

// the DropSrc of dropout site `which` (0 CNN features, 1 command embeddings, 2 target embeddings): an explicit mask, or the
// Philox stream of (seed, call offset, site)
DropSrc make_drop(const float* mask, const gscan_dropout* rng, int which) {
  DropSrc d(mask);
  if (mask != nullptr || rng == nullptr) return d;
  const float p = which == 0 ? rng->p_cnn : (which == 1 ? rng->p_enc : rng->p_dec);
  if (!(p > 0.f)) return d;
  d.rng = 1;
  d.k0 = (unsigned int)rng->seed;
  d.k1 = (unsigned int)(rng->seed >> 32) ^ (unsigned int)(rng->offset >> 32);
  d.c2 = (unsigned int)which;
  d.c3 = (unsigned int)rng->offset;
  const double t = (double)p * 4294967296.0;
  d.thresh = t >= 4294967295.0 ? 0xffffffffu : (unsigned int)t;
  d.scale = 1.f / (1.f - p);
  return d;
}

int forward_impl(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                 const float* situations, const int64_t* targets, DropSrc drop_cnn, DropSrc drop_enc, DropSrc drop_dec,
                 float* ws, size_t ws_floats, float* logp, float* aux_logp, void* stream,
                 const float* d_logp_early = nullptr, cudaEvent_t d_logp_ready = nullptr);

// gscan_forward_train: which workspaces hold an output-head backward pass already (dpre, dU, dW_h2o formed inside the
// forward call from a d_logp known in advance), and for which d_logp pointer.  The backward call on the same workspace
// with the same pointer skips that stage; any other d_logp recomputes it.
std::mutex& early_head_mutex() { static std::mutex m; return m; }
std::map<const float*, const float*>& early_head_map() { static std::map<const float*, const float*> m; return m; }
void early_head_set(const float* ws, const float* d_logp) {
  std::lock_guard<std::mutex> g(early_head_mutex());
  if (d_logp) early_head_map()[ws] = d_logp;
  else early_head_map().erase(ws);
}
bool early_head_take(const float* ws, const float* d_logp) {
  std::lock_guard<std::mutex> g(early_head_mutex());
  auto it = early_head_map().find(ws);
  if (it == early_head_map().end()) return false;
  const bool same = it->second == d_logp;
  early_head_map().erase(it);
  return same;
}
int backward_impl(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                  const float* situations, const int64_t* targets, DropSrc drop_cnn, DropSrc drop_enc, DropSrc drop_dec,
                  float* ws, size_t ws_floats, const float* d_logp, const float* d_aux_logp, float* const* G, void* stream);

}  // namespace

// =================================================================================================
extern "C" {

int gscan_abi_version(void) { return GSCAN_ABI_VERSION; }

unsigned long long gscan_launch_count(void) { return launch_counter(); }

int gscan_profile(int enable) {
  g_profile = enable != 0;
  for (auto& s : g_stage_seen) s = false;
  return GSCAN_OK;
}

int gscan_profile_read(float* stage_ms) {
  if (!stage_ms) return GSCAN_E_BADARG;
  // stage i spans events i -> i+1, except the forward/backward seam (4 -> 5)
  for (int i = 0; i < kNumStages; ++i) {
    stage_ms[i] = -1.f;
    if (i == 4 || !g_stage_seen[i] || !g_stage_seen[i + 1]) continue;
    TRYCUDA(cudaEventSynchronize(g_events[i + 1]));
    float ms = 0.f;
    TRYCUDA(cudaEventElapsedTime(&ms, g_events[i], g_events[i + 1]));
    stage_ms[i] = ms;
  }
  return GSCAN_OK;
}

int gscan_check_dims(const gscan_dims* d) {
  if (!d) return GSCAN_E_BADARG;
  if (d->B < 1 || d->Ti < 1 || d->Tt < 1 || d->G < 1 || d->C < 1 || d->F < 1 || d->K3 < 1 || d->E < 1 || d->H < 4 ||
      d->Vi < 1 || d->V < 1 || d->Ti_stride < d->Ti)
    return GSCAN_E_BADARG;
  if (d->H % 4 != 0 || d->H > kRecThreads) return GSCAN_E_UNSUPPORTED;
  if ((d->K3 & 1) == 0) return GSCAN_E_UNSUPPORTED;            // 'same' padding k//2 needs odd k
  if (d->V > 32 * kMaxVPerLane) return GSCAN_E_UNSUPPORTED;
  if ((size_t)d->G * d->G * d->C * 8 > kMaxSmemBytes) return GSCAN_E_UNSUPPORTED;
  Layout L = make_layout(*d, false);
  if (pick_nb(*d, L.RA, L.RB, true, false) == 0) return GSCAN_E_UNSUPPORTED;
  if (pick_nb(*d, L.RA, L.RB, false, d->V <= 32) == 0) return GSCAN_E_UNSUPPORTED;
  return GSCAN_OK;
}

size_t gscan_workspace_floats(const gscan_dims* d) { return d ? make_layout(*d, true).total : 0; }
size_t gscan_encode_workspace_floats(const gscan_dims* d) {
  if (!d) return 0;
  gscan_dims e = *d;
  e.Tt = 1;
  return make_layout(e, false).total;
}
size_t gscan_step_workspace_floats(const gscan_dims* d) {
  if (!d) return 0;
  gscan_dims e = *d;
  e.Tt = 1;
  return make_layout(e, false).total;
}
size_t gscan_greedy_workspace_floats(const gscan_dims* d) {
  if (!d) return 0;
  gscan_dims e = *d;
  e.Tt = 1;
  // + tables: XeTab [V][4H], Wout [V][4H], OutE [V][V], Wo_t [3H][Vp]
  size_t extra = 2 * (size_t)d->V * 4 * d->H + (size_t)d->V * d->V + 3 * (size_t)d->H * pad4(d->V) + 64;
  return make_layout(e, false).total + extra;
}

int gscan_forward(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                  const float* situations, const int64_t* targets, const float* drop_cnn, const float* drop_enc,
                  const float* drop_dec, float* ws, size_t ws_floats, float* logp, float* aux_logp, void* stream) {
  return forward_impl(d, P, commands, cmd_len, situations, targets, DropSrc(drop_cnn), DropSrc(drop_enc), DropSrc(drop_dec), ws,
                      ws_floats, logp, aux_logp, stream);
}

int gscan_forward_rng(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                      const float* situations, const int64_t* targets, const gscan_dropout* rng, float* ws,
                      size_t ws_floats, float* logp, float* aux_logp, void* stream) {
  return forward_impl(d, P, commands, cmd_len, situations, targets, make_drop(nullptr, rng, 0), make_drop(nullptr, rng, 1),
                      make_drop(nullptr, rng, 2), ws, ws_floats, logp, aux_logp, stream);
}

int gscan_forward_train(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                        const float* situations, const int64_t* targets, const float* drop_cnn, const float* drop_enc,
                        const float* drop_dec, const gscan_dropout* rng, float* ws, size_t ws_floats, float* logp,
                        float* aux_logp, const float* d_logp, void* d_logp_ready, void* stream) {
  if (!d_logp) return GSCAN_E_BADARG;
  return forward_impl(d, P, commands, cmd_len, situations, targets, make_drop(drop_cnn, rng, 0), make_drop(drop_enc, rng, 1),
                      make_drop(drop_dec, rng, 2), ws, ws_floats, logp, aux_logp, stream, d_logp, (cudaEvent_t)d_logp_ready);
}

int gscan_backward_rng(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                       const float* situations, const int64_t* targets, const gscan_dropout* rng, float* ws,
                       size_t ws_floats, const float* d_logp, const float* d_aux_logp, float* const* G, void* stream) {
  return backward_impl(d, P, commands, cmd_len, situations, targets, make_drop(nullptr, rng, 0), make_drop(nullptr, rng, 1),
                       make_drop(nullptr, rng, 2), ws, ws_floats, d_logp, d_aux_logp, G, stream);
}

int gscan_dropout_mask(const gscan_dropout* rng, int32_t which, size_t n, float* out, void* stream) {
  if (!rng || !out || which < 0 || which > 2) return GSCAN_E_BADARG;
  if (n == 0) return GSCAN_OK;
  DropSrc src = make_drop(nullptr, rng, which);
  if (!src.rng) {   // p == 0: the mask is all ones
    src.rng = 1; src.thresh = 0; src.scale = 1.f;
  }
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, out, (long)n);
  GSCAN_CHECK_LAUNCH();
  return GSCAN_OK;
}

}  // extern "C"

namespace {

int forward_impl(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                 const float* situations, const int64_t* targets, DropSrc drop_cnn, DropSrc drop_enc, DropSrc drop_dec,
                 float* ws, size_t ws_floats, float* logp, float* aux_logp, void* stream,
                 const float* d_logp_early, cudaEvent_t d_logp_ready) {
  TRY(check_common(d, P));
  if (!commands || !cmd_len || !situations || !targets || !ws || !logp) return GSCAN_E_BADARG;
  if (d->auxiliary_task && !aux_logp) return GSCAN_E_BADARG;
  if (!aligned16(ws)) return GSCAN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  Layout L = make_layout(*d, true);
  if (ws_floats < L.total) return GSCAN_E_WORKSPACE;
  const int B = d->B, Tt = d->Tt, M = d->G * d->G, H = d->H, V = d->V;
  const long long* cmds = reinterpret_cast<const long long*>(commands);
  const long long* tgts = reinterpret_cast<const long long*>(targets);

  prof_mark(0, st);
  chain_mark("fwd_start", st);
  // gscan_forward_train: the output-head backward runs inside this call (fused kernel only: V <= kHeadMaxV, H <= 128)
  early_head_set(ws, nullptr);
  const bool early_head = d_logp_early != nullptr && V <= kHeadMaxV && H <= 128 && getenv("GSCAN_HEAD_UNFUSED") == nullptr;
  // Three chains before the sweep: the command encoder (a long chain of small kernels: high-priority helper stream 1,
  // issued first), the decoder prelude (depends on targets and weights only; one big GEMM, capped: high-priority
  // helper stream 0) and the situation CNN (wide kernels of small CTAs that co-reside with the GEMM's: the caller's
  // stream, which fills whatever the other two leave).  Measured orders: prelude on a helper stream and encoder side
  // on the caller's 213 us; encoder side on helper streams and prelude on the caller's 183 us.
  SideStreams* S = side_streams();
  cudaStream_t se = S ? S->s[1] : st, sp = S ? S->s[0] : st;
  // GSCAN_FWD_SCHED: 0 = round-1 schedule (K^T, h0, P one after the other behind the encoder);  1 = the three products
  // that hang off the encoder outputs run side by side (K^T on the encoder's stream, h0 on the prelude's, P - straight
  // from enc_out through Wcomb - on the caller's).  (Holding the prelude GEMM back until the encoder is done - its
  // persistent CTAs and the encoder's 100 CTAs exclude each other on an SM, and the encoder then runs in one wave, 31
  // instead of 55 us - ended later overall: profiles/r02_negative_results.md.)
  static const int sched = env_int("GSCAN_FWD_SCHED", 1);
  const bool par_tail = S && sched >= 1 && v3_shape_ok(*d);
  if (S) {
    TRY(fork_side(S, 1, st));
    TRY(fork_side(S, 0, st));
  }
  // (behind the forks: the helper chains do not wait for these fills)
  TRYCUDA(cudaMemsetAsync(ws + L.progress_f, 0, 4 * sizeof(unsigned int), st));   // progress words of the sweep (far ahead of it)
  if (early_head) TRYCUDA(cudaMemsetAsync(ws + L.dWh2o, 0, sizeof(float) * (size_t)V * H, st));
  if (par_tail) {   // weights only, ahead of the prelude (which has slack)
    TRY(compute_Wcomb(*d, P, ws + L.Wcomb, sp));
    TRYCUDA(cudaEventRecord(S->aux_ev[1], sp));
  }
  TRY(run_encoder_side(*d, P, cmds, cmd_len, situations, drop_cnn, drop_enc, ws, L, true, se, st, true,
                       par_tail ? S->aux_ev[0] : nullptr));
  if (par_tail) {   // the caller's stream: CNN, visual keys, then P as soon as the encoder is done
    TRYCUDA(cudaStreamWaitEvent(st, S->aux_ev[0], 0));
    TRYCUDA(cudaStreamWaitEvent(st, S->aux_ev[1], 0));
    TRY(compute_PT_direct(*d, ws + L.enc_out, ws + L.Wcomb, ws + L.PT, L.RB, st));
  }
  TRY(pack_decoder_weights(*d, P, ws, L, sp));
  // target embeddings straight into the e-block of U (time-major rows, group 0 reserved for h_{-1})
  {
    long n = (long)B * Tt * H;
    const bool vec4 = (H & 3) == 0 && aligned16(P[GSCAN_P_DEC_EMB]) && (!drop_dec.mask || aligned16(drop_dec.mask));
    if (vec4)
      embed4_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, sp>>>(tgts, Tt, P[GSCAN_P_DEC_EMB], H, drop_dec, ws + L.U,
                                                                     4 * H, B, Tt, 1);
    else
      embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, sp>>>(tgts, Tt, P[GSCAN_P_DEC_EMB], H, drop_dec, ws + L.U,
                                                                4 * H, B, Tt, 1);
    GSCAN_CHECK_LAUNCH();
  }
  float* U1 = ws + L.U + (size_t)B * 4 * H;
  // input-gate pre-activations of every step at once: Xe = E . W_ih[:, :H]^T + b_ih + b_hh
  {
    tc::ScopedSmCap cap(S ? cap_prelude() : 0);   // leave SMs to the command encoder running beside it
    TRY(linear(U1, 4 * H, P[GSCAN_P_DEC_WIH], 3 * H, ws + L.Xe, 4 * H, Tt * B, 4 * H, H, P[GSCAN_P_DEC_BIH],
               P[GSCAN_P_DEC_BHH], 0, sp));
  }
  if (par_tail) {   // initial decoder state, beside K^T and P
    TRYCUDA(cudaStreamWaitEvent(sp, S->aux_ev[0], 0));
    TRY(linear(ws + L.h_enc, H, P[GSCAN_P_E2D_W], H, ws + L.h0, H, B, H, H, P[GSCAN_P_E2D_B], nullptr, 1, sp));
  }
  chain_mark("s0:prelude", sp);
  if (S) TRY(join_side(S, 0, st));
  if (S) TRY(join_side(S, 1, st));
  prof_mark(1, st);
  chain_mark("m:joined", st);
  DecFwdP p{};
  fill_dec_fwd_common(*d, P, ws, L, p);
  p.T = Tt;
  p.KT = ws + L.KT; p.KV = ws + L.KV; p.cmd_len = cmd_len;
  p.h_init = ws + L.h0; p.c_init = ws + L.h0;
  p.Xe = ws + L.Xe;
  p.U = ws + L.U; p.Cs = ws + L.Cs; p.gates = ws + L.gates; p.alpha = ws + L.alpha; p.beta = ws + L.beta;
  p.Qp = ws + L.Qp; p.qT = ws + L.qT; p.qV = ws + L.qV; p.beta_sum = ws + L.beta_sum;
  prof_mark(2, st);
  const ClusterCfg cc = v3_shape_ok(*d) ? ClusterCfg{} : pick_cluster_cfg(*d);
  bool v3_done = false;
  // output projection + log-softmax of the steps [t0, t1) (rows t0*B .. t1*B of the time-major lists)
  const size_t head_smem = (size_t)V * (H + 1) * sizeof(float);
  TRY(set_smem(out_logsoftmax_kernel, head_smem > 80 * 1024 ? head_smem : (size_t)80 * 1024));
  // keep_off_sweep_sms: a shadow launch must not land on the SMs of the sweep (its CTAs would share their issue slots
  // with the latency-bound recurrence: measured +45 us on the sweep) - asking for 48 KB of shared memory makes the
  // CTAs fit only where the sweep (185 KB) is not resident
  auto head_rows = [&](int t0, int t1, cudaStream_t s_, bool keep_off_sweep_sms = false) -> int {
    const long r0 = (long)t0 * B, r1 = (long)t1 * B;
    if (r1 <= r0) return 0;
    // (the forward sweep holds 150 KB of an SM's 227 KB: 80 KB do not fit beside it)
    const size_t head_smem_l = keep_off_sweep_sms && head_smem < 80 * 1024 ? (size_t)80 * 1024 : head_smem;
    TRY(linear(U1 + (size_t)r0 * 4 * H, 4 * H, P[GSCAN_P_O2H_W], 4 * H, ws + L.pre + (size_t)r0 * H, H, (int)(r1 - r0), H,
               4 * H, nullptr, nullptr, 0, s_));
    // a shadow launch fits two CTAs on an idle SM (the shared-memory request above): 1024 threads each for full occupancy
    const int hthreads = keep_off_sweep_sms ? 1024 : 256;
    const int blocks = min(ceil_div((int)(r1 - r0), hthreads / 32), 8 * num_sms());
    // (the caller's copy of the log-probabilities is written here too: no copy kernel between the head and the loss)
    out_logsoftmax_kernel<<<blocks, hthreads, head_smem_l, s_>>>(ws + L.pre, P[GSCAN_P_H2O_W], H, V, B, Tt, ws + L.logp, nullptr,
                                                            r0, r1, logp);
    GSCAN_CHECK_LAUNCH();
    return 0;
  };
  int head_t0 = 0;               // first step whose output head is still to do after the sweep
  bool fwd_shadow_used = false;
  // output-head backward of the steps [t0, t1) from a d_logp known in advance: log-softmax backward, dpre and the
  // hidden_to_output weight gradient (into the workspace: the caller's gradient buffer is not known here) in one pass
  // over the rows, then dU = dpre . W_o2h for them
  const size_t hb_smem = 2 * sizeof(float) * (size_t)V * H;
  if (early_head) TRY(set_smem(head_bwd_fused_kernel, hb_smem > 80 * 1024 ? hb_smem : (size_t)80 * 1024));
  auto head_bwd_rows = [&](int t0, int t1, cudaStream_t s_, bool keep_off_sweep_sms) -> int {
    const long r0 = (long)t0 * B, r1 = (long)t1 * B;
    if (r1 <= r0) return 0;
    const size_t sm_l = keep_off_sweep_sms && hb_smem < 80 * 1024 ? (size_t)80 * 1024 : hb_smem;
    head_bwd_fused_kernel<<<min(ceil_div((int)(r1 - r0), 8), 2 * num_sms()), 256, sm_l, s_>>>(
        d_logp_early, ws + L.logp, ws + L.pre, P[GSCAN_P_H2O_W], H, V, B, Tt, ws + L.dpre, ws + L.dWh2o, r0, r1);
    GSCAN_CHECK_LAUNCH();
    TRY(matmul_nn(ws + L.dpre + (size_t)r0 * H, H, P[GSCAN_P_O2H_W], 4 * H, ws + L.dU + (size_t)r0 * 4 * H, 4 * H,
                  (int)(r1 - r0), 4 * H, H, 0, s_));
    return 0;
  };
  int head_bwd_t0 = 0;           // first step whose output-head backward is still to do after the sweep
  if (v3_shape_ok(*d)) {
    v3::DecFwd3P p3{};
    p3.B = B; p3.T = Tt; p3.Ti = d->Ti;
    p3.vT = p.vT; p3.vV = p.vV; p3.bc = p.bc;
    p3.KT = p.KT; p3.KV = p.KV; p3.cmd_len = cmd_len; p3.h_init = p.h_init; p3.c_init = p.c_init; p3.Xe = p.Xe;
    p3.U = p.U; p3.Cs = p.Cs; p3.gates = p.gates; p3.alpha = p.alpha; p3.beta = p.beta;
    p3.Qp = p.Qp; p3.qT = p.qT; p3.qV = p.qV; p3.beta_sum = p.beta_sum;
    // Shadow schedule of the output head: the sweep holds 125 of the 148 SMs for ~0.67 ms and completes the rows of
    // [e | h | c_T | c_V] step by step.  Once every CTA has published the steps < cut (a counter in the workspace,
    // awaited by helper stream 2 through cuStreamWaitValue32) the output projection and the log-softmax of THOSE rows
    // run on the idle SMs beside the rest of the sweep; only the steps >= the last cut are left for after it.
    // (Round 1 tried this with the compute warps signalling: the signal code cost the sweep 33 us.  Now the I/O warps,
    // off the critical path, fence and one thread adds at a barrier that exists anyway.)
    int f_cut[4];
    int nf = 0;
    {
      const char* spec = getenv("GSCAN_FWD_SHADOW_CUTS");
      if (!spec) spec = "20,44,68,92";
      int prev = 0;
      for (const char* q = spec; *q && nf < 4;) {
        const int pct = atoi(q);
        const int t = pct * Tt / 100;
        if (pct > 0 && pct < 100 && t > prev && t < Tt) { f_cut[nf++] = t; prev = t; }
        while (*q && *q != ',') ++q;
        if (*q == ',') ++q;
      }
    }
    const int sweep_ctas_f = ceil_div(B, v3::kNB) * v3::kC;
    const int idle_f = num_sms() - sweep_ctas_f;
    bool fshadow = S && stream_wait_value_fn() && env_int("GSCAN_FWD_SHADOW", 1) != 0 && Tt >= 16 && idle_f >= 8 && nf > 0;
    unsigned int* progress_f = reinterpret_cast<unsigned int*>(ws + L.progress_f);
    if (fshadow) {
      p3.progress = progress_f;   // (zeroed at the top of the call)
      p3.n_signals = nf;
      for (int k = 0; k < nf; ++k) p3.t_signal[k] = f_cut[k] - 1;
      TRY(fork_side(S, 2, st));   // helper stream 2 starts from here, NOT from the end of the sweep
    }
    int rc = launch_dec_fwd_v3(*d, P, ws, L, p3, false, st, par_tail);
    if (rc == 0) v3_done = true;
    else if (rc != GSCAN_E_UNSUPPORTED) return rc;
    if (fshadow && !v3_done) {   // nobody will signal
      fshadow = false;
      TRY(join_side(S, 2, st));
    }
    if (fshadow) {
      cudaStream_t sh = S->s[2];
      tc::ScopedSmCap cap(idle_f);
      for (int k = 0; k < nf; ++k) {
        if (stream_wait_value_fn()((CUstream)sh, (CUdeviceptr)(progress_f + k), (cuuint32_t)sweep_ctas_f, 0u /* GEQ */) !=
            CUDA_SUCCESS) {
          if (k > 0) { join_side(S, 2, st); return GSCAN_E_UNSUPPORTED; }
          fshadow = false;          // stream memory operations unavailable: everything after the sweep
          TRY(join_side(S, 2, st));
          break;
        }
        TRY(head_rows(k == 0 ? 0 : f_cut[k - 1], f_cut[k], sh, true));
        // ... and, with d_logp known in advance, their backward pass too - for the first chunks only: the 23 idle SMs
        // cannot take all of it before the sweep ends (GSCAN_FWD_SHADOW_BWD_CHUNKS)
        static const int bwd_chunks = env_int("GSCAN_FWD_SHADOW_BWD_CHUNKS", 3);   // of 4: the last one would end after the sweep
        if (early_head && k < bwd_chunks) {
          if (k == 0 && d_logp_ready) TRYCUDA(cudaStreamWaitEvent(sh, d_logp_ready, 0));
          TRY(head_bwd_rows(k == 0 ? 0 : f_cut[k - 1], f_cut[k], sh, true));
          head_bwd_t0 = f_cut[k];
        }
      }
      if (fshadow) head_t0 = f_cut[nf - 1];
      else head_bwd_t0 = 0;
    }
    fwd_shadow_used = fshadow;
  }
  if (v3_done) {
  } else if (cc.C) {
    DecFwd2P p2{};
    p2.B = B; p2.T = Tt; p2.Ti = d->Ti; p2.M = M; p2.H = H; p2.cond = d->conditional_attention;
    p2.vT = p.vT; p2.vV = p.vV; p2.bc = p.bc;
    p2.KT = p.KT; p2.KV = p.KV; p2.cmd_len = cmd_len; p2.h_init = p.h_init; p2.c_init = p.c_init; p2.Xe = p.Xe;
    p2.U = p.U; p2.Cs = p.Cs; p2.gates = p.gates; p2.alpha = p.alpha; p2.beta = p.beta;
    p2.Qp = p.Qp; p2.qT = p.qT; p2.qV = p.qV; p2.beta_sum = p.beta_sum;
    TRY(launch_dec_fwd_cluster(*d, P, ws, L, cc, p2, st));
  } else {
    TRY(launch_dec_fwd(*d, p, false, st));
  }
  // every byte of the tag word = version of the sweep that produced the saved activations of this workspace
  TRYCUDA(cudaMemsetAsync(ws + L.tag, v3_done ? 3 : (cc.C ? 2 : 1), sizeof(unsigned int), st));
  prof_mark(3, st);
  chain_mark("m:sweep_done", st);
  // output projection for all steps at once, then log-softmax.  (Chunks of it in the shadow of the forward sweep, the
  // way the backward pass does it, were a net loss: DESIGN.md 4.5 - and the signal code alone cost the sweep 33 us.)
  TRY(head_rows(head_t0, Tt, st));
  if (fwd_shadow_used) TRY(join_side(S, 2, st));
  if (early_head) {
    if (d_logp_ready) TRYCUDA(cudaStreamWaitEvent(st, d_logp_ready, 0));
    TRY(head_bwd_rows(head_bwd_t0, Tt, st, false));
    early_head_set(ws, d_logp_early);
  }
  if (d->auxiliary_task) {
    row_logsoftmax_kernel<<<ceil_div(B, 8), 256, 0, st>>>(ws + L.beta_sum, M, B, ws + L.aux_logp);
    GSCAN_CHECK_LAUNCH();
    TRYCUDA(cudaMemcpyAsync(aux_logp, ws + L.aux_logp, sizeof(float) * (size_t)B * M, cudaMemcpyDeviceToDevice, st));
  }
  prof_mark(4, st);
  chain_mark("m:fwd_end", st);
  chain_report("forward");
  return GSCAN_OK;
}

}  // namespace

extern "C" {

int gscan_backward(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                   const float* situations, const int64_t* targets, const float* drop_cnn, const float* drop_enc,
                   const float* drop_dec, float* ws, size_t ws_floats, const float* d_logp, const float* d_aux_logp,
                   float* const* G, void* stream) {
  return backward_impl(d, P, commands, cmd_len, situations, targets, DropSrc(drop_cnn), DropSrc(drop_enc), DropSrc(drop_dec), ws,
                       ws_floats, d_logp, d_aux_logp, G, stream);
}

}  // extern "C"

namespace {

int backward_impl(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                  const float* situations, const int64_t* targets, DropSrc drop_cnn, DropSrc drop_enc, DropSrc drop_dec,
                  float* ws, size_t ws_floats, const float* d_logp, const float* d_aux_logp, float* const* G, void* stream) {
  TRY(check_common(d, P));
  if (!commands || !cmd_len || !situations || !targets || !ws || !d_logp || !G) return GSCAN_E_BADARG;
  for (int i = 0; i < GSCAN_NUM_PARAMS; ++i) {
    bool optional = (i == GSCAN_P_COND_W || i == GSCAN_P_COND_B) && !d->conditional_attention;
    if (!optional && !G[i]) return GSCAN_E_BADARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Layout L = make_layout(*d, true);
  if (ws_floats < L.total) return GSCAN_E_WORKSPACE;
  const int B = d->B, Ti = d->Ti, Tt = d->Tt, M = d->G * d->G, D = 3 * d->F, H = d->H, E = d->E, V = d->V;
  const int R = Tt * B;
  const int sms = num_sms();
  const long long* cmds = reinterpret_cast<const long long*>(commands);
  const long long* tgts = reinterpret_cast<const long long*>(targets);
  float* U0 = ws + L.U;
  float* U1 = ws + L.U + (size_t)B * 4 * H;

  SideStreams* S = side_streams();
  // helper stream 2 starts here: it zeroes the destinations of the weight-gradient GEMMs beside the output-head work
  // (the caller's stream waits for that before the sweep), then runs the shadow launches
  if (S) TRY(fork_side(S, 2, st));
  // B1: log-softmax backward, hidden_to_output
  prof_mark(5, st);
  const bool head_fused = V <= kHeadMaxV && H <= 128 && getenv("GSCAN_HEAD_UNFUSED") == nullptr;
  // gscan_forward_train has done B1 and B2 already for this d_logp: only the weight gradient has to reach the caller
  const bool head_done = early_head_take(ws, d_logp) && head_fused;
  if (head_done) {
    TRYCUDA(cudaMemcpyAsync(G[GSCAN_P_H2O_W], ws + L.dWh2o, sizeof(float) * (size_t)V * H, cudaMemcpyDeviceToDevice, st));
  } else if (head_fused) {
    // log-softmax backward, dpre and the hidden_to_output weight gradient in one pass over the rows
    TRYCUDA(cudaMemsetAsync(G[GSCAN_P_H2O_W], 0, sizeof(float) * (size_t)V * H, st));
    head_bwd_fused_kernel<<<min(ceil_div(R, 8), 2 * sms), 256, 2 * sizeof(float) * (size_t)V * H, st>>>(
        d_logp, ws + L.logp, ws + L.pre, P[GSCAN_P_H2O_W], H, V, B, Tt, ws + L.dpre, G[GSCAN_P_H2O_W]);
    GSCAN_CHECK_LAUNCH();
  } else {
    int blocks = min(ceil_div(R, 8), 8 * sms);
    logsoftmax_bwd_kernel<<<blocks, 256, 0, st>>>(d_logp, ws + L.logp, V, B, Tt, ws + L.dlogits);
    GSCAN_CHECK_LAUNCH();
    TRY(matmul_nn(ws + L.dlogits, V, P[GSCAN_P_H2O_W], H, ws + L.dpre, H, R, H, V, 0, st));
  }
  // B2: output_to_hidden.  The two output-head weight gradients are not needed by the sweep: they are issued after
  // it, on helper stream 1 beside the other post-sweep chains.  (Running them on a helper stream BEFORE the sweep
  // delays the cluster launch of the sweep behind the persistent GEMM: measured 1.11 -> 1.45 ms for the sweep.)
  if (!head_done) TRY(matmul_nn(ws + L.dpre, H, P[GSCAN_P_O2H_W], 4 * H, ws + L.dU, 4 * H, R, 4 * H, H, 0, st));
  // B3: auxiliary head
  const float* dbeta_aux = nullptr;
  if (d->auxiliary_task && d_aux_logp) {
    row_logsoftmax_bwd_kernel<<<ceil_div(B, 8), 256, 0, st>>>(d_aux_logp, ws + L.aux_logp, M, B, ws + L.dbeta_aux);
    GSCAN_CHECK_LAUNCH();
    dbeta_aux = ws + L.dbeta_aux;
  }
  // B4: reverse-time sweep
  TRYCUDA(cudaMemsetAsync(ws + L.dvec, 0, sizeof(float) * 2 * H, st));
  DecBwdP bp{};
  bp.B = B; bp.T = Tt; bp.Ti = Ti; bp.M = M; bp.H = H; bp.cond = d->conditional_attention;
  bp.W_ih = P[GSCAN_P_DEC_WIH]; bp.W_hh = P[GSCAN_P_DEC_WHH]; bp.W_qV = P[GSCAN_P_VIS_QUERY_W];
  bp.W_c = P[GSCAN_P_COND_W]; bp.W_qT = P[GSCAN_P_TXT_QUERY_W];
  bp.vT = P[GSCAN_P_TXT_ENERGY_W]; bp.vV = P[GSCAN_P_VIS_ENERGY_W];
  bp.KT = ws + L.KT; bp.KV = ws + L.KV; bp.cmd_len = cmd_len;
  bp.Cs = ws + L.Cs; bp.gates = ws + L.gates; bp.alpha = ws + L.alpha; bp.beta = ws + L.beta;
  bp.Qp = ws + L.Qp; bp.qT = ws + L.qT; bp.qV = ws + L.qV;
  bp.dU = ws + L.dU; bp.dbeta_aux = dbeta_aux;
  bp.dgates = ws + L.dgates; bp.dd = ws + L.dd; bp.dqV = ws + L.dqV; bp.dqT = ws + L.dqT;
  bp.dKT = ws + L.dKT; bp.dKV = ws + L.dKV; bp.dh0 = ws + L.dh0;
  bp.dvT = ws + L.dvec; bp.dvV = ws + L.dvec + H;
  // decoder weight gradients (and the output_to_hidden one) as "TN" products over all R = Tt*B rows, one grouped
  // tcgen05 launch; the list is needed before the sweep for the shadow schedule below
  const float* Hprev = U0 + H;       // rows t*B+b hold h_{t-1}
  tc::GroupProblem gp[tc::MAXG];
  int ngp = 0;
  gp[ngp++] = {ws + L.dgates, 4 * H, U1, 4 * H, G[GSCAN_P_DEC_WIH], 3 * H, 4 * H, H};
  gp[ngp++] = {ws + L.dgates, 4 * H, U1 + 2 * H, 4 * H, G[GSCAN_P_DEC_WIH] + H, 3 * H, 4 * H, 2 * H};
  gp[ngp++] = {ws + L.dgates, 4 * H, Hprev, 4 * H, G[GSCAN_P_DEC_WHH], H, 4 * H, H};
  gp[ngp++] = {ws + L.dpre, H, U1, 4 * H, G[GSCAN_P_O2H_W], 4 * H, H, 4 * H};
  gp[ngp++] = {ws + L.dqT, H, Hprev, 4 * H, G[GSCAN_P_TXT_QUERY_W], H, H, H};
  gp[ngp++] = {ws + L.dqV, H, ws + L.Qp, H, G[GSCAN_P_VIS_QUERY_W], H, H, H};
  if (d->conditional_attention) {
    gp[ngp++] = {ws + L.dd, H, Hprev, 4 * H, G[GSCAN_P_COND_W], 2 * H, H, H};
    gp[ngp++] = {ws + L.dd, H, U1 + 2 * H, 4 * H, G[GSCAN_P_COND_W] + H, 2 * H, H, H};
  }
  // Shadow schedule: the sweep occupies 125 of the 148 SMs for ~1.1 ms and produces the rows of the GEMM operands
  // from the last step backwards.  Once every CTA has published the rows t >= t_sig (a counter in the workspace,
  // awaited by helper stream 2 through cuStreamWaitValue32) the grouped weight-gradient GEMM over THOSE rows runs on
  // the idle SMs beside the rest of the sweep; only the rows t < t_sig are left for after it.  Both partial launches
  // add into the destinations (zeroed before the sweep) with split-K atomics.
  const int sweep_ctas = ceil_div(B, v3::kNB) * v3::kC;
  const int idle_sms = sms - sweep_ctas;
  // cut points in percent of Tt, descending: chunk k = steps [cut_k, cut_{k-1}) is launched at signal k.  Measured at
  // B = 200: one signal at 45 % ends 24 us before the sweep does; three signals put 85 % of the rows into the shadow.
  int t_cut[4];
  int n_cut = 0;
  {
    const char* spec = getenv("GSCAN_SHADOW_CUTS");
    if (!spec) spec = "70,40,15";
    int prev = Tt;
    for (const char* q = spec; *q && n_cut < 4;) {
      const int pct = atoi(q);
      const int t = pct * Tt / 100;
      if (pct > 0 && pct < 100 && t > 0 && t < prev) { t_cut[n_cut++] = t; prev = t; }
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
  }
  bool shadow = S && v3_bwd_shape_ok(*d) && stream_wait_value_fn() && env_int("GSCAN_SHADOW", 1) != 0 && Tt >= 16 &&
                idle_sms >= 8 && n_cut > 0;
  for (int k = 0; shadow && k <= n_cut; ++k) {   // every chunk, and what is left for after the sweep, must take the tcgen05 path
    const int rows = ((k == 0 ? Tt : t_cut[k - 1]) - (k == n_cut ? 0 : t_cut[k])) * B;
    for (int i = 0; shadow && i < ngp; ++i) shadow = rows >= 1024 && tc::group_eligible(gp[i], rows);
  }
  // one launch of the value-path Z kernel (decoder_v3_bwd.cuh) over the steps [t0, t1)
  const int zNC = (d->conditional_attention ? 6 : 5) * H;
  auto launch_z = [&](int t0, int t1, bool accumulate, cudaStream_t s_, bool keep_off_sweep_sms = false) -> int {
    if (t1 <= t0) return 0;
    v3::ValueZP zp{};
    zp.dgates = ws + L.dgates; zp.dpre = ws + L.dpre; zp.dd = ws + L.dd;
    zp.alpha = ws + L.alpha; zp.beta = ws + L.beta;
    zp.B = B; zp.T = Tt; zp.Ti = Ti; zp.H = H; zp.NC = zNC;
    zp.ZV = ws + L.ZV; zp.ZT = ws + L.ZT; zp.ldv = 5 * H; zp.ldt = zNC;
    zp.t_begin = t0; zp.t_end = t1; zp.accumulate = accumulate ? 1 : 0;
    size_t smem = v3::value_z_smem_bytes(t1 - t0, Ti);
    // a shadow launch must not become resident next to a sweep CTA (194 KB of the SM's 227 KB): ask for at least 48 KB
    if (keep_off_sweep_sms && smem < 48 * 1024) smem = 48 * 1024;
    // tensor-core form (GSCAN_Z_MMA=0: the FFMA2 kernel): 36 + Ti weight rows in 3 or 4 m-tiles of 16
    static const bool z_mma = env_int("GSCAN_Z_MMA", 1) != 0;
    if (z_mma && H == v3::kH && M == v3::kM) {
      const int MT = (v3::kM + Ti + 15) / 16;
      size_t zsm = v3::value_zm_smem_bytes(t1 - t0, MT);
      if (keep_off_sweep_sms && zsm < 48 * 1024) zsm = 48 * 1024;
      if (MT <= 4 && zsm <= kMaxSmemBytes) {
        const dim3 g3(B, 3);
        if (MT <= 3) {
          if (zsm > 48 * 1024) TRY(set_smem(v3::attn_value_zm_kernel<3>, zsm));
          v3::attn_value_zm_kernel<3><<<g3, v3::kZmThreads, zsm, s_>>>(zp);
        } else {
          if (zsm > 48 * 1024) TRY(set_smem(v3::attn_value_zm_kernel<4>, zsm));
          v3::attn_value_zm_kernel<4><<<g3, v3::kZmThreads, zsm, s_>>>(zp);
        }
        GSCAN_CHECK_LAUNCH();
        return 0;
      }
    }
    const dim3 zgrid(B, ceil_div(zNC / 4, 64));
    if (v3::value_z_qw(Ti) == 3) {
      if (v3::value_z_smem_bytes(Tt, Ti) > 48 * 1024) TRY(set_smem(v3::attn_value_z_kernel<3>, v3::value_z_smem_bytes(Tt, Ti)));
      v3::attn_value_z_kernel<3><<<zgrid, 256, smem, s_>>>(zp);
    } else {
      if (v3::value_z_smem_bytes(Tt, Ti) > 48 * 1024) TRY(set_smem(v3::attn_value_z_kernel<4>, v3::value_z_smem_bytes(Tt, Ti)));
      v3::attn_value_z_kernel<4><<<zgrid, 256, smem, s_>>>(zp);
    }
    GSCAN_CHECK_LAUNCH();
    return 0;
  };
  // What goes into the shadow: the grouped GEMM of every chunk.  GSCAN_SHADOW_Z=1 also puts the Z kernel of every
  // chunk there (it heads the chain everything after the sweep waits for) and the grouped GEMM of only the first
  // GSCAN_SHADOW_GROUP_CHUNKS chunks - measured a LOSS (2.975 ms/step with 2 GEMM chunks, 2.884 with 1, against 2.848
  // without): on 23 SMs the latency-bound Z launches run past the end of the sweep and the last partial sum, which
  // the whole post-sweep chain waits for, arrives later than the single full-chip launch does.  Off by default.
  const bool shadow_z = env_int("GSCAN_SHADOW_Z", 0) != 0;
  const int group_chunks = shadow_z ? min(n_cut, env_int("GSCAN_SHADOW_GROUP_CHUNKS", 2)) : n_cut;
  unsigned int* progress = reinterpret_cast<unsigned int*>(ws + L.progress);
  if (shadow) {
    cudaStream_t z = S->s[2];     // forked at the top of the call, NOT from the end of the sweep
    TRYCUDA(cudaMemsetAsync(progress, 0, 4 * sizeof(unsigned int), z));
    TRY(tc::launch_group_zero(gp, ngp, z));
    TRYCUDA(cudaEventRecord(S->aux_ev[1], z));
    TRYCUDA(cudaStreamWaitEvent(st, S->aux_ev[1], 0));   // before the sweep (and the post-sweep group launch)
  }
  prof_mark(6, st);
  bool bwd_v3 = false;
  if (v3_bwd_shape_ok(*d)) {
    v3::DecBwd3P b3{};
    if (shadow) {
      b3.progress = progress;
      b3.n_signals = n_cut;
      for (int k = 0; k < n_cut; ++k) b3.t_signal[k] = t_cut[k];
    }
    b3.fwd_tag = reinterpret_cast<const unsigned int*>(ws + L.tag);
    b3.B = B; b3.T = Tt; b3.Ti = Ti;
    b3.W_ih = bp.W_ih; b3.W_hh = bp.W_hh; b3.W_qV = bp.W_qV; b3.W_c = bp.W_c; b3.W_qT = bp.W_qT;
    b3.PT = ws + L.PT;   // computed by the forward call on this workspace
    b3.vT = bp.vT; b3.vV = bp.vV; b3.KT = bp.KT; b3.KV = bp.KV; b3.cmd_len = cmd_len;
    b3.Cs = bp.Cs; b3.gates = bp.gates; b3.alpha = bp.alpha; b3.beta = bp.beta; b3.Qp = bp.Qp; b3.qT = bp.qT; b3.qV = bp.qV;
    b3.dU = bp.dU; b3.dbeta_aux = dbeta_aux;
    b3.dgates = bp.dgates; b3.dd = bp.dd; b3.dqV = bp.dqV; b3.dqT = bp.dqT;
    b3.dKT = bp.dKT; b3.dKV = bp.dKV; b3.dh0 = bp.dh0; b3.dvT = bp.dvT; b3.dvV = bp.dvV;
    int rc = launch_dec_bwd_v3(*d, b3, st);
    if (rc == 0) bwd_v3 = true;
    else if (rc != GSCAN_E_UNSUPPORTED) return rc;
    if (shadow && !bwd_v3) {      // nobody will signal: undo (the zeroing is harmless)
      shadow = false;
      TRY(join_side(S, 2, st));
    }
    if (shadow) {
      cudaStream_t sh = S->s[2];
      tc::ScopedSmCap cap(idle_sms);
      for (int k = 0; k < n_cut; ++k) {
        if (stream_wait_value_fn()((CUstream)sh, (CUdeviceptr)(progress + k), (cuuint32_t)sweep_ctas, 0u /* GEQ */) !=
            CUDA_SUCCESS) {
          if (k > 0) {   // (a driver that accepted the first wait accepts the next.)  Nothing of this call may stay in
            join_side(S, 2, st);   // flight on a helper stream when it returns: the caller owns the workspace again
            return GSCAN_E_UNSUPPORTED;
          }
          shadow = false;                          // stream memory operations unavailable: everything after the sweep
          TRY(join_side(S, 2, st));
          break;
        }
        const int t0 = t_cut[k], t1 = k == 0 ? Tt : t_cut[k - 1];
        if (shadow_z) TRY(launch_z(t0, t1, k > 0, sh, true));
        if (k >= group_chunks) continue;
        tc::GroupProblem part[tc::MAXG];
        const size_t r0 = (size_t)t0 * B;
        for (int i = 0; i < ngp; ++i) {
          part[i] = gp[i];
          part[i].X += r0 * gp[i].ldx;
          part[i].Y += r0 * gp[i].ldy;
        }
        TRY(tc::launch_group_tn(part, ngp, (t1 - t0) * B, sh, false, idle_sms));   // units = tiles x idle SMs: even rounds
      }
      if (shadow && shadow_z) TRYCUDA(cudaEventRecord(S->join_ev[2], sh));   // partial Z sums of the steps >= the last cut
      chain_mark("s2:shadow_end", sh);
    }
  }
  // Three independent chains from here, joined before returning:
  //   caller's stream  decoder weight gradients, decoder embedding
  //   helper stream 0  value path of both attentions (dc_T, dc_V -> dK^V, dK^T), then visual keys -> CNN
  //   helper stream 1  (after the value path) textual keys -> initial state -> command encoder
  cudaStream_t sv = S ? S->s[0] : st, stx = S ? S->s[1] : st;
  bool wait_value_path = false;
  if (bwd_v3) {
    prof_mark(7, st);   // the sweep kernel alone; what follows counts as batched weight-gradient work
    chain_mark("sweep_end", st);
    if (S) TRY(fork_side(S, 0, st));
    // value path of both attentions, outside the recurrence and summed over time first (decoder_v3_bwd.cuh):
    // Z = sum_t w_t (x) [dgates | dpre | dd], then dK += Z . Wst
    {
      const int NC = (d->conditional_attention ? 6 : 5) * H;
      const int nst = 5 * H * H + NC * H;
      v3::value_weight_stack_kernel<<<ceil_div(nst, 256), 256, 0, sv>>>(
          P[GSCAN_P_DEC_WIH], P[GSCAN_P_O2H_W], d->conditional_attention ? P[GSCAN_P_COND_W] : nullptr, H, ws + L.WstV,
          ws + L.WstT);
      GSCAN_CHECK_LAUNCH();
      if (shadow && shadow_z) {   // the steps the shadow launches did not cover, added to their partial sums
        TRYCUDA(cudaStreamWaitEvent(sv, S->join_ev[2], 0));
        TRY(launch_z(0, t_cut[n_cut - 1], true, sv));
      } else {
        TRY(launch_z(0, Tt, false, sv));
      }
      chain_mark("s0:Z", sv);
      // split-K (atomic adds onto the key-path part the sweep stored): short K loops on more SMs
      TRY(launch_gemm(ws + L.ZV, 5 * H, 1, ws + L.WstV, H, 1, ws + L.dKV, H, B * M, H, 5 * H, nullptr, nullptr, 0, 0, 2, sv));
      TRY(launch_gemm(ws + L.ZT, NC, 1, ws + L.WstT, H, 1, ws + L.dKT, H, Ti * B, H, NC, nullptr, nullptr, 0, 0, 4, sv));
      chain_mark("s0:value_path", sv);
    }
    // helper stream 1: the output-head weight gradients right after the sweep, then (after the value path) the text chain
    if (S) TRY(fork_side(S, 1, st));
    if (!head_fused) TRY(launch_grad_gemm(ws + L.dlogits, V, ws + L.pre, H, G[GSCAN_P_H2O_W], H, V, H, R, sms, stx));
    if (S) {
      TRYCUDA(cudaEventRecord(S->join_ev[0], sv));   // value path done: dK^T, dK^V complete
      wait_value_path = true;
    }
  } else {
    TRY(launch_dec_bwd(*d, bp, st));
    prof_mark(7, st);
    if (S) {
      TRY(fork_side(S, 0, st));
      TRY(fork_side(S, 1, st));
    }
    if (!head_fused) TRY(launch_grad_gemm(ws + L.dlogits, V, ws + L.pre, H, G[GSCAN_P_H2O_W], H, V, H, R, sms, stx));
  }
  TRYCUDA(cudaMemcpyAsync(G[GSCAN_P_TXT_ENERGY_W], ws + L.dvec, sizeof(float) * H, cudaMemcpyDeviceToDevice, st));
  TRYCUDA(cudaMemcpyAsync(G[GSCAN_P_VIS_ENERGY_W], ws + L.dvec + H, sizeof(float) * H, cudaMemcpyDeviceToDevice, st));
  static const bool main_waits_value = env_int("GSCAN_MAIN_WAITS_VALUE", 1) != 0;
  // B5: decoder weight gradients: what the shadow launch left (rows t < t_sig), or everything
  {
    tc::ScopedSmCap cap(S ? cap_post() : 0);   // the helper chains (CNN, encoder) need SMs meanwhile
    // everything downstream on both helper streams hangs off the value path: it gets the chip first
    if (wait_value_path && main_waits_value) TRYCUDA(cudaStreamWaitEvent(st, S->join_ev[0], 0));
    if (shadow) TRY(tc::launch_group_tn(gp, ngp, t_cut[group_chunks - 1] * B, st, false, 0));
    else TRY(launch_grad_group(gp, ngp, R, sms, st));
  }
  chain_mark("m:group", st);
  // decoder embedding: dE = dU[:, :H] + dgates . W_ih[:, :H], then scatter by token
  {
    tc::ScopedSmCap cap(S ? cap_post() : 0);
    TRY(matmul_nn(ws + L.dgates, 4 * H, P[GSCAN_P_DEC_WIH], 3 * H, ws + L.dU, 4 * H, R, H, 4 * H, 1, st));
  }
  {
    TRYCUDA(cudaMemsetAsync(G[GSCAN_P_DEC_EMB], 0, sizeof(float) * (size_t)V * H, st));
    const size_t tab = (size_t)embed_bwd_groups(H, 256) * V * H * sizeof(float);
    int use_smem = tab > 0 && tab <= 48 * 1024;
    int rpb = 64;
    embed_bwd_kernel<<<ceil_div(R, rpb), 256, use_smem ? tab : 0, st>>>(
        tgts, Tt, ws + L.dU, 4 * H, drop_dec, G[GSCAN_P_DEC_EMB], H, V, d->pad_idx_out, B, Tt, 1, rpb, use_smem);
    GSCAN_CHECK_LAUNCH();
  }
  // text-key weight gradient: dK^T is complete once the value path is, which this stream waited for when it can
  if (wait_value_path && !main_waits_value) TRYCUDA(cudaStreamWaitEvent(st, S->join_ev[0], 0));
  TRY(launch_grad_gemm(ws + L.dKT, H, ws + L.enc_out, H, G[GSCAN_P_TXT_KEY_W], H, H, H, Ti * B, sms, st));
  chain_mark("m:dE_embed", st);
  prof_mark(8, st);
  // B6: visual keys -> CNN (helper stream 0)
  TRY(launch_grad_gemm(ws + L.dKV, H, ws + L.feat, D, G[GSCAN_P_VIS_KEY_W], D, H, D, B * M, sms, sv));
  TRY(matmul_nn(ws + L.dKV, H, P[GSCAN_P_VIS_KEY_W], D, ws + L.dfeat, D, B * M, D, H, 0, sv));
  {
    long n = (long)B * M * D;
    cnn_dact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, sv>>>(ws + L.dfeat, ws + L.feat, drop_cnn, ws + L.dconv, n);
    GSCAN_CHECK_LAUNCH();
    CnnShape cs{B, d->G, d->C, d->F, d->K3};
    TRYCUDA(cudaMemsetAsync(ws + L.dWt_cnn, 0, sizeof(float) * cs.wtotal(), sv));
    size_t smem = (size_t)B * 8;
    if (smem > 48 * 1024) TRY(set_smem(cnn_wgrad_kernel, smem));
    const int ysplit = max(1, min(4, ceil_div(M * D, 5 * 256)));
    cnn_wgrad_kernel<<<dim3(M * d->C, ysplit), 256, smem, sv>>>(cs, situations, ws + L.dconv, ws + L.dWt_cnn);
    GSCAN_CHECK_LAUNCH();
    cnn_relayout_kernel<<<ceil_div(cs.wtotal(), 256), 256, 0, sv>>>(cs, G[GSCAN_P_CONV1_W], G[GSCAN_P_CONV2_W],
                                                                   G[GSCAN_P_CONV3_W], ws + L.dWt_cnn, 0);
    GSCAN_CHECK_LAUNCH();
    TRY(launch_colsum(ws + L.dconv, D, B * M, d->F, G[GSCAN_P_CONV1_B], sv));
    TRY(launch_colsum(ws + L.dconv + d->F, D, B * M, d->F, G[GSCAN_P_CONV2_B], sv));
    TRY(launch_colsum(ws + L.dconv + 2 * d->F, D, B * M, d->F, G[GSCAN_P_CONV3_B], sv));
  }
  // this chain also takes the bias column sums of the decoder (the caller's chain was the last to finish)
  TRY(launch_colsum(ws + L.dgates, 4 * H, R, 4 * H, G[GSCAN_P_DEC_BIH], sv));
  TRYCUDA(cudaMemcpyAsync(G[GSCAN_P_DEC_BHH], G[GSCAN_P_DEC_BIH], sizeof(float) * 4 * H, cudaMemcpyDeviceToDevice, sv));
  if (d->conditional_attention) TRY(launch_colsum(ws + L.dd, H, R, H, G[GSCAN_P_COND_B], sv));
  chain_mark("s0:kv_cnn", sv);
  // B7: initial state and textual keys (helper stream 1).  What needs only dh0 goes first; the products on dK^T wait
  // for the value path of helper stream 0.
  TRY(launch_tanh_bwd(ws + L.dh0, ws + L.h0, ws + L.dpre0, (long)B * H, stx));
  TRY(launch_grad_gemm(ws + L.dpre0, H, ws + L.h_enc, H, G[GSCAN_P_E2D_W], H, H, H, B, sms, stx));
  TRY(launch_colsum(ws + L.dpre0, H, B, H, G[GSCAN_P_E2D_B], stx));
  TRY(matmul_nn(ws + L.dpre0, H, P[GSCAN_P_E2D_W], H, ws + L.dh_enc, H, B, H, H, 0, stx));
  const int RE = B * Ti;
  TRYCUDA(cudaMemsetAsync(ws + L.denc_x, 0, sizeof(float) * (size_t)RE * E, stx));
  chain_mark("s1:h2o_e2d", stx);
  if (wait_value_path) TRYCUDA(cudaStreamWaitEvent(stx, S->join_ev[0], 0));
  TRY(matmul_nn(ws + L.dKT, H, P[GSCAN_P_TXT_KEY_W], H, ws + L.denc_out, H, Ti * B, H, H, 0, stx));
  // B8: encoder BPTT
  EncP ep{};
  ep.B = B; ep.Ti = Ti; ep.H = H;
  ep.W_hh[0] = P[GSCAN_P_ENC_WHH]; ep.W_hh[1] = P[GSCAN_P_ENC_WHH_R];
  for (int i = 0; i < 2; ++i) {
    ep.enc_h[i] = ws + L.enc_h[i];
    ep.enc_c[i] = ws + L.enc_c[i];
    ep.enc_g[i] = ws + L.enc_g[i];
    ep.dga[i] = ws + L.dga[i];
    ep.hprev[i] = ws + L.hprev[i];
  }
  ep.len = cmd_len;
  ep.denc_out = ws + L.denc_out;
  ep.dh_enc = ws + L.dh_enc;
  TRY(launch_enc(*d, ep, true, stx));
  chain_mark("s1:enc_bwd", stx);
  // From here the chain splits: the embedding gradient stays on helper stream 1, the encoder weight gradients (which
  // nothing waits for) move to helper stream 2, idle since its shadow launches ended.  (One stream: joined at +423 us
  // after the sweep, 57 us after the caller's own chain.)
  cudaStream_t sw = stx;
  static const bool split_tail = env_int("GSCAN_SPLIT_ENC_TAIL", 1) != 0;
  if (S && split_tail) {
    sw = S->s[2];
    TRYCUDA(cudaEventRecord(S->fork_ev[2], stx));
    TRYCUDA(cudaStreamWaitEvent(sw, S->fork_ev[2], 0));
  }
  const int wih[2] = {GSCAN_P_ENC_WIH, GSCAN_P_ENC_WIH_R}, whh[2] = {GSCAN_P_ENC_WHH, GSCAN_P_ENC_WHH_R};
  const int bih[2] = {GSCAN_P_ENC_BIH, GSCAN_P_ENC_BIH_R}, bhh[2] = {GSCAN_P_ENC_BHH, GSCAN_P_ENC_BHH_R};
  // embedding gradient first (it is what the rest of this chain waits for): both directions add into the zeroed
  // denc_x through split-K atomics, so neither waits for the other and each fills 8x more CTAs than one K-pass
  for (int i = 0; i < 2; ++i)
    TRY(launch_gemm(ws + L.dga[i], 4 * H, 1, P[wih[i]], E, 1, ws + L.denc_x, E, RE, E, 4 * H, nullptr, nullptr, 0, 0, 8, stx));
  {
    TRYCUDA(cudaMemsetAsync(G[GSCAN_P_ENC_EMB], 0, sizeof(float) * (size_t)d->Vi * E, stx));
    const size_t tab = (size_t)embed_bwd_groups(E, 256) * d->Vi * E * sizeof(float);
    int use_smem = tab > 0 && tab <= 48 * 1024;
    int rpb = 64;
    embed_bwd_kernel<<<ceil_div(RE, rpb), 256, use_smem ? tab : 0, stx>>>(
        cmds, d->Ti_stride, ws + L.denc_x, E, drop_enc, G[GSCAN_P_ENC_EMB], E, d->Vi, d->pad_idx_in, B, Ti, 0, rpb,
        use_smem);
    GSCAN_CHECK_LAUNCH();
  }
  {
    tc::GroupProblem gp[5];
    int n = 0;
    for (int i = 0; i < 2; ++i) {
      gp[n++] = {ws + L.dga[i], 4 * H, ws + L.hprev[i], H, G[whh[i]], H, 4 * H, H};
      gp[n++] = {ws + L.dga[i], 4 * H, ws + L.enc_x, E, G[wih[i]], E, 4 * H, E};
    }
    TRY(launch_grad_group(gp, n, RE, sms, sw));
    for (int i = 0; i < 2; ++i) {
      TRY(launch_colsum(ws + L.dga[i], 4 * H, RE, 4 * H, G[bih[i]], sw));
      TRYCUDA(cudaMemcpyAsync(G[bhh[i]], G[bih[i]], sizeof(float) * 4 * H, cudaMemcpyDeviceToDevice, sw));
    }
  }
  chain_mark("enc_wgrad", sw);
  chain_mark("s1:end", stx);
  if (S) {
    TRY(join_side(S, 0, st));
    TRY(join_side(S, 1, st));
    TRY(join_side(S, 2, st));   // forked at the top of the call
  }
  prof_mark(9, st);
  chain_mark("joined", st);
  chain_report("backward after the sweep");
  return GSCAN_OK;
}

}  // namespace

extern "C" {

int gscan_encode(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                 const float* situations, const float* drop_cnn, const float* drop_enc, float* ws, size_t ws_floats,
                 float* feat, float* enc_out, float* hidden, void* stream) {
  TRY(check_common(d, P));
  if (!commands || !cmd_len || !situations || !ws || !feat || !enc_out || !hidden) return GSCAN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  gscan_dims e = *d;
  e.Tt = 1;
  Layout L = make_layout(e, false);
  if (ws_floats < L.total) return GSCAN_E_WORKSPACE;
  const size_t B = d->B, Ti = d->Ti, M = (size_t)d->G * d->G, D = 3 * (size_t)d->F, H = d->H;
  TRY(run_encoder_side(e, P, reinterpret_cast<const long long*>(commands), cmd_len, situations, drop_cnn, drop_enc,
                       ws, L, false, st));
  TRYCUDA(cudaMemcpyAsync(feat, ws + L.feat, sizeof(float) * B * M * D, cudaMemcpyDeviceToDevice, st));
  TRYCUDA(cudaMemcpyAsync(enc_out, ws + L.enc_out, sizeof(float) * Ti * B * H, cudaMemcpyDeviceToDevice, st));
  TRYCUDA(cudaMemcpyAsync(hidden, ws + L.h_enc, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, st));
  return GSCAN_OK;
}

int gscan_decoder_step(const gscan_dims* d, const float* const* P, const int64_t* tokens, const float* h_in,
                       const float* c_in, const float* keys_text, const int32_t* cmd_len, const float* keys_vis,
                       const float* drop_dec, float* ws, size_t ws_floats, float* logits, float* h_out, float* c_out,
                       float* alpha, float* beta, void* stream) {
  TRY(check_common(d, P));
  if (!tokens || !h_in || !c_in || !keys_text || !cmd_len || !keys_vis || !ws || !logits || !h_out || !c_out)
    return GSCAN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  gscan_dims e = *d;
  e.Tt = 1;
  Layout L = make_layout(e, false);
  if (ws_floats < L.total) return GSCAN_E_WORKSPACE;
  const int B = d->B, H = d->H, V = d->V;
  TRY(pack_decoder_weights(e, P, ws, L, st));
  {
    long n = (long)B * H;
    embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const long long*>(tokens), 1,
                                                              P[GSCAN_P_DEC_EMB], H, drop_dec, ws + L.U, 4 * H, B, 1, 1);
    GSCAN_CHECK_LAUNCH();
  }
  float* U1 = ws + L.U + (size_t)B * 4 * H;
  TRY(linear(U1, 4 * H, P[GSCAN_P_DEC_WIH], 3 * H, ws + L.Xe, 4 * H, B, 4 * H, H, P[GSCAN_P_DEC_BIH],
             P[GSCAN_P_DEC_BHH], 0, st));
  DecFwdP p{};
  fill_dec_fwd_common(e, P, ws, L, p);
  p.T = 1;
  p.KT = keys_text; p.KV = keys_vis; p.cmd_len = cmd_len;
  p.h_init = h_in; p.c_init = c_in;
  p.Xe = ws + L.Xe;
  p.U = ws + L.U;
  p.alpha = alpha; p.beta = beta;
  p.h_out = h_out; p.c_out = c_out;
  TRY(launch_dec_fwd(e, p, false, st));
  TRY(linear(U1, 4 * H, P[GSCAN_P_O2H_W], 4 * H, ws + L.pre, H, B, H, 4 * H, nullptr, nullptr, 0, st));
  size_t smem = (size_t)V * (H + 1) * sizeof(float);
  if (smem > 48 * 1024) TRY(set_smem(out_logsoftmax_kernel, smem));
  out_logsoftmax_kernel<<<min(ceil_div(B, 8), 8 * num_sms()), 256, smem, st>>>(ws + L.pre, P[GSCAN_P_H2O_W], H, V, B, 1,
                                                                               nullptr, logits);
  GSCAN_CHECK_LAUNCH();
  return GSCAN_OK;
}

int gscan_greedy_decode(const gscan_dims* d, const float* const* P, const int64_t* commands, const int32_t* cmd_len,
                        const float* situations, int32_t max_decoding_steps, int32_t sos, int32_t eos, float* ws,
                        size_t ws_floats, int64_t* out_tokens, int32_t* out_len, int32_t* out_steps, float* beta_sum,
                        float* aux_logp, float* alphas, float* betas, void* stream) {
  TRY(check_common(d, P));
  if (!commands || !cmd_len || !situations || !ws || !out_tokens || !out_len || !out_steps || !beta_sum)
    return GSCAN_E_BADARG;
  if (max_decoding_steps < 0 || d->V > 32) return GSCAN_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  gscan_dims e = *d;
  e.Tt = 1;
  Layout L = make_layout(e, false);
  if (ws_floats < gscan_greedy_workspace_floats(d)) return GSCAN_E_WORKSPACE;
  const int B = d->B, M = d->G * d->G, H = d->H, V = d->V, Vp = pad4(V);
  const int T = max_decoding_steps + 1;
  float* XeTab = ws + L.total;
  float* Wout = XeTab + pad4(V * 4 * H);
  float* OutE = Wout + pad4(V * 4 * H);
  float* Wo_t = OutE + pad4(V * V);
  // everything that depends on the weights only (packed decoder weights, eval-mode tables, the token buffer fill)
  // runs on helper stream 1 beside the encoder side
  SideStreams* S = side_streams();
  cudaStream_t sp = S ? S->s[1] : st;
  if (S) TRY(fork_side(S, 1, st));
  TRY(pack_decoder_weights(e, P, ws, L, sp));
  // eval-mode tables: the embedding-dependent parts of the gates and of the logits have only V rows
  TRY(linear(P[GSCAN_P_DEC_EMB], H, P[GSCAN_P_DEC_WIH], 3 * H, XeTab, 4 * H, V, 4 * H, H, P[GSCAN_P_DEC_BIH],
             P[GSCAN_P_DEC_BHH], 0, sp));
  TRY(matmul_nn(P[GSCAN_P_H2O_W], H, P[GSCAN_P_O2H_W], 4 * H, Wout, 4 * H, V, 4 * H, H, 0, sp));
  TRY(linear(P[GSCAN_P_DEC_EMB], H, Wout, 4 * H, OutE, V, V, V, H, nullptr, nullptr, 0, sp));
  TRYCUDA(cudaMemsetAsync(Wo_t, 0, sizeof(float) * 3 * H * Vp, sp));
  {
    PackTable tab;
    tab.n = 1;
    tab.d[0] = PackDesc{Wout, 4 * H, H, Wo_t, Vp, 0, V, 3 * H};
    TRY(launch_pack(tab, sp));
  }
  {
    long n = (long)B * T;
    fill_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, sp>>>(reinterpret_cast<long long*>(out_tokens), n, -1);
    GSCAN_CHECK_LAUNCH();
  }
  TRY(run_encoder_side(e, P, reinterpret_cast<const long long*>(commands), cmd_len, situations, nullptr, nullptr, ws,
                       L, true, st));
  if (S) TRY(join_side(S, 1, st));
  DecFwdP p{};
  fill_dec_fwd_common(e, P, ws, L, p);
  p.T = T;
  p.KT = ws + L.KT; p.KV = ws + L.KV; p.cmd_len = cmd_len;
  p.h_init = ws + L.h0; p.c_init = ws + L.h0;
  p.beta_sum = beta_sum;
  p.XeTab = XeTab; p.OutE = OutE; p.Wo_t = Wo_t; p.Vp = Vp; p.sos = sos; p.eos = eos;
  p.out_tokens = reinterpret_cast<long long*>(out_tokens);
  p.out_len = out_len; p.out_steps = out_steps;
  p.g_alphas = alphas; p.g_betas = betas;
  bool greedy_v3 = false;
  if (v3_shape_ok(e) && V <= 32 && !getenv("GSCAN_GREEDY_V1")) {
    v3::DecFwd3P p3{};
    p3.B = B; p3.T = T; p3.Ti = e.Ti;
    p3.vT = p.vT; p3.vV = p.vV; p3.bc = p.bc;
    p3.KT = p.KT; p3.KV = p.KV; p3.cmd_len = cmd_len; p3.h_init = p.h_init; p3.c_init = p.c_init;
    p3.beta_sum = beta_sum;
    p3.XeTab = XeTab; p3.OutE = OutE; p3.Wo_t = Wo_t; p3.V = V; p3.Vp = Vp; p3.sos = sos; p3.eos = eos;
    p3.out_tokens = p.out_tokens; p3.out_len = out_len; p3.out_steps = out_steps;
    p3.g_alphas = alphas; p3.g_betas = betas;
    int rc = launch_dec_fwd_v3(e, P, ws, L, p3, true, st);
    if (rc == 0) greedy_v3 = true;
    else if (rc != GSCAN_E_UNSUPPORTED) return rc;
  }
  if (!greedy_v3) TRY(launch_dec_fwd(e, p, true, st));
  if (aux_logp) {
    row_logsoftmax_kernel<<<ceil_div(B, 8), 256, 0, st>>>(beta_sum, M, B, aux_logp);
    GSCAN_CHECK_LAUNCH();
  }
  return GSCAN_OK;
}

int gscan_nll_forward(const float* logp, const int64_t* targets, int32_t B, int32_t Tt, int32_t V, int32_t pad_idx,
                      int32_t shift, float* loss_out, void* stream) {
  if (!logp || !targets || !loss_out || B < 1 || Tt < 1 || V < 1 || shift < 0) return GSCAN_E_BADARG;
  static_assert(kNllBlocks <= 32 && 2 + 2 * kNllBlocks <= GSCAN_NLL_OUT_FLOATS, "scratch behind loss_out");
  nll_forward_kernel<<<kNllBlocks, 256, 0, (cudaStream_t)stream>>>(logp, reinterpret_cast<const long long*>(targets), B, Tt,
                                                                  V, pad_idx, shift, loss_out + 2);
  GSCAN_CHECK_LAUNCH();
  nll_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_out, kNllBlocks);
  GSCAN_CHECK_LAUNCH();
  return GSCAN_OK;
}

int gscan_nll_count(const int64_t* targets, int32_t B, int32_t Tt, int32_t pad_idx, int32_t shift, float* loss_out,
                    void* stream) {
  if (!targets || !loss_out || shift < 0 || B < 1 || Tt < 1) return GSCAN_E_BADARG;
  nll_count_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(targets), B, Tt, pad_idx, shift,
                                                         loss_out);
  GSCAN_CHECK_LAUNCH();
  return GSCAN_OK;
}

int gscan_nll_backward(const int64_t* targets, int32_t B, int32_t Tt, int32_t V, int32_t pad_idx, int32_t shift,
                       const float* loss_out, const float* d_loss, float* d_logp, void* stream) {
  if (!targets || !loss_out || !d_loss || !d_logp || shift < 0) return GSCAN_E_BADARG;
  long n = (long)B * Tt * V;
  nll_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(targets), B, Tt, V, pad_idx, shift, loss_out, d_loss, d_logp);
  GSCAN_CHECK_LAUNCH();
  return GSCAN_OK;
}

int gscan_metrics(const float* logp, const int64_t* targets, int32_t B, int32_t Tt, int32_t V, int32_t pad_idx,
                  int32_t* counts, void* stream) {
  if (!logp || !targets || !counts) return GSCAN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  TRYCUDA(cudaMemsetAsync(counts, 0, 3 * sizeof(int32_t), st));
  metrics_kernel<<<ceil_div(B, 8), 256, 0, st>>>(logp, reinterpret_cast<const long long*>(targets), B, Tt, V, pad_idx,
                                                 counts);
  GSCAN_CHECK_LAUNCH();
  return GSCAN_OK;
}

int gscan_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                        float beta1, float beta2, float eps, int32_t step, float grad_scale, const float* grad_denom,
                        void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || step < 1) return GSCAN_E_BADARG;
  if (n == 0) return GSCAN_OK;
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr,
                                                                            beta1, beta2, eps, bc1, bc2s, grad_scale,
                                                                            grad_denom);
  GSCAN_CHECK_LAUNCH();
  return GSCAN_OK;
}

int gscan_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                    float beta1, float beta2, float eps, int32_t step, float grad_scale, void* stream) {
  return gscan_adam_step_dev(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale, nullptr,
                             stream);
}

int gscan_sgemm(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs, float* C,
                int64_t ldc, int32_t M, int32_t N, int32_t K, const float* bias, int32_t act, int32_t accumulate,
                void* stream) {
  if (!A || !B || !C || M < 0 || N < 0 || K < 0) return GSCAN_E_BADARG;
  return launch_gemm(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, bias, nullptr, act, accumulate, 1,
                     (cudaStream_t)stream);
}

int gscan_sgemm_path(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs, float* C,
                     int64_t ldc, int32_t M, int32_t N, int32_t K, const float* bias, int32_t act, int32_t accumulate,
                     int32_t ksplit, int32_t path, void* stream) {
  if (!A || !B || !C || M < 0 || N < 0 || K < 0 || ksplit < 1 || (path != 0 && path != 1)) return GSCAN_E_BADARG;
  if (ksplit > 1 && (bias || act)) return GSCAN_E_BADARG;
  if (path == 0)
    return launch_sgemm(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, bias, nullptr, act, accumulate, ksplit,
                        (cudaStream_t)stream);
  if (!tc::eligible(A, a_rs, a_cs, B, b_rs, b_cs, M, N, K)) return GSCAN_E_UNSUPPORTED;
  return tc::launch(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, bias, nullptr, act, accumulate, ksplit,
                    (cudaStream_t)stream);
}

int gscan_cnn_forward(const gscan_dims* d, const float* const* P, const float* situations, const float* drop_cnn,
                      float* ws, size_t ws_floats, float* feat, void* stream) {
  if (!d || !P || !situations || !ws || !feat) return GSCAN_E_BADARG;
  CnnShape cs{d->B, d->G, d->C, d->F, d->K3};
  if (ws_floats < (size_t)cs.wtotal()) return GSCAN_E_WORKSPACE;
  return run_cnn_forward(*d, P, situations, drop_cnn, ws, feat, (cudaStream_t)stream);
}

}  // extern "C"
