// Recurrent sweeps of the path: one launch runs ALL time steps (no per-step launches, no host
// syncs - the reference issues ~35 launches and 2 syncs per step, seq2seq_model.py:473-480).
//
// Work split: one CTA owns NB examples for the whole sequence.  Everything an example needs
// between steps (h, c, its projected keys K^T / K^V, attention scratch) stays in shared memory;
// the recurrent weight matrices (~640 KB fp32 for H=100, more than one SM's smem) are streamed
// from L2 every step through 128-bit read-only loads and consumed with packed fp32 FMAs (FFMA2).
// Each matrix-vector stage is split over output-row quads x K-slices so all 512 threads work,
// with partial sums combined through shared memory.
#pragma once
#include "common.cuh"

namespace gscan {

constexpr int kRecThreads = 512;

struct Bump {
  float* p;
  __device__ __forceinline__ float* take(int n) {
    float* r = p;
    p += (n + 3) & ~3;
    return r;
  }
};
__host__ __device__ __forceinline__ int pad4(int n) { return (n + 3) & ~3; }

// Number of K-slices for a stage with R output rows reduced over K, given nthreads.
__host__ __device__ __forceinline__ int matvec_splits(int R, int K, int nthreads) {
  int rq = R >> 2;
  int ks = nthreads / (rq > 0 ? rq : 1);
  int kmax = K / 8;
  if (ks > kmax) ks = kmax;
  if (ks > 32) ks = 32;
  return ks < 1 ? 1 : ks;
}

// part[(s*NB + n)*R + r] = sum_{k in slice s} Wt[k*ldw + r] * x_s[n*ldx + k]
// Wt is "reduction-major": row k holds the R outputs' weights contiguously (R % 4 == 0, 16B aligned).
template <int NB>
__device__ __forceinline__ void matvec_partial(const float* __restrict__ Wt, int ldw, int R, int K,
                                               const float* x_s, int ldx, float* part, int KS) {
  const int RQ = R >> 2;
  for (int item = threadIdx.x; item < RQ * KS; item += blockDim.x) {
    const int s = item / RQ, q = item - s * RQ;
    const int k0 = (K * s) / KS, k1 = (K * (s + 1)) / KS;
    float2 acc[NB][2];
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[n][0] = acc[n][1] = make_float2(0.f, 0.f);
    const float* wp = Wt + (long)k0 * ldw + 4 * q;
#pragma unroll 4
    for (int k = k0; k < k1; ++k, wp += ldw) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(wp));
      const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float xv = x_s[n * ldx + k];
        const float2 xx = make_float2(xv, xv);
        fma2(acc[n][0], w01, xx);
        fma2(acc[n][1], w23, xx);
      }
    }
#pragma unroll
    for (int n = 0; n < NB; ++n)
      *reinterpret_cast<float4*>(&part[(s * NB + n) * R + 4 * q]) =
          make_float4(acc[n][0].x, acc[n][0].y, acc[n][1].x, acc[n][1].y);
  }
}

template <int NB>
__device__ __forceinline__ float part_sum(const float* part, int KS, int R, int n, int r) {
  float v = 0.f;
  for (int s = 0; s < KS; ++s) v += part[(s * NB + n) * R + r];
  return v;
}

// scores[n*N + j] = v . tanh(q[n] + K[n][j]) for all (n, j); one warp per pair.
template <int NB>
__device__ __forceinline__ void attn_scores(const float* q_s, int ldq, const float* K_s, int N, int H,
                                            const float* v_s, const int* len_s, bool masked, float* sc_s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  for (int pidx = warp; pidx < NB * N; pidx += NW) {
    const int n = pidx / N, j = pidx - n * N;
    const float* kp = K_s + (long)pidx * H;
    const float* qp = q_s + n * ldq;
    float s = 0.f;
    for (int h = lane; h < H; h += 32) s = fmaf(v_s[h], act_tanh(qp[h] + kp[h]), s);
    s = warp_sum(s);
    if (lane == 0) sc_s[pidx] = (masked && j >= len_s[n]) ? -INFINITY : s;
  }
}

// In-place softmax of sc_s[n*N .. n*N+N) for n < NB, one warp per example.
template <int NB>
__device__ __forceinline__ void attn_softmax(float* sc_s, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < NB) {
    float* s = sc_s + warp * N;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, s[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      float e = __expf(s[j] - mx);
      s[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < N; j += 32) s[j] *= inv;
  }
}

// ---------------------------------------------------------------------------------------------
// Decoder forward sweep (reference seq2seq_model.py:359-428 per step, loop at 473-480;
// greedy variant: predict.py:97-117).
// ---------------------------------------------------------------------------------------------
struct DecFwdP {
  int B, T, Ti, M, H, V, cond;
  // reduction-major (transposed) weight packs, see pack_decoder_weights in gscan_api.cu
  const float* WA_t; int RA;   // [H][RA]  rows: W_qT | (W_c[:, :H]) | W_hh
  const float* WB_t; int RB;   // [H][RB]  rows: (W_c[:, H:]) | W_ih[:, H:2H]
  const float* WC_t;           // [H][H]   W_qV
  const float* WD_t;           // [H][4H]  W_ih[:, 2H:3H]
  const float *vT, *vV, *bc;
  const float* KT;             // [Ti][B][H]
  const float* KV;             // [B][M][H]
  const int* cmd_len;          // [B]
  const float *h_init, *c_init;  // [B][H]
  const float* Xe;             // [T][B][4H]  teacher-forced input-gate pre-activations (incl. biases)
  // saved activations (any may be null)
  float* U;      // [(T+1)][B][4H]  row = [e | h | cT | cV]
  float* Cs;     // [(T+1)][B][H]
  float* gates;  // [T][B][4H]
  float* alpha;  // [T][B][Ti]
  float* beta;   // [T][B][M]
  float *Qp, *qT, *qV;  // [T][B][H]
  float* beta_sum;      // [B][M]
  float *h_out, *c_out; // [B][H] final state
  // greedy decoding
  const float* XeTab;   // [V][4H]   Emb . W_ih[:, :H]^T + b_ih + b_hh
  const float* OutE;    // [V][V]    OutE[tok][v] = Wout[v, :H] . Emb[tok]
  const float* Wo_t;    // [3H][Vp]  transposed Wout[:, H:4H], Wout = W_h2o . W_o2h
  int Vp, sos, eos;
  long long* out_tokens;  // [B][T]
  int *out_len, *out_steps;
  float *g_alphas, *g_betas;  // [B][T][Ti], [B][T][M]
};

template <int NB>
size_t dec_fwd_smem_floats(int Ti, int M, int H, int RA, int RB, int Vp, bool greedy, int nthreads) {
  size_t n = 0;
  n += pad4(NB * M * H) + pad4(NB * Ti * H);        // KV_s, KT_s
  n += pad4(NB * 3 * H) + pad4(NB * H);             // u_s, c_s
  n += 4 * pad4(NB * H) + pad4(NB * 4 * H);         // qT_s, ch_s, qp_s, qV_s, gacc_s
  n += pad4(NB * Ti) + 2 * pad4(NB * M);            // scT_s, scV_s, bsum_s
  n += 3 * pad4(H);                                 // vT_s, vV_s, bc_s
  size_t part = 0;
  auto upd = [&](int R, int K) { size_t v = (size_t)matvec_splits(R, K, nthreads) * NB * R; if (v > part) part = v; };
  upd(RA, H); upd(RB, H); upd(H, H); upd(4 * H, H);
  if (greedy) upd(Vp, 3 * H);
  n += pad4((int)part);
  n += 16;                                          // len_s, tok_s, alive_s, flags
  return n;
}

template <int NB, bool GREEDY>
__global__ void __launch_bounds__(kRecThreads, 1) decoder_fwd_kernel(DecFwdP p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, H3 = 3 * H, H4 = 4 * H, M = p.M, Ti = p.Ti, B = p.B;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int b0 = blockIdx.x * NB;
  const int nb = min(NB, B - b0);
  Bump bump{smem};
  float* KV_s = bump.take(NB * M * H);
  float* KT_s = bump.take(NB * Ti * H);
  float* u_s = bump.take(NB * H3);   // [h | cT | cV] per example
  float* c_s = bump.take(NB * H);
  float* qT_s = bump.take(NB * H);
  float* ch_s = bump.take(NB * H);
  float* qp_s = bump.take(NB * H);
  float* qV_s = bump.take(NB * H);
  float* gacc_s = bump.take(NB * H4);
  float* scT_s = bump.take(NB * Ti);
  float* scV_s = bump.take(NB * M);
  float* bsum_s = bump.take(NB * M);
  float* vT_s = bump.take(H);
  float* vV_s = bump.take(H);
  float* bc_s = bump.take(H);
  const int KSA = matvec_splits(p.RA, H, NT), KSB = matvec_splits(p.RB, H, NT);
  const int KSC = matvec_splits(H, H, NT), KSD = matvec_splits(H4, H, NT);
  const int KSE = GREEDY ? matvec_splits(p.Vp, H3, NT) : 1;
  float* part = bump.p;
  {
    int part_n = max(max(KSA * p.RA, KSB * p.RB), max(KSC * H, KSD * H4));
    if (GREEDY) part_n = max(part_n, KSE * p.Vp);
    bump.take(NB * part_n);
  }
  int* len_s = reinterpret_cast<int*>(bump.take(16));
  int* tok_s = len_s + NB;       // NB <= 4
  int* alive_s = tok_s + NB;
  int* flag_s = alive_s + NB;

  // ---- one-time loads -------------------------------------------------------------------
  for (int i = tid; i < NB * M * H; i += NT) {
    int n = i / (M * H);
    KV_s[i] = (n < nb) ? __ldg(p.KV + (long)b0 * M * H + i) : 0.f;
  }
  for (int i = tid; i < NB * Ti * H; i += NT) {
    int n = i / (Ti * H), r = i - n * Ti * H;
    int j = r / H, h = r - j * H;
    KT_s[i] = (n < nb) ? __ldg(p.KT + ((long)j * B + b0 + n) * H + h) : 0.f;
  }
  for (int i = tid; i < NB * H; i += NT) {
    int n = i / H, h = i - n * H;
    float hv = 0.f, cv = 0.f;
    if (n < nb) {
      hv = __ldg(p.h_init + (long)(b0 + n) * H + h);
      cv = __ldg(p.c_init + (long)(b0 + n) * H + h);
      if (p.U) p.U[(long)(b0 + n) * H4 + H + h] = hv;   // row group 0 carries h_{-1}
      if (p.Cs) p.Cs[(long)(b0 + n) * H + h] = cv;
    }
    u_s[n * H3 + h] = hv;
    c_s[i] = cv;
  }
  for (int i = tid; i < NB * M; i += NT) bsum_s[i] = 0.f;
  for (int h = tid; h < H; h += NT) {
    vT_s[h] = __ldg(p.vT + h);
    vV_s[h] = __ldg(p.vV + h);
    bc_s[h] = p.cond ? __ldg(p.bc + h) : 0.f;
  }
  if (tid < NB) {
    int l = (tid < nb) ? p.cmd_len[b0 + tid] : 1;
    len_s[tid] = max(1, min(l, Ti));
    tok_s[tid] = p.sos;
    alive_s[tid] = (tid < nb) ? 1 : 0;
  }
  int my_len = 0, my_steps = 0;   // greedy bookkeeping, thread n < NB
  __syncthreads();

  for (int t = 0; t < p.T; ++t) {
    // ---- stage A: everything that depends only on h_{t-1} --------------------------------
    matvec_partial<NB>(p.WA_t, p.RA, p.RA, H, u_s, H3, part, KSA);
    __syncthreads();
    {
      const int goff = p.RA - H4;
      for (int i = tid; i < NB * p.RA; i += NT) {
        int n = i / p.RA, r = i - n * p.RA;
        float v = part_sum<NB>(part, KSA, p.RA, n, r);
        if (r < H) {
          qT_s[n * H + r] = v;
          if (p.qT && n < nb) p.qT[((long)t * B + b0 + n) * H + r] = v;
          if (!p.cond) qp_s[n * H + r] = u_s[n * H3 + r];
        } else if (r < goff) {
          ch_s[n * H + r - H] = v;
        } else {
          gacc_s[n * H4 + r - goff] = v;
        }
      }
    }
    __syncthreads();
    // ---- textual attention --------------------------------------------------------------
    attn_scores<NB>(qT_s, H, KT_s, Ti, H, vT_s, len_s, true, scT_s);
    __syncthreads();
    attn_softmax<NB>(scT_s, Ti);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      float c = 0.f;
      for (int j = 0; j < Ti; ++j) c = fmaf(scT_s[n * Ti + j], KT_s[(n * Ti + j) * H + h], c);
      u_s[n * H3 + H + h] = c;
      if (p.U && n < nb) p.U[((long)(t + 1) * B + b0 + n) * H4 + 2 * H + h] = c;
    }
    for (int i = tid; i < NB * Ti; i += NT) {
      int n = i / Ti, j = i - n * Ti;
      if (n < nb) {
        if (p.alpha) p.alpha[((long)t * B + b0 + n) * Ti + j] = scT_s[i];
        if (GREEDY && p.g_alphas && alive_s[n]) p.g_alphas[((long)(b0 + n) * p.T + t) * Ti + j] = scT_s[i];
      }
    }
    __syncthreads();
    // ---- stage B: everything that depends on c_T ------------------------------------------
    matvec_partial<NB>(p.WB_t, p.RB, p.RB, H, u_s + H, H3, part, KSB);
    __syncthreads();
    {
      const int goff = p.RB - H4;
      for (int i = tid; i < NB * p.RB; i += NT) {
        int n = i / p.RB, r = i - n * p.RB;
        float v = part_sum<NB>(part, KSB, p.RB, n, r);
        if (r < goff) {   // conditional query q' = tanh(W_c [h; cT] + b_c)
          float q = act_tanh(ch_s[n * H + r] + v + bc_s[r]);
          qp_s[n * H + r] = q;
        } else {
          gacc_s[n * H4 + r - goff] += v;
        }
      }
    }
    __syncthreads();
    if (p.Qp)
      for (int i = tid; i < NB * H; i += NT) {
        int n = i / H, h = i - n * H;
        if (n < nb) p.Qp[((long)t * B + b0 + n) * H + h] = qp_s[i];
      }
    // ---- stage C: visual query -----------------------------------------------------------
    matvec_partial<NB>(p.WC_t, H, H, H, qp_s, H, part, KSC);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      float v = part_sum<NB>(part, KSC, H, n, h);
      qV_s[i] = v;
      if (p.qV && n < nb) p.qV[((long)t * B + b0 + n) * H + h] = v;
    }
    __syncthreads();
    // ---- visual attention ------------------------------------------------------------------
    attn_scores<NB>(qV_s, H, KV_s, M, H, vV_s, len_s, false, scV_s);
    __syncthreads();
    attn_softmax<NB>(scV_s, M);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      float c = 0.f;
      for (int m = 0; m < M; ++m) c = fmaf(scV_s[n * M + m], KV_s[(n * M + m) * H + h], c);
      u_s[n * H3 + 2 * H + h] = c;
      if (p.U && n < nb) p.U[((long)(t + 1) * B + b0 + n) * H4 + 3 * H + h] = c;
    }
    for (int i = tid; i < NB * M; i += NT) {
      int n = i / M, m = i - n * M;
      if (n < nb) {
        float w = scV_s[i];
        if (p.beta) p.beta[((long)t * B + b0 + n) * M + m] = w;
        if (!GREEDY || alive_s[n]) bsum_s[i] += w;
        if (GREEDY && p.g_betas && alive_s[n]) p.g_betas[((long)(b0 + n) * p.T + t) * M + m] = w;
      }
    }
    __syncthreads();
    // ---- stage D: c_V contribution to the gates, then the LSTM cell ------------------------
    matvec_partial<NB>(p.WD_t, H4, H4, H, u_s + 2 * H, H3, part, KSD);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      float a[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        a[g] = gacc_s[n * H4 + g * H + h] + part_sum<NB>(part, KSD, H4, n, g * H + h);
        if (n < nb) {
          if (GREEDY) a[g] += __ldg(p.XeTab + (long)tok_s[n] * H4 + g * H + h);
          else a[g] += __ldg(p.Xe + ((long)t * B + b0 + n) * H4 + g * H + h);
        }
      }
      float ig = act_sigmoid(a[0]), fg = act_sigmoid(a[1]), gg = act_tanh(a[2]), og = act_sigmoid(a[3]);
      float cn = fmaf(fg, c_s[i], ig * gg);
      float hn = og * act_tanh(cn);
      c_s[i] = cn;
      u_s[n * H3 + h] = hn;
      if (n < nb) {
        long row = (long)t * B + b0 + n;
        if (p.gates) {
          float* gp = p.gates + row * H4 + h;
          gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
        }
        if (p.U) p.U[((long)(t + 1) * B + b0 + n) * H4 + H + h] = hn;
        if (p.Cs) p.Cs[((long)(t + 1) * B + b0 + n) * H + h] = cn;
      }
    }
    __syncthreads();
    if (GREEDY) {
      // ---- stage E: logits = OutE[tok] + Wout[:, H:] . [h; cT; cV]; argmax; feed back ------
      matvec_partial<NB>(p.Wo_t, p.Vp, p.Vp, H3, u_s, H3, part, KSE);
      __syncthreads();
      if (tid < NB) {
        const int n = tid;
        if (alive_s[n]) {
          const int tok = tok_s[n];
          float l[128 / 4];   // V <= 32 in greedy mode (checked on the host)
          float mx = -INFINITY;
          for (int v = 0; v < p.V; ++v) {
            l[v] = __ldg(p.OutE + tok * p.V + v) + part_sum<NB>(part, KSE, p.Vp, n, v);
            mx = fmaxf(mx, l[v]);
          }
          float sum = 0.f;
          for (int v = 0; v < p.V; ++v) sum += expf(l[v] - mx);
          const float lse = logf(sum);
          // first maximum of the log-softmax values, as F.log_softmax(...).max(dim=-1) gives
          float best = -INFINITY;
          int arg = 0;
          for (int v = 0; v < p.V; ++v) {
            float lp = (l[v] - mx) - lse;
            if (lp > best) { best = lp; arg = v; }
          }
          my_steps++;
          if (arg == p.eos) {
            alive_s[n] = 0;
          } else {
            p.out_tokens[(long)(b0 + n) * p.T + my_len] = arg;
            my_len++;
          }
          tok_s[n] = arg;
        }
      }
      __syncthreads();
      if (tid == 0) {
        int any = 0;
        for (int n = 0; n < NB; ++n) any |= alive_s[n];
        flag_s[0] = any;
      }
      __syncthreads();
      if (!flag_s[0]) break;
    }
  }

  // ---- epilogue -----------------------------------------------------------------------------
  if (p.beta_sum)
    for (int i = tid; i < NB * M; i += NT) {
      int n = i / M;
      if (n < nb) p.beta_sum[(long)b0 * M + i] = bsum_s[i];
    }
  if (p.h_out)
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      if (n < nb) {
        p.h_out[(long)(b0 + n) * H + h] = u_s[n * H3 + h];
        p.c_out[(long)(b0 + n) * H + h] = c_s[i];
      }
    }
  if (GREEDY && tid < nb) {
    p.out_len[b0 + tid] = my_len;
    p.out_steps[b0 + tid] = my_steps;
  }
}

// ---------------------------------------------------------------------------------------------
// Decoder backward sweep (BPTT; SURVEY.md A.6 / oracle/manual_backward.py stage B4).
// Produces per-step pre-activation gradients (consumed by batched weight-gradient GEMMs) and the
// gradients of the keys and of the initial state.
// ---------------------------------------------------------------------------------------------
struct DecBwdP {
  int B, T, Ti, M, H, cond;
  const float *W_ih, *W_hh, *W_qV, *W_c, *W_qT;   // original row-major parameters
  const float *vT, *vV;
  const float *KT, *KV;
  const int* cmd_len;
  const float *Cs, *gates, *alpha, *beta, *Qp, *qT, *qV;
  const float* dU;          // [T][B][4H]
  const float* dbeta_aux;   // [B][M] or null
  float *dgates, *dd, *dqV, *dqT;   // [T][B][4H], [T][B][H] x3
  float *dKT, *dKV, *dh0;   // [Ti][B][H], [B][M][H], [B][H]
  float *dvT, *dvV;         // [H] each, atomically accumulated (zeroed by the host)
};

template <int NB>
size_t dec_bwd_smem_floats(int Ti, int M, int H, int nthreads) {
  size_t n = 0;
  n += 2 * (size_t)pad4(NB * M * H) + 2 * (size_t)pad4(NB * Ti * H);  // KV_s, dKV_s, KT_s, dKT_s
  n += 7 * (size_t)pad4(NB * H);                               // dh, dc, dcT, dcV, dqV, dd, dqT
  n += pad4(NB * 4 * H);                                       // da_s
  n += 3 * (size_t)pad4(NB * M) + 3 * (size_t)pad4(NB * Ti);   // db, dr, bt for both attentions
  n += 2 * pad4(H);
  size_t p1 = (size_t)matvec_splits(2 * H, 4 * H, nthreads) * NB * 2 * H +
              (size_t)matvec_splits(H, 4 * H, nthreads) * NB * H;
  size_t p2 = (size_t)matvec_splits(H, H, nthreads) * NB * H;
  size_t p3 = (size_t)matvec_splits(2 * H, H, nthreads) * NB * 2 * H;
  size_t part = p1 > p2 ? p1 : p2;
  if (p3 > part) part = p3;
  n += pad4((int)part) + 16;
  return n;
}

// d(score) -> accumulate dK, dq, dv for one attention.  Thread (n,h) walks the N keys.
template <int NB>
__device__ __forceinline__ void attn_bwd_keys(const float* q_g, long q_stride_n, const float* K_s, float* dK_s,
                                              const float* w_s, const float* dr_s, const float* dc_s,
                                              const float* v_s, int N, int H, int nb, float* dq_s,
                                              float& dv_acc) {
  for (int i = threadIdx.x; i < NB * H; i += blockDim.x) {
    int n = i / H, h = i - n * H;
    float q = (n < nb) ? __ldg(q_g + n * q_stride_n + h) : 0.f;
    float dcv = dc_s[i], v = v_s[h], dq = 0.f;
    for (int j = 0; j < N; ++j) {
      int kj = (n * N + j) * H + h;
      float z = act_tanh(q + K_s[kj]);
      float dr = dr_s[n * N + j];
      float g = dr * v * (1.f - z * z);
      dK_s[kj] += fmaf(w_s[n * N + j], dcv, g);
      dq += g;
      dv_acc = fmaf(dr, z, dv_acc);
    }
    dq_s[i] = dq;
  }
}

template <int NB>
__global__ void __launch_bounds__(kRecThreads, 1) decoder_bwd_kernel(DecBwdP p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, H2 = 2 * H, H4 = 4 * H, M = p.M, Ti = p.Ti, B = p.B;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
  const int b0 = blockIdx.x * NB;
  const int nb = min(NB, B - b0);
  Bump bump{smem};
  float* KV_s = bump.take(NB * M * H);
  float* dKV_s = bump.take(NB * M * H);
  float* KT_s = bump.take(NB * Ti * H);
  float* dKT_s = bump.take(NB * Ti * H);
  float* dh_s = bump.take(NB * H);
  float* dc_s = bump.take(NB * H);
  float* dcT_s = bump.take(NB * H);
  float* dcV_s = bump.take(NB * H);
  float* dqV_s = bump.take(NB * H);
  float* dd_s = bump.take(NB * H);
  float* dqT_s = bump.take(NB * H);
  float* da_s = bump.take(NB * H4);
  float* dbV_s = bump.take(NB * M);
  float* drV_s = bump.take(NB * M);
  float* btV_s = bump.take(NB * M);
  float* dbT_s = bump.take(NB * Ti);
  float* drT_s = bump.take(NB * Ti);
  float* btT_s = bump.take(NB * Ti);
  float* vT_s = bump.take(H);
  float* vV_s = bump.take(H);
  const int KS1 = matvec_splits(H2, H4, NT), KS2 = matvec_splits(H, H4, NT);
  const int KS3 = matvec_splits(H, H, NT), KS4 = matvec_splits(H2, H, NT);
  float* part = bump.p;
  float* part2 = part + (size_t)KS1 * NB * H2;

  for (int i = tid; i < NB * M * H; i += NT) {
    int n = i / (M * H);
    KV_s[i] = (n < nb) ? __ldg(p.KV + (long)b0 * M * H + i) : 0.f;
    dKV_s[i] = 0.f;
  }
  for (int i = tid; i < NB * Ti * H; i += NT) {
    int n = i / (Ti * H), r = i - n * Ti * H;
    int j = r / H, h = r - j * H;
    KT_s[i] = (n < nb) ? __ldg(p.KT + ((long)j * B + b0 + n) * H + h) : 0.f;
    dKT_s[i] = 0.f;
  }
  for (int i = tid; i < NB * H; i += NT) { dh_s[i] = 0.f; dc_s[i] = 0.f; }
  for (int h = tid; h < H; h += NT) { vT_s[h] = __ldg(p.vT + h); vV_s[h] = __ldg(p.vV + h); }
  float dvV_acc = 0.f, dvT_acc = 0.f;   // valid for tid < NB*H (requires NB*H <= blockDim)
  __syncthreads();

  for (int t = p.T - 1; t >= 0; --t) {
    // ---- 1. LSTM cell backward -------------------------------------------------------------
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      float da0 = 0.f, da1 = 0.f, da2 = 0.f, da3 = 0.f;
      if (n < nb) {
        long row = (long)t * B + b0 + n;
        const float* gp = p.gates + row * H4 + h;
        float ig = __ldg(gp), fg = __ldg(gp + H), gg = __ldg(gp + 2 * H), og = __ldg(gp + 3 * H);
        float c_prev = __ldg(p.Cs + row * H + h);
        float c_new = __ldg(p.Cs + (row + B) * H + h);
        float dh_t = dh_s[i] + __ldg(p.dU + row * H4 + H + h);
        float tc = act_tanh(c_new);
        float d_o = dh_t * tc;
        float dc_t = fmaf(dh_t * og, 1.f - tc * tc, dc_s[i]);
        da0 = dc_t * gg * ig * (1.f - ig);
        da1 = dc_t * c_prev * fg * (1.f - fg);
        da2 = dc_t * ig * (1.f - gg * gg);
        da3 = d_o * og * (1.f - og);
        dc_s[i] = dc_t * fg;
        float* dg = p.dgates + row * H4 + h;
        dg[0] = da0; dg[H] = da1; dg[2 * H] = da2; dg[3 * H] = da3;
      }
      float* dp = da_s + n * H4 + h;
      dp[0] = da0; dp[H] = da1; dp[2 * H] = da2; dp[3 * H] = da3;
    }
    __syncthreads();
    // ---- 2. gate gradients back to [cT; cV] (W_ih[:, H:3H]^T) and to h_{t-1} (W_hh^T) --------
    matvec_partial<NB>(p.W_ih + H, 3 * H, H2, H4, da_s, H4, part, KS1);
    matvec_partial<NB>(p.W_hh, H, H, H4, da_s, H4, part2, KS2);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      float dct = part_sum<NB>(part, KS1, H2, n, h);
      float dcv = part_sum<NB>(part, KS1, H2, n, H + h);
      if (n < nb) {
        long row = (long)t * B + b0 + n;
        dct += __ldg(p.dU + row * H4 + 2 * H + h);
        dcv += __ldg(p.dU + row * H4 + 3 * H + h);
      }
      dcT_s[i] = dct;
      dcV_s[i] = dcv;
      dh_s[i] = part_sum<NB>(part2, KS2, H, n, h);
    }
    __syncthreads();
    // ---- 3. visual attention backward ----------------------------------------------------------
    for (int pidx = warp; pidx < NB * M; pidx += NW) {
      int n = pidx / M, m = pidx - n * M;
      const float* kp = KV_s + (long)pidx * H;
      float s = 0.f;
      for (int h = lane; h < H; h += 32) s = fmaf(dcV_s[n * H + h], kp[h], s);
      s = warp_sum(s);
      if (lane == 0) {
        float bt = 0.f;
        if (n < nb) {
          bt = __ldg(p.beta + ((long)t * B + b0 + n) * M + m);
          if (p.dbeta_aux) s += __ldg(p.dbeta_aux + (long)(b0 + n) * M + m);
        }
        dbV_s[pidx] = s;
        btV_s[pidx] = bt;
      }
    }
    __syncthreads();
    if (warp < NB) {
      float dot = 0.f;
      for (int m = lane; m < M; m += 32) dot = fmaf(dbV_s[warp * M + m], btV_s[warp * M + m], dot);
      dot = warp_sum(dot);
      for (int m = lane; m < M; m += 32) drV_s[warp * M + m] = btV_s[warp * M + m] * (dbV_s[warp * M + m] - dot);
    }
    __syncthreads();
    attn_bwd_keys<NB>(p.qV + ((long)t * B + b0) * H, H, KV_s, dKV_s, btV_s, drV_s, dcV_s, vV_s, M, H, nb,
                      dqV_s, dvV_acc);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      if (n < nb) p.dqV[((long)t * B + b0 + n) * H + h] = dqV_s[i];
    }
    // ---- 4. through W_qV to the visual query, 5. through the conditional layer ------------------
    matvec_partial<NB>(p.W_qV, H, H, H, dqV_s, H, part, KS3);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      float dqp = part_sum<NB>(part, KS3, H, n, h);
      if (p.cond) {
        float q = (n < nb) ? __ldg(p.Qp + ((long)t * B + b0 + n) * H + h) : 0.f;
        float d = dqp * (1.f - q * q);
        dd_s[i] = d;
        if (n < nb) p.dd[((long)t * B + b0 + n) * H + h] = d;
      } else {
        dh_s[i] += dqp;
      }
    }
    __syncthreads();
    if (p.cond) {
      matvec_partial<NB>(p.W_c, H2, H2, H, dd_s, H, part, KS4);
      __syncthreads();
      for (int i = tid; i < NB * H; i += NT) {
        int n = i / H, h = i - n * H;
        dh_s[i] += part_sum<NB>(part, KS4, H2, n, h);
        dcT_s[i] += part_sum<NB>(part, KS4, H2, n, H + h);
      }
      __syncthreads();
    }
    // ---- 6. textual attention backward ------------------------------------------------------------
    for (int pidx = warp; pidx < NB * Ti; pidx += NW) {
      int n = pidx / Ti, j = pidx - n * Ti;
      const float* kp = KT_s + (long)pidx * H;
      float s = 0.f;
      for (int h = lane; h < H; h += 32) s = fmaf(dcT_s[n * H + h], kp[h], s);
      s = warp_sum(s);
      if (lane == 0) {
        dbT_s[pidx] = s;
        btT_s[pidx] = (n < nb) ? __ldg(p.alpha + ((long)t * B + b0 + n) * Ti + j) : 0.f;
      }
    }
    __syncthreads();
    if (warp < NB) {
      float dot = 0.f;
      for (int j = lane; j < Ti; j += 32) dot = fmaf(dbT_s[warp * Ti + j], btT_s[warp * Ti + j], dot);
      dot = warp_sum(dot);
      for (int j = lane; j < Ti; j += 32) drT_s[warp * Ti + j] = btT_s[warp * Ti + j] * (dbT_s[warp * Ti + j] - dot);
    }
    __syncthreads();
    attn_bwd_keys<NB>(p.qT + ((long)t * B + b0) * H, H, KT_s, dKT_s, btT_s, drT_s, dcT_s, vT_s, Ti, H, nb,
                      dqT_s, dvT_acc);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      if (n < nb) p.dqT[((long)t * B + b0 + n) * H + h] = dqT_s[i];
    }
    // ---- 7. through W_qT back to h_{t-1} ------------------------------------------------------------
    matvec_partial<NB>(p.W_qT, H, H, H, dqT_s, H, part, KS3);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      dh_s[i] += part_sum<NB>(part, KS3, H, n, h);
    }
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------------------
  for (int i = tid; i < NB * M * H; i += NT) {
    int n = i / (M * H);
    if (n < nb) p.dKV[(long)b0 * M * H + i] = dKV_s[i];
  }
  for (int i = tid; i < NB * Ti * H; i += NT) {
    int n = i / (Ti * H), r = i - n * Ti * H;
    int j = r / H, h = r - j * H;
    if (n < nb) p.dKT[((long)j * B + b0 + n) * H + h] = dKT_s[i];
  }
  for (int i = tid; i < NB * H; i += NT) {
    int n = i / H, h = i - n * H;
    if (n < nb) p.dh0[(long)(b0 + n) * H + h] = dh_s[i] + dc_s[i];
    if (n < nb) {   // one (n,h) per thread: tid == i because NB*H <= blockDim
      atomicAdd(p.dvV + h, dvV_acc);
      atomicAdd(p.dvT + h, dvT_acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Command encoder: bidirectional LSTM over the valid tokens of each command
// (reference seq2seq_model.py:47-89).  grid = (ceil(B/NB), 2 directions).
// ---------------------------------------------------------------------------------------------
struct EncP {
  int B, Ti, H;
  const float* Whh_t[2];   // [H][4H] transposed W_hh per direction (forward sweep)
  const float* W_hh[2];    // [4H][H] original (backward sweep)
  const float* xg[2];      // [B][Ti][4H] input-gate pre-activations incl. both biases
  const int* len;
  float* enc_h[2];         // [Ti][B][H] state after position t (carried through pads)
  float* enc_c[2];
  float* enc_g[2];         // [Ti][B][4H] gate activations
  float* enc_out;          // [Ti][B][H]  sum of directions, zero at pads (zeroed by host; atomicAdd)
  float* h_enc;            // [B][H]      sum of final states (zeroed by host; atomicAdd)
  // backward
  const float* denc_out;   // [Ti][B][H]
  const float* dh_enc;     // [B][H]
  float* dga[2];           // [B][Ti][4H]
  float* hprev[2];         // [B][Ti][H]
};

template <int NB>
size_t enc_smem_floats(int H, int nthreads, bool bwd) {
  size_t n = 3 * (size_t)pad4(NB * H) + pad4(NB * 4 * H);
  size_t part = bwd ? (size_t)matvec_splits(H, 4 * H, nthreads) * NB * H
                    : (size_t)matvec_splits(4 * H, H, nthreads) * NB * 4 * H;
  return n + pad4((int)part) + 16;
}

template <int NB>
__global__ void __launch_bounds__(kRecThreads, 1) encoder_fwd_kernel(EncP p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, H4 = 4 * H, Ti = p.Ti, B = p.B;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int b0 = blockIdx.x * NB, d = blockIdx.y;
  const int nb = min(NB, B - b0);
  Bump bump{smem};
  float* h_s = bump.take(NB * H);
  float* c_s = bump.take(NB * H);
  int* len_s = reinterpret_cast<int*>(bump.take(16));
  const int KS = matvec_splits(H4, H, NT);
  float* part = bump.p;
  for (int i = tid; i < NB * H; i += NT) { h_s[i] = 0.f; c_s[i] = 0.f; }
  if (tid < NB) len_s[tid] = (tid < nb) ? max(1, min(p.len[b0 + tid], Ti)) : 0;
  __syncthreads();
  for (int step = 0; step < Ti; ++step) {
    const int t = d == 0 ? step : Ti - 1 - step;
    matvec_partial<NB>(p.Whh_t[d], H4, H4, H, h_s, H, part, KS);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      if (n >= nb) continue;
      const bool valid = t < len_s[n];
      const float* xp = p.xg[d] + ((long)(b0 + n) * Ti + t) * H4 + h;
      float a0 = __ldg(xp) + part_sum<NB>(part, KS, H4, n, h);
      float a1 = __ldg(xp + H) + part_sum<NB>(part, KS, H4, n, H + h);
      float a2 = __ldg(xp + 2 * H) + part_sum<NB>(part, KS, H4, n, 2 * H + h);
      float a3 = __ldg(xp + 3 * H) + part_sum<NB>(part, KS, H4, n, 3 * H + h);
      float ig = act_sigmoid(a0), fg = act_sigmoid(a1), gg = act_tanh(a2), og = act_sigmoid(a3);
      float cn = fmaf(fg, c_s[i], ig * gg);
      float hn = og * act_tanh(cn);
      long row = (long)t * B + b0 + n;
      float* gp = p.enc_g[d] + row * H4 + h;
      gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
      if (valid) {
        h_s[i] = hn;
        c_s[i] = cn;
        atomicAdd(p.enc_out + row * H + h, hn);   // two commutative adds onto zero: deterministic
      }
      p.enc_h[d][row * H + h] = h_s[i];
      p.enc_c[d][row * H + h] = c_s[i];
    }
    __syncthreads();
  }
  for (int i = tid; i < NB * H; i += NT) {
    int n = i / H, h = i - n * H;
    if (n < nb) atomicAdd(p.h_enc + (long)(b0 + n) * H + h, h_s[i]);
  }
}

template <int NB>
__global__ void __launch_bounds__(kRecThreads, 1) encoder_bwd_kernel(EncP p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, H4 = 4 * H, Ti = p.Ti, B = p.B;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int b0 = blockIdx.x * NB, d = blockIdx.y;
  const int nb = min(NB, B - b0);
  Bump bump{smem};
  float* dh_s = bump.take(NB * H);
  float* dc_s = bump.take(NB * H);
  float* da_s = bump.take(NB * H4);
  int* len_s = reinterpret_cast<int*>(bump.take(16));
  const int KS = matvec_splits(H, H4, NT);
  float* part = bump.p;
  for (int i = tid; i < NB * H; i += NT) {
    int n = i / H, h = i - n * H;
    dh_s[i] = (n < nb) ? __ldg(p.dh_enc + (long)(b0 + n) * H + h) : 0.f;
    dc_s[i] = 0.f;
  }
  if (tid < NB) len_s[tid] = (tid < nb) ? max(1, min(p.len[b0 + tid], Ti)) : 0;
  __syncthreads();
  for (int step = Ti - 1; step >= 0; --step) {
    const int t = d == 0 ? step : Ti - 1 - step;
    const int tp = d == 0 ? t - 1 : t + 1;   // position visited before t in this direction
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H, h = i - n * H;
      float da0 = 0.f, da1 = 0.f, da2 = 0.f, da3 = 0.f;
      if (n < nb) {
        const bool valid = t < len_s[n];
        long row = (long)t * B + b0 + n;
        float h_prev = 0.f, c_prev = 0.f;
        if (step > 0) {
          long rp = (long)tp * B + b0 + n;
          h_prev = __ldg(p.enc_h[d] + rp * H + h);
          c_prev = __ldg(p.enc_c[d] + rp * H + h);
        }
        p.hprev[d][((long)(b0 + n) * Ti + t) * H + h] = h_prev;
        if (valid) {
          const float* gp = p.enc_g[d] + row * H4 + h;
          float ig = __ldg(gp), fg = __ldg(gp + H), gg = __ldg(gp + 2 * H), og = __ldg(gp + 3 * H);
          float c_new = fmaf(fg, c_prev, ig * gg);
          float tc = act_tanh(c_new);
          float dh_t = dh_s[i] + __ldg(p.denc_out + row * H + h);
          float d_o = dh_t * tc;
          float dc_t = fmaf(dh_t * og, 1.f - tc * tc, dc_s[i]);
          da0 = dc_t * gg * ig * (1.f - ig);
          da1 = dc_t * c_prev * fg * (1.f - fg);
          da2 = dc_t * ig * (1.f - gg * gg);
          da3 = d_o * og * (1.f - og);
          dc_s[i] = dc_t * fg;
        }
        float* dg = p.dga[d] + ((long)(b0 + n) * Ti + t) * H4 + h;
        dg[0] = da0; dg[H] = da1; dg[2 * H] = da2; dg[3 * H] = da3;
      }
      float* dp = da_s + n * H4 + h;
      dp[0] = da0; dp[H] = da1; dp[2 * H] = da2; dp[3 * H] = da3;
    }
    __syncthreads();
    matvec_partial<NB>(p.W_hh[d], H, H, H4, da_s, H4, part, KS);
    __syncthreads();
    for (int i = tid; i < NB * H; i += NT) {
      int n = i / H;
      if (n < nb && t < len_s[n]) dh_s[i] = part_sum<NB>(part, KS, H, n, i - n * H);
    }
    __syncthreads();
  }
}

}  // namespace gscan
