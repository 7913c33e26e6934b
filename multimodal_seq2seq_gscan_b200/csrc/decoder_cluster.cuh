// Decoder sweeps, version 2: recurrent weights resident in the shared memory of a thread-block
// cluster (reference hot loop: seq2seq_model.py:473-480 calling forward_step 359-428).
//
// A cluster of C CTAs owns NB examples for the whole sequence.  CTA `r` of the cluster owns the
// hidden slice S_r = [r*hs, (r+1)*hs), hs = H/C, of EVERY H-sized quantity of those examples
// (h, c, q_T, q', q_V, the four gates, the key columns K[:, S_r]) and keeps the rows of all
// recurrent weight matrices that produce its slice in shared memory for all T steps, so no weight
// is re-read from L2 inside the time loop (v1 re-read ~640 KB per CTA per step).
// Per step the CTAs exchange only activations, by storing into each other's shared memory
// (DSMEM) followed by a cluster barrier:
//     X1  partial text-attention scores   (sum over ranks)      [NB][Ti] per rank
//     X3  q' slice (conditional query)    (all-gather)          [NB][hs] per rank
//     X4  partial visual-attention scores (sum over ranks)      [NB][M]  per rank
//     X5  c_V slice                       (all-gather)          [NB][hs] per rank
//     X6  h_t slice                       (all-gather)          [NB][hs] per rank
// The text context never has to be exchanged: everything linear in c_T = sum_j alpha_j K^T_j is
// computed as sum_j alpha_j P_j with P_j = W . K^T_j precomputed once per sequence (Ti <= ~10
// MACs per row instead of H, and alpha is known to every rank after X1).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "recurrent.cuh"

namespace gscan {

namespace cg = cooperative_groups;

constexpr int kClThreads = 512;
constexpr int kClWarps = kClThreads / 32;
constexpr int kClMaxNB = 8;

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// ---- weight slices, packed per rank so that each CTA loads one contiguous block ----------------
//   WA [C][H][RAl]   columns: q_T slice | (W_c[:, :H] slice) | W_hh gate i,f,g,o slices     (input h)
//   WC [C][H][hs]    W_qV slice                                                          (input q')
//   WD [C][H][4hs]   W_ih[:, 2H:3H] gate slices                                          (input c_V)
struct ClusterPackP {
  const float *W_qT, *W_c, *W_hh, *W_qV, *W_ih;
  float *WA, *WC, *WD;
  int H, C, hs, cond;
};

__global__ void pack_cluster_kernel(ClusterPackP p) {
  const int H = p.H, hs = p.hs, C = p.C;
  const int RAl = hs * (5 + p.cond);
  const long nA = (long)C * H * RAl, nC = (long)C * H * hs, nD = (long)C * H * 4 * hs;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < nA + nC + nD; idx += (long)gridDim.x * blockDim.x) {
    if (idx < nA) {
      int col = idx % RAl;
      long rk = idx / RAl;
      int k = rk % H, r = rk / H;
      int type = col / hs, i = col - type * hs, h = r * hs + i;
      float v;
      if (type == 0) v = p.W_qT[(long)h * H + k];
      else if (p.cond && type == 1) v = p.W_c[(long)h * 2 * H + k];
      else v = p.W_hh[((long)(type - 1 - p.cond) * H + h) * H + k];
      p.WA[idx] = v;
    } else if (idx < nA + nC) {
      long j = idx - nA;
      int i = j % hs;
      long rk = j / hs;
      int k = rk % H, r = rk / H;
      p.WC[j] = p.W_qV[(long)(r * hs + i) * H + k];
    } else {
      long j = idx - nA - nC;
      int col = j % (4 * hs);
      long rk = j / (4 * hs);
      int k = rk % H, r = rk / H;
      int g = col / hs, i = col - g * hs;
      p.WD[j] = p.W_ih[((long)g * H + r * hs + i) * 3 * H + 2 * H + k];
    }
  }
}

struct DecFwd2P {
  int B, T, Ti, M, H, cond, C, hs;
  const float *WA, *WC, *WD;
  const float* PT;   // [Ti][B][RB], RB = H*(cond + 4): rows of [W_c[:, H:2H] ; W_ih[:, H:2H]] applied to K^T
  const float *vT, *vV, *bc;
  const float* KT;   // [Ti][B][H]
  const float* KV;   // [B][M][H]
  const int* cmd_len;
  const float *h_init, *c_init;   // [B][H]
  const float* Xe;                // [T][B][4H]
  float *U, *Cs, *gates, *alpha, *beta, *Qp, *qT, *qV, *beta_sum;   // saved activations (see DecFwdP)
  long long* timeline;   // debug: [T][16] clock64 stamps of CTA 0 (GSCAN_TIMELINE=1), else null
};

// k-split plan of one mat-vec stage: warps = QG quad-groups (8 row-quads each) x KG k-groups,
// and 4 k-lanes inside a warp => 4*KG k-slices
struct MvPlan { int QG, KG; };
__host__ __device__ __forceinline__ MvPlan mv_plan(int R, int K) {
  MvPlan m;
  m.QG = ((R >> 2) + 7) >> 3;
  int kg = kClWarps / (m.QG > 0 ? m.QG : 1);
  int kmax = ((K >> 2) + 3) >> 2;
  if (kg > kmax) kg = kmax;
  m.KG = kg < 1 ? 1 : kg;
  return m;
}

// part[(kg*NB + n)*R + r] = sum over the k-slices of k-group kg of W_s[k*R + r] * x_s[n*ldx + k]
// W_s is k-major ([K][R], R % 4 == 0), x_s rows are 16-byte aligned, K % 4 == 0.
template <int NB>
__device__ __forceinline__ void matvec_smem(const float* __restrict__ W_s, int R, int K, const float* __restrict__ x_s,
                                            int ldx, float* __restrict__ part, MvPlan pl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= pl.QG * pl.KG) return;
  const int qg = warp % pl.QG, kg = warp / pl.QG;
  const int kl = lane >> 3, ql = lane & 7;
  const int RQ = R >> 2, KQ = K >> 2, KS = pl.KG * 4;
  const int q = qg * 8 + ql;
  const int s = kg * 4 + kl;
  const int kq0 = (KQ * s) / KS, kq1 = (KQ * (s + 1)) / KS;
  float2 acc[NB][2];
#pragma unroll
  for (int n = 0; n < NB; ++n) acc[n][0] = acc[n][1] = make_float2(0.f, 0.f);
  if (q < RQ) {
    for (int kq = kq0; kq < kq1; ++kq) {
      const float* wp = W_s + (4 * kq) * R + 4 * q;
      const float4 w0 = *reinterpret_cast<const float4*>(wp);
      const float4 w1 = *reinterpret_cast<const float4*>(wp + R);
      const float4 w2 = *reinterpret_cast<const float4*>(wp + 2 * R);
      const float4 w3 = *reinterpret_cast<const float4*>(wp + 3 * R);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float4 xv = *reinterpret_cast<const float4*>(x_s + n * ldx + 4 * kq);
        fma2(acc[n][0], make_float2(w0.x, w0.y), make_float2(xv.x, xv.x));
        fma2(acc[n][1], make_float2(w0.z, w0.w), make_float2(xv.x, xv.x));
        fma2(acc[n][0], make_float2(w1.x, w1.y), make_float2(xv.y, xv.y));
        fma2(acc[n][1], make_float2(w1.z, w1.w), make_float2(xv.y, xv.y));
        fma2(acc[n][0], make_float2(w2.x, w2.y), make_float2(xv.z, xv.z));
        fma2(acc[n][1], make_float2(w2.z, w2.w), make_float2(xv.z, xv.z));
        fma2(acc[n][0], make_float2(w3.x, w3.y), make_float2(xv.w, xv.w));
        fma2(acc[n][1], make_float2(w3.z, w3.w), make_float2(xv.w, xv.w));
      }
    }
  }
  // reduce over the 4 k-lanes with a halving butterfly: lane kl ends up owning row 4q + kl
  const bool hi1 = (kl & 2) != 0, hi0 = (kl & 1) != 0;
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    float2 keep = hi1 ? acc[n][1] : acc[n][0];
    float2 send = hi1 ? acc[n][0] : acc[n][1];
    keep.x += __shfl_xor_sync(0xffffffffu, send.x, 16);
    keep.y += __shfl_xor_sync(0xffffffffu, send.y, 16);
    float mine = hi0 ? keep.y : keep.x;
    float give = hi0 ? keep.x : keep.y;
    mine += __shfl_xor_sync(0xffffffffu, give, 8);
    if (q < RQ) part[(kg * NB + n) * R + 4 * q + kl] = mine;
  }
}

template <int NB>
__device__ __forceinline__ float mv_sum(const float* part, int KG, int R, int n, int r) {
  float v = 0.f;
  for (int g = 0; g < KG; ++g) v += part[(g * NB + n) * R + r];
  return v;
}

// partial scores over this rank's hidden slice, 4 lanes per (example, key) pair; the result is
// stored into slot `rank` of the exchange buffer of every CTA of the cluster
template <int NB>
__device__ __forceinline__ void cluster_partial_scores(cg::cluster_group& cluster, const float* q_s, const float* K_s,
                                                       const float* v_s, int N, int hs, float* x_buf, int rank,
                                                       int C) {
  const int lane = threadIdx.x & 31;
  const int total = NB * N * 4;
  for (int base = (threadIdx.x >> 5) * 32; base < total; base += kClThreads) {
    const int item = base + lane;
    const int pair = item >> 2, u = item & 3;
    float s = 0.f;
    if (item < total) {
      const int n = pair / N;
      const float* kp = K_s + pair * hs;
      const float* qp = q_s + n * hs;
      for (int h = u; h < hs; h += 4) s = fmaf(v_s[h], act_tanh(qp[h] + kp[h]), s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (item < total) {
      for (int dst = u; dst < C; dst += 4) {
        float* remote = cluster.map_shared_rank(x_buf, dst);
        remote[rank * NB * N + pair] = s;
      }
    }
  }
}

// softmax over the summed partial scores; warp n handles example n (NB <= number of warps)
template <int NB>
__device__ __forceinline__ void cluster_softmax(const float* x_buf, int C, int N, const int* len_s, bool masked,
                                                float* w_s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= NB) return;
  const int n = warp;
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) {
    float s = 0.f;
    for (int r = 0; r < C; ++r) s += x_buf[(r * NB + n) * N + j];
    if (masked && j >= len_s[n]) s = -INFINITY;
    w_s[n * N + j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < N; j += 32) {
    float e = __expf(w_s[n * N + j] - mx);
    w_s[n * N + j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < N; j += 32) w_s[n * N + j] *= inv;
}

struct ClFwdSmem {
  size_t WA, WC, WD, P, KT, KV, hfull, qpfull, cvfull, xT, xV, qT, ch, qV, g, c, al, be, bsum, vT, vV, bc, part, len;
  size_t total;
};

inline ClFwdSmem cl_fwd_smem(int NB, int C, int H, int Ti, int M, int cond) {
  ClFwdSmem s{};
  const int hs = H / C, RAl = hs * (5 + cond), RBl = hs * (4 + cond);
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~size_t(3); return r; };
  s.WA = take((size_t)H * RAl);
  s.WC = take((size_t)H * hs);
  s.WD = take((size_t)H * 4 * hs);
  s.P = take((size_t)NB * Ti * RBl);
  s.KT = take((size_t)NB * Ti * hs);
  s.KV = take((size_t)NB * M * hs);
  s.hfull = take((size_t)NB * H);
  s.qpfull = take((size_t)NB * H);
  s.cvfull = take((size_t)NB * H);
  s.xT = take((size_t)C * NB * Ti);
  s.xV = take((size_t)C * NB * M);
  s.qT = take((size_t)NB * hs);
  s.ch = take((size_t)NB * hs);
  s.qV = take((size_t)NB * hs);
  s.g = take((size_t)NB * 4 * hs);
  s.c = take((size_t)NB * hs);
  s.al = take((size_t)NB * Ti);
  s.be = take((size_t)NB * M);
  s.bsum = take((size_t)NB * M);
  s.vT = take(hs);
  s.vV = take(hs);
  s.bc = take(hs);
  MvPlan a = mv_plan(RAl, H), c = mv_plan(hs, H), d = mv_plan(4 * hs, H);
  size_t pa = (size_t)a.KG * NB * RAl, pc = (size_t)c.KG * NB * hs, pd = (size_t)d.KG * NB * 4 * hs;
  s.part = take(pa > pc ? (pa > pd ? pa : pd) : (pc > pd ? pc : pd));
  s.len = take(16);
  s.total = o;
  return s;
}

#define GSCAN_STAMP(k)                                                                   \
  do {                                                                                   \
    if (p.timeline && blockIdx.x == 0 && threadIdx.x == 0) p.timeline[t * 16 + (k)] = clock64(); \
  } while (0)

template <int NB>
__global__ void __launch_bounds__(kClThreads, 1) decoder_fwd_cluster_kernel(DecFwd2P p, ClFwdSmem L) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = p.C, hs = p.hs, H = p.H, H4 = 4 * H, M = p.M, Ti = p.Ti, B = p.B, cond = p.cond;
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / C) * NB;
  const int nb = min(NB, B - b0);
  const int S0 = rank * hs;
  const int RAl = hs * (5 + cond), RBl = hs * (4 + cond), G4 = 4 * hs;
  const int tid = threadIdx.x, NT = kClThreads;

  float* WA_s = smem + L.WA;
  float* WC_s = smem + L.WC;
  float* WD_s = smem + L.WD;
  float* P_s = smem + L.P;
  float* KT_s = smem + L.KT;
  float* KV_s = smem + L.KV;
  float* hfull_s = smem + L.hfull;
  float* qpfull_s = smem + L.qpfull;
  float* cvfull_s = smem + L.cvfull;
  float* xT_s = smem + L.xT;
  float* xV_s = smem + L.xV;
  float* qT_s = smem + L.qT;
  float* ch_s = smem + L.ch;
  float* qV_s = smem + L.qV;
  float* g_s = smem + L.g;
  float* c_s = smem + L.c;
  float* al_s = smem + L.al;
  float* be_s = smem + L.be;
  float* bsum_s = smem + L.bsum;
  float* vT_s = smem + L.vT;
  float* vV_s = smem + L.vV;
  float* bc_s = smem + L.bc;
  float* part = smem + L.part;
  int* len_s = reinterpret_cast<int*>(smem + L.len);

  // ---- one-time loads ----------------------------------------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(p.WA + (size_t)rank * H * RAl);
    float4* dst = reinterpret_cast<float4*>(WA_s);
    for (int i = tid; i < H * RAl / 4; i += NT) dst[i] = __ldg(src + i);
    src = reinterpret_cast<const float4*>(p.WC + (size_t)rank * H * hs);
    dst = reinterpret_cast<float4*>(WC_s);
    for (int i = tid; i < H * hs / 4; i += NT) dst[i] = __ldg(src + i);
    src = reinterpret_cast<const float4*>(p.WD + (size_t)rank * H * G4);
    dst = reinterpret_cast<float4*>(WD_s);
    for (int i = tid; i < H * G4 / 4; i += NT) dst[i] = __ldg(src + i);
  }
  {
    const int RB = H * (4 + cond);
    for (int i = tid; i < NB * Ti * RBl; i += NT) {
      int col = i % RBl, nj = i / RBl;
      int j = nj % Ti, n = nj / Ti;
      int type = col / hs, ii = col - type * hs;
      // global column: conditional rows first (if any), then gate rows g*H + h
      int gcol = (cond && type == 0) ? (S0 + ii) : (cond * H + (type - cond) * H + S0 + ii);
      P_s[i] = (n < nb) ? __ldg(p.PT + ((size_t)j * B + b0 + n) * RB + gcol) : 0.f;
    }
  }
  for (int i = tid; i < NB * Ti * hs; i += NT) {
    int h = i % hs, nj = i / hs;
    int j = nj % Ti, n = nj / Ti;
    KT_s[i] = (n < nb) ? __ldg(p.KT + ((size_t)j * B + b0 + n) * H + S0 + h) : 0.f;
  }
  for (int i = tid; i < NB * M * hs; i += NT) {
    int h = i % hs, nm = i / hs;
    int n = nm / M;
    KV_s[i] = (n < nb) ? __ldg(p.KV + ((size_t)b0 * M + nm) * H + S0 + h) : 0.f;
  }
  for (int i = tid; i < NB * H; i += NT) {
    int n = i / H, h = i - n * H;
    float hv = (n < nb) ? __ldg(p.h_init + (size_t)(b0 + n) * H + h) : 0.f;
    hfull_s[i] = hv;
    qpfull_s[i] = 0.f;
    cvfull_s[i] = 0.f;
    if (n < nb && h >= S0 && h < S0 + hs && p.U) p.U[(size_t)(b0 + n) * H4 + H + h] = hv;   // row group 0: h_{-1}
  }
  for (int i = tid; i < NB * hs; i += NT) {
    int n = i / hs, h = i - n * hs;
    float cv = (n < nb) ? __ldg(p.c_init + (size_t)(b0 + n) * H + S0 + h) : 0.f;
    c_s[i] = cv;
    if (n < nb && p.Cs) p.Cs[(size_t)(b0 + n) * H + S0 + h] = cv;
  }
  for (int i = tid; i < NB * M; i += NT) bsum_s[i] = 0.f;
  for (int h = tid; h < hs; h += NT) {
    vT_s[h] = __ldg(p.vT + S0 + h);
    vV_s[h] = __ldg(p.vV + S0 + h);
    bc_s[h] = cond ? __ldg(p.bc + S0 + h) : 0.f;
  }
  if (tid < NB) len_s[tid] = (tid < nb) ? max(1, min(p.cmd_len[b0 + tid], Ti)) : 1;
  const MvPlan plA = mv_plan(RAl, H), plC = mv_plan(hs, H), plD = mv_plan(G4, H);
  // all CTAs of the cluster must have started (and initialised their buffers) before any remote store
  cluster_sync_all();

  for (int t = 0; t < p.T; ++t) {
    const size_t row0 = (size_t)t * B + b0;   // + n
    GSCAN_STAMP(0);
    // prefetch the embedding part of the gate pre-activations of this step (consumed after stage A)
    float xe_pref[2] = {0.f, 0.f};
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int i = tid + it * NT;
      if (i < NB * G4) {
        int n = i / G4, col = i - n * G4;
        int g = col / hs, ii = col - g * hs;
        if (n < nb) xe_pref[it] = __ldg(p.Xe + (row0 + n) * H4 + g * H + S0 + ii);
      }
    }
    // ---- stage A: everything that depends only on h_{t-1} ----------------------------------------
    matvec_smem<NB>(WA_s, RAl, H, hfull_s, H, part, plA);
    __syncthreads();
    GSCAN_STAMP(1);
    for (int i = tid; i < NB * (RAl - G4); i += NT) {     // q_T and (conditional) W_c[:, :H] h
      const int W2 = RAl - G4;
      int n = i / W2, col = i - n * W2;
      float v = mv_sum<NB>(part, plA.KG, RAl, n, col);
      if (col < hs) {
        qT_s[n * hs + col] = v;
        if (p.qT && n < nb) p.qT[(row0 + n) * H + S0 + col] = v;
      } else {
        ch_s[n * hs + col - hs] = v;
      }
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {                      // hidden-to-hidden part of the gates
      int i = tid + it * NT;
      if (i < NB * G4) {
        int n = i / G4, col = i - n * G4;
        g_s[i] = mv_sum<NB>(part, plA.KG, RAl, n, RAl - G4 + col) + xe_pref[it];
      }
    }
    __syncthreads();
    GSCAN_STAMP(2);
    // ---- textual attention: partial scores over the local slice, summed over ranks (X1) --------
    cluster_partial_scores<NB>(cluster, qT_s, KT_s, vT_s, Ti, hs, xT_s, rank, C);
    GSCAN_STAMP(3);
    cluster_sync_all();
    GSCAN_STAMP(4);
    cluster_softmax<NB>(xT_s, C, Ti, len_s, true, al_s);
    __syncthreads();
    GSCAN_STAMP(5);
    // ---- stage B through P_j = W K^T_j: q' slice, gate contributions, c_T slice ---------------
    {
      const int QB = RBl >> 2, QH = hs >> 2;
      const int n1 = NB * QB, n2 = NB * QH;
      for (int i = tid; i < n1 + n2; i += NT) {
        if (i < n1) {
          int n = i / QB, cq = i - n * QB;
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          const float* pp = P_s + (size_t)n * Ti * RBl + 4 * cq;
          for (int j = 0; j < Ti; ++j) {
            const float a = al_s[n * Ti + j];
            const float4 v = *reinterpret_cast<const float4*>(pp + j * RBl);
            o.x = fmaf(a, v.x, o.x); o.y = fmaf(a, v.y, o.y); o.z = fmaf(a, v.z, o.z); o.w = fmaf(a, v.w, o.w);
          }
          const int col = 4 * cq;
          if (cond && col < hs) {
            const float4 chv = *reinterpret_cast<const float4*>(ch_s + n * hs + col);
            const float4 bcv = *reinterpret_cast<const float4*>(bc_s + col);
            float4 q;
            q.x = act_tanh(chv.x + o.x + bcv.x);
            q.y = act_tanh(chv.y + o.y + bcv.y);
            q.z = act_tanh(chv.z + o.z + bcv.z);
            q.w = act_tanh(chv.w + o.w + bcv.w);
            for (int dst = 0; dst < C; ++dst) {
              float* remote = cluster.map_shared_rank(qpfull_s, dst);
              *reinterpret_cast<float4*>(remote + n * H + S0 + col) = q;
            }
            if (p.Qp && n < nb) *reinterpret_cast<float4*>(p.Qp + (row0 + n) * H + S0 + col) = q;
          } else {
            float4* gp = reinterpret_cast<float4*>(g_s + n * G4 + col - cond * hs);
            float4 gv = *gp;
            gv.x += o.x; gv.y += o.y; gv.z += o.z; gv.w += o.w;
            *gp = gv;
          }
        } else {
          int k = i - n1;
          int n = k / QH, hq = k - n * QH;
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          const float* kp = KT_s + (size_t)n * Ti * hs + 4 * hq;
          for (int j = 0; j < Ti; ++j) {
            const float a = al_s[n * Ti + j];
            const float4 v = *reinterpret_cast<const float4*>(kp + j * hs);
            o.x = fmaf(a, v.x, o.x); o.y = fmaf(a, v.y, o.y); o.z = fmaf(a, v.z, o.z); o.w = fmaf(a, v.w, o.w);
          }
          if (p.U && n < nb) *reinterpret_cast<float4*>(p.U + (row0 + B + n) * H4 + 2 * H + S0 + 4 * hq) = o;
        }
      }
      if (p.alpha && rank == 0)
        for (int i = tid; i < nb * Ti; i += NT) {
          int n = i / Ti, j = i - n * Ti;
          p.alpha[(row0 + n) * Ti + j] = al_s[i];
        }
    }
    GSCAN_STAMP(6);
    if (cond) cluster_sync_all();   // X3: q' gathered
    else __syncthreads();
    GSCAN_STAMP(7);
    // ---- stage C: visual query slice -----------------------------------------------------------
    matvec_smem<NB>(WC_s, hs, H, cond ? qpfull_s : hfull_s, H, part, plC);
    __syncthreads();
    for (int i = tid; i < NB * hs; i += NT) {
      int n = i / hs, col = i - n * hs;
      float v = mv_sum<NB>(part, plC.KG, hs, n, col);
      qV_s[i] = v;
      if (p.qV && n < nb) p.qV[(row0 + n) * H + S0 + col] = v;
      if (!cond && p.Qp && n < nb) p.Qp[(row0 + n) * H + S0 + col] = hfull_s[n * H + S0 + col];
    }
    __syncthreads();
    // ---- visual attention: partial scores (X4), softmax, c_V slice gathered (X5) ----------------
    GSCAN_STAMP(8);
    cluster_partial_scores<NB>(cluster, qV_s, KV_s, vV_s, M, hs, xV_s, rank, C);
    GSCAN_STAMP(9);
    cluster_sync_all();
    GSCAN_STAMP(10);
    cluster_softmax<NB>(xV_s, C, M, len_s, false, be_s);
    __syncthreads();
    {
      // 4 lanes per (example, hidden quad): each sums a quarter of the cells, then a butterfly
      const int QH = hs >> 2;
      const int total = NB * QH * 4;
      const int lane = tid & 31;
      for (int base = (tid >> 5) * 32; base < total; base += NT) {
        const int item = base + lane;
        const int k = item >> 2, u = item & 3;
        const int n = k / QH, hq = k - n * QH;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < total) {
          const float* kp = KV_s + (size_t)n * M * hs + 4 * hq;
          for (int m = u; m < M; m += 4) {
            const float a = be_s[n * M + m];
            const float4 v = *reinterpret_cast<const float4*>(kp + m * hs);
            o.x = fmaf(a, v.x, o.x); o.y = fmaf(a, v.y, o.y); o.z = fmaf(a, v.z, o.z); o.w = fmaf(a, v.w, o.w);
          }
        }
#pragma unroll
        for (int sh = 1; sh <= 2; sh <<= 1) {
          o.x += __shfl_xor_sync(0xffffffffu, o.x, sh);
          o.y += __shfl_xor_sync(0xffffffffu, o.y, sh);
          o.z += __shfl_xor_sync(0xffffffffu, o.z, sh);
          o.w += __shfl_xor_sync(0xffffffffu, o.w, sh);
        }
        if (item < total) {
          for (int dst = u; dst < C; dst += 4) {
            float* remote = cluster.map_shared_rank(cvfull_s, dst);
            *reinterpret_cast<float4*>(remote + n * H + S0 + 4 * hq) = o;
          }
          if (u == 0 && p.U && n < nb) *reinterpret_cast<float4*>(p.U + (row0 + B + n) * H4 + 3 * H + S0 + 4 * hq) = o;
        }
      }
      for (int i = tid; i < NB * M; i += NT) {
        const float w = be_s[i];
        bsum_s[i] += w;
        int n = i / M, m = i - n * M;
        if (p.beta && rank == 0 && n < nb) p.beta[(row0 + n) * M + m] = w;
      }
    }
    GSCAN_STAMP(11);
    cluster_sync_all();   // X5
    GSCAN_STAMP(12);
    // ---- stage D: c_V contribution to the gates, then the LSTM cell -----------------------------
    matvec_smem<NB>(WD_s, G4, H, cvfull_s, H, part, plD);
    __syncthreads();
    GSCAN_STAMP(13);
    for (int i = tid; i < NB * hs; i += NT) {
      int n = i / hs, h = i - n * hs;
      float a[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) a[g] = g_s[n * G4 + g * hs + h] + mv_sum<NB>(part, plD.KG, G4, n, g * hs + h);
      const float ig = act_sigmoid(a[0]), fg = act_sigmoid(a[1]), gg = act_tanh(a[2]), og = act_sigmoid(a[3]);
      const float cn = fmaf(fg, c_s[i], ig * gg);
      const float hn = og * act_tanh(cn);
      c_s[i] = cn;
      for (int dst = 0; dst < C; ++dst) {
        float* remote = cluster.map_shared_rank(hfull_s, dst);
        remote[n * H + S0 + h] = hn;
      }
      if (n < nb) {
        const size_t row = row0 + n;
        if (p.gates) {
          float* gp = p.gates + row * H4 + S0 + h;
          gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
        }
        if (p.U) p.U[(row + B) * H4 + H + S0 + h] = hn;
        if (p.Cs) p.Cs[(row + B) * H + S0 + h] = cn;
      }
    }
    GSCAN_STAMP(14);
    cluster_sync_all();   // X6
    GSCAN_STAMP(15);
  }

  if (p.beta_sum && rank == 0)
    for (int i = tid; i < nb * M; i += NT) p.beta_sum[(size_t)b0 * M + i] = bsum_s[i];
}

}  // namespace gscan
