// Decoder sweeps, version 3 (reference hot loop: seq2seq_model.py:473-480 calling forward_step
// 359-428; backward = autograd of the same).  Compile-time specialised for the gSCAN paper shape
// H = 100, 6x6 grid: a cluster of 5 CTAs owns 8 examples for the whole sequence, CTA r owns the
// hidden slice [20r, 20r+20) of every H-sized quantity.
//
//  * Recurrent weights live in REGISTERS for the whole sweep: every thread owns a fixed
//    (row pair, k-slice) tile of each mat-vec stage, so the time loop reads no weights at all
//    (v1 streamed ~640 KB/step/CTA from L2, v2 re-read 88 KB/step/CTA from shared memory).
//  * CTAs exchange activations with one-sided stores into each other's shared memory
//    (st.async ... mbarrier::complete_tx): the receiver waits on a local mbarrier whose
//    transaction count covers the bytes of all 5 senders.  There is no cluster-wide barrier in
//    the time loop (v2: 5 barrier.cluster per step at ~750-1200 cycles each).
//  * Everything linear in the text context is evaluated as sum_j alpha_j P_j with
//    P_j = W K^T_j precomputed per sequence (Ti <= ~10 terms instead of H, no exchange of c_T).
#pragma once
#include "common.cuh"

namespace gscan {
namespace v3 {

constexpr int kH = 100, kC = 5, kHS = 20, kM = 36, kG4 = 80, kNB = 8, kThreads = 256;
constexpr int kXS = 100;   // row stride of the gathered activation vectors
constexpr int kGS = 84;    // row stride of the gate pre-activation scratch (bank spread)
constexpr int kMaxTi = 16;

// ---- PTX helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// one-sided store into the shared memory of a CTA of the cluster; the bytes are counted on that
// CTA's mbarrier, whose completion makes them visible to the waiting threads
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr),
               "r"(__float_as_uint(v)), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_async_f32x4(uint32_t raddr, float4 v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
               "r"(__float_as_uint(v.w)), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

// ---- shared-memory layout (float offsets) --------------------------------------------------------
struct FwdSmem {
  int hfull, qpfull, cvfull, xT, xV, P, KT, KV, qT, ch, qV, g, al, be, vT, vV, bc, len, bars, total;
};
__host__ __device__ inline FwdSmem fwd_smem(int Ti, int cond) {
  FwdSmem s{};
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  const int RBl = kHS * (4 + cond);
  s.hfull = take(kNB * kXS);
  s.qpfull = take(kNB * kXS);
  s.cvfull = take(kNB * kXS);
  s.xT = take(kC * kNB * Ti);
  s.xV = take(kC * kNB * kM);
  s.P = take(kNB * Ti * RBl);
  s.KT = take(kNB * Ti * kHS);
  s.KV = take(kNB * kM * kHS);
  s.qT = take(kNB * kHS);
  s.ch = take(kNB * kHS);
  s.qV = take(kNB * kHS);
  s.g = take(kNB * kGS);
  s.al = take(kNB * Ti);
  s.be = take(kNB * kM);
  s.vT = take(kHS);
  s.vV = take(kHS);
  s.bc = take(kHS);
  s.len = take(kNB);
  s.bars = take(16);   // 5 mbarriers (8 bytes each)
  s.total = o;
  return s;
}

struct DecFwd3P {
  int B, T, Ti;
  const float *W_qT, *W_c, *W_hh, *W_qV, *W_ih;   // original row-major parameters
  const float* PT;   // [Ti][B][RB], RB = H*(cond + 4): [W_c[:, H:2H] ; W_ih[:, H:2H]] . K^T_j
  const float *vT, *vV, *bc;
  const float* KT;   // [Ti][B][H]
  const float* KV;   // [B][M][H]
  const int* cmd_len;
  const float *h_init, *c_init;   // [B][H]
  const float* Xe;                // [T][B][4H]
  float *U, *Cs, *gates, *alpha, *beta, *Qp, *qT, *qV, *beta_sum;   // saved activations (recurrent.cuh DecFwdP)
  long long* timeline;   // debug: [T][16] clock64 stamps of CTA 0, else null
};

#define GSCAN3_STAMP(k)                                                                            \
  do {                                                                                             \
    if (p.timeline && blockIdx.x == 0 && threadIdx.x == 0) p.timeline[t * 16 + (k)] = clock64();   \
  } while (0)

// 4-lane mat-vec tile: rows (2 per thread) x 7 k-quads (quad 4i+ks) x 8 examples.
// Returns in o[0..3] the full dot products of row (2*rp + (ks>>1)) for examples 4*(ks&1)+m.
__device__ __forceinline__ void mv_rowpair(const float4 (&w0)[7], const float4 (&w1)[7], const float* __restrict__ x,
                                           int ks, float (&o)[4]) {
  float2 a0[kNB], a1[kNB];
#pragma unroll
  for (int n = 0; n < kNB; ++n) a0[n] = a1[n] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int q = min(4 * i + ks, kH / 4 - 1);   // clamped quads carry zero weights
    const float* xp = x + 4 * q;
#pragma unroll
    for (int n = 0; n < kNB; ++n) {
      const float4 xv = lds4(xp + n * kXS);
      fma2(a0[n], lo2(w0[i]), lo2(xv));
      fma2(a0[n], hi2(w0[i]), hi2(xv));
      fma2(a1[n], lo2(w1[i]), lo2(xv));
      fma2(a1[n], hi2(w1[i]), hi2(xv));
    }
  }
  const bool up = (ks & 2) != 0, odd = (ks & 1) != 0;
  float keep[kNB];
#pragma unroll
  for (int n = 0; n < kNB; ++n) {
    const float r0 = a0[n].x + a0[n].y, r1 = a1[n].x + a1[n].y;
    const float mine = up ? r1 : r0, give = up ? r0 : r1;
    keep[n] = mine + __shfl_xor_sync(0xffffffffu, give, 2);
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const float mine = odd ? keep[4 + m] : keep[m], give = odd ? keep[m] : keep[4 + m];
    o[m] = mine + __shfl_xor_sync(0xffffffffu, give, 1);
  }
}

// partial attention scores over this CTA's hidden slice: 4 lanes per (example, key) pair, 5 hidden
// units each; the pair's partial sum is stored into slot `rank` of the score buffer of every CTA
template <int NKEYS_CT>
__device__ __forceinline__ void partial_scores(const float* __restrict__ q_s, const float* __restrict__ K_s,
                                               const float* __restrict__ v_s, int nkeys, int xoff_floats, int rank,
                                               uint32_t rb_u, uint32_t rb_4, uint32_t bar_off) {
  const int lane = threadIdx.x & 31, u = lane & 3;
  const int N = NKEYS_CT > 0 ? NKEYS_CT : nkeys;
  const int total = kNB * N * 4;
  float v[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) v[i] = v_s[5 * u + i];
  for (int base = (threadIdx.x >> 5) * 32; base < total; base += kThreads) {
    const int item = base + lane;
    const int pair = item >> 2;
    float s = 0.f;
    if (item < total) {
      const int n = pair / N;
      const float* kp = K_s + pair * kHS + 5 * u;
      const float* qp = q_s + n * kHS + 5 * u;
#pragma unroll
      for (int i = 0; i < 5; ++i) s = fmaf(v[i], act_tanh(qp[i] + kp[i]), s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (item < total) {
      const uint32_t off = (uint32_t)(xoff_floats + rank * kNB * N + pair) * 4u;
      st_async_f32(rb_u + off, s, rb_u + bar_off);
      if (u == 0) st_async_f32(rb_4 + off, s, rb_4 + bar_off);
    }
  }
}

template <bool COND>
__global__ void __cluster_dims__(kC, 1, 1) __launch_bounds__(kThreads, 1) dec_fwd_v3_kernel(DecFwd3P p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RBl = kHS * (4 + (COND ? 1 : 0));
  constexpr int QB = RBl / 4;
  constexpr int H4 = 4 * kH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int B = p.B, Ti = p.Ti;
  const int b0 = (blockIdx.x / kC) * kNB;
  const int nb = min(kNB, B - b0);
  const int S0 = rank * kHS;
  const FwdSmem L = fwd_smem(Ti, COND ? 1 : 0);

  float* hfull_s = smem + L.hfull;
  float* qpfull_s = smem + L.qpfull;
  float* cvfull_s = smem + L.cvfull;
  float* xT_s = smem + L.xT;
  float* xV_s = smem + L.xV;
  float* P_s = smem + L.P;
  float* KT_s = smem + L.KT;
  float* KV_s = smem + L.KV;
  float* qT_s = smem + L.qT;
  float* ch_s = smem + L.ch;
  float* qV_s = smem + L.qV;
  float* g_s = smem + L.g;
  float* al_s = smem + L.al;
  float* be_s = smem + L.be;
  float* vT_s = smem + L.vT;
  float* vV_s = smem + L.vV;
  float* bc_s = smem + L.bc;
  int* len_s = reinterpret_cast<int*>(smem + L.len);

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar0 = smem_base + (uint32_t)L.bars * 4u;   // [0] xT  [1] qp  [2] xV  [3] cv  [4] h
  uint32_t rb[kC];
#pragma unroll
  for (int d = 0; d < kC; ++d) rb[d] = mapa_u32(smem_base, (uint32_t)d);
  const uint32_t rb_u = mapa_u32(smem_base, (uint32_t)(lane & 3)), rb_4 = rb[4];
  const uint32_t boff = (uint32_t)L.bars * 4u;

  // ---- register-resident weights -----------------------------------------------------------------
  const int ks = tid & 3, ks8 = tid & 7;
  const bool actA = tid < 240, actC = COND && tid < 160, actD = tid >= 96;
  float4 wA0[7], wA1[7], wD0[7], wD1[7], wC[4];
  int lrA = 0;   // local output row of stage A owned after the butterfly
  {
    const int rp = tid >> 2;
    auto rowA = [&](int lr) -> const float* {
      const int type = lr / kHS, i = lr - type * kHS, hr = S0 + i;
      if (type == 0) return p.W_qT + (size_t)hr * kH;
      if (type == 1) return COND ? p.W_c + (size_t)hr * 2 * kH : p.W_qV + (size_t)hr * kH;
      return p.W_hh + (size_t)((type - 2) * kH + hr) * kH;
    };
    const float* r0 = rowA(actA ? 2 * rp : 0);
    const float* r1 = rowA(actA ? 2 * rp + 1 : 0);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int q = 4 * i + ks;
      const bool ok = actA && q < kH / 4;
      wA0[i] = ok ? ldg4(r0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      wA1[i] = ok ? ldg4(r1 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    lrA = 2 * rp + (ks >> 1);
    const int rpD = actD ? (tid - 96) >> 2 : 0;
    auto rowD = [&](int lr) -> const float* {
      const int g = lr / kHS, i = lr - g * kHS;
      return p.W_ih + (size_t)(g * kH + S0 + i) * 3 * kH + 2 * kH;
    };
    const float* d0 = rowD(2 * rpD);
    const float* d1 = rowD(2 * rpD + 1);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int q = 4 * i + ks;
      const bool ok = actD && q < kH / 4;
      wD0[i] = ok ? ldg4(d0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      wD1[i] = ok ? ldg4(d1 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int rC = actC ? tid >> 3 : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = 8 * i + ks8;
      const bool ok = actC && q < kH / 4;
      wC[i] = ok ? ldg4(p.W_qV + (size_t)(S0 + rC) * kH + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const int lrD = actD ? 2 * ((tid - 96) >> 2) + (ks >> 1) : 0;   // gate row owned after the stage-D butterfly
  const int nA = 4 * (ks & 1);                                     // first example owned after a 4-lane butterfly

  // ---- one-time loads ------------------------------------------------------------------------------
  {
    constexpr int RB = kH * (4 + (COND ? 1 : 0));
    for (int i = tid; i < kNB * Ti * RBl; i += kThreads) {
      const int col = i % RBl, nj = i / RBl;
      const int j = nj % Ti, n = nj / Ti;
      const int type = col / kHS, ii = col - type * kHS;
      const int gcol = type * kH + S0 + ii;   // conditional rows first (if any), then gate rows g*H + h
      P_s[i] = (n < nb) ? __ldg(p.PT + ((size_t)j * B + b0 + n) * RB + gcol) : 0.f;
    }
    for (int i = tid; i < kNB * Ti * kHS; i += kThreads) {
      const int h = i % kHS, nj = i / kHS;
      const int j = nj % Ti, n = nj / Ti;
      KT_s[i] = (n < nb) ? __ldg(p.KT + ((size_t)j * B + b0 + n) * kH + S0 + h) : 0.f;
    }
    for (int i = tid; i < kNB * kM * kHS; i += kThreads) {
      const int h = i % kHS, nm = i / kHS;
      const int n = nm / kM;
      KV_s[i] = (n < nb) ? __ldg(p.KV + ((size_t)b0 * kM + nm) * kH + S0 + h) : 0.f;
    }
    for (int i = tid; i < kNB * kH; i += kThreads) {
      const int n = i / kH, h = i - n * kH;
      const float hv = (n < nb) ? __ldg(p.h_init + (size_t)(b0 + n) * kH + h) : 0.f;
      hfull_s[n * kXS + h] = hv;
      qpfull_s[n * kXS + h] = 0.f;
      cvfull_s[n * kXS + h] = 0.f;
      if (n < nb && h >= S0 && h < S0 + kHS) p.U[(size_t)(b0 + n) * H4 + kH + h] = hv;   // row group 0: h_{-1}
    }
    for (int i = tid; i < kNB * kGS; i += kThreads) g_s[i] = 0.f;
    if (tid < kHS) {
      vT_s[tid] = __ldg(p.vT + S0 + tid);
      vV_s[tid] = __ldg(p.vV + S0 + tid);
      bc_s[tid] = COND ? __ldg(p.bc + S0 + tid) : 0.f;
    }
    if (tid < kNB) len_s[tid] = (tid < nb) ? max(1, min(p.cmd_len[b0 + tid], Ti)) : 1;
    if (tid == 0) {
      for (int k = 0; k < 5; ++k) mbar_init(bar0 + 8u * k, 1);
      fence_mbar_init();
    }
  }
  // cell state of (example tid/20, hidden S0 + tid%20), thread-private for the whole sweep
  const int cn = tid / kHS, chh = tid - cn * kHS;
  float c_reg = 0.f;
  if (tid < kNB * kHS) {
    if (cn < nb) {
      c_reg = __ldg(p.c_init + (size_t)(b0 + cn) * kH + S0 + chh);
      p.Cs[(size_t)(b0 + cn) * kH + S0 + chh] = c_reg;
    }
  }
  float bs0 = 0.f, bs1 = 0.f;   // sum over steps of beta[warp][lane], beta[warp][lane + 32]
  // Xe prefetch for the gate outputs this lane owns after the stage-A butterfly
  const bool gateA = actA && lrA >= 2 * kHS;
  const int xe_col = gateA ? ((lrA - 2 * kHS) / kHS) * kH + S0 + (lrA % kHS) : 0;
  float xe[4] = {0.f, 0.f, 0.f, 0.f};
  if (gateA) {
#pragma unroll
    for (int m = 0; m < 4; ++m)
      if (nA + m < nb) xe[m] = __ldg(p.Xe + (size_t)(b0 + nA + m) * H4 + xe_col);
  }
  // all CTAs of the cluster must have initialised their barriers and buffers before any remote store
  __syncthreads();
  cluster_barrier();

  const uint32_t bytes_xT = (uint32_t)(kC * kNB * Ti * 4), bytes_vec = (uint32_t)(kNB * kH * 4),
                 bytes_xV = (uint32_t)(kC * kNB * kM * 4);

  for (int t = 0; t < p.T; ++t) {
    const size_t row0 = (size_t)t * B + b0;   // + n
    const uint32_t par = (uint32_t)(t & 1);
    GSCAN3_STAMP(0);
    if (t > 0) mbar_wait(bar0 + 8u * 4, par ^ 1u);   // h_{t-1} gathered (X6 of the previous step)
    if (tid == 0) {
      mbar_arm(bar0 + 8u * 0, bytes_xT);
      if (COND) mbar_arm(bar0 + 8u * 1, bytes_vec);
      mbar_arm(bar0 + 8u * 2, bytes_xV);
      mbar_arm(bar0 + 8u * 3, bytes_vec);
      if (t + 1 < p.T) mbar_arm(bar0 + 8u * 4, bytes_vec);
    }
    GSCAN3_STAMP(1);
    // ---- stage A: everything that depends only on h_{t-1} ---------------------------------------------
    {
      float o[4];
      mv_rowpair(wA0, wA1, hfull_s, ks, o);   // all lanes take part in the butterfly; idle ones carry zero weights
      const int type = lrA / kHS, i = lrA - type * kHS;
      if (actA) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int n = nA + m;
        if (type == 0) {
          qT_s[n * kHS + i] = o[m];
          if (n < nb) p.qT[(row0 + n) * kH + S0 + i] = o[m];
        } else if (type == 1) {
          if (COND) {
            ch_s[n * kHS + i] = o[m];
          } else {
            qV_s[n * kHS + i] = o[m];
            if (n < nb) {
              p.qV[(row0 + n) * kH + S0 + i] = o[m];
              p.Qp[(row0 + n) * kH + S0 + i] = hfull_s[n * kXS + S0 + i];
            }
          }
        } else {
          g_s[n * kGS + lrA - 2 * kHS] = o[m] + xe[m];
        }
      }
      }
      if (gateA && t + 1 < p.T) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (nA + m < nb) xe[m] = __ldg(p.Xe + (row0 + B + nA + m) * H4 + xe_col);
      }
    }
    __syncthreads();
    GSCAN3_STAMP(2);
    // ---- textual attention: partial scores over the local slice, summed over ranks (X1) ---------------
    partial_scores<0>(qT_s, KT_s, vT_s, Ti, L.xT, rank, rb_u, rb_4, boff + 0u);
    GSCAN3_STAMP(3);
    mbar_wait(bar0 + 8u * 0, par);
    GSCAN3_STAMP(4);
    {
      const int n = warp;   // kNB == number of warps
      float s = -INFINITY;
      if (lane < Ti) {
        s = 0.f;
#pragma unroll
        for (int r = 0; r < kC; ++r) s += xT_s[(r * kNB + n) * Ti + lane];
        if (lane >= len_s[n]) s = -INFINITY;
      }
      const float mx = warp_max(s);
      const float e = (lane < Ti) ? __expf(s - mx) : 0.f;
      const float sum = warp_sum(e);
      const float a = e * (1.0f / sum);
      if (lane < Ti) {
        al_s[n * Ti + lane] = a;
        if (n < nb && rank == n % kC) p.alpha[(row0 + n) * Ti + lane] = a;
      }
    }
    __syncthreads();
    GSCAN3_STAMP(5);
    // ---- everything linear in c_T through P_j = W K^T_j: q' slice (X3), gate contributions, c_T slice ----
    if (tid < kNB * (QB + 5)) {
      const int n = tid / (QB + 5), q = tid - n * (QB + 5);
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = (q < QB) ? P_s + (size_t)n * Ti * RBl + 4 * q : KT_s + (size_t)n * Ti * kHS + 4 * (q - QB);
      const int stride = (q < QB) ? RBl : kHS;
      for (int j = 0; j < Ti; ++j) {
        const float a = al_s[n * Ti + j];
        const float4 v = lds4(src + j * stride);
        o.x = fmaf(a, v.x, o.x); o.y = fmaf(a, v.y, o.y); o.z = fmaf(a, v.z, o.z); o.w = fmaf(a, v.w, o.w);
      }
      if (COND && q < 5) {
        const float4 chv = lds4(ch_s + n * kHS + 4 * q);
        const float4 bcv = lds4(bc_s + 4 * q);
        float4 qq;
        qq.x = act_tanh(chv.x + o.x + bcv.x);
        qq.y = act_tanh(chv.y + o.y + bcv.y);
        qq.z = act_tanh(chv.z + o.z + bcv.z);
        qq.w = act_tanh(chv.w + o.w + bcv.w);
        const uint32_t off = (uint32_t)(L.qpfull + n * kXS + S0 + 4 * q) * 4u;
#pragma unroll
        for (int d = 0; d < kC; ++d) st_async_f32x4(rb[d] + off, qq, rb[d] + boff + 8u * 1);
        if (n < nb) *reinterpret_cast<float4*>(p.Qp + (row0 + n) * kH + S0 + 4 * q) = qq;
      } else if (q < QB) {
        float4* gp = reinterpret_cast<float4*>(g_s + n * kGS + 4 * q - (COND ? kHS : 0));
        float4 gv = *gp;
        gv.x += o.x; gv.y += o.y; gv.z += o.z; gv.w += o.w;
        *gp = gv;
      } else if (n < nb) {
        *reinterpret_cast<float4*>(p.U + (row0 + B + n) * H4 + 2 * kH + S0 + 4 * (q - QB)) = o;
      }
    }
    GSCAN3_STAMP(6);
    // ---- stage C: visual query slice ---------------------------------------------------------------------
    if (COND) {
      mbar_wait(bar0 + 8u * 1, par);
      GSCAN3_STAMP(7);
      if (actC) {
        float2 acc[kNB];
#pragma unroll
        for (int n = 0; n < kNB; ++n) acc[n] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int q = min(8 * i + ks8, kH / 4 - 1);
#pragma unroll
          for (int n = 0; n < kNB; ++n) {
            const float4 xv = lds4(qpfull_s + n * kXS + 4 * q);
            fma2(acc[n], lo2(wC[i]), lo2(xv));
            fma2(acc[n], hi2(wC[i]), hi2(xv));
          }
        }
        // 8-lane halving butterfly: lane ks8 ends up with example n = ks8
        float k4[4], k2[2];
        {
          const bool up = (ks8 & 4) != 0;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const float lo = acc[m].x + acc[m].y, hi = acc[4 + m].x + acc[4 + m].y;
            k4[m] = (up ? hi : lo) + __shfl_xor_sync(0xffffffffu, up ? lo : hi, 4);
          }
        }
        {
          const bool up = (ks8 & 2) != 0;
#pragma unroll
          for (int m = 0; m < 2; ++m) k2[m] = (up ? k4[2 + m] : k4[m]) + __shfl_xor_sync(0xffffffffu, up ? k4[m] : k4[2 + m], 2);
        }
        const bool up = (ks8 & 1) != 0;
        const float v = (up ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, up ? k2[0] : k2[1], 1);
        const int r = tid >> 3, n = ks8;
        qV_s[n * kHS + r] = v;
        if (n < nb) p.qV[(row0 + n) * kH + S0 + r] = v;
      }
    }
    __syncthreads();
    GSCAN3_STAMP(8);
    // ---- visual attention: partial scores (X4), softmax, c_V slice gathered (X5) -------------------------
    partial_scores<kM>(qV_s, KV_s, vV_s, kM, L.xV, rank, rb_u, rb_4, boff + 8u * 2);
    GSCAN3_STAMP(9);
    mbar_wait(bar0 + 8u * 2, par);
    GSCAN3_STAMP(10);
    {
      const int n = warp;
      float s0 = 0.f, s1 = -INFINITY;
#pragma unroll
      for (int r = 0; r < kC; ++r) s0 += xV_s[(r * kNB + n) * kM + lane];
      if (lane < kM - 32) {
        s1 = 0.f;
#pragma unroll
        for (int r = 0; r < kC; ++r) s1 += xV_s[(r * kNB + n) * kM + 32 + lane];
      }
      const float mx = warp_max(fmaxf(s0, s1));
      const float e0 = __expf(s0 - mx), e1 = (lane < kM - 32) ? __expf(s1 - mx) : 0.f;
      const float inv = 1.0f / warp_sum(e0 + e1);
      const float w0 = e0 * inv, w1 = e1 * inv;
      be_s[n * kM + lane] = w0;
      bs0 += w0;
      if (lane < kM - 32) {
        be_s[n * kM + 32 + lane] = w1;
        bs1 += w1;
      }
      if (n < nb && rank == n % kC) {
        p.beta[(row0 + n) * kM + lane] = w0;
        if (lane < kM - 32) p.beta[(row0 + n) * kM + 32 + lane] = w1;
      }
    }
    __syncthreads();
    if (tid < kNB * 5 * 4) {
      // 4 lanes per (example, hidden quad): each sums 9 of the 36 cells, then a butterfly all-reduce
      const int k = tid >> 2, u = tid & 3;
      const int n = k / 5, hq = k - n * 5;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* kp = KV_s + (size_t)n * kM * kHS + 4 * hq;
#pragma unroll
      for (int mm = 0; mm < kM / 4; ++mm) {
        const int m = 4 * mm + u;
        const float a = be_s[n * kM + m];
        const float4 v = lds4(kp + m * kHS);
        o.x = fmaf(a, v.x, o.x); o.y = fmaf(a, v.y, o.y); o.z = fmaf(a, v.z, o.z); o.w = fmaf(a, v.w, o.w);
      }
#pragma unroll
      for (int sh = 1; sh <= 2; sh <<= 1) {
        o.x += __shfl_xor_sync(0xffffffffu, o.x, sh);
        o.y += __shfl_xor_sync(0xffffffffu, o.y, sh);
        o.z += __shfl_xor_sync(0xffffffffu, o.z, sh);
        o.w += __shfl_xor_sync(0xffffffffu, o.w, sh);
      }
      const uint32_t off = (uint32_t)(L.cvfull + n * kXS + S0 + 4 * hq) * 4u;
      st_async_f32x4(rb_u + off, o, rb_u + boff + 8u * 3);
      if (u == 0) st_async_f32x4(rb_4 + off, o, rb_4 + boff + 8u * 3);
      if (u == 1 && n < nb) *reinterpret_cast<float4*>(p.U + (row0 + B + n) * H4 + 3 * kH + S0 + 4 * hq) = o;
    }
    GSCAN3_STAMP(11);
    mbar_wait(bar0 + 8u * 3, par);
    GSCAN3_STAMP(12);
    // ---- stage D: c_V contribution to the gates, then the LSTM cell ---------------------------------------
    if (actD) {
      float o[4];
      mv_rowpair(wD0, wD1, cvfull_s, ks, o);
#pragma unroll
      for (int m = 0; m < 4; ++m) g_s[(nA + m) * kGS + lrD] += o[m];
    }
    __syncthreads();
    GSCAN3_STAMP(13);
    if (tid < kNB * kHS) {
      const float* gp = g_s + cn * kGS + chh;
      const float ig = act_sigmoid(gp[0]), fg = act_sigmoid(gp[kHS]), gg = act_tanh(gp[2 * kHS]), og = act_sigmoid(gp[3 * kHS]);
      c_reg = fmaf(fg, c_reg, ig * gg);
      const float hn = og * act_tanh(c_reg);
      if (t + 1 < p.T) {
        const uint32_t off = (uint32_t)(L.hfull + cn * kXS + S0 + chh) * 4u;
#pragma unroll
        for (int d = 0; d < kC; ++d) st_async_f32(rb[d] + off, hn, rb[d] + boff + 8u * 4);
      }
      if (cn < nb) {
        const size_t row = row0 + cn;
        float* go = p.gates + row * H4 + S0 + chh;
        go[0] = ig; go[kH] = fg; go[2 * kH] = gg; go[3 * kH] = og;
        p.U[(row + B) * H4 + kH + S0 + chh] = hn;
        p.Cs[(row + B) * kH + S0 + chh] = c_reg;
      }
    }
    GSCAN3_STAMP(14);
    GSCAN3_STAMP(15);
  }

  if (warp < nb && rank == warp % kC) {
    p.beta_sum[(size_t)(b0 + warp) * kM + lane] = bs0;
    if (lane < kM - 32) p.beta_sum[(size_t)(b0 + warp) * kM + 32 + lane] = bs1;
  }
  // no CTA may exit while stores from its peers can still be in flight towards it
  cluster_barrier();
}

}  // namespace v3
}  // namespace gscan
