// Decoder sweeps, version 3 (reference hot loop: seq2seq_model.py:473-480 calling forward_step
// 359-428; backward = autograd of the same).  Compile-time specialised for the gSCAN paper shape
// H = 100, 6x6 grid: a cluster of 5 CTAs owns 8 examples for the whole sequence, CTA r owns the
// hidden slice [20r, 20r+20) of every H-sized quantity.
//
//  * Recurrent weights are resident for the whole sweep as tensor-core operand fragments: every
//    mat-vec stage is a [16 rows x K] x [K x 8 examples] product on mma.sync.m16n8k8 (tf32) in
//    split precision (3xTF32: W_hi x_hi + W_lo x_hi + W_hi x_lo, fp32 accumulate, ~2^-21 relative
//    error), the W_hi fragments live in registers, the W_lo fragments in shared memory
//    (v1 streamed ~640 KB/step/CTA from L2, v2 re-read 88 KB/step/CTA from shared memory; the
//    fp32 FFMA2 tile kept below for tools/ubench_mv2.cu is operand-bandwidth bound at ~50 FMA/clk/SM).
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

constexpr int kH = 100, kC = 5, kHS = 20, kM = 36, kG4 = 80, kNB = 8, kThreads = 512;
constexpr int kXS = 112;   // row stride of the gathered activation vectors (7 k-steps of 16; pad stays zero; 112 = 16 mod 32
                           // words: the 16-byte B-operand loads of two rows x four lanes touch every bank once)
constexpr int kKSteps = 13;   // k-steps of 8 of the tf32 tile (mv_tile: micro-benchmarks, backward sweep helpers)
constexpr int kK16 = 7;       // k-steps of 16 of the f16 tile (mv_tile16: the forward sweep)
constexpr int kGS = 84;    // row stride of the gate pre-activation scratch (bank spread)
constexpr int kMaxTi = 16;
constexpr int kXeBuf = kNB * kG4 + 4;   // one Xe staging buffer: [8][80] + four zero words (the "no Xe term" slot)
constexpr int kOutRow = 120;    // [gates i f g o: 80 | h_t: 20 | c_t: 20] per example, staged by the cell threads
constexpr int kIoWarp0 = 14, kIoThreads = 64;   // backward sweep: warps 14, 15 move the per-step global traffic
constexpr int kIoWarp0F = 8, kIoThreadsF = 256; // forward sweep: warps 8-15 (stage D / C owners and the spare warp, idle
                                                // outside their own stage) - one or two 16-byte accesses per thread and site
constexpr int kTlStamps = 32;   // phase stamps per step of the instrumented (TL) instantiations: 0-15 phases, 16+ sub-phases

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
// The bytes counted on the barrier were stored into THIS CTA's shared memory by st.async and become visible
// with the phase completion (the complete_tx contract, as for TMA), so the default cta-scope acquire is the
// right one: a cluster-scope acquire makes ptxas add an L1 invalidate (CCTL.IVALL) after every wait, which
// the ncu source view showed as ~17 % of all stall samples of the sweep.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
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

// ---- split-precision tensor-core mat-vec ----------------------------------------------------------
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// round-to-nearest tf32 part of x (ties away from zero), so that the remainder x - hi has either sign and
// |lo| <= 2^-12 |x|.  Integer add + mask on the ALU pipe: cvt.rna.tf32.f32 does the same but measured ~12 %
// slower sweeps on B200 (the conversion runs at a fraction of the ALU rate).
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x) { return __float_as_uint(x - __uint_as_float(tf32_hi(x))); }

// One 16-row tile: o = W[16 x 104] . x[8 examples][104]^T.  Lane (g = lane>>2, t = lane&3) supplies
// for k-step s the operand slots (k = t, t+4) from the physical columns (8s + 2t, 8s + 2t + 1), for A
// and B alike, so both come from 8-byte loads.  Result: o[0], o[1] = row g, examples 2t, 2t+1;
// o[2], o[3] = row g + 8, same examples.
__device__ __forceinline__ void mv_tile(const uint32_t (&whi)[kKSteps][4], const float4* __restrict__ wlo_lane,
                                        const float* __restrict__ x_lane, float (&o)[4]) {
  float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int s = 0; s < kKSteps; ++s) {
    const float2 xv = *reinterpret_cast<const float2*>(x_lane + 8 * s);
    const float4 lo = wlo_lane[s * 32];
    const uint32_t bh0 = tf32_hi(xv.x), bh1 = tf32_hi(xv.y);
    const uint32_t bl0 = __float_as_uint(xv.x - __uint_as_float(bh0)), bl1 = __float_as_uint(xv.y - __uint_as_float(bh1));
    mma_tf32(d0, whi[s][0], whi[s][1], whi[s][2], whi[s][3], bh0, bh1);
    mma_tf32(d1, __float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), __float_as_uint(lo.w), bh0, bh1);
    mma_tf32(d2, whi[s][0], whi[s][1], whi[s][2], whi[s][3], bl0, bl1);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = d0[j] + (d1[j] + d2[j]);
}

// ---- split-precision mat-vec on f16 operands (round 2, forward sweep) ----------------------------------------------
// Every mma.sync shape issues at 8.0 cycles per instruction per scheduler on B200, whatever it computes
// (tools/ubench_hmma_rates.cu: tf32 m16n8k8 = 1024 MAC, f16 m16n8k16 = 2048 MAC, same rate), and the three mat-vec
// stages of a step are bound by exactly that (tools/ubench_mvtile.cu).  An fp32 value splits into two f16 operands the
// same way it splits into two tf32 ones - 11 significant bits each: hi = f16(x), lo' = f16((x - hi) * 2^11), the
// scaling keeps lo' in f16's normal range - so the product costs 3 x 7 instructions of K = 16 instead of 3 x 13 of K = 8:
//     W x  ~=  W_hi x_hi + 2^-11 (W_lo' x_hi + W_hi x_lo')          (dropped: W_lo x_lo ~ 2^-22 relative, as before)
// Forward only: weights and activations (h, q' in (-1, 1); c_V a convex combination of keys) are far inside f16's
// range, and values below f16's normal range keep their precision through lo' (|x - hi| <= 2^-25 absolute).  The
// backward sweep multiplies gradients of arbitrary magnitude and stays on tf32.
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;
// (x0, x1) -> packed f16 pair of the rounded values (x0 in the low half) and of the scaled remainders
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  float h0, h1;
  asm("{\n.reg .f16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}" : "=f"(h0), "=f"(h1) : "r"(hi));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"((x1 - h1) * kLoScale), "f"((x0 - h0) * kLoScale));
}
// One 16-row tile: o = W[16 x 112] . x[8 examples][112]^T.  Lane (g = lane>>2, t = lane&3) supplies for k-step s the
// operand slots (k = 2t, 2t+1 | 2t+8, 2t+9) from the physical columns (16s + 4t, +1 | +2, +3), for A and B alike, so
// that B comes from ONE 16-byte load.  Result layout as mv_tile: o[0], o[1] = row g, examples 2t, 2t+1; o[2], o[3] = row g + 8.
// PIPE: loads of step s + 1 ahead of the MMAs of step s (training sweep: -7 us; the greedy sweep, with its other register
// budget, lost 1 % with it and keeps the plain order).
template <bool PIPE = true>
__device__ __forceinline__ void mv_tile16(const uint32_t (&whi)[kK16][4], const uint4* __restrict__ wlo_lane,
                                          const float* __restrict__ x_lane, float (&o)[4]) {
  float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
  // software-pipelined by one k-step: the two 16-byte shared-memory loads of step s + 1 are issued before the MMAs of
  // step s (left to itself the compiler put every load right in front of its first use: a load-to-use stall per k-step,
  // ~800 cycles per stage in the per-warp timeline against ~330 of tensor time)
  float4 xv = lds4(x_lane);
  uint4 lo = wlo_lane[0];
#pragma unroll
  for (int s = 0; s < kK16; ++s) {
    float4 xn = xv;
    uint4 ln = lo;
    if (PIPE && s + 1 < kK16) {
      xn = lds4(x_lane + 16 * (s + 1));
      ln = wlo_lane[(s + 1) * 32];
    }
    if (!PIPE && s > 0) {
      xv = lds4(x_lane + 16 * s);
      lo = wlo_lane[s * 32];
      xn = xv;
      ln = lo;
    }
    uint32_t bh0, bl0, bh1, bl1;
    split_f16x2(xv.x, xv.y, bh0, bl0);
    split_f16x2(xv.z, xv.w, bh1, bl1);
    mma_f16(d0, whi[s][0], whi[s][1], whi[s][2], whi[s][3], bh0, bh1);
    mma_f16(d1, lo.x, lo.y, lo.z, lo.w, bh0, bh1);
    mma_f16(d2, whi[s][0], whi[s][1], whi[s][2], whi[s][3], bl0, bl1);
    xv = xn;
    lo = ln;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = fmaf(kLoInv, d1[j] + d2[j], d0[j]);
}

// ---- shared-memory layout (float offsets) --------------------------------------------------------
struct FwdSmem {
  int hfull, qpfull, cvfull, xT, xV, P, KT, KV, qT, ch, qV, g, al, be, vT, vV, bc, len, bars, wlo, total;
  int xeTab, outE, wo, u, xL, tok;   // greedy decoding only
  int xe, cT, out;                   // training only: staging of Xe (double buffered), c_T, [gates | h | c] (see I/O warps)
};
__host__ __device__ inline FwdSmem fwd_smem(int Ti, int cond, int greedy_V = 0) {
  FwdSmem s{};
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  const int RBl = kHS * (4 + cond);
  s.hfull = take(kNB * kXS);
  s.qpfull = take(kNB * kXS);
  s.cvfull = take(kNB * kXS);
  s.xT = take(kC * kNB * kMaxTi);   // partial text scores of every rank, kMaxTi slots per example
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
  s.wlo = take(15 * kK16 * 32 * 4);   // W_lo' fragments: [role warp][k-step][lane] four packed f16 pairs
  if (greedy_V > 0) {
    const int V = greedy_V, Vp = (V + 3) & ~3;
    s.xeTab = take(V * kG4);       // this CTA's columns of Emb . W_ih[:, :H]^T + b_ih + b_hh
    s.outE = take(V * V);          // OutE[tok][v] = Wout[v, :H] . Emb[tok]
    s.wo = take(3 * kHS * Vp);     // rows of Wout[:, H:4H]^T for this CTA's slices of h, c_T, c_V
    s.u = take(kNB * 3 * kHS);     // [h | c_T | c_V] slices of the current step
    s.xL = take(kC * kNB * Vp);    // partial logits from every rank
    s.tok = take(3 * kNB + 4);     // tok, alive, flag (ints)
  } else {
    s.xe = take(2 * kXeBuf);       // Xe rows of this and of the next step (filled by cp.async one step ahead) + zero pad
    s.cT = take(kNB * kHS);        // c_T slice of this step
    s.out = take(kNB * kOutRow);   // activated gates, h_t, c_t of this step
  }
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
  long long* timeline;   // debug: [T][16 stamps][16 warps] clock64 stamps of CTA 0 (lane 0 of every warp), else null
  // progress signals (training sweep): every CTA adds 1 to progress[k] once all its stores of the steps t <= t_signal[k]
  // (row groups <= t_signal[k] + 1 of U) are visible: the output head of those rows then runs beside the rest of the sweep
  unsigned int* progress;   // [n_signals] words, zeroed by the host; null: no signals
  int n_signals;
  int t_signal[4];          // ascending
  // greedy decoding (predict.py:97-117); tables as in recurrent.cuh DecFwdP
  const float *XeTab, *OutE, *Wo_t;
  int V, Vp, sos, eos;
  long long* out_tokens;   // [B][T]
  int *out_len, *out_steps;
  float *g_alphas, *g_betas;   // [B][T][Ti], [B][T][M] or null
};

// clock64 stamps of CTA 0 / thread 0 between the phases of a step: compiled in only for the TL = true
// instantiations (GSCAN_TIMELINE=1), the production kernels carry no trace of them
#define GSCAN3_STAMP(k)                                                                              \
  do {                                                                                               \
    if constexpr (TL) {                                                                              \
      if (p.timeline && blockIdx.x == 0 && (threadIdx.x & 31) == 0)                                  \
        p.timeline[((size_t)t * kTlStamps + (k)) * 16 + (threadIdx.x >> 5)] = clock64();             \
    }                                                                                                \
  } while (0)

// 4-lane mat-vec tile: rows (2 per thread) x 7 k-quads (quad 4i+ks) x 8 examples.
// Returns in o[0..3] the full dot products of row (2*rp + (ks>>1)) for examples 4*(ks&1)+m.
__device__ __forceinline__ void mv_rowpair(const float4 (&w0)[7], const float4 (&w1)[7], const float* __restrict__ x,
                                           int ks, float (&o)[4]) {
  // examples are processed in two blocks of 4 with 16 independent FFMA2 chains each: the
  // dependent-issue latency of the packed FMA, not its throughput, bounds a 2-chain schedule
  float r0[kNB], r1[kNB];
#pragma unroll
  for (int nb4 = 0; nb4 < kNB; nb4 += 4) {
    float2 a0l[4], a0h[4], a1l[4], a1h[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) a0l[n] = a0h[n] = a1l[n] = a1h[n] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int q = min(4 * i + ks, kH / 4 - 1);   // clamped quads carry zero weights
      const float* xp = x + nb4 * kXS + 4 * q;
      float4 xv[4];
#pragma unroll
      for (int n = 0; n < 4; ++n) xv[n] = lds4(xp + n * kXS);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        fma2(a0l[n], lo2(w0[i]), lo2(xv[n]));
        fma2(a0h[n], hi2(w0[i]), hi2(xv[n]));
        fma2(a1l[n], lo2(w1[i]), lo2(xv[n]));
        fma2(a1h[n], hi2(w1[i]), hi2(xv[n]));
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      r0[nb4 + n] = (a0l[n].x + a0h[n].x) + (a0l[n].y + a0h[n].y);
      r1[nb4 + n] = (a1l[n].x + a1h[n].x) + (a1l[n].y + a1h[n].y);
    }
  }
  const bool up = (ks & 2) != 0, odd = (ks & 1) != 0;
  float keep[kNB];
#pragma unroll
  for (int n = 0; n < kNB; ++n) {
    const float mine = up ? r1[n] : r0[n], give = up ? r0[n] : r1[n];
    keep[n] = mine + __shfl_xor_sync(0xffffffffu, give, 2);
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const float mine = odd ? keep[4 + m] : keep[m], give = odd ? keep[m] : keep[4 + m];
    o[m] = mine + __shfl_xor_sync(0xffffffffu, give, 1);
  }
}

// 4-lane mat-vec tile, one row per thread: returns the dot products of the row for examples
// 4*(ks>>1) + 2*(ks&1) + {0, 1}
__device__ __forceinline__ void mv_row(const float4 (&w)[7], const float* __restrict__ x, int ks, float (&o)[2]) {
  float2 al[kNB], ah[kNB];
#pragma unroll
  for (int n = 0; n < kNB; ++n) al[n] = ah[n] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int q = min(4 * i + ks, kH / 4 - 1);
    const float* xp = x + 4 * q;
    float4 xv[kNB];
#pragma unroll
    for (int n = 0; n < kNB; ++n) xv[n] = lds4(xp + n * kXS);
#pragma unroll
    for (int n = 0; n < kNB; ++n) {
      fma2(al[n], lo2(w[i]), lo2(xv[n]));
      fma2(ah[n], hi2(w[i]), hi2(xv[n]));
    }
  }
  float r[kNB];
#pragma unroll
  for (int n = 0; n < kNB; ++n) r[n] = (al[n].x + ah[n].x) + (al[n].y + ah[n].y);
  const bool up = (ks & 2) != 0, odd = (ks & 1) != 0;
  float k4[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) k4[m] = (up ? r[4 + m] : r[m]) + __shfl_xor_sync(0xffffffffu, up ? r[m] : r[4 + m], 2);
#pragma unroll
  for (int m = 0; m < 2; ++m) o[m] = (odd ? k4[2 + m] : k4[m]) + __shfl_xor_sync(0xffffffffu, odd ? k4[m] : k4[2 + m], 1);
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

// Round 2: LANES threads per (example, key) pair, each covering 20 / LANES contiguous hidden units with vector loads
// (rows of K are 80 bytes: 16-byte aligned; lanes at a stride of 20 floats hit all 32 banks once per quarter-warp).
// The 4-lane form above spends ~145 warp-instructions per 32 items on scalar loads, index arithmetic and two
// shuffles - the visual scores alone were 29 % of all instructions issued per decoder step (ncu, profiles/r02_*);
// one thread per pair issues ~2.5x fewer and needs no shuffle.  LANES = 2 halves the dependent chain for the short
// text attention (8 x Ti pairs), where latency, not issue slots, is what counts.
template <int NKEYS_CT, int LANES>
__device__ __forceinline__ void partial_scores_vec(const float* __restrict__ q_s, const float* __restrict__ K_s,
                                                   const float* __restrict__ v_s, int nkeys, int xoff_floats, int rank,
                                                   uint32_t rb_k, uint32_t rb_4, uint32_t bar_off, int w0 = 0,
                                                   int nw = kThreads / 32) {
  // rb_k: shared-memory window of CTA (lane & 3) for LANES = 1, of CTA min(lane & 7, 4) for LANES = 2; rb_4: of CTA 4;
  // the items are spread over the warps w0 .. w0 + nw - 1 (whole warps: the shuffles below use the full mask)
  static_assert(LANES == 1 || LANES == 2, "20 hidden units per CTA: 1 x 20 or 2 x 10");
  const int N = NKEYS_CT > 0 ? NKEYS_CT : nkeys;
  const int total = kNB * N * LANES;
  const int wrel = (int)(threadIdx.x >> 5) - w0;
  if (wrel < 0 || wrel >= nw) return;
  for (int base = wrel * 32; base < total; base += nw * 32) {
    const int item = base + (threadIdx.x & 31);
    const int pair = item / LANES, u = item - pair * LANES;
    float s = 0.f;
    if (item < total) {
      const int n = pair / N;
      float s1 = 0.f;
      if (LANES == 1) {
        const float4* kp = reinterpret_cast<const float4*>(K_s + pair * kHS);
        const float4* qp = reinterpret_cast<const float4*>(q_s + n * kHS);
        const float4* vp = reinterpret_cast<const float4*>(v_s);
#pragma unroll
        for (int i = 0; i < kHS / 4; ++i) {
          const float4 k = kp[i], q = qp[i], v = vp[i];
          s = fmaf(v.x, act_tanh(q.x + k.x), s);
          s1 = fmaf(v.y, act_tanh(q.y + k.y), s1);
          s = fmaf(v.z, act_tanh(q.z + k.z), s);
          s1 = fmaf(v.w, act_tanh(q.w + k.w), s1);
        }
      } else {
        const float2* kp = reinterpret_cast<const float2*>(K_s + pair * kHS + 10 * u);
        const float2* qp = reinterpret_cast<const float2*>(q_s + n * kHS + 10 * u);
        const float2* vp = reinterpret_cast<const float2*>(v_s + 10 * u);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const float2 k = kp[i], q = qp[i], v = vp[i];
          s = fmaf(v.x, act_tanh(q.x + k.x), s);
          s1 = fmaf(v.y, act_tanh(q.y + k.y), s1);
        }
      }
      s += s1;
    }
    if (LANES == 2) s += __shfl_xor_sync(0xffffffffu, s, 1);
    // Four consecutive pairs travel as ONE 16-byte st.async per destination: every st.async is one transaction on the
    // receiver's mbarrier, and the exchanges of a step are bound by the NUMBER of those, not by their bytes
    // (kNB * N is a multiple of 4 and the score buffers are 16-byte aligned, so a group is all valid or all padding).
    const int lane = threadIdx.x & 31;
    constexpr int G = 4 * LANES;                  // lanes that hold one group of four pair sums
    const int g0 = lane & ~(G - 1), k = lane & (G - 1);
    float4 v;
    v.x = __shfl_sync(0xffffffffu, s, g0);
    v.y = __shfl_sync(0xffffffffu, s, g0 + LANES);
    v.z = __shfl_sync(0xffffffffu, s, g0 + 2 * LANES);
    v.w = __shfl_sync(0xffffffffu, s, g0 + 3 * LANES);
    if (item < total) {
      const uint32_t off = (uint32_t)(xoff_floats + rank * kNB * N + (base + g0) / LANES) * 4u;
      if (LANES == 1) {   // lane k of the group -> CTA k, lane 0 also -> CTA 4
        st_async_f32x4(rb_k + off, v, rb_k + bar_off);
        if (k == 0) st_async_f32x4(rb_4 + off, v, rb_4 + bar_off);
      } else if (k < kC) {   // lanes 0..4 of the group of eight -> CTA 0..4
        st_async_f32x4(rb_k + off, v, rb_k + bar_off);
      }
    }
  }
}

// Four lanes per (example, key) pair like partial_scores (the tanh work - 2 MUFU each - stays spread evenly over all 16
// warps / 4 SFU pipes; one thread per pair puts 3 of the 9 busy warps on one scheduler: measured slower), but lane u
// covers hidden units {4u .. 4u+3, 16+u}: one 16-byte and one 4-byte load per operand instead of five scalar ones,
// and four consecutive pair sums leave as ONE 16-byte st.async per destination (lanes 0..4 of each group of 16 lanes
// -> CTA 0..4): 360 instead of 1440 mbarrier transactions per receiver for the visual scores
// (tools/ubench_exchange.cu: 1207 -> 899 cycles per exchange).  rb_16 = window of CTA min(lane & 15, 4).
template <int NKEYS_CT>
__device__ __forceinline__ void partial_scores_q4(const float* __restrict__ q_s, const float* __restrict__ K_s,
                                                  const float* __restrict__ v_s, int nkeys, int xoff_floats, int rank,
                                                  uint32_t rb_16, uint32_t bar_off) {
  const int lane = threadIdx.x & 31, u = lane & 3;
  const int N = NKEYS_CT > 0 ? NKEYS_CT : nkeys;
  const int total = kNB * N * 4;   // a multiple of 32: whole warps
  const float4 v4 = lds4(v_s + 4 * u);
  const float v1 = v_s[16 + u];
  const int g0 = lane & 16, k = lane & 15;
  for (int base = (threadIdx.x >> 5) * 32; base < total; base += kThreads) {
    const int pair = (base + lane) >> 2;
    const int n = pair / N;
    const float* kp = K_s + pair * kHS;
    const float* qp = q_s + n * kHS;
    const float4 k4 = lds4(kp + 4 * u), q4 = lds4(qp + 4 * u);
    const float k1 = kp[16 + u], q1 = qp[16 + u];
    float s = v4.x * act_tanh(q4.x + k4.x);
    float s1 = v4.y * act_tanh(q4.y + k4.y);
    s = fmaf(v4.z, act_tanh(q4.z + k4.z), s);
    s1 = fmaf(v4.w, act_tanh(q4.w + k4.w), s1);
    s = fmaf(v1, act_tanh(q1 + k1), s);
    s += s1;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    float4 o;
    o.x = __shfl_sync(0xffffffffu, s, g0);
    o.y = __shfl_sync(0xffffffffu, s, g0 + 4);
    o.z = __shfl_sync(0xffffffffu, s, g0 + 8);
    o.w = __shfl_sync(0xffffffffu, s, g0 + 12);
    if (k < kC) {
      const uint32_t off = (uint32_t)(xoff_floats + rank * kNB * N + ((base + g0) >> 2)) * 4u;
      st_async_f32x4(rb_16 + off, o, rb_16 + bar_off);
    }
  }
}

// Text scores, one example per warp (warps w0 .. w0 + 7): lane = 2 * key + half, ten hidden units per lane.  No index
// arithmetic beyond shifts (the generic routine above spends ~270 warp-instructions per 32 items, a quarter of them on
// the division by the run-time Ti and on loop control), key slots >= Ti idle.  The partial scores travel in a layout
// padded to kMaxTi keys per example, four keys per 16-byte st.async: lanes 0..4 of each group of eight -> CTA 0..4.
// rb_8 = window of CTA min(lane & 7, 4).
__device__ __forceinline__ void text_scores_warp(const float* __restrict__ q_s, const float* __restrict__ K_s,
                                                 const float* __restrict__ v_s, int Ti, int xoff_floats, int rank,
                                                 uint32_t rb_8, uint32_t bar_off, int w0) {
  const int n = (int)(threadIdx.x >> 5) - w0, lane = threadIdx.x & 31;
  if (n < 0 || n >= kNB) return;
  const int j = lane >> 1, u = lane & 1;
  float s = 0.f;
  if (j < Ti) {
    const float2* kp = reinterpret_cast<const float2*>(K_s + (n * Ti + j) * kHS + 10 * u);
    const float2* qp = reinterpret_cast<const float2*>(q_s + n * kHS + 10 * u);
    const float2* vp = reinterpret_cast<const float2*>(v_s + 10 * u);
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float2 k = kp[i], q = qp[i], v = vp[i];
      s = fmaf(v.x, act_tanh(q.x + k.x), s);
      s1 = fmaf(v.y, act_tanh(q.y + k.y), s1);
    }
    s += s1;
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  const int g0 = lane & ~7, k = lane & 7;
  float4 o;
  o.x = __shfl_sync(0xffffffffu, s, g0);
  o.y = __shfl_sync(0xffffffffu, s, g0 + 2);
  o.z = __shfl_sync(0xffffffffu, s, g0 + 4);
  o.w = __shfl_sync(0xffffffffu, s, g0 + 6);
  if ((g0 >> 1) < Ti && k < kC) {
    const uint32_t off = (uint32_t)(xoff_floats + (rank * kNB + n) * kMaxTi + (g0 >> 1)) * 4u;
    st_async_f32x4(rb_8 + off, o, rb_8 + bar_off);
  }
}

template <bool COND, bool GREEDY, bool TL = false>
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
  const FwdSmem L = fwd_smem(Ti, COND ? 1 : 0, GREEDY ? p.V : 0);

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
  float* xeTab_s = smem + L.xeTab;
  float* outE_s = smem + L.outE;
  float* wo_s = smem + L.wo;
  float* u_s = smem + L.u;
  float* xL_s = smem + L.xL;
  int* tok_s = reinterpret_cast<int*>(smem + L.tok);
  int* alive_s = tok_s + kNB;
  int* flag_s = alive_s + kNB;

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar0 = smem_base + (uint32_t)L.bars * 4u;   // [0] xT  [1] qp  [2] xV  [3] cv  [4] h  [5] logits
  // shared-memory windows of the peer CTAs: one `mapa` where they are used (cheaper than seven registers held for the
  // whole sweep: the kernel sits at the 128-register limit)
#define RB(d) mapa_u32(smem_base, (uint32_t)(d))
#define RB_U RB(lane & 3)
#define RB_4 RB(4)
#define RB_8 RB(min(lane & 7, kC - 1))
  const uint32_t boff = (uint32_t)L.bars * 4u;

  // ---- resident weight fragments; the content depends on the warp's role ---------------------------------
  //   warps 0-7   stage A  tile w    of [q_T | W_c[:, :H] h (or q_V) | W_hh gates]   (120 rows, input h)
  //   warps 8-12  stage D  tile w-8  of the 80 gate rows of W_ih[:, 2H:3H]            (input c_V)
  //   warps 13-14 stage C  tile w-13 of W_qV (20 rows, conditional attention only)    (input q')
  const int fg = lane >> 2, ft = lane & 3;   // fragment coordinates
  const bool roleA = warp < 8, roleD = warp >= 8 && warp < 13, roleC = COND && (warp == 13 || warp == 14);
  uint32_t whi[kK16][4];
  uint4* wlo_lane = reinterpret_cast<uint4*>(smem + L.wlo) + (size_t)min(warp, 14) * kK16 * 32 + lane;
  int lr0 = 0;   // local output row of o[0], o[1]; o[2], o[3] belong to row lr0 + 8
  {
    const float *r0 = nullptr, *r1 = nullptr;
    auto rowA = [&](int lr) -> const float* {
      if (lr >= 6 * kHS) return nullptr;
      const int type = lr / kHS, i = lr - type * kHS, hr = S0 + i;
      if (type == 0) return p.W_qT + (size_t)hr * kH;
      if (type == 1) return COND ? p.W_c + (size_t)hr * 2 * kH : p.W_qV + (size_t)hr * kH;
      return p.W_hh + (size_t)((type - 2) * kH + hr) * kH;
    };
    auto rowD = [&](int lr) -> const float* {
      const int g = lr / kHS, i = lr - g * kHS;
      return p.W_ih + (size_t)(g * kH + S0 + i) * 3 * kH + 2 * kH;
    };
    auto rowC = [&](int lr) -> const float* { return lr < kHS ? p.W_qV + (size_t)(S0 + lr) * kH : nullptr; };
    if (roleA) {
      // tile 0 (the q_T rows, which the text scores wait for) on warp 7: the scheduler favours the higher warp index
      lr0 = 16 * (7 - warp) + fg;
      r0 = rowA(lr0);
      r1 = rowA(lr0 + 8);
    } else if (roleD) {
      lr0 = 16 * (warp - 8) + fg;
      r0 = rowD(lr0);
      r1 = rowD(lr0 + 8);
    } else if (roleC) {
      lr0 = 16 * (warp - 13) + fg;
      r0 = rowC(lr0);
      r1 = rowC(lr0 + 8);
    }
#pragma unroll
    for (int s = 0; s < kK16; ++s) {
      const int k = 16 * s + 4 * ft;   // four consecutive weights of rows lr0 and lr0 + 8 (rows are 16-byte aligned)
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 && k < kH) a = ldg4(r0 + k);
      if (r1 && k < kH) b = ldg4(r1 + k);
      uint32_t la, lb, lc, ld;
      split_f16x2(a.x, a.y, whi[s][0], la);
      split_f16x2(b.x, b.y, whi[s][1], lb);
      split_f16x2(a.z, a.w, whi[s][2], lc);
      split_f16x2(b.z, b.w, whi[s][3], ld);
      if (warp < 15) wlo_lane[s * 32] = make_uint4(la, lb, lc, ld);
    }
  }
  const int nF = 2 * ft;   // examples nF, nF + 1 are owned after a tile product

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
    for (int i = tid; i < kNB * kXS; i += kThreads) {
      const int n = i / kXS, h = i - n * kXS;
      const float hv = (n < nb && h < kH) ? __ldg(p.h_init + (size_t)(b0 + n) * kH + h) : 0.f;
      hfull_s[i] = hv;
      qpfull_s[i] = 0.f;
      cvfull_s[i] = 0.f;
      if (!GREEDY && n < nb && h >= S0 && h < S0 + kHS) p.U[(size_t)(b0 + n) * H4 + kH + h] = hv;   // row group 0: h_{-1}
    }
    for (int i = tid; i < kNB * kGS; i += kThreads) g_s[i] = 0.f;
    if (tid < kHS) {
      vT_s[tid] = __ldg(p.vT + S0 + tid);
      vV_s[tid] = __ldg(p.vV + S0 + tid);
      bc_s[tid] = COND ? __ldg(p.bc + S0 + tid) : 0.f;
    }
    if (tid < kNB) len_s[tid] = (tid < nb) ? max(1, min(p.cmd_len[b0 + tid], Ti)) : 1;
    if (GREEDY) {
      const int V = p.V, Vp = p.Vp;
      for (int i = tid; i < V * kG4; i += kThreads) {
        const int v = i / kG4, c = i - v * kG4;
        xeTab_s[i] = __ldg(p.XeTab + (size_t)v * H4 + (c / kHS) * kH + S0 + (c % kHS));
      }
      for (int i = tid; i < V * V; i += kThreads) outE_s[i] = __ldg(p.OutE + i);
      for (int i = tid; i < 3 * kHS * Vp; i += kThreads) {
        const int k = i / Vp, v = i - k * Vp;   // k = part * 20 + hidden offset
        wo_s[i] = __ldg(p.Wo_t + (size_t)((k / kHS) * kH + S0 + (k % kHS)) * Vp + v);
      }
      if (tid < kNB) {
        tok_s[tid] = p.sos;
        alive_s[tid] = tid < nb ? 1 : 0;
      }
    }
    if (tid == 0) {
      for (int k = 0; k < 6; ++k) mbar_init(bar0 + 8u * k, 1);
      fence_mbar_init();
    }
  }
  // cell state of (example tid/20, hidden S0 + tid%20), thread-private for the whole sweep
  const int cn = tid / kHS, chh = tid - cn * kHS;
  float c_reg = 0.f;
  if (tid < kNB * kHS) {
    if (cn < nb) {
      c_reg = __ldg(p.c_init + (size_t)(b0 + cn) * kH + S0 + chh);
      if (!GREEDY) p.Cs[(size_t)(b0 + cn) * kH + S0 + chh] = c_reg;
    }
  }
  int my_len = 0, my_steps = 0;   // greedy bookkeeping of lane n < kNB of warp 15
  float bs0 = 0.f, bs1 = 0.f;   // sum over steps of beta[warp][lane], beta[warp][lane + 32]
  // Round 2 (training sweep): everything a step reads from or writes to global memory goes through shared memory
  // and the two I/O warps (14, 15: stage C apart, idle) in 16-byte coalesced accesses.  Round 1 had every thread of the
  // stage-A epilogue, the softmaxes, the c_T / c_V combinations and the LSTM cell store its own 4-byte words (saved
  // activations for the backward pass) and prefetch its own 4 words of Xe: taking those accesses out of the kernel
  // shortened it from 0.88 to 0.75 ms (experiment) - address arithmetic and LSU slots on the critical
  // path of issue- and latency-bound phases.
  float* xe_s = smem + L.xe;
  float* cT_s = smem + L.cT;
  float* out_s = smem + L.out;
  const bool ioT = !GREEDY && warp >= kIoWarp0F;
  const int io = tid - kIoWarp0F * 32;
  // I/O threads [first, first + kNB * nq): one float4 each, shared src_s[n * sstride + 4q] -> global
  // dst[(grow0 + n) * gstride + S0 + col(q)], col(q) = 4q, or for the gate block (q / 5) * H + 4 (q % 5)
  auto io_rows = [&](int first, const float* src_s, int sstride, int nq, float* dst, size_t grow0, int gstride,
                     bool gate_cols) {
    const int f = io - first;
    if (!ioT || f < 0 || f >= kNB * nq) return;
    const int n = f / nq, q = f - n * nq;
    if (n < nb) {
      const int gcol = gate_cols ? (q / 5) * kH + 4 * (q % 5) : 4 * q;
      *reinterpret_cast<float4*>(dst + (grow0 + n) * gstride + S0 + gcol) = lds4(src_s + n * sstride + 4 * q);
    }
  };
  // Xe rows of step `t` -> staging buffer `buf` by the I/O threads [first, first + 160): 16-byte asynchronous copies,
  // awaited by xe_wait() before a later barrier
  auto xe_issue = [&](int first, int t, int buf) {
    const int f = io - first;
    if (!ioT) return;
    if (f >= 0 && f < kNB * (kG4 / 4)) {
      const int n = f / (kG4 / 4), q = f - n * (kG4 / 4);
      if (n < nb) {
        const uint32_t dst = smem_u32(xe_s + buf * kXeBuf + n * kG4 + 4 * q);
        const float* src = p.Xe + ((size_t)t * B + b0 + n) * H4 + (q / 5) * kH + S0 + 4 * (q % 5);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto xe_wait = [&]() {
    if (ioT) asm volatile("cp.async.wait_group 0;" ::: "memory");
  };
  // Stage-A epilogue of the training sweep: where each of the four tile results of this lane goes (float offset in
  // shared memory, low half) and which staged Xe word is added to it (offset in a staging buffer, high half; the zero
  // pad for rows that are not LSTM gates).  Decoded once: the epilogue is 4 x (LDS, FADD, STS) without branches
  // (round 1: row-type decode, divergent branches and 64-bit global addresses per result, ~800 cycles per step).
  uint32_t eo[4] = {0u, 0u, 0u, 0u};
  if (!GREEDY && roleA) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int lr = lr0 + 8 * (j >> 1), n = nF + (j & 1);
      const int type = lr / kHS, i = lr - type * kHS;
      int dst, x = kNB * kG4;   // zero pad
      if (type == 0) dst = L.qT + n * kHS + i;
      else if (type == 1) dst = (COND ? L.ch : L.qV) + n * kHS + i;
      else if (type < 6) { dst = L.g + n * kGS + lr - 2 * kHS; x = n * kG4 + lr - 2 * kHS; }
      else dst = L.g + n * kGS + kG4;   // rows 120..127 of the last tile: a pad column of the gate scratch
      eo[j] = (uint32_t)dst | ((uint32_t)x << 16);
    }
  }
  if (!GREEDY) {
    for (int i = tid; i < 2 * kXeBuf; i += kThreads) xe_s[i] = 0.f;   // rows of absent examples and the pads stay zero
    __syncthreads();
    xe_issue(0, 0, 0);
    xe_wait();
  }
  // all CTAs of the cluster must have initialised their barriers and buffers before any remote store
  __syncthreads();
  cluster_barrier();

  const uint32_t bytes_xT = (uint32_t)(kC * kNB * ((Ti + 3) / 4) * 16), bytes_vec = (uint32_t)(kNB * kH * 4),
                 bytes_xV = (uint32_t)(kC * kNB * kM * 4);
  bool finished_early = false;   // greedy: every sequence of this cluster ended before the step limit
  // greedy (predict.py:106-117), run by warp 15: wait for the logit partial sums of step s (X7), pick the token of each
  // live example, update tok / alive / the any-alive flag; lane-private my_len / my_steps of lanes < kNB.
  // Four lanes per example (lane = n + 8 u takes the vocabulary entries u, u + 4, ...): the pick runs beside the stage-A
  // mat-vecs of the next step and must not outlast them (one lane per example: 45 dependent shared-memory loads).
  auto greedy_pick = [&](int s) {
    mbar_wait(bar0 + 8u * 5, (uint32_t)(s & 1));
    if (lane == 0 && s + 1 < p.T) mbar_arm(bar0 + 8u * 5, (uint32_t)(kC * kNB * p.V * 4));   // for step s + 1
    const int n = lane & (kNB - 1), u = lane >> 3, V = p.V;
    const bool live = alive_s[n] != 0;
    const int tok = tok_s[n];
    float l[8];   // V <= 32 in greedy mode (checked on the host): <= 8 entries per lane
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = u + 4 * i;
      float a = -INFINITY;
      if (live && v < V) {
        a = outE_s[tok * V + v];
#pragma unroll
        for (int r = 0; r < kC; ++r) a += xL_s[(r * kNB + n) * V + v];
      }
      l[i] = a;
      mx = fmaxf(mx, a);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (live && u + 4 * i < V) sum += expf(l[i] - mx);
    sum += __shfl_xor_sync(0xffffffffu, sum, 8);
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    const float lse = logf(sum);
    float best = -INFINITY;   // first maximum of the log-softmax values, as F.log_softmax(...).max(dim=-1) gives
    int arg = 1 << 30;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = u + 4 * i;
      if (live && v < V) {
        const float lp = (l[i] - mx) - lse;
        if (lp > best) { best = lp; arg = v; }
      }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    __syncwarp();   // every lane has read alive / tok of its example
    int alive_now = 0;
    if (lane < kNB) {
      if (live) {
        my_steps++;
        if (arg == p.eos) {
          alive_s[n] = 0;
        } else {
          if (rank == 0) p.out_tokens[(size_t)(b0 + n) * p.T + my_len] = arg;
          my_len++;
        }
        tok_s[n] = arg;
      }
      alive_now = alive_s[n];
    }
    const unsigned any = __ballot_sync(0xffffffffu, alive_now != 0);
    if (lane == 0) flag_s[0] = any != 0u;
  };

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
      if (GREEDY && t == 0) mbar_arm(bar0 + 8u * 5, (uint32_t)(kC * kNB * p.V * 4));   // later steps: armed by greedy_pick
    }
    GSCAN3_STAMP(1);
    // ---- stage A: everything that depends only on h_{t-1} ---------------------------------------------
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (roleA) mv_tile16<!GREEDY>(whi, wlo_lane, hfull_s + fg * kXS + 4 * ft, o);
    GSCAN3_STAMP(16);
    if (GREEDY) {
      // The token of the previous step is picked HERE, by the otherwise idle warp 15, while the role-A warps run the
      // mat-vecs on h (which do not depend on the token); only the embedding term added below needs it.
      if (warp == 15 && t > 0) greedy_pick(t - 1);
      __syncthreads();
      if (t > 0 && !flag_s[0]) {   // the same decision in every CTA of the cluster: all of them computed the same tokens
        finished_early = true;
        break;
      }
    }
    if (roleA) {
      if (GREEDY) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int lr = lr0 + 8 * (j >> 1), n = nF + (j & 1);
          const int type = lr / kHS, i = lr - type * kHS;
          if (type == 0) qT_s[n * kHS + i] = o[j];
          else if (type == 1) (COND ? ch_s : qV_s)[n * kHS + i] = o[j];
          else if (type < 6) g_s[n * kGS + lr - 2 * kHS] = o[j] + xeTab_s[tok_s[n] * kG4 + lr - 2 * kHS];
        }
      } else {
        const float* xb = xe_s + (t & 1) * kXeBuf;
#pragma unroll
        for (int j = 0; j < 4; ++j) smem[eo[j] & 0xffffu] = o[j] + xb[eo[j] >> 16];
      }
      GSCAN3_STAMP(17);
    }
    GSCAN3_STAMP(18);
    // Only q_T (tiles 0 and 1: warps 7 and 6) is needed now; W_c h is first read after the text softmax, the gate rows by
    // the cell.  Training sweep with conditional attention: no block barrier here - warps 6 and 7 hand q_T over to warps
    // 8-15 (named barrier 2), which form the text scores and start the X1 exchange while the other stage-A tiles finish.
    constexpr bool kEarlyText = COND && !GREEDY;
    if (kEarlyText) {
      if (warp == 6 || warp == 7) asm volatile("bar.arrive 2, 320;" ::: "memory");
      else if (warp >= 8) asm volatile("bar.sync 2, 320;" ::: "memory");
    } else {
      __syncthreads();
    }
    GSCAN3_STAMP(2);
    if (!GREEDY) {
      // I/O warps: q_T of this step (and q_V, q' = h_{t-1} without conditional attention)
      io_rows(0, qT_s, kHS, 5, p.qT, row0, kH, false);
      if (!COND) {
        io_rows(40, qV_s, kHS, 5, p.qV, row0, kH, false);
        io_rows(80, hfull_s + S0, kXS, 5, p.Qp, row0, kH, false);
      }
    }
    // ---- textual attention: partial scores over the local slice, summed over ranks (X1) ---------------
    text_scores_warp(qT_s, KT_s, vT_s, Ti, L.xT, rank, RB_8, boff + 0u, kEarlyText ? 8 : 0);
    GSCAN3_STAMP(3);
    mbar_wait(bar0 + 8u * 0, par);
    GSCAN3_STAMP(4);
    if (warp < kNB) {
      const int n = warp;
      float s = -INFINITY;
      if (lane < Ti) {
        s = 0.f;
#pragma unroll
        for (int r = 0; r < kC; ++r) s += xT_s[(r * kNB + n) * kMaxTi + lane];
        if (lane >= len_s[n]) s = -INFINITY;
      }
      GSCAN3_STAMP(24);
      const float mx = warp_max_redux(s);
      GSCAN3_STAMP(25);
      const float e = (lane < Ti) ? __expf(s - mx) : 0.f;
      const float sum = warp_sum(e);
      GSCAN3_STAMP(26);
      const float a = e * (1.0f / sum);
      if (lane < Ti) {
        al_s[n * Ti + lane] = a;
        if (GREEDY && p.g_alphas && n < nb && rank == n % kC && alive_s[n])
          p.g_alphas[((size_t)(b0 + n) * p.T + t) * Ti + lane] = a;
      }
    }
    GSCAN3_STAMP(19);
    __syncthreads();
    GSCAN3_STAMP(5);
    if (ioT) {   // alpha rows of the examples this rank writes (Ti floats each: no 16-byte granularity in general)
      for (int f = io; f < kNB * Ti; f += kIoThreadsF) {
        const int n = f / Ti;
        if (n < nb && rank == n % kC) p.alpha[(row0 + n) * Ti + (f - n * Ti)] = al_s[f];
      }
    }
    // ---- everything linear in c_T through P_j = W K^T_j: q' slice (X3), gate contributions, c_T slice ----
    // Only q' (5 float4 per example: 40 outputs) is on the critical path of the step - it feeds the X3 exchange and
    // stage C; the gate contributions and the c_T slice (25 float4 per example) are first read by the cell / the I/O
    // warps, two block barriers later.  The 40 critical outputs get four lanes each (positions j = u, u + 4, ...:
    // three terms at Ti = 10 instead of ten), two shuffle rounds, and the five sends spread over the four lanes; the 200
    // others run one thread each on warps 5-11, off the critical path.  (One thread per output for all 240, as in
    // round 1, made this phase ~1,700 cycles of a 12,000-cycle step in the per-warp timeline.  Putting the critical
    // group on the warps the scheduler favours, 11-15, measured slower: 0.741 vs 0.721 ms.)
    if (COND && tid < 4 * kNB * 5) {
      const int idx = tid >> 2, u = tid & 3;
      const int n = idx / 5, q = idx - n * 5;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = P_s + (size_t)n * Ti * RBl + 4 * q;
#pragma unroll 4
      for (int j = u; j < Ti; j += 4) {
        const float a = al_s[n * Ti + j];
        const float4 v = lds4(src + j * RBl);
        o.x = fmaf(a, v.x, o.x); o.y = fmaf(a, v.y, o.y); o.z = fmaf(a, v.z, o.z); o.w = fmaf(a, v.w, o.w);
      }
#pragma unroll
      for (int sh = 1; sh <= 2; sh <<= 1) {
        o.x += __shfl_xor_sync(0xffffffffu, o.x, sh);
        o.y += __shfl_xor_sync(0xffffffffu, o.y, sh);
        o.z += __shfl_xor_sync(0xffffffffu, o.z, sh);
        o.w += __shfl_xor_sync(0xffffffffu, o.w, sh);
      }
      GSCAN3_STAMP(23);
      const float4 chv = lds4(ch_s + n * kHS + 4 * q);
      const float4 bcv = lds4(bc_s + 4 * q);
      float4 qq;
      qq.x = act_tanh(chv.x + o.x + bcv.x);
      qq.y = act_tanh(chv.y + o.y + bcv.y);
      qq.z = act_tanh(chv.z + o.z + bcv.z);
      qq.w = act_tanh(chv.w + o.w + bcv.w);
      const uint32_t off = (uint32_t)(L.qpfull + n * kXS + S0 + 4 * q) * 4u;
      const uint32_t wu = RB_U;   // lane u of the group -> CTA u, lane 0 also -> CTA 4
      st_async_f32x4(wu + off, qq, wu + boff + 8u * 1);
      if (u == 0) {
        const uint32_t w4 = RB_4;
        st_async_f32x4(w4 + off, qq, w4 + boff + 8u * 1);
      }
    } else if (tid >= 4 * kNB * 5 && tid < 4 * kNB * 5 + kNB * QB) {
      // (without conditional attention QB = 20: the gate contributions and c_T only, nothing critical)
      const int idx = tid - 4 * kNB * 5;
      const int n = idx / QB, q = (COND ? 5 : 0) + idx - n * QB;   // q in [5, 30) with COND, [0, 20) + c_T below without
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = (q < QB) ? P_s + (size_t)n * Ti * RBl + 4 * q : KT_s + (size_t)n * Ti * kHS + 4 * (q - QB);
      const int stride = (q < QB) ? RBl : kHS;
#pragma unroll 2
      for (int j = 0; j < Ti; ++j) {
        const float a = al_s[n * Ti + j];
        const float4 v = lds4(src + j * stride);
        o.x = fmaf(a, v.x, o.x); o.y = fmaf(a, v.y, o.y); o.z = fmaf(a, v.z, o.z); o.w = fmaf(a, v.w, o.w);
      }
      if (q < QB) {
        float4* gp = reinterpret_cast<float4*>(g_s + n * kGS + 4 * q - (COND ? kHS : 0));
        float4 gv = *gp;
        gv.x += o.x; gv.y += o.y; gv.z += o.z; gv.w += o.w;
        *gp = gv;
      } else if (GREEDY) {
        *reinterpret_cast<float4*>(u_s + n * 3 * kHS + kHS + 4 * (q - QB)) = o;
      } else {
        *reinterpret_cast<float4*>(cT_s + n * kHS + 4 * (q - QB)) = o;
      }
    }
    if (!COND && tid >= 4 * kNB * 5 + kNB * QB && tid < 4 * kNB * 5 + kNB * QB + kNB * 5) {
      // c_T slice without conditional attention (the five float4 per example that COND covers as q in [25, 30) above)
      const int idx = tid - (4 * kNB * 5 + kNB * QB);
      const int n = idx / 5, q = idx - n * 5;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = KT_s + (size_t)n * Ti * kHS + 4 * q;
      for (int j = 0; j < Ti; ++j) {
        const float a = al_s[n * Ti + j];
        const float4 v = lds4(src + j * kHS);
        o.x = fmaf(a, v.x, o.x); o.y = fmaf(a, v.y, o.y); o.z = fmaf(a, v.z, o.z); o.w = fmaf(a, v.w, o.w);
      }
      if (GREEDY) *reinterpret_cast<float4*>(u_s + n * 3 * kHS + kHS + 4 * q) = o;
      else *reinterpret_cast<float4*>(cT_s + n * kHS + 4 * q) = o;
    }
    GSCAN3_STAMP(6);
    // ---- stage C: visual query slice ---------------------------------------------------------------------
    if (COND) {
      mbar_wait(bar0 + 8u * 1, par);
      GSCAN3_STAMP(7);
      if (roleC) {
        float o[4];
        mv_tile16<!GREEDY>(whi, wlo_lane, qpfull_s + fg * kXS + 4 * ft, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = lr0 + 8 * (j >> 1), n = nF + (j & 1);
          if (r < kHS) qV_s[n * kHS + r] = o[j];
        }
      }
    }
    __syncthreads();
    GSCAN3_STAMP(8);
    if (!GREEDY) {
      // I/O warps: c_T (row group t + 1 of U), and with conditional attention q_V and q' (this CTA's slice of q' came
      // back through its own X3 store)
      io_rows(0, cT_s, kHS, 5, p.U + 2 * kH, row0 + B, H4, false);
      if (COND) {
        io_rows(40, qV_s, kHS, 5, p.qV, row0, kH, false);
        io_rows(80, qpfull_s + S0, kXS, 5, p.Qp, row0, kH, false);
      }
      // the Xe rows of the next step into the other staging buffer (last read by the stage-A epilogue of step t - 1)
      if (t + 1 < p.T) xe_issue(96, t + 1, (t + 1) & 1);
    }
    // ---- visual attention: partial scores (X4), softmax, c_V slice gathered (X5) -------------------------
    partial_scores_vec<kM, 1>(qV_s, KV_s, vV_s, kM, L.xV, rank, RB_U, RB_4, boff + 8u * 2);
    GSCAN3_STAMP(9);
    mbar_wait(bar0 + 8u * 2, par);
    GSCAN3_STAMP(10);
    if (warp < kNB) {
      const int n = warp;
      float s0 = 0.f, s1 = -INFINITY;
#pragma unroll
      for (int r = 0; r < kC; ++r) s0 += xV_s[(r * kNB + n) * kM + lane];
      if (lane < kM - 32) {
        s1 = 0.f;
#pragma unroll
        for (int r = 0; r < kC; ++r) s1 += xV_s[(r * kNB + n) * kM + 32 + lane];
      }
      const float mx = warp_max_redux(fmaxf(s0, s1));
      const float e0 = __expf(s0 - mx), e1 = (lane < kM - 32) ? __expf(s1 - mx) : 0.f;
      const float inv = 1.0f / warp_sum(e0 + e1);
      const float w0 = e0 * inv, w1 = e1 * inv;
      const bool counted = !GREEDY || alive_s[n] != 0;   // predict.py sums beta over generated steps only
      be_s[n * kM + lane] = w0;
      if (counted) bs0 += w0;
      if (lane < kM - 32) {
        be_s[n * kM + 32 + lane] = w1;
        if (counted) bs1 += w1;
      }
      if (GREEDY && p.g_betas && n < nb && rank == n % kC && counted) {
        float* gb = p.g_betas + ((size_t)(b0 + n) * p.T + t) * kM;
        gb[lane] = w0;
        if (lane < kM - 32) gb[32 + lane] = w1;
      }
    }
    GSCAN3_STAMP(20);
    __syncthreads();
    GSCAN3_STAMP(21);
    if (ioT) {   // beta rows (36 floats = 9 x 16 bytes) of the examples this rank writes
      if (io < kNB * (kM / 4)) {
        const int n = io / (kM / 4), q = io - n * (kM / 4);
        if (n < nb && rank == n % kC)
          *reinterpret_cast<float4*>(p.beta + (row0 + n) * kM + 4 * q) = lds4(be_s + n * kM + 4 * q);
      }
    }
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
      const uint32_t wu = RB_U;
      st_async_f32x4(wu + off, o, wu + boff + 8u * 3);
      if (u == 0) {
        const uint32_t w4 = RB_4;
        st_async_f32x4(w4 + off, o, w4 + boff + 8u * 3);
      }
      if (GREEDY && u == 1) *reinterpret_cast<float4*>(u_s + n * 3 * kHS + 2 * kHS + 4 * hq) = o;
    }
    GSCAN3_STAMP(11);
    mbar_wait(bar0 + 8u * 3, par);
    GSCAN3_STAMP(12);
    // ---- stage D: c_V contribution to the gates, then the LSTM cell ---------------------------------------
    if (roleD) {
      float o[4];
      mv_tile16<!GREEDY>(whi, wlo_lane, cvfull_s + fg * kXS + 4 * ft, o);
#pragma unroll
      for (int j = 0; j < 4; ++j) g_s[(nF + (j & 1)) * kGS + lr0 + 8 * (j >> 1)] += o[j];
    }
    GSCAN3_STAMP(22);
    __syncthreads();
    GSCAN3_STAMP(13);
    // I/O warps: c_V (this CTA's slice came back through its own X5 store), row group t + 1 of U
    if (!GREEDY) {
      io_rows(0, cvfull_s + S0, kXS, 5, p.U + 3 * kH, row0 + B, H4, false);
      xe_wait();   // issued after stage C; published by the barriers before the next stage-A epilogue
    }
    if (tid < kNB * kHS) {
      const float* gp = g_s + cn * kGS + chh;
      const float ig = act_sigmoid(gp[0]), fg = act_sigmoid(gp[kHS]), gg = act_tanh(gp[2 * kHS]), og = act_sigmoid(gp[3 * kHS]);
      c_reg = fmaf(fg, c_reg, ig * gg);
      const float hn = og * act_tanh(c_reg);
      if (t + 1 < p.T) {
        // four consecutive hidden units of one example (20 per example: a group never straddles two) as one 16-byte store
        // per destination: lane k of the group -> CTA k, lane 0 also -> CTA 4 (tid < 160 = five full warps)
        const int g0 = lane & ~3, k = lane & 3;
        float4 v;
        v.x = __shfl_sync(0xffffffffu, hn, g0);
        v.y = __shfl_sync(0xffffffffu, hn, g0 + 1);
        v.z = __shfl_sync(0xffffffffu, hn, g0 + 2);
        v.w = __shfl_sync(0xffffffffu, hn, g0 + 3);
        const uint32_t off = (uint32_t)(L.hfull + cn * kXS + S0 + (chh & ~3)) * 4u;
        const uint32_t wu = RB_U;
        st_async_f32x4(wu + off, v, wu + boff + 8u * 4);
        if (k == 0) {
          const uint32_t w4 = RB_4;
          st_async_f32x4(w4 + off, v, w4 + boff + 8u * 4);
        }
      }
      if (GREEDY) u_s[cn * 3 * kHS + chh] = hn;
      if (!GREEDY) {   // staged; written out by the I/O warps after the first barrier of the next step (or after the loop)
        float* op = out_s + cn * kOutRow + chh;
        op[0] = ig; op[kHS] = fg; op[2 * kHS] = gg; op[3 * kHS] = og; op[4 * kHS] = hn; op[5 * kHS] = c_reg;
        asm volatile("bar.arrive 1, %0;" ::"n"(kNB * kHS + kIoThreadsF) : "memory");   // hand-off to the I/O warps
      }
    }
    if (ioT) {
      // the I/O warps take the staged cell outputs as soon as the 160 cell threads have written them (named barrier 1:
      // nobody else waits) and write them out in the shadow of the next step's stage A
      asm volatile("bar.sync 1, %0;" ::"n"(kNB * kHS + kIoThreadsF) : "memory");
      io_rows(0, out_s, kOutRow, 20, p.gates, row0, H4, true);
      io_rows(160, out_s + 4 * kHS, kOutRow, 5, p.U + kH, row0 + B, H4, false);   // h_t: row group t + 1 of U
      io_rows(200, out_s + 5 * kHS, kOutRow, 5, p.Cs, row0 + B, kH, false);      // c_t: row group t + 1 of Cs
      // Progress signal: these were the last stores of step t, and every store of the step was issued by an I/O thread.
      // Fence by the writers, a barrier among the 256 I/O threads only (named barrier 3), one add per CTA - all of it in
      // the I/O warps' shadow.  (At a block barrier of the next step the same code cost the sweep 18 us: a fence in the
      // instruction stream of the compute warps, taken or not, pins the memory operations around it.)
      if (p.progress != nullptr) {
        int sig = -1;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < p.n_signals && t == p.t_signal[k]) sig = k;
        if (sig >= 0) {
          __threadfence();
          asm volatile("bar.sync 3, %0;" ::"n"(kIoThreadsF) : "memory");
          if (io == 0) atomicAdd(p.progress + sig, 1u);
        }
      }
    }
    GSCAN3_STAMP(14);
    if (GREEDY) {
      // ---- logits = OutE[tok] + Wout[:, H:4H] . [h; c_T; c_V]: partial sums over this CTA's slices (X7), argmax, feed back ----
      __syncthreads();
      const int V = p.V, Vp = p.Vp;
      const int total = kNB * V * 4;
      for (int base = warp * 32; base < total; base += kThreads) {
        const int item = base + lane, pair = item >> 2, u = item & 3;
        float s = 0.f;
        if (item < total) {
          const int n = pair / V, v = pair - n * V;
          const float* up = u_s + n * 3 * kHS + 15 * u;
          const float* wp = wo_s + (15 * u) * Vp + v;
#pragma unroll
          for (int i = 0; i < 15; ++i) s = fmaf(up[i], wp[i * Vp], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (item < total) {
          const uint32_t off = (uint32_t)(L.xL + rank * kNB * V + pair) * 4u;
          const uint32_t wu = RB_U;
          st_async_f32(wu + off, s, wu + boff + 8u * 5);
          if (u == 0) {
            const uint32_t w4 = RB_4;
            st_async_f32(w4 + off, s, w4 + boff + 8u * 5);
          }
        }
      }
    }
    GSCAN3_STAMP(15);
  }
  if (GREEDY) {
    // (an early exit happens right after the h of the last executed step has been received, before anything of the
    //  next step is sent, so nothing is in flight towards this CTA; after a full run the last token is still to be picked)
    if (!finished_early && warp == 15) greedy_pick(p.T - 1);
    if (rank == 0 && warp == 15 && lane < nb) {
      p.out_len[b0 + lane] = my_len;
      p.out_steps[b0 + lane] = my_steps;
    }
  }

  if (warp < nb && rank == warp % kC) {
    p.beta_sum[(size_t)(b0 + warp) * kM + lane] = bs0;
    if (lane < kM - 32) p.beta_sum[(size_t)(b0 + warp) * kM + 32 + lane] = bs1;
  }
  // no CTA may exit while stores from its peers can still be in flight towards it
  cluster_barrier();
}

#undef RB
#undef RB_U
#undef RB_4
#undef RB_8

}  // namespace v3
}  // namespace gscan
