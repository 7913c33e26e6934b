// Decoder backward sweep, version 3 (BPTT of seq2seq_model.py:359-428, reverse time; recipe:
// SURVEY.md A.6 / oracle/manual_backward.py stage B4).  Same organisation as the forward sweep in
// decoder_v3.cuh: a cluster of 5 CTAs owns 8 examples, CTA r owns the hidden slice [20r, 20r+20) of
// every H-sized gradient, CTAs exchange with st.async + mbarrier, mat-vecs run on mma.sync 3xTF32.
//
// The transposed products are INPUT-sliced: CTA r multiplies its own slice of da / dd / dq_V / dq_T
// by the matching columns of W^T and the partial results are reduce-scattered to the owners
//     X_e  partial dc_V   = W_ih[:, 2H:3H]^T da          [8][100] -> owners
//     X_a  partial dbeta  = dc_V . K^V_m over the slice  [8][36]  -> all
//     X_b  partial dq'    = W_qV^T dq_V                  [8][100] -> owners
//     X_c  partial dalpha (through P_j = W K^T_j)        [8][Ti]  -> all
//     X_d  partial dh     = W_hh^T da + W_c[:, :H]^T dd + W_qT^T dq_T   [8][100] -> owners
// What does not have to be inside the recurrence is left to batched kernels after the sweep: the
// "value path" of both attentions, dK_m += w_m dc (dc_T, dc_V are linear in the saved da, dd, dU).
#pragma once
#include "decoder_v3.cuh"

namespace gscan {
namespace v3 {

constexpr int kDaS = 88;    // row stride of da_s (80 gate pre-activation gradients + pad)
constexpr int kVs = 24;     // row stride of the 20-wide mma B operands (dq_V, dd, dq_T); pad stays zero
constexpr int kTiles = 7;   // ceil(100 / 16) output tiles of the transposed products
// weight units (16 rows x 8 columns each) per output tile, in shared memory fragment order
constexpr int kUcV = 0, kUhh = 10, kUc = 20, kUqT = 23, kUqV = 26, kUnitsPerTile = 29;
constexpr int kMaxTiB = 16;   // text positions per attention thread: 8 registers
// Saved activations of a step are staged through shared memory by the two I/O warps (14, 15): one row of kStageRow
// floats per example, filled with 16-byte coalesced loads one step ahead, double buffered
//   [gates i f g o: 80 | c_{t-1}: 20 | dU_h: 20 | dU_cT: 20 | dU_cV: 20 | q': 20 | q_V: 20 | q_T: 20 | beta: 36]
constexpr int kStageRow = 256, kStageQ = kStageRow / 4;
constexpr int kSgC = 80, kSgDUh = 100, kSgDUcT = 120, kSgDUcV = 140, kSgQp = 160, kSgQV = 180, kSgQT = 200, kSgBeta = 220;


struct BwdSmem {
  int W, KV, KT, P, da, dqV, dd, dqT, dUcT, dhpart, dcV, xcV, xbe, xqp, xal, xdh, drV, drT, a1, vT, vV, len, bars, stage, total;
};
__host__ __device__ inline BwdSmem bwd_smem(int Ti, int cond) {
  BwdSmem s{};
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  const int RBl = kHS * (4 + cond);
  s.W = take(kTiles * kUnitsPerTile * 32 * 4);
  s.KV = take(kNB * kM * kHS);
  s.KT = take(kNB * Ti * kHS);
  s.P = take(kNB * Ti * RBl);
  s.da = take(kNB * kDaS);
  s.dqV = take(kNB * kVs);
  s.dd = take(kNB * kVs);
  s.dqT = take(kNB * kVs);
  s.dUcT = take(kNB * kHS);
  s.dhpart = take(kNB * kH);
  s.dcV = take(kNB * kHS);
  s.xcV = take(kC * kNB * kHS);
  s.xbe = take(kC * kNB * kM);
  s.xqp = take(kC * kNB * kHS);
  s.xal = take(kC * kNB * Ti);
  s.xdh = take(kC * kNB * kHS);
  s.drV = take(kNB * kM);
  s.drT = take(kNB * Ti);
  s.a1 = take(kNB * Ti);
  s.vT = take(kHS);
  s.vV = take(kHS);
  s.len = take(kNB);
  s.bars = take(16);
  s.stage = take(2 * kNB * kStageRow);
  s.total = o;
  return s;
}

struct DecBwd3P {
  int B, T, Ti;
  const float *W_ih, *W_hh, *W_qV, *W_c, *W_qT;   // original row-major parameters
  const float* PT;                                // [Ti][B][RB] (see DecFwd3P)
  const float *vT, *vV;
  const float *KT, *KV;
  const int* cmd_len;
  const float *Cs, *gates, *alpha, *beta, *Qp, *qT, *qV;   // saved by the forward sweep
  const float* dU;          // [T][B][4H] gradient of [e | h | c_T | c_V] from the output projection
  const float* dbeta_aux;   // [B][M] or null
  float *dgates, *dd, *dqV, *dqT;   // [T][B][4H], [T][B][H] x3: operands of the weight-gradient GEMMs
  float *dKT, *dKV;         // key-path part only; the value path is added after the sweep
  float* dh0;               // [B][H]
  float *dvT, *dvV;         // [H] each, atomically accumulated (zeroed by the host)
  long long* timeline;
  // progress signals: every CTA adds 1 to progress[k] once all its global stores of the steps t >= t_signal[k] are visible
  // (the host lets work on those rows start in the shadow of the rest of the sweep: cuStreamWaitValue32)
  const unsigned int* fwd_tag;   // workspace word the forward call set to the sweep version it ran (3 = decoder_v3.cuh), or null
  unsigned int* progress;   // [n_signals] words, zeroed by the host
  int n_signals;
  int t_signal[4];          // descending
};

// o = W_tile[16 x 8*NSTEPS] . x^T: fp32 fragments from shared memory, split into tf32 hi/lo on the fly.
// Truncation split: the tensor core reads only the upper 19 bits of a tf32 operand, so the fp32 word ITSELF is the hi
// operand (no instruction) and lo = x - trunc(x) costs one LOP + one FADD (round 1 rounded hi to nearest: three
// instructions per element, ~2,400 warp-instructions per decoder step spent splitting constant weights).  |lo| < 2^-10 |x|
// and the hardware truncates lo to 11 bits in turn: 2^-20 relative per product term instead of 2^-22.
__device__ __forceinline__ uint32_t tf32_lo_trunc(float x) {
  return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
}
template <int NSTEPS>
__device__ __forceinline__ void mv_units(const float4* __restrict__ w_lane, const float* __restrict__ x_lane,
                                         float (&o)[4]) {
  float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int s = 0; s < NSTEPS; ++s) {
    const float2 xv = *reinterpret_cast<const float2*>(x_lane + 8 * s);
    const float4 a = w_lane[s * 32];
    const uint32_t ah0 = __float_as_uint(a.x), ah1 = __float_as_uint(a.y), ah2 = __float_as_uint(a.z), ah3 = __float_as_uint(a.w);
    const uint32_t al0 = tf32_lo_trunc(a.x), al1 = tf32_lo_trunc(a.y), al2 = tf32_lo_trunc(a.z), al3 = tf32_lo_trunc(a.w);
    const uint32_t bh0 = __float_as_uint(xv.x), bh1 = __float_as_uint(xv.y);
    const uint32_t bl0 = tf32_lo_trunc(xv.x), bl1 = tf32_lo_trunc(xv.y);
    mma_tf32(d0, ah0, ah1, ah2, ah3, bh0, bh1);
    mma_tf32(d1, al0, al1, al2, al3, bh0, bh1);
    mma_tf32(d2, ah0, ah1, ah2, ah3, bl0, bl1);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = d0[j] + (d1[j] + d2[j]);
}

template <bool COND, bool TL = false>
__global__ void __cluster_dims__(kC, 1, 1) __launch_bounds__(kThreads, 1) dec_bwd_v3_kernel(DecBwd3P p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RBl = kHS * (4 + (COND ? 1 : 0));
  constexpr int COFF = COND ? kHS : 0;   // first gate column of a P row
  constexpr int H4 = 4 * kH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int B = p.B, Ti = p.Ti;
  const int b0 = (blockIdx.x / kC) * kNB;
  const int nb = min(kNB, B - b0);
  const int S0 = rank * kHS;
  const BwdSmem L = bwd_smem(Ti, COND ? 1 : 0);

  float4* W_s = reinterpret_cast<float4*>(smem + L.W);
  float* KV_s = smem + L.KV;
  float* KT_s = smem + L.KT;
  float* P_s = smem + L.P;
  float* da_s = smem + L.da;
  float* dqV_s = smem + L.dqV;
  float* dd_s = smem + L.dd;
  float* dqT_s = smem + L.dqT;
  float* dUcT_s = smem + L.dUcT;
  float* dhpart_s = smem + L.dhpart;
  float* dcV_s = smem + L.dcV;
  float* xcV_s = smem + L.xcV;
  float* xbe_s = smem + L.xbe;
  float* xqp_s = smem + L.xqp;
  float* xal_s = smem + L.xal;
  float* xdh_s = smem + L.xdh;
  float* drV_s = smem + L.drV;
  float* drT_s = smem + L.drT;
  float* a1_s = smem + L.a1;
  float* vT_s = smem + L.vT;
  float* vV_s = smem + L.vV;
  int* len_s = reinterpret_cast<int*>(smem + L.len);

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t boff = (uint32_t)L.bars * 4u;
  const uint32_t bar0 = smem_base + boff;   // [0] X_e  [1] X_a  [2] X_b  [3] X_c  [4] X_d
  uint32_t rb[kC];
#pragma unroll
  for (int d = 0; d < kC; ++d) rb[d] = mapa_u32(smem_base, (uint32_t)d);
  const uint32_t rb_u = mapa_u32(smem_base, (uint32_t)(lane & 3)), rb_4 = rb[4];

  // The saved activations and P = W K^T must come from the v3 forward sweep of the same workspace: anything else
  // (a forward call that fell back to another kernel, a foreign workspace) stops the kernel with an error
  // instead of producing gradients from uninitialised memory.
  if (p.fwd_tag != nullptr && tid == 0 && __ldg(p.fwd_tag) != 0x03030303u) __trap();   // cudaMemsetAsync(.., 3, 4)
  // ---- one-time staging ------------------------------------------------------------------------------
  // transposed weight slices as mma A fragments: unit (tile, step) holds rows h = 16*tile + {g, g+8},
  // columns k = 8*step + {2t, 2t+1} of A[h][k] = W[input k of this CTA's slice][output h]
  for (int idx = tid; idx < kTiles * kUnitsPerTile * 32; idx += kThreads) {
    const int ln = idx & 31, unit = idx >> 5;
    const int tile = unit / kUnitsPerTile, u = unit - tile * kUnitsPerTile;
    const int g = ln >> 2, tt = ln & 3;
    auto A = [&](int h, int kk) -> float {
      if (h >= kH) return 0.f;
      if (u < kUc) {   // inputs: the 80 gate pre-activation gradients of the slice
        const int k = 8 * (u < kUhh ? u - kUcV : u - kUhh) + kk;
        const int gk = k / kHS, ik = k - gk * kHS;
        const size_t row = (size_t)gk * kH + S0 + ik;
        return u < kUhh ? __ldg(p.W_ih + row * 3 * kH + 2 * kH + h) : __ldg(p.W_hh + row * kH + h);
      }
      const int grp = (u - kUc) / 3;
      const int k = 8 * ((u - kUc) - 3 * grp) + kk;
      if (k >= kHS) return 0.f;
      if (grp == 0) return COND ? __ldg(p.W_c + (size_t)(S0 + k) * 2 * kH + h) : 0.f;
      if (grp == 1) return __ldg(p.W_qT + (size_t)(S0 + k) * kH + h);
      return __ldg(p.W_qV + (size_t)(S0 + k) * kH + h);
    };
    const int h0 = 16 * tile + g;
    W_s[idx] = make_float4(A(h0, 2 * tt), A(h0 + 8, 2 * tt), A(h0, 2 * tt + 1), A(h0 + 8, 2 * tt + 1));
  }
  {
    constexpr int RB = kH * (4 + (COND ? 1 : 0));
    for (int i = tid; i < kNB * Ti * RBl; i += kThreads) {
      const int col = i % RBl, nj = i / RBl;
      const int j = nj % Ti, n = nj / Ti;
      const int type = col / kHS, ii = col - type * kHS;
      P_s[i] = (n < nb) ? __ldg(p.PT + ((size_t)j * B + b0 + n) * RB + type * kH + S0 + ii) : 0.f;
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
    for (int i = tid; i < kNB * kDaS; i += kThreads) da_s[i] = 0.f;
    for (int i = tid; i < kNB * kVs; i += kThreads) { dqV_s[i] = 0.f; dd_s[i] = 0.f; dqT_s[i] = 0.f; }
    for (int i = tid; i < kNB * kH; i += kThreads) dhpart_s[i] = 0.f;
    for (int i = tid; i < 2 * kNB * kStageRow; i += kThreads) smem[L.stage + i] = 0.f;   // rows of absent examples stay zero
    if (tid < kHS) {
      vT_s[tid] = __ldg(p.vT + S0 + tid);
      vV_s[tid] = __ldg(p.vV + S0 + tid);
    }
    if (tid < kNB) len_s[tid] = (tid < nb) ? max(1, min(p.cmd_len[b0 + tid], Ti)) : 1;
    if (tid == 0) {
      for (int k = 0; k < 5; ++k) mbar_init(bar0 + 8u * k, 1);
      fence_mbar_init();
    }
  }

  // ---- thread roles ------------------------------------------------------------------------------------
  const int fg = lane >> 2, ft = lane & 3, nF = 2 * ft;   // mma fragment coordinates
  const bool cellT = tid < kNB * kHS;                     // (example cn, hidden S0 + chh)
  const int cn = tid / kHS, chh = tid - cn * kHS;
  // 320 attention threads (example an, hidden S0 + ah), keys 2k + mg: warps 0-6 and 13-15.  Their 26 tanh per step
  // are recomputed at the top of a step, right after the critical warps 7-13 have sent the last piece of dh (B12): only
  // warp 13 has both jobs (round 1 / first half of round 2: warps 0-9, three warps double-booked, ~450 cycles per step)
  const bool attT = warp < kTiles || warp >= 13;
  const int aidx = warp < kTiles ? tid : tid - (13 - kTiles) * 32;
  const int an = attT ? (aidx >> 1) / kHS : 0, ah = attT ? (aidx >> 1) % kHS : 0, mg = tid & 1;
  const bool an_ok = attT && an < nb;
  const bool cn_ok = cellT && cn < nb;

  // attention accumulators, thread-private for the whole sweep
  float dKV_acc[kM / 2], dKT_acc[kMaxTiB / 2], dvV_acc = 0.f, dvT_acc = 0.f;
#pragma unroll
  for (int k = 0; k < kM / 2; ++k) dKV_acc[k] = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxTiB / 2; ++k) dKT_acc[k] = 0.f;
  const float my_vV = attT ? __ldg(p.vV + S0 + ah) : 0.f, my_vT = attT ? __ldg(p.vT + S0 + ah) : 0.f;

  // recurrent state of the cell threads
  float dc_carry = 0.f, dh_extra = 0.f;
  // Round 2: the saved activations of a step reach their consumers through shared memory.  In round 1 every cell /
  // attention / softmax thread fetched its own 4-byte words one step ahead (9 + 2 + 3 scattered loads per thread and
  // step) and stored its own gradients (7 scattered stores): taking exactly these loads and stores out of the kernel
  // shortened the sweep from 1.12 to 0.80 ms (experiment) - the address arithmetic and the LSU slots
  // sit on the critical path of phases that are bound by issue and latency.  Now warps 14 and 15, idle for most of
  // a step, move the same bytes with 16-byte coalesced accesses.
  float* stage_s = smem + L.stage;
  const bool ioT = warp >= kIoWarp0;
  const int io = tid - kIoWarp0 * 32;          // 0..63 for the I/O threads: float4 column of the staging row
  const float* io_src = nullptr;               // source of that column in row t*B + b (global), and its row stride
  size_t io_stride = 0;
  if (ioT) {
    const int q = io;
    if (q < 20) { io_src = p.gates + (q / 5) * kH + S0 + 4 * (q % 5); io_stride = H4; }
    else if (q < 25) { io_src = p.Cs + S0 + 4 * (q - 20); io_stride = kH; }
    else if (q < 40) { io_src = p.dU + (1 + (q - 25) / 5) * kH + S0 + 4 * ((q - 25) % 5); io_stride = H4; }
    else if (q < 45) { io_src = COND ? p.Qp + S0 + 4 * (q - 40) : nullptr; io_stride = kH; }
    else if (q < 50) { io_src = p.qV + S0 + 4 * (q - 45); io_stride = kH; }
    else if (q < 55) { io_src = p.qT + S0 + 4 * (q - 50); io_stride = kH; }
    else { io_src = p.beta + 4 * (q - 55); io_stride = kM; }
  }
  // 16-byte asynchronous copies global -> shared (LDGSTS: no registers held while the data is in flight); io_commit
  // waits for them, the block barrier that follows publishes the buffer
  auto io_issue = [&](int t, int buf) {
    if (!ioT || io_src == nullptr) return;
#pragma unroll
    for (int n = 0; n < kNB; ++n) {
      if (n < nb) {
        const uint32_t dst = smem_u32(stage_s + (buf * kNB + n) * kStageRow + 4 * io);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(io_src + ((size_t)t * B + b0 + n) * io_stride) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto io_commit = [&]() {
    if (ioT) asm volatile("cp.async.wait_group 0;" ::: "memory");
  };
  // rows [r_lo, ...) of a [kNB][stride] shared-memory array, `nq` float4 per row, to global rows (row0 + n) * gstride + goff(q)
  auto io_store = [&](const float* src_s, int sstride, int nq, float* dst, size_t row0_, int gstride, bool gate_cols) {
    if (!ioT) return;
    for (int f = io; f < kNB * nq; f += kIoThreads) {
      const int n = f / nq, q = f - n * nq;
      if (n < nb) {
        const int gcol = gate_cols ? (q / 5) * kH + 4 * (q % 5) : 4 * q;
        *reinterpret_cast<float4*>(dst + (row0_ + n) * gstride + S0 + gcol) = lds4(src_s + n * sstride + 4 * q);
      }
    }
  };
  float c_cnew = 0.f, s_al = 0.f, s_aux0 = 0.f, s_aux1 = 0.f;
  if (cn_ok) c_cnew = __ldg(p.Cs + ((size_t)p.T * B + b0 + cn) * kH + S0 + chh);   // c_{T-1}: row group T of Cs
  auto load_soft = [&](int t) {
    if (warp >= nb) return;
    s_al = (lane < Ti) ? __ldg(p.alpha + ((size_t)t * B + b0 + warp) * Ti + lane) : 0.f;
  };
  load_soft(p.T - 1);
  __syncthreads();   // the staging buffers have been zeroed
  io_issue(p.T - 1, 0);
  io_commit();
  if (p.dbeta_aux && warp < nb) {
    s_aux0 = __ldg(p.dbeta_aux + (size_t)(b0 + warp) * kM + lane);
    s_aux1 = (lane < kM - 32) ? __ldg(p.dbeta_aux + (size_t)(b0 + warp) * kM + 32 + lane) : 0.f;
  }
  __syncthreads();
  cluster_barrier();

  const uint32_t bytes_vec = (uint32_t)(kC * kNB * kHS * 4), bytes_be = (uint32_t)(kC * kNB * kM * 4),
                 bytes_al = (uint32_t)(kC * kNB * Ti * 4);
  const float4* w_tile = W_s + (size_t)(warp < kTiles ? warp : (warp < 2 * kTiles ? warp - kTiles : 0)) * kUnitsPerTile * 32 + lane;
  const int tile = warp < kTiles ? warp : warp - kTiles;
  // Which warps own the tensor-core work.  The scheduler favours the HIGHER warp index (measured: the nine HMMA of the
  // critical W_qV^T dq_V product on warps 0-6 took 840 cycles while warps 7-13 ran the deferred W_hh^T da product on the
  // same pipes), so the products the step waits for sit on warps 7-13 and the deferred ones on warps 0-6.
  const bool mmaCrit = warp >= kTiles && warp < 2 * kTiles, mmaDefer = warp < kTiles;

  // send the four values of a tile result to the owners of their hidden rows (reduce-scatter)
  auto scatter_tile = [&](const float (&o)[4], int xoff, uint32_t bar) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = 16 * tile + fg + 8 * (j >> 1), n = nF + (j & 1);
      if (h < kH) {
        const int owner = h / kHS, hh = h - owner * kHS;
        const uint32_t dst = mapa_u32(smem_base, (uint32_t)owner);
        st_async_f32(dst + (uint32_t)(xoff + (rank * kNB + n) * kHS + hh) * 4u, o[j], dst + boff + 8u * bar);
      }
    }
  };

  int it = 0;
  for (int t = p.T - 1; t >= 0; --t, ++it) {
    const size_t row0 = (size_t)t * B + b0;
    const uint32_t par = (uint32_t)(it & 1);
    const float* sg = stage_s + (it & 1) * kNB * kStageRow;   // this step's saved activations
    GSCAN3_STAMP(0);
    // ---- tanh of both attentions for this step: depends only on saved activations, overlaps the X_d wait ----
    float zV[kM / 2], zT[kMaxTiB / 2];
    if (attT) {
      const float a_qv = sg[an * kStageRow + kSgQV + ah], a_qt = sg[an * kStageRow + kSgQT + ah];
#pragma unroll
      for (int k = 0; k < kM / 2; ++k) zV[k] = act_tanh(a_qv + KV_s[(an * kM + 2 * k + mg) * kHS + ah]);
#pragma unroll
      for (int k = 0; k < kMaxTiB / 2; ++k) {
        const int j = 2 * k + mg;
        zT[k] = (j < Ti) ? act_tanh(a_qt + KT_s[(an * Ti + j) * kHS + ah]) : 0.f;
      }
    }
    GSCAN3_STAMP(1);
    if (it > 0) mbar_wait(bar0 + 8u * 4, par ^ 1u);   // dh partials of the previous iteration (X_d)
    if (tid == 0) {
      mbar_arm(bar0 + 8u * 0, bytes_vec);
      mbar_arm(bar0 + 8u * 1, bytes_be);
      mbar_arm(bar0 + 8u * 2, bytes_vec);
      mbar_arm(bar0 + 8u * 3, bytes_al);
      mbar_arm(bar0 + 8u * 4, bytes_vec);
    }
    GSCAN3_STAMP(2);
    // ---- B1: LSTM cell backward ------------------------------------------------------------------------
    if (cellT) {
      const float* sc = sg + cn * kStageRow + chh;
      const float c_gi = sc[0], c_gf = sc[kHS], c_gg = sc[2 * kHS], c_go = sc[3 * kHS], c_cprev = sc[kSgC];
      const float c_dUh = sc[kSgDUh], c_dUcT = sc[kSgDUcT];
      float dh_t = dh_extra + c_dUh;
      if (it > 0) {
#pragma unroll
        for (int r = 0; r < kC; ++r) dh_t += xdh_s[(r * kNB + cn) * kHS + chh];
      }
      const float tc = act_tanh(c_cnew);
      const float d_o = dh_t * tc;
      const float dc_t = fmaf(dh_t * c_go, 1.f - tc * tc, dc_carry);
      const float da0 = dc_t * c_gg * c_gi * (1.f - c_gi);
      const float da1 = dc_t * c_cprev * c_gf * (1.f - c_gf);
      const float da2 = dc_t * c_gi * (1.f - c_gg * c_gg);
      const float da3 = d_o * c_go * (1.f - c_go);
      dc_carry = dc_t * c_gf;
      float* dp = da_s + cn * kDaS + chh;
      dp[0] = da0; dp[kHS] = da1; dp[2 * kHS] = da2; dp[3 * kHS] = da3;
      dUcT_s[cn * kHS + chh] = c_dUcT;
      c_cnew = c_cprev;   // c_{t-1} is the "new" cell state of the step about to come
    }
    __syncthreads();
    GSCAN3_STAMP(3);
    // I/O warps: dgates of this step out (da_s is final and untouched until the next cell backward), the saved
    // activations of the next step (t - 1) in: asynchronous copies into the other staging buffer, awaited two barriers later
    io_store(da_s, kDaS, 20, p.dgates, row0, H4, true);
    if (t > 0) io_issue(t - 1, (it + 1) & 1);   // (everybody left the other buffer at the end of the previous step)
    // ---- B2: partial dc_V (X_e) and the W_hh^T da piece of dh ----------------------------------------------
    // (the W_hh^T da piece of dh is not needed before B12: it runs later, in the shadow of the X_b exchange, so that
    //  the seven tiles of this critical product have the tensor pipe to themselves)
    if (mmaCrit) {
      float o[4];
      mv_units<10>(w_tile + kUcV * 32, da_s + fg * kDaS + 2 * ft, o);
      GSCAN3_STAMP(16);
      scatter_tile(o, L.xcV, 0);
    }
    GSCAN3_STAMP(4);
    // ---- B3: the part of dalpha that does not depend on dd: da . P_gates + dU_cT . K^T --------------------------
    {
      const int total = kNB * Ti * 4;
      for (int base = warp * 32; base < total; base += kThreads) {
        const int item = base + lane, pair = item >> 2, u = item & 3;
        float s = 0.f;
        if (item < total) {
          const int n = pair / Ti;
          const float* pp = P_s + (size_t)pair * RBl + COFF + 20 * u;
          const float* dp = da_s + n * kDaS + 20 * u;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float4 a = lds4(dp + 4 * i), b = lds4(pp + 4 * i);
            s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
          }
          const float* kp = KT_s + (size_t)pair * kHS + 5 * u;
          const float* up = dUcT_s + n * kHS + 5 * u;
#pragma unroll
          for (int i = 0; i < 5; ++i) s = fmaf(up[i], kp[i], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (item < total) {
          if (COND) {
            if (u == 0) a1_s[pair] = s;
          } else {   // nothing else contributes: this is already the partial dalpha (X_c)
            const uint32_t off = (uint32_t)(L.xal + rank * kNB * Ti + pair) * 4u;
            st_async_f32(rb_u + off, s, rb_u + boff + 8u * 3);
            if (u == 0) st_async_f32(rb_4 + off, s, rb_4 + boff + 8u * 3);
          }
        }
      }
    }
    GSCAN3_STAMP(5);
    // ---- B4: dc_V slice assembled, partial dbeta over the slice (X_a) ---------------------------------------------
    mbar_wait(bar0 + 8u * 0, par);
    if (cellT) {
      float v = sg[cn * kStageRow + kSgDUcV + chh];
#pragma unroll
      for (int r = 0; r < kC; ++r) v += xcV_s[(r * kNB + cn) * kHS + chh];
      dcV_s[cn * kHS + chh] = v;
    }
    __syncthreads();
    GSCAN3_STAMP(6);
    if (tid < kNB * kM) {
      // one thread per (example, key) pair (288 = nine full warps): 16-byte loads of the key row and of dc_V, no
      // shuffle reduction, and four consecutive pair sums leave as ONE 16-byte st.async per destination (lane k of a
      // group of four -> CTA k, lane 0 also -> CTA 4): 360 instead of 1440 transactions on every receiver's mbarrier
      // (tools/ubench_exchange.cu: 1207 -> 899 cycles for this exchange)
      const int pair = tid, n = pair / kM;
      const float4* kp = reinterpret_cast<const float4*>(KV_s + pair * kHS);
      const float4* dp = reinterpret_cast<const float4*>(dcV_s + n * kHS);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int i = 0; i < kHS / 4; ++i) {
        const float4 k = kp[i], d = dp[i];
        s0 = fmaf(d.x, k.x, s0); s1 = fmaf(d.y, k.y, s1); s0 = fmaf(d.z, k.z, s0); s1 = fmaf(d.w, k.w, s1);
      }
      const float s = s0 + s1;
      const int g0 = lane & ~3, k = lane & 3;
      float4 v;
      v.x = __shfl_sync(0xffffffffu, s, g0);
      v.y = __shfl_sync(0xffffffffu, s, g0 + 1);
      v.z = __shfl_sync(0xffffffffu, s, g0 + 2);
      v.w = __shfl_sync(0xffffffffu, s, g0 + 3);
      const uint32_t off = (uint32_t)(L.xbe + rank * kNB * kM + (pair & ~3)) * 4u;
      st_async_f32x4(rb_u + off, v, rb_u + boff + 8u * 1);
      if (k == 0) st_async_f32x4(rb_4 + off, v, rb_4 + boff + 8u * 1);
    }
    GSCAN3_STAMP(7);
    // ---- B5: softmax backward of the visual attention ---------------------------------------------------------------
    mbar_wait(bar0 + 8u * 1, par);
    if (warp < kNB) {
      const int n = warp;
      const float s_b0 = sg[n * kStageRow + kSgBeta + lane];
      const float s_b1 = (lane < kM - 32) ? sg[n * kStageRow + kSgBeta + 32 + lane] : 0.f;
      float d0 = s_aux0, d1 = s_aux1;
#pragma unroll
      for (int r = 0; r < kC; ++r) d0 += xbe_s[(r * kNB + n) * kM + lane];
      if (lane < kM - 32) {
#pragma unroll
        for (int r = 0; r < kC; ++r) d1 += xbe_s[(r * kNB + n) * kM + 32 + lane];
      }
      const float dot = warp_sum(fmaf(s_b0, d0, s_b1 * d1));
      drV_s[n * kM + lane] = s_b0 * (d0 - dot);
      if (lane < kM - 32) drV_s[n * kM + 32 + lane] = s_b1 * (d1 - dot);
    }
    __syncthreads();
    GSCAN3_STAMP(8);
    // ---- B6: key path of the visual attention: dK^V, dq_V, dv_V ---------------------------------------------------------
    if (attT) {
      float dq = 0.f;
#pragma unroll
      for (int k = 0; k < kM / 2; ++k) {
        const float dr = drV_s[an * kM + 2 * k + mg];
        const float z = zV[k];
        const float g = dr * my_vV * (1.f - z * z);
        dKV_acc[k] += g;
        dq += g;
        dvV_acc = fmaf(dr, z, dvV_acc);
      }
      dq += __shfl_xor_sync(0xffffffffu, dq, 1);
      if (mg == 0) dqV_s[an * kVs + ah] = dq;
    }
    __syncthreads();
    GSCAN3_STAMP(9);
    io_store(dqV_s, kVs, 5, p.dqV, row0, kH, false);
    io_commit();   // the copies issued two barriers ago have landed; published by the barriers that follow
    // ---- B7: partial dq' = W_qV^T dq_V (X_b) ------------------------------------------------------------------------------
    if (mmaCrit) {
      float o[4];
      mv_units<3>(w_tile + kUqV * 32, dqV_s + fg * kVs + 2 * ft, o);
      GSCAN3_STAMP(17);
      scatter_tile(o, L.xqp, 2);
    } else if (mmaDefer) {
      // the W_hh^T da piece of dh (da_s is untouched until the next cell backward), while X_b is in flight
      float o[4];
      mv_units<10>(w_tile + kUhh * 32, da_s + fg * kDaS + 2 * ft, o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int h = 16 * tile + fg + 8 * (j >> 1);
        if (h < kH) dhpart_s[(nF + (j & 1)) * kH + h] = o[j];
      }
    }
    GSCAN3_STAMP(10);
    mbar_wait(bar0 + 8u * 2, par);
    if (cellT) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < kC; ++r) v += xqp_s[(r * kNB + cn) * kHS + chh];
      if (COND) {
        const float c_qp = sg[cn * kStageRow + kSgQp + chh];
        dd_s[cn * kVs + chh] = v * (1.f - c_qp * c_qp);
      } else {
        dh_extra = v;   // q' = h_{t-1}: joins dh at the next cell backward
      }
    }
    GSCAN3_STAMP(11);
    if (COND) {
      __syncthreads();
      io_store(dd_s, kVs, 5, p.dd, row0, kH, false);
      // ---- B9: the W_c[:, :H]^T dd piece of dh, and the partial dalpha completed with dd . P_cond (X_c) -------------
      if (mmaDefer) {
        float o[4];
        mv_units<3>(w_tile + kUc * 32, dd_s + fg * kVs + 2 * ft, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int h = 16 * tile + fg + 8 * (j >> 1);
          if (h < kH) dhpart_s[(nF + (j & 1)) * kH + h] += o[j];
        }
      } else {
        // warps 7-15: 9 warps, 4 lanes per (example, position) pair
        const int w9 = warp - kTiles;
        const int total = kNB * Ti * 4;
        for (int base = w9 * 32; base < total; base += 9 * 32) {
          const int item = base + lane, pair = item >> 2, u = item & 3;
          float s = 0.f;
          if (item < total) {
            const int n = pair / Ti;
            const float* pp = P_s + (size_t)pair * RBl + 5 * u;
            const float* dp = dd_s + n * kVs + 5 * u;
#pragma unroll
            for (int i = 0; i < 5; ++i) s = fmaf(dp[i], pp[i], s);
            if (u == 0) s += a1_s[pair];
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          if (item < total) {
            const uint32_t off = (uint32_t)(L.xal + rank * kNB * Ti + pair) * 4u;
            st_async_f32(rb_u + off, s, rb_u + boff + 8u * 3);
            if (u == 0) st_async_f32(rb_4 + off, s, rb_4 + boff + 8u * 3);
          }
        }
      }
    }
    GSCAN3_STAMP(12);
    // ---- B10: softmax backward of the textual attention ---------------------------------------------------------------------
    mbar_wait(bar0 + 8u * 3, par);
    if (warp < kNB) {
      const int n = warp;
      float d = 0.f;
      if (lane < Ti) {
#pragma unroll
        for (int r = 0; r < kC; ++r) d += xal_s[(r * kNB + n) * Ti + lane];
      }
      const float dot = warp_sum(s_al * d);   // alpha is zero at masked positions and beyond Ti
      if (lane < Ti) drT_s[n * Ti + lane] = s_al * (d - dot);
      if (t > 0) load_soft(t - 1);
    }
    __syncthreads();
    GSCAN3_STAMP(13);
    // ---- B11: key path of the textual attention: dK^T, dq_T, dv_T ---------------------------------------------------------------
    if (attT) {
      float dq = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxTiB / 2; ++k) {
        const int j = 2 * k + mg;
        if (j < Ti) {
          const float dr = drT_s[an * Ti + j];
          const float z = zT[k];
          const float g = dr * my_vT * (1.f - z * z);
          dKT_acc[k] += g;
          dq += g;
          dvT_acc = fmaf(dr, z, dvT_acc);
        }
      }
      dq += __shfl_xor_sync(0xffffffffu, dq, 1);
      if (mg == 0) dqT_s[an * kVs + ah] = dq;
    }
    __syncthreads();
    // dq_T is the last global store of the step, and every store of the step was issued by an I/O thread: "rows
    // t >= t_signal are complete" is published right here by the I/O warps alone - fence by the writers, a barrier among
    // the 64 I/O threads (named barrier 1), one add per CTA.  (At the first block barrier of the next step, as in round 1,
    // the fence sat in the instruction stream of the compute warps and pinned their memory operations, taken or not.)
    io_store(dqT_s, kVs, 5, p.dqT, row0, kH, false);
    if (p.progress != nullptr && ioT) {
      int sig = -1;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < p.n_signals && t == p.t_signal[k]) sig = k;
      if (sig >= 0) {
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(kIoThreads) : "memory");
        if (io == 0) atomicAdd(p.progress + sig, 1u);
      }
    }
    GSCAN3_STAMP(14);
    // ---- B12: last piece of dh, W_qT^T dq_T, added to the earlier pieces and reduce-scattered (X_d) ------------------------------
    if (mmaCrit) {
      float o[4];
      mv_units<3>(w_tile + kUqT * 32, dqT_s + fg * kVs + 2 * ft, o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int h = 16 * tile + fg + 8 * (j >> 1);
        if (h < kH) o[j] += dhpart_s[(nF + (j & 1)) * kH + h];
      }
      GSCAN3_STAMP(18);
      scatter_tile(o, L.xdh, 4);
    }
    GSCAN3_STAMP(15);
  }

  // ---- epilogue -------------------------------------------------------------------------------------------
  mbar_wait(bar0 + 8u * 4, (uint32_t)((it - 1) & 1));
  if (cn_ok) {
    float dh = dh_extra;
#pragma unroll
    for (int r = 0; r < kC; ++r) dh += xdh_s[(r * kNB + cn) * kHS + chh];
    p.dh0[(size_t)(b0 + cn) * kH + S0 + chh] = dh + dc_carry;   // h_{-1} = c_{-1} = the same tensor
  }
  if (an_ok) {
#pragma unroll
    for (int k = 0; k < kM / 2; ++k) p.dKV[((size_t)(b0 + an) * kM + 2 * k + mg) * kH + S0 + ah] = dKV_acc[k];
#pragma unroll
    for (int k = 0; k < kMaxTiB / 2; ++k) {
      const int j = 2 * k + mg;
      if (j < Ti) p.dKT[((size_t)j * B + b0 + an) * kH + S0 + ah] = dKT_acc[k];
    }
    atomicAdd(p.dvV + S0 + ah, dvV_acc);
    atomicAdd(p.dvT + S0 + ah, dvT_acc);
  }
  cluster_barrier();
}

// ---- after the sweep: value path of both attentions ------------------------------------------------------
//   dK^V[b, m, :] += sum_t beta[t, b, m]  * dc_V[t, b, :]
//   dK^T[j, b, :] += sum_t alpha[t, b, j] * dc_T[t, b, :]
// with dc = [dc_T | dc_V] rows of width 2H at stride ldc.  grid = (B, 2): y = 0 visual, y = 1 textual.
__global__ void __launch_bounds__(256) attn_value_bwd_kernel(const float* __restrict__ dc, long ldc, const float* __restrict__ beta,
                                                             const float* __restrict__ alpha, int B, int T, int Ti, int M, int H,
                                                             float* __restrict__ dKV, float* __restrict__ dKT) {
  // per example a [N x T] . [T x H] product with both operands staged in shared memory:
  // w_s [T][NP] attention weights (NP = N rounded up to 4), dc_s [T][H] context gradients
  extern __shared__ __align__(16) float av_s[];
  const int b = blockIdx.x, vis = blockIdx.y == 0;
  const int N = vis ? M : Ti, NP = (N + 3) & ~3;
  const float* w = vis ? beta : alpha;
  float* w_s = av_s;
  float* dc_s = av_s + (size_t)T * NP;
  for (int i = threadIdx.x; i < T * NP; i += blockDim.x) {
    const int t = i / NP, k = i - t * NP;
    w_s[i] = k < N ? __ldg(w + ((size_t)t * B + b) * N + k) : 0.f;
  }
  const float* dcb = dc + (vis ? H : 0);
  for (int i = threadIdx.x; i < T * H; i += blockDim.x) {
    const int t = i / H, h = i - t * H;
    dc_s[i] = __ldg(dcb + ((size_t)t * B + b) * ldc + h);
  }
  __syncthreads();
  // thread = (hidden h, key quad group): 4 keys x 1 hidden per accumulator set, keys strided by the group count
  const int HP = (H + 31) & ~31;                 // lanes per key group, warp aligned
  const int groups = blockDim.x / HP;            // 2 for H = 100
  const int grp = threadIdx.x / HP, h = threadIdx.x - grp * HP;
  if (grp >= groups || h >= H) return;
  for (int k0 = 4 * grp; k0 < NP; k0 += 4 * groups) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float4 wv = *reinterpret_cast<const float4*>(w_s + t * NP + k0);
      const float d = dc_s[t * H + h];
      acc.x = fmaf(wv.x, d, acc.x); acc.y = fmaf(wv.y, d, acc.y); acc.z = fmaf(wv.z, d, acc.z); acc.w = fmaf(wv.w, d, acc.w);
    }
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u;
      if (k < N) {
        if (vis) dKV[((size_t)b * M + k) * H + h] += a[u];
        else dKT[((size_t)k * B + b) * H + h] += a[u];
      }
    }
  }
}

// ---- after the sweep: value path of both attentions, reordered so that nothing of size [Tt*B] is multiplied ----
//   dK^V[b, m, :] += sum_t beta[t, b, m]  * dc_V[t, b, :],   dK^T[j, b, :] += sum_t alpha[t, b, j] * dc_T[t, b, :]
// where dc_V[t, b, :] = X[t, b, :] . Wst_V and dc_T[t, b, :] = X[t, b, :] . Wst_T are linear in the row
// X = [dgates (4H) | dpre (H) | dd (H)] (LSTM input block, output head, conditional query).  Summing over t FIRST,
//   Z_V[b, m, :] = sum_t beta[t, b, m] X[t, b, :]        (per example a [36 x Tt] . [Tt x 5H] product)
//   Z_T[j, b, :] = sum_t alpha[t, b, j] X[t, b, :]       ([Ti x Tt] . [Tt x 6H])
// turns the two [Tt*B x 4H] x [4H x 2H] products + attn_value_bwd_kernel of the first version (166 us on the
// critical path after the sweep) into this one memory-bound pass over X plus two small products
// dK^V += Z_V . Wst_V ([B*36 x 5H] . [5H x H]) and dK^T += Z_T . Wst_T ([Ti*B x 6H] . [6H x H]).
struct ValueZP {
  const float *dgates, *dpre, *dd;   // [T*B][4H], [T*B][H], [T*B][H] (dd null without conditional attention)
  const float *alpha, *beta;         // [T][B][Ti], [T][B][36]
  int B, T, Ti, H, NC;               // NC = 5H or 6H columns of X
  float *ZV, *ZT;                    // [B*36][ldv], [Ti*B][ldt]
  int ldv, ldt;
  int t_begin, t_end;                // steps summed by this launch (0, T for all of them)
  int accumulate;                    // add to ZV / ZT (partial sums of an earlier launch over other steps)
};

// grid = (B, ceil(NC/4 / 64)), 256 threads.  Thread = (column quad, weight group wg): lane = 8 column quads x 4
// groups, group wg owns the weight quads {wg, wg+4, wg+8, wg+12} of the 13 (so <= 16 weights x 4 columns = 32 packed
// accumulators): per step one 16-byte quad of X and <= 4 shared-memory weight quads feed 32 FFMA2.
// X is a strided gather (one 1 KB row segment per step, rows B*ld floats apart), so the kernel is bound by how many
// loads it keeps in flight: each group of 4 lanes streams its column quad through a private shared-memory ring with
// cp.async, kZRing rounds (of 4 steps) ahead - no registers held by loads in flight and no block barrier in the loop.
// (First versions: register loads 4 steps ahead, 117-128 us; the 58 MB of X cost ~10 us at HBM speed.)
constexpr int kZRing = 8;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// QW = weight quads per lane group: 3 (48 weight slots: 36 beta + Ti <= 12 alpha) or 4 (64 slots, Ti <= 16); the
// unused slots hold zeros, so the inner loop is branch-free and the same for every lane.  Steps are padded to a
// multiple of 4 with zero weights AND zero X (a stale ring slot could hold a NaN).
inline int value_z_qw(int Ti) { return Ti <= 12 ? 3 : 4; }
inline size_t value_z_smem_bytes(int T, int Ti) {   // T = steps of one launch
  const int Tp = (T + 3) & ~3;
  return sizeof(float) * ((size_t)Tp * 16 * value_z_qw(Ti) + (size_t)kZRing * 4 * 64 * 4);
}

template <int QW>
__global__ void __launch_bounds__(256, 2) attn_value_z_kernel(ValueZP p) {
  constexpr int ZS = 16 * QW;                      // weight slots per step
  extern __shared__ __align__(16) float zw_s[];   // [Tp][ZS] weights, then the ring [kZRing][4 steps][64 quads] float4
  const int b = blockIdx.x, B = p.B, T = p.t_end - p.t_begin, Ti = p.Ti, H = p.H;   // local step tl = t - t_begin
  const int Tp = (T + 3) & ~3, nr = Tp / 4;
  float4* ring = reinterpret_cast<float4*>(zw_s + (size_t)Tp * ZS);
  const int wg = threadIdx.x & 3, ql = threadIdx.x >> 2;
  const int c = 4 * (blockIdx.y * 64 + ql);   // first of the 4 columns
  const bool active = c < p.NC;
  const float* src = p.dgates;
  size_t ld = 4 * (size_t)H;
  if (active) {
    if (c < 4 * H) { src = p.dgates + c; }
    else if (c < 5 * H) { src = p.dpre + (c - 4 * H); ld = H; }
    else { src = p.dd + (c - 5 * H); ld = H; }
    src += ((size_t)p.t_begin * B + b) * ld;
  }
  const size_t step = (size_t)B * ld;
  // lane wg of a quad's 4 lanes fetches step 4r + wg of round r; all 4 lanes consume all 4 steps
  auto issue = [&](int r) {
    const int t = 4 * r + wg;
    if (r < nr) {
      float4* dst = &ring[((r % kZRing) * 4 + wg) * 64 + ql];
      if (active && t < T) cp_async16(dst, src + (size_t)t * step);
      else *dst = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };
  // the attention weights of this example go in first, as one more (oldest) cp.async group: a dependent
  // load -> store loop here cost 30 us of exposed latency
  for (int i = threadIdx.x; i < Tp * ZS; i += blockDim.x) {
    const int t = i / ZS, k = i - t * ZS;
    if (t < T && k < kM) cp_async4(zw_s + i, p.beta + ((size_t)(p.t_begin + t) * B + b) * kM + k);
    else if (t < T && k - kM < Ti) cp_async4(zw_s + i, p.alpha + ((size_t)(p.t_begin + t) * B + b) * Ti + (k - kM));
    else zw_s[i] = 0.f;
  }
  cp_async_commit();
#pragma unroll
  for (int r = 0; r < kZRing; ++r) issue(r);
  cp_async_wait<kZRing>();
  __syncthreads();
  float2 acc[QW][4][2];                            // [owned quad][weight in quad][column pair]
#pragma unroll
  for (int q = 0; q < QW; ++q)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[q][k][0] = acc[q][k][1] = make_float2(0.f, 0.f);
  for (int r = 0; r < nr; ++r) {
    cp_async_wait<kZRing - 1>();
    __syncwarp();
    const float4* xr = ring + ((r % kZRing) * 4) * 64 + ql;
    const float4* wr = reinterpret_cast<const float4*>(zw_s + (size_t)(4 * r) * ZS) + wg;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 x = xr[u * 64];
      const float2 x01 = make_float2(x.x, x.y), x23 = make_float2(x.z, x.w);
#pragma unroll
      for (int q = 0; q < QW; ++q) {
        const float4 wv = wr[u * (ZS / 4) + 4 * q];
        const float ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 ww = make_float2(ws[k], ws[k]);
          fma2(acc[q][k][0], ww, x01);
          fma2(acc[q][k][1], ww, x23);
        }
      }
    }
    __syncwarp();
    issue(r + kZRing);
  }
  if (!active) return;
#pragma unroll
  for (int q = 0; q < QW; ++q) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int widx = 4 * (wg + 4 * q) + k;     // 0..35 beta, 36.. alpha
      float4 o = make_float4(acc[q][k][0].x, acc[q][k][0].y, acc[q][k][1].x, acc[q][k][1].y);
      float4* dst = nullptr;
      if (widx < kM) {
        if (c < p.ldv) dst = reinterpret_cast<float4*>(p.ZV + ((size_t)b * kM + widx) * p.ldv + c);
      } else if (widx - kM < Ti) {
        dst = reinterpret_cast<float4*>(p.ZT + ((size_t)(widx - kM) * B + b) * p.ldt + c);
      }
      if (dst) {
        if (p.accumulate) { const float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
        *dst = o;
      }
    }
  }
}

// ---- the same sums on the tensor cores ------------------------------------------------------------------------------
// Per example the two sums are ONE small product  Z_b [MT*16 x 6H] = W_b [MT*16 x Tt] . X_b [Tt x 6H]  with the rows of
// W_b = (36 beta columns | Ti alpha columns | zeros) transposed: mma.sync m16n8k8 tf32 in split precision (W_lo X_hi +
// W_hi X_lo + W_hi X_hi, truncation split: the raw word is the hi operand).  The FFMA2 kernel above issues 29 M warp
// instructions (14 M of them FFMA2) and is bound by issue (64 us); this one issues 2.2 M MMAs + ~3 M others.
//   grid (B, 3): CTA = (example, third of the 600 columns = 25 n-tiles);  160 threads = 5 warps x 5 n-tiles x MT m-tiles.
//   W_b sits in shared memory ([MT*16][Tp + 4]: fragment loads conflict-free; split in registers), X streams through a
//   per-warp cp.async rings of kZmStages stages of 8 steps x 40 columns (row pitch 40 = 8 mod 32: B fragments conflict-free).
constexpr int kZmStages = 6, kZmCols = 200, kZmThreads = 160;
inline size_t value_zm_smem_bytes(int T, int MT) {
  const int Tp = (T + 7) & ~7;
  return sizeof(float) * ((size_t)MT * 16 * (Tp + 4) + (size_t)kZmStages * 8 * kZmCols);
}
template <int MT>
__global__ void __launch_bounds__(kZmThreads, 3) attn_value_zm_kernel(ValueZP p) {
  extern __shared__ __align__(16) float zm_s[];
  const int b = blockIdx.x, third = blockIdx.y, B = p.B, T = p.t_end - p.t_begin, Ti = p.Ti;
  const int Tp = (T + 7) & ~7, nk = Tp / 8, ldw = Tp + 4;
  float* w_s = zm_s;                                   // [MT*16][ldw]
  float* ring = zm_s + (size_t)MT * 16 * ldw;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  // Every warp streams ITS 40 columns through a private ring (no block barrier in the loop: the warps drift apart and
  // hide each other's waits).  Stage = steps [8 ks, 8 ks + 8) x 40 columns = 80 16-byte chunks, <= 3 per lane, the same
  // (row in stage, column quad) every stage: source pointers and pitches are set up once (zero beyond T and for an
  // absent dd: the ring slot is cleared instead)
  constexpr int kWCols = 40, kChunks = 8 * (kWCols / 4), kPer = (kChunks + 31) / 32;
  const int warp_ = tid >> 5, lane_ = tid & 31;
  float* wring = ring + (size_t)warp_ * kZmStages * 8 * kWCols;
  // Steps beyond T are NOT cleared: their weights are zero and the row read instead (the chunk's last valid step) is
  // finite, so the loop body is one LDGSTS and one pointer add per chunk.  Columns without a source (dd of a model
  // without conditional attention) are cleared once in every stage and never loaded.
  const float* csrc[kPer];
  size_t cstep[kPer];      // floats between consecutive stages (8 steps)
  int coff[kPer], clast[kPer];   // clast: last stage whose row of this chunk is < T
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = lane_ + j * 32;
    const int r = i / (kWCols / 4), c4 = i - r * (kWCols / 4);
    const int col = third * kZmCols + warp_ * kWCols + 4 * c4;     // of the 600
    coff[j] = r * kWCols + 4 * c4;
    // rows of this chunk: r, r + 8, ...; if even the first is beyond T, read step 0 of the launch (its weight is zero)
    const int r0 = r < T ? r : 0;
    clast[j] = r < T ? (T - 1 - r) / 8 : -1;
    const size_t row = (size_t)(p.t_begin + r0) * B + b;
    csrc[j] = nullptr;
    cstep[j] = 0;
    if (i < kChunks) {
      if (col < 4 * kH) { csrc[j] = p.dgates + row * (4 * kH) + col; cstep[j] = (size_t)8 * B * 4 * kH; }
      else if (col < 5 * kH) { csrc[j] = p.dpre + row * kH + (col - 4 * kH); cstep[j] = (size_t)8 * B * kH; }
      else if (p.NC > 5 * kH) { csrc[j] = p.dd + row * kH + (col - 5 * kH); cstep[j] = (size_t)8 * B * kH; }
      if (!csrc[j])
        for (int sgi = 0; sgi < kZmStages; ++sgi)
          *reinterpret_cast<float4*>(wring + (size_t)sgi * 8 * kWCols + coff[j]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  auto issue = [&](int ks) {
    if (ks < nk) {
      float* st = wring + (size_t)(ks % kZmStages) * 8 * kWCols;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        if (csrc[j]) {
          cp_async16(st + coff[j], csrc[j]);
          if (ks < clast[j]) csrc[j] += cstep[j];
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kZmStages - 1; ++s) issue(s);
  // attention weights of this example, transposed (rows >= 36 + Ti and steps >= T are zero): 4-byte cp.async, one more
  // group in flight beside the first stages of X - a load -> store loop here is ~40 dependent DRAM round trips.
  // (t, m) advance by kZmThreads elements per iteration without a division.
  {
    constexpr int W = MT * 16, dT = kZmThreads / W, dM = kZmThreads - dT * W;
    int t = tid / W, m = tid - t * W;
    for (; t < Tp; ) {
      float* dst = w_s + m * ldw + t;
      if (t < T && m < kM) cp_async4(dst, p.beta + ((size_t)(p.t_begin + t) * B + b) * kM + m);
      else if (t < T && m - kM < Ti) cp_async4(dst, p.alpha + ((size_t)(p.t_begin + t) * B + b) * Ti + (m - kM));
      else *dst = 0.f;
      t += dT; m += dM;
      if (m >= W) { m -= W; ++t; }
    }
  }
  cp_async_commit();
  float acc[MT][5][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 5; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
  const int ncol0 = warp * 40;   // this warp's first column inside the CTA's 200
  cp_async_wait<0>();
  __syncthreads();                 // the weights (and the first stages) have landed for everybody
  for (int ks = 0; ks < nk; ++ks) {
    // groups in flight: this warp's stages ks .. ks + kZmStages - 2
    cp_async_wait<kZmStages - 2>();
    __syncwarp();                  // stage ks landed for the whole warp; stage ks - 1 is free (the warp computed it)
    issue(ks + kZmStages - 1);
    const float* st = wring + (size_t)(ks % kZmStages) * 8 * kWCols;
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int o = (16 * mt + g) * ldw + 8 * ks + t4;
      const float wv[4] = {w_s[o], w_s[o + 8 * ldw], w_s[o + 4], w_s[o + 8 * ldw + 4]};
#pragma unroll
      for (int c = 0; c < 4; ++c) {   // the tensor core reads the top 19 bits of the raw word: it IS the hi operand
        ah[mt][c] = __float_as_uint(wv[c]);
        al[mt][c] = __float_as_uint(wv[c] - __uint_as_float(ah[mt][c] & 0xffffe000u));
      }
    }
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
      const float x0 = st[t4 * kWCols + 8 * nt + g], x1 = st[(t4 + 4) * kWCols + 8 * nt + g];
      const uint32_t bh0 = __float_as_uint(x0), bh1 = __float_as_uint(x1);
      const uint32_t bl0 = __float_as_uint(x0 - __uint_as_float(bh0 & 0xffffe000u));
      const uint32_t bl1 = __float_as_uint(x1 - __uint_as_float(bh1 & 0xffffe000u));
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][nt], al[mt][0], al[mt][1], al[mt][2], al[mt][3], bh0, bh1);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bl0, bl1);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bh0, bh1);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 5; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = 16 * mt + g + 8 * h;
        const int c = third * kZmCols + ncol0 + 8 * nt + 2 * t4;
        float2* dst = nullptr;
        if (row < kM) {
          if (c < p.ldv) dst = reinterpret_cast<float2*>(p.ZV + ((size_t)b * kM + row) * p.ldv + c);
        } else if (row - kM < Ti) {
          if (c < p.NC) dst = reinterpret_cast<float2*>(p.ZT + ((size_t)(row - kM) * B + b) * p.ldt + c);
        }
        if (dst) {
          float2 o = make_float2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
          if (p.accumulate) { const float2 old = *dst; o.x += old.x; o.y += old.y; }
          *dst = o;
        }
      }
}

// Wst_V [5H][H] = [W_ih[:, 2H:3H] ; W_o2h[:, 3H:4H]],  Wst_T [6H][H] = [W_ih[:, H:2H] ; W_o2h[:, 2H:3H] ; W_c[:, H:2H]]
__global__ void value_weight_stack_kernel(const float* __restrict__ W_ih, const float* __restrict__ W_o2h,
                                          const float* __restrict__ W_c, int H, float* __restrict__ WstV,
                                          float* __restrict__ WstT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nV = 5 * H * H, nT = (W_c ? 6 : 5) * H * H;
  if (i < nV) {
    const int r = i / H, j = i - r * H;
    WstV[i] = r < 4 * H ? __ldg(W_ih + (size_t)r * 3 * H + 2 * H + j) : __ldg(W_o2h + (size_t)(r - 4 * H) * 4 * H + 3 * H + j);
  } else if (i < nV + nT) {
    const int k = i - nV, r = k / H, j = k - r * H;
    float v;
    if (r < 4 * H) v = __ldg(W_ih + (size_t)r * 3 * H + H + j);
    else if (r < 5 * H) v = __ldg(W_o2h + (size_t)(r - 4 * H) * 4 * H + 2 * H + j);
    else v = __ldg(W_c + (size_t)(r - 5 * H) * 2 * H + H + j);
    WstT[k] = v;
  }
}

}  // namespace v3
}  // namespace gscan
