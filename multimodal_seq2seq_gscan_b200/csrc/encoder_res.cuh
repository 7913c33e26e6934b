// Command encoder sweeps with the recurrent weights RESIDENT in shared memory
// (reference seq2seq_model.py:47-89: packed bidirectional LSTM; backward = autograd of it).
//
// encoder_fwd_kernel / encoder_bwd_kernel (recurrent.cuh) stream the 160 KB W_hh of their direction from L2 in
// every step and for every CTA (200 CTAs x 10 steps x 160 KB = 320 MB of L2 reads, 7.7 us per step).  Here one
// CTA owns NB = 4 examples of one direction, copies W_hh once into shared memory (160 KB of the 227 KB) and runs
// the whole sequence from there; whatever a step needs from HBM (input-gate pre-activations, saved activations)
// is fetched one step ahead into registers, so no global-memory latency sits on the recurrent chain.
// Same outputs, element for element, as the streaming kernels (they remain the fallback for shapes whose
// W_hh does not fit or whose K-slices are not multiples of 4).
#pragma once
#include "recurrent.cuh"

namespace gscan {

// part[(s*NB + n)*R + r] = sum_{k in slice s} Wt_s[k*ldw + r] * x_s[n*ldx + k], Wt_s and x_s in shared memory.
// Slices are multiples of 4 wide (checked on the host) so that x is read as broadcast 128-bit loads.
template <int NB>
__device__ __forceinline__ void matvec_partial_s(const float* __restrict__ Wt_s, int ldw, int R, int K,
                                                 const float* __restrict__ x_s, int ldx, float* part, int KS) {
  const int RQ = R >> 2;
  for (int item = threadIdx.x; item < RQ * KS; item += blockDim.x) {
    const int s = item / RQ, q = item - s * RQ;
    const int k0 = (K * s) / KS, k1 = (K * (s + 1)) / KS;
    float2 acc[NB][2];
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[n][0] = acc[n][1] = make_float2(0.f, 0.f);
    const float* wp = Wt_s + (long)k0 * ldw + 4 * q;
    for (int k = k0; k < k1; k += 4, wp += 4 * ldw) {
      float4 w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(wp + j * ldw);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float4 x = *reinterpret_cast<const float4*>(x_s + n * ldx + k);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 xx = make_float2(xs[j], xs[j]);
          fma2(acc[n][0], make_float2(w[j].x, w[j].y), xx);
          fma2(acc[n][1], make_float2(w[j].z, w[j].w), xx);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NB; ++n)
      *reinterpret_cast<float4*>(&part[(s * NB + n) * R + 4 * q]) =
          make_float4(acc[n][0].x, acc[n][0].y, acc[n][1].x, acc[n][1].y);
  }
}

// shape test shared by host and kernels: slices of both directions of the product must be multiples of 4
__host__ __device__ inline bool enc_res_slices_ok(int H, int nthreads) {
  const int ksf = matvec_splits(4 * H, H, nthreads), ksb = matvec_splits(H, 4 * H, nthreads);
  return H % 4 == 0 && H % ksf == 0 && (H / ksf) % 4 == 0 && (4 * H) % ksb == 0 && ((4 * H) / ksb) % 4 == 0;
}

template <int NB>
size_t enc_res_smem_floats(int H, int nthreads, bool bwd) {
  const size_t w = (size_t)H * 4 * H;
  if (bwd) return w + pad4(NB * 4 * H) + 16 + (size_t)matvec_splits(H, 4 * H, nthreads) * NB * H;
  return w + pad4(NB * H) + 16 + (size_t)matvec_splits(4 * H, H, nthreads) * NB * 4 * H;
}

// grid = (ceil(B/NB), 2 directions); requires NB*H <= blockDim.x: thread i owns (example i/H, hidden unit i%H)
template <int NB>
__global__ void __launch_bounds__(kRecThreads, 1) encoder_fwd_res_kernel(EncP p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, H4 = 4 * H, Ti = p.Ti, B = p.B;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int b0 = blockIdx.x * NB, d = blockIdx.y;
  const int nb = min(NB, B - b0);
  Bump bump{smem};
  float* W_s = bump.take(H * H4);
  float* h_s = bump.take(NB * H);
  int* len_s = reinterpret_cast<int*>(bump.take(16));
  const int KS = matvec_splits(H4, H, NT);
  float* part = bump.p;
  {
    const float4* src = reinterpret_cast<const float4*>(p.Whh_t[d]);
    float4* dst = reinterpret_cast<float4*>(W_s);
    for (int i = tid; i < H * H; i += NT) dst[i] = __ldg(src + i);   // H*4H/4 quads
  }
  for (int i = tid; i < NB * H; i += NT) h_s[i] = 0.f;
  if (tid < NB) len_s[tid] = (tid < nb) ? max(1, min(p.len[b0 + tid], Ti)) : 0;
  const int n = tid / H, h = tid - n * H;
  const bool owner = tid < NB * H && n < nb;
  float c_reg = 0.f, h_reg = 0.f;
  float x[4] = {0.f, 0.f, 0.f, 0.f};
  auto fetch = [&](int step) {
    const int t = d == 0 ? step : Ti - 1 - step;
    const float* xp = p.xg[d] + ((long)(b0 + n) * Ti + t) * H4 + h;
#pragma unroll
    for (int g = 0; g < 4; ++g) x[g] = __ldg(xp + g * H);
  };
  if (owner) fetch(0);
  __syncthreads();
  const int my_len = owner ? len_s[n] : 0;
  for (int step = 0; step < Ti; ++step) {
    const int t = d == 0 ? step : Ti - 1 - step;
    matvec_partial_s<NB>(W_s, H4, H4, H, h_s, H, part, KS);
    __syncthreads();
    if (owner) {
      const float a0 = x[0] + part_sum<NB>(part, KS, H4, n, h);
      const float a1 = x[1] + part_sum<NB>(part, KS, H4, n, H + h);
      const float a2 = x[2] + part_sum<NB>(part, KS, H4, n, 2 * H + h);
      const float a3 = x[3] + part_sum<NB>(part, KS, H4, n, 3 * H + h);
      if (step + 1 < Ti) fetch(step + 1);
      const float ig = act_sigmoid(a0), fg = act_sigmoid(a1), gg = act_tanh(a2), og = act_sigmoid(a3);
      const float cn = fmaf(fg, c_reg, ig * gg);
      const float hn = og * act_tanh(cn);
      const long row = (long)t * B + b0 + n;
      float* gp = p.enc_g[d] + row * H4 + h;
      gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
      if (t < my_len) {
        h_reg = hn;
        c_reg = cn;
        h_s[tid] = hn;
        atomicAdd(p.enc_out + row * H + h, hn);   // two commutative adds onto zero: deterministic
      }
      p.enc_h[d][row * H + h] = h_reg;
      p.enc_c[d][row * H + h] = c_reg;
    }
    __syncthreads();
  }
  if (owner) atomicAdd(p.h_enc + (long)(b0 + n) * H + h, h_reg);
}

template <int NB>
__global__ void __launch_bounds__(kRecThreads, 1) encoder_bwd_res_kernel(EncP p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, H4 = 4 * H, Ti = p.Ti, B = p.B;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int b0 = blockIdx.x * NB, d = blockIdx.y;
  const int nb = min(NB, B - b0);
  Bump bump{smem};
  float* W_s = bump.take(H4 * H);    // original [4H][H]: row k = gate row, the reduction index of W_hh^T da
  float* da_s = bump.take(NB * H4);
  int* len_s = reinterpret_cast<int*>(bump.take(16));
  const int KS = matvec_splits(H, H4, NT);
  float* part = bump.p;
  {
    const float4* src = reinterpret_cast<const float4*>(p.W_hh[d]);
    float4* dst = reinterpret_cast<float4*>(W_s);
    for (int i = tid; i < H * H; i += NT) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < NB * H4; i += NT) da_s[i] = 0.f;   // rows of absent examples stay zero
  if (tid < NB) len_s[tid] = (tid < nb) ? max(1, min(p.len[b0 + tid], Ti)) : 0;
  const int n = tid / H, h = tid - n * H;
  const bool owner = tid < NB * H && n < nb;
  float dh = owner ? __ldg(p.dh_enc + (long)(b0 + n) * H + h) : 0.f;
  float dc = 0.f;
  // saved activations of the step about to be processed, fetched one step ahead
  float g4[4] = {0.f, 0.f, 0.f, 0.f}, h_prev = 0.f, c_prev = 0.f, dout = 0.f;
  auto fetch = [&](int step) {
    const int t = d == 0 ? step : Ti - 1 - step;
    const int tp = d == 0 ? t - 1 : t + 1;   // position visited before t in this direction
    const long row = (long)t * B + b0 + n;
    h_prev = 0.f;
    c_prev = 0.f;
    if (step > 0) {
      const long rp = (long)tp * B + b0 + n;
      h_prev = __ldg(p.enc_h[d] + rp * H + h);
      c_prev = __ldg(p.enc_c[d] + rp * H + h);
    }
    const float* gp = p.enc_g[d] + row * H4 + h;
#pragma unroll
    for (int g = 0; g < 4; ++g) g4[g] = __ldg(gp + g * H);
    dout = __ldg(p.denc_out + row * H + h);
  };
  if (owner) fetch(Ti - 1);
  __syncthreads();
  const int my_len = owner ? len_s[n] : 0;
  for (int step = Ti - 1; step >= 0; --step) {
    const int t = d == 0 ? step : Ti - 1 - step;
    const bool valid = t < my_len;
    if (owner) {
      float da0 = 0.f, da1 = 0.f, da2 = 0.f, da3 = 0.f;
      p.hprev[d][((long)(b0 + n) * Ti + t) * H + h] = h_prev;
      if (valid) {
        const float ig = g4[0], fg = g4[1], gg = g4[2], og = g4[3];
        const float c_new = fmaf(fg, c_prev, ig * gg);
        const float tc = act_tanh(c_new);
        const float dh_t = dh + dout;
        const float d_o = dh_t * tc;
        const float dc_t = fmaf(dh_t * og, 1.f - tc * tc, dc);
        da0 = dc_t * gg * ig * (1.f - ig);
        da1 = dc_t * c_prev * fg * (1.f - fg);
        da2 = dc_t * ig * (1.f - gg * gg);
        da3 = d_o * og * (1.f - og);
        dc = dc_t * fg;
      }
      float* dg = p.dga[d] + ((long)(b0 + n) * Ti + t) * H4 + h;
      dg[0] = da0; dg[H] = da1; dg[2 * H] = da2; dg[3 * H] = da3;
      float* dp = da_s + n * H4 + h;
      dp[0] = da0; dp[H] = da1; dp[2 * H] = da2; dp[3 * H] = da3;
      if (step > 0) fetch(step - 1);
    }
    __syncthreads();
    matvec_partial_s<NB>(W_s, H, H, H4, da_s, H4, part, KS);
    __syncthreads();
    if (owner && valid) dh = part_sum<NB>(part, KS, H, n, h);
  }
}

}  // namespace gscan
