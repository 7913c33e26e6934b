// Situation CNN (reference cnn_model.py:22-36): relu(cat[conv1x1, conv5x5, conv k3xk3]) * dropout.
//
// gSCAN situations are {0,1} grids with a handful of occupied cells (about 30 non-zeros out of
// G*G*C = 576), so the kernels are zero-skipping direct convolutions: work is proportional to the
// number of non-zero inputs, and for dense inputs they degrade gracefully to the dense operation
// count.  Results are exact fp32 sums of weight taps (skipping x == 0 terms changes nothing).
//
// Weights are first re-laid "tap major": Wt[conv][dr][dc][ch][f] (f contiguous), where dr / dc
// are the grid ROW / COLUMN offsets.  The reference convolves the transposed grid
// (cnn_model.py:28), so reference tap w[f][ch][i][j] multiplies x[row + j - p][col + i - p]:
// dr <-> j, dc <-> i.
#pragma once
#include "common.cuh"

namespace gscan {

struct CnnShape {
  int B, G, C, F, K3;
  __host__ __device__ int M() const { return G * G; }
  __host__ __device__ int D() const { return 3 * F; }
  __host__ __device__ int ksize(int conv) const { return conv == 0 ? 1 : (conv == 1 ? 5 : K3); }
  __host__ __device__ int woff(int conv) const {  // float offset of conv's block in the tap-major buffer
    int o = 0;
    for (int c = 0; c < conv; ++c) o += ksize(c) * ksize(c) * C * F;
    return o;
  }
  __host__ __device__ int wtotal() const { return woff(3); }
};

// to_tap_major = 1: Wt <- w (forward prep);  = 0: w <- Wt (gradient un-prep)
__global__ void cnn_relayout_kernel(CnnShape s, float* w1, float* w2, float* w3, float* Wt, int to_tap_major) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= s.wtotal()) return;
  int conv = idx >= s.woff(2) ? 2 : (idx >= s.woff(1) ? 1 : 0);
  int k = s.ksize(conv);
  int r = idx - s.woff(conv);
  int f = r % s.F; r /= s.F;
  int ch = r % s.C; r /= s.C;
  int dc = r % k;
  int dr = r / k;
  float* w = conv == 0 ? w1 : (conv == 1 ? w2 : w3);
  long src = (((long)f * s.C + ch) * k + dc) * k + dr;   // w[f][ch][i=dc][j=dr]
  if (to_tap_major) Wt[idx] = w[src];
  else w[src] = Wt[idx];
}

// Ordered (deterministic) compaction of the non-zeros of v[0..n) (stride `stride`) by warp 0.
// Writes indices / values to smem lists; returns the count through *count_s.
__device__ __forceinline__ void compact_nonzero(const float* __restrict__ v, long stride, int n,
                                                int* idx_s, float* val_s, int* count_s) {
  if (threadIdx.x < 32) {
    int lane = threadIdx.x;
    int count = 0;
    for (int base = 0; base < n; base += 32) {
      int i = base + lane;
      float x = (i < n) ? __ldg(v + (long)i * stride) : 0.f;
      unsigned m = __ballot_sync(0xffffffffu, x != 0.f);
      if (x != 0.f) {
        int pos = count + __popc(m & ((1u << lane) - 1u));
        idx_s[pos] = i;
        val_s[pos] = x;
      }
      count += __popc(m);
    }
    if (lane == 0) *count_s = count;
  }
}

// One CTA per example.  dynamic smem: (M*C) ints + (M*C) floats
__global__ void __launch_bounds__(256) cnn_forward_kernel(CnnShape s, const float* __restrict__ x,
                                                          const float* __restrict__ Wt,
                                                          const float* __restrict__ b1,
                                                          const float* __restrict__ b2,
                                                          const float* __restrict__ b3,
                                                          DropSrc drop,
                                                          float* __restrict__ feat) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = s.M(), D = s.D(), MC = M * s.C;
  int* idx_s = reinterpret_cast<int*>(smem_raw);
  float* val_s = reinterpret_cast<float*>(idx_s + MC);
  __shared__ int count_s;
  const int b = blockIdx.x;
  compact_nonzero(x + (long)b * MC, 1, MC, idx_s, val_s, &count_s);
  __syncthreads();
  const int nnz = count_s;
  // decode each non-zero once into (row | col << 8 | channel << 16): no integer division in the tap loop
  for (int z = threadIdx.x; z < nnz; z += blockDim.x) {
    const int e = idx_s[z];
    const int ci_cell = e / s.C, ch = e - ci_cell * s.C;
    const int ri = ci_cell / s.G, ci = ci_cell - ri * s.G;
    idx_s[z] = ri | (ci << 8) | (ch << 16);
  }
  __syncthreads();
  const int CF = s.C * s.F;
  // grid.y slices the outputs of the example (the gathers below are L2-latency chains: the more threads, the
  // shorter each chain); the non-zero loop is unrolled by 4 with predicated loads so that 4 gathers are in flight.
  // Terms outside the window enter as fmaf(0, 0, acc) = acc: the sum and its order are those of the plain loop.
  for (int o = blockIdx.y * blockDim.x + threadIdx.x; o < M * D; o += gridDim.y * blockDim.x) {
    const int cell = o / D, n = o - cell * D;
    const int conv = n / s.F, f = n - conv * s.F;
    const int k = s.ksize(conv), p = k >> 1;
    const int ro = cell / s.G, co = cell - ro * s.G;
    const float* w = Wt + s.woff(conv) + f;
    float acc = __ldg((conv == 0 ? b1 : (conv == 1 ? b2 : b3)) + f);
    const int rbase = p - ro, cbase = p - co;
    for (int z0 = 0; z0 < nnz; z0 += 4) {
      float wv[4], xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int z = z0 + u;
        const int e = z < nnz ? idx_s[z] : 0;
        const int dr = (e & 255) + rbase, dc = ((e >> 8) & 255) + cbase;
        const bool in = z < nnz && (unsigned)dr < (unsigned)k && (unsigned)dc < (unsigned)k;
        wv[u] = in ? __ldg(w + (dr * k + dc) * CF + (e >> 16) * s.F) : 0.f;
        xv[u] = in ? val_s[z] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fmaf(xv[u], wv[u], acc);
    }
    acc = fmaxf(acc, 0.f);
    long oi = (long)b * M * D + o;
    if (drop.active()) acc *= drop_at(drop, oi);
    feat[oi] = acc;
  }
}

// dconv = dfeat * [feat > 0] * dropmask   (ReLU + dropout backward), elementwise
__global__ void cnn_dact_kernel(const float* __restrict__ dfeat, const float* __restrict__ feat,
                                DropSrc drop, float* __restrict__ dconv, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = feat[i] > 0.f ? dfeat[i] : 0.f;
  if (drop.active()) g *= drop_at(drop, i);
  dconv[i] = g;
}

// Weight gradient.  One CTA per input element position (cell_i, ch): gathers the examples whose
// x[b, cell_i, ch] != 0, forms sum_b x * dconv[b, :, :] for all M*D outputs, and adds each into
// the tap it belongs to.  dWt must be zeroed.  dynamic smem: B ints + B floats.
__global__ void __launch_bounds__(256) cnn_wgrad_kernel(CnnShape s, const float* __restrict__ x,
                                                        const float* __restrict__ dconv,
                                                        float* __restrict__ dWt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = s.M(), D = s.D(), MC = M * s.C;
  int* idx_s = reinterpret_cast<int*>(smem_raw);
  float* val_s = reinterpret_cast<float*>(idx_s + s.B);
  __shared__ int count_s;
  const int e = blockIdx.x;
  const int ci_cell = e / s.C, ch = e - ci_cell * s.C;
  const int ri = ci_cell / s.G, ci = ci_cell - ri * s.G;
  compact_nonzero(x + e, MC, s.B, idx_s, val_s, &count_s);
  __syncthreads();
  const int nnz = count_s;
  if (nnz == 0) return;
  for (int o = blockIdx.y * blockDim.x + threadIdx.x; o < M * D; o += gridDim.y * blockDim.x) {
    int cell = o / D, n = o - cell * D;
    int conv = n / s.F, f = n - conv * s.F;
    int k = s.ksize(conv), p = k >> 1;
    int ro = cell / s.G, co = cell - ro * s.G;
    int dr = ri - ro + p, dc = ci - co + p;
    if (dr < 0 || dr >= k || dc < 0 || dc >= k) continue;
    float acc = 0.f;
    for (int z0 = 0; z0 < nnz; z0 += 4) {   // 4 gathers in flight; same sum, same order
      float gv[4], xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int z = z0 + u;
        const bool in = z < nnz;
        gv[u] = in ? __ldg(dconv + (long)idx_s[z] * M * D + o) : 0.f;
        xv[u] = in ? val_s[z] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fmaf(xv[u], gv[u], acc);
    }
    atomicAdd(dWt + s.woff(conv) + (long)((dr * k + dc) * s.C + ch) * s.F + f, acc);
  }
}

}  // namespace gscan
