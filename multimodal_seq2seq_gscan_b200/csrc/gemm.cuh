// Generic fp32 GEMM on the FMA pipe (FFMA2), used for every batched (off-the-recurrence)
// contraction of the path: key projections, input-gate pre-activations of all steps at once,
// the output projection, and all weight gradients (as split-K "TN" products).
//
//   C[i,j] (ldc) (+)= act( sum_k A(i,k) * B(k,j) + bias[j] + bias2[j] )
//   A(i,k) = A[i*a_rs + k*a_cs],  B(k,j) = B[k*b_rs + j*b_cs]
//
// so NT / NN / TN forms are all the same kernel with different strides; the tile loaders pick the
// lane mapping that makes the contiguous axis the coalesced one.
#pragma once
#include "common.cuh"

namespace gscan {

struct GemmP {
  const float* A; long a_rs, a_cs;
  const float* B; long b_rs, b_cs;
  float* C; long ldc;
  int M, N, K;
  const float* bias; const float* bias2;
  int act;         // 0 none, 1 tanh, 2 relu
  int accumulate;  // C += result
  int kchunk;      // K range per blockIdx.z (multiple of BK); gridDim.z > 1 => atomicAdd epilogue
};

constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;

__global__ void __launch_bounds__(GTHREADS) sgemm_kernel(GemmP p) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int k_begin = blockIdx.z * p.kchunk;
  const int k_end = min(p.K, k_begin + p.kchunk);
  const bool a_kfast = (p.a_cs == 1);
  const bool b_kfast = (p.b_rs == 1);

  float a_reg[8], b_reg[4];
  auto load_tiles = [&](int kt) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int e = tid + i * GTHREADS;
      int kk = a_kfast ? (e & (GBK - 1)) : (e / GBM);
      int mm = a_kfast ? (e / GBK) : (e % GBM);
      int gm = m0 + mm, gk = kt + kk;
      a_reg[i] = (gm < p.M && gk < k_end) ? __ldg(p.A + (long)gm * p.a_rs + (long)gk * p.a_cs) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * GTHREADS;
      int kk = b_kfast ? (e & (GBK - 1)) : (e / GBN);
      int nn = b_kfast ? (e / GBK) : (e % GBN);
      int gn = n0 + nn, gk = kt + kk;
      b_reg[i] = (gn < p.N && gk < k_end) ? __ldg(p.B + (long)gk * p.b_rs + (long)gn * p.b_cs) : 0.f;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int e = tid + i * GTHREADS;
      int kk = a_kfast ? (e & (GBK - 1)) : (e / GBM);
      int mm = a_kfast ? (e / GBK) : (e % GBM);
      As[kk][mm] = a_reg[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * GTHREADS;
      int kk = b_kfast ? (e & (GBK - 1)) : (e / GBN);
      int nn = b_kfast ? (e / GBK) : (e % GBN);
      Bs[kk][nn] = b_reg[i];
    }
  };

  float2 acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);

  if (k_begin < k_end) {
    load_tiles(k_begin);
    store_tiles();
  }
  __syncthreads();
  for (int kt = k_begin; kt < k_end; kt += GBK) {
    const bool more = (kt + GBK < k_end);
    if (more) load_tiles(kt + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float2 b01 = make_float2(b.x, b.y), b23 = make_float2(b.z, b.w);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 aa = make_float2(av[i], av[i]);
        fma2(acc[i][0], aa, b01);
        fma2(acc[i][1], aa, b23);
      }
    }
    __syncthreads();
    if (more) {
      store_tiles();
      __syncthreads();
    }
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= p.M) continue;
    float v[4] = {acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= p.N) continue;
      float* c = p.C + (long)gm * p.ldc + gn;
      if (split) {
        atomicAdd(c, v[j]);
      } else {
        float r = v[j];
        if (p.bias) r += __ldg(p.bias + gn);
        if (p.bias2) r += __ldg(p.bias2 + gn);
        if (p.act == 1) r = act_tanh(r);
        else if (p.act == 2) r = fmaxf(r, 0.f);
        if (p.accumulate) r += *c;
        *c = r;
      }
    }
  }
}

// Launch helper.  ksplit > 1 requires C to be initialised (zero, or the value to accumulate
// onto) and forbids bias / act.
inline int launch_sgemm(const float* A, long a_rs, long a_cs, const float* B, long b_rs, long b_cs,
                        float* C, long ldc, int M, int N, int K, const float* bias, const float* bias2,
                        int act, int accumulate, int ksplit, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  GemmP p{A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, bias, bias2, act, accumulate, 0};
  if (ksplit < 1) ksplit = 1;
  int kchunk = ceil_div(ceil_div(K, ksplit), GBK) * GBK;
  if (kchunk < GBK) kchunk = GBK;
  ksplit = K > 0 ? ceil_div(K, kchunk) : 1;
  p.kchunk = (ksplit == 1) ? max(K, 1) : kchunk;
  dim3 grid(ceil_div(N, GBN), ceil_div(M, GBM), ksplit);
  sgemm_kernel<<<grid, GTHREADS, 0, st>>>(p);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

// Weight-gradient form: C[N1,N2] (ldc) = sum_r X[r, i] * Y[r, j] over R rows (R large).
// Chooses a split so the grid covers the chip; C is zeroed here first.
inline int launch_grad_gemm(const float* X, long ldx, const float* Y, long ldy, float* C, long ldc,
                            int N1, int N2, int R, int num_sms, cudaStream_t st) {
  // zero the destination block (it may be a column block of a wider matrix)
  cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, N2 * sizeof(float), N1, st);
  if (e != cudaSuccess) return (int)e;
  int tiles = ceil_div(N1, GBM) * ceil_div(N2, GBN);
  int ksplit = max(1, min(ceil_div(R, 4 * GBK), ceil_div(2 * num_sms, tiles)));
  return launch_sgemm(X, 1, ldx, Y, ldy, 1, C, ldc, N1, N2, R, nullptr, nullptr, 0, 0, ksplit, st);
}

// out[j] = sum_r X[r*ldx + j]  for j < N  (bias gradients).  out is overwritten.
__global__ void colsum_kernel(const float* __restrict__ X, long ldx, int R, int N, int rows_per_block,
                              float* __restrict__ out) {
  __shared__ float red[8][33];
  int tx = threadIdx.x, ty = threadIdx.y;
  int j = blockIdx.x * 32 + tx;
  int r0 = blockIdx.y * rows_per_block;
  int r1 = min(R, r0 + rows_per_block);
  float s = 0.f;
  if (j < N)
    for (int r = r0 + ty; r < r1; r += 8) s += __ldg(X + (long)r * ldx + j);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(out + j, t);
  }
}

inline int launch_colsum(const float* X, long ldx, int R, int N, float* out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * N, st);
  if (e != cudaSuccess) return (int)e;
  if (R <= 0) return 0;
  int rows_per_block = 256;
  dim3 grid(ceil_div(N, 32), ceil_div(R, rows_per_block));
  colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(X, ldx, R, N, rows_per_block, out);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace gscan
