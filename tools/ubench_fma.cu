// Micro-benchmarks behind the decoder sweep design: issue rate of packed
// fp32 FMA (FFMA2) vs scalar FFMA as a function of independent chains and warps per SM
// sub-partition, and of broadcast LDS.128.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, bool PACKED>
__global__ void k_fma(float* out, long long* cyc, int iters, float a, float b) {
  float2 acc[CHAINS];
  float2 w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = make_float2(a + i, a - i);
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = make_float2(threadIdx.x * 1e-3f + c, b);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        if (PACKED) acc[c] = __ffma2_rn(acc[c], w[r], w[(r + c) & 7]);
        else { acc[c].x = fmaf(acc[c].x, w[r].x, w[(r + c) & 7].x); acc[c].y = fmaf(acc[c].y, w[r].y, w[(r + c) & 7].y); }
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c].x + acc[c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds(float* out, long long* cyc, int iters, int mode) {
  __shared__ __align__(16) float sm[8 * 100 + 64];
  for (int i = threadIdx.x; i < 864; i += blockDim.x) sm[i] = i * 0.001f;
  __syncthreads();
  const int ks = threadIdx.x & 3;
  float4 acc = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      int q = mode == 0 ? (4 * i + ks < 24 ? 4 * i + ks : 24) : (mode == 1 ? i : ((threadIdx.x & 31) % 25));
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        float4 v = *reinterpret_cast<const float4*>(sm + n * 100 + 4 * q);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS, bool PACKED>
void run_fma(int threads, float* out, long long* cyc) {
  const int iters = 200;
  k_fma<CHAINS, PACKED><<<1, threads>>>(out, cyc, iters, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)c / (iters * 8.0 * CHAINS);   // cycles per FFMA2 (or FFMA pair) instruction per warp
  int wps = threads / 128;
  printf("%s chains=%2d warps/SMSP=%d : %.2f cyc per warp-instr, => %.1f FMA/clk/SM\n", PACKED ? "FFMA2" : "2xFFMA", CHAINS,
         wps, per, 64.0 * (threads / 32) / per / 1.0);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  for (int threads : {128, 256, 512, 1024}) {
    run_fma<1, true>(threads, out, cyc);
    run_fma<2, true>(threads, out, cyc);
    run_fma<4, true>(threads, out, cyc);
    run_fma<8, true>(threads, out, cyc);
    run_fma<16, true>(threads, out, cyc);
    run_fma<4, false>(threads, out, cyc);
    run_fma<8, false>(threads, out, cyc);
    run_fma<16, false>(threads, out, cyc);
  }
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256, 512}) {
      k_lds<<<1, threads>>>(out, cyc, 100, mode);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("LDS.128 mode %d threads %d: %.2f cyc per warp-LDS (all warps: %.2f cyc per LDS SM-wide)\n", mode, threads,
             (double)c / (100 * 56.0), (double)c / (100 * 56.0 * (threads / 32)));
    }
  return 0;
}
