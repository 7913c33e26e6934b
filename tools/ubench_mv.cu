// Micro-benchmark of the register-resident mat-vec tile of decoder_v3.cuh (cycles per call per CTA).
#include <cstdio>
#include <cuda_runtime.h>
#include "../multimodal_seq2seq_gscan_b200/csrc/decoder_v3.cuh"
using namespace gscan;
using namespace gscan::v3;

__global__ void __launch_bounds__(512, 1) k_mv(const float* wsrc, float* out, long long* cyc, int iters, int active_warps, int sync_each) {
  __shared__ __align__(16) float x[kNB * kXS];
  for (int i = threadIdx.x; i < kNB * kXS; i += blockDim.x) x[i] = 0.001f * i;
  const int tid = threadIdx.x, ks = tid & 3, warp = tid >> 5;
  float4 w0[7], w1[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    w0[i] = ldg4(wsrc + ((tid * 14 + i) % 1000) * 4);
    w1[i] = ldg4(wsrc + ((tid * 14 + 7 + i) % 1000) * 4);
  }
  __syncthreads();
  float s = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (warp < active_warps) {
      float o[4];
      mv_rowpair(w0, w1, x, ks, o);
      s += o[0] + o[1] + o[2] + o[3];
      if (sync_each) x[(tid * 4) % (kNB * kXS)] = s * 1e-9f;
    }
    if (sync_each) __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(512, 1) k_tanh(float* out, long long* cyc, int iters, int rounds) {
  __shared__ float q[kNB * kHS], K[kNB * kM * kHS], v[kHS];
  for (int i = threadIdx.x; i < kNB * kM * kHS; i += blockDim.x) K[i] = 0.001f * (i % 777);
  for (int i = threadIdx.x; i < kNB * kHS; i += blockDim.x) q[i] = 0.01f * i;
  if (threadIdx.x < kHS) v[threadIdx.x] = 0.1f * threadIdx.x;
  __syncthreads();
  const int lane = threadIdx.x & 31, u = lane & 3;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int total = kNB * kM * 4;
    for (int base = (threadIdx.x >> 5) * 32; base < total; base += blockDim.x) {
      const int item = base + lane, pair = item >> 2;
      float s = 0.f;
      if (item < total) {
        const int n = pair / kM;
        const float* kp = K + pair * kHS + 5 * u;
        const float* qp = q + n * kHS + 5 * u;
#pragma unroll
        for (int i = 0; i < 5; ++i) s = fmaf(v[5 * u + i], act_tanh(qp[i] + kp[i]), s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      acc += s;
    }
    __syncthreads();
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float *w, *out; long long* cyc;
  cudaMalloc(&w, 16000 * 4); cudaMemset(w, 0, 16000 * 4); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 200;
  for (int sync_each = 0; sync_each < 2; ++sync_each)
    for (int threads : {256, 512})
      for (int aw : {4, 5, 8, 16}) {
        if (aw * 32 > threads) continue;
        k_mv<<<1, threads>>>(w, out, cyc, iters, aw, sync_each);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mv_rowpair threads=%d active_warps=%d sync=%d: %.0f cycles per call\n", threads, aw, sync_each, (double)c / iters);
      }
  for (int threads : {256, 512}) {
    k_tanh<<<1, threads>>>(out, cyc, iters, 0);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("visual partial scores (5760 tanh) threads=%d: %.0f cycles per call\n", threads, (double)c / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
