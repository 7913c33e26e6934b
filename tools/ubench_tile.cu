// Cycles per call of the split-precision tensor-core mat-vec tile of decoder_v3.cuh.
#include <cstdio>
#include <cuda_runtime.h>
#include "../multimodal_seq2seq_gscan_b200/csrc/decoder_v3.cuh"
using namespace gscan;
using namespace gscan::v3;

__global__ void __launch_bounds__(512, 1) k_tile(const float* wsrc, float* out, long long* cyc, int iters, int active_warps, int sync_each) {
  extern __shared__ __align__(16) float sm[];
  float* x = sm;                                  // [8][104]
  float4* wlo = reinterpret_cast<float4*>(sm + kNB * kXS);   // [16][13][32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kNB * kXS; i += blockDim.x) x[i] = 0.001f * i;
  for (int i = tid; i < 16 * kKSteps * 32; i += blockDim.x) wlo[i] = make_float4(1e-5f * i, 2e-5f, 3e-5f, 4e-5f);
  uint32_t whi[kKSteps][4];
#pragma unroll
  for (int s = 0; s < kKSteps; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) whi[s][j] = __float_as_uint(__ldg(wsrc + (tid * 52 + s * 4 + j) % 4000));
  __syncthreads();
  const float4* wlo_lane = wlo + warp * kKSteps * 32 + lane;
  const float* x_lane = x + (lane >> 2) * kXS + 2 * (lane & 3);
  float s = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (warp < active_warps) {
      float o[4];
      mv_tile(whi, wlo_lane, x_lane, o);
      s += o[0] + o[1] + o[2] + o[3];
      if (sync_each) x[(tid * 4) % (kNB * kXS)] = s * 1e-9f;
    }
    if (sync_each) __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float *w, *out; long long* cyc;
  cudaMalloc(&w, 16000 * 4); cudaMemset(w, 0, 16000 * 4); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 200;
  size_t bytes = (kNB * kXS + 16 * kKSteps * 32 * 4) * 4;
  cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  for (int sync_each = 0; sync_each < 2; ++sync_each)
    for (int aw : {1, 2, 4, 5, 8, 16}) {
      k_tile<<<1, 512, bytes>>>(w, out, cyc, iters, aw, sync_each);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("mv_tile (39 HMMA) active_warps=%2d sync=%d: %.0f cycles per call\n", aw, sync_each, (double)c / iters);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
