// Variants of the register-resident mat-vec tile: which instruction order reaches the FFMA2 pipe rate?
#include <cstdio>
#include <cuda_runtime.h>
#include "../multimodal_seq2seq_gscan_b200/csrc/decoder_v3.cuh"
using namespace gscan;
using namespace gscan::v3;

__device__ __forceinline__ void ffma2_v(float2& acc, float2 a, float2 b) {
  unsigned long long d, ua, ub, uc;
  ua = *reinterpret_cast<unsigned long long*>(&a);
  ub = *reinterpret_cast<unsigned long long*>(&b);
  uc = *reinterpret_cast<unsigned long long*>(&acc);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(ua), "l"(ub), "l"(uc));
  acc = *reinterpret_cast<float2*>(&d);
}

// variant 1: weight-major order pinned with asm volatile (w reused by 8 consecutive FFMA2)
__device__ __forceinline__ void mv_v1(const float4 (&w0)[7], const float4 (&w1)[7], const float* __restrict__ x, int ks, float (&o)[4]) {
  float2 a0[kNB], a1[kNB];
#pragma unroll
  for (int n = 0; n < kNB; ++n) a0[n] = a1[n] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int q = min(4 * i + ks, kH / 4 - 1);
    const float* xp = x + 4 * q;
    float4 xv[kNB];
#pragma unroll
    for (int n = 0; n < kNB; ++n) xv[n] = lds4(xp + n * kXS);
#pragma unroll
    for (int n = 0; n < kNB; ++n) ffma2_v(a0[n], lo2(w0[i]), lo2(xv[n]));
#pragma unroll
    for (int n = 0; n < kNB; ++n) ffma2_v(a1[n], lo2(w1[i]), lo2(xv[n]));
#pragma unroll
    for (int n = 0; n < kNB; ++n) ffma2_v(a0[n], hi2(w0[i]), hi2(xv[n]));
#pragma unroll
    for (int n = 0; n < kNB; ++n) ffma2_v(a1[n], hi2(w1[i]), hi2(xv[n]));
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) o[m] = a0[m].x + a0[m].y + a1[m].x + a1[m].y + a0[m + 4].x + a0[m + 4].y + a1[m + 4].x + a1[m + 4].y;
}
// variant 2: as the library version but no LDS in the loop (x preloaded) -> pure FMA cost of this operand pattern
__device__ __forceinline__ void mv_v2(const float4 (&w0)[7], const float4 (&w1)[7], const float4 (&xr)[8], int ks, float (&o)[4]) {
  float2 a0[kNB], a1[kNB];
#pragma unroll
  for (int n = 0; n < kNB; ++n) a0[n] = a1[n] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
#pragma unroll
    for (int n = 0; n < kNB; ++n) {
      fma2(a0[n], lo2(w0[i]), lo2(xr[n]));
      fma2(a0[n], hi2(w0[i]), hi2(xr[n]));
      fma2(a1[n], lo2(w1[i]), lo2(xr[n]));
      fma2(a1[n], hi2(w1[i]), hi2(xr[n]));
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) o[m] = a0[m].x + a0[m].y + a1[m].x + a1[m].y + a0[m + 4].x + a0[m + 4].y + a1[m + 4].x + a1[m + 4].y;
}
// variant 3: scalar FFMA (no packing), row-pair, x broadcast from LDS.128
__device__ __forceinline__ void mv_v3(const float4 (&w0)[7], const float4 (&w1)[7], const float* __restrict__ x, int ks, float (&o)[4]) {
  float a0[kNB], a1[kNB];
#pragma unroll
  for (int n = 0; n < kNB; ++n) a0[n] = a1[n] = 0.f;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int q = min(4 * i + ks, kH / 4 - 1);
    const float* xp = x + 4 * q;
#pragma unroll
    for (int n = 0; n < kNB; ++n) {
      const float4 xv = lds4(xp + n * kXS);
      a0[n] = fmaf(w0[i].x, xv.x, a0[n]); a0[n] = fmaf(w0[i].y, xv.y, a0[n]); a0[n] = fmaf(w0[i].z, xv.z, a0[n]); a0[n] = fmaf(w0[i].w, xv.w, a0[n]);
      a1[n] = fmaf(w1[i].x, xv.x, a1[n]); a1[n] = fmaf(w1[i].y, xv.y, a1[n]); a1[n] = fmaf(w1[i].z, xv.z, a1[n]); a1[n] = fmaf(w1[i].w, xv.w, a1[n]);
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) o[m] = a0[m] + a1[m] + a0[m + 4] + a1[m + 4];
}

template <int V>
__global__ void __launch_bounds__(512, 1) k_mv(const float* wsrc, float* out, long long* cyc, int iters, int active_warps) {
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
  float4 xr[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) xr[n] = lds4(x + n * kXS + 4 * ks);
  float s = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (warp < active_warps) {
      float o[4];
      if (V == 0) mv_rowpair(w0, w1, x, ks, o);
      if (V == 1) mv_v1(w0, w1, x, ks, o);
      if (V == 2) { mv_v2(w0, w1, xr, ks, o); xr[it & 7].x += o[0]; }
      if (V == 3) mv_v3(w0, w1, x, ks, o);
      s += o[0] + o[1] + o[2] + o[3];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const float* w, float* out, long long* cyc) {
  const int iters = 200;
  for (int aw : {4, 8, 16}) {
    k_mv<V><<<1, 512>>>(w, out, cyc, iters, aw);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("variant %d active_warps=%2d: %.0f cycles per call (%.2f per FFMA2-equivalent per SMSP)\n", V, aw, (double)c / iters,
           (double)c / iters / (224.0 * aw / 4));
  }
}

int main() {
  float *w, *out; long long* cyc;
  cudaMalloc(&w, 16000 * 4); cudaMemset(w, 0, 16000 * 4); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  run<0>(w, out, cyc); run<1>(w, out, cyc); run<2>(w, out, cyc); run<3>(w, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
