// What bounds one mat-vec stage of the sweeps (decoder_v3.cuh mv_tile: [16 x 104] x [104 x 8] on mma.sync m16n8k8
// tf32, 3xTF32, W_hi in registers, W_lo and x from shared memory)?  One CTA per SM, `nw` warps run the stage back to back
// `iters` times, each iteration depending on the previous one through x (as consecutive decoder steps do).
//   variant 0  mv_tile as shipped
//   variant 1  two accumulator sets (even / odd k-steps)
//   variant 2  x pre-split into tf32 hi / lo in shared memory (no ALU between the loads and the HMMA)
//   variant 3  HMMA only (operands in registers): the pipe / latency floor
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_mvtile tools/ubench_mvtile.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../multimodal_seq2seq_gscan_b200/csrc/decoder_v3.cuh"
using namespace gscan;
using namespace gscan::v3;

template <int VAR>
__global__ void __launch_bounds__(512, 1) k(int nw, int iters, long long* cyc, float* sink) {
  extern __shared__ __align__(16) float smem[];
  float* x_s = smem;
  float* xh_s = smem + kNB * kXS;
  float* xl_s = smem + 2 * kNB * kXS;
  float4* wlo_s = reinterpret_cast<float4*>(smem + 3 * kNB * kXS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, fg = lane >> 2, ft = lane & 3;
  for (int i = tid; i < kNB * kXS; i += 512) { x_s[i] = 0.01f * (i % 37); xh_s[i] = x_s[i]; xl_s[i] = 1e-5f * (i % 11); }
  for (int i = tid; i < 16 * kKSteps * 32; i += 512) wlo_s[i] = make_float4(1e-4f, 2e-4f, 3e-4f, 4e-4f);
  uint32_t whi[kKSteps][4];
#pragma unroll
  for (int s = 0; s < kKSteps; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) whi[s][j] = tf32_hi(0.01f * (float)((lane + s + j) % 13));
  __syncthreads();
  const float4* wlo_lane = wlo_s + warp * kKSteps * 32 + lane;
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  long long t0 = clock64();
  if (warp < nw) {
    for (int it = 0; it < iters; ++it) {
      if (VAR == 0) {
        mv_tile(whi, wlo_lane, x_s + fg * kXS + 2 * ft, o);
      } else if (VAR == 1) {
        float d0[2][4] = {}, d1[2][4] = {}, d2[2][4] = {};
        const float* x_lane = x_s + fg * kXS + 2 * ft;
#pragma unroll
        for (int s = 0; s < kKSteps; ++s) {
          const float2 xv = *reinterpret_cast<const float2*>(x_lane + 8 * s);
          const float4 lo = wlo_lane[s * 32];
          const uint32_t bh0 = tf32_hi(xv.x), bh1 = tf32_hi(xv.y);
          const uint32_t bl0 = __float_as_uint(xv.x - __uint_as_float(bh0)), bl1 = __float_as_uint(xv.y - __uint_as_float(bh1));
          mma_tf32(d0[s & 1], whi[s][0], whi[s][1], whi[s][2], whi[s][3], bh0, bh1);
          mma_tf32(d1[s & 1], __float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), __float_as_uint(lo.w), bh0, bh1);
          mma_tf32(d2[s & 1], whi[s][0], whi[s][1], whi[s][2], whi[s][3], bl0, bl1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (d0[0][j] + d0[1][j]) + ((d1[0][j] + d1[1][j]) + (d2[0][j] + d2[1][j]));
      } else if (VAR == 2) {
        float d0[4] = {}, d1[4] = {}, d2[4] = {};
        const float* xh = xh_s + fg * kXS + 2 * ft;
        const float* xl = xl_s + fg * kXS + 2 * ft;
#pragma unroll
        for (int s = 0; s < kKSteps; ++s) {
          const float2 h = *reinterpret_cast<const float2*>(xh + 8 * s), l = *reinterpret_cast<const float2*>(xl + 8 * s);
          const float4 lo = wlo_lane[s * 32];
          mma_tf32(d0, whi[s][0], whi[s][1], whi[s][2], whi[s][3], __float_as_uint(h.x), __float_as_uint(h.y));
          mma_tf32(d1, __float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), __float_as_uint(lo.w), __float_as_uint(h.x), __float_as_uint(h.y));
          mma_tf32(d2, whi[s][0], whi[s][1], whi[s][2], whi[s][3], __float_as_uint(l.x), __float_as_uint(l.y));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = d0[j] + (d1[j] + d2[j]);
      } else {
        float d0[4] = {}, d1[4] = {}, d2[4] = {};
        const uint32_t b0 = __float_as_uint(o[0]) & 0xffffe000u, b1 = __float_as_uint(o[1]) & 0xffffe000u;
#pragma unroll
        for (int s = 0; s < kKSteps; ++s) {
          mma_tf32(d0, whi[s][0], whi[s][1], whi[s][2], whi[s][3], b0, b1);
          mma_tf32(d1, whi[s][1], whi[s][0], whi[s][3], whi[s][2], b0, b1);
          mma_tf32(d2, whi[s][0], whi[s][1], whi[s][2], whi[s][3], b1, b0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = d0[j] + (d1[j] + d2[j]);
      }
      // next step's x depends on this step's result (through shared memory, as in the sweep)
      x_s[(fg * kXS + 2 * ft + 8 * (it % kKSteps))] = o[0] * 1e-6f;
      xh_s[(fg * kXS + 2 * ft + 8 * (it % kKSteps))] = o[1] * 1e-6f;
      __syncwarp();
    }
  }
  long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (o[0] + o[1] + o[2] + o[3] == 123.f) sink[0] = o[0];
}

template <int VAR>
void run(const char* what, int nw, int iters, long long* cyc, float* sink) {
  const size_t bytes = (3 * kNB * kXS + 16 * kKSteps * 32 * 4) * sizeof(float);
  cudaFuncSetAttribute(k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  k<VAR><<<148, 512, bytes>>>(nw, iters, cyc, sink);
  k<VAR><<<148, 512, bytes>>>(nw, iters, cyc, sink);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < 148; ++i) m += (double)h[i];
  printf("%-58s warps %2d  %7.0f cycles per stage\n", what, nw, m / 148 / iters);
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int nw : {1, 2, 4, 5, 8, 16}) {
    run<0>("mv_tile as shipped", nw, iters, cyc, sink);
    run<1>("two accumulator sets", nw, iters, cyc, sink);
    run<2>("x pre-split (no ALU between LDS and HMMA)", nw, iters, cyc, sink);
    run<3>("HMMA only (operands in registers)", nw, iters, cyc, sink);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return 0;
}
