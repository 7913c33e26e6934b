// Issue rate of the legacy (mma.sync) tensor instructions on B200, one to four warps per scheduler, operands in registers:
//   tf32 m16n8k8 (1024 MAC), f16 m16n8k16 (2048 MAC), bf16 m16n8k16 (2048 MAC), tf32 m16n8k4 (512 MAC)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_hmma_rates tools/ubench_hmma_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int KIND>
__global__ void __launch_bounds__(512, 1) k(int nw, int iters, long long* cyc, float* sink) {
  const int warp = threadIdx.x >> 5;
  float d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
  uint32_t a0 = threadIdx.x * 3u, a1 = a0 ^ 0x3c003c00u, a2 = a0 + 77u, a3 = a1 + 5u, b0 = 0x3c003c00u, b1 = 0x38003800u;
  long long t0 = clock64();
  if (warp < nw) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < 13; ++s) {
        if (KIND == 0) {
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(d0[0]), "+f"(d0[1]), "+f"(d0[2]), "+f"(d0[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(d1[0]), "+f"(d1[1]), "+f"(d1[2]), "+f"(d1[3]) : "r"(a1), "r"(a0), "r"(a3), "r"(a2), "r"(b0), "r"(b1));
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(d2[0]), "+f"(d2[1]), "+f"(d2[2]), "+f"(d2[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b1), "r"(b0));
        } else if (KIND == 1) {
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(d0[0]), "+f"(d0[1]), "+f"(d0[2]), "+f"(d0[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(d1[0]), "+f"(d1[1]), "+f"(d1[2]), "+f"(d1[3]) : "r"(a1), "r"(a0), "r"(a3), "r"(a2), "r"(b0), "r"(b1));
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(d2[0]), "+f"(d2[1]), "+f"(d2[2]), "+f"(d2[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b1), "r"(b0));
        } else if (KIND == 2) {
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(d0[0]), "+f"(d0[1]), "+f"(d0[2]), "+f"(d0[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(d1[0]), "+f"(d1[1]), "+f"(d1[2]), "+f"(d1[3]) : "r"(a1), "r"(a0), "r"(a3), "r"(a2), "r"(b0), "r"(b1));
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(d2[0]), "+f"(d2[1]), "+f"(d2[2]), "+f"(d2[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b1), "r"(b0));
        } else {
          asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+f"(d0[0]), "+f"(d0[1]), "+f"(d0[2]), "+f"(d0[3]) : "r"(a0), "r"(a1), "r"(b0));
          asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+f"(d1[0]), "+f"(d1[1]), "+f"(d1[2]), "+f"(d1[3]) : "r"(a1), "r"(a0), "r"(b0));
          asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+f"(d2[0]), "+f"(d2[1]), "+f"(d2[2]), "+f"(d2[3]) : "r"(a0), "r"(a1), "r"(b1));
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (d0[0] + d1[1] + d2[2] == 123.f) sink[0] = d0[0];
}
template <int KIND> void run(const char* what, int nw, long long* cyc, float* sink) {
  const int iters = 2000;
  k<KIND><<<148, 512>>>(nw, iters, cyc, sink);
  k<KIND><<<148, 512>>>(nw, iters, cyc, sink);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0; for (int i = 0; i < 148; ++i) m += (double)h[i];
  const double per = m / 148 / iters / 39.0;   // cycles per instruction of ONE warp's stream
  const int per_smsp = (nw + 3) / 4;
  printf("%-22s warps/SM %2d (%d per scheduler): %5.2f cycles per MMA per warp -> %5.2f cycles per MMA per scheduler\n", what, nw, per_smsp, per, per / per_smsp);
}
int main() {
  long long* cyc; float* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  for (int nw : {4, 8, 16}) {
    run<0>("tf32 m16n8k8", nw, cyc, sink);
    run<3>("tf32 m16n8k4", nw, cyc, sink);
    run<1>("f16 m16n8k16", nw, cyc, sink);
    run<2>("bf16 m16n8k16", nw, cyc, sink);
  }
  return 0;
}
