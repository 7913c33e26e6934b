// Issue rate / latency of the legacy tensor path (mma.sync m16n8k8 tf32, m16n8k16 bf16) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS, bool BF16>
__global__ void k(float* out, long long* cyc, int iters) {
  unsigned a[8][4], b[2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = __float_as_uint(1.0f + 0.001f * (threadIdx.x + i * 4 + j)) & 0xffffe000u;
  b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
  float d[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[c][j] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        if (BF16) mma_bf16(d[c], a[i], b); else mma_tf32(d[c], a[i], b);
      }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS, bool BF16>
void run(int threads, float* out, long long* cyc) {
  const int iters = 100;
  k<CHAINS, BF16><<<1, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)c / (iters * 8.0 * CHAINS);
  int warps = threads / 32;
  double macs = BF16 ? 2048.0 : 1024.0;
  printf("%s chains=%d warps/SMSP=%d: %.2f cyc per mma per warp -> %.0f MAC/clk/SM\n", BF16 ? "bf16 m16n8k16" : "tf32 m16n8k8", CHAINS,
         warps / 4, per, macs * warps / per);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  for (int threads : {128, 256, 512}) {
    run<1, false>(threads, out, cyc); run<2, false>(threads, out, cyc); run<3, false>(threads, out, cyc); run<6, false>(threads, out, cyc);
    run<1, true>(threads, out, cyc); run<3, true>(threads, out, cyc); run<6, true>(threads, out, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
