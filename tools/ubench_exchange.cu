// Micro-benchmark of the cluster exchanges of the decoder sweeps (decoder_v3.cuh): what ONE all-to-all exchange
// inside a 5-CTA cluster costs on B200, for the store patterns the sweeps use or could use.  25 clusters of 5 CTAs x
// 512 threads run concurrently (the geometry of the real sweep).  Every iteration is two dependent exchanges
// (separate mbarriers and buffers, like consecutive phases of a decoder step); reported: cycles per exchange.
//
//   mode 0  ping: one thread sends 4 bytes to each of the 5 CTAs - the floor (st.async + complete_tx + try_wait)
//   mode 1  N floats per CTA to all 5 CTAs, 4-byte st.async, four lanes per item (lane u -> CTA u, lane 0 also -> CTA 4):
//           the round-1 pattern of the attention-score exchanges (N = 288 visual, 80 text)
//   mode 2  the same bytes as 16-byte st.async (N/4 groups: lane k -> CTA k, lane 0 also -> CTA 4): round 2
//   mode 3  the same bytes written to LOCAL shared memory, block barrier, then 5 bulk DSMEM copies
//           (cp.async.bulk.shared::cluster.shared::cta) issued by 5 threads
//   mode 4  16-byte st.async, one lane per group sends to all 5 CTAs
//   mode 5  gather: 160 threads each send one float to all 5 CTAs (800 scalar stores: the round-1 h / q' / c_V gathers)
//   mode 6  gather as 16-byte stores (40 groups x 5)
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_exchange tools/ubench_exchange.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kC = 5, kThreads = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_async_f32(uint32_t a, float v, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(a), "r"(__float_as_uint(v)), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async_f32x4(uint32_t a, float4 v, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(a),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_copy_s2s(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

struct P { int mode, N, iters; long long* cycles; float* sink; };

__global__ void __cluster_dims__(kC, 1, 1) __launch_bounds__(kThreads, 1) exchange_kernel(P p) {
  extern __shared__ __align__(16) float smem[];
  // layout: recv[2][kC][N] | local[2][N] | bars
  const int N = p.N;
  float* recv = smem;
  float* local = smem + 2 * kC * N;
  const uint32_t base = smem_u32(smem);
  const uint32_t bar_off = (uint32_t)(2 * kC * N + 2 * N) * 4u;
  const uint32_t bar0 = base + bar_off;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  uint32_t rb[kC];
#pragma unroll
  for (int d = 0; d < kC; ++d) rb[d] = mapa_u32(base, d);
  const uint32_t rb_u = mapa_u32(base, lane & 3), rb_4 = rb[4];
  if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = tid; i < 2 * kC * N + 2 * N; i += kThreads) smem[i] = 0.f;
  __syncthreads();
  cluster_barrier();
  float acc = 0.f;
  uint32_t bytes = 0;
  switch (p.mode) {
    case 0: bytes = kC * 4; break;
    default: bytes = (uint32_t)(kC * N * 4); break;
  }
  long long t0 = clock64();
  for (int it = 0; it < p.iters; ++it) {
    const uint32_t par = (uint32_t)(it & 1);
#pragma unroll 1
    for (int x = 0; x < 2; ++x) {
      const uint32_t bar = bar0 + 8u * x, boff = bar_off + 8u * x;
      const uint32_t roff = (uint32_t)(x * kC * N + rank * N) * 4u;   // this rank's slot in every receiver
      if (tid == 0) mbar_arm(bar, bytes);
      const float val = acc + (float)(it + tid);
      if (p.mode == 0) {
        if (tid < kC) st_async_f32(rb[0] * 0 + mapa_u32(base, tid) + roff, val, mapa_u32(base, tid) + boff);
      } else if (p.mode == 1) {
        const int total = N * 4, u = lane & 3;
        for (int b = warp * 32; b < total; b += kThreads) {
          const int item = b + lane, pair = item >> 2;
          if (item < total) {
            st_async_f32(rb_u + roff + pair * 4u, val, rb_u + boff);
            if (u == 0) st_async_f32(rb_4 + roff + pair * 4u, val, rb_4 + boff);
          }
        }
      } else if (p.mode == 2) {
        const int k = lane & 3;
        for (int b = warp * 32; b < N; b += kThreads) {   // one lane per item, groups of four lanes
          const int item = b + lane;
          if (item < N) {
            const float4 v = make_float4(val, val, val, val);
            const uint32_t off = roff + (uint32_t)(item & ~3) * 4u;
            st_async_f32x4(rb_u + off, v, rb_u + boff);
            if (k == 0) st_async_f32x4(rb_4 + off, v, rb_4 + boff);
          }
        }
      } else if (p.mode == 3) {
        for (int i = tid; i < N; i += kThreads) local[x * N + i] = val;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < kC) {
          const uint32_t dst = mapa_u32(base, tid);
          bulk_copy_s2s(dst + roff, base + (uint32_t)(2 * kC * N + x * N) * 4u, (uint32_t)(N * 4), dst + boff);
        }
      } else if (p.mode == 4) {
        for (int g = tid; g < N / 4; g += kThreads) {
          const float4 v = make_float4(val, val, val, val);
#pragma unroll
          for (int d = 0; d < kC; ++d) st_async_f32x4(rb[d] + roff + g * 16u, v, rb[d] + boff);
        }
      } else if (p.mode == 5) {
        if (tid < N) {
#pragma unroll
          for (int d = 0; d < kC; ++d) st_async_f32(rb[d] + roff + tid * 4u, val, rb[d] + boff);
        }
      } else if (p.mode == 6) {
        if (tid < N) {
          const int k = lane & 3;
          const float4 v = make_float4(val, val, val, val);
          const uint32_t off = roff + (uint32_t)(tid & ~3) * 4u;
          st_async_f32x4(rb_u + off, v, rb_u + boff);
          if (k == 0) st_async_f32x4(rb_4 + off, v, rb_4 + boff);
        }
      }
      mbar_wait(bar, par);
      // consume: every thread reads one received word (keeps the dependence chain of a real phase)
      acc += recv[x * kC * N + (tid % (kC * (p.mode == 0 ? 1 : N)))] * 1e-30f;
      if (p.mode == 3) __syncthreads();   // local[] may be rewritten only after everybody has read (as a real phase would)
    }
  }
  long long t1 = clock64();
  if (tid == 0) p.cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.f) p.sink[0] = acc;
  cluster_barrier();
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  const int nclusters = 25;
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, sizeof(long long) * nclusters * kC);
  cudaMalloc(&sink, 4);
  struct Case { int mode, N; const char* what; };
  const Case cases[] = {
      {0, 4, "ping (5 x 4 B)"},
      {1, 288, "visual scores, 4-byte st.async, 4 lanes/item (round 1)"},
      {2, 288, "visual scores, 16-byte st.async"},
      {3, 288, "visual scores, local STS + barrier + 5 bulk DSMEM copies"},
      {4, 288, "visual scores, 16-byte st.async, one lane -> all 5"},
      {1, 80, "text scores, 4-byte st.async (round 1)"},
      {2, 80, "text scores, 16-byte st.async"},
      {3, 80, "text scores, bulk copies"},
      {5, 160, "gather of 160 floats/CTA, 4-byte st.async x 5 (round 1: h, q', c_V)"},
      {6, 160, "gather of 160 floats/CTA, 16-byte st.async"},
      {3, 160, "gather of 160 floats/CTA, bulk copies"},
  };
  for (const Case& c : cases) {
    P p{c.mode, c.N, iters, cyc, sink};
    size_t smem = (size_t)(2 * kC * c.N + 2 * c.N) * 4 + 64;
    cudaFuncSetAttribute(exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    exchange_kernel<<<nclusters * kC, kThreads, smem>>>(p);   // warm-up
    exchange_kernel<<<nclusters * kC, kThreads, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d N %d: %s\n", c.mode, c.N, cudaGetErrorString(e)); return 1; }
    long long h[nclusters * kC];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0, mx = 0;
    for (int i = 0; i < nclusters * kC; ++i) { mean += (double)h[i]; if ((double)h[i] > mx) mx = (double)h[i]; }
    mean /= nclusters * kC;
    printf("mode %d N %3d  %-72s  %7.0f cycles/exchange (max CTA %7.0f)\n", c.mode, c.N, c.what, mean / (2.0 * iters), mx / (2.0 * iters));
  }
  return 0;
}
