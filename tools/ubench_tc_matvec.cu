// Micro-benchmark: ONE recurrent mat-vec stage of the decoder sweeps on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulator in TMEM), to compare with the mma.sync stage the sweeps ship
// (tools/ubench_mvtile.cu: 515 cycles for the 3xTF32 tile, fewer for the f16x3 one).
//
// The stage is the recurrence  x_{t+1} = tanh(W x_t)  for a batch of N = 16 vectors (a cluster of the sweep holds 8
// examples; 16 is the smallest N of an M = 128 tcgen05.mma), W = [128 x K] resident in shared memory (the per-CTA
// weight slice of the sweep is 80 x 104, padded here to the MMA's M = 128 and K = 104 or 128), fp32 accuracy by the
// 3xTF32 split (W_lo x_hi + W_hi x_lo + W_hi x_hi; truncation split: the raw fp32 word is the hi operand).
// Per step, exactly what a sweep stage would have to do:
//   1. 128 threads write the new x (hi and lo words) into the K-major 128-byte-swizzled B tile,
//      fence.proxy.async (generic -> async proxy), tcgen05.fence::before_thread_sync, block barrier;
//   2. one thread issues PASSES x K/8 tcgen05.mma (M = 128, N = 16, K = 8 each) and a tcgen05.commit on an mbarrier;
//   3. everybody waits on the mbarrier, tcgen05.ld 32x32b.x16 (row m of D for the 16 vectors), tcgen05.wait::ld;
//   4. tanh, next step.
// Reported: cycles per step and where they go (thread 0's clock: barrier -> MMAs issued -> commit seen -> TMEM read).
// The first step is checked against an fp64 host product.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_tc_matvec tools/ubench_tc_matvec.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int M = 128, N = 16, KMAX = 128;
constexpr int A_CHUNK = M * 128;          // bytes of one 32-column K chunk of A (128 rows x 128 B)
constexpr int B_CHUNK = N * 128;
constexpr int TMEM_COLS = 128;   // up to 8 accumulators of 16 columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major operand, 128-byte swizzle: rows of 128 B (32 fp32 of K), 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((1024u >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
// byte offset of element (row r, column k) of a K-major swizzled operand whose 32-column chunks are `chunk` bytes apart
__host__ __device__ __forceinline__ uint32_t sw_off(int r, int k, int chunk) {
  const int c = k >> 5, kk = k & 31;
  return (uint32_t)(c * chunk + (r >> 3) * 1024 + (r & 7) * 128 + ((((kk >> 2) ^ (r & 7)) & 7) << 4) + (kk & 3) * 4);
}
__device__ __forceinline__ uint64_t koff_a(int k) { return (uint64_t)(((k >> 2) * A_CHUNK + (k & 3) * 32) >> 4); }
__device__ __forceinline__ uint64_t koff_b(int k) { return (uint64_t)(((k >> 2) * B_CHUNK + (k & 3) * 32) >> 4); }
__device__ __forceinline__ float fast_tanh(float x) { float e = __expf(2.f * x); return 1.f - __fdividef(2.f, 1.f + e); }

struct P {
  const float* W;    // [M][KMAX]
  const float* x0;   // [N][KMAX]
  float* D0;         // [M][N]  product of the first step (validation)
  long long* cyc;    // [grid][6]: total, write+barrier, issue, commit wait, tmem read, tanh
  int iters, variant;   // variant 1: no tanh (x = D / 2); 2: no MMA at all (TMEM read only)
};

template <int K, int PASSES, int NACC>
__global__ void __launch_bounds__(128, 1) tc_matvec_kernel(P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_hi = smem;                       // 4 chunks x 16 KB
  uint8_t* a_lo = a_hi + 4 * A_CHUNK;
  uint8_t* b_hi = a_lo + 4 * A_CHUNK;         // 4 chunks x 2 KB
  uint8_t* b_lo = b_hi + 4 * B_CHUNK;
  uint64_t* bar = (uint64_t*)(b_lo + 4 * B_CHUNK);
  uint32_t* tslot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_a = smem_u32(bar);

  for (int i = tid; i < M * KMAX; i += 128) {
    const int r = i / KMAX, k = i % KMAX;
    const float w = k < K ? p.W[i] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    *(float*)(a_hi + sw_off(r, k, A_CHUNK)) = w;          // the tensor core truncates to tf32 by itself
    *(float*)(a_lo + sw_off(r, k, A_CHUNK)) = w - hi;
  }
  if (tid == 0) { mbar_init(bar_a, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = *tslot;

  // thread tid owns row m = tid of D and column k = tid of the next x
  float xv[N];
#pragma unroll
  for (int n = 0; n < N; ++n) xv[n] = tid < K ? p.x0[n * KMAX + tid] : 0.f;
  constexpr int KSTEPS = (K + 7) / 8;
  const uint64_t dA_hi = make_desc(smem_u32(a_hi)), dA_lo = make_desc(smem_u32(a_lo));
  const uint64_t dB_hi = make_desc(smem_u32(b_hi)), dB_lo = make_desc(smem_u32(b_lo));
  long long c_write = 0, c_issue = 0, c_commit = 0, c_read = 0, c_tanh = 0;
  const long long t_begin = clock64();
  for (int it = 0; it < p.iters; ++it) {
    const long long t0 = clock64();
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const uint32_t o = sw_off(n, tid, B_CHUNK);
      const float hi = __uint_as_float(__float_as_uint(xv[n]) & 0xffffe000u);
      *(float*)(b_hi + o) = xv[n];
      *(float*)(b_lo + o) = xv[n] - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const long long t1 = clock64();
    if (warp == 0 && p.variant != 2) {
      uint32_t elected = 0;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
      if (elected) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // small terms first: W_lo x_hi, W_hi x_lo, W_hi x_hi  (a single pass is W_hi x_hi); descriptors differ only in
        // the 14-bit start-address field.  MMA j accumulates into TMEM accumulator j % NACC: MMAs on ONE accumulator
        // are a dependent chain.  Everything is unrolled with compile-time offsets: an MMA is ~3 uniform instructions.
#pragma unroll
        for (int j = 0; j < PASSES * KSTEPS; ++j) {
          const int pass = PASSES == 1 ? 2 : j / KSTEPS, k = j % KSTEPS;
          mma_tf32(tacc + 16u * (j % NACC), (pass == 0 ? dA_lo : dA_hi) + koff_a(k), (pass == 1 ? dB_lo : dB_hi) + koff_b(k), kIdesc, j >= NACC ? 1u : 0u);
        }
        mma_commit(bar_a);
      }
      __syncwarp();
    }
    const long long t2 = clock64();
    if (p.variant != 2) mbar_wait(bar_a, (uint32_t)(it & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t3 = clock64();
    uint32_t d[16];
    tmem_ld16(tacc + ((uint32_t)(warp * 32) << 16), d);
    float dv[N];
#pragma unroll
    for (int n = 0; n < N; ++n) dv[n] = __uint_as_float(d[n]);
    for (int a = 1; a < NACC; ++a) {
      tmem_ld16(tacc + ((uint32_t)(warp * 32) << 16) + 16u * a, d);
#pragma unroll
      for (int n = 0; n < N; ++n) dv[n] += __uint_as_float(d[n]);
    }
    const long long t3b = clock64();
    if (it == 0) {
#pragma unroll
      for (int n = 0; n < N; ++n) p.D0[tid * N + n] = dv[n];
    }
    if (p.variant) {
#pragma unroll
      for (int n = 0; n < N; ++n) xv[n] = 0.5f * dv[n];
    } else {
      float e[N];
#pragma unroll
      for (int n = 0; n < N; ++n) e[n] = 1.f + __expf(2.f * dv[n]);
#pragma unroll
      for (int n = 0; n < N; ++n) xv[n] = 1.f - __fdividef(2.f, e[n]);
    }
    if (tid >= K) {
#pragma unroll
      for (int n = 0; n < N; ++n) xv[n] = 0.f;
    }
    const long long t4 = clock64();
    c_write += t1 - t0; c_issue += t2 - t1; c_commit += t3 - t2; c_read += t3b - t3; c_tanh += t4 - t3b;
  }
  const long long t_end = clock64();
  if (tid == 0) {
    long long* c = p.cyc + (size_t)blockIdx.x * 6;
    c[0] = t_end - t_begin; c[1] = c_write; c[2] = c_issue; c[3] = c_commit; c[4] = c_read; c[5] = c_tanh;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tacc), "r"(TMEM_COLS) : "memory");
  }
}

template <int K, int PASSES, int NACC>
static int run_case(const P& p, int variant, int sms, size_t smem, const std::vector<float>& W, const std::vector<float>& x0, const char* what) {
  P q = p;
  q.variant = variant;
  cudaFuncSetAttribute(tc_matvec_kernel<K, PASSES, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_matvec_kernel<K, PASSES, NACC><<<sms, 128, smem>>>(q);
  tc_matvec_kernel<K, PASSES, NACC><<<sms, 128, smem>>>(q);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); return 1; }
  std::vector<long long> h(6 * sms);
  std::vector<float> D(M * N);
  cudaMemcpy(h.data(), q.cyc, h.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), q.D0, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)W[m * KMAX + k] * (double)x0[n * KMAX + k];
      worst = fmax(worst, fabs(s - (double)D[m * N + n]));
    }
  double m5[6] = {0, 0, 0, 0, 0, 0};
  for (int b = 0; b < sms; ++b)
    for (int j = 0; j < 6; ++j) m5[j] += (double)h[b * 6 + j] / sms / q.iters;
  printf("%-92s %6.0f cycles/step  (write x + fences + barrier %4.0f, issue %4.0f, wait for the commit %4.0f, tcgen05.ld %4.0f, tanh %4.0f)  max |err| of step 0 %.2e\n",
         what, m5[0], m5[1], m5[2], m5[3], m5[4], m5[5], worst);
  return 0;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  std::vector<float> W(M * KMAX), x0(N * KMAX);
  srand(1);
  for (auto& v : W) v = ((float)rand() / RAND_MAX - 0.5f) * 0.2f;
  for (auto& v : x0) v = ((float)rand() / RAND_MAX - 0.5f) * 2.f;
  float *dW, *dx, *dD;
  long long* dc;
  cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dx, x0.size() * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dc, sizeof(long long) * 6 * sms);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dx, x0.data(), x0.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 8 * A_CHUNK + 8 * B_CHUNK + 64 + 1024;
  P p{dW, dx, dD, dc, iters, 0};
  int rc = 0;
  rc |= run_case<104, 3, 1>(p, 0, sms, smem, W, x0, "K = 104 (the sweep's [x|h] row), 3xTF32: 39 MMAs on one TMEM accumulator");
  rc |= run_case<104, 3, 3>(p, 0, sms, smem, W, x0, "  the same on 3 TMEM accumulators (summed after tcgen05.ld)");
  rc |= run_case<104, 3, 6>(p, 0, sms, smem, W, x0, "  the same on 6 TMEM accumulators");
  rc |= run_case<104, 3, 8>(p, 0, sms, smem, W, x0, "  the same on 8 TMEM accumulators");
  rc |= run_case<104, 3, 8>(p, 1, sms, smem, W, x0, "  8 accumulators, x = D / 2 instead of tanh");
  rc |= run_case<128, 3, 8>(p, 0, sms, smem, W, x0, "K = 128, 3xTF32: 48 MMAs, 8 accumulators");
  rc |= run_case<104, 1, 1>(p, 0, sms, smem, W, x0, "K = 104, one TF32 pass: 13 MMAs, one accumulator (accuracy 1e-3, not shippable)");
  rc |= run_case<104, 1, 4>(p, 0, sms, smem, W, x0, "  the same on 4 accumulators");
  rc |= run_case<8, 1, 1>(p, 0, sms, smem, W, x0, "K = 8, ONE MMA: the floor of write + fence + barrier + issue + commit + TMEM read + tanh");
  rc |= run_case<8, 1, 1>(p, 1, sms, smem, W, x0, "K = 8, one MMA, x = D / 2 instead of tanh");
  rc |= run_case<8, 1, 1>(p, 2, sms, smem, W, x0, "no MMA: write x + fences + barrier + tcgen05.ld only");
  return rc;
}
