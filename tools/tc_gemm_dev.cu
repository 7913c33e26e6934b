// Stand-alone check + timing of the tcgen05 3xTF32 GEMM (csrc/gemm_tc.cuh) against fp64 on the host
// and against the mma.sync kernel (csrc/gemm.cuh).  Build: see tools/build_tc_gemm_dev.sh
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cmath>
#include <cstring>
#include "../include/gscan_b200.h"
#include "../multimodal_seq2seq_gscan_b200/csrc/gemm.cuh"
#include "../multimodal_seq2seq_gscan_b200/csrc/gemm_tc.cuh"

using namespace gscan;

struct Case { const char* name; int M, N, K; bool ak, bk; int ksplit; bool bias; int act; long lda, ldb, ldc; };

static double check(const Case& c, const std::vector<float>& A, const std::vector<float>& B, const std::vector<float>& bias,
                    const std::vector<float>& C, long a_rs, long a_cs, long b_rs, long b_cs, double* scale_out) {
  std::mt19937 rng(7);
  double worst = 0, scale = 0;
  int samples = 4000;
  for (int s = 0; s < samples; ++s) {
    int i = rng() % c.M, j = rng() % c.N;
    if (s < 8) { i = (s & 1) ? c.M - 1 : 0; j = (s & 2) ? c.N - 1 : 0; }
    double acc = 0, mag = 0;
    for (int k = 0; k < c.K; ++k) {
      double t = (double)A[i * a_rs + k * a_cs] * (double)B[k * b_rs + j * b_cs];
      acc += t; mag += fabs(t);
    }
    if (c.bias) acc += bias[j];
    if (c.act == 1) acc = tanh(acc);
    double err = fabs(acc - (double)C[(long)i * c.ldc + j]);
    worst = fmax(worst, err / fmax(mag / sqrt((double)c.K), 1e-30));   // relative to the typical sum magnitude
    scale = fmax(scale, fabs(acc));
  }
  *scale_out = scale;
  return worst;
}

int main(int argc, char** argv) {
  cudaStream_t st; cudaStreamCreate(&st);
  if (getenv("MN_LAYOUT")) tc::mn_config().layout = atoi(getenv("MN_LAYOUT"));
  if (getenv("MN_SBO")) tc::mn_config().sbo = atoi(getenv("MN_SBO"));
  if (getenv("MN_LBO")) tc::mn_config().lbo = atoi(getenv("MN_LBO"));
  if (getenv("MN_SWZ")) tc::mn_config().tma_swizzle = atoi(getenv("MN_SWZ"));
  const char* only = getenv("ONLY");
  Case cases[] = {
    {"NT out_proj    ", 24200, 100, 400, true, true, 1, false, 0, 400, 400, 100},
    {"NT xe  bias    ", 24200, 400, 100, true, true, 1, true, 0, 400, 100, 400},
    {"NT tanh small  ", 200, 100, 100, true, true, 1, true, 1, 100, 100, 100},
    {"NT K=152 tail  ", 7200, 100, 150, true, true, 1, false, 0, 152, 152, 100},
    {"NN dU          ", 24200, 400, 100, true, false, 1, false, 0, 100, 400, 400},
    {"TN wgrad split ", 400, 100, 24200, false, false, 37, false, 0, 400, 400, 300},
    {"TN wgrad 400x200", 400, 200, 24200, false, false, 18, false, 0, 400, 400, 300},
    {"TN 100x100     ", 100, 100, 24200, false, false, 148, false, 0, 100, 400, 100},
    {"TK mixed       ", 300, 260, 1000, false, true, 1, false, 0, 304, 1000, 260},
  };
  for (const Case& c : cases) {
    if (only && !strstr(c.name, only)) continue;
    long a_rs = c.ak ? c.lda : 1, a_cs = c.ak ? 1 : c.lda;
    long b_rs = c.bk ? 1 : c.ldb, b_cs = c.bk ? c.ldb : 1;
    size_t na = c.ak ? (size_t)c.M * c.lda : (size_t)c.K * c.lda;
    size_t nb = c.bk ? (size_t)c.N * c.ldb : (size_t)c.K * c.ldb;
    size_t nc = (size_t)c.M * c.ldc;
    std::vector<float> A(na), B(nb), bias(c.N), C(nc), C2(nc);
    std::mt19937 rng(123);
    std::normal_distribution<float> nd(0.f, 1.f);
    for (auto& x : A) x = nd(rng);
    for (auto& x : B) x = nd(rng) * 0.1f;
    for (auto& x : bias) x = nd(rng);
    float *dA, *dB, *dC, *dC2, *dbias;
    cudaMalloc(&dA, na * 4); cudaMalloc(&dB, nb * 4); cudaMalloc(&dC, nc * 4); cudaMalloc(&dC2, nc * 4); cudaMalloc(&dbias, c.N * 4);
    cudaMemcpy(dA, A.data(), na * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), nb * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dbias, bias.data(), c.N * 4, cudaMemcpyHostToDevice);
    if (!tc::eligible(dA, a_rs, a_cs, dB, b_rs, b_cs, c.M, c.N, c.K)) { printf("%s not eligible\n", c.name); continue; }
    cudaMemset(dC, 0, nc * 4); cudaMemset(dC2, 0, nc * 4);
    int rc = tc::launch(dA, a_rs, a_cs, dB, b_rs, b_cs, dC, c.ldc, c.M, c.N, c.K, c.bias ? dbias : nullptr, nullptr, c.act, 0, c.ksplit, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc || e != cudaSuccess) { printf("%s FAILED rc=%d err=%s\n", c.name, rc, cudaGetErrorString(e)); return 1; }
    launch_sgemm(dA, a_rs, a_cs, dB, b_rs, b_cs, dC2, c.ldc, c.M, c.N, c.K, c.bias ? dbias : nullptr, nullptr, c.act, 0, c.ksplit, st);
    cudaStreamSynchronize(st);
    cudaMemcpy(C.data(), dC, nc * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(C2.data(), dC2, nc * 4, cudaMemcpyDeviceToHost);
    double scale, scale2;
    double err = check(c, A, B, bias, C, a_rs, a_cs, b_rs, b_cs, &scale);
    double err2 = check(c, A, B, bias, C2, a_rs, a_cs, b_rs, b_cs, &scale2);
    double maxdiff = 0;
    for (int i = 0; i < c.M; ++i) for (int j = 0; j < c.N; ++j) maxdiff = fmax(maxdiff, fabs((double)C[(long)i * c.ldc + j] - (double)C2[(long)i * c.ldc + j]));
    // timing
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_tc, ms_old; const int reps = 20;
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; ++r) tc::launch(dA, a_rs, a_cs, dB, b_rs, b_cs, dC, c.ldc, c.M, c.N, c.K, nullptr, nullptr, 0, 0, c.ksplit, st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_tc, e0, e1);
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; ++r) launch_sgemm(dA, a_rs, a_cs, dB, b_rs, b_cs, dC2, c.ldc, c.M, c.N, c.K, nullptr, nullptr, 0, 0, c.ksplit, st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_old, e0, e1);
#ifdef GSCAN_TC_TIMELINE
    {
      long long* dtl; cudaMalloc(&dtl, 5 * 128 * 8); cudaMemset(dtl, 0, 5 * 128 * 8);
      tc::mn_config().timeline = dtl;
      tc::launch(dA, a_rs, a_cs, dB, b_rs, b_cs, dC, c.ldc, c.M, c.N, c.K, nullptr, nullptr, 0, 0, c.ksplit, st);
      cudaStreamSynchronize(st);
      tc::mn_config().timeline = nullptr;
      std::vector<long long> tl(640);
      cudaMemcpy(tl.data(), dtl, 640 * 8, cudaMemcpyDeviceToHost);
      long long t0 = tl[0];
      const char* names[5] = {"tma issue ", "mma start ", "cvt start ", "acc window", "write-out "};
      for (int r = 0; r < 5; ++r) {
        printf("  %s:", names[r]);
        for (int i = 0; i < 28 && tl[r * 128 + i]; ++i) printf(" %lld", tl[r * 128 + i] - t0);
        if (r == 4) { printf("\n   after STS:"); for (int i = 65; i < 65 + 20; ++i) printf(" %lld", tl[r * 128 + i] ? tl[r * 128 + i] - t0 : 0); }
        printf("\n");
      }
      cudaFree(dtl);
    }
#endif
    printf("%s M=%d N=%d K=%d split=%d | tcgen05 err %.2e (mma.sync err %.2e) maxdiff %.2e scale %.2f | %.1f us vs %.1f us\n",
           c.name, c.M, c.N, c.K, c.ksplit, err, err2, maxdiff, scale, 1e3 * ms_tc / reps, 1e3 * ms_old / reps);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dC2); cudaFree(dbias);
  }
  return 0;
}
