// Generic fp32-accurate GEMM on the tensor cores, used for every batched (off-the-recurrence)
// contraction of the path: key projections, input-gate pre-activations of all steps at once,
// the output projection, and all weight gradients (as split-K "TN" products).
//
//   C[i,j] (ldc) (+)= act( sum_k A(i,k) * B(k,j) + bias[j] + bias2[j] )
//   A(i,k) = A[i*a_rs + k*a_cs],  B(k,j) = B[k*b_rs + j*b_cs]
//
// so NT / NN / TN forms are all the same kernel with different strides; the tile loaders pick the
// lane mapping that makes the contiguous axis the coalesced one.
//
// Arithmetic: mma.sync.m16n8k8 tf32 in split precision (3xTF32): every fp32 operand x is used as
// x_hi = top 19 bits (what the tensor core reads) and x_lo = x - x_hi, and A.B is accumulated in fp32
// as A_hi B_hi + A_lo B_hi + A_hi B_lo.  The dropped A_lo B_lo term is ~2^-21 relative, which keeps
// the 1e-4 parity bar against the fp32 reference with two orders of magnitude to spare.
// (The operands are fp32 activations / gradients produced on the fly, so a tcgen05 + TMA pipeline
// would need a separate hi/lo staging pass per operand tile)
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace gscan {

struct GemmP {
  const float* A; long a_rs, a_cs;
  const float* B; long b_rs, b_cs;
  float* C; long ldc;
  int M, N, K;
  const float* bias; const float* bias2;
  int act;         // 0 none, 1 tanh, 2 relu
  int accumulate;  // C += result
  int kchunk;      // K range per blockIdx.z (multiple of BK); gridDim.z > 1 => atomicAdd epilogue
};

constexpr int GBM = 128, GBN = 64, GBK = 32, GTHREADS = 256, GSTAGES = 3;
constexpr int GLDK = GBK + 4;    // row stride of a k-contiguous operand tile  [rows][GLDK]
constexpr int GLDAM = GBM + 8;   // row stride of an m-contiguous A tile       [GBK][GLDAM]
constexpr int GLDBN = GBN + 8;   // row stride of an n-contiguous B tile       [GBK][GLDBN]
// (GLDK % 32 == 4, GLDAM % 32 == 8, GLDBN % 32 == 8: every fragment load below is bank-conflict free)

__device__ __forceinline__ void gemm_mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// round-to-nearest tf32 part of x (integer add + mask; cvt.rna.tf32.f32 is several times slower to issue);
// the remainder x - hi is the low part (|lo| <= 2^-12 |x|, either sign,
// so the dropped lo*lo terms stay ~2^-24 relative and do not accumulate a bias over a long K)
__device__ __forceinline__ uint32_t tf32_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void cp_async16(float* dst, const float* src, int bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, int bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__host__ __device__ constexpr int gemm_a_stage(bool ak) { return ak ? GBM * GLDK : GBK * GLDAM; }
__host__ __device__ constexpr int gemm_b_stage(bool bk) { return bk ? GBN * GLDK : GBK * GLDBN; }

// AK / BKF: the operand is contiguous along k in global memory (else along m / n).
// VEC: 16-byte copies are legal (base pointers and leading dimensions 16-byte aligned).
template <bool AK, bool BKF, bool VEC>
__global__ void __launch_bounds__(GTHREADS, 2) sgemm_kernel(GemmP p) {
  extern __shared__ __align__(16) float gsm[];
  constexpr int ASZ = gemm_a_stage(AK), BSZ = gemm_b_stage(BKF);
  float* As = gsm;
  float* Bs = gsm + GSTAGES * ASZ;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;   // 4 x 2 warps, 32 x 32 outputs each
  const int fg = lane >> 2, ft = lane & 3;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int k_begin = blockIdx.z * p.kchunk;
  const int k_end = min(p.K, k_begin + p.kchunk);
  const int nk = k_end > k_begin ? (k_end - k_begin + GBK - 1) / GBK : 0;
  const long lda = AK ? p.a_rs : p.a_cs, ldb = BKF ? p.b_cs : p.b_rs;

  auto issue = [&](int stage, int kt) {
    float* as = As + stage * ASZ;
    float* bs = Bs + stage * BSZ;
    if (AK) {   // [GBM rows][GBK k]
      if (VEC) {
#pragma unroll
        for (int i = 0; i < GBM * GBK / 4 / GTHREADS; ++i) {
          const int c = tid + i * GTHREADS, r = c / (GBK / 4), kq = c % (GBK / 4);
          const int gm = m0 + r, gk = kt + 4 * kq;
          const int bytes = gm < p.M ? max(0, min(16, (k_end - gk) * 4)) : 0;
          cp_async16(as + r * GLDK + 4 * kq, bytes ? p.A + (long)gm * lda + gk : p.A, bytes);
        }
      } else {
#pragma unroll
        for (int i = 0; i < GBM * GBK / GTHREADS; ++i) {
          const int e = tid + i * GTHREADS, r = e / GBK, k = e % GBK;
          const int gm = m0 + r, gk = kt + k;
          const int bytes = (gm < p.M && gk < k_end) ? 4 : 0;
          cp_async4(as + r * GLDK + k, bytes ? p.A + (long)gm * lda + gk : p.A, bytes);
        }
      }
    } else {    // [GBK k][GBM m]
      if (VEC) {
#pragma unroll
        for (int i = 0; i < GBM * GBK / 4 / GTHREADS; ++i) {
          const int c = tid + i * GTHREADS, k = c / (GBM / 4), mq = c % (GBM / 4);
          const int gm = m0 + 4 * mq, gk = kt + k;
          const int bytes = gk < k_end ? max(0, min(16, (p.M - gm) * 4)) : 0;
          cp_async16(as + k * GLDAM + 4 * mq, bytes ? p.A + (long)gk * lda + gm : p.A, bytes);
        }
      } else {
#pragma unroll
        for (int i = 0; i < GBM * GBK / GTHREADS; ++i) {
          const int e = tid + i * GTHREADS, k = e / GBM, m = e % GBM;
          const int gm = m0 + m, gk = kt + k;
          const int bytes = (gm < p.M && gk < k_end) ? 4 : 0;
          cp_async4(as + k * GLDAM + m, bytes ? p.A + (long)gk * lda + gm : p.A, bytes);
        }
      }
    }
    if (BKF) {  // [GBN rows][GBK k]
      if (VEC) {
#pragma unroll
        for (int i = 0; i < GBN * GBK / 4 / GTHREADS; ++i) {
          const int c = tid + i * GTHREADS, r = c / (GBK / 4), kq = c % (GBK / 4);
          const int gn = n0 + r, gk = kt + 4 * kq;
          const int bytes = gn < p.N ? max(0, min(16, (k_end - gk) * 4)) : 0;
          cp_async16(bs + r * GLDK + 4 * kq, bytes ? p.B + (long)gn * ldb + gk : p.B, bytes);
        }
      } else {
#pragma unroll
        for (int i = 0; i < GBN * GBK / GTHREADS; ++i) {
          const int e = tid + i * GTHREADS, r = e / GBK, k = e % GBK;
          const int gn = n0 + r, gk = kt + k;
          const int bytes = (gn < p.N && gk < k_end) ? 4 : 0;
          cp_async4(bs + r * GLDK + k, bytes ? p.B + (long)gn * ldb + gk : p.B, bytes);
        }
      }
    } else {    // [GBK k][GBN n]
      if (VEC) {
#pragma unroll
        for (int i = 0; i < GBN * GBK / 4 / GTHREADS; ++i) {
          const int c = tid + i * GTHREADS, k = c / (GBN / 4), nq = c % (GBN / 4);
          const int gn = n0 + 4 * nq, gk = kt + k;
          const int bytes = gk < k_end ? max(0, min(16, (p.N - gn) * 4)) : 0;
          cp_async16(bs + k * GLDBN + 4 * nq, bytes ? p.B + (long)gk * ldb + gn : p.B, bytes);
        }
      } else {
#pragma unroll
        for (int i = 0; i < GBN * GBK / GTHREADS; ++i) {
          const int e = tid + i * GTHREADS, k = e / GBN, n = e % GBN;
          const int gn = n0 + n, gk = kt + k;
          const int bytes = (gn < p.N && gk < k_end) ? 4 : 0;
          cp_async4(bs + k * GLDBN + n, bytes ? p.B + (long)gk * ldb + gn : p.B, bytes);
        }
      }
    }
  };

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) {
    if (s < nk) issue(s, k_begin + s * GBK);
    cp_async_commit();
  }
  for (int it = 0; it < nk; ++it) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();   // tile `it` has landed for everyone; the stage refilled below was consumed in iteration it-1
    if (it + GSTAGES - 1 < nk) issue((it + GSTAGES - 1) % GSTAGES, k_begin + (it + GSTAGES - 1) * GBK);
    cp_async_commit();
    const float* as = As + (it % GSTAGES) * ASZ;
    const float* bs = Bs + (it % GSTAGES) * BSZ;
    // The tensor core adds into its fp32 accumulator with truncation, which over a K of thousands grows into a
    // bias of ~K/8 half-ulps of the running sum.  So each 32-deep tile is accumulated from zero on the tensor
    // core and folded into the running sum with a round-to-nearest fp32 add.
    float tmp[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) tmp[i][j][c] = 0.f;
#pragma unroll
    for (int k8 = 0; k8 < GBK; k8 += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int mrow = wm * 32 + mt * 16 + fg;
        float v[4];
        if (AK) {
          v[0] = as[mrow * GLDK + k8 + ft]; v[1] = as[(mrow + 8) * GLDK + k8 + ft];
          v[2] = as[mrow * GLDK + k8 + ft + 4]; v[3] = as[(mrow + 8) * GLDK + k8 + ft + 4];
        } else {
          v[0] = as[(k8 + ft) * GLDAM + mrow]; v[1] = as[(k8 + ft) * GLDAM + mrow + 8];
          v[2] = as[(k8 + ft + 4) * GLDAM + mrow]; v[3] = as[(k8 + ft + 4) * GLDAM + mrow + 8];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ah[mt][c] = tf32_rna(v[c]);
          al[mt][c] = __float_as_uint(v[c] - __uint_as_float(ah[mt][c]));
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int ncol = wn * 32 + nt * 8 + fg;
        float v[2];
        if (BKF) { v[0] = bs[ncol * GLDK + k8 + ft]; v[1] = bs[ncol * GLDK + k8 + ft + 4]; }
        else { v[0] = bs[(k8 + ft) * GLDBN + ncol]; v[1] = bs[(k8 + ft + 4) * GLDBN + ncol]; }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          bh[nt][c] = tf32_rna(v[c]);
          bl[nt][c] = __float_as_uint(v[c] - __uint_as_float(bh[nt][c]));
        }
      }
      // term-major order: 8 independent accumulators between two dependent MMAs
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) gemm_mma_tf32(tmp[mt][nt], al[mt], bh[nt]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) gemm_mma_tf32(tmp[mt][nt], ah[mt], bl[nt]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) gemm_mma_tf32(tmp[mt][nt], ah[mt], bh[nt]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][j][c] += tmp[i][j][c];
  }
  cp_async_wait<0>();

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int gm = m0 + wm * 32 + mt * 16 + fg + 8 * (c >> 1);
        const int gn = n0 + wn * 32 + nt * 8 + 2 * ft + (c & 1);
        if (gm >= p.M || gn >= p.N) continue;
        float* cp = p.C + (long)gm * p.ldc + gn;
        if (split) {
          atomicAdd(cp, acc[mt][nt][c]);
        } else {
          float r = acc[mt][nt][c];
          if (p.bias) r += __ldg(p.bias + gn);
          if (p.bias2) r += __ldg(p.bias2 + gn);
          if (p.act == 1) r = act_tanh(r);
          else if (p.act == 2) r = fmaxf(r, 0.f);
          if (p.accumulate) r += *cp;
          *cp = r;
        }
      }
}

template <bool AK, bool BKF, bool VEC>
inline int launch_sgemm_t(const GemmP& p, dim3 grid, cudaStream_t st) {
  constexpr size_t bytes = sizeof(float) * GSTAGES * (gemm_a_stage(AK) + gemm_b_stage(BKF));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sgemm_kernel<AK, BKF, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  sgemm_kernel<AK, BKF, VEC><<<grid, GTHREADS, bytes, st>>>(p);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

// Launch helper.  ksplit > 1 requires C to be initialised (zero, or the value to accumulate
// onto) and forbids bias / act.  Each operand must be contiguous along one of its two axes.
inline int launch_sgemm(const float* A, long a_rs, long a_cs, const float* B, long b_rs, long b_cs,
                        float* C, long ldc, int M, int N, int K, const float* bias, const float* bias2,
                        int act, int accumulate, int ksplit, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  GemmP p{A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, bias, bias2, act, accumulate, 0};
  if (ksplit < 1) ksplit = 1;
  int kchunk = ceil_div(ceil_div(K, ksplit), GBK) * GBK;
  if (kchunk < GBK) kchunk = GBK;
  ksplit = K > 0 ? ceil_div(K, kchunk) : 1;
  p.kchunk = (ksplit == 1) ? max(K, 1) : kchunk;
  dim3 grid(ceil_div(N, GBN), ceil_div(M, GBM), ksplit);
  const bool ak = (a_cs == 1), bk = (b_rs == 1);
  if ((!ak && a_rs != 1) || (!bk && b_cs != 1)) return -2;   // GSCAN_E_UNSUPPORTED: no unit stride
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const long lda = ak ? a_rs : a_cs, ldb = bk ? b_cs : b_rs;
  const bool vec = al16(A) && al16(B) && (lda % 4 == 0) && (ldb % 4 == 0);
  if (ak && bk) return vec ? launch_sgemm_t<true, true, true>(p, grid, st) : launch_sgemm_t<true, true, false>(p, grid, st);
  if (ak && !bk) return vec ? launch_sgemm_t<true, false, true>(p, grid, st) : launch_sgemm_t<true, false, false>(p, grid, st);
  if (!ak && !bk) return vec ? launch_sgemm_t<false, false, true>(p, grid, st) : launch_sgemm_t<false, false, false>(p, grid, st);
  return vec ? launch_sgemm_t<false, true, true>(p, grid, st) : launch_sgemm_t<false, true, false>(p, grid, st);
}

// Path choice for one product: the tcgen05 kernel (gemm_tc.cuh) when TMA can address both operands and the
// product is large enough to amortise its pipeline fill; the mma.sync kernel above otherwise.
inline bool use_tc(const float* A, long a_rs, long a_cs, const float* B, long b_rs, long b_cs, int M, int N, int K) {
  return (long)M * N * K >= (1L << 26) && tc::eligible(A, a_rs, a_cs, B, b_rs, b_cs, M, N, K);
}
inline int launch_gemm(const float* A, long a_rs, long a_cs, const float* B, long b_rs, long b_cs,
                       float* C, long ldc, int M, int N, int K, const float* bias, const float* bias2,
                       int act, int accumulate, int ksplit, cudaStream_t st) {
  if (use_tc(A, a_rs, a_cs, B, b_rs, b_cs, M, N, K))
    return tc::launch(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, bias, bias2, act, accumulate, ksplit, st);
  return launch_sgemm(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, bias, bias2, act, accumulate, ksplit, st);
}

// Weight-gradient form: C[N1,N2] (ldc) = sum_r X[r, i] * Y[r, j] over R rows (R large).
// Chooses a split so the grid covers the chip; C is zeroed here first.
inline int launch_grad_gemm(const float* X, long ldx, const float* Y, long ldy, float* C, long ldc,
                            int N1, int N2, int R, int num_sms, cudaStream_t st) {
  // zero the destination block (it may be a column block of a wider matrix)
  cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, N2 * sizeof(float), N1, st);
  if (e != cudaSuccess) return (int)e;
  if (use_tc(X, 1, ldx, Y, ldy, 1, N1, N2, R)) {
    const int tiles = ceil_div(N1, tc::BM) * ceil_div(N2, tc::BN);
    const int ksplit = max(1, min(ceil_div(R, 4 * tc::BK), num_sms / tiles));   // one persistent wave
    return tc::launch(X, 1, ldx, Y, ldy, 1, C, ldc, N1, N2, R, nullptr, nullptr, 0, 0, ksplit, st);
  }
  int tiles = ceil_div(N1, GBM) * ceil_div(N2, GBN);
  int ksplit = max(1, min(ceil_div(R, 4 * GBK), (2 * num_sms) / tiles));   // one wave at 2 CTAs per SM
  return launch_sgemm(X, 1, ldx, Y, ldy, 1, C, ldc, N1, N2, R, nullptr, nullptr, 0, 0, ksplit, st);
}

// A family of weight-gradient products over the same R rows: the ones the tcgen05 kernel can address go out as ONE
// grouped launch (tc::launch_group_tn), the rest one by one.
inline int launch_grad_group(const tc::GroupProblem* probs, int n, int R, int num_sms, cudaStream_t st) {
  tc::GroupProblem grouped[tc::MAXG];
  int ng = 0;
  static const bool no_group = getenv("GSCAN_NO_GROUP_GEMM") != nullptr;
  for (int i = 0; i < n; ++i) {
    const tc::GroupProblem& q = probs[i];
    if (!no_group && ng < tc::MAXG && use_tc(q.X, 1, q.ldx, q.Y, q.ldy, 1, q.N1, q.N2, R)) {
      grouped[ng++] = q;
    } else {
      int rc = launch_grad_gemm(q.X, q.ldx, q.Y, q.ldy, q.C, q.ldc, q.N1, q.N2, R, num_sms, st);
      if (rc) return rc;
    }
  }
  if (ng == 1) return launch_grad_gemm(grouped[0].X, grouped[0].ldx, grouped[0].Y, grouped[0].ldy, grouped[0].C,
                                       grouped[0].ldc, grouped[0].N1, grouped[0].N2, R, num_sms, st);
  return tc::launch_group_tn(grouped, ng, R, st);
}

// out[j] = sum_r X[r*ldx + j]  for j < N  (bias gradients).  out is overwritten.
__global__ void colsum_kernel(const float* __restrict__ X, long ldx, int R, int N, int rows_per_block,
                              float* __restrict__ out) {
  __shared__ float red[8][33];
  int tx = threadIdx.x, ty = threadIdx.y;
  int j = blockIdx.x * 32 + tx;
  int r0 = blockIdx.y * rows_per_block;
  int r1 = min(R, r0 + rows_per_block);
  float s = 0.f;
  if (j < N)
    for (int r = r0 + ty; r < r1; r += 8) s += __ldg(X + (long)r * ldx + j);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(out + j, t);
  }
}

inline int launch_colsum(const float* X, long ldx, int R, int N, float* out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * N, st);
  if (e != cudaSuccess) return (int)e;
  if (R <= 0) return 0;
  int rows_per_block = 256;
  dim3 grid(ceil_div(N, 32), ceil_div(R, rows_per_block));
  colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(X, ldx, R, N, rows_per_block, out);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace gscan
