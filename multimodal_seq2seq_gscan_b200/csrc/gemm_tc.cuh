// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05, kind::tf32) for the large batched
// contractions of the path: the output projection over all decoder steps, the input-gate pre-GEMM,
// the data gradients of those, and every decoder weight gradient (split-K "TN" products with
// K = Tt*B = 24,200).  Same contract as sgemm_kernel (gemm.cuh):
//
//   C[i,j] (ldc) (+)= act( sum_k A(i,k) * B(k,j) + bias[j] + bias2[j] )
//
// with each operand contiguous along one axis (K-major or MN-major); both majors are consumed
// directly from the layout TMA lands them in (128-byte swizzle), so nothing is transposed in HBM.
//
// Arithmetic: 3xTF32.  The operands are fp32 activations / gradients produced on the fly, so the
// hi / lo split happens in shared memory: TMA lands the raw fp32 tile, which the tensor core reads as
// hi = trunc_tf32(x) as it is; four converter warps write lo = rn_tf32(x - hi) to a twin buffer at the
// SAME byte offset (element-wise, hence swizzle-agnostic), and the MMA thread issues lo*hi + hi*lo + hi*hi
// into a TMEM accumulator.  The dropped lo*lo term is ~2^-20 relative.  The tensor core adds into its fp32
// accumulator with truncation (measured: error grows linearly with the number of MMAs), so every FLUSH
// K-blocks the window sum is folded into register accumulators with round-to-nearest adds.
//
// What bounds it: shared-memory bandwidth, not the tensor pipe.  Per 128x128x32 K-block the SM moves
// 32 KB (TMA in) + 64 KB (converter read + lo write) + 96 KB (12 SS-mode MMAs x 8 KB of operand fetch)
// = 192 KB through a 128 B/clk pipe = 1500 clk, against 768 clk of MMA issue (tools/tc_gemm_dev.cu timeline).
//
// Persistent kernel, one CTA per SM walking a static tile list (m-tile, n-tile, k-split).
// Roles (320 threads, 3-stage ring of 64 KB stages, two TMEM accumulator buffers):
//   warp 0      TMA producer        empty[s]    -> full_raw[s] (expect_tx)
//   warps 2-5   hi/lo converters    full_raw[s] -> full_cvt[s] (fence.proxy.async + arrive)
//   warp 1      MMA issuer          full_cvt[s], acc_empty[b] -> tcgen05.mma x12 -> tcgen05.commit -> empty[s], acc_full[b]
//   warps 6-9   accumulate/epilogue acc_full[b] -> tcgen05.ld -> fp32 RN add into registers -> acc_empty[b];
//                                   after the last K-block: bias/act -> coalesced global stores (atomicAdd for split-K)
#pragma once
#include <cuda.h>
#include <cstring>
#include <cstdlib>
#include "common.cuh"

namespace gscan {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;          // BK fp32 = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;              // 16 KB (A and B tiles have the same size: BM == BN)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;          // A_hi | A_lo | B_hi | B_lo
constexpr int THREADS = 320;
constexpr int CVT_THREADS = 128;
constexpr int EPI_THREADS = 128;
#ifndef GSCAN_TC_FLUSH
#define GSCAN_TC_FLUSH 2
#endif
constexpr int FLUSH = GSCAN_TC_FLUSH;                             // K-blocks accumulated in TMEM before the sum is folded into registers
constexpr int ACC_BUFS = 4;                          // TMEM accumulator buffers (4 x 128 columns = all of TMEM)
constexpr int TMEM_COLS = ACC_BUFS * BN;
constexpr int EPI_BYTES = 4 * 32 * 36 * 4;           // per-warp 32x36 transpose tiles for the write-out
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct Params {
  float* C; long ldc;
  int M, N, K;
  const float* bias; const float* bias2;
  int act;          // 0 none, 1 tanh, 2 relu
  int accumulate;   // C += result (single split only)
  int kb_per;       // K-blocks per split
  int kb_total;
  int ksplit;
  uint32_t mn_layout, mn_sbo, mn_lbo;   // descriptor fields of an MN-major operand
  long long* timeline;                  // optional [4 roles][128] clock64 stamps of CTA 0 (tools/tc_gemm_dev.cu)
};
// Grouped launch: several products that share K (and the operand majors) walk ONE persistent tile list, so that a
// family of small weight-gradient products costs one launch, one pipeline fill and one wave of split-K atomics.
constexpr int MAXG = 12;
struct alignas(64) GroupMaps { CUtensorMap a[MAXG]; CUtensorMap b[MAXG]; };
struct GroupTable {
  float* C[MAXG]; long ldc[MAXG];
  int M[MAXG], N[MAXG];
  int n_tiles[MAXG];   // N tiles of problem g
  int mn_end[MAXG];    // cumulative count of (m, n) tiles up to and including problem g
  int ng;
};
#ifndef TC_VEC_STORE
#define TC_VEC_STORE 0
#endif
#ifdef GSCAN_TC_TIMELINE
#define TC_STAMP(role, idx) do { if (p.timeline && blockIdx.x == 0 && (idx) < 128) p.timeline[(role) * 128 + (idx)] = clock64(); } while (0)
#else
#define TC_STAMP(role, idx) do { } while (0)
#endif
// MN-major tf32 operands exist only in the 32-byte-atom flavour of the 128-byte swizzle (layout type 1, TMA
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 4 k-rows of 128 B per swizzle atom, atoms SBO = 512 B apart.
struct MnConfig { uint32_t layout = 1, sbo = 512, lbo = BK * 128; int tma_swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; long long* timeline = nullptr; };
inline MnConfig& mn_config() { static MnConfig c; return c; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1", 128-byte swizzle).
//   K-major : rows of 128 B (32 fp32 of K), 8-row swizzle atoms 1024 B apart (SBO); LBO unused.
//             One MMA consumes K = 8 fp32 = 32 B: advance the start address by 32 B per k-step.
//   MN-major: rows of 128 B (32 fp32 of M/N) per k; 8 k-rows = one 1024-B atom (SBO between atoms);
//             the next 32 M/N elements start LBO = BK*128 B later (one TMA box per 32 columns).
//             One MMA consumes 8 k-rows = one atom: advance the start address by 1024 B per k-step.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// instruction descriptor: D fp32, A/B tf32, M = 128, N = n (multiple of 16), majors as given
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// hi part of the split.  Default: truncation, which is what the tensor core does by itself when it reads an fp32
// word as tf32 - so the raw tile IS the hi operand and only lo = x - trunc(x) has to be written (measured:
// 5e-6 worst error relative to the typical |sum| at K = 400, against 3e-6 for the round-to-nearest split, which
// costs another 32 KB of shared-memory writes per K-block; -DGSCAN_TC_RN_SPLIT selects it).
#ifdef GSCAN_TC_RN_SPLIT
__device__ __forceinline__ uint32_t tf32_hi_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
#else
__device__ __forceinline__ uint32_t tf32_hi_bits(float x) { return __float_as_uint(x) & 0xffffe000u; }
#endif
__device__ __forceinline__ uint32_t tf32_rn_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// One lane's column of a 32-row block: 8 independent shared loads, then 8 global stores (128 B per warp each).
// MODE 0 store, 1 tanh, 2 relu, 3 accumulate onto C, 4 atomic add (split-K partial sums).
template <int MODE>
__device__ __forceinline__ void store_rows(float* dst, long ldc, const float* src, int rows, float bsum) {
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = src[(r0 + i) * 33] + bsum;
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (r0 + i < rows) v[i] += dst[(long)(r0 + i) * ldc];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (r0 + i < rows) {
        float* d = dst + (long)(r0 + i) * ldc;
        if (MODE == 1) *d = act_tanh(v[i]);
        else if (MODE == 2) *d = fmaxf(v[i], 0.f);
        else if (MODE == 4) atomicAdd(d, v[i]);
        else *d = v[i];
      }
    }
  }
}

// Vector form: the warp covers 4 rows x 128 B per instruction (lane = row r0 + lane/8, 16-byte chunk lane%8);
// `src` is the warp's transpose tile with a row pitch of 36 floats.  Needs 16-byte aligned rows of C.
template <int MODE>
__device__ __forceinline__ void store_rows_v4(float* dst, long ldc, const float* tile, int lane, int rows, int cols,
                                              float4 bsum) {
  const int rr = lane >> 3, cc = (lane & 7) * 4;
  if (cc >= cols) return;
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 16) {
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = *reinterpret_cast<const float4*>(tile + (r0 + 4 * i + rr) * 36 + cc);
      v[i].x += bsum.x; v[i].y += bsum.y; v[i].z += bsum.z; v[i].w += bsum.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + 4 * i + rr;
      if (r < rows) {
        float* d = dst + (long)r * ldc + cc;
        if (MODE == 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(v[i].x), "f"(v[i].y), "f"(v[i].z),
                       "f"(v[i].w) : "memory");
        } else {
          if (MODE == 3) {
            const float4 o = *reinterpret_cast<const float4*>(d);
            v[i].x += o.x; v[i].y += o.y; v[i].z += o.z; v[i].w += o.w;
          } else if (MODE == 1) {
            v[i].x = act_tanh(v[i].x); v[i].y = act_tanh(v[i].y); v[i].z = act_tanh(v[i].z); v[i].w = act_tanh(v[i].w);
          } else if (MODE == 2) {
            v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f); v[i].z = fmaxf(v[i].z, 0.f); v[i].w = fmaxf(v[i].w, 0.f);
          }
          *reinterpret_cast<float4*>(d) = v[i];
        }
      }
    }
  }
}

// AK / BKM: operand contiguous along K in global memory (else along M / N).  GROUP: tmA / tmB are arrays indexed by
// the problem number and the per-problem shapes come from `gt` (Params then carries only the shared K split).
template <bool AK, bool BKM, bool GROUP>
__device__ __forceinline__ void tc_body(const CUtensorMap* tmA, const CUtensorMap* tmB, const Params& p,
                                        const GroupTable* gt) {
  extern __shared__ uint8_t tc_smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = tc_smem_raw + (base - smem_u32(tc_smem_raw));
  const uint32_t bar0 = base + STAGES * STAGE_BYTES + EPI_BYTES;
  auto full_raw = [&](int s) { return bar0 + 8u * s; };
  auto full_cvt = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto empty = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (3 * STAGES + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (3 * STAGES + ACC_BUFS + b); };
  const uint32_t tmem_slot = bar0 + 8u * (3 * STAGES + 2 * ACC_BUFS);

  if (threadIdx.x == 0) {
    const int nmaps = GROUP ? gt->ng : 1;
    for (int g = 0; g < nmaps; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(tmA + g) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(tmB + g) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_raw(s), 1);
      mbar_init(full_cvt(s), CVT_THREADS);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < ACC_BUFS; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  // persistent static schedule: tile = ((m-tile * n_tiles + n-tile) * ksplit + z); every role walks the same list
  const int n_tiles = ceil_div(p.N, BN);
  const int total_tiles = (GROUP ? gt->mn_end[gt->ng - 1] : ceil_div(p.M, BM) * n_tiles) * p.ksplit;
  auto tile_coords = [&](int tile, int& g, int& m0, int& n0, int& kb0, int& nkb) {
    const int z = tile % p.ksplit;
    int mn = tile / p.ksplit;
    g = 0;
    int nt = n_tiles;
    if (GROUP) {
      while (mn >= gt->mn_end[g]) ++g;
      if (g > 0) mn -= gt->mn_end[g - 1];
      nt = gt->n_tiles[g];
    }
    m0 = (mn / nt) * BM;
    n0 = (mn % nt) * BN;
    kb0 = z * p.kb_per;
    nkb = min(p.kb_total, kb0 + p.kb_per) - kb0;
  };
  auto prob_M = [&](int g) { return GROUP ? gt->M[g] : p.M; };
  auto prob_N = [&](int g) { return GROUP ? gt->N[g] : p.N; };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int g = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int g_, m0, n0, kb0, nkb;
        tile_coords(tile, g_, m0, n0, kb0, nkb);
        for (int it = 0; it < nkb; ++it, ++g) {
          const int s = g % STAGES;
          mbar_wait(empty(s), ((g / STAGES) & 1) ^ 1);
          TC_STAMP(0, g);
          const uint32_t st = base + s * STAGE_BYTES;
          const int k = (kb0 + it) * BK;
          mbar_arrive_expect_tx(full_raw(s), 2 * TILE_BYTES);
          if (AK) {
            tma_load_2d(st, tmA + g_, k, m0, full_raw(s));                       // box [32 k][128 rows]
          } else {
#pragma unroll
            for (int q = 0; q < BM / 32; ++q)                                // 4 boxes [32 m][32 k]
              tma_load_2d(st + q * (BK * 128), tmA + g_, m0 + 32 * q, k, full_raw(s));
          }
          if (BKM) {
            tma_load_2d(st + 2 * TILE_BYTES, tmB + g_, k, n0, full_raw(s));
          } else {
#pragma unroll
            for (int q = 0; q < BN / 32; ++q)
              tma_load_2d(st + 2 * TILE_BYTES + q * (BK * 128), tmB + g_, n0 + 32 * q, k, full_raw(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      int g = 0, w = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int g_, m0, n0, kb0, nkb;
        tile_coords(tile, g_, m0, n0, kb0, nkb);
        const int n_valid = min(BN, prob_N(g_) - n0);
        const uint32_t idesc = make_idesc(!AK, !BKM, min(BN, (n_valid + 15) & ~15));
        int in_win = 0;
        for (int it = 0; it < nkb; ++it, ++g) {
          const int s = g % STAGES;
          const int ab = w % ACC_BUFS;
          const uint32_t tacc = tmem_base + (uint32_t)(ab * BN);
          if (in_win == 0) {   // the accumulate warps must have drained this TMEM buffer (ACC_BUFS windows ago)
            mbar_wait(acc_empty(ab), ((w / ACC_BUFS) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          mbar_wait(full_cvt(s), (g / STAGES) & 1);
          TC_STAMP(1, g);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = base + s * STAGE_BYTES;
          const uint32_t a_hi = st, a_lo = st + TILE_BYTES, b_hi = st + 2 * TILE_BYTES, b_lo = st + 3 * TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint32_t ao = AK ? 32u * k : 1024u * k, bo = BKM ? 32u * k : 1024u * k;
            const uint32_t a_lbo = AK ? 16u : p.mn_lbo, b_lbo = BKM ? 16u : p.mn_lbo;
            const uint32_t a_sbo = AK ? 1024u : p.mn_sbo, b_sbo = BKM ? 1024u : p.mn_sbo;
            const uint32_t a_lay = AK ? 2u : p.mn_layout, b_lay = BKM ? 2u : p.mn_layout;
            const uint64_t dah = make_desc(a_hi + ao, a_lbo, a_sbo, a_lay), dal = make_desc(a_lo + ao, a_lbo, a_sbo, a_lay);
            const uint64_t dbh = make_desc(b_hi + bo, b_lbo, b_sbo, b_lay), dbl = make_desc(b_lo + bo, b_lbo, b_sbo, b_lay);
            mma_tf32(tacc, dal, dbh, idesc, (in_win > 0 || k > 0) ? 1u : 0u);   // small terms first
            mma_tf32(tacc, dah, dbl, idesc, 1u);
            mma_tf32(tacc, dah, dbh, idesc, 1u);
          }
          mma_commit(empty(s));               // frees the stage once these MMAs have read it
          ++in_win;
          if (in_win == FLUSH || it == nkb - 1) {
            mma_commit(acc_full(ab));         // window complete: hand the TMEM buffer to the accumulate warps
            ++w;
            in_win = 0;
          }
        }
      }
    }
  } else if (warp < 2 + CVT_THREADS / 32) {
    // ===== hi / lo converters (warps 2..5) =====
    const int ct = threadIdx.x - 64;
    int g = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int g_, m0, n0, kb0, nkb;
      tile_coords(tile, g_, m0, n0, kb0, nkb);
      for (int it = 0; it < nkb; ++it, ++g) {
        const int s = g % STAGES;
        mbar_wait(full_raw(s), (g / STAGES) & 1);
        if (ct == 0) TC_STAMP(2, g);
        uint8_t* st = gen_base + (size_t)s * STAGE_BYTES;
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // A then B
          float4* hi = reinterpret_cast<float4*>(st + half * 2 * TILE_BYTES);
          float4* lo = reinterpret_cast<float4*>(st + half * 2 * TILE_BYTES + TILE_BYTES);
          float4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = hi[ct + i * CVT_THREADS];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 h, l;
            h.x = __uint_as_float(tf32_hi_bits(v[i].x)); l.x = __uint_as_float(tf32_rn_bits(v[i].x - h.x));
            h.y = __uint_as_float(tf32_hi_bits(v[i].y)); l.y = __uint_as_float(tf32_rn_bits(v[i].y - h.y));
            h.z = __uint_as_float(tf32_hi_bits(v[i].z)); l.z = __uint_as_float(tf32_rn_bits(v[i].z - h.z));
            h.w = __uint_as_float(tf32_hi_bits(v[i].w)); l.w = __uint_as_float(tf32_rn_bits(v[i].w - h.w));
#ifdef GSCAN_TC_RN_SPLIT
            hi[ct + i * CVT_THREADS] = h;
#endif
            lo[ct + i * CVT_THREADS] = l;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        mbar_arrive(full_cvt(s));
      }
    }
  } else {
    // ===== accumulate + epilogue (warps 6..9) =====
    // thread = one accumulator row (TMEM lane); a warp may only touch lanes 32*(warp%4)..+31.  The tensor core adds
    // into its fp32 accumulator with truncation, so every FLUSH K-blocks the window sum is folded into registers
    // with round-to-nearest adds (same policy as sgemm_kernel).
    const int q = warp & 3;
    float* stage_out = reinterpret_cast<float*>(gen_base + STAGES * STAGE_BYTES) + (warp - 6) * (32 * 36);
    const bool split = p.ksplit > 1;
    int w = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int g_, m0, n0, kb0, nkb;
      tile_coords(tile, g_, m0, n0, kb0, nkb);
      const int n_valid = min(BN, prob_N(g_) - n0);
      const int nwin = ceil_div(nkb, FLUSH);
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
      for (int i = 0; i < nwin; ++i, ++w) {
        const int ab = w % ACC_BUFS;
        mbar_wait(acc_full(ab), (w / ACC_BUFS) & 1);
        if (threadIdx.x == 192) TC_STAMP(3, w);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(ab * BN);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          if (c * 32 < n_valid) {            // warp-uniform
            uint32_t v[32];
            tmem_ld32(tacc + (uint32_t)(c * 32), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(v[j]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(acc_empty(ab));
      }
      // write-out: per-warp 32x32 transpose through shared memory so that every global access is a full 128-B row segment
      const int mode = split ? 4 : (p.accumulate ? 3 : p.act);     // warp-uniform
      const int rows = min(32, prob_M(g_) - (m0 + 32 * q));
      float* const Cg = GROUP ? gt->C[g_] : p.C;
      const long ldcg = GROUP ? gt->ldc[g_] : p.ldc;
      const int Ng = prob_N(g_);
      const bool vec = TC_VEC_STORE && ((ldcg & 3) == 0) && ((Ng & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cg) & 15) == 0);
      int tl_i = 5 * ((tile - blockIdx.x) / gridDim.x);
      if (threadIdx.x == 192) TC_STAMP(4, tl_i);
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        if (c * 32 < n_valid && rows > 0) {
          __syncwarp();
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(stage_out + lane * 36 + j) =
                  make_float4(acc[c * 32 + j], acc[c * 32 + j + 1], acc[c * 32 + j + 2], acc[c * 32 + j + 3]);
            __syncwarp();
            const int gn = n0 + c * 32 + (lane & 7) * 4;
            float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!split && gn < Ng) {
              if (p.bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + gn)); bsum.x += b.x; bsum.y += b.y; bsum.z += b.z; bsum.w += b.w; }
              if (p.bias2) { const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias2 + gn)); bsum.x += b.x; bsum.y += b.y; bsum.z += b.z; bsum.w += b.w; }
            }
            float* dst = Cg + (long)(m0 + 32 * q) * ldcg + n0 + c * 32;
            const int cols = n_valid - c * 32;
            switch (mode) {
              case 0: store_rows_v4<0>(dst, ldcg, stage_out, lane, rows, cols, bsum); break;
              case 1: store_rows_v4<1>(dst, ldcg, stage_out, lane, rows, cols, bsum); break;
              case 2: store_rows_v4<2>(dst, ldcg, stage_out, lane, rows, cols, bsum); break;
              case 3: store_rows_v4<3>(dst, ldcg, stage_out, lane, rows, cols, bsum); break;
              default: store_rows_v4<4>(dst, ldcg, stage_out, lane, rows, cols, bsum); break;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) stage_out[lane * 33 + j] = acc[c * 32 + j];
            __syncwarp();
            if (threadIdx.x == 192) TC_STAMP(4, 64 + tl_i + 1 + c);
            const int gn = n0 + c * 32 + lane;
            if (gn < Ng) {
              float bsum = 0.f;
              if (!split) {
                if (p.bias) bsum += __ldg(p.bias + gn);
                if (p.bias2) bsum += __ldg(p.bias2 + gn);
              }
              float* dst = Cg + (long)(m0 + 32 * q) * ldcg + gn;
              const float* src = stage_out + lane;
              switch (mode) {
                case 0: store_rows<0>(dst, ldcg, src, rows, bsum); break;
                case 1: store_rows<1>(dst, ldcg, src, rows, bsum); break;
                case 2: store_rows<2>(dst, ldcg, src, rows, bsum); break;
                case 3: store_rows<3>(dst, ldcg, src, rows, bsum); break;
                default: store_rows<4>(dst, ldcg, src, rows, bsum); break;
              }
            }
          }
          if (threadIdx.x == 192) TC_STAMP(4, tl_i + 1 + c);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <bool AK, bool BKM>
__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  tc_body<AK, BKM, false>(&tmA, &tmB, p, nullptr);
}

// "TN" weight-gradient family: C_g[M_g, N_g] += sum_r X_g[r, i] * Y_g[r, j] for every problem g, all over the same R rows
__global__ void __launch_bounds__(THREADS, 1)
tc_group_tn_kernel(const __grid_constant__ GroupMaps maps, const __grid_constant__ GroupTable gt, const Params p) {
  tc_body<false, false, true>(maps.a, maps.b, p, &gt);
}

// zero the destination blocks of a group before its split-K atomics (one launch for all of them)
__global__ void group_zero_kernel(const __grid_constant__ GroupTable gt) {
  const int g = blockIdx.y;
  if (g >= gt.ng) return;
  const int M = gt.M[g], N = gt.N[g];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M * N; i += gridDim.x * blockDim.x)
    gt.C[g][(long)(i / N) * gt.ldc[g] + (i % N)] = 0.f;
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor map: `inner` contiguous elements per row, `outer` rows `ld` floats apart; box [box_inner=32][box_outer]
inline int make_map(CUtensorMap* m, const float* ptr, long inner, long outer, long ld, int box_outer,
                    CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -2;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

inline bool eligible(const float* A, long a_rs, long a_cs, const float* B, long b_rs, long b_cs, int M, int N, int K) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool ak = (a_cs == 1), bk = (b_rs == 1);
  if ((!ak && a_rs != 1) || (!bk && b_cs != 1)) return false;
  const long lda = ak ? a_rs : a_cs, ldb = bk ? b_cs : b_rs;
  if (!al16(A) || !al16(B) || (lda & 3) || (ldb & 3)) return false;
  if (M < 64 || N < 32 || K < 32) return false;     // tiny products stay on the mma.sync kernel
  return encode_fn() != nullptr;
}

inline int num_sms() {
  static int per_device[16] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
  int& n = per_device[dev];
  if (!n && (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)) n = 148;
  return n;
}

// SM budget of the persistent kernels.  One CTA of this kernel takes a whole SM (197 KB of shared memory), and a
// persistent wave never gives it back: while a GEMM holds all SMs, the small kernels of concurrent streams wait for
// its end.  Callers that have latency-critical chains running beside a GEMM cap its grid for the scope of the launch.
inline int& sm_cap() { static thread_local int cap = 0; return cap; }
struct ScopedSmCap {
  int prev;
  explicit ScopedSmCap(int cap) : prev(sm_cap()) { sm_cap() = cap; }
  ~ScopedSmCap() { sm_cap() = prev; }
};
inline int usable_sms() {
  const int n = num_sms(), cap = sm_cap();
  return (cap > 0 && cap < n) ? cap : n;
}

template <bool AK, bool BKM>
inline int launch_t(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, dim3 grid, cudaStream_t st) {
  static bool configured_d[16] = {};   // per device: the attribute is a per-device setting
  int dev_ = 0;
  if (cudaGetDevice(&dev_) != cudaSuccess || dev_ < 0 || dev_ >= 16) return GSCAN_E_UNSUPPORTED;
  bool& configured = configured_d[dev_];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<AK, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  tc_gemm_kernel<AK, BKM><<<grid, THREADS, SMEM_BYTES, st>>>(ta, tb, p);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

// Same argument meaning as launch_sgemm (gemm.cuh).  ksplit > 1 requires C initialised and forbids bias / act.
inline int launch(const float* A, long a_rs, long a_cs, const float* B, long b_rs, long b_cs, float* C, long ldc,
                  int M, int N, int K, const float* bias, const float* bias2, int act, int accumulate, int ksplit,
                  cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  const bool ak = (a_cs == 1), bk = (b_rs == 1);
  const long lda = ak ? a_rs : a_cs, ldb = bk ? b_cs : b_rs;
  CUtensorMap ta, tb;
  const MnConfig& mc = mn_config();
  const CUtensorMapSwizzle mn_swz = (CUtensorMapSwizzle)mc.tma_swizzle;
  int rc = ak ? make_map(&ta, A, K, M, lda, BM) : make_map(&ta, A, M, K, lda, BK, mn_swz);
  if (rc) return rc;
  rc = bk ? make_map(&tb, B, K, N, ldb, BN) : make_map(&tb, B, N, K, ldb, BK, mn_swz);
  if (rc) return rc;
  Params p{C, ldc, M, N, K, bias, bias2, act, accumulate, 0, 0, 1, mc.layout, mc.sbo, mc.lbo, mc.timeline};
  p.kb_total = ceil_div(K, BK);
  if (ksplit < 1) ksplit = 1;
  p.kb_per = ceil_div(p.kb_total, ksplit);
  p.ksplit = ceil_div(p.kb_total, p.kb_per);
  const int tiles = ceil_div(N, BN) * ceil_div(M, BM) * p.ksplit;
  dim3 grid(min(tiles, usable_sms()));
  if (ak && bk) return launch_t<true, true>(ta, tb, p, grid, st);
  if (ak && !bk) return launch_t<true, false>(ta, tb, p, grid, st);
  if (!ak && bk) return launch_t<false, true>(ta, tb, p, grid, st);
  return launch_t<false, false>(ta, tb, p, grid, st);
}

// One weight-gradient problem of a group: C[N1, N2] (ldc) = sum_r X[r, i] * Y[r, j], X [R][ldx], Y [R][ldy].
struct GroupProblem { const float* X; long ldx; const float* Y; long ldy; float* C; long ldc; int N1, N2; };

inline bool group_eligible(const GroupProblem& q, int R) { return eligible(q.X, 1, q.ldx, q.Y, q.ldy, 1, q.N1, q.N2, R); }

// All problems in ONE persistent launch (plus one launch that zeroes the destinations).  The tensor maps are
// rebuilt only when a pointer or shape changed since the previous call (the workspace is stable across steps).
inline int launch_group_zero(const GroupProblem* probs, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  if (n > MAXG) return -3;
  GroupTable gt{};
  for (int g = 0; g < n; ++g) { gt.C[g] = probs[g].C; gt.ldc[g] = probs[g].ldc; gt.M[g] = probs[g].N1; gt.N[g] = probs[g].N2; }
  gt.ng = n;
  group_zero_kernel<<<dim3(16, n), 256, 0, st>>>(gt);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

// zero_first: zero the destinations here (else the caller did, e.g. before two partial launches over row ranges);
// ksplit_hint > 0 fixes the number of K-splits (default: one persistent wave over the SM budget).
inline int launch_group_tn(const GroupProblem* probs, int n, int R, cudaStream_t st, bool zero_first = true,
                           int ksplit_hint = 0) {
  if (n <= 0) return 0;
  if (n > MAXG) return -3;
  struct Entry { GroupProblem key[MAXG]; int n = 0, R = 0; GroupMaps maps; };
  struct Cache { Entry e[8]; int next = 0; };
  static thread_local Cache store;
  const MnConfig& mc = mn_config();
  const CUtensorMapSwizzle mn_swz = (CUtensorMapSwizzle)mc.tma_swizzle;
  Entry* hit = nullptr;
  for (auto& e : store.e) {
    bool same = e.n == n && e.R == R;
    for (int g = 0; same && g < n; ++g) same = memcmp(&e.key[g], &probs[g], sizeof(GroupProblem)) == 0;
    if (same) { hit = &e; break; }
  }
  if (!hit) {
    Entry& e = store.e[store.next];
    store.next = (store.next + 1) % 8;
    e.n = 0;
    for (int g = 0; g < n; ++g) {
      int rc = make_map(&e.maps.a[g], probs[g].X, probs[g].N1, R, probs[g].ldx, BK, mn_swz);
      if (rc) return rc;
      rc = make_map(&e.maps.b[g], probs[g].Y, probs[g].N2, R, probs[g].ldy, BK, mn_swz);
      if (rc) return rc;
      memset(&e.key[g], 0, sizeof(GroupProblem));
      e.key[g] = probs[g];
    }
    e.n = n;
    e.R = R;
    hit = &e;
  }
  Entry& cache = *hit;
  GroupTable gt{};
  int mn = 0;
  for (int g = 0; g < n; ++g) {
    gt.C[g] = probs[g].C; gt.ldc[g] = probs[g].ldc; gt.M[g] = probs[g].N1; gt.N[g] = probs[g].N2;
    gt.n_tiles[g] = ceil_div(probs[g].N2, BN);
    mn += ceil_div(probs[g].N1, BM) * gt.n_tiles[g];
    gt.mn_end[g] = mn;
  }
  gt.ng = n;
  Params p{nullptr, 0, 0, 0, R, nullptr, nullptr, 0, 0, 0, 0, 1, mc.layout, mc.sbo, mc.lbo, nullptr};
  p.kb_total = ceil_div(R, BK);
  // one persistent wave over the SM budget of the caller
  const int sms = ksplit_hint > 0 ? usable_sms() : max(usable_sms(), mn);
  int ksplit = max(1, min(ceil_div(R, 4 * BK), sms / mn));
  if (ksplit_hint > 0) ksplit = min(ksplit_hint, ceil_div(R, 4 * BK));
  p.kb_per = ceil_div(p.kb_total, ksplit);
  p.ksplit = ceil_div(p.kb_total, p.kb_per);
  if (!zero_first && p.ksplit == 1 && p.kb_total >= 2) {   // partial launches may overlap in time: atomics only
    p.kb_per = ceil_div(p.kb_total, 2);
    p.ksplit = ceil_div(p.kb_total, p.kb_per);
  }
  if (zero_first) {
    group_zero_kernel<<<dim3(16, n), 256, 0, st>>>(gt);
    GSCAN_CHECK_LAUNCH();
  } else if (p.ksplit == 1) {
    p.accumulate = 1;   // destinations hold the other partial sums
  }
  static bool configured_d[16] = {};   // per device: the attribute is a per-device setting
  int dev_ = 0;
  if (cudaGetDevice(&dev_) != cudaSuccess || dev_ < 0 || dev_ >= 16) return GSCAN_E_UNSUPPORTED;
  bool& configured = configured_d[dev_];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_group_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  tc_group_tn_kernel<<<min(mn * p.ksplit, sms), THREADS, SMEM_BYTES, st>>>(cache.maps, gt, p);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

}  // namespace tc
}  // namespace gscan
