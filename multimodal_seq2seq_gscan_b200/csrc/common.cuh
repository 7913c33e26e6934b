// Shared device helpers for the gSCAN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

// every kernel launch of the library goes through this macro, which also counts launches
// (gscan_launch_count) so that callers can state how many of OUR kernels ran in a timed region
namespace gscan { inline unsigned long long& launch_counter() { static unsigned long long c = 0; return c; } }
#define GSCAN_CHECK_LAUNCH()                         \
  do {                                               \
    ++gscan::launch_counter();                       \
    cudaError_t e__ = cudaGetLastError();            \
    if (e__ != cudaSuccess) return (int)e__;         \
  } while (0)

namespace gscan {

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// maximum over the warp in ONE instruction (redux.sync.max.f32 -> CREDUX.MAX.F32, new on sm_100a) instead of five
// shuffle + max pairs; -inf is handled, the softmaxes never see NaN
__device__ __forceinline__ float warp_max_redux(float v) {
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}

// tanh / sigmoid with ~1e-7 absolute error: one ex2.approx + one rcp.approx on the SFU instead
// of libm's ~25-instruction tanhf.  The attention scores need (Ti + G*G) * H tanh per example per
// decoder step, which makes the SFU the second-busiest pipe of the recurrent sweep.
#ifdef GSCAN_PRECISE_MATH
__device__ __forceinline__ float act_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ float act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
#else
__device__ __forceinline__ float act_tanh(float x) {
  // 1 - 2/(1+e^{2x}); saturates correctly: e -> inf gives 1, e -> 0 gives -1
  float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}
__device__ __forceinline__ float act_sigmoid(float x) {
  return __fdividef(1.0f, 1.0f + __expf(-x));
}
#endif

// N tanh evaluations sharing ONE reciprocal: tanh(x_i) = 1 - 2 / d_i with d_i = 1 + e^{2 x_i}, and
// 1/d_i = (prod_{j != i} d_j) / (prod_j d_j).  N ex2 + 1 rcp on the SFU instead of 2N, the products go to the
// FMA pipe.  The clamp keeps prod_j d_j finite for N <= 5 (d <= e^17.2) and costs nothing: tanh(8.6) rounds to 1.
#ifdef GSCAN_PRECISE_MATH
template <int N>
__device__ __forceinline__ void act_tanh_n(const float (&x)[N], float (&y)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = tanhf(x[i]);
}
#else
template <int N>
__device__ __forceinline__ void act_tanh_n(const float (&x)[N], float (&y)[N]) {
  static_assert(N >= 2 && N <= 5, "product of N denominators must stay below FLT_MAX");
  float d[N], pre[N], suf[N];
#pragma unroll
  for (int i = 0; i < N; ++i) d[i] = 1.0f + __expf(2.0f * fminf(fmaxf(x[i], -8.6f), 8.6f));
  pre[0] = 1.0f;
#pragma unroll
  for (int i = 1; i < N; ++i) pre[i] = pre[i - 1] * d[i - 1];      // prod_{j < i} d_j
  suf[N - 1] = 1.0f;
#pragma unroll
  for (int i = N - 2; i >= 0; --i) suf[i] = suf[i + 1] * d[i + 1];  // prod_{j > i} d_j
  const float r = __fdividef(2.0f, pre[N - 1] * d[N - 1]);          // 2 / prod_j d_j
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = fmaf(-r, pre[i] * suf[i], 1.0f);
}
#endif

// ---- dropout masks: explicit (parity tests, the oracle's masks) or drawn in the consuming kernel -------------------
// Round 1 drew the three masks of a step with three ATen kernels before the forward pass and read them back in five
// kernels (25 MB per step through HBM).  A DropSrc names the mask of one dropout site instead: either a pointer to an
// explicit, already scaled mask, or a Philox4x32-10 stream (key = seed, counter = element index / 4, site, call
// number): the forward and the backward kernels of a step regenerate the same bits from the same counter.
struct DropSrc {
  const float* mask = nullptr;
  unsigned int k0 = 0, k1 = 0, c2 = 0, c3 = 0;
  unsigned int thresh = 0;   // keep iff the 32 random bits are >= thresh (= p * 2^32)
  float scale = 1.f;         // 1 / (1 - p)
  int rng = 0;
  __host__ __device__ DropSrc() {}
  __host__ __device__ DropSrc(const float* m) : mask(m) {}
  __host__ __device__ bool active() const { return mask != nullptr || rng != 0; }
};
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int h0 = __umulhi(0xD2511F53u, c.x), l0 = 0xD2511F53u * c.x;
    const unsigned int h1 = __umulhi(0xCD9E8D57u, c.z), l1 = 0xCD9E8D57u * c.z;
    c = make_uint4(h1 ^ c.y ^ k.x, l1, h0 ^ c.w ^ k.y, l0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// mask values of the four elements i .. i + 3 (i a multiple of 4)
__device__ __forceinline__ float4 drop_at4(const DropSrc& d, long i) {
  if (d.mask) return __ldg(reinterpret_cast<const float4*>(d.mask + i));
  const uint4 r = philox4x32_10(make_uint4((unsigned int)(i >> 2), (unsigned int)(i >> 34), d.c2, d.c3), make_uint2(d.k0, d.k1));
  return make_float4(r.x >= d.thresh ? d.scale : 0.f, r.y >= d.thresh ? d.scale : 0.f, r.z >= d.thresh ? d.scale : 0.f,
                     r.w >= d.thresh ? d.scale : 0.f);
}
__device__ __forceinline__ float drop_at(const DropSrc& d, long i) {
  if (d.mask) return __ldg(d.mask + i);
  const uint4 r = philox4x32_10(make_uint4((unsigned int)(i >> 2), (unsigned int)(i >> 34), d.c2, d.c3), make_uint2(d.k0, d.k1));
  const int j = (int)(i & 3);
  const unsigned int v = j == 0 ? r.x : (j == 1 ? r.y : (j == 2 ? r.z : r.w));
  return v >= d.thresh ? d.scale : 0.f;
}

// Packed fp32 FMA (FFMA2, new on sm_100): two independent IEEE fp32 FMAs per issue slot.
__device__ __forceinline__ void fma2(float2& acc, const float2 a, const float2 b) {
  acc = __ffma2_rn(a, b, acc);
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace gscan
