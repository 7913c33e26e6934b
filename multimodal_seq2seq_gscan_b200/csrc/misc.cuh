// Small kernels around the sweeps: embedding gather/scatter, output log-softmax, auxiliary
// head, loss / metrics, fused Adam, weight re-layout.
#pragma once
#include "common.cuh"

namespace gscan {

// ---------------------------------------------------------------------------------------------
// Weight packing: dst[k*ldd + r0 + r] = src[r*lds + c0 + k]   (transpose of a column block)
// ---------------------------------------------------------------------------------------------
struct PackDesc { const float* src; int lds, c0; float* dst; int ldd, r0, R, K; };
constexpr int kMaxPack = 12;
struct PackTable { PackDesc d[kMaxPack]; int n; };

__global__ void pack_transpose_kernel(PackTable tab) {
  __shared__ float tile[32][33];
  const PackDesc d = tab.d[blockIdx.z];
  const int r_base = blockIdx.y * 32, k_base = blockIdx.x * 32;
  if (r_base >= d.R || k_base >= d.K) return;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r_base + i, k = k_base + threadIdx.x;
    tile[i][threadIdx.x] = (r < d.R && k < d.K) ? __ldg(d.src + (long)r * d.lds + d.c0 + k) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int k = k_base + i, r = r_base + threadIdx.x;
    if (r < d.R && k < d.K) d.dst[(long)k * d.ldd + d.r0 + r] = tile[threadIdx.x][i];
  }
}

inline int launch_pack(const PackTable& tab, cudaStream_t st) {
  int maxR = 0, maxK = 0;
  for (int i = 0; i < tab.n; ++i) { maxR = max(maxR, tab.d[i].R); maxK = max(maxK, tab.d[i].K); }
  dim3 grid(ceil_div(maxK, 32), ceil_div(maxR, 32), tab.n);
  pack_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(tab);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Embedding gather: out[row*ldo + h] = W[tok][h] * mask[mrow*Wd + h]
//   row_mode 0 (encoder): rows r = b*T + t, tok = tokens[b*tok_stride + t], out row = r, mask row = r
//   row_mode 1 (decoder): tok = tokens[b*tok_stride + t], out row = (t + 1)*B + b (time-major U
//                         rows, group 0 reserved), mask row = b*T + t
// ---------------------------------------------------------------------------------------------
__global__ void embed_kernel(const long long* __restrict__ tokens, int tok_stride, const float* __restrict__ W,
                             int Wd, DropSrc mask, float* __restrict__ out, long ldo,
                             int B, int T, int row_mode) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * T * Wd) return;
  int h = idx % Wd;
  long r = idx / Wd;
  int b = r / T, t = r - (long)b * T;
  long long tok = tokens[(long)b * tok_stride + t];
  float v = __ldg(W + tok * Wd + h);
  if (mask.active()) v *= drop_at(mask, r * Wd + h);
  long orow = row_mode == 0 ? r : ((long)(t + 1) * B + b);
  out[orow * ldo + h] = v;
}

// 4 elements per thread (Wd % 4 == 0, 16-byte aligned rows): a quarter of the threads and 128-bit accesses
__global__ void embed4_kernel(const long long* __restrict__ tokens, int tok_stride, const float* __restrict__ W,
                              int Wd, DropSrc mask, float* __restrict__ out, long ldo,
                              int B, int T, int row_mode) {
  const int Wq = Wd >> 2;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * T * Wq) return;
  int hq = idx % Wq;
  long r = idx / Wq;
  int b = r / T, t = r - (long)b * T;
  long long tok = tokens[(long)b * tok_stride + t];
  float4 v = __ldg(reinterpret_cast<const float4*>(W + tok * Wd) + hq);
  if (mask.active()) {
    const float4 m = drop_at4(mask, r * Wd + 4 * hq);
    v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
  }
  long orow = row_mode == 0 ? r : ((long)(t + 1) * B + b);
  reinterpret_cast<float4*>(out + orow * ldo)[hq] = v;
}

// Embedding scatter-add: dW[tok][h] += mask * dX[row][h] for tok != pad.  dW zeroed by host.
// Each CTA walks a chunk of rows.  use_smem: thread (g, h) = (tid / Wd, tid % Wd) owns column h of a private copy g of
// the table in shared memory ([groups][Vsz][Wd] floats) and walks the rows r0 + g, r0 + g + groups, ... with plain
// read-modify-writes - no atomics inside the CTA (the first version issued one shared-memory atomic per element and
// took 28 us for the decoder table); the copies are summed at the end and added to dW with one atomic per entry.
__host__ __device__ inline int embed_bwd_groups(int Wd, int nthreads) { return Wd <= nthreads ? nthreads / Wd : 0; }

__global__ void __launch_bounds__(256) embed_bwd_kernel(const long long* __restrict__ tokens, int tok_stride,
                                                        const float* __restrict__ dX, long ldx,
                                                        DropSrc mask, float* __restrict__ dW,
                                                        int Wd, int Vsz, int pad, int B, int T, int row_mode,
                                                        int rows_per_block, int use_smem) {
  extern __shared__ __align__(16) float acc_s[];
  const long R = (long)B * T;
  long r0 = (long)blockIdx.x * rows_per_block;
  long r1 = min(R, r0 + rows_per_block);
  if (use_smem) {
    const int groups = embed_bwd_groups(Wd, blockDim.x);
    for (int i = threadIdx.x; i < groups * Vsz * Wd; i += blockDim.x) acc_s[i] = 0.f;
    __syncthreads();
    const int g = threadIdx.x / Wd, h = threadIdx.x - g * Wd;
    if (g < groups) {
      float* mine = acc_s + (size_t)g * Vsz * Wd + h;
      constexpr int kU = 8;   // rows in flight per thread
      for (long rb = r0 + g; rb < r1; rb += (long)kU * groups) {
        long long tok[kU];
        float x[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const long r = rb + (long)u * groups;
          tok[u] = pad;
          x[u] = 0.f;
          if (r < r1) {
            const int b = r / T, t = r - (long)b * T;
            tok[u] = tokens[(long)b * tok_stride + t];
            if (tok[u] != pad) {
              const long xrow = row_mode == 0 ? r : ((long)t * B + b);
              x[u] = __ldg(dX + xrow * ldx + h);
              if (mask.active()) x[u] *= drop_at(mask, r * Wd + h);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (tok[u] != pad) mine[tok[u] * Wd] += x[u];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Vsz * Wd; i += blockDim.x) {
      float v = 0.f;
      for (int k = 0; k < groups; ++k) v += acc_s[(size_t)k * Vsz * Wd + i];
      if (v != 0.f) atomicAdd(dW + i, v);
    }
    return;
  }
  // table too large for shared memory: one warp per row, global atomics
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (long r = r0 + warp; r < r1; r += nwarps) {
    int b = r / T, t = r - (long)b * T;
    long long tok = tokens[(long)b * tok_stride + t];
    if (tok == pad) continue;
    long xrow = row_mode == 0 ? r : ((long)t * B + b);
    for (int h = lane; h < Wd; h += 32) {
      float g = __ldg(dX + xrow * ldx + h);
      if (mask.active()) g *= drop_at(mask, r * Wd + h);
      atomicAdd(dW + tok * Wd + h, g);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// logits = pre . W_h2o^T ; logp = log_softmax(logits)   (seq2seq_model.py:424, model.py:203)
// pre rows are time-major (row = t*B + b); logp is written batch-major [B][T][V].
// One warp per row, lane v owns vocabulary entry v (V <= 32*4).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxVPerLane = 4;

__global__ void __launch_bounds__(1024) out_logsoftmax_kernel(const float* __restrict__ pre, const float* __restrict__ Wh2o,
                                                             int H, int V, int B, int T, float* __restrict__ logp,
                                                             float* __restrict__ logits_out /* [R][V] or null */,
                                                             long row_begin = 0, long row_end = -1,
                                                             float* __restrict__ logp2 = nullptr /* second copy of logp */) {
  extern __shared__ __align__(16) float w_s[];   // [V][H+1]
  for (int i = threadIdx.x; i < V * H; i += blockDim.x) w_s[(i / H) * (H + 1) + (i % H)] = __ldg(Wh2o + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long R = row_end >= 0 ? row_end : (long)B * T;   // rows [row_begin, R) of the time-major list
  for (long row = row_begin + (long)blockIdx.x * (blockDim.x >> 5) + warp; row < R; row += (long)gridDim.x * (blockDim.x >> 5)) {
    const float* x = pre + row * H;
    float l[kMaxVPerLane];
#pragma unroll
    for (int i = 0; i < kMaxVPerLane; ++i) l[i] = -INFINITY;
    if (H <= 128) {
      // lanes split the hidden axis (coalesced row load, 4 values per lane); one butterfly sum per vocabulary entry.
      // (One lane per entry walking all H terms left 23 of 32 lanes idle for V = 9: 34 us for the 24,200 rows.)
      float xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = (lane + 32 * i < H) ? __ldg(x + lane + 32 * i) : 0.f;
      for (int v = 0; v < V; ++v) {
        const float* w = w_s + v * (H + 1);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (lane + 32 * i < H) s = fmaf(xv[i], w[lane + 32 * i], s);
        s = warp_sum(s);
        if ((v & 31) == lane) {
#pragma unroll
          for (int i = 0; i < kMaxVPerLane; ++i)
            if (i == (v >> 5)) l[i] = s;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kMaxVPerLane; ++i) {
        int v = lane + 32 * i;
        if (v < V) {
          const float* w = w_s + v * (H + 1);
          float s = 0.f;
          for (int h = 0; h < H; ++h) s = fmaf(__ldg(x + h), w[h], s);
          l[i] = s;
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMaxVPerLane; ++i) mx = fmaxf(mx, l[i]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVPerLane; ++i)
      if (lane + 32 * i < V) sum += expf(l[i] - mx);
    sum = warp_sum(sum);
    const float lse = logf(sum);
    int t = row / B, b = row - (long)t * B;
#pragma unroll
    for (int i = 0; i < kMaxVPerLane; ++i) {
      int v = lane + 32 * i;
      if (v < V) {
        if (logp) logp[((long)b * T + t) * V + v] = (l[i] - mx) - lse;
        if (logp2) logp2[((long)b * T + t) * V + v] = (l[i] - mx) - lse;
        if (logits_out) logits_out[row * V + v] = l[i];
      }
    }
  }
}

// Output head backward in one pass over the rows (V <= kHeadMaxV, H <= 128), replacing logsoftmax_bwd_kernel + the
// [R x V] . [V x H] product + the [V x R] . [R x H] weight-gradient product:
//   dlogits[row] = dlogp - exp(logp) * sum_v dlogp           (lane v holds entry v)
//   dpre[row]    = dlogits[row] . W_h2o                       (lanes split H; dlogits broadcast by shuffles)
//   dW_h2o      += dlogits[row]^T (x) pre[row]                (per-warp register accumulators, one atomic per CTA entry)
// dW must be zeroed by the host.  dlogits is not materialised.
constexpr int kHeadMaxV = 16;
__global__ void __launch_bounds__(256) head_bwd_fused_kernel(const float* __restrict__ dlogp, const float* __restrict__ logp,
                                                             const float* __restrict__ pre, const float* __restrict__ Wh2o,
                                                             int H, int V, int B, int T, float* __restrict__ dpre,
                                                             float* __restrict__ dW, long row_begin = 0, long row_end = -1) {
  extern __shared__ __align__(16) float hb_s[];   // W [V][H] then dW accumulator [V][H]
  float* w_s = hb_s;
  float* acc_s = hb_s + V * H;
  for (int i = threadIdx.x; i < V * H; i += blockDim.x) { w_s[i] = __ldg(Wh2o + i); acc_s[i] = 0.f; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long R = row_end >= 0 ? row_end : (long)B * T;   // rows [row_begin, R) of the time-major list
  float acc[kHeadMaxV][4];
#pragma unroll
  for (int v = 0; v < kHeadMaxV; ++v)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[v][i] = 0.f;
  for (long row = row_begin + (long)blockIdx.x * (blockDim.x >> 5) + warp; row < R; row += (long)gridDim.x * (blockDim.x >> 5)) {
    const int t = row / B, b = row - (long)t * B;
    const long off = ((long)b * T + t) * V;
    const float dl = lane < V ? __ldg(dlogp + off + lane) : 0.f;
    const float lp = lane < V ? __ldg(logp + off + lane) : 0.f;
    float xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = (lane + 32 * i < H) ? __ldg(pre + row * H + lane + 32 * i) : 0.f;
    const float s = warp_sum(dl);
    const float dlog = lane < V ? dl - expf(lp) * s : 0.f;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int v = 0; v < kHeadMaxV; ++v) {
      if (v < V) {
        const float dv = __shfl_sync(0xffffffffu, dlog, v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (lane + 32 * i < H) d[i] = fmaf(dv, w_s[v * H + lane + 32 * i], d[i]);
          acc[v][i] = fmaf(dv, xv[i], acc[v][i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (lane + 32 * i < H) dpre[row * H + lane + 32 * i] = d[i];
  }
#pragma unroll
  for (int v = 0; v < kHeadMaxV; ++v)
    if (v < V)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (lane + 32 * i < H) atomicAdd(acc_s + v * H + lane + 32 * i, acc[v][i]);
  __syncthreads();
  for (int i = threadIdx.x; i < V * H; i += blockDim.x) atomicAdd(dW + i, acc_s[i]);
}

// dlogits[row][v] = dlogp[b][t][v] - exp(logp[b][t][v]) * sum_v dlogp[b][t][v]; row = t*B + b
__global__ void __launch_bounds__(256) logsoftmax_bwd_kernel(const float* __restrict__ dlogp, const float* __restrict__ logp,
                                                             int V, int B, int T, float* __restrict__ dlogits) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long R = (long)B * T;
  for (long row = (long)blockIdx.x * (blockDim.x >> 5) + warp; row < R; row += (long)gridDim.x * (blockDim.x >> 5)) {
    int t = row / B, b = row - (long)t * B;
    const float* dl = dlogp + ((long)b * T + t) * V;
    const float* lp = logp + ((long)b * T + t) * V;
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += __ldg(dl + v);
    s = warp_sum(s);
    for (int v = lane; v < V; v += 32) dlogits[row * V + v] = __ldg(dl + v) - expf(__ldg(lp + v)) * s;
  }
}

// aux head: out[b] = log_softmax(x[b])  (model.py:166-170); one warp per row
__global__ void __launch_bounds__(256) row_logsoftmax_kernel(const float* __restrict__ x, int N, int R,
                                                             float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long row = (long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= R) return;
  const float* xr = x + row * N;
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) mx = fmaxf(mx, xr[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < N; j += 32) sum += expf(xr[j] - mx);
  sum = warp_sum(sum);
  float lse = logf(sum);
  for (int j = lane; j < N; j += 32) out[row * N + j] = (xr[j] - mx) - lse;
}

// d x[b] = dy[b] - exp(y[b]) * sum(dy[b])
__global__ void __launch_bounds__(256) row_logsoftmax_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                 int N, int R, float* __restrict__ dx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long row = (long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= R) return;
  float s = 0.f;
  for (int j = lane; j < N; j += 32) s += dy[row * N + j];
  s = warp_sum(s);
  for (int j = lane; j < N; j += 32) dx[row * N + j] = dy[row * N + j] - expf(y[row * N + j]) * s;
}

// ---------------------------------------------------------------------------------------------
// Loss / metrics (model.py:108-160): targets shifted left by one, pad ignored.  kNllBlocks CTAs write partial
// (sum, count) pairs, one warp adds them in a fixed order => bitwise reproducible (round 1: a single CTA walked all
// B * Tt rows, 16 us on the path between the forward and the backward pass).
// ---------------------------------------------------------------------------------------------
constexpr int kNllBlocks = 32;
__global__ void __launch_bounds__(256) nll_forward_kernel(const float* __restrict__ logp, const long long* __restrict__ tgt,
                                                          int B, int T, int V, int pad, int shift,
                                                          float* __restrict__ partial /* [gridDim.x][2] */) {
  __shared__ float s_sum[32];
  __shared__ float s_cnt[32];
  float sum = 0.f, cnt = 0.f;
  const long R = (long)B * T;
  constexpr int kU = 4;   // rows in flight per thread: all targets first, then all log-probs (two dependent latencies per pass)
  for (long rb = (long)blockIdx.x * blockDim.x + threadIdx.x; rb < R; rb += (long)kU * blockDim.x * gridDim.x) {
    long long y[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long r = rb + (long)u * blockDim.x * gridDim.x;
      y[u] = pad;
      if (r < R) {
        int b = r / T, t = r - (long)b * T;
        // beyond the shifted sequence the reference appends a literal 0 (remove_start_of_sequence, model.py:108-115):
        // ignored when pad == 0 (the gSCAN vocabularies), scored as class 0 otherwise - as NLLLoss(ignore_index) does
        y[u] = (t + shift < T) ? tgt[(long)b * T + t + shift] : (shift > 0 ? 0 : pad);
      }
    }
    float lp[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long r = rb + (long)u * blockDim.x * gridDim.x;
      lp[u] = y[u] != pad ? __ldg(logp + r * V + y[u]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (y[u] != pad) { sum -= lp[u]; cnt += 1.f; }
  }
  sum = warp_sum(sum); cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x < 32) {
    int nw = blockDim.x >> 5;
    sum = threadIdx.x < nw ? s_sum[threadIdx.x] : 0.f;
    cnt = threadIdx.x < nw ? s_cnt[threadIdx.x] : 0.f;
    sum = warp_sum(sum); cnt = warp_sum(cnt);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sum; partial[2 * blockIdx.x + 1] = cnt; }
  }
}
// out[0] = mean, out[1] = count from the partial pairs behind them (out + 2): one warp, fixed order
__global__ void nll_final_kernel(float* __restrict__ out, int nblocks) {
  const int lane = threadIdx.x;
  float sum = lane < nblocks ? out[2 + 2 * lane] : 0.f, cnt = lane < nblocks ? out[3 + 2 * lane] : 0.f;
  sum = warp_sum(sum); cnt = warp_sum(cnt);
  if (lane == 0) { out[0] = sum / cnt; out[1] = cnt; }
}

// out[1] = number of scored entries (what nll_forward_kernel counts), out[0] = 0: the targets alone decide it, so the
// gradient of the loss with respect to the log-probabilities can be formed before the forward pass has run
__global__ void __launch_bounds__(1024) nll_count_kernel(const long long* __restrict__ tgt, int B, int T, int pad, int shift,
                                                         float* __restrict__ out) {
  __shared__ float s_cnt[32];
  float cnt = 0.f;
  const long R = (long)B * T;
  for (long r = threadIdx.x; r < R; r += blockDim.x) {
    const int b = r / T, t = r - (long)b * T;
    const long long y = (t + shift < T) ? tgt[(long)b * T + t + shift] : (shift > 0 ? 0 : pad);   // see nll_forward_kernel
    cnt += y != pad ? 1.f : 0.f;
  }
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x < 32) {
    float c = threadIdx.x < (blockDim.x >> 5) ? s_cnt[threadIdx.x] : 0.f;
    c = warp_sum(c);   // integers below 2^24: exact in any order
    if (threadIdx.x == 0) { out[0] = 0.f; out[1] = c; }
  }
}

__global__ void nll_backward_kernel(const long long* __restrict__ tgt, int B, int T, int V, int pad, int shift,
                                    const float* __restrict__ loss_out, const float* __restrict__ d_loss,
                                    float* __restrict__ d_logp) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * T * V) return;
  int v = idx % V;
  long r = idx / V;
  int b = r / T, t = r - (long)b * T;
  float g = 0.f;
  {
    const long long y = (t + shift < T) ? tgt[(long)b * T + t + shift] : (shift > 0 ? 0 : pad);   // see nll_forward_kernel
    if (y != pad && y == v) g = -__ldg(d_loss) / __ldg(loss_out + 1);
  }
  d_logp[idx] = g;
}

// counts[0] = matching non-pad tokens, [1] = non-pad tokens, [2] = exactly matching sequences
__global__ void __launch_bounds__(256) metrics_kernel(const float* __restrict__ logp, const long long* __restrict__ tgt,
                                                      int B, int T, int V, int pad, int* __restrict__ counts) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= B) return;
  int match = 0, total = 0;
  for (int t = lane; t < T; t += 32) {
    long long y = (t + 1 < T) ? tgt[(long)b * T + t + 1] : 0;   // the appended literal 0 (model.py:108-115,121)
    if (y == pad) continue;
    const float* lp = logp + ((long)b * T + t) * V;
    int arg = 0;
    float best = lp[0];
    for (int v = 1; v < V; ++v)
      if (lp[v] > best) { best = lp[v]; arg = v; }   // first maximum, as tensor.max(dim)[1]
    total++;
    match += (arg == y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    match += __shfl_xor_sync(0xffffffffu, match, o);
    total += __shfl_xor_sync(0xffffffffu, total, o);
  }
  if (lane == 0) {
    atomicAdd(counts + 0, match);
    atomicAdd(counts + 1, total);
    if (match == total) atomicAdd(counts + 2, 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused Adam (torch.optim.Adam semantics, train.py:67-70,110-113)
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt, float gscale, const float* __restrict__ gdenom) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (gdenom) gscale = gscale / __ldg(gdenom);   // summed token count of the data-parallel all-reduce
  float gi = g[i] * gscale;
  float mi = b1 * m[i] + (1.f - b1) * gi;
  float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// the mask of a dropout site, materialised (parity tests of the in-kernel draw)
__global__ void dropout_mask_kernel(DropSrc src, float* __restrict__ out, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = drop_at(src, i);
}

// d_pre = d_out * (1 - out^2)
__global__ void tanh_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dpre,
                                long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dpre[i] = dout[i] * (1.f - out[i] * out[i]);
}
inline int launch_tanh_bwd(const float* dout, const float* out, float* dpre, long n, cudaStream_t st) {
  tanh_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dout, out, dpre, n);
  GSCAN_CHECK_LAUNCH();
  return 0;
}

__global__ void fill_i64_kernel(long long* p, long n, long long v) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace gscan
