// nms.cu -- batched hard NMS over score-sorted segments (tf.image.non_max_suppression
// semantics: IoU in the TF CPU kernel's division form, strict '>', greedy in
// (score desc, index asc) order, stop at max_output_size).
//
// Two exact formulations, chosen per call:
//  A. bitmask (RPN-sized segments, large cap): all (segment, 64-row block) CTAs
//     compute the upper-triangular IoU>thr matrix as 64-bit words in parallel
//     (phase 1, spread over every SM); one warp per segment then sweeps 64 rows at
//     a time: the diagonal 64x64 tile is resolved serially from shared memory,
//     the surviving rows are OR-ed into a per-lane "removed" word set, and the
//     sweep stops at the cap (phase 2).
//  B. capped lazy sweep (huge candidate lists, small cap: Fast R-CNN / RetinaNet):
//     one CTA per segment keeps the selected boxes in shared memory and, per
//     64-candidate block, tests candidates against the kept list, resolves the
//     diagonal tile with ballots, and exits once `cap` boxes are kept.  Only
//     O(n * kept) IoUs are ever evaluated and no n^2 mask exists.
// Both are latency/dependency-bound rather than HBM-bound.
#include "kernels.cuh"

namespace d2b {
namespace {

typedef unsigned long long u64;

// ------------------------------------------------------------------ A: bitmask
constexpr int kMaskThreads = 256;

// grid (row_blocks, S).  mask[seg][i][w] bit c: box (w*64+c) is suppressed by box i (only j > i).
__global__ void __launch_bounds__(kMaskThreads) nms_mask_kernel(const float4* boxes, const int32_t* counts, int n,
                                                                 int W, float thr, u64* mask) {
  const int seg = blockIdx.y, rb = blockIdx.x;
  const int cnt = counts ? min(counts[seg], n) : n;
  if (rb * 64 >= cnt) return;
  const float4* b = boxes + (size_t)seg * n;
  const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int i = rb * 64 + r;
  const bool live = i < cnt;
  const float4 bi = live ? b[i] : make_float4(0, 0, 0, 0);
  u64* mrow = mask + ((size_t)seg * W * 64 + i) * W;
  const int nb = (cnt + 63) >> 6;
  for (int cb = rb + q; cb < nb; cb += kMaskThreads / 64) {
    u64 bits = 0;
    const int j0 = cb * 64;
    const int jn = min(64, cnt - j0);
    for (int c = 0; c < jn; ++c) {
      const int j = j0 + c;
      const float4 bj = __ldg(b + j);  // warp-uniform address: one broadcast transaction
      if (j > i && d2b_iou(bi, bj) > thr) bits |= (1ull << c);
    }
    if (live) mrow[cb] = bits;
  }
}

// One warp per segment.  WPL = removed-words per lane (W <= 32*WPL).
template <int WPL>
__global__ void __launch_bounds__(32) nms_sweep_kernel(const int32_t* counts, int n, int W, int max_out,
                                                        const u64* mask, int32_t* keep, int32_t* num_keep) {
  __shared__ u64 diag[64];
  const int seg = blockIdx.x, lane = threadIdx.x;
  const int cnt = counts ? min(counts[seg], n) : n;
  const u64* m = mask + (size_t)seg * W * 64 * W;
  int32_t* kp = keep + (size_t)seg * max_out;
  u64 removed[WPL];
#pragma unroll
  for (int s = 0; s < WPL; ++s) removed[s] = 0;
  int kept = 0;
  const int nb = (cnt + 63) >> 6;
  for (int b = 0; b < nb && kept < max_out; ++b) {
    // removed word b lives in lane (b & 31), slot (b >> 5)
    u64 rem = 0;
#pragma unroll
    for (int s = 0; s < WPL; ++s)
      if ((b >> 5) == s) rem = removed[s];
    rem = __shfl_sync(0xffffffffu, rem, b & 31);
    const int rows = min(64, cnt - b * 64);
    if (rows < 64) rem |= ~0ull << rows;
    // diagonal tile -> shared memory
    for (int t = lane; t < 64; t += 32) diag[t] = t < rows ? m[((size_t)b * 64 + t) * W + b] : 0ull;
    __syncwarp();
    u64 keepm = 0;
    if (lane == 0) {
      int left = max_out - kept;
      for (int t = 0; t < rows && left > 0; ++t) {
        if (!((rem >> t) & 1ull)) {
          keepm |= 1ull << t;
          rem |= diag[t];
          --left;
        }
      }
    }
    keepm = __shfl_sync(0xffffffffu, keepm, 0);
    __syncwarp();
    // emit kept positions in order
    const unsigned lo = (unsigned)keepm, hi = (unsigned)(keepm >> 32);
    if ((lo >> lane) & 1u) kp[kept + __popc(lo & ((1u << lane) - 1u))] = b * 64 + lane;
    if ((hi >> lane) & 1u) kp[kept + __popc(lo) + __popc(hi & ((1u << lane) - 1u))] = b * 64 + 32 + lane;
    kept += __popcll(keepm);
    // OR the kept rows into the removed set (words beyond b only matter)
    u64 km = keepm;
    while (km) {
      const int t = __ffsll((long long)km) - 1;
      km &= km - 1;
      const u64* row = m + ((size_t)b * 64 + t) * W;
#pragma unroll
      for (int s = 0; s < WPL; ++s) {
        const int w = s * 32 + lane;
        if (w > b && w < W) removed[s] |= __ldg(row + w);
      }
    }
  }
  for (int j = kept + lane; j < max_out; j += 32) kp[j] = -1;
  if (lane == 0) num_keep[seg] = kept;
}

// ------------------------------------------------------------------ B: capped lazy sweep
constexpr int kLazyThreads = 256;

__global__ void __launch_bounds__(kLazyThreads) nms_lazy_kernel(const float4* boxes, const int32_t* counts, int n,
                                                                 int max_out, float thr, int32_t* keep,
                                                                 int32_t* num_keep) {
  extern __shared__ float4 s_kept[];  // [max_out]
  __shared__ float4 s_blk[64];
  __shared__ u64 s_diag[64];
  __shared__ unsigned s_dead[2];
  __shared__ int s_kept_n;
  const int seg = blockIdx.x, tid = threadIdx.x;
  const int cnt = counts ? min(counts[seg], n) : n;
  const float4* b = boxes + (size_t)seg * n;
  int32_t* kp = keep + (size_t)seg * max_out;
  if (tid == 0) s_kept_n = 0;
  __syncthreads();
  const int r = tid & 63, q = tid >> 6;  // candidate r of the block, quarter q of the kept list / columns
  for (int b0 = 0; b0 < cnt; b0 += 64) {
    const int kept = s_kept_n;
    if (kept >= max_out) break;
    const int rows = min(64, cnt - b0);
    if (tid < 64) s_blk[tid] = tid < rows ? b[b0 + tid] : make_float4(0, 0, 0, 0);
    if (tid < 2) s_dead[tid] = 0;
    if (tid < 64) s_diag[tid] = 0;
    __syncthreads();
    const float4 bi = s_blk[r];
    // (1) candidate r vs kept list (strided by quarter)
    bool dead = false;
    if (r < rows)
      for (int j = q; j < kept; j += kLazyThreads / 64)
        if (d2b_iou(bi, s_kept[j]) > thr) { dead = true; break; }
    if (dead) atomicOr(&s_dead[r >> 5], 1u << (r & 31));
    // (2) diagonal tile: bit c of row r set if box r suppresses box c (c > r); quarter q covers 16 columns
    u64 bits = 0;
    if (r < rows)
      for (int c = max(q * 16, r + 1); c < min(q * 16 + 16, rows); ++c)
        if (d2b_iou(bi, s_blk[c]) > thr) bits |= 1ull << c;
    if (bits) atomicOr(&s_diag[r], bits);
    __syncthreads();
    // (3) serial resolve by one thread
    if (tid == 0) {
      u64 rem = ((u64)s_dead[1] << 32) | (u64)s_dead[0];
      int k = kept;
      for (int t = 0; t < rows && k < max_out; ++t) {
        if (!((rem >> t) & 1ull)) {
          s_kept[k] = s_blk[t];
          kp[k] = b0 + t;
          ++k;
          rem |= s_diag[t];
        }
      }
      s_kept_n = k;
    }
    __syncthreads();
  }
  const int kept = s_kept_n;
  for (int j = kept + tid; j < max_out; j += kLazyThreads) kp[j] = -1;
  if (tid == 0) num_keep[seg] = kept;
}

constexpr int kLazyMaxOut = 8192;      // 128 KB of kept boxes in shared memory
constexpr int kBitmaskMaxN = 65536;    // 32 removed-words per lane

bool use_lazy(int n, int max_out) {
  if (max_out > kLazyMaxOut) return false;
  if (n > kBitmaskMaxN) return true;
  return (long long)max_out * 8 <= n;  // cap much smaller than the candidate list
}

}  // namespace

size_t nms_sorted_workspace_bytes(int S, int n, int max_out) {
  if (S <= 0 || n <= 0 || use_lazy(n, max_out)) return 0;
  const size_t W = (n + 63) / 64;
  return ws_slice((size_t)S * W * 64 * W * sizeof(u64));
}

int nms_sorted(const float* boxes, const int32_t* counts, int S, int n, int max_out, float thr, int32_t* keep,
               int32_t* num_keep, void* ws, cudaStream_t st) {
  if (S <= 0) return D2B_OK;
  D2B_REQUIRE(max_out >= 0 && n >= 0, "nms: negative sizes");
  if (max_out == 0) {
    D2B_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t) * S, st));
    return D2B_OK;
  }
  if (n == 0) {
    D2B_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t) * S, st));
    D2B_CUDA(cudaMemsetAsync(keep, 0xff, sizeof(int32_t) * (size_t)S * max_out, st));
    return D2B_OK;
  }
  const float4* b4 = reinterpret_cast<const float4*>(boxes);
  if (use_lazy(n, max_out)) {
    const size_t smem = (size_t)max_out * sizeof(float4);
    if (smem > 48 * 1024)
      D2B_CUDA(cudaFuncSetAttribute(nms_lazy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_lazy_kernel<<<S, kLazyThreads, smem, st>>>(b4, counts, n, max_out, thr, keep, num_keep);
    D2B_LAUNCH_CHECK();
    return D2B_OK;
  }
  D2B_REQUIRE(n <= kBitmaskMaxN, "nms: n=%d with max_output_size=%d is not supported (n <= %d or cap <= %d)", n,
              max_out, kBitmaskMaxN, kLazyMaxOut);
  const int W = (n + 63) / 64;
  u64* mask = static_cast<u64*>(ws);
  nms_mask_kernel<<<dim3(W, S), kMaskThreads, 0, st>>>(b4, counts, n, W, thr, mask);
  D2B_LAUNCH_CHECK();
  const int wpl = (W + 31) / 32;
  if (wpl <= 1) nms_sweep_kernel<1><<<S, 32, 0, st>>>(counts, n, W, max_out, mask, keep, num_keep);
  else if (wpl <= 2) nms_sweep_kernel<2><<<S, 32, 0, st>>>(counts, n, W, max_out, mask, keep, num_keep);
  else if (wpl <= 4) nms_sweep_kernel<4><<<S, 32, 0, st>>>(counts, n, W, max_out, mask, keep, num_keep);
  else if (wpl <= 8) nms_sweep_kernel<8><<<S, 32, 0, st>>>(counts, n, W, max_out, mask, keep, num_keep);
  else if (wpl <= 16) nms_sweep_kernel<16><<<S, 32, 0, st>>>(counts, n, W, max_out, mask, keep, num_keep);
  else nms_sweep_kernel<32><<<S, 32, 0, st>>>(counts, n, W, max_out, mask, keep, num_keep);
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}

}  // namespace d2b
