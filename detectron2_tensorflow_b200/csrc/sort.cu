// sort.cu -- segmented bitonic sort of 64-bit composite keys, descending.
//
// Used to order top-k winners / NMS candidates by (score desc, index asc): the
// composite key is (order-preserving score key << 32) | ~index, so one unsigned
// 64-bit descending sort realises TF's tie rule.
// Segments up to kTile keys are sorted by one CTA in shared memory; longer ones
// alternate global compare-exchange steps (stride >= kTile) with shared-memory
// tails, the classic tiled bitonic schedule.  Launch count depends only on P.
#include "kernels.cuh"

namespace d2b {
namespace {

constexpr int kTile = 4096;     // keys per CTA tile (32 KB of shared memory)
constexpr int kSortThreads = 512;

typedef unsigned long long u64;

__device__ __forceinline__ void cmpx(u64& a, u64& b, bool desc) {
  const bool sw = desc ? (a < b) : (a > b);
  if (sw) { u64 t = a; a = b; b = t; }
}

// Sort stages k = 2 .. min(P, tile) entirely in shared memory (k_from == 2), or run the
// tail substages j = tile/2 .. 1 of a single outer stage `k_only` (k_from == 0).
__global__ void __launch_bounds__(kSortThreads) sort_local(u64* keys, int P, int tile, const int32_t* seg_len,
                                                            int k_only, int mask_dead) {
  extern __shared__ u64 s[];
  const int seg = blockIdx.y;
  const int t0 = blockIdx.x * tile;
  u64* g = keys + (size_t)seg * P + t0;
  const int live = seg_len ? seg_len[seg] : P;
  for (int i = threadIdx.x; i < tile; i += kSortThreads) {
    u64 v = g[i];
    if (mask_dead && t0 + i >= live) v = 0ull;
    s[i] = v;
  }
  __syncthreads();
  if (k_only == 0) {
    for (int k = 2; k <= tile; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int p = threadIdx.x; p < tile / 2; p += kSortThreads) {
          const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
          const bool desc = (((t0 + i) & k) == 0);
          cmpx(s[i], s[i | j], desc);
        }
        __syncthreads();
      }
  } else {
    const int k = k_only;
    for (int j = tile >> 1; j > 0; j >>= 1) {
      for (int p = threadIdx.x; p < tile / 2; p += kSortThreads) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const bool desc = (((t0 + i) & k) == 0);
        cmpx(s[i], s[i | j], desc);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < tile; i += kSortThreads) g[i] = s[i];
}

// One global compare-exchange step (stride j >= kTile) of outer stage k.
__global__ void sort_global_step(u64* keys, int P, int k, int j) {
  const int seg = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P / 2) return;
  const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
  u64* g = keys + (size_t)seg * P;
  u64 a = g[i], b = g[i | j];
  const bool desc = ((i & k) == 0);
  const bool sw = desc ? (a < b) : (a > b);
  if (sw) { g[i] = b; g[i | j] = a; }
}

}  // namespace

int sort_segments_desc(unsigned long long* keys, int S, int P, const int32_t* seg_len, cudaStream_t st) {
  if (S <= 0 || P <= 1) {
    if (S > 0 && P == 1 && seg_len) {
      // single-slot segments: nothing to order; dead slots are ignored by consumers via counts
    }
    return D2B_OK;
  }
  D2B_REQUIRE((P & (P - 1)) == 0, "sort: P=%d is not a power of two", P);
  const int tile = P < kTile ? P : kTile;
  const dim3 grid(P / tile, S);
  const size_t smem = (size_t)tile * sizeof(u64);
  sort_local<<<grid, kSortThreads, smem, st>>>(keys, P, tile, seg_len, 0, seg_len != nullptr);
  D2B_LAUNCH_CHECK();
  for (int k = tile << 1; k <= P; k <<= 1) {
    for (int j = k >> 1; j >= tile; j >>= 1) {
      const dim3 g2((P / 2 + 255) / 256, S);
      sort_global_step<<<g2, 256, 0, st>>>(keys, P, k, j);
      D2B_LAUNCH_CHECK();
    }
    sort_local<<<grid, kSortThreads, smem, st>>>(keys, P, tile, nullptr, k, 0);
    D2B_LAUNCH_CHECK();
  }
  return D2B_OK;
}

}  // namespace d2b
