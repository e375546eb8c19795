// sort.cu -- segmented bitonic sort of 64-bit composite keys, descending.
//
// Used to order top-k winners / NMS candidates by (score desc, index asc): the
// composite key is (order-preserving score key << 32) | ~index, so one unsigned
// 64-bit descending sort realises TF's tie rule.
//
// Two launches whatever the size: (1) every tile of up to kTile keys is fully sorted in
// shared memory; (2) for segments longer than one tile a single CTA per segment finishes the
// bitonic network in global memory.  The EFFECTIVE length of a segment (next power of two of its
// live count, read on the device) bounds the work: tiles and stages beyond it are skipped, so a
// padded capacity (e.g. Rmax*K candidates of Fast R-CNN) costs nothing when few entries are live.
// Inside a tile a thread keeps 1-8 consecutive keys in registers: strides inside the thread and inside the warp
// (shuffles) need no barrier; only strides of 32 threads or more go through shared memory.
#include "kernels.cuh"
#include "sort_tile.cuh"

namespace d2b {
namespace {

constexpr int kTile = 8192;  // keys per CTA tile (64 KB of shared memory)
constexpr int kSortThreads = 1024;

typedef unsigned long long u64;

__device__ __forceinline__ int eff_len(const int32_t* seg_len, int seg, int P) {
  if (!seg_len) return P;
  int c = seg_len[seg];
  if (c < 1) c = 1;
  if (c > P) c = P;
  int e = 1;
  while (e < c) e <<= 1;
  return e;
}

__global__ void __launch_bounds__(kSortThreads) sort_local(u64* keys, int P, int tile, const int32_t* seg_len,
                                                           const int32_t* skip) {
  grid_dep_sync();
  extern __shared__ u64 s[];
  const int seg = blockIdx.y;
  if (skip && skip[seg]) return;  // already in order
  const int t0 = blockIdx.x * tile;
  const int Pe = eff_len(seg_len, seg, P);
  if (t0 >= Pe) return;
  const int tl = tile < Pe ? tile : Pe;  // live part of this tile (a power of two)
  u64* g = keys + (size_t)seg * P + t0;
  const int live = seg_len ? seg_len[seg] : P;
  if (tl >= 8 * kSortThreads) sort_tile<8>(s, g, g, live, tl, t0);
  else if (tl >= 4 * kSortThreads) sort_tile<4>(s, g, g, live, tl, t0);
  else if (tl >= 2 * kSortThreads) sort_tile<2>(s, g, g, live, tl, t0);
  else sort_tile<1>(s, g, g, live, tl, t0);
}

// Segments longer than one tile: one CTA finishes stages k = 2*tile .. Pe in global memory.
__global__ void __launch_bounds__(kSortThreads) sort_big(u64* keys, int P, int tile, const int32_t* seg_len,
                                                         const int32_t* skip) {
  grid_dep_sync();
  const int seg = blockIdx.x;
  if (skip && skip[seg]) return;
  const int Pe = eff_len(seg_len, seg, P);
  if (Pe <= tile) return;
  u64* g = keys + (size_t)seg * P;
  for (int k = tile << 1; k <= Pe; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = threadIdx.x; p < Pe / 2; p += kSortThreads) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const bool desc = ((i & k) == 0);
        const u64 a = g[i], b = g[i | j];
        if (desc ? (a < b) : (a > b)) { g[i] = b; g[i | j] = a; }
      }
      __syncthreads();
    }
}

}  // namespace

int sort_segments_desc(unsigned long long* keys, int S, int P, const int32_t* seg_len, cudaStream_t st,
                       const int32_t* skip) {
  if (S <= 0 || P <= 1) return D2B_OK;
  D2B_REQUIRE((P & (P - 1)) == 0, "sort: P=%d is not a power of two", P);
  const int tile = P < kTile ? P : kTile;
  const size_t smem = (size_t)(tile + tile / 2) * sizeof(u64);  // + the pad slots of sort_phys (at most 1 per 2 keys)
  if (smem > 48 * 1024)
    D2B_CUDA(cudaFuncSetAttribute(sort_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  D2B_CUDA(launch_pdl(sort_local, dim3(P / tile, S), dim3(kSortThreads), smem, st, 0, keys, P, tile, seg_len, skip));
  D2B_LAUNCH_CHECK();
  if (P > tile) {
    D2B_CUDA(launch_pdl(sort_big, dim3(S), dim3(kSortThreads), 0, st, 0, keys, P, tile, seg_len, skip));
    D2B_LAUNCH_CHECK();
  }
  return D2B_OK;
}

}  // namespace d2b
