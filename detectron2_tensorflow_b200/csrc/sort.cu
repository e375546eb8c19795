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

// One compare-exchange of the bitonic network on register values.
__device__ __forceinline__ void cmpx(u64& lo, u64& hi, bool desc) {
  const u64 a = lo, b = hi;
  if (desc ? (a < b) : (a > b)) { lo = b; hi = a; }
}

// Steps j = jstart .. 1 of merge phase k on the E consecutive keys r[] a thread holds (index base + e): strides
// >= E through warp shuffles (jstart <= 16 E), strides < E inside the thread.  No barrier, no shared memory.
template <int E>
__device__ __forceinline__ void low_steps(u64 (&r)[E], int base, int t0, int k, int jstart) {
  for (int j = jstart; j >= E; j >>= 1) {
    const int lane_mask = j / E;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const u64 mine = r[e];
      const u64 other = __shfl_xor_sync(0xffffffffu, mine, lane_mask);
      const int i = base + e;
      const bool keep_max = ((((t0 + i) & k) == 0) == ((i & j) == 0));  // descending pair: the lower index keeps the max
      const u64 mx = mine > other ? mine : other, mn = mine > other ? other : mine;
      r[e] = keep_max ? mx : mn;
    }
  }
#pragma unroll
  for (int j = E / 2; j > 0; j >>= 1) {
    if (j <= jstart) {
#pragma unroll
      for (int e = 0; e < E; ++e)
        if ((e & j) == 0) cmpx(r[e], r[e | j], ((t0 + base + e) & k) == 0);
    }
  }
}

// Shared-memory slot of key i: a thread's E consecutive keys are E * 8 bytes apart from its neighbour's, which would be
// an E-way bank conflict on every register load / store; one pad slot per E keys makes consecutive threads hit
// consecutive bank pairs (stride E + 1, odd).
template <int E>
__device__ __forceinline__ int sort_phys(int i) { return E > 1 ? i + i / E : i; }

// Full bitonic sort of the tl keys g_in[0..tl) (entries at or beyond `live` count as 0) into g_out, staged in s[]
// (tl = a power of two, tl / E <= blockDim threads hold E keys each).  Merge phases up to k = 32 E run entirely in
// registers; later phases do their long strides (>= 32 E) in shared memory two levels per barrier and the rest in
// registers: ~20 barriers for 4,096 keys instead of 78.
template <int E>
__device__ __forceinline__ void sort_tile(u64* s, const u64* g_in, u64* g_out, const int live, const int tl, const int t0) {
  const int t = threadIdx.x;
  auto ph = [](int i) { return sort_phys<E>(i); };
  for (int i = t; i < tl; i += blockDim.x) s[ph(i)] = (t0 + i < live) ? g_in[i] : 0ull;
  __syncthreads();
  const int nact = tl / E;
  const bool warp_on = (t & ~31) < nact;
  const int base = t * E;
  u64 r[E];
  if (warp_on) {
#pragma unroll
    for (int e = 0; e < E; ++e) r[e] = (base + e < tl) ? s[ph(base + e)] : 0ull;
    const int kA = tl < 32 * E ? tl : 32 * E;
    for (int k = 2; k <= kA; k <<= 1) low_steps<E>(r, base, t0, k, k >> 1);
    if (base < tl) {
#pragma unroll
      for (int e = 0; e < E; ++e) s[ph(base + e)] = r[e];
    }
  }
  __syncthreads();
  for (int k = 64 * E; k <= tl; k <<= 1) {
    int j = k >> 1;
    for (; j >= 64 * E; j >>= 2) {  // strides j and j/2 on quads
      const int j2 = j >> 1;
      for (int q = t; q < tl / 4; q += blockDim.x) {
        const int i = ((q & ~(j2 - 1)) << 2) | (q & (j2 - 1));
        const bool desc = (((t0 + i) & k) == 0);
        u64 a0 = s[ph(i)], a1 = s[ph(i | j2)], a2 = s[ph(i | j)], a3 = s[ph(i | j | j2)];
        cmpx(a0, a2, desc); cmpx(a1, a3, desc);
        cmpx(a0, a1, desc); cmpx(a2, a3, desc);
        s[ph(i)] = a0; s[ph(i | j2)] = a1; s[ph(i | j)] = a2; s[ph(i | j | j2)] = a3;
      }
      __syncthreads();
    }
    if (j >= 32 * E) {  // one long stride left
      for (int p = t; p < tl / 2; p += blockDim.x) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        u64 a = s[ph(i)], b = s[ph(i | j)];
        cmpx(a, b, ((t0 + i) & k) == 0);
        s[ph(i)] = a; s[ph(i | j)] = b;
      }
      __syncthreads();
      j >>= 1;
    }
    if (warp_on) {
#pragma unroll
      for (int e = 0; e < E; ++e) r[e] = (base + e < tl) ? s[ph(base + e)] : 0ull;
      low_steps<E>(r, base, t0, k, j);
      if (base < tl) {
#pragma unroll
        for (int e = 0; e < E; ++e) s[ph(base + e)] = r[e];
      }
    }
    __syncthreads();
  }
  for (int i = t; i < tl; i += blockDim.x) g_out[i] = s[ph(i)];
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
