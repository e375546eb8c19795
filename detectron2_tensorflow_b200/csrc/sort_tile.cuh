// sort_tile.cuh -- the in-CTA bitonic sort of 64-bit keys shared by sort.cu (segment tiles) and topk.cu (the k winners
// of a RetinaNet row): keys in registers, strides inside a thread / a warp without barriers, long strides through
// shared memory two levels per barrier.
#pragma once
#include "common.cuh"

namespace d2b {

typedef unsigned long long sort_key_t;

// One compare-exchange of the bitonic network on register values.
__device__ __forceinline__ void cmpx(sort_key_t& lo, sort_key_t& hi, bool desc) {
  const sort_key_t a = lo, b = hi;
  if (desc ? (a < b) : (a > b)) { lo = b; hi = a; }
}

// Steps j = jstart .. 1 of merge phase k on the E consecutive keys r[] a thread holds (index base + e): strides
// >= E through warp shuffles (jstart <= 16 E), strides < E inside the thread.  No barrier, no shared memory.
template <int E>
__device__ __forceinline__ void low_steps(sort_key_t (&r)[E], int base, int t0, int k, int jstart) {
  for (int j = jstart; j >= E; j >>= 1) {
    const int lane_mask = j / E;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const sort_key_t mine = r[e];
      const sort_key_t other = __shfl_xor_sync(0xffffffffu, mine, lane_mask);
      const int i = base + e;
      const bool keep_max = ((((t0 + i) & k) == 0) == ((i & j) == 0));  // descending pair: the lower index keeps the max
      const sort_key_t mx = mine > other ? mine : other, mn = mine > other ? other : mine;
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
template <int E, bool PAD>
__device__ __forceinline__ int sort_phys(int i) { return (PAD && E > 1) ? i + i / E : i; }

// Full bitonic sort of the tl keys g_in[0..tl) (entries at or beyond `live` count as 0) into g_out, staged in s[]
// (tl = a power of two, tl / E <= blockDim threads hold E keys each).  Merge phases up to k = 32 E run entirely in
// registers; later phases do their long strides (>= 32 E) in shared memory two levels per barrier and the rest in
// registers: ~20 barriers for 4,096 keys instead of 78.
template <int E, bool PAD>
__device__ __forceinline__ void sort_tile_core(sort_key_t* s, const int tl, const int t0) {
  const int t = threadIdx.x;
  auto ph = [](int i) { return sort_phys<E, PAD>(i); };
  const int nact = tl / E;
  const bool warp_on = (t & ~31) < nact;
  const int base = t * E;
  sort_key_t r[E];
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
        sort_key_t a0 = s[ph(i)], a1 = s[ph(i | j2)], a2 = s[ph(i | j)], a3 = s[ph(i | j | j2)];
        cmpx(a0, a2, desc); cmpx(a1, a3, desc);
        cmpx(a0, a1, desc); cmpx(a2, a3, desc);
        s[ph(i)] = a0; s[ph(i | j2)] = a1; s[ph(i | j)] = a2; s[ph(i | j | j2)] = a3;
      }
      __syncthreads();
    }
    if (j >= 32 * E) {  // one long stride left
      for (int p = t; p < tl / 2; p += blockDim.x) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        sort_key_t a = s[ph(i)], b = s[ph(i | j)];
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
}

// g_in[0..tl) (entries at or beyond `live` count as 0) -> sorted descending (ascending when t0 & tl) -> g_out, staged
// in the padded shared buffer s[] (tl + tl / 2 entries at most).
template <int E>
__device__ __forceinline__ void sort_tile(sort_key_t* s, const sort_key_t* g_in, sort_key_t* g_out, const int live, const int tl, const int t0) {
  const int t = threadIdx.x;
  for (int i = t; i < tl; i += blockDim.x) s[sort_phys<E, true>(i)] = (t0 + i < live) ? g_in[i] : 0ull;
  __syncthreads();
  sort_tile_core<E, true>(s, tl, t0);
  for (int i = t; i < tl; i += blockDim.x) g_out[i] = s[sort_phys<E, true>(i)];
}

// Keys already in shared memory, unpadded, sorted in place (tl <= 2 * blockDim.x).  Ends with a barrier.
__device__ __forceinline__ void sort_smem_desc(sort_key_t* s, const int tl) {
  if (tl > (int)blockDim.x) sort_tile_core<2, false>(s, tl, 0);
  else sort_tile_core<1, false>(s, tl, 0);
}


}  // namespace d2b
