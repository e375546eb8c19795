// topk.cu -- segmented radix-select top-k over (image, level) rows.
//
// Replaces tf.nn.top_k at rpn_outputs.py:70,106 and retinanet.py:326 (with the
// sigmoid of retinanet.py:322 fused into the key load).  Tie rule of TF's TopKV2:
// among equal values the lower index wins -- realised by ranking 64-bit
// composites (order-preserving value key << 32 | ~index), which are unique.
//
// All rows of all levels are processed by the same launches (no per-image loop):
//   1. radix-select passes over the composite, most significant digit first
//      (11/11/10 bits of the value key, then 11/11/10 bits of ~index).  A pass is
//      one streaming read of the live rows: shared-memory histogram per CTA chunk,
//      merged into a per-row global histogram; the last CTA of a row to finish
//      locates the digit that holds the k-th largest element.  A row stops as soon
//      as its boundary bucket is taken whole, so the index passes only ever run
//      for rows whose k-th value is tied (CTAs of finished rows exit at once).
//   2. collect: one more streaming read gathers every element with composite >=
//      the row's threshold (exactly k_r of them) through a warp-aggregated slot
//      counter.
//   3. sort the k_r winners (sort.cu) and emit values / indices.
// SIGMOID rows (RetinaNet) first run a histogram over the raw LOGIT keys (no transform) to get a cutoff
// c <= k-th largest logit; elements below c - margin can never reach the top-k, so the bit-exact sigmoid
// (ALU-heavy) is evaluated only for the few elements above it.  The margin keeps the computed sigmoid
// strictly ordered across the gap (true ratio >= 1 + 3e-5 vs <= 4e-7 evaluation error) and the shortcut is
// disabled in the saturated / underflowing tails (c outside (-80, 8)) where plateaus make index ties matter.
// HBM-bound scan (4 B per candidate score per pass); no tensor cores.
#include "kernels.cuh"
#include "sort_tile.cuh"

namespace d2b {
namespace {

typedef unsigned long long u64;

constexpr int kBins = 2048;
constexpr int kPasses = 6;
constexpr int kHistThreads = 256;
constexpr int kChunk = 8192;        // elements per CTA per pass (rows up to 1 M elements)
constexpr int kCandCap = 65536;     // candidate-list capacity per SIGMOID row (beyond it the row is re-scanned)
#ifndef D2B_CHUNK_LONG
#define D2B_CHUNK_LONG 65536
#endif
constexpr int kChunkLong = D2B_CHUNK_LONG;   // ... for longer rows (RetinaNet class scores): amortises the per-CTA setup
constexpr int kBatch = 8;       // independent loads in flight per thread

__constant__ int c_shift[kPasses] = {53, 42, 32, 21, 10, 0};
__constant__ int c_bits[kPasses] = {11, 11, 10, 11, 11, 10};

struct RowState {
  u64 prefix;        // digits resolved so far (right-aligned)
  u64 threshold;     // final: take every composite >= threshold
  unsigned k_rem;    // winners still to be found inside the current prefix
  unsigned active;   // 1 while more passes are needed
  unsigned k_r;      // min(k, len, k_limit)
  unsigned out_count;  // collect slot counter
  unsigned cut_key;    // SIGMOID rows: logit keys below this can never reach the top-k (0 = evaluate all)
  unsigned pre_done;   // CTAs of the logit pre-histogram that have finished
  unsigned cand_count; // SIGMOID rows with a cutoff: composites appended to the candidate list by pass 0
  unsigned compact;    // 1 => the candidate list holds every element that can matter: later passes read it
  unsigned redo;       // 1 => the sampled cutoff left fewer than k_r elements: pass 0 is repeated without a cutoff
  unsigned cut_hi;     // key of the cutoff c itself (cut_key = key of c - margin)
  unsigned hi_count;   // elements with logit >= c counted by pass 0: the exactness argument needs >= k_r of them
  unsigned done[kPasses];
};

struct TopkArgs {
  TopkDesc d;
  int chunks[D2B_MAX_LEVELS];      // CTAs per row in group g
  int chunk_elems[D2B_MAX_LEVELS]; // elements per CTA in group g (multiple of 4)
  int cta_begin[D2B_MAX_LEVELS + 1];  // first CTA of group g
  RowState* state;
  unsigned* hist;  // [rows][kPasses][kBins]
  int P;           // padded k
  unsigned* prehist;  // SIGMOID only: [rows][kBins] histogram of raw logit keys (top 11 bits)
  u64* cand;          // SIGMOID only: [rows][kCandCap] composites of the elements above the cutoff
};

__device__ __forceinline__ bool locate(const TopkArgs& a, int cta, int& g, int& img, int& chunk) {
  g = 0;
  while (g + 1 < a.d.G && cta >= a.cta_begin[g + 1]) ++g;
  const int rem = cta - a.cta_begin[g];
  img = rem / a.chunks[g];
  chunk = rem - img * a.chunks[g];
  return img < a.d.rows_per_group;
}

__device__ __forceinline__ uint32_t value_key(float x, int transform) {
  if (transform == D2B_TOPK_SIGMOID) x = d2b_sigmoidf(x);
  return float_to_key(x);
}
__device__ __forceinline__ u64 composite_of(uint32_t key, unsigned idx) {
  return ((u64)key << 32) | (u64)(0xffffffffu - idx);
}

// Calls f(value, index, valid) for every element of [beg, end) -- and, with valid == false, for the padding
// slots -- with warp-uniform control flow (f may use warp collectives).  16-byte loads when the chunk is
// 16-byte aligned, kBatch/4 vectors (or kBatch scalars) in flight per thread.
// kStream: the row is read exactly once (RetinaNet rows with a sampled cutoff) -> evict-first loads.
template <int kV = kBatch / 4, bool kStream = false, typename F>
__device__ __forceinline__ void for_each_elem(const float* x, long long beg, long long end, F f) {
  const bool vec = ((reinterpret_cast<uintptr_t>(x + beg) & 15) == 0);
  if (vec) {
    const long long nvec = (end - beg) >> 2;
    const float4* xv = reinterpret_cast<const float4*>(x + beg);
    const long long step = (long long)kV * kHistThreads;
    auto load = [&](float4* q, long long v0) {
#pragma unroll
      for (int u = 0; u < kV; ++u) {
        const long long vi = v0 + (long long)u * kHistThreads + threadIdx.x;
        q[u] = vi < nvec ? (kStream ? __ldcs(xv + vi) : __ldg(xv + vi)) : make_float4(0, 0, 0, 0);
      }
    };
    auto process = [&](const float4* q, long long v0) {
#pragma unroll
      for (int u = 0; u < kV; ++u) {
        const long long vi = v0 + (long long)u * kHistThreads + threadIdx.x;
        const bool ok = vi < nvec;
        const long long i = beg + 4 * vi;
        f(q[u].x, i, ok); f(q[u].y, i + 1, ok); f(q[u].z, i + 2, ok); f(q[u].w, i + 3, ok);
      }
    };
    // (measured: keeping the loads of trip t+1 in flight while trip t is processed does not help -- 0.7395 -> 0.748 ms
    // for the RetinaNet scan -- it only costs registers)
    for (long long v0 = 0; v0 < nvec; v0 += step) {  // block-uniform trip count
      float4 q[kV];
      load(q, v0);
      process(q, v0);
    }
    const long long t0 = beg + 4 * nvec;  // < 4 leftover elements
    if (t0 < end) {
      const long long i = t0 + threadIdx.x;
      const bool ok = i < end;
      f(ok ? __ldg(x + i) : 0.0f, i, ok);
    }
  } else {
    for (long long i0 = beg; i0 < end; i0 += (long long)kBatch * kHistThreads) {
      float q[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const long long i = i0 + (long long)u * kHistThreads + threadIdx.x;
        q[u] = i < end ? __ldg(x + i) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const long long i = i0 + (long long)u * kHistThreads + threadIdx.x;
        f(q[u], i, i < end);
      }
    }
  }
}

__global__ void topk_init(TopkArgs a, int rows, int32_t* seg_len, int32_t* out_counts) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int g = r % a.d.G;
  long long kr = a.d.k;
  if (a.d.row_len[g] < kr) kr = a.d.row_len[g];
  if (a.d.k_limit[g] > 0 && a.d.k_limit[g] < kr) kr = a.d.k_limit[g];
  if (kr < 0) kr = 0;
  RowState s;
  s.prefix = 0; s.k_rem = (unsigned)kr; s.k_r = (unsigned)kr; s.out_count = 0;
  for (int p = 0; p < kPasses; ++p) s.done[p] = 0;
  if (kr == 0) { s.active = 0; s.threshold = ~0ull; }            // take nothing
  else if (kr == a.d.row_len[g]) { s.active = 0; s.threshold = 0ull; }  // take the whole row
  else { s.active = 1; s.threshold = 0ull; }
  s.cut_key = 0u; s.pre_done = 0u; s.cand_count = 0u; s.compact = 0u; s.redo = 0u; s.cut_hi = 0u; s.hi_count = 0u;
  a.state[r] = s;
  seg_len[r] = (int32_t)kr;
  if (out_counts) out_counts[r] = (int32_t)kr;
}

// The end of a radix pass for one CTA of `row`: count the arrival; the LAST CTA of the row to finish locates the
// digit of the row's global histogram `gh` that holds the k_rem-th largest element (or, for pass 0 of a row with a
// sampled cutoff, discards the pass when the cutoff cannot be trusted).  Called by all kHistThreads threads.
__device__ __forceinline__ void pass_tail(const TopkArgs& a, RowState* st, unsigned* gh, int pass, unsigned expect,
                                          bool compact, unsigned cut, u64 prefix, int shift, int bits, int redo) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&st->done[pass], 1u);
    s_last = (t == expect - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0 && a.cand && !compact && pass == 0 && cut != 0u)
    st->compact = (*reinterpret_cast<volatile unsigned*>(&st->cand_count) <= (unsigned)kCandCap) ? 1u : 0u;
  // ---- last CTA of the row: find the digit holding the k_rem-th largest element
  const unsigned k_rem = st->k_rem;
  constexpr int kPer = kBins / kHistThreads;  // 8 bins per thread, thread t owns the t-th highest group
  unsigned loc[kPer];
  unsigned sum = 0;
  const int top = kBins - 1 - threadIdx.x * kPer;  // this thread's highest bin
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    loc[j] = __ldcg(gh + (top - j));
    sum += loc[j];
  }
  __shared__ unsigned scan[kHistThreads];
  scan[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < kHistThreads; off <<= 1) {  // inclusive scan from the top bins down
    unsigned v = threadIdx.x >= off ? scan[threadIdx.x - off] : 0;
    __syncthreads();
    scan[threadIdx.x] += v;
    __syncthreads();
  }
  const unsigned incl = scan[threadIdx.x], excl = incl - sum;
  if (pass == 0 && cut != 0u &&
      (scan[kHistThreads - 1] < k_rem || *reinterpret_cast<volatile unsigned*>(&st->hi_count) < k_rem)) {
    // The cutoff came from a SAMPLE of the row (topk_prehist) and kept fewer than k_r elements -- or fewer than k_r
    // of them reach the cutoff c itself, in which case the k-th largest logit may sit inside the margin below c,
    // where the computed-sigmoid order of an excluded neighbour is not guaranteed (header comment): throw this
    // pass away and let the repeat launch histogram the whole row without a cutoff.
    for (int j = 0; j < kPer; ++j) gh[top - j] = 0u;
    if (threadIdx.x == 0) {
      st->cut_key = 0u; st->cand_count = 0u; st->compact = 0u; st->done[0] = 0u; st->redo = 1u; st->hi_count = 0u;
    }
    return;
  }
  if (redo && threadIdx.x == 0) st->redo = 0u;
  if (excl < k_rem && k_rem <= incl) {
    unsigned above = excl;
    for (int j = 0; j < kPer; ++j) {
      if (above + loc[j] >= k_rem) {
        const unsigned digit = (unsigned)(top - j);
        D2B_BOUND(digit, 1u << bits);
        const unsigned need = k_rem - above;
        const u64 np = (prefix << bits) | digit;
        st->prefix = np;
        st->k_rem = need;
        if (loc[j] == need || pass == kPasses - 1) {  // bucket taken whole: row resolved
          st->threshold = np << shift;
          st->active = 0;
        }
        break;
      }
      above += loc[j];
    }
  }
}

constexpr int kCopies = 4;  // replicated pass-0 histograms (lane & 3) to spread same-bin atomics
__global__ void __launch_bounds__(kHistThreads) topk_hist(TopkArgs a, int pass, int redo, int skip_cut) {
  __shared__ unsigned sh[kCopies][kBins];
  int g, img, chunk;
  if (!locate(a, blockIdx.x, g, img, chunk)) return;
  const int row = img * a.d.G + g;
  RowState* st = a.state + row;
  if (!st->active) return;  // uniform per CTA
  if (redo && !st->redo) return;
  if (skip_cut && st->cut_key != 0u) return;  // pass 0 of this row is topk_scan_cut_kernel's  // the repeat launch only serves rows whose sampled cutoff failed
  const u64 prefix = st->prefix;
  const int shift = c_shift[pass], bits = c_bits[pass];
  const unsigned mask = (1u << bits) - 1u;
  // pass 0 sees every element: shared-memory histogram.  Later passes only count the few elements
  // inside the current prefix: they go straight to the row's global histogram -- and so does pass 0 of a row with a
  // sampled cutoff (only ~3k of its elements pass the compare; zeroing and merging a 32 KB shared histogram per
  // 32 KB of input cost as much as the streaming read itself).
  const bool use_smem = (pass == 0) && (st->cut_key == 0u);
  unsigned* gh = a.hist + ((size_t)row * kPasses + pass) * kBins;
  if (use_smem) {
    for (int i = threadIdx.x; i < kCopies * kBins; i += kHistThreads) (&sh[0][0])[i] = 0;
    __syncthreads();
  }
  unsigned* my = use_smem ? sh[threadIdx.x & (kCopies - 1)] : gh;
  const long long len = a.d.row_len[g];
  const float* x = a.d.scores[g] + (size_t)img * len;
  const long long beg = (long long)chunk * a.chunk_elems[g];
  const long long end = beg + a.chunk_elems[g] < len ? beg + a.chunk_elems[g] : len;
  const int hi_shift = shift + bits;  // bits above the current digit
  const unsigned cut = st->cut_key;
  const int transform = a.d.transform;
  const bool compact = st->compact != 0;  // set by pass 0; block-uniform
  unsigned expect = (unsigned)a.chunks[g];
  if (compact) {
    // every element that can still matter is in the candidate list: one CTA walks it
    if (chunk != 0) return;
    expect = 1u;
    const u64* cand = a.cand + (size_t)row * kCandCap;
    const unsigned n = st->cand_count;
    for (unsigned j0 = 0; j0 < n; j0 += kBatch * kHistThreads) {  // kBatch independent loads in flight per thread
      u64 c[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const unsigned j = j0 + u * kHistThreads + threadIdx.x;
        c[u] = j < n ? __ldcg(cand + j) : 0ull;  // 0 never matches a live prefix bucket below
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const unsigned j = j0 + u * kHistThreads + threadIdx.x;
        if (j < n && (c[u] >> hi_shift) == prefix) atomicAdd(gh + ((unsigned)(c[u] >> shift) & mask), 1u);
      }
    }
  } else {
    // Candidate list: pass 0 of a row with a logit cutoff remembers every element above the cutoff
    // (measured: doing the same after pass 1 for rows without a cutoff does not pay at RPN sizes).
    const bool cand0 = (pass == 0 && cut != 0u);
    u64* cand = (cand0 && a.cand) ? a.cand + (size_t)row * kCandCap : nullptr;
    // key(v) >= cut  <=>  v >= cut_f for cut != 0 (keys are monotone in the float order and NaN has key 0): one
    // compare per element instead of the ~8-instruction key -- pass 0 of the RetinaNet rows was issue-bound on it
    const float cut_f = key_to_float(cut);
    const bool use_f = (cut != 0u) && (cut_f == cut_f);
    const unsigned cut_hi = st->cut_hi;
    auto body = [&](float v, long long i, bool ok) {
      bool in = ok && (use_f ? (v >= cut_f) : (float_to_key(v) >= cut));
      unsigned digit = 0;
      if (in) {
        const u64 c = composite_of(value_key(v, transform), (unsigned)i);
        if (cand) {
          const unsigned slot = atomicAdd(&st->cand_count, 1u);
          if (slot < (unsigned)kCandCap) cand[slot] = c;
        }
        if (cand0 && float_to_key(v) >= cut_hi) atomicAdd(&st->hi_count, 1u);  // (rare path: ~3 k_r elements per row)
        in = (pass == 0) || ((c >> hi_shift) == prefix);
        digit = (unsigned)(c >> shift) & mask;
      }
      D2B_BOUND(digit, kBins);
      if (in) atomicAdd(my + digit, 1u);
    };
    // rows with a cutoff do one compare per element: keep 4 x 16 B per thread in flight to stay HBM-bound
    if (cand0) {  // (only the redo-free generic route gets here: cutoff rows normally run topk_scan_cut_kernel)
      for_each_elem<4, true>(x, beg, end, body);
    } else {
      for_each_elem(x, beg, end, body);
    }
  }
  __syncthreads();
  if (use_smem) {
    for (int i = threadIdx.x; i < kBins; i += kHistThreads) {
      unsigned v = 0;
#pragma unroll
      for (int c = 0; c < kCopies; ++c) v += sh[c][i];
      if (v) atomicAdd(gh + i, v);
    }
  }
  pass_tail(a, st, gh, pass, expect, compact, cut, prefix, shift, bits, redo);
}

// Pass 0 of the rows that have a sampled cutoff (the RetinaNet class-logit rows: the one full read of 2 GB at
// N = 32) as a kernel of its own: ONE float compare per element in straight-line code -- key(v) >= cut <=> v >= cut_f
// for cut != 0 (keys are monotone in the float order, NaN has key 0) -- and everything else (key, candidate append,
// counters, global histogram) behind the rarely taken branch.  No shared histogram and few registers, so the
// scan runs at full occupancy; rows without a cutoff stay with topk_hist(pass 0).
#ifndef D2B_SCAN_KV
#define D2B_SCAN_KV 4    // 16-byte loads in flight per thread (A/B: 8 -> 400 us vs 383 us)
#endif
#ifndef D2B_SCAN_MINB
#define D2B_SCAN_MINB 8  // CTAs per SM the register budget must allow (A/B: 6 / 4 / 1 are slower)
#endif
__global__ void __launch_bounds__(kHistThreads, D2B_SCAN_MINB) topk_scan_cut_kernel(TopkArgs a) {
  int g, img, chunk;
  if (!locate(a, blockIdx.x, g, img, chunk)) return;
  const int row = img * a.d.G + g;
  RowState* st = a.state + row;
  const unsigned cut = st->cut_key;
  if (!st->active || cut == 0u) return;
  const float cut_f = key_to_float(cut);  // (cut != 0 comes from a finite logit: never NaN)
  const unsigned cut_hi = st->cut_hi;
  const int transform = a.d.transform;
  unsigned* gh = a.hist + ((size_t)row * kPasses + 0) * kBins;
  const long long len = a.d.row_len[g];
  const float* x = a.d.scores[g] + (size_t)img * len;
  const long long beg = (long long)chunk * a.chunk_elems[g];
  const long long end = beg + a.chunk_elems[g] < len ? beg + a.chunk_elems[g] : len;
  u64* cand = a.cand ? a.cand + (size_t)row * kCandCap : nullptr;
  const int sh0 = c_shift[0];
  const unsigned mk0 = (1u << c_bits[0]) - 1u;
  // the rare path (~3 k_r elements per row), ONE copy of it: the element is re-read by index (an L1/L2 hit) so that
  // the hot loop needs no per-element code -- with the slow path inlined per element the loop was 25 KB of SASS
  // and the scan stalled on instruction fetch
  auto slow = [&](long long i) {
    const float v = __ldg(x + i);
    if (v >= cut_f) {
      const u64 c = composite_of(value_key(v, transform), (unsigned)i);
      if (cand) {
        const unsigned slot = atomicAdd(&st->cand_count, 1u);
        if (slot < (unsigned)kCandCap) cand[slot] = c;
      }
      if (float_to_key(v) >= cut_hi) atomicAdd(&st->hi_count, 1u);
      atomicAdd(gh + ((unsigned)(c >> sh0) & mk0), 1u);
    }
  };
  if ((reinterpret_cast<uintptr_t>(x + beg) & 15) == 0) {
    const long long nvec = (end - beg) >> 2;
    const float4* xv = reinterpret_cast<const float4*>(x + beg);
    constexpr int kV = D2B_SCAN_KV;
    const long long step = (long long)kV * kHistThreads;
    for (long long v0 = 0; v0 < nvec; v0 += step) {
      float4 q[kV];
#pragma unroll
      for (int u = 0; u < kV; ++u) {
        const long long vi = v0 + (long long)u * kHistThreads + threadIdx.x;
        q[u] = vi < nvec ? __ldcs(xv + vi) : make_float4(0, 0, 0, 0);
      }
      unsigned hit = 0;  // bit u: some element of vector u reaches the cutoff (max ignores NaN, which never does)
#pragma unroll
      for (int u = 0; u < kV; ++u) {
        const long long vi = v0 + (long long)u * kHistThreads + threadIdx.x;
        const float m = fmaxf(fmaxf(q[u].x, q[u].y), fmaxf(q[u].z, q[u].w));
        hit |= (vi < nvec && m >= cut_f) ? (1u << u) : 0u;
      }
      while (hit) {  // rare
        const int u = __ffs(hit) - 1;
        hit &= hit - 1;
        const long long i = beg + 4 * (v0 + (long long)u * kHistThreads + threadIdx.x);
        for (int c = 0; c < 4; ++c) slow(i + c);
      }
    }
    for (long long i = beg + 4 * nvec + threadIdx.x; i < end; i += kHistThreads) slow(i);  // < 4 leftover elements
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += kHistThreads) slow(i);  // unaligned rows: plain scalar scan
  }
  __syncthreads();
  pass_tail(a, st, gh, 0, (unsigned)a.chunks[g], false, cut, 0ull, c_shift[0], c_bits[0], 0);
}

__global__ void __launch_bounds__(kHistThreads) topk_collect(TopkArgs a, u64* out_keys) {
  int g, img, chunk;
  if (!locate(a, blockIdx.x, g, img, chunk)) return;
  const int row = img * a.d.G + g;
  RowState* st = a.state + row;
  if (st->k_r == 0) return;
  const u64 thr = st->threshold;
  const long long len = a.d.row_len[g];
  const float* x = a.d.scores[g] + (size_t)img * len;
  const long long beg = (long long)chunk * a.chunk_elems[g];
  const long long end = beg + a.chunk_elems[g] < len ? beg + a.chunk_elems[g] : len;
  u64* out = out_keys + (size_t)row * a.P;
  const int lane = threadIdx.x & 31;
  const unsigned cut = st->cut_key;
  const int transform = a.d.transform;
  const int P = a.P;
  auto emit = [&](u64 c, bool take) {  // warp-uniform call
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(&st->out_count, (unsigned)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (take) {
        const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
        D2B_BOUND(slot, P);  // exactly k_r <= P composites reach the threshold
        if (slot < (unsigned)P) out[slot] = c;
      }
    }
  };
  if (st->compact) {
    if (chunk != 0) return;
    const u64* cand = a.cand + (size_t)row * kCandCap;
    const unsigned n = st->cand_count;
    for (unsigned j0 = 0; j0 < n; j0 += kBatch * kHistThreads) {  // block-uniform trip count
      u64 c[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const unsigned j = j0 + u * kHistThreads + threadIdx.x;
        c[u] = j < n ? __ldcg(cand + j) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const unsigned j = j0 + u * kHistThreads + threadIdx.x;
        emit(c[u], j < n && c[u] >= thr);
      }
    }
    return;
  }
  for_each_elem(x, beg, end, [&](float v, long long i, bool ok) {
    u64 c = 0;
    bool take = false;
    if (ok && float_to_key(v) >= cut) {
      c = composite_of(value_key(v, transform), (unsigned)i);
      take = c >= thr;
    }
    emit(c, take);
  });
}

// SIGMOID rows: histogram of the raw logit keys (sign + exponent = one bin per binade), then the cutoff (see the
// header comment).  Long rows are SAMPLED: the first kPreSample elements of every chunk (1/32 of the row; 1/16 measured 30 us slower), one
// CTA per kPreGroup chunks.  The cutoff only has to be a lower bound of the k-th largest logit, and pass 0
// verifies it by exact count (and repeats without a cutoff otherwise), so the second full read of the class
// logits disappears.
constexpr int kPreBits = 9;
constexpr int kPreBins = 1 << kPreBits;
constexpr int kPreCopies = 16;       // [bin][lane & 15]: at most 2 lanes of a warp share a word
constexpr int kPreSample = 1024;     // sampled elements per chunk of a long row (1/64 of a 64 K-element chunk)
constexpr int kPreGroup = 8;         // chunks served by one CTA in sampled mode
constexpr int kPreMinChunks = 2;     // rows shorter than this many chunks are histogrammed in full (one CTA reading a 7-chunk row serially was the long pole of the launch)
// Launched on the kPreGroup-times coarser grid (`a` = fill_args(d, ., kPreGroup)): CTA `chunk` of a row serves the
// fine chunks [chunk * kPreGroup, (chunk + 1) * kPreGroup), so every CTA of the launch has work.
__global__ void __launch_bounds__(kHistThreads) topk_prehist(TopkArgs a) {
  __shared__ unsigned sh[kPreBins * kPreCopies];
  __shared__ unsigned s_scan[kHistThreads];
  __shared__ int s_last;
  int g, img, chunk;
  if (!locate(a, blockIdx.x, g, img, chunk)) return;
  const int row = img * a.d.G + g;
  RowState* st = a.state + row;
  if (!st->active) return;
  const long long len = a.d.row_len[g];
  const long long fine = a.chunk_elems[g] / kPreGroup;  // elements of a fine chunk (what pass 0 gives one CTA)
  const int fine_chunks = (int)((len + fine - 1) / fine);
  const bool sampled = fine_chunks >= kPreMinChunks && fine > kPreSample;
  for (int i = threadIdx.x; i < kPreBins * kPreCopies; i += kHistThreads) sh[i] = 0;
  __syncthreads();
  const float* x = a.d.scores[g] + (size_t)img * len;
  const int lane = threadIdx.x & (kPreCopies - 1);
  auto count = [&](float v, long long, bool ok) {
    D2B_BOUND((float_to_key(v) >> (32 - kPreBits)) * kPreCopies + lane, kPreBins * kPreCopies);
    if (ok) atomicAdd(&sh[(float_to_key(v) >> (32 - kPreBits)) * kPreCopies + lane], 1u);
  };
  if (sampled) {
    static_assert(kPreSample == 4 * kHistThreads, "one 16-byte vector per thread and fine chunk");
    const int c0 = chunk * kPreGroup;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (fine % 4 == 0);
    if (vec) {  // the kPreGroup samples are independent loads: all in flight at once (one round trip, not eight)
      float4 q[kPreGroup];
#pragma unroll
      for (int u = 0; u < kPreGroup; ++u) {
        const long long i = (long long)(c0 + u) * fine + 4 * threadIdx.x;
        q[u] = (c0 + u < fine_chunks && i + 3 < len) ? __ldg(reinterpret_cast<const float4*>(x + i))
                                                     : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < kPreGroup; ++u) {
        const long long i = (long long)(c0 + u) * fine + 4 * threadIdx.x;
        const bool ok = c0 + u < fine_chunks && i + 3 < len;
        count(q[u].x, i, ok); count(q[u].y, i + 1, ok); count(q[u].z, i + 2, ok); count(q[u].w, i + 3, ok);
      }
    } else {
      for (int c = c0; c < c0 + kPreGroup && c < fine_chunks; ++c) {
        const long long beg = (long long)c * fine;
        const long long end = beg + kPreSample < len ? beg + kPreSample : len;
        for_each_elem<4>(x, beg, end, count);
      }
    }
  } else {
    const long long beg = (long long)chunk * a.chunk_elems[g];
    const long long end = beg + a.chunk_elems[g] < len ? beg + a.chunk_elems[g] : len;
    for_each_elem(x, beg, end, count);
  }
  __syncthreads();
  unsigned* gh = a.prehist + (size_t)row * kBins;
  for (int b = threadIdx.x; b < kPreBins; b += kHistThreads) {
    unsigned v = 0;
#pragma unroll
    for (int l = 0; l < kPreCopies; ++l) v += sh[b * kPreCopies + ((l + b) & (kPreCopies - 1))];
    if (v) atomicAdd(gh + b, v);
  }
  __threadfence();
  __syncthreads();
  const unsigned expect = (unsigned)a.chunks[g];
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->pre_done, 1u) == expect - 1u);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA of the row: the first bin from the top at which the running count covers the target rank.  Sampled
  // rows aim 3x past the scaled rank so that the full row holds >= k_r elements above the cutoff with overwhelming
  // probability (verified exactly by pass 0).  Thread t owns the bins kPreBins-1-2t and kPreBins-2-2t; a block scan
  // of the pair sums replaces the serial walk (512 dependent shared-memory loads, ~6 us).
  static_assert(kPreBins == 2 * kHistThreads, "two bins per thread");
  const unsigned target = sampled ? (unsigned)(((u64)st->k_r * 3ull * kPreSample + fine - 1) / fine) + 16u : st->k_r;
  const int top = kPreBins - 1 - 2 * threadIdx.x;
  const unsigned h0 = __ldcg(gh + top), h1 = __ldcg(gh + top - 1);
  s_scan[threadIdx.x] = h0 + h1;
  __syncthreads();
  for (int off = 1; off < kHistThreads; off <<= 1) {  // inclusive scan from the top bins down
    const unsigned v = threadIdx.x >= off ? s_scan[threadIdx.x - off] : 0;
    __syncthreads();
    s_scan[threadIdx.x] += v;
    __syncthreads();
  }
  const unsigned incl = s_scan[threadIdx.x], excl = incl - (h0 + h1);
  int bin = -1;
  if (excl < target && target <= incl) bin = (excl + h0 >= target) ? top : top - 1;
  if (threadIdx.x == kHistThreads - 1 && incl < target) bin = 0;  // fewer elements than the target: the lowest bin
  if (bin < 0) return;
  const float c = key_to_float((unsigned)bin << (32 - kPreBits));  // lower edge of that bucket: c <= k-th largest logit
  unsigned cut = 0u;
  if (c == c && c > -80.0f && c < 8.0f) {
    const float cm = c - (0.01f * fmaxf(1.0f, fabsf(c)) + 0.01f);
    cut = float_to_key(cm);
    st->cut_hi = float_to_key(c);
  }
  st->cut_key = cut;
}

// SIGMOID rows after pass 0: ONE CTA per row runs the remaining radix passes, the collect and (k <= kFinSortMax) the
// sort of the winners.  A row's pass 0 normally leaves a complete candidate list (every element that can matter:
// a few thousand composites), and each later pass over such a list was one CTA's work anyway -- five launches of
// mostly idle grids plus collect plus sort (82 us of launch latency at N = 32).  Rows WITHOUT a complete list (no
// usable cutoff: k-th logit outside (-80, 8); or more than kCandCap elements above the cutoff) are finished by the
// same CTA re-reading the row: correct, but serial -- such rows do not occur with detector outputs.
constexpr int kFinThreads = 1024;
constexpr int kFinSortMax = 2048;
__global__ void __launch_bounds__(kFinThreads) topk_finish_kernel(TopkArgs a, u64* out_keys, int32_t* sorted_flag) {
  __shared__ unsigned sh[kBins];
  __shared__ unsigned s_warp[kFinThreads / 32];
  __shared__ unsigned s_sel[3];  // digit, need, bucket
  __shared__ unsigned s_cnt;
  __shared__ u64 s_out[kFinSortMax];
  const int row = blockIdx.x;
  const int g = row % a.d.G, img = row / a.d.G;
  RowState* st = a.state + row;
  const unsigned kr = st->k_r;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64* out = out_keys + (size_t)row * a.P;
  const bool sort_here = a.P <= kFinSortMax;
  if (tid == 0 && sorted_flag) sorted_flag[row] = sort_here ? 1 : 0;
  if (kr == 0) {
    for (int i = tid; i < a.P; i += kFinThreads) out[i] = 0ull;
    return;
  }
  const bool compact = st->compact != 0;
  const unsigned n_list = st->cand_count;
  const u64* cand = a.cand + (size_t)row * kCandCap;
  const long long len = a.d.row_len[g];
  const float* x = a.d.scores[g] + (size_t)img * len;
  const unsigned cut = st->cut_key;
  const int transform = a.d.transform;
  // f(composite) for every element of the row that can still matter
  auto for_each_comp = [&](auto f) {
    if (compact) {
      for (unsigned j0 = 0; j0 < n_list; j0 += 4 * kFinThreads) {
        u64 c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned j = j0 + u * kFinThreads + tid;
          c[u] = j < n_list ? __ldcg(cand + j) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u * kFinThreads + tid < n_list) f(c[u]);
      }
    } else {
      for (long long i = tid; i < len; i += kFinThreads) {
        const float v = __ldg(x + i);
        if (float_to_key(v) >= cut) f(composite_of(value_key(v, transform), (unsigned)i));
      }
    }
  };
  u64 prefix = st->prefix, thr = st->threshold;
  unsigned k_rem = st->k_rem;
  bool active = st->active != 0;
  for (int pass = 1; pass < kPasses && active; ++pass) {
    const int shift = c_shift[pass], bits = c_bits[pass];
    const unsigned mask = (1u << bits) - 1u;
    const int hi_shift = shift + bits;
    for (int i = tid; i < kBins; i += kFinThreads) sh[i] = 0;
    __syncthreads();
    for_each_comp([&](u64 c) {
      if ((c >> hi_shift) == prefix) atomicAdd(&sh[(unsigned)(c >> shift) & mask], 1u);
    });
    __syncthreads();
    // thread t owns bins 2047 - 2t and 2046 - 2t: inclusive scan from the top bin down
    const unsigned b0 = sh[kBins - 1 - 2 * tid], b1 = sh[kBins - 2 - 2 * tid];
    unsigned inc = b0 + b1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const unsigned incl = before + inc, excl = incl - (b0 + b1);
    if (excl < k_rem && k_rem <= incl) {
      if (excl + b0 >= k_rem) { s_sel[0] = kBins - 1 - 2 * tid; s_sel[1] = k_rem - excl; s_sel[2] = b0; }
      else { s_sel[0] = kBins - 2 - 2 * tid; s_sel[1] = k_rem - excl - b0; s_sel[2] = b1; }
    }
    __syncthreads();
    prefix = (prefix << bits) | s_sel[0];
    k_rem = s_sel[1];
    if (s_sel[2] == k_rem || pass == kPasses - 1) {  // bucket taken whole: row resolved
      thr = prefix << shift;
      active = false;
    }
    __syncthreads();
  }
  // collect the k_r composites >= thr
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  u64* dst = sort_here ? s_out : out;
  for_each_comp([&](u64 c) {
    if (c >= thr) {
      const unsigned slot = atomicAdd(&s_cnt, 1u);
      D2B_BOUND(slot, kr);
      if (slot < (unsigned)a.P) dst[slot] = c;
    }
  });
  __syncthreads();
  if (!sort_here) {  // the segment sort runs as a separate launch
    for (int i = (int)kr + tid; i < a.P; i += kFinThreads) out[i] = 0ull;
    return;
  }
  int Pe = 1;
  while (Pe < (int)kr) Pe <<= 1;
  for (int i = (int)kr + tid; i < Pe; i += kFinThreads) s_out[i] = 0ull;
  __syncthreads();
  sort_smem_desc(s_out, Pe);  // registers + shuffles (sort_tile.cuh): ~15 barriers for 1,024 keys instead of 55
  for (int i = tid; i < a.P; i += kFinThreads) out[i] = i < (int)kr ? s_out[i] : 0ull;
}

__global__ void topk_emit(TopkArgs a, const u64* keys, float* out_values, int32_t* out_indices,
                          int32_t* out_counts, int rows) {
  const int row = blockIdx.y;
  const unsigned kr = a.state[row].k_r;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.d.k) return;
  float v = 0.0f;
  int32_t idx = -1;
  if ((unsigned)j < kr) {
    const u64 c = keys[(size_t)row * a.P + j];
    v = key_to_float((uint32_t)(c >> 32));
    idx = (int32_t)(0xffffffffu - (uint32_t)c);
  }
  if (out_values) out_values[(size_t)row * a.d.k + j] = v;
  if (out_indices) out_indices[(size_t)row * a.d.k + j] = idx;
}

int fill_args(const TopkDesc& d, TopkArgs& a, int coarsen = 1) {
  a.d = d;
  int cta = 0;
  for (int g = 0; g < d.G; ++g) {
    a.chunk_elems[g] = (d.row_len[g] > (1ll << 20) ? kChunkLong : kChunk) * coarsen;
    a.chunks[g] = (int)((d.row_len[g] + a.chunk_elems[g] - 1) / a.chunk_elems[g]);
    if (a.chunks[g] < 1) a.chunks[g] = 1;
    a.cta_begin[g] = cta;
    cta += a.chunks[g] * d.rows_per_group;
  }
  for (int g = d.G; g <= D2B_MAX_LEVELS; ++g) a.cta_begin[g] = cta;
  for (int g = d.G; g < D2B_MAX_LEVELS; ++g) { a.chunks[g] = 1; a.chunk_elems[g] = kChunk; }
  a.P = topk_padded_k(d.k);
  a.prehist = nullptr;
  a.cand = nullptr;
  return cta;
}

}  // namespace

int topk_padded_k(int k) {
  int P = 1;
  while (P < k) P <<= 1;
  return P;
}

size_t topk_workspace_bytes(const TopkDesc& d) {
  const size_t rows = (size_t)d.G * d.rows_per_group;
  return ws_slice(rows * sizeof(RowState)) + ws_slice(rows * kPasses * kBins * sizeof(unsigned)) +
         ws_slice(rows * sizeof(int32_t)) +
         (d.transform == D2B_TOPK_SIGMOID
              ? ws_slice(rows * (size_t)kCandCap * sizeof(u64)) + ws_slice(rows * kBins * sizeof(unsigned))
              : 0);
}

int topk_run(const TopkDesc& d, unsigned long long* out_keys, float* out_values, int32_t* out_indices,
             int32_t* out_counts, void* ws, cudaStream_t st) {
  D2B_REQUIRE(d.G >= 1 && d.G <= D2B_MAX_LEVELS, "top-k: num_groups=%d out of range", d.G);
  D2B_REQUIRE(d.k >= 0 && d.k <= kTopkMaxK, "top-k: k=%d out of [0,%d]", d.k, kTopkMaxK);
  D2B_REQUIRE(d.rows_per_group >= 0, "top-k: negative rows_per_group");
  for (int g = 0; g < d.G; ++g)
    D2B_REQUIRE(d.row_len[g] >= 0 && d.row_len[g] < (1ll << 32) - 1, "top-k: row_len[%d] out of range", g);
  const int rows = d.G * d.rows_per_group;
  if (rows == 0 || d.k == 0) return D2B_OK;
  TopkArgs a;
  const int ctas = fill_args(d, a);
  Workspace w(ws);
  a.state = w.take<RowState>(rows);
  a.hist = w.take<unsigned>((size_t)rows * kPasses * kBins);
  int32_t* seg_len = w.take<int32_t>(rows);
  if (d.transform == D2B_TOPK_SIGMOID) {
    a.cand = w.take<u64>((size_t)rows * kCandCap);
    a.prehist = w.take<unsigned>((size_t)rows * kBins);
    D2B_CUDA(cudaMemsetAsync(a.prehist, 0, (size_t)rows * kBins * sizeof(unsigned), st));
  }
  D2B_CUDA(cudaMemsetAsync(a.hist, 0, (size_t)rows * kPasses * kBins * sizeof(unsigned), st));
  topk_init<<<(rows + 127) / 128, 128, 0, st>>>(a, rows, seg_len, out_counts);
  D2B_LAUNCH_CHECK();
  if (d.transform == D2B_TOPK_SIGMOID) {
    TopkArgs c = a;
    const int ctas_c = fill_args(d, c, kPreGroup);
    c.state = a.state; c.hist = a.hist; c.prehist = a.prehist; c.cand = a.cand;
    topk_prehist<<<ctas_c, kHistThreads, 0, st>>>(c);
    D2B_LAUNCH_CHECK();
  }
  // SIGMOID rows are normally resolved from their candidate lists by one CTA per row after pass 0: the later
  // launches use 16x coarser chunks so that the (mostly idle) grids cost ~1/16 of the CTA launches.
  TopkArgs b = a;
  int ctas_b = ctas;
  if (d.transform == D2B_TOPK_SIGMOID) {
    ctas_b = fill_args(d, b, 16);
    b.state = a.state; b.hist = a.hist; b.prehist = a.prehist; b.cand = a.cand;
  }
  int rc;
  if (d.transform == D2B_TOPK_SIGMOID) {
    // pass 0: rows with a cutoff (normally all of them) by the lean streaming scan, the rest -- and the rows whose
    // sampled cutoff failed the exact count -- by the generic kernel on the coarse grid; then ONE CTA per row finishes
    topk_scan_cut_kernel<<<ctas, kHistThreads, 0, st>>>(a);
    D2B_LAUNCH_CHECK();
    topk_hist<<<ctas_b, kHistThreads, 0, st>>>(b, 0, 0, 1);
    D2B_LAUNCH_CHECK();
    topk_hist<<<ctas_b, kHistThreads, 0, st>>>(b, 0, 1, 0);
    D2B_LAUNCH_CHECK();
    topk_finish_kernel<<<rows, kFinThreads, 0, st>>>(b, out_keys, nullptr);
    D2B_LAUNCH_CHECK();
    if (a.P > kFinSortMax) {
      rc = sort_segments_desc(out_keys, rows, a.P, seg_len, st);
      if (rc != D2B_OK) return rc;
    }
  } else {
    for (int p = 0; p < kPasses; ++p) {
      if (p == 0) topk_hist<<<ctas, kHistThreads, 0, st>>>(a, p, 0, 0);
      else topk_hist<<<ctas_b, kHistThreads, 0, st>>>(b, p, 0, 0);
      D2B_LAUNCH_CHECK();
    }
    topk_collect<<<ctas_b, kHistThreads, 0, st>>>(b, out_keys);
    D2B_LAUNCH_CHECK();
    rc = sort_segments_desc(out_keys, rows, a.P, seg_len, st);
    if (rc != D2B_OK) return rc;
  }
  if (out_values || out_indices) {
    const dim3 grid((d.k + 255) / 256, rows);
    topk_emit<<<grid, 256, 0, st>>>(a, out_keys, out_values, out_indices, out_counts, rows);
    D2B_LAUNCH_CHECK();
  }
  return D2B_OK;
}

}  // namespace d2b

using namespace d2b;

static int to_desc(const d2b_segmented_topk_params* p, TopkDesc& d) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->num_groups >= 1 && p->num_groups <= D2B_MAX_LEVELS, "num_groups=%d out of range", p->num_groups);
  D2B_REQUIRE(p->transform == D2B_TOPK_IDENTITY || p->transform == D2B_TOPK_SIGMOID, "unknown transform");
  for (int g = 0; g < D2B_MAX_LEVELS; ++g) {
    d.scores[g] = g < p->num_groups ? p->scores[g] : nullptr;
    d.row_len[g] = g < p->num_groups ? p->row_len[g] : 0;
    d.k_limit[g] = g < p->num_groups ? p->k_limit[g] : 0;
  }
  d.G = p->num_groups;
  d.rows_per_group = p->rows_per_group;
  d.k = p->k;
  d.transform = p->transform;
  return D2B_OK;
}

extern "C" size_t d2b_segmented_topk_workspace_bytes(const d2b_segmented_topk_params* p) {
  TopkDesc d;
  if (!p || to_desc(p, d) != D2B_OK) return 0;
  const size_t rows = (size_t)d.G * d.rows_per_group;
  return topk_workspace_bytes(d) + ws_slice(rows * (size_t)topk_padded_k(d.k) * sizeof(unsigned long long));
}

extern "C" int d2b_segmented_topk(const d2b_segmented_topk_params* p, void* workspace, size_t workspace_bytes,
                                  d2b_stream_t stream) {
  TopkDesc d;
  int rc = to_desc(p, d);
  if (rc != D2B_OK) return rc;
  const int rows = d.G * d.rows_per_group;
  if (rows == 0 || d.k == 0) return D2B_OK;
  for (int g = 0; g < d.G; ++g) D2B_REQUIRE(d.scores[g] != nullptr || d.row_len[g] == 0, "scores[%d] is NULL", g);
  if (workspace == nullptr || workspace_bytes < d2b_segmented_topk_workspace_bytes(p)) {
    set_last_error("segmented_topk needs %zu workspace bytes", d2b_segmented_topk_workspace_bytes(p));
    return D2B_EWORKSPACE;
  }
  Workspace w(workspace);
  unsigned long long* keys = w.take<unsigned long long>((size_t)rows * topk_padded_k(d.k));
  return topk_run(d, keys, p->out_values, p->out_indices, p->out_counts, w.base + w.off,
                  static_cast<cudaStream_t>(stream));
}
