// rpn_fused.cu -- the front of the RPN proposal stage as ONE thread-block-cluster launch:
//   tf.nn.top_k per (image, level) row (rpn_outputs.py:67-70)  ->  order (score desc, index asc)
//   -> gather + Box2BoxTransform.apply_deltas (:403-426, only the winners) -> clip_to_window (:77-80)
//   -> prune_small_boxes (:83-87), ordered.
// One cluster of kSelCluster CTAs owns one row (a row of P2 logits is 806 KB: each CTA streams 1/8 of it).
//   select   radix-select over the order-preserving 32-bit value keys, 8 bits per pass: every CTA histograms its
//            chunk into warp-private shared-memory bins, the 8 per-CTA histograms are summed by every CTA through
//            distributed shared memory (2 KB per peer), and a block scan finds the digit that holds the k-th
//            largest key.  The cluster barrier replaces the kernel boundary of a multi-launch radix select; the
//            row is read from HBM once and from L2 afterwards.  A pass whose boundary bucket is taken whole ends
//            the selection early.
//   ties     if the k-th value is tied, the lower indices win (TF TopKV2): chunks are contiguous index ranges in
//            rank order, so per-CTA tie counts give each CTA a quota, and only the one CTA with a partial quota
//            walks its chunk in index order.
//   collect  winners go to a CTA-local list (shared-memory atomics), one remote atomic per CTA reserves a range
//            of the leader's list, and the local list is copied there through DSMEM.
//   leader   bitonic sort of the <= 4096 composite keys (value key << 32 | ~index: unique, so a plain descending
//            sort realises the tie rule), then decode / clip / ordered prune into the NMS input buffers.
// Replaces 11 launches (memset, init, 6 histogram passes, collect, sort, decode) of the generic chain in
// pipelines.cu / topk.cu, which remains for k > kRpnFusedMaxK.
#include <cooperative_groups.h>

#include "nms.cuh"
#include "rpn.cuh"

namespace cg = cooperative_groups;

namespace d2b {
namespace {

constexpr int kSelThreads = 512;
constexpr int kSelCluster = 8;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kHistCopies = 8;  // two warps share one copy of the 256 bins

struct SelCtl {
  unsigned digit, need, bucket;  // result of one pass (identical in every CTA of the cluster)
  unsigned local_cnt;            // winners in this CTA's local list
  unsigned tie_cnt;              // elements equal to the k-th value in this CTA's chunk
  unsigned total;                // leader only: slots handed out in the final list
  unsigned base;                 // this CTA's range in the leader's list
  unsigned ties[kSelCluster];    // tie counts of every CTA (written remotely)
};

// Dynamic shared memory: [hist 256][whist kHistCopies*256][scan 16][ctl][local P u64][final P u64]
__device__ __forceinline__ unsigned key_of(float v) { return float_to_key(v); }

// Calls f(value, index) for the elements of [beg, end) of row x; 16-byte loads when aligned, 4 vectors in flight.
template <typename F>
__device__ __forceinline__ void scan_chunk(const float* __restrict__ x, long long beg, long long end, F f) {
  if (end <= beg) return;
  const bool vec = ((reinterpret_cast<uintptr_t>(x + beg) & 15) == 0);
  if (vec) {
    const long long nvec = (end - beg) >> 2;
    const float4* xv = reinterpret_cast<const float4*>(x + beg);
    constexpr int kV = 4;
    for (long long v0 = 0; v0 < nvec; v0 += (long long)kV * kSelThreads) {  // block-uniform trip count
      float4 q[kV];
#pragma unroll
      for (int u = 0; u < kV; ++u) {
        const long long vi = v0 + (long long)u * kSelThreads + threadIdx.x;
        q[u] = vi < nvec ? __ldg(xv + vi) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < kV; ++u) {
        const long long vi = v0 + (long long)u * kSelThreads + threadIdx.x;
        const bool ok = vi < nvec;
        const long long i = beg + 4 * vi;
        f(q[u].x, i, ok); f(q[u].y, i + 1, ok); f(q[u].z, i + 2, ok); f(q[u].w, i + 3, ok);
      }
    }
    const long long t0 = beg + 4 * nvec;
    for (long long i0 = t0; i0 < end; i0 += kSelThreads) {  // < 4 leftover elements: one uniform trip
      const long long i = i0 + threadIdx.x;
      const bool ok = i < end;
      f(ok ? __ldg(x + i) : 0.0f, i, ok);
    }
  } else {
    for (long long i0 = beg; i0 < end; i0 += kSelThreads) {
      const long long i = i0 + threadIdx.x;
      const bool ok = i < end;
      f(ok ? __ldg(x + i) : 0.0f, i, ok);
    }
  }
}

// Exclusive scan of `v` over the block in thread order; `total` gets the block sum.  s_w: kSelWarps ints.
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* s_w, unsigned& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  unsigned before = 0;
  total = 0;
#pragma unroll
  for (int w = 0; w < kSelWarps; ++w) {
    const unsigned c = s_w[w];
    if (w < warp) before += c;
    total += c;
  }
  __syncthreads();
  return before + inc - v;
}

__global__ void __launch_bounds__(kSelThreads) rpn_select_kernel(RpnArgs a, float4* seg_boxes, float* seg_scores,
                                                                  int32_t* seg_count, int32_t* img_done,
                                                                  u64* nms_in_total) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  unsigned* hist = reinterpret_cast<unsigned*>(s_raw);            // [256] merged histogram of this CTA (peers read it)
  unsigned* whist = hist + 256;                                    // [kHistCopies][256]
  unsigned* s_scan = whist + kHistCopies * 256;                    // [32]
  SelCtl* ctl = reinterpret_cast<SelCtl*>(s_scan + 32);
  u64* s_local = reinterpret_cast<u64*>(s_raw + 256 * 4 + kHistCopies * 256 * 4 + 32 * 4 + 128);  // [P]
  u64* s_final = s_local + a.P;                                                                    // [P]

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int cid = blockIdx.x / kSelCluster;  // level-major: the long P2 rows are scheduled first
  const int l = cid / a.N, n = cid - l * a.N;
  const int row = n * a.L + l;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long len = a.hwa[l];
  const float* x = a.logits[l] + (size_t)n * len;
  const unsigned kr = (unsigned)(len < (long long)a.k ? len : (long long)a.k);

  if (tid == 0) {
    ctl->local_cnt = 0; ctl->tie_cnt = 0; ctl->total = 0; ctl->base = 0;
    if (rank == 0 && l == 0 && img_done) img_done[n] = 0;
  }
  cluster.sync();  // every CTA of the cluster runs and has initialised its control block
  if (kr == 0) {
    if (rank == 0 && tid == 0) seg_count[row] = 0;
    return;
  }
  // contiguous chunk of the row, multiple of 4 elements so that 16-byte loads stay aligned
  long long per = (len + kSelCluster - 1) / kSelCluster;
  per = (per + 3) & ~3ll;
  const long long beg = per * rank < len ? per * rank : len;
  const long long end = beg + per < len ? beg + per : len;

  // ------------------------------------------------------------------ select
  unsigned prefix = 0;      // resolved high bits of the k-th largest key
  unsigned k_rem = kr;      // winners still to be found inside the prefix bucket
  bool all = (kr == (unsigned)len);
  bool ties = false;        // k-th value tied: only `k_rem` of the elements equal to `prefix` are taken
  unsigned cut = 0;         // take key > cut (plus the tie quota)
  if (!all) {
    bool resolved = false;
    int pass = 0;
    for (; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < kHistCopies * 256; i += kSelThreads) whist[i] = 0;
      __syncthreads();
      unsigned* my = whist + (warp >> 1) * 256;
      if (pass == 0) {
        scan_chunk(x, beg, end, [&](float v, long long, bool ok) {
          if (ok) atomicAdd(my + (key_of(v) >> 24), 1u);
        });
      } else {
        const int hs = shift + 8;
        scan_chunk(x, beg, end, [&](float v, long long, bool ok) {
          const unsigned key = key_of(v);
          if (ok && (key >> hs) == prefix) atomicAdd(my + ((key >> shift) & 255u), 1u);
        });
      }
      __syncthreads();
      if (pass > 0) cluster.barrier_wait();  // every peer has finished reading `hist` of the previous pass
      if (tid < 256) {
        unsigned v = 0;
#pragma unroll
        for (int c = 0; c < kHistCopies; ++c) v += whist[c * 256 + tid];
        hist[tid] = v;
      }
      cluster.sync();  // the 8 per-CTA histograms are visible cluster-wide
      // every CTA sums the peers' histograms (thread t owns bin 255 - t) and scans from the top bin down
      unsigned tot = 0;
      if (tid < 256) {
#pragma unroll
        for (unsigned r = 0; r < kSelCluster; ++r) tot += cluster.map_shared_rank(hist, r)[255 - tid];
      }
      unsigned total;
      const unsigned excl = block_excl_scan(tot, s_scan, total);
      if (tid < 256 && excl < k_rem && k_rem <= excl + tot) {
        ctl->digit = 255u - (unsigned)tid;
        ctl->need = k_rem - excl;
        ctl->bucket = tot;
      }
      __syncthreads();
      cluster.barrier_arrive();  // done with the peers' `hist`
      prefix = (prefix << 8) | ctl->digit;
      k_rem = ctl->need;
      const unsigned bucket = ctl->bucket;
      __syncthreads();  // ctl fields are rewritten by the next pass
      if (bucket == k_rem) {  // boundary bucket taken whole: every key >= prefix << shift wins
        const unsigned thr = prefix << shift;
        if (thr == 0u) all = true; else cut = thr - 1u;
        resolved = true;
        break;
      }
    }
    cluster.barrier_wait();  // pairs with the last barrier_arrive
    if (!resolved) {  // all 32 value bits fixed and the bucket holds more than k_rem equal keys
      ties = true;
      cut = prefix;
    }
  }

  // ------------------------------------------------------------------ collect
  auto append_local = [&](u64 c, bool take) {  // warp-uniform call
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(&ctl->local_cnt, (unsigned)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (take) {
        const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < (unsigned)a.P) s_local[slot] = c;
      }
    }
  };
  auto flush_local = [&]() {  // reserve a range of the leader's list and copy the local winners there
    __syncthreads();
    if (tid == 0) {
      const unsigned c = ctl->local_cnt;
      ctl->base = c ? atomicAdd(&cluster.map_shared_rank(ctl, 0)->total, c) : 0u;
    }
    __syncthreads();
    const unsigned c = ctl->local_cnt, base = ctl->base;
    u64* dst = cluster.map_shared_rank(s_final, 0);
    for (unsigned i = tid; i < c; i += kSelThreads)
      if (base + i < (unsigned)a.P) dst[base + i] = s_local[i];
    __syncthreads();
    if (tid == 0) ctl->local_cnt = 0;
    __syncthreads();
  };
  unsigned my_ties = 0;
  scan_chunk(x, beg, end, [&](float v, long long i, bool ok) {
    const unsigned key = key_of(v);
    const bool take = ok && (all || key > cut);
    if (ties) my_ties += (ok && key == cut) ? 1u : 0u;
    append_local(((u64)key << 32) | (u64)(0xffffffffu - (unsigned)i), take);
  });
  flush_local();
  if (ties) {  // block-uniform (identical in the whole cluster)
    // per-CTA tie counts -> every CTA's table; chunks are contiguous index ranges in rank order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_ties += __shfl_xor_sync(0xffffffffu, my_ties, o);
    if (lane == 0 && my_ties) atomicAdd(&ctl->tie_cnt, my_ties);
    __syncthreads();
    if (tid < kSelCluster) cluster.map_shared_rank(ctl, tid)->ties[rank] = ctl->tie_cnt;
    cluster.sync();
    unsigned before = 0;
    for (unsigned r = 0; r < rank; ++r) before += ctl->ties[r];
    const unsigned mine = ctl->ties[rank];
    const unsigned quota = before >= k_rem ? 0u : (k_rem - before < mine ? k_rem - before : mine);
    if (quota == mine) {  // all of this chunk's ties (possibly none)
      if (mine)
        scan_chunk(x, beg, end, [&](float v, long long i, bool ok) {
          const unsigned key = key_of(v);
          append_local(((u64)key << 32) | (u64)(0xffffffffu - (unsigned)i), ok && key == cut);
        });
    } else if (quota > 0) {
      // the one CTA with a partial quota: its `quota` lowest-index ties, found by walking the chunk in index order
      // (thread t owns 4 consecutive elements per trip; an ordered block scan gives every tie its ordinal)
      unsigned running = 0;
      for (long long i0 = beg; i0 < end && running < quota; i0 += 4ll * kSelThreads) {
        const long long i = i0 + 4ll * tid;
        unsigned keyv[4];
        unsigned cnt = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool ok = i + c < end;
          keyv[c] = ok ? key_of(__ldg(x + i + c)) : 0u;
          if (!ok || keyv[c] != cut) keyv[c] = cut + 1u;  // marks "not a tie" (cut + 1 != cut even on wrap-around)
          else ++cnt;
        }
        unsigned total;
        unsigned ord = running + block_excl_scan(cnt, s_scan, total);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (keyv[c] == cut) {
            if (ord < quota) {
              const unsigned slot = atomicAdd(&ctl->local_cnt, 1u);
              if (slot < (unsigned)a.P) s_local[slot] = ((u64)cut << 32) | (u64)(0xffffffffu - (unsigned)(i + c));
            }
            ++ord;
          }
        }
        running += total;
      }
    }
    flush_local();
  }
  cluster.sync();  // the leader's list is complete; no CTA touches a peer's shared memory after this point
  if (rank != 0) return;

  // ------------------------------------------------------------------ leader: sort, decode, clip, prune
  int Pe = 1;
  while (Pe < (int)kr) Pe <<= 1;
  for (int i = (int)kr + tid; i < Pe; i += kSelThreads) s_final[i] = 0ull;
  __syncthreads();
  for (int k = 2; k <= Pe; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = tid; p < Pe / 2; p += kSelThreads) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const bool desc = ((i & k) == 0);
        const u64 va = s_final[i], vb = s_final[i | j];
        if (desc ? (va < vb) : (va > vb)) { s_final[i] = vb; s_final[i | j] = va; }
      }
      __syncthreads();
    }
  const float h = (float)a.shapes[2 * n], w = (float)a.shapes[2 * n + 1];
  const size_t rbase = (size_t)n * len;
  int* s_warp = reinterpret_cast<int*>(s_scan);
  int base = 0;
  for (int j0 = 0; j0 < (int)kr; j0 += kSelThreads) {
    const int j = j0 + tid;
    bool ok = false;
    float4 box = make_float4(0, 0, 0, 0);
    float score = 0.0f;
    if (j < (int)kr) {
      const unsigned idx = key_index(s_final[j]);
      score = __ldg(a.logits[l] + rbase + idx);
      if (a.proposals[l]) box = __ldg(a.proposals[l] + rbase + idx);
      else box = d2b_decode(__ldg(a.deltas[l] + rbase + idx), a.anchors[l].at(idx), a.w[0], a.w[1], a.w[2], a.w[3], a.clampv);
      box = d2b_clip(box, h, w);  // rpn_outputs.py:77-80
      ok = true;
      if (a.min_len > 0.0f) {     // prune_small_boxes, :83-87
        const float bh = box.z - box.x, bw = box.w - box.y;
        ok = (bw >= a.min_len) && (bh >= a.min_len);
      }
    }
    const int slot = block_compact<kSelThreads>(ok, base, s_warp);
    if (slot >= 0) {
      seg_boxes[(size_t)row * a.k + slot] = box;
      seg_scores[(size_t)row * a.k + slot] = score;
    }
  }
  if (tid == 0) {
    seg_count[row] = base;
    if (nms_in_total) atomicAdd(nms_in_total, (u64)base);
  }
}

// One CTA per (image, level) segment sweeps its suppression mask (nms.cuh); the CTA that finishes an image's last
// segment then merges the L survivor lists into the image's top `post` proposals (rpn_outputs.py:101-114,
// rpn_merge_rank_body).  img_done [N] was zeroed by rpn_select_kernel.
__global__ void __launch_bounds__(kColSweepThreads) rpn_sweep_merge_kernel(
    RpnArgs a, const int32_t* seg_count, int W, const u64* mask, const float4* seg_boxes, const float* seg_scores,
    int32_t* keep, int32_t* num_keep, int32_t* img_done, uint32_t* gkeys, int use_smem, float4* out_boxes,
    float* out_logits, uint8_t* out_valid, int32_t* out_num) {
  extern __shared__ uint32_t s_merge_keys[];
  __shared__ int s_last;
  const int seg = blockIdx.x;
  const int n = seg / a.L;
  const int cnt = min(seg_count[seg], a.k);
  const int kept = nms_sweep_columns(cnt, W, a.post, mask + (size_t)seg * W * 64 * W, keep + (size_t)seg * a.post);
  if (threadIdx.x == 0) {
    num_keep[seg] = kept;
    __threadfence();  // this segment's keep list and count are visible before the arrival is
    s_last = (atomicAdd(img_done + n, 1) == a.L - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  rpn_merge_rank_body(a, n, seg_boxes, seg_scores, keep, num_keep, gkeys, s_merge_keys, use_smem, out_boxes,
                      out_logits, out_valid, out_num);
}

size_t select_smem_bytes(int P) {
  return 256 * 4 + kHistCopies * 256 * 4 + 32 * 4 + 128 + 2 * (size_t)P * sizeof(u64);
}

}  // namespace

int rpn_select_fused(const RpnArgs& a, float4* seg_boxes, float* seg_scores, int32_t* seg_count, int32_t* img_done,
                     unsigned long long* nms_in_total, cudaStream_t st) {
  static_assert(sizeof(SelCtl) <= 128, "control block must fit its slot");
  D2B_REQUIRE(a.k <= kRpnFusedMaxK, "fused proposal stage: k=%d > %d", a.k, kRpnFusedMaxK);
  const int rows = a.L * a.N;
  if (rows == 0) return D2B_OK;
  const size_t smem = select_smem_bytes(a.P);
  if (smem > 48 * 1024)
    D2B_CUDA(cudaFuncSetAttribute(rpn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)rows * kSelCluster, 1, 1);
  cfg.blockDim = dim3(kSelThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kSelCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  D2B_CUDA(cudaLaunchKernelEx(&cfg, rpn_select_kernel, a, seg_boxes, seg_scores, seg_count, img_done,
                              reinterpret_cast<u64*>(nms_in_total)));
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}

int rpn_sweep_merge_fused(const RpnArgs& a, const int32_t* seg_count, const unsigned long long* mask,
                          const float4* seg_boxes, const float* seg_scores, int32_t* keep, int32_t* num_keep,
                          int32_t* img_done, uint32_t* gkeys, float4* out_boxes, float* out_logits, uint8_t* out_valid,
                          int32_t* out_num, cudaStream_t st) {
  const int rows = a.L * a.N;
  if (rows == 0) return D2B_OK;
  const int W = (a.k + 63) / 64;
  D2B_REQUIRE(W <= kColSweepMaxW, "fused sweep: k=%d too large", a.k);
  const size_t merge_smem = (size_t)a.P2 * sizeof(uint32_t);
  const int in_smem = merge_smem <= 160 * 1024;
  if (in_smem && merge_smem > 48 * 1024)
    D2B_CUDA(cudaFuncSetAttribute(rpn_sweep_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem));
  rpn_sweep_merge_kernel<<<rows, kColSweepThreads, in_smem ? merge_smem : 0, st>>>(
      a, seg_count, W, mask, seg_boxes, seg_scores, keep, num_keep, img_done, gkeys, in_smem, out_boxes, out_logits,
      out_valid, out_num);
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}

}  // namespace d2b
