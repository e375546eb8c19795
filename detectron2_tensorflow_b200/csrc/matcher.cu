// matcher.cu -- training-side neighbours of the proposal path:
//   pairwise_iou                       lib/structures/box_list_ops.py:295-334
//   Matcher (+ low-quality matches)    lib/modeling/matcher.py:57-174
//   RPNOutputs._get_ground_truth       lib/modeling/proposal_generator/rpn_outputs.py:245-304
//   ROIHeads.label_and_sample_proposals (matching part)  lib/modeling/roi_heads/roi_heads.py:100-165
//   Box2BoxTransform.get_deltas        lib/modeling/box_regression.py:38-74
//   single-level decode+clip+filter    rpn_outputs.py:403-426, 77-86
//
// The reference materialises three [G, P] IoU matrices per image (valid / crowd / difficult GT x
// 268 K anchors), reduces them along both axes and stitches labels with tf.where/dynamic_stitch.
// Here the matrices are never written: the GT lists live in shared memory, each thread owns one
// prediction box, and the only HBM traffic is the box read and the 32 B/box of results, so the
// label kernel is bound by its output write.  Low-quality matches need the per-GT maximum over
// ALL predictions first, so that case runs a first pass that reduces per-GT maxima
// (warp REDUX -> smem atomicMax -> global atomicMax on order-preserving keys).
// Every IoU is evaluated by the same inlined function in both passes => the equality test
// `iou == highest_quality_foreach_gt` (matcher.py:157-160) is exact.
#include "kernels.cuh"

namespace d2b {
namespace {

constexpr int kThreads = 256;
constexpr int kPredsPerThread = 4;  // first pass only

// box_list_ops.py:295-334, iou_type='iou'.  g = (y1,x1,y2,x2) of the GT row, ga its area, p / pa the
// prediction column.  where(unions == 0, 0, inter / unions); pairs with an empty intersection and a
// finite union are +-0 by either branch and skip the division.
__device__ __forceinline__ float pair_iou(const float4 g, const float ga, const float4 p, const float pa) {
  const float ih = fmaxf(0.0f, fminf(g.z, p.z) - fmaxf(g.x, p.x));
  const float iw = fmaxf(0.0f, fminf(g.w, p.w) - fmaxf(g.y, p.y));
  const float inter = ih * iw;
  float u = ga + pa;
  u = u - inter;
  if (inter == 0.0f && fabsf(u) < __int_as_float(0x7f800000)) return 0.0f;
  return (u == 0.0f) ? 0.0f : inter / u;
}
__device__ __forceinline__ float box_area(const float4 b) { return (b.z - b.x) * (b.w - b.y); }

// Box2BoxTransform.get_deltas, box_regression.py:38-74 (src p -> target t), one rounding per written op
__device__ __forceinline__ float4 encode_deltas(const float4 p, const float4 t, float wy, float wx, float wh, float ww) {
  const float sh = p.z - p.x, sw = p.w - p.y;
  float scy = 0.5f * sh; scy = p.x + scy;
  float scx = 0.5f * sw; scx = p.y + scx;
  const float th = t.z - t.x, tw = t.w - t.y;
  float tcy = 0.5f * th; tcy = t.x + tcy;
  float tcx = 0.5f * tw; tcx = t.y + tcx;
  float dy = tcy - scy; dy = wy * dy; dy = dy / sh;
  float dx = tcx - scx; dx = wx * dx; dx = dx / sw;
  float dh = th / sh; dh = d2b_logf(dh); dh = wh * dh;
  float dw = tw / sw; dw = d2b_logf(dw); dw = ww * dw;
  return make_float4(dy, dx, dh, dw);
}
__global__ void get_deltas_kernel(const float4* src, const float4* tgt, long long n, float wy, float wx, float wh,
                                  float ww, float4* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = encode_deltas(__ldg(src + i), __ldg(tgt + i), wy, wx, wh, ww);
}

// box_list_ops.py:335-371 (iou_type giou / diou / ciou), the reference's op order; boxes are (y1, x1, y2, x2)
__device__ __forceinline__ float pair_iou_variant(const float4 a, const float4 b, int type) {
  const float ih = fmaxf(0.0f, fminf(a.z, b.z) - fmaxf(a.x, b.x));
  const float iw = fmaxf(0.0f, fminf(a.w, b.w) - fmaxf(a.y, b.y));
  const float inter = ih * iw;
  const float h1 = a.z - a.x, w1 = a.w - a.y, h2 = b.z - b.x, w2 = b.w - b.y;
  float u = h1 * w1 + h2 * w2;
  u = u - inter;
  const float iou = (u == 0.0f) ? 0.0f : inter / u;
  const float dy = fmaxf(a.z, b.z) - fminf(a.x, b.x);
  const float dx = fmaxf(a.w, b.w) - fminf(a.y, b.y);
  if (type == D2B_GIOU) {
    const float convex = fmaxf(0.0f, dy) * iw;  // (sic) :344
    float t = convex - u;
    t = (convex == 0.0f) ? 0.0f : t / convex;
    return iou - t;
  }
  const float diag2 = dy * dy + dx * dx;
  const float cx = (a.y + a.w) / 2.0f - (b.y + b.w) / 2.0f;
  const float cy = (a.x + a.z) / 2.0f - (b.x + b.z) / 2.0f;
  const float cd2 = cx * cx + cy * cy;
  const float diou = iou - ((diag2 == 0.0f) ? 0.0f : cd2 / diag2);
  if (type == D2B_DIOU) return diou;
  const float d = atanf(w1 / h1) - atanf(w2 / h2);
  const float v = 0.40528473456935109f * (d * d);  // 4 / pi^2
  float den = 1.0f - iou;
  den = den + v;
  return diou - (v / den) * v;
}
__global__ void pairwise_iou_variant_kernel(const float4* b1, const float4* b2, long long n2, int type, float* out) {
  const long long i = blockIdx.y;
  const float4 g = __ldg(b1 + i);
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += (long long)gridDim.x * blockDim.x)
    __stcs(out + i * n2 + j, pair_iou_variant(g, __ldg(b2 + j), type));
}

__global__ void pairwise_iou_kernel(const float4* b1, long long n1, const float4* b2, long long n2, float* out) {
  // one row of boxes1 per blockIdx.y, columns strided over the block: coalesced 4 B/pair stores
  const long long i = blockIdx.y;
  const float4 g = __ldg(b1 + i);
  const float ga = box_area(g);
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += (long long)gridDim.x * blockDim.x) {
    const float4 p = __ldg(b2 + j);
    __stcs(out + i * n2 + j, pair_iou(g, ga, p, box_area(p)));
  }
}

struct LabelArgs {
  const float4* pred;
  int pred_shared;
  const int32_t* pred_counts;
  int N, P, G;
  const float4* gt;
  const uint8_t* valid;
  const uint8_t* crowd;
  const uint8_t* difficult;
  float thr[D2B_MATCH_MAX_THRESHOLDS];
  int nt;
  int labels[D2B_MATCH_MAX_THRESHOLDS + 1];
  int allow_lq;
  float boundary;
  const int32_t* shapes;
  int deltas_on;
  float wy, wx, wh, ww;
  long long* out_matches;
  long long* out_labels;
  float4* out_deltas;
  unsigned* gtmax;  // [N, G] keys (workspace), only when allow_lq
};

// Order-preserving compaction of one image's GT flags into shared lists (boolean_mask order).
// s_box[0..nv) valid, s_box[G..G+nc) crowd, s_box[2G..2G+nd) difficult; areas alongside.
struct GtLists {
  int nv, nc, nd;
};
__device__ __forceinline__ GtLists load_gt(const LabelArgs& a, int img, float4* s_box, float* s_area, int* s_cnt) {
  const int G = a.G;
  if (threadIdx.x < 32) {
    // warp 0 scans the flags in order (G <= 1024: at most 32 rounds)
    const int lane = threadIdx.x;
    int nv = 0, nc = 0, nd = 0;
    for (int base = 0; base < G; base += 32) {
      const int g = base + lane;
      bool v = false, c = false, d = false;
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < G) {
        const size_t o = (size_t)img * G + g;
        v = a.valid[o] != 0;
        c = a.crowd ? a.crowd[o] != 0 : false;
        d = a.difficult ? a.difficult[o] != 0 : false;
        b = __ldg(a.gt + o);
      }
      const bool vv = v && !c && !d;
      const unsigned mv = __ballot_sync(0xffffffffu, vv), mc = __ballot_sync(0xffffffffu, c),
                     md = __ballot_sync(0xffffffffu, d);
      const unsigned below = (1u << lane) - 1u;
      if (vv) { const int k = nv + __popc(mv & below); s_box[k] = b; s_area[k] = box_area(b); }
      if (c) { const int k = G + nc + __popc(mc & below); s_box[k] = b; s_area[k] = box_area(b); }
      if (d) { const int k = 2 * G + nd + __popc(md & below); s_box[k] = b; s_area[k] = box_area(b); }
      nv += __popc(mv); nc += __popc(mc); nd += __popc(md);
    }
    if (lane == 0) { s_cnt[0] = nv; s_cnt[1] = nc; s_cnt[2] = nd; }
  }
  __syncthreads();
  GtLists r;
  r.nv = s_cnt[0]; r.nc = s_cnt[1]; r.nd = s_cnt[2];
  return r;
}

// Spatial culling.  A GT box whose intersection with the bounding box of ALL prediction boxes of this
// CTA is empty has IoU exactly 0 (inter == 0, finite union) with every one of them, so it can only
// matter through "everything ties at 0" cases, which the callers handle in closed form.  RPN anchors are
// laid out (y, x, a)-major, so 256 consecutive anchors cover a thin strip of the image and most GT boxes
// drop out.  Lists keep GT order (argmax ties go to the lower index).  Boxes with NaN coordinates are
// outside this argument (the reference's result for them is NaN-comparison noise).
struct Cull {
  int nv, nc, nd;  // lengths of the filtered valid / crowd / difficult lists
};
__device__ __forceinline__ Cull cull_gt(const GtLists L, int G, const float4* s_box, const float* s_area, float y1,
                                        float x1, float y2, float x2, bool any_live, int* s_sel, unsigned* s_bb,
                                        int* s_cnt) {
  // s_bb: [0]=min y1, [1]=min x1 (as keys), [2]=max y2, [3]=max x2 over the CTA's live predictions
  if (threadIdx.x == 0) { s_bb[0] = 0xffffffffu; s_bb[1] = 0xffffffffu; s_bb[2] = 0u; s_bb[3] = 0u; }
  __syncthreads();
  unsigned k0 = any_live ? float_to_key(y1) : 0xffffffffu, k1 = any_live ? float_to_key(x1) : 0xffffffffu;
  unsigned k2 = any_live ? float_to_key(y2) : 0u, k3 = any_live ? float_to_key(x2) : 0u;
  k0 = __reduce_min_sync(0xffffffffu, k0); k1 = __reduce_min_sync(0xffffffffu, k1);
  k2 = __reduce_max_sync(0xffffffffu, k2); k3 = __reduce_max_sync(0xffffffffu, k3);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(s_bb + 0, k0); atomicMin(s_bb + 1, k1); atomicMax(s_bb + 2, k2); atomicMax(s_bb + 3, k3);
  }
  __syncthreads();
  const float by1 = key_to_float(s_bb[0]), bx1 = key_to_float(s_bb[1]);
  const float by2 = key_to_float(s_bb[2]), bx2 = key_to_float(s_bb[3]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 3) {  // one warp per list, ordered compaction
    const int n = warp == 0 ? L.nv : (warp == 1 ? L.nc : L.nd);
    const int base = warp * G;
    int cnt = 0;
    for (int c = 0; c < n; c += 32) {
      const int g = c + lane;
      bool hit = false;
      if (g < n) {
        const float4 b = s_box[base + g];
        const bool apart = (b.x >= by2) || (b.z <= by1) || (b.y >= bx2) || (b.w <= bx1);
        const bool finite = fabsf(s_area[base + g]) < __int_as_float(0x7f800000);
        hit = !(apart && finite);
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) s_sel[base + cnt + __popc(m & ((1u << lane) - 1u))] = g;
      cnt += __popc(m);
    }
    if (lane == 0) s_cnt[4 + warp] = cnt;
  }
  __syncthreads();
  Cull r;
  r.nv = s_cnt[4]; r.nc = s_cnt[5]; r.nd = s_cnt[6];
  return r;
}

// Pass 1 (allow_low_quality_matches only): highest_quality_foreach_gt = reduce_max(q, axis=1), matcher.py:152.
__global__ void __launch_bounds__(kThreads) label_gtmax_kernel(const LabelArgs a) {
  extern __shared__ float4 smem[];
  const int G = a.G;
  float4* s_box = smem;                                           // [3G]
  float* s_area = reinterpret_cast<float*>(smem + 3 * G);         // [3G]
  unsigned* s_max = reinterpret_cast<unsigned*>(s_area + 3 * G);  // [G]
  int* s_sel = reinterpret_cast<int*>(s_max + G);                 // [3G]
  int* s_cnt = s_sel + 3 * G;                                     // [8]
  unsigned* s_bb = reinterpret_cast<unsigned*>(s_cnt + 8);        // [4]
  const int img = blockIdx.y;
  for (int g = threadIdx.x; g < G; g += kThreads) s_max[g] = 0u;
  const GtLists L = load_gt(a, img, s_box, s_area, s_cnt);
  if (L.nv == 0) return;
  const int cnt = a.pred_counts ? min(a.pred_counts[img], a.P) : a.P;
  const float4* pb = a.pred + (a.pred_shared ? 0 : (size_t)img * a.P);
  const int j0 = blockIdx.x * (kThreads * kPredsPerThread) + threadIdx.x;
  float4 p[kPredsPerThread];
  float pa[kPredsPerThread];
  bool live[kPredsPerThread];
  bool any = false;
  const float inf = __int_as_float(0x7f800000);
  float y1 = inf, x1 = inf, y2 = -inf, x2 = -inf;
#pragma unroll
  for (int r = 0; r < kPredsPerThread; ++r) {
    const int j = j0 + r * kThreads;
    live[r] = j < cnt;
    p[r] = live[r] ? __ldg(pb + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    pa[r] = box_area(p[r]);
    any |= live[r];
    if (live[r]) { y1 = fminf(y1, p[r].x); x1 = fminf(x1, p[r].y); y2 = fmaxf(y2, p[r].z); x2 = fmaxf(x2, p[r].w); }
  }
  const Cull Cn = cull_gt(L, G, s_box, s_area, y1, x1, y2, x2, any, s_sel, s_bb, s_cnt);
  const bool warp_any = __any_sync(0xffffffffu, any);
  if (warp_any) {
    for (int q = 0; q < Cn.nv; ++q) {
      const int g = s_sel[q];
      const float4 gb = s_box[g];
      const float ga = s_area[g];
      unsigned k = 0u;
#pragma unroll
      for (int r = 0; r < kPredsPerThread; ++r)
        if (live[r]) k = max(k, float_to_key(pair_iou(gb, ga, p[r], pa[r])));
      k = __reduce_max_sync(0xffffffffu, k);
      if ((threadIdx.x & 31) == 0 && k > s_max[g]) atomicMax(s_max + g, k);
    }
  }
  __syncthreads();
  // culled GT contribute exactly 0 here; the label kernel floors every maximum at key(0) instead
  for (int q = threadIdx.x; q < Cn.nv; q += kThreads) {
    const int g = s_sel[q];
    const unsigned k = s_max[g];
    if (k > 0x80000000u) atomicMax(a.gtmax + (size_t)img * G + g, k);
  }
}

// Pass 2: per prediction column argmax/max over the valid GT, threshold labels, low-quality override,
// crowd / difficult ignore, inside_window, get_deltas -> 32 B/box of output.
__global__ void __launch_bounds__(kThreads) label_kernel(const LabelArgs a) {
  extern __shared__ float4 smem[];
  const int G = a.G;
  float4* s_box = smem;
  float* s_area = reinterpret_cast<float*>(smem + 3 * G);
  unsigned* s_max = reinterpret_cast<unsigned*>(s_area + 3 * G);
  int* s_sel = reinterpret_cast<int*>(s_max + G);
  int* s_cnt = s_sel + 3 * G;
  unsigned* s_bb = reinterpret_cast<unsigned*>(s_cnt + 8);
  const int img = blockIdx.y;
  const GtLists L = load_gt(a, img, s_box, s_area, s_cnt);
  const int cnt = a.pred_counts ? min(a.pred_counts[img], a.P) : a.P;
  const unsigned key0 = 0x80000000u;  // float_to_key(0.0f)
  if (a.allow_lq) {
    // every per-GT maximum is >= 0 once there is one live prediction (IoU is never negative)
    if (threadIdx.x == 0) s_cnt[7] = 0;
    __syncthreads();
    bool zero = false;
    for (int g = threadIdx.x; g < L.nv; g += kThreads) {
      const unsigned k = max(a.gtmax[(size_t)img * G + g], cnt > 0 ? key0 : 0u);
      s_max[g] = k;
      zero |= (k == key0);
    }
    if (zero) s_cnt[7] = 1;  // some GT overlaps nothing: every prediction ties with its maximum 0
  }
  const int j = blockIdx.x * kThreads + threadIdx.x;
  const size_t o = (size_t)img * a.P + j;
  const bool live = j < cnt;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) p = __ldg(a.pred + (a.pred_shared ? (size_t)j : o));
  const Cull Cn = cull_gt(L, G, s_box, s_area, p.x, p.y, p.z, p.w, live, s_sel, s_bb, s_cnt);
  if (j >= a.P) return;
  if (!live) {
    a.out_matches[o] = 0;
    a.out_labels[o] = -1;
    if (a.out_deltas) a.out_deltas[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float pa = box_area(p);
  long long match = 0;
  int label = 0;
  if (L.nv > 0) {
    // tf.argmax / tf.reduce_max over axis 0: first maximum wins (matcher.py:93-94).  Culled GT are exact zeros,
    // so the scan starts from (0, index 0) and only a strictly larger candidate moves it.
    float best = 0.0f;
    int bi = 0;
    bool lowq = a.allow_lq && s_cnt[7] != 0;
    for (int q = 0; q < Cn.nv; ++q) {
      const int g = s_sel[q];
      const float v = pair_iou(s_box[g], s_area[g], p, pa);
      if (v > best) { best = v; bi = g; }
      if (a.allow_lq) {
        const unsigned k = float_to_key(v);
        lowq |= (k != 0u && k == s_max[g]);
      }
    }
    match = bi;
    for (int t = 0; t <= a.nt; ++t) {  // matcher.py:97-107 (intervals [low, high))
      const float low = (t == 0) ? __int_as_float(0xff800000) : a.thr[t - 1];
      const float high = (t == a.nt) ? __int_as_float(0x7f800000) : a.thr[t];
      if (best >= low && best < high) label = a.labels[t];
    }
    if (lowq) label = 1;  // dynamic_stitch: the low-quality indices come last and win (matcher.py:109-115)
  }
  if (a.crowd) {  // matcher.py:124-134
    float mx = 0.0f;
    for (int q = 0; q < Cn.nc; ++q) {
      const int g = G + s_sel[G + q];
      mx = fmaxf(mx, pair_iou(s_box[g], s_area[g], p, pa));
    }
    if (label == 0 && L.nc > 0 && mx > 1e-3f) label = -1;
  }
  if (a.difficult) {  // matcher.py:136-148
    float mx = 0.0f;
    for (int q = 0; q < Cn.nd; ++q) {
      const int g = 2 * G + s_sel[2 * G + q];
      mx = fmaxf(mx, pair_iou(s_box[g], s_area[g], p, pa));
    }
    if (label == 0 && L.nd > 0 && mx > a.thr[0]) label = -1;
  }
  if (a.boundary >= 0.0f) {  // rpn_outputs.py:268-278, box_list_ops.py:150-161
    const float wy1 = 0.0f - a.boundary, wx1 = 0.0f - a.boundary;
    const float wy2 = (float)a.shapes[2 * img] + a.boundary, wx2 = (float)a.shapes[2 * img + 1] + a.boundary;
    if ((p.x < wy1) || (p.y < wx1) || (p.z > wy2) || (p.w > wx2)) label = -1;
  }
  a.out_matches[o] = match;
  a.out_labels[o] = label;
  if (a.out_deltas) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (label > 0 && L.nv > 0) {  // box_regression.py:38-74
      d = encode_deltas(p, s_box[(int)match], a.wy, a.wx, a.wh, a.ww);
    }
    a.out_deltas[o] = d;
  }
}

// rpn_outputs.py:403-426 + 77-86 on one level: decode all anchors, clip to the image, flag small boxes
__global__ void decode_clip_filter_kernel(const float4* deltas, const float4* anchors, long long n,
                                          const int32_t* shapes, float wy, float wx, float wh, float ww, float clampv,
                                          float min_side, float4* out, uint8_t* keep) {
  const int img = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t o = (size_t)img * n + i;
  float4 b = d2b_decode(__ldg(deltas + o), __ldg(anchors + i), wy, wx, wh, ww, clampv);
  b = d2b_clip(b, (float)shapes[2 * img], (float)shapes[2 * img + 1]);
  out[o] = b;
  if (keep) keep[o] = (min_side > 0.0f) ? (((b.w - b.y) >= min_side && (b.z - b.x) >= min_side) ? 1 : 0) : 1;
}

// ---- Matcher.__call__ on materialised matrices (the reference signature, matcher.py:57-150)
// row pass: highest_quality_foreach_gt; one CTA per (row, column chunk), 16-byte loads when aligned
__global__ void matcher_rowmax_kernel(const float* q, long long N, unsigned* rowmax) {
  const int i = blockIdx.y;
  const float* row = q + (size_t)i * N;
  unsigned k = 0u;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (long long)gridDim.x * blockDim.x)
    k = max(k, float_to_key(__ldg(row + j)));
  k = __reduce_max_sync(0xffffffffu, k);
  if ((threadIdx.x & 31) == 0 && k) atomicMax(rowmax + i, k);
}
// column pass: thread per prediction, rows streamed (coalesced across the warp)
__global__ void matcher_col_kernel(const float* q, int M, long long N, const float* crowd, int Mc, const float* diff,
                                   int Md, int has_crowd, int has_diff, const unsigned* rowmax, LabelArgs a,
                                   long long* matches, long long* labels) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  long long match = 0;
  int label = 0;
  if (M > 0) {
    float best = __ldg(q + j);
    int bi = 0;
    bool lowq = a.allow_lq && float_to_key(best) != 0u && float_to_key(best) == rowmax[0];
    for (int i = 1; i < M; ++i) {
      const float v = __ldg(q + (size_t)i * N + j);
      if (v > best) { best = v; bi = i; }
      if (a.allow_lq) {
        const unsigned k = float_to_key(v);
        lowq |= (k != 0u && k == rowmax[i]);
      }
    }
    match = bi;
    for (int t = 0; t <= a.nt; ++t) {
      const float low = (t == 0) ? __int_as_float(0xff800000) : a.thr[t - 1];
      const float high = (t == a.nt) ? __int_as_float(0x7f800000) : a.thr[t];
      if (best >= low && best < high) label = a.labels[t];
    }
    if (lowq) label = 1;
  }
  if (has_crowd) {
    bool cb = false;
    if (Mc > 0) {
      float mx = __ldg(crowd + j);
      for (int i = 1; i < Mc; ++i) mx = fmaxf(mx, __ldg(crowd + (size_t)i * N + j));
      cb = mx > 1e-3f;
    }
    if (label == 0 && cb) label = -1;
  }
  if (has_diff) {
    bool db = false;
    if (Md > 0) {
      float mx = __ldg(diff + j);
      for (int i = 1; i < Md; ++i) mx = fmaxf(mx, __ldg(diff + (size_t)i * N + j));
      db = mx > a.thr[0];
    }
    if (label == 0 && db) label = -1;
  }
  matches[j] = match;
  labels[j] = label;
}

size_t label_smem_bytes(int G) { return (size_t)3 * G * 16 + (size_t)3 * G * 4 + (size_t)G * 4 + (size_t)3 * G * 4 + 48; }

}  // namespace
}  // namespace d2b

using namespace d2b;

extern "C" size_t d2b_get_deltas_workspace_bytes(const d2b_get_deltas_params*) { return 0; }
extern "C" int d2b_get_deltas(const d2b_get_deltas_params* p, void*, size_t, d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->n >= 0 && p->n < (1ll << 31) * 256, "get_deltas: bad n");
  if (p->n == 0) return D2B_OK;
  D2B_REQUIRE(p->src_boxes && p->target_boxes && p->out, "get_deltas: NULL pointer");
  get_deltas_kernel<<<(unsigned)((p->n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(p->src_boxes), reinterpret_cast<const float4*>(p->target_boxes), p->n,
      p->weights[0], p->weights[1], p->weights[2], p->weights[3], reinterpret_cast<float4*>(p->out));
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}

extern "C" size_t d2b_pairwise_iou_workspace_bytes(const d2b_pairwise_iou_params*) { return 0; }
extern "C" int d2b_pairwise_iou(const d2b_pairwise_iou_params* p, void*, size_t, d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->n1 >= 0 && p->n2 >= 0 && p->n1 < 65536ll * 32768, "pairwise_iou: bad sizes");
  if (p->n1 == 0 || p->n2 == 0) return D2B_OK;
  D2B_REQUIRE(p->boxes1 && p->boxes2 && p->out, "pairwise_iou: NULL pointer");
  D2B_REQUIRE(p->n1 <= 65535, "pairwise_iou: n1=%lld > 65535 rows (put the larger set in boxes2)", (long long)p->n1);
  D2B_REQUIRE(p->iou_type >= D2B_IOU && p->iou_type <= D2B_CIOU, "pairwise_iou: unknown iou_type %d", p->iou_type);
  const unsigned gx = (unsigned)((p->n2 + 4 * 256 - 1) / (4 * 256));
  if (p->iou_type != D2B_IOU)
    pairwise_iou_variant_kernel<<<dim3(gx, (unsigned)p->n1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(p->boxes1), reinterpret_cast<const float4*>(p->boxes2), p->n2, p->iou_type, p->out);
  else
  pairwise_iou_kernel<<<dim3(gx, (unsigned)p->n1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(p->boxes1), p->n1, reinterpret_cast<const float4*>(p->boxes2), p->n2, p->out);
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}

extern "C" size_t d2b_label_boxes_workspace_bytes(const d2b_label_boxes_params* p) {
  if (!p || !p->allow_low_quality_matches || p->num_images <= 0 || p->max_gt <= 0) return 0;
  return ws_slice((size_t)p->num_images * p->max_gt * sizeof(unsigned));
}

extern "C" int d2b_label_boxes(const d2b_label_boxes_params* p, void* workspace, size_t workspace_bytes,
                               d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->num_images >= 0 && p->num_preds >= 0 && p->max_gt >= 0, "label_boxes: negative sizes");
  D2B_REQUIRE(p->max_gt <= D2B_MATCH_MAX_GT, "label_boxes: max_gt=%d > %d", p->max_gt, D2B_MATCH_MAX_GT);
  D2B_REQUIRE(p->num_thresholds >= 1 && p->num_thresholds <= D2B_MATCH_MAX_THRESHOLDS, "label_boxes: 1..%d thresholds",
              D2B_MATCH_MAX_THRESHOLDS);
  for (int t = 0; t + 1 < p->num_thresholds; ++t)  // matcher.py:49 assert low <= high
    D2B_REQUIRE(p->thresholds[t] <= p->thresholds[t + 1], "label_boxes: thresholds must be ascending");
  for (int t = 0; t <= p->num_thresholds; ++t)  // matcher.py:50
    D2B_REQUIRE(p->labels[t] >= -1 && p->labels[t] <= 1, "label_boxes: labels must be in {-1,0,1}");
  if (p->num_images == 0 || p->num_preds == 0) return D2B_OK;
  D2B_REQUIRE(p->num_images <= 65535, "label_boxes: too many images");
  D2B_REQUIRE(p->pred_boxes && p->out_matches && p->out_labels, "label_boxes: NULL pointer");
  D2B_REQUIRE(p->max_gt == 0 || (p->gt_boxes && p->gt_valid), "label_boxes: gt_boxes/gt_valid must be non-NULL");
  D2B_REQUIRE(p->boundary_threshold < 0.0f || p->image_shapes, "label_boxes: boundary_threshold needs image_shapes");
  D2B_REQUIRE(!p->compute_deltas || p->out_deltas, "label_boxes: compute_deltas needs out_deltas");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LabelArgs a;
  a.pred = reinterpret_cast<const float4*>(p->pred_boxes);
  a.pred_shared = p->pred_shared ? 1 : 0;
  a.pred_counts = p->pred_counts;
  a.N = p->num_images; a.P = p->num_preds; a.G = p->max_gt;
  a.gt = reinterpret_cast<const float4*>(p->gt_boxes);
  a.valid = p->gt_valid; a.crowd = p->gt_crowd; a.difficult = p->gt_difficult;
  for (int t = 0; t < D2B_MATCH_MAX_THRESHOLDS; ++t) a.thr[t] = t < p->num_thresholds ? p->thresholds[t] : 0.0f;
  a.nt = p->num_thresholds;
  for (int t = 0; t <= D2B_MATCH_MAX_THRESHOLDS; ++t) a.labels[t] = t <= p->num_thresholds ? p->labels[t] : 0;
  a.allow_lq = (p->allow_low_quality_matches && p->max_gt > 0) ? 1 : 0;
  a.boundary = p->boundary_threshold;
  a.shapes = p->image_shapes;
  a.deltas_on = p->compute_deltas;
  a.wy = p->weights[0]; a.wx = p->weights[1]; a.wh = p->weights[2]; a.ww = p->weights[3];
  a.out_matches = reinterpret_cast<long long*>(p->out_matches);
  a.out_labels = reinterpret_cast<long long*>(p->out_labels);
  a.out_deltas = p->compute_deltas ? reinterpret_cast<float4*>(p->out_deltas) : nullptr;
  a.gtmax = nullptr;
  const size_t smem = label_smem_bytes(p->max_gt);
  if (smem > 48 * 1024) {
    D2B_CUDA(cudaFuncSetAttribute(label_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    D2B_CUDA(cudaFuncSetAttribute(label_gtmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (a.allow_lq) {
    const size_t need = d2b_label_boxes_workspace_bytes(p);
    if (workspace == nullptr || workspace_bytes < need) {
      set_last_error("label_boxes needs %zu workspace bytes", need);
      return D2B_EWORKSPACE;
    }
    a.gtmax = static_cast<unsigned*>(workspace);
    D2B_CUDA(cudaMemsetAsync(a.gtmax, 0, (size_t)p->num_images * p->max_gt * sizeof(unsigned), st));
    const int per = kThreads * kPredsPerThread;
    label_gtmax_kernel<<<dim3((p->num_preds + per - 1) / per, p->num_images), kThreads, smem, st>>>(a);
    D2B_LAUNCH_CHECK();
  }
  label_kernel<<<dim3((p->num_preds + kThreads - 1) / kThreads, p->num_images), kThreads, smem, st>>>(a);
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}

extern "C" size_t d2b_matcher_workspace_bytes(const d2b_matcher_params* p) {
  if (!p || !p->allow_low_quality_matches || p->num_gt <= 0) return 0;
  return ws_slice((size_t)p->num_gt * sizeof(unsigned));
}
extern "C" int d2b_matcher(const d2b_matcher_params* p, void* workspace, size_t workspace_bytes, d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->num_gt >= 0 && p->num_preds >= 0 && p->num_crowd >= 0 && p->num_difficult >= 0, "matcher: negative sizes");
  D2B_REQUIRE(p->num_gt <= 65535, "matcher: num_gt=%d > 65535", p->num_gt);
  D2B_REQUIRE(p->num_thresholds >= 1 && p->num_thresholds <= D2B_MATCH_MAX_THRESHOLDS, "matcher: 1..%d thresholds",
              D2B_MATCH_MAX_THRESHOLDS);
  for (int t = 0; t + 1 < p->num_thresholds; ++t)
    D2B_REQUIRE(p->thresholds[t] <= p->thresholds[t + 1], "matcher: thresholds must be ascending");
  for (int t = 0; t <= p->num_thresholds; ++t)
    D2B_REQUIRE(p->labels[t] >= -1 && p->labels[t] <= 1, "matcher: labels must be in {-1,0,1}");
  if (p->num_preds == 0) return D2B_OK;
  D2B_REQUIRE(p->out_matches && p->out_labels, "matcher: NULL output");
  D2B_REQUIRE(p->num_gt == 0 || p->match_quality_matrix, "matcher: match_quality_matrix is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LabelArgs a = {};
  for (int t = 0; t < D2B_MATCH_MAX_THRESHOLDS; ++t) a.thr[t] = t < p->num_thresholds ? p->thresholds[t] : 0.0f;
  a.nt = p->num_thresholds;
  for (int t = 0; t <= D2B_MATCH_MAX_THRESHOLDS; ++t) a.labels[t] = t <= p->num_thresholds ? p->labels[t] : 0;
  a.allow_lq = (p->allow_low_quality_matches && p->num_gt > 0) ? 1 : 0;
  unsigned* rowmax = nullptr;
  if (a.allow_lq) {
    const size_t need = d2b_matcher_workspace_bytes(p);
    if (workspace == nullptr || workspace_bytes < need) {
      set_last_error("matcher needs %zu workspace bytes", need);
      return D2B_EWORKSPACE;
    }
    rowmax = static_cast<unsigned*>(workspace);
    D2B_CUDA(cudaMemsetAsync(rowmax, 0, (size_t)p->num_gt * sizeof(unsigned), st));
    const unsigned gx = (unsigned)((p->num_preds + 8 * 256 - 1) / (8 * 256));
    matcher_rowmax_kernel<<<dim3(gx, p->num_gt), 256, 0, st>>>(p->match_quality_matrix, p->num_preds, rowmax);
    D2B_LAUNCH_CHECK();
  }
  matcher_col_kernel<<<(unsigned)((p->num_preds + 255) / 256), 256, 0, st>>>(
      p->match_quality_matrix, p->num_gt, p->num_preds, p->crowd_matrix, p->num_crowd, p->difficult_matrix,
      p->num_difficult, p->use_crowd, p->use_difficult, rowmax, a, reinterpret_cast<long long*>(p->out_matches),
      reinterpret_cast<long long*>(p->out_labels));
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}

extern "C" size_t d2b_decode_clip_filter_workspace_bytes(const d2b_decode_clip_filter_params*) { return 0; }
extern "C" int d2b_decode_clip_filter(const d2b_decode_clip_filter_params* p, void*, size_t, d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->num_images >= 0 && p->n >= 0 && p->num_images <= 65535 && p->n < (1ll << 31) * 256,
              "decode_clip_filter: bad sizes");
  if (p->num_images == 0 || p->n == 0) return D2B_OK;
  D2B_REQUIRE(p->deltas && p->anchors && p->image_shapes && p->out_boxes, "decode_clip_filter: NULL pointer");
  decode_clip_filter_kernel<<<dim3((unsigned)((p->n + 255) / 256), p->num_images), 256, 0,
                              static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(p->deltas), reinterpret_cast<const float4*>(p->anchors), p->n, p->image_shapes,
      p->weights[0], p->weights[1], p->weights[2], p->weights[3], p->scale_clamp, p->min_box_side_len,
      reinterpret_cast<float4*>(p->out_boxes), p->out_keep);
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}
