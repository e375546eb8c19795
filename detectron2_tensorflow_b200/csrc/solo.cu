// solo.cu -- SOLOv2 neighbours of Matrix-NMS (SURVEY.md 8f "next" #4).
#include <math.h>

#include "kernels.cuh"

// ------------------------------------------------------------------ point_nms (solo_v2.py:29-40)
namespace d2b {
namespace {
typedef unsigned long long u64;

// solo_v2.py:513-517,530-533 fused: sigmoid -> threshold -> bit-pack (+ exact mask sum, + sum of the scores
// under the mask).  sigmoid is monotone up to a few ulp, so `sigmoid(x) > thr` is decided by comparing x with
// logit(thr) outside a guard band [lo, hi] and by the bit-exact sigmoid inside it (and for every set bit, whose
// score is needed anyway).  Same pack layout as mnms_pack4: lane = one float4, 16 lanes = one 64-bit word.
struct EncodeArgs {
  const float* logits;
  const int32_t* counts;
  int B, n;
  long long hw;
  int Wd;
  float thr, lo, hi;
  u64* packed;
  float* sum_masks;
  float* score_sums;
};
constexpr int kEncSteps = 8;
#define kPadNaN __int_as_float(0x7fc00000)  // pixels past hw: NaN never sets a bit

__device__ __forceinline__ unsigned enc_bit(float x, const EncodeArgs& a, float& acc) {
  if (!(x >= a.lo)) return 0u;                       // far below logit(thr) (or NaN): sigmoid(x) > thr is false
  const float s = d2b_sigmoidf(x);
  const bool on = (x > a.hi) ? true : (s > a.thr);   // inside the band the exact sigmoid decides
  if (on) acc = acc + s;
  return on ? 1u : 0u;
}

template <bool VEC>
__global__ void __launch_bounds__(256) solo_encode_kernel(EncodeArgs a) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int rows = a.counts ? min(a.counts[b], a.n) : a.n;
  if (i >= rows) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* m = a.logits + ((size_t)b * a.n + i) * a.hw;
  const int w_first = blockIdx.x * (8 * kEncSteps * 2) + warp * (kEncSteps * 2);
  float4 v[kEncSteps];
#pragma unroll
  for (int s = 0; s < kEncSteps; ++s) {
    const long long q = (long long)(w_first + 2 * s) * 16 + lane;  // float4 index
    if (VEC) {
      v[s] = q < (a.hw >> 2) ? __ldcs(reinterpret_cast<const float4*>(m) + q) : make_float4(kPadNaN, kPadNaN, kPadNaN, kPadNaN);
    } else {
      const long long e = q * 4;
      v[s].x = e + 0 < a.hw ? __ldg(m + e + 0) : kPadNaN;
      v[s].y = e + 1 < a.hw ? __ldg(m + e + 1) : kPadNaN;
      v[s].z = e + 2 < a.hw ? __ldg(m + e + 2) : kPadNaN;
      v[s].w = e + 3 < a.hw ? __ldg(m + e + 3) : kPadNaN;
    }
  }
  unsigned total = 0;
  float acc = 0.0f;
#pragma unroll
  for (int s = 0; s < kEncSteps; ++s) {
    const unsigned nib = enc_bit(v[s].x, a, acc) | (enc_bit(v[s].y, a, acc) << 1) | (enc_bit(v[s].z, a, acc) << 2) |
                         (enc_bit(v[s].w, a, acc) << 3);
    u64 w = (u64)nib << (4 * (lane & 15));
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) w |= __shfl_xor_sync(0xffffffffu, w, o);
    const int word = w_first + 2 * s + (lane >> 4);
    if ((lane & 15) == 0 && word < a.Wd) {
      a.packed[((size_t)b * a.n + i) * a.Wd + word] = w;
      total += __popcll(w);
    }
  }
  total += __shfl_xor_sync(0xffffffffu, total, 16);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (total) atomicAdd(a.sum_masks + (size_t)b * a.n + i, (float)total);  // integer-valued partial sums < 2^24: exact
    if (acc != 0.0f) atomicAdd(a.score_sums + (size_t)b * a.n + i, acc);
  }
}
}  // namespace
}  // namespace d2b

namespace d2b {
namespace {
__global__ void point_nms_kernel(const float4* x, int H, int W, int C4, float4* out, long long total) {
  // one thread per 4 channels of one cell (NHWC, C % 4 == 0): neighbours are -C, -W*C, -(W+1)*C away
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long cell = t / C4;
  const int xx = (int)(cell % W);
  const int y = (int)((cell / W) % H);
  const float4 v = __ldg(x + t);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 lf = xx > 0 ? __ldg(x + t - C4) : z;
  const float4 up = y > 0 ? __ldg(x + t - (long long)W * C4) : z;
  const float4 ul = (xx > 0 && y > 0) ? __ldg(x + t - (long long)(W + 1) * C4) : z;
  auto one = [](float a, float b, float c, float d) {
    const float m = fmaxf(fmaxf(a, b), fmaxf(c, d));
    return a * (a == m ? 1.0f : 0.0f);  // inputs * cast(equal(inputs, pooled)) (:38-40)
  };
  out[t] = make_float4(one(v.x, up.x, lf.x, ul.x), one(v.y, up.y, lf.y, ul.y), one(v.z, up.z, lf.z, ul.z),
                       one(v.w, up.w, lf.w, ul.w));
}
__global__ void point_nms_scalar_kernel(const float* x, int H, int W, int C, float* out, long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long cell = t / C;
  const int xx = (int)(cell % W);
  const int y = (int)((cell / W) % H);
  const float v = __ldg(x + t);
  const float lf = xx > 0 ? __ldg(x + t - C) : 0.0f;
  const float up = y > 0 ? __ldg(x + t - (long long)W * C) : 0.0f;
  const float ul = (xx > 0 && y > 0) ? __ldg(x + t - (long long)(W + 1) * C) : 0.0f;
  const float m = fmaxf(fmaxf(v, up), fmaxf(lf, ul));
  out[t] = v * (v == m ? 1.0f : 0.0f);
}
}  // namespace
}  // namespace d2b

namespace d2b {
namespace {
// ------------------------------------------------------------------ SOLOv2 inference tail (solo_v2.py:507-558)
__device__ __forceinline__ u64 solo_key(float score, unsigned idx) {
  return ((u64)float_to_key(score) << 32) | (u64)(0xffffffffu - idx);
}
// :520-533  keep = sum_masks > strides; pred_scores *= score_sums / sum_masks; candidates become sort keys
__global__ void solo_score_kernel(const float* scores, const float* strides, const float* sum_masks,
                                  const float* score_sums, const int32_t* counts, int n, int P, u64* keys,
                                  float* rescored, int32_t* nvalid) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int cnt = counts ? min(counts[b], n) : n;
  bool ok = false;
  if (i < P) {
    u64 key = 0ull;
    if (i < cnt) {
      const size_t o = (size_t)b * n + i;
      const float sm = sum_masks[o];
      if (sm > strides[o]) {
        float ms = score_sums[o] / sm;
        const float s = scores[o] * ms;
        rescored[o] = s;
        key = solo_key(s, (unsigned)i);
        ok = true;
      }
    }
    keys[(size_t)b * P + i] = key;
  }
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(nvalid + b, __popc(m));
}
// :536-540  the top-k rows (already sorted) gathered into a dense [kcap] block per image
__global__ void solo_gather_kernel(const u64* keys, int P, const int32_t* nvalid, int pre, int n, int kcap, int Wd,
                                   const u64* packed, const float* sum_masks, const long long* classes,
                                   const float* rescored, u64* packed2, float* sum2, long long* cls2, float* sc2,
                                   int32_t* kcount) {
  const int b = blockIdx.y, r = blockIdx.x;
  const int kc = min(min(pre, nvalid[b]), kcap);
  if (r == 0 && threadIdx.x == 0) kcount[b] = kc;
  if (r >= kc) return;
  const unsigned idx = 0xffffffffu - (unsigned)keys[(size_t)b * P + r];
  const u64* src = packed + ((size_t)b * n + idx) * Wd;
  u64* dst = packed2 + ((size_t)b * kcap + r) * Wd;
  for (int w = threadIdx.x; w < Wd; w += blockDim.x) dst[w] = src[w];
  if (threadIdx.x == 0) {
    const size_t o = (size_t)b * n + idx, q = (size_t)b * kcap + r;
    sum2[q] = sum_masks[o];
    cls2[q] = classes[o];
    sc2[q] = rescored[o];
  }
}
// :549-556  keep = updated > update_score_threshold (order kept), pad / clip to max_det.  One CTA per image.
__global__ void __launch_bounds__(256) solo_emit_kernel(const float* upd, const long long* cls2, const int32_t* kcount,
                                                        int kcap, float thr, int D, long long* out_classes,
                                                        float* out_scores, uint8_t* out_valid, int32_t* src_row,
                                                        int32_t* out_num) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const int kc = kcount[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int r0 = 0; r0 < kc; r0 += 256) {
    const int r = r0 + threadIdx.x;
    const bool keep = r < kc && upd[(size_t)b * kcap + r] > thr;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { before += (w < warp) ? s_warp[w] : 0; total += s_warp[w]; }
    const int slot = s_base + before + __popc(m & ((1u << lane) - 1u));
    if (keep && slot < D) {
      const size_t o = (size_t)b * D + slot;
      out_classes[o] = cls2[(size_t)b * kcap + r];
      out_scores[o] = upd[(size_t)b * kcap + r];
      out_valid[o] = 1;
      src_row[o] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
  const int nv = min(s_base, D);
  for (int q = nv + threadIdx.x; q < D; q += 256) {
    const size_t o = (size_t)b * D + q;
    out_classes[o] = 0; out_scores[o] = 0.0f; out_valid[o] = 0; src_row[o] = -1;
  }
  if (threadIdx.x == 0 && out_num) out_num[b] = nv;
}
// masks of the kept detections: packed copy and / or fp32 0/1 expansion (zeros past the valid count)
__global__ void __launch_bounds__(256) solo_unpack_kernel(const u64* packed2, const int32_t* src_row, int kcap, int Wd,
                                                          long long hw, int D, u64* out_packed, float* out_masks) {
  const int b = blockIdx.z, q = blockIdx.y;
  const int r = src_row[(size_t)b * D + q];
  const u64* src = r >= 0 ? packed2 + ((size_t)b * kcap + r) * Wd : nullptr;
  const int w = blockIdx.x * 256 + threadIdx.x;  // one word (64 pixels) per thread
  if (w >= Wd) return;
  const u64 bits = src ? src[w] : 0ull;
  if (out_packed) out_packed[((size_t)b * D + q) * Wd + w] = bits;
  if (out_masks) {
    float* o = out_masks + ((size_t)b * D + q) * hw + (long long)w * 64;
    const long long left = hw - (long long)w * 64;
    if (left >= 64 && (hw & 3) == 0) {
#pragma unroll
      for (int v = 0; v < 16; ++v)
        __stcs(reinterpret_cast<float4*>(o) + v,
               make_float4((bits >> (4 * v)) & 1 ? 1.f : 0.f, (bits >> (4 * v + 1)) & 1 ? 1.f : 0.f,
                           (bits >> (4 * v + 2)) & 1 ? 1.f : 0.f, (bits >> (4 * v + 3)) & 1 ? 1.f : 0.f));
    } else {
      for (int c = 0; c < 64 && c < left; ++c) o[c] = (bits >> c) & 1 ? 1.0f : 0.0f;
    }
  }
}

struct SoloPlan {
  int P, kcap, Wd;
  size_t bytes, o_packed, o_sum, o_ssum, o_resc, o_keys, o_nvalid, o_kcount, o_packed2, o_sum2, o_cls2, o_sc2, o_upd,
      o_src, o_mnms, o_dyn, dyn_bytes;
  bool dynamic;
  d2b_matrix_nms_params mp;
};
int solo_plan(const d2b_solo_postprocess_params* p, SoloPlan& pl) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->batch >= 0 && p->n >= 0 && p->hw >= 0, "solo_postprocess: negative sizes");
  D2B_REQUIRE(p->batch <= 65535 && p->n <= 65535, "solo_postprocess: batch / n too large");
  D2B_REQUIRE(p->hw < (1ll << 24), "solo_postprocess: hw >= 2^24");
  D2B_REQUIRE(p->pre_nms_topk >= 1 && p->max_detections >= 1 && p->max_detections <= 65535,
              "solo_postprocess: bad pre_nms_topk / max_detections");
  D2B_REQUIRE(p->kernel == D2B_MNMS_GAUSSIAN || p->kernel == D2B_MNMS_LINEAR, "NMS kernel %d not implemented yet.", p->kernel);
  const size_t B = p->batch, n = p->n > 0 ? p->n : 1;
  pl.Wd = (int)((p->hw + 63) / 64);
  if (pl.Wd < 1) pl.Wd = 1;
  pl.P = 1;
  while ((size_t)pl.P < n) pl.P <<= 1;
  pl.kcap = (int)(p->pre_nms_topk < (int)n ? p->pre_nms_topk : n);
  const size_t Wd = pl.Wd, kc = pl.kcap, D = p->max_detections;
  size_t o = 0;
  pl.o_packed = o; o += ws_slice(B * n * Wd * 8);
  pl.o_sum = o; o += ws_slice(B * n * 4);
  pl.o_ssum = o; o += ws_slice(B * n * 4);
  pl.o_resc = o; o += ws_slice(B * n * 4);
  pl.o_keys = o; o += ws_slice(B * pl.P * 8);
  pl.o_nvalid = o; o += ws_slice(B * 4);
  pl.o_kcount = o; o += ws_slice(B * 4);
  pl.o_packed2 = o; o += ws_slice(B * kc * Wd * 8);
  pl.o_sum2 = o; o += ws_slice(B * kc * 4);
  pl.o_cls2 = o; o += ws_slice(B * kc * 8);
  pl.o_sc2 = o; o += ws_slice(B * kc * 4);
  pl.o_upd = o; o += ws_slice(B * kc * 4);
  pl.o_src = o; o += ws_slice(B * D * 4);
  // masks from the dynamic conv (solo_v2.py:499-511) when no logits are given
  pl.dynamic = p->mask_logits == nullptr && p->mask_features != nullptr;
  pl.o_dyn = o;
  pl.dyn_bytes = 0;
  if (pl.dynamic) {
    D2B_REQUIRE(p->channels >= 4 && p->channels % 4 == 0 && p->channels <= 4096,
                "solo_postprocess: channels=%d must be a multiple of 4 in [4, 4096]", p->channels);
    pl.dyn_bytes = solo_dynamic_masks_ws(p->batch, p->n, p->channels);
    o += pl.dyn_bytes;
  }
  pl.o_mnms = o;
  pl.mp = d2b_matrix_nms_params();
  pl.mp.batch = p->batch; pl.mp.n = pl.kcap; pl.mp.hw = p->hw; pl.mp.kernel = p->kernel; pl.mp.sigma = p->sigma;
  pl.mp.packed_masks = reinterpret_cast<const uint64_t*>(8);  // non-NULL: size the packed-input workspace
  o += d2b_matrix_nms_workspace_bytes(&pl.mp);
  pl.bytes = o;
  return D2B_OK;
}
}  // namespace
}  // namespace d2b

using namespace d2b;

extern "C" size_t d2b_point_nms_workspace_bytes(const d2b_point_nms_params*) { return 0; }
extern "C" int d2b_point_nms(const d2b_point_nms_params* p, void*, size_t, d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->num_images >= 0 && p->height >= 0 && p->width >= 0 && p->channels >= 0, "point_nms: negative sizes");
  const long long total = (long long)p->num_images * p->height * p->width * p->channels;
  if (total == 0) return D2B_OK;
  D2B_REQUIRE(p->scores && p->out, "point_nms: NULL pointer");
  D2B_REQUIRE(p->scores != p->out, "point_nms: in-place operation is not supported (neighbours are re-read)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->channels % 4 == 0 && (reinterpret_cast<uintptr_t>(p->scores) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(p->out) & 15) == 0) {
    const long long t4 = total / 4;
    point_nms_kernel<<<(unsigned)((t4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(p->scores),
                                                                   p->height, p->width, p->channels / 4,
                                                                   reinterpret_cast<float4*>(p->out), t4);
  } else {
    point_nms_scalar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p->scores, p->height, p->width,
                                                                             p->channels, p->out, total);
  }
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}

extern "C" size_t d2b_solo_mask_encode_workspace_bytes(const d2b_solo_mask_encode_params*) { return 0; }
extern "C" int d2b_solo_mask_encode(const d2b_solo_mask_encode_params* p, void*, size_t, d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->batch >= 0 && p->n >= 0 && p->hw >= 0, "solo_mask_encode: negative sizes");
  if (p->batch == 0 || p->n == 0 || p->hw == 0) return D2B_OK;
  D2B_REQUIRE(p->n <= 65535 && p->batch <= 65535, "solo_mask_encode: n / batch too large");
  D2B_REQUIRE(p->hw < (1ll << 24), "solo_mask_encode: hw=%lld >= 2^24 (mask sums must stay exact in fp32)", (long long)p->hw);
  D2B_REQUIRE(p->mask_logits && p->packed_masks && p->sum_masks && p->score_sums, "solo_mask_encode: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EncodeArgs a;
  a.logits = p->mask_logits; a.counts = p->counts; a.B = p->batch; a.n = p->n; a.hw = p->hw;
  a.Wd = (int)((p->hw + 63) / 64);
  a.thr = p->mask_threshold;
  // guard band around logit(thr); thresholds near 0/1 (or outside) always take the exact path
  const double t = (double)p->mask_threshold;
  if (t > 1e-3 && t < 1.0 - 1e-3) {
    const double x0 = log(t / (1.0 - t));
    const double d = 1e-3 * (fabs(x0) > 1.0 ? fabs(x0) : 1.0);
    a.lo = (float)(x0 - d);
    a.hi = (float)(x0 + d);
  } else {
    a.lo = -INFINITY;
    a.hi = INFINITY;
  }
  a.packed = reinterpret_cast<u64*>(p->packed_masks);
  a.sum_masks = p->sum_masks;
  a.score_sums = p->score_sums;
  D2B_CUDA(cudaMemsetAsync(p->sum_masks, 0, sizeof(float) * (size_t)p->batch * p->n, st));
  D2B_CUDA(cudaMemsetAsync(p->score_sums, 0, sizeof(float) * (size_t)p->batch * p->n, st));
  if (p->counts)  // rows past the valid prefix: defined (empty) masks
    D2B_CUDA(cudaMemsetAsync(p->packed_masks, 0, sizeof(u64) * (size_t)p->batch * p->n * a.Wd, st));
  const dim3 grid((a.Wd + 8 * kEncSteps * 2 - 1) / (8 * kEncSteps * 2), a.n, a.B);
  if (a.hw % 4 == 0 && (reinterpret_cast<uintptr_t>(a.logits) & 15) == 0)
    solo_encode_kernel<true><<<grid, 256, 0, st>>>(a);
  else
    solo_encode_kernel<false><<<grid, 256, 0, st>>>(a);
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}

extern "C" size_t d2b_solo_postprocess_workspace_bytes(const d2b_solo_postprocess_params* p) {
  SoloPlan pl;
  if (solo_plan(p, pl) != D2B_OK) return 0;
  return pl.bytes;
}

extern "C" int d2b_solo_postprocess(const d2b_solo_postprocess_params* p, void* workspace, size_t workspace_bytes,
                                    d2b_stream_t stream) {
  SoloPlan pl;
  int rc = solo_plan(p, pl);
  if (rc != D2B_OK) return rc;
  if (p->batch == 0) return D2B_OK;
  D2B_REQUIRE(p->out_classes && p->out_scores && p->out_valid, "solo_postprocess: NULL output");
  D2B_REQUIRE(p->n == 0 || ((p->mask_logits || (p->mask_features && p->mask_kernels)) && p->scores && p->classes && p->strides),
              "solo_postprocess: NULL input");
  if (workspace == nullptr || workspace_bytes < pl.bytes) {
    set_last_error("solo_postprocess needs %zu workspace bytes", pl.bytes);
    return D2B_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  const int B = p->batch, n = p->n, D = p->max_detections, kcap = pl.kcap, Wd = pl.Wd, P = pl.P;
  u64* packed = reinterpret_cast<u64*>(ws + pl.o_packed);
  float* sum_masks = reinterpret_cast<float*>(ws + pl.o_sum);
  float* score_sums = reinterpret_cast<float*>(ws + pl.o_ssum);
  float* rescored = reinterpret_cast<float*>(ws + pl.o_resc);
  u64* keys = reinterpret_cast<u64*>(ws + pl.o_keys);
  int32_t* nvalid = reinterpret_cast<int32_t*>(ws + pl.o_nvalid);
  int32_t* kcount = reinterpret_cast<int32_t*>(ws + pl.o_kcount);
  u64* packed2 = reinterpret_cast<u64*>(ws + pl.o_packed2);
  float* sum2 = reinterpret_cast<float*>(ws + pl.o_sum2);
  long long* cls2 = reinterpret_cast<long long*>(ws + pl.o_cls2);
  float* sc2 = reinterpret_cast<float*>(ws + pl.o_sc2);
  float* upd = reinterpret_cast<float*>(ws + pl.o_upd);
  int32_t* src_row = reinterpret_cast<int32_t*>(ws + pl.o_src);
  D2B_CUDA(cudaMemsetAsync(nvalid, 0, sizeof(int32_t) * B, st));
  D2B_CUDA(cudaMemsetAsync(kcount, 0, sizeof(int32_t) * B, st));
  if (n > 0 && p->hw > 0) {
    if (pl.dynamic) {
      d2b_solo_dynamic_masks_params e = {};
      e.mask_features = p->mask_features; e.mask_kernels = p->mask_kernels; e.counts = p->counts;
      e.batch = B; e.n = n; e.channels = p->channels; e.hw = p->hw;
      e.mask_threshold = p->mask_threshold;
      e.packed_masks = reinterpret_cast<uint64_t*>(packed); e.sum_masks = sum_masks; e.score_sums = score_sums;
      rc = d2b_solo_dynamic_masks(&e, ws + pl.o_dyn, pl.dyn_bytes, stream);
    } else {
      d2b_solo_mask_encode_params e = {};
      e.mask_logits = p->mask_logits; e.counts = p->counts; e.batch = B; e.n = n; e.hw = p->hw;
      e.mask_threshold = p->mask_threshold;
      e.packed_masks = reinterpret_cast<uint64_t*>(packed); e.sum_masks = sum_masks; e.score_sums = score_sums;
      rc = d2b_solo_mask_encode(&e, nullptr, 0, stream);
    }
    if (rc != D2B_OK) return rc;
    solo_score_kernel<<<dim3((P + 255) / 256, B), 256, 0, st>>>(p->scores, p->strides, sum_masks, score_sums, p->counts,
                                                               n, P, keys, rescored, nvalid);
    D2B_LAUNCH_CHECK();
    rc = sort_segments_desc(keys, B, P, nullptr, st);
    if (rc != D2B_OK) return rc;
    solo_gather_kernel<<<dim3(kcap, B), 128, 0, st>>>(keys, P, nvalid, p->pre_nms_topk, n, kcap, Wd, packed, sum_masks,
                                                     reinterpret_cast<const long long*>(p->classes), rescored, packed2,
                                                     sum2, cls2, sc2, kcount);
    D2B_LAUNCH_CHECK();
    d2b_matrix_nms_params m = pl.mp;
    m.packed_masks = reinterpret_cast<const uint64_t*>(packed2);
    m.masks = nullptr;
    m.classes = reinterpret_cast<const int64_t*>(cls2);
    m.scores = sc2;
    m.sum_masks = sum2;
    m.counts = kcount;
    m.out = upd;
    rc = d2b_matrix_nms(&m, ws + pl.o_mnms, pl.bytes - pl.o_mnms, stream);
    if (rc != D2B_OK) return rc;
  }
  solo_emit_kernel<<<B, 256, 0, st>>>(upd, cls2, kcount, kcap, p->update_score_threshold, D,
                                      reinterpret_cast<long long*>(p->out_classes), p->out_scores, p->out_valid, src_row,
                                      p->out_num);
  D2B_LAUNCH_CHECK();
  if ((p->out_masks || p->out_packed_masks) && p->hw > 0) {
    solo_unpack_kernel<<<dim3((Wd + 255) / 256, D, B), 256, 0, st>>>(packed2, src_row, kcap, Wd, p->hw, D,
                                                                     reinterpret_cast<u64*>(p->out_packed_masks),
                                                                     p->out_masks);
    D2B_LAUNCH_CHECK();
  }
  return D2B_OK;
}

// ------------------------------------------------------------------ candidate selection (solo_v2.py:481-497)
//   keep_inds = tf.where(pred_scores > score_threshold)      (row-major over [G cells, K classes])
//   scores = gather_nd, classes = keep_inds[:, 1], kernels = gather(pred_kernels, keep_inds[:, 0]), strides likewise
// Three small launches: kept scores per 4096-score chunk, an exclusive scan of the chunk counts per image, and a
// scatter in which every kept element's rank = chunk base + rank inside the chunk, i.e. the compaction is ORDERED
// (tf.where order) without a sort.  Candidates past `cap` are dropped (the
// reference has no cap; `out_total` reports how many passed so the caller can detect it).
namespace d2b {
namespace {
constexpr int kSelThreads = 256, kSelPer = 16;          // one CTA covers 4096 consecutive scores of one image
constexpr int kSelChunk = kSelThreads * kSelPer;
// pass 1: kept scores per chunk
__global__ void __launch_bounds__(kSelThreads) solo_select_count_kernel(const float* scores, long long total, float thr,
                                                                       int n_chunks, int32_t* chunk_counts) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* sc = scores + (size_t)b * total;
  const long long i0 = (long long)chunk * kSelChunk + (long long)threadIdx.x * kSelPer;
  int c = 0;
#pragma unroll
  for (int e = 0; e < kSelPer; ++e) {
    const long long i = i0 + e;
    c += (i < total && __ldg(sc + i) > thr) ? 1 : 0;
  }
  __shared__ int s_w[kSelThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kSelThreads / 32; ++w) t += s_w[w];
    chunk_counts[(size_t)b * n_chunks + chunk] = t;
  }
}
// pass 2: exclusive scan of the chunk counts of one image (one warp), counts / totals out
__global__ void solo_select_scan_kernel(int32_t* chunk_counts, int n_chunks, int cap, int32_t* out_counts, int32_t* out_total) {
  const int b = blockIdx.x, lane = threadIdx.x;
  int32_t* cc = chunk_counts + (size_t)b * n_chunks;
  int run = 0;
  for (int c0 = 0; c0 < n_chunks; c0 += 32) {
    const int c = c0 + lane;
    const int v = c < n_chunks ? cc[c] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (c < n_chunks) cc[c] = run + inc - v;  // exclusive prefix
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) {
    out_counts[b] = min(run, cap);
    if (out_total) out_total[b] = run;
  }
}
// pass 3: every chunk writes its kept elements at base + rank (thread = 16 consecutive scores, so ranks follow the
// row-major tf.where order); the last chunk's CTA also pads the rows past the count
__global__ void __launch_bounds__(kSelThreads) solo_select_scatter_kernel(const float* scores, const float* cell_strides,
                                                                         long long total, int K, float thr, int cap,
                                                                         int n_chunks, const int32_t* chunk_base,
                                                                         const int32_t* counts, float* out_scores,
                                                                         long long* out_classes, float* out_strides,
                                                                         int32_t* out_cells) {
  __shared__ int s_w[kSelThreads / 32];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* sc = scores + (size_t)b * total;
  const long long i0 = (long long)chunk * kSelChunk + (long long)threadIdx.x * kSelPer;
  float v[kSelPer];
  unsigned keep = 0;
#pragma unroll
  for (int e = 0; e < kSelPer; ++e) {
    const long long i = i0 + e;
    v[e] = i < total ? __ldg(sc + i) : 0.0f;
    keep |= (i < total && v[e] > thr) ? (1u << e) : 0u;
  }
  const int mine = __popc(keep);
  int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  int before = 0;
#pragma unroll
  for (int w = 0; w < kSelThreads / 32; ++w) before += w < warp ? s_w[w] : 0;
  int slot = chunk_base[(size_t)b * n_chunks + chunk] + before + inc - mine;
#pragma unroll
  for (int e = 0; e < kSelPer; ++e) {
    if ((keep >> e) & 1u) {
      if (slot < cap) {
        const long long i = i0 + e;
        const int cell = (int)(i / K);
        const size_t o = (size_t)b * cap + slot;
        out_scores[o] = v[e];
        out_classes[o] = (long long)(i - (long long)cell * K);
        out_strides[o] = cell_strides[cell];
        out_cells[o] = cell;
      }
      ++slot;
    }
  }
  if (chunk == n_chunks - 1) {  // defined padding
    for (int q = counts[b] + threadIdx.x; q < cap; q += kSelThreads) {
      const size_t o = (size_t)b * cap + q;
      out_scores[o] = 0.0f; out_classes[o] = 0; out_strides[o] = 1.0f; out_cells[o] = 0;
    }
  }
}
// kernels [B, G, E] rows of the kept cells -> [B, cap, E] (zeros past the count); one warp per row, 16-byte copies
__global__ void solo_gather_kernels_kernel(const float4* kernels, const int32_t* cells, const int32_t* counts, int G, int E4,
                                           int cap, float4* out) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= cap) return;
  const int lane = threadIdx.x & 31;
  const bool live = row < counts[b];
  const float4* src = kernels + ((size_t)b * G + (live ? cells[(size_t)b * cap + row] : 0)) * E4;
  float4* dst = out + ((size_t)b * cap + row) * E4;
  for (int c = lane; c < E4; c += 32) dst[c] = live ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
}
}  // namespace
}  // namespace d2b

static size_t solo_select_chunks(const d2b_solo_select_params* p) {
  const long long total = (long long)p->num_cells * p->num_classes;
  const long long n = (total + kSelChunk - 1) / kSelChunk;
  return (size_t)(n > 0 ? n : 1);
}
extern "C" size_t d2b_solo_select_workspace_bytes(const d2b_solo_select_params* p) {
  if (!p || p->batch < 0 || p->max_candidates < 1 || p->num_cells < 0 || p->num_classes < 1) return 0;
  return ws_slice(sizeof(int32_t) * (size_t)p->batch * p->max_candidates) +
         ws_slice(sizeof(int32_t) * (size_t)p->batch * solo_select_chunks(p));
}
extern "C" int d2b_solo_select(const d2b_solo_select_params* p, void* workspace, size_t workspace_bytes,
                               d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->batch >= 0 && p->num_cells >= 0 && p->num_classes >= 1 && p->channels >= 0, "solo_select: negative sizes");
  D2B_REQUIRE(p->max_candidates >= 1 && p->max_candidates <= 65535, "solo_select: max_candidates must be in [1, 65535]");
  D2B_REQUIRE(p->channels % 4 == 0, "solo_select: channels=%d must be a multiple of 4", p->channels);
  D2B_REQUIRE(p->batch <= 65535, "solo_select: batch too large");
  if (p->batch == 0) return D2B_OK;
  D2B_REQUIRE(p->scores && p->cell_strides && p->out_scores && p->out_classes && p->out_strides && p->out_counts,
              "solo_select: NULL pointer");
  D2B_REQUIRE(!p->out_kernels || p->kernels, "solo_select: out_kernels needs kernels");
  const size_t need = d2b_solo_select_workspace_bytes(p);
  if (workspace == nullptr || workspace_bytes < need) {
    set_last_error("solo_select needs %zu workspace bytes", need);
    return D2B_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* cells = static_cast<int32_t*>(workspace);
  int32_t* chunk_counts = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) +
                                                     ws_slice(sizeof(int32_t) * (size_t)p->batch * p->max_candidates));
  const long long total = (long long)p->num_cells * p->num_classes;
  const int n_chunks = (int)solo_select_chunks(p);
  const dim3 grid(n_chunks, p->batch);
  solo_select_count_kernel<<<grid, kSelThreads, 0, st>>>(p->scores, total, p->score_threshold, n_chunks, chunk_counts);
  D2B_LAUNCH_CHECK();
  solo_select_scan_kernel<<<p->batch, 32, 0, st>>>(chunk_counts, n_chunks, p->max_candidates, p->out_counts, p->out_total);
  D2B_LAUNCH_CHECK();
  solo_select_scatter_kernel<<<grid, kSelThreads, 0, st>>>(p->scores, p->cell_strides, total, p->num_classes,
                                                          p->score_threshold, p->max_candidates, n_chunks, chunk_counts,
                                                          p->out_counts, p->out_scores,
                                                          reinterpret_cast<long long*>(p->out_classes), p->out_strides, cells);
  D2B_LAUNCH_CHECK();
  if (p->out_kernels && p->channels > 0) {
    D2B_REQUIRE((reinterpret_cast<uintptr_t>(p->kernels) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->out_kernels) & 15) == 0,
                "solo_select: kernels must be 16-byte aligned");
    const dim3 grid((p->max_candidates + 7) / 8, p->batch);
    solo_gather_kernels_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(p->kernels), cells, p->out_counts,
                                                     p->num_cells, p->channels / 4, p->max_candidates,
                                                     reinterpret_cast<float4*>(p->out_kernels));
    D2B_LAUNCH_CHECK();
  }
  return D2B_OK;
}
