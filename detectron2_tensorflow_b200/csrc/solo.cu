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
