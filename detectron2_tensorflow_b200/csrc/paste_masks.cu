// paste_masks.cu -- mask paste-back (SURVEY.md 8f "next" #1):
//   reframe_box_masks_to_image_masks   lib/structures/mask_ops.py:7-56
//   (called by detector_postprocess    lib/modeling/postprocessing.py:9-59)
// The reference crops every box mask to the FULL image with tf.image.crop_and_resize on inverted boxes,
// materialising [M, H, W, 1] fp32 (430 MB per 100 detections at 800x1344), then thresholds to uint8.
// Here the bilinear sample and the threshold are fused and only the uint8 plane is written: the op is
// bound by that write (1 B/pixel).  Pixels outside the box are zeros by the extrapolation rule, so a CTA
// whose rows miss the box stores zeros without touching the mask.
#include <stdlib.h>

#include "kernels.cuh"

namespace d2b {
namespace {

constexpr int kPasteThreads = 256;
constexpr int kPx = 16;      // output pixels (bytes) per thread per iteration: one 16-byte store
constexpr int kIters = 8;    // iterations per CTA
constexpr int kMaxMaskSmem = 16384;  // floats

struct PasteArgs {
  const float* masks;  // [M, mh, mw]
  const float* boxes;  // [M, 4] absolute yxyx in output-image pixels
  long long M;
  int mh, mw, H, W;
  float thr;
  uint8_t* out;  // [M, H, W]
  int vec;       // 16-byte stores allowed
  int smem_mask;
};

struct Axis {  // tf.image.crop_and_resize coordinate rule along one axis
  float n1, step, mid;
  int crop, dim;
  __device__ __forceinline__ float at(int i) const { return crop > 1 ? n1 * (float)(dim - 1) + (float)i * step : mid; }
  // the same value from the index held as a float (exact below 2^24): the int -> float conversion per pixel ran on the
  // quarter-rate XU pipe, which was the kernel's limiter
  __device__ __forceinline__ float at_f(float fi) const { return crop > 1 ? n1 * (float)(dim - 1) + fi * step : mid; }
};
__device__ __forceinline__ Axis make_axis(float lo, float hi, int crop, int dim) {
  Axis a;
  a.crop = crop; a.dim = dim;
  // reverse box: ([0,1] - min) / (max - min)   mask_ops.py:44-49
  const float n1 = (0.0f - lo) / (hi - lo);
  const float n2 = (1.0f - lo) / (hi - lo);
  a.n1 = n1;
  a.step = crop > 1 ? (n2 - n1) * (float)(dim - 1) / (float)(crop - 1) : 0.0f;
  a.mid = 0.5f * (n1 + n2) * (float)(dim - 1);
  return a;
}

__global__ void __launch_bounds__(kPasteThreads) paste_masks_kernel(PasteArgs a) {
  extern __shared__ float s_mask[];
  const long long m = blockIdx.y;
  const unsigned plane = (unsigned)a.H * (unsigned)a.W;  // < 2^31 (checked on the host): 32-bit index math
  const unsigned cta_first = blockIdx.x * (unsigned)(kPasteThreads * kPx * kIters);
  if (cta_first >= plane) return;
  const unsigned cta_last = min(plane, cta_first + (unsigned)(kPasteThreads * kPx * kIters)) - 1u;
  const float4 bx = __ldg(reinterpret_cast<const float4*>(a.boxes) + m);
  // to_normalized_coordinates: scale by 1/height, 1/width (box_list_ops.py:829-839)
  const float ys = 1.0f / (float)a.H, xs = 1.0f / (float)a.W;
  const Axis ay = make_axis(ys * bx.x, ys * bx.z, a.H, a.mh);
  const Axis ax = make_axis(xs * bx.y, xs * bx.w, a.W, a.mw);
  const float ymax_in = (float)(a.mh - 1), xmax_in = (float)(a.mw - 1);
  // does any output row of this CTA fall inside the mask?
  const int yA = (int)(cta_first / (unsigned)a.W), yB = (int)(cta_last / (unsigned)a.W);
  // at(i) is monotone in i under fp32 rounding (monotone product + monotone sum), so the rows of this
  // CTA map into [min(at(yA), at(yB)), max(...)]; NaN coordinates compare false -> zeros, as the exact rule.
  const float e0 = ay.at(yA), e1 = ay.at(yB);
  const bool any = fminf(e0, e1) <= ymax_in && fmaxf(e0, e1) >= 0.0f && e0 == e0 && e1 == e1;
  const float* mk = a.masks + (size_t)m * a.mh * a.mw;
  if (any && a.smem_mask) {  // block-uniform
    for (int i = threadIdx.x; i < a.mh * a.mw; i += kPasteThreads) s_mask[i] = __ldg(mk + i);
    __syncthreads();
    mk = s_mask;
  }
  uint8_t* o = a.out + (size_t)m * (size_t)plane;
  for (int it = 0; it < kIters; ++it) {
    const unsigned p0 = cta_first + (unsigned)(it * kPasteThreads + threadIdx.x) * kPx;
    if (p0 >= plane) break;
    int y = (int)(p0 / (unsigned)a.W), x = (int)(p0 - (unsigned)y * (unsigned)a.W);
    unsigned pk[4] = {0u, 0u, 0u, 0u};
    // A run of kPx pixels inside one row whose row, or whose whole column span, misses the mask is zeros by the
    // extrapolation rule: decided from the row coordinate and the two END columns (at() is monotone in the index, see
    // above; a NaN coordinate compares false and takes the per-pixel path).  With ~1 % of the pixels inside a box the
    // per-pixel loop was the bound (84 % issue-active at 2.3 TB/s written); now it only runs where a box is.
    bool work = any;
    if (any && x + kPx <= a.W) {
      const float in_y = ay.at(y);
      if (in_y < 0.0f || in_y > ymax_in) work = false;
      else {
        const float c0 = ax.at(x), c1 = ax.at(x + kPx - 1);
        if (fmaxf(c0, c1) < 0.0f || fminf(c0, c1) > xmax_in) work = (c0 != c0) || (c1 != c1);
      }
    }
    if (work) {
      int top = 0, bot = 0;
      float ly = 0.0f;
      bool vy = false;
      int cur_y = -1;
      float xf = (float)x;
#pragma unroll
      for (int j = 0; j < kPx; ++j) {
        if (p0 + j < plane) {
          if (y != cur_y) {
            const float in_y = ay.at(y);
            vy = in_y >= 0.0f && in_y <= ymax_in;
            const float f = floorf(in_y);
            top = (int)f; bot = (int)ceilf(in_y);
            ly = in_y - f;
            cur_y = y;
          }
          if (vy) {
            const float in_x = ax.at_f(xf);
            if (in_x >= 0.0f && in_x <= xmax_in) {
              // floor / ceil of a value in [0, mw - 1]: truncation, and floor + 1 unless the value is an integer
              const int left = (int)in_x;
              const float f = (float)left;
              const float lx = in_x - f;
              const int right = left + (lx > 0.0f ? 1 : 0);
              const float tl = mk[top * a.mw + left], tr = mk[top * a.mw + right];
              const float bl = mk[bot * a.mw + left], br = mk[bot * a.mw + right];
              float t = tr - tl; t = t * lx; t = tl + t;
              float bb = br - bl; bb = bb * lx; bb = bl + bb;
              float r = bb - t; r = r * ly; r = t + r;
              if (r > a.thr) pk[j >> 2] |= 1u << (8 * (j & 3));
            }
          }
          if (++x == a.W) { x = 0; ++y; xf = 0.0f; } else { xf = xf + 1.0f; }
        }
      }
    }
    if (a.vec && p0 + kPx <= plane) {
      __stcs(reinterpret_cast<uint4*>(o + p0), make_uint4(pk[0], pk[1], pk[2], pk[3]));
    } else {
#pragma unroll
      for (int j = 0; j < kPx; ++j)  // (static indices: a run-time index would push pk[] to local memory)
        if (p0 + j < plane) o[p0 + j] = (uint8_t)((pk[j >> 2] >> (8 * (j & 3))) & 0xffu);
    }
  }
}

// ---- two-phase variant (the default): the output is zero-filled at memset speed (7.3 TB/s on a B200; the fused kernel
// above wrote 2.3-2.5 TB/s because the warps that straddle a box run the 16-pixel body with most lanes idle), then
// this kernel visits only a conservative bounding rectangle of each box -- the index range the inverse of the affine
// crop_and_resize map gives, widened by 2 -- and applies the SAME exact per-pixel rule there, storing the ones.
#ifndef D2B_PASTE_RECT_CTAS
#define D2B_PASTE_RECT_CTAS 8  // A/B at 1,600 masks: 32 -> 0.515 ms, 16 -> 0.452, 8 -> 0.431, 4 -> 0.441, 2 -> 0.476
#endif
constexpr int kRectCtas = D2B_PASTE_RECT_CTAS;  // CTAs per mask: rows y0 + blockIdx.x, + kRectCtas, ...

// index range [lo, hi] of the output axis (n samples) whose input coordinate can fall inside [0, max_in]
__device__ __forceinline__ void axis_range(const Axis& ax, int n, float max_in, int& lo, int& hi) {
  lo = 0; hi = n - 1;
  if (ax.crop <= 1) return;
  const float c = ax.n1 * (float)(ax.dim - 1), st = ax.step;
  if (!(st > 0.0f) && !(st < 0.0f)) return;                  // 0 or NaN: every index has the same / no coordinate
  if (!(fabsf(c) < 1e30f) || !(fabsf(st) < 1e30f)) return;   // inf / NaN: leave it to the exact per-pixel rule
  float a = (0.0f - c) / st, b = (max_in - c) / st;
  if (a > b) { const float t = a; a = b; b = t; }
  a = fminf(fmaxf(a, -4.0f), (float)n + 4.0f);
  b = fminf(fmaxf(b, -4.0f), (float)n + 4.0f);
  lo = max(0, (int)floorf(a) - 2);
  hi = min(n - 1, (int)ceilf(b) + 2);
}

__global__ void __launch_bounds__(kPasteThreads) paste_boxes_kernel(PasteArgs a) {
  extern __shared__ float s_mask[];
  const long long m = blockIdx.y;
  const float4 bx = __ldg(reinterpret_cast<const float4*>(a.boxes) + m);
  const float ys = 1.0f / (float)a.H, xs = 1.0f / (float)a.W;
  const Axis ay = make_axis(ys * bx.x, ys * bx.z, a.H, a.mh);
  const Axis ax = make_axis(xs * bx.y, xs * bx.w, a.W, a.mw);
  const float ymax_in = (float)(a.mh - 1), xmax_in = (float)(a.mw - 1);
  int y0, y1, x0, x1;
  axis_range(ay, a.H, ymax_in, y0, y1);
  axis_range(ax, a.W, xmax_in, x0, x1);
  if (y0 + (int)blockIdx.x > y1 || x0 > x1) return;  // block-uniform
  const float* mk = a.masks + (size_t)m * a.mh * a.mw;
  if (a.smem_mask) {
    for (int i = threadIdx.x; i < a.mh * a.mw; i += kPasteThreads) s_mask[i] = __ldg(mk + i);
    __syncthreads();
    mk = s_mask;
  }
  uint8_t* o = a.out + (size_t)m * ((size_t)a.H * a.W);
  for (int y = y0 + (int)blockIdx.x; y <= y1; y += kRectCtas) {
    const float in_y = ay.at(y);
    if (!(in_y >= 0.0f && in_y <= ymax_in)) continue;
    const float fy = floorf(in_y);
    const int top = (int)fy, bot = (int)ceilf(in_y);
    const float ly = in_y - fy;
    for (int x = x0 + (int)threadIdx.x; x <= x1; x += kPasteThreads) {
      const float in_x = ax.at(x);
      if (in_x >= 0.0f && in_x <= xmax_in) {
        const float f = floorf(in_x);
        const int left = (int)f, right = (int)ceilf(in_x);
        const float lx = in_x - f;
        const float tl = mk[top * a.mw + left], tr = mk[top * a.mw + right];
        const float bl = mk[bot * a.mw + left], br = mk[bot * a.mw + right];
        float t = tr - tl; t = t * lx; t = tl + t;
        float bb = br - bl; bb = bb * lx; bb = bl + bb;
        float r = bb - t; r = r * ly; r = t + r;
        if (r > a.thr) o[(size_t)y * a.W + x] = 1;
      }
    }
  }
}

}  // namespace
}  // namespace d2b

using namespace d2b;

extern "C" size_t d2b_paste_masks_workspace_bytes(const d2b_paste_masks_params*) { return 0; }

extern "C" int d2b_paste_masks(const d2b_paste_masks_params* p, void*, size_t, d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->num_masks >= 0 && p->num_masks <= 65535ll * 16, "paste_masks: num_masks=%lld out of range",
              (long long)p->num_masks);
  D2B_REQUIRE(p->mask_h >= 1 && p->mask_w >= 1 && p->image_h >= 1 && p->image_w >= 1, "paste_masks: bad sizes");
  D2B_REQUIRE((long long)p->image_h * p->image_w < (1ll << 31) - 65536, "paste_masks: image too large");
  if (p->num_masks == 0) return D2B_OK;  // mask_ops.py:25-28: zeros [0, H, W]
  D2B_REQUIRE(p->box_masks && p->boxes && p->out, "paste_masks: NULL pointer");
  PasteArgs a;
  a.masks = p->box_masks; a.boxes = p->boxes; a.mh = p->mask_h; a.mw = p->mask_w;
  a.H = p->image_h; a.W = p->image_w; a.thr = p->mask_threshold;
  const long long plane = (long long)a.H * a.W;
  a.vec = (plane % 16 == 0) && ((reinterpret_cast<uintptr_t>(p->out) & 15) == 0);
  a.smem_mask = (a.mh * a.mw <= kMaxMaskSmem);
  const size_t smem = a.smem_mask ? sizeof(float) * a.mh * a.mw : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (smem > 48 * 1024)
    D2B_CUDA(cudaFuncSetAttribute(paste_masks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned gx = (unsigned)((plane + kPasteThreads * kPx * kIters - 1) / (kPasteThreads * kPx * kIters));
  const char* fused_env = getenv("D2B_PASTE_FUSED");  // tests / A-B: the one-pass kernel
  const bool two_phase = !(fused_env && fused_env[0] == '1');
  if (two_phase) D2B_CUDA(cudaMemsetAsync(p->out, 0, (size_t)p->num_masks * (size_t)plane, st));
  if (two_phase && smem > 48 * 1024)
    D2B_CUDA(cudaFuncSetAttribute(paste_boxes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (long long m0 = 0; m0 < p->num_masks; m0 += 65535) {  // grid.y limit
    const long long cnt = p->num_masks - m0 < 65535 ? p->num_masks - m0 : 65535;
    a.masks = p->box_masks + (size_t)m0 * a.mh * a.mw;
    a.boxes = p->boxes + 4 * m0;
    a.out = p->out + (size_t)m0 * plane;
    a.M = cnt;
    if (two_phase) paste_boxes_kernel<<<dim3(kRectCtas, (unsigned)cnt), kPasteThreads, smem, st>>>(a);
    else paste_masks_kernel<<<dim3(gx, (unsigned)cnt), kPasteThreads, smem, st>>>(a);
    D2B_LAUNCH_CHECK();
  }
  return D2B_OK;
}

// ------------------------------------------------------------------ mask_rcnn_inference (mask_head.py:71-103)
//   pred_mask_logits [M, Hm, Wm, C] (NHWC) -> transpose to NCHW -> gather_nd([m, pred_classes[m]]) -> sigmoid
// The reference transposes the whole tensor (C = 80: 400 MB at 1,600 masks of 28x28) to pick one channel per mask;
// here a thread reads exactly the element it needs (one 32-byte sector per pixel) and applies the bit-exact sigmoid.
namespace d2b {
namespace {
__global__ void mask_select_kernel(const float* logits, const long long* classes, long long total, int hw, int C,
                                   float* out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long m = t / hw;
  long long c = C == 1 ? 0 : classes[m];
  float v = 0.0f;  // tf.gather_nd on the GPU returns 0 for an out-of-range index
  if (c >= 0 && c < C) v = __ldg(logits + t * C + c);
  out[t] = d2b_sigmoidf(v);
}
}  // namespace
}  // namespace d2b

extern "C" size_t d2b_mask_rcnn_inference_workspace_bytes(const d2b_mask_rcnn_inference_params*) { return 0; }
extern "C" int d2b_mask_rcnn_inference(const d2b_mask_rcnn_inference_params* p, void*, size_t, d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->num_masks >= 0 && p->mask_h >= 0 && p->mask_w >= 0 && p->num_classes >= 1, "mask_rcnn_inference: bad sizes");
  const long long total = (long long)p->num_masks * p->mask_h * p->mask_w;
  if (total == 0) return D2B_OK;
  D2B_REQUIRE(p->mask_logits && p->out && (p->num_classes == 1 || p->pred_classes), "mask_rcnn_inference: NULL pointer");
  D2B_REQUIRE(total < (1ll << 31) * 256, "mask_rcnn_inference: too many elements");
  d2b::mask_select_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p->mask_logits, reinterpret_cast<const long long*>(p->pred_classes), total, p->mask_h * p->mask_w, p->num_classes, p->out);
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}
