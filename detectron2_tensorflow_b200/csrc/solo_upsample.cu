// solo_upsample.cu -- what MaskKernelBranch.inference does after the per-image tail (solo_v2.py:599-627):
//   resize_images(pred_masks, image_shape) (bilinear)  ->  > mask_threshold  ->  boxes from masks.
// The reference materialises [N, D, H, W] fp32 twice (resize output, thresholded masks) plus the yy / xx products
// (427 MB per image each at D=100, 800x1333).  Here the kept masks arrive BIT-PACKED from d2b_solo_postprocess; one
// kernel re-samples them, thresholds, writes uint8 (or packed) image masks and reduces the box statistics on the fly.
//   * flat 16-pixel runs per thread (16-byte stores); the few source rows a CTA needs are expanded to fp32 0/1 in
//     shared memory together with each row's [first, last] set column and the per-output-column taps;
//   * a run whose source window is empty in both rows is written as zeros without sampling (most of an image);
//   * box statistics are exact integers (count, sum y, sum x, min / max over pixels with y > 0 / x > 0), reduced per
//     warp and accumulated with 64-bit atomics; a finalise kernel applies the reference's mean / where rule.
// TF ResizeBilinear arithmetic in the reference's op order (no FMA), both coordinate conventions the reference can
// end up with (functional.py:21-35): half-pixel centres (tf.compat.v2.image.resize) and align_corners=True.
#include <math.h>

#include "kernels.cuh"

namespace d2b {
namespace {
typedef unsigned long long u64;

constexpr int kUpThreads = 256;
constexpr int kRun = 16;  // output pixels per thread per step

struct Tap {
  unsigned short lo, hi;
  float lerp;
};
struct Stats {  // per (image, detection); zero-initialised except the min fields (0x7fffffff)
  u64 count, sum_y, sum_x;
  int min_y, max_y, min_x, max_x;
};
struct UpArgs {
  const u64* packed;
  int B, D, h, w, H, W, Wd_in, Wd_out;
  int align_corners;
  float scale_y, scale_x, thr;
  int iters;     // runs per thread: a CTA covers iters * 256 * 16 flat output pixels
  int rows_cap;  // source rows staged per CTA
  uint8_t* out_masks;
  u64* out_packed;
  Stats* stats;
};

__device__ __forceinline__ void resize_tap(int i, int in_size, float scale, int align_corners, int& lo, int& hi, float& lerp) {
  float src;
  if (align_corners) {
    src = (float)i * scale;
  } else {
    src = (float)i + 0.5f;
    src = src * scale;
    src = src - 0.5f;
  }
  const float f = floorf(src);
  lo = max((int)f, 0);
  hi = min((int)ceilf(src), in_size - 1);
  lerp = src - f;
}

__global__ void up_init_stats(Stats* s, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Stats z;
  z.count = z.sum_y = z.sum_x = 0ull;
  z.min_y = z.min_x = 0x7fffffff;
  z.max_y = z.max_x = -1;
  s[i] = z;
}

__global__ void __launch_bounds__(kUpThreads) solo_upsample_kernel(const UpArgs a) {
  extern __shared__ __align__(16) unsigned char up_smem[];
  float* rows = reinterpret_cast<float*>(up_smem);                                 // [rows_cap][w]
  int2* span = reinterpret_cast<int2*>(rows + (((size_t)a.rows_cap * a.w + 1) & ~(size_t)1));  // [rows_cap] first / last set column
  Tap* xt = reinterpret_cast<Tap*>(span + a.rows_cap);                             // [W]

  const int b = blockIdx.z, d = blockIdx.y;
  const long long hw_out = (long long)a.H * a.W;
  const long long chunk = (long long)a.iters * kUpThreads * kRun;
  const long long p_begin = (long long)blockIdx.x * chunk;
  const long long p_end = min(p_begin + chunk, hw_out);
  const int y_first = (int)(p_begin / a.W), y_last = (int)((p_end - 1) / a.W);
  int r0, r1, tmp;
  float tf;
  resize_tap(y_first, a.h, a.scale_y, a.align_corners, r0, tmp, tf);
  resize_tap(y_last, a.h, a.scale_y, a.align_corners, tmp, r1, tf);
  const int n_rows = r1 - r0 + 1;  // <= rows_cap by construction (host)
  const u64* src = a.packed + ((size_t)b * a.D + d) * a.Wd_in;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* om = a.out_masks ? a.out_masks + ((size_t)b * a.D + d) * hw_out : nullptr;
  u64* op = a.out_packed ? a.out_packed + ((size_t)b * a.D + d) * a.Wd_out : nullptr;
  const bool vec16 = om && ((reinterpret_cast<uintptr_t>(om) & 15) == 0);  // flat runs start at multiples of 16
  // every shortcut below assumes "all four corners 0 -> off, all four corners 1 -> on", i.e. 0 <= thr < 1
  const bool shortcuts = a.thr >= 0.0f && a.thr < 1.0f;
  if (shortcuts) {
    // ---- cheap rejection first: OR of the words that hold source rows r0..r1 (edge words may carry neighbouring
    // rows' bits: conservative).  Most chunks of an image see no mask at all and leave as zeros right here.
    const long long w_first = ((long long)r0 * a.w) >> 6, w_last = (((long long)(r1 + 1) * a.w - 1) >> 6);
    u64 acc = 0;
    for (long long q = w_first + tid; q <= w_last; q += kUpThreads) acc |= __ldg(src + q);
    if (!__syncthreads_or(acc != 0ull)) {
      const int len = (int)(p_end - p_begin);
      if (om) {
        if (vec16) {  // p_begin is a multiple of 16: whole 16-byte stores, then the (image-end) tail
          uint4* dst = reinterpret_cast<uint4*>(om + p_begin);
          for (int v = tid; v < (len >> 4); v += kUpThreads) dst[v] = make_uint4(0, 0, 0, 0);
          for (int q = (len & ~15) + tid; q < len; q += kUpThreads) om[p_begin + q] = 0;
        } else {
          for (int q = tid; q < len; q += kUpThreads) om[p_begin + q] = 0;
        }
      }
      if (op) {
        u64* dst = op + (p_begin >> 6);  // p_begin is a multiple of 64
        for (int v = tid; v < ((len + 63) >> 6); v += kUpThreads) dst[v] = 0ull;
      }
      return;
    }
  }

  // ---- stage the source rows: bits -> fp32 0/1, and each row's span of set columns
  for (int r = warp; r < n_rows; r += kUpThreads / 32) {
    const long long bit0 = (long long)(r0 + r) * a.w;
    int first = 0x7fffffff, last = -1;
    for (int x = lane; x < a.w; x += 32) {
      const long long q = bit0 + x;
      const unsigned on = (unsigned)((__ldg(src + (q >> 6)) >> (q & 63)) & 1ull);
      rows[(size_t)r * a.w + x] = on ? 1.0f : 0.0f;
      if (on) { first = min(first, x); last = max(last, x); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    }
    if (lane == 0) span[r] = make_int2(first, last);
  }
  __syncthreads();
  for (int x = tid; x < a.W; x += kUpThreads) {
    int lo, hi;
    float l;
    resize_tap(x, a.w, a.scale_x, a.align_corners, lo, hi, l);
    Tap t;
    t.lo = (unsigned short)lo; t.hi = (unsigned short)hi; t.lerp = l;
    xt[x] = t;
  }
  __syncthreads();

  // per-thread statistics fit 32 bits (<= iters * 16 pixels per thread, coordinates < 65536); widened at the reduction
  unsigned cnt32 = 0, sum_y32 = 0, sum_x32 = 0;
  int min_y = 0x7fffffff, max_y = -1, min_x = 0x7fffffff, max_x = -1;
  // A warp owns 512 consecutive flat pixels per step; in sub-step j lane l samples pixel base + 32 j + l, so the tap
  // table and the source rows are read without bank conflicts (neighbouring pixels share or neighbour their taps) and
  // the 32 decisions leave as one ballot = 32 packed mask bits.  Lane j keeps ballot j; a shuffle then hands every
  // lane the 16 bits of ITS 16-pixel run for the 16-byte uint8 store.
  const int chunk_len = (int)(p_end - p_begin);
  for (int it = 0; it < a.iters; ++it) {
    const int off = (it * (kUpThreads / 32) + warp) * (32 * kRun);  // offset of the warp's 512 pixels in the chunk
    if (off >= chunk_len) break;  // warp-uniform
    const long long base = p_begin + off;
    const int span_len = min(32 * kRun, chunk_len - off);
    const int ya = (int)(base / a.W), xa = (int)(base - (long long)ya * a.W);
    const int yb = (int)((base + span_len - 1) / a.W);
    unsigned mine = 0;
    // The span is cut by output row (1-2 rows for W >= 512): everything that depends on the row -- its two source
    // rows, their lerp weight and set-column spans -- is warp-uniform inside the sub-step loop.
    for (int y = ya; y <= yb; ++y) {
      const int f0 = y == ya ? 0 : (y - ya) * a.W - xa;                 // flat range of this row inside the span
      const int f1 = min(span_len, (y - ya + 1) * a.W - xa);
      int ylo, yhi;
      float yl;
      resize_tap(y, a.h, a.scale_y, a.align_corners, ylo, yhi, yl);
      const int2 s0 = span[ylo - r0], s1 = span[yhi - r0];
      const int xbase = xa - (y - ya) * a.W;                             // x = xbase + f
      {  // row-level rejection over the columns this row's part can touch
        const int c0 = xt[xbase + f0].lo, c1 = xt[xbase + f1 - 1].hi;
        if (shortcuts && (s0.y < c0 || s0.x > c1) && (s1.y < c0 || s1.x > c1)) continue;
      }
      const float* q0 = rows + (ylo - r0) * a.w;
      const float* q1 = rows + (yhi - r0) * a.w;
      unsigned row_cnt = 0;
      for (int j = f0 >> 5; j <= (f1 - 1) >> 5; ++j) {
        const int f = lane + 32 * j;
        bool on = false;
        if (f >= f0 && f < f1) {
          const int x = xbase + f;
          const Tap t = xt[x];
          const int lo = t.lo, hi = t.hi;
          // (tried: a warp-uniform all-zeros / all-ones test of the 32 pixels' source window on row-aligned bit words,
          //  so that only mask edges are sampled: its ~45 instructions per sub-step cost more than they save -- object
          //  masks 0.81 -> 1.23 ms, noise 4.2 -> 7.6 ms)
          const bool empty = shortcuts && (s0.y < lo || s0.x > hi) && (s1.y < lo || s1.x > hi);
          if (!empty) {
            const float tl = q0[lo], tr = q0[hi], bl = q1[lo], br = q1[hi];
            float top = tr - tl; top = top * t.lerp; top = tl + top;
            float bot = br - bl; bot = bot * t.lerp; bot = bl + bot;
            float v = bot - top; v = v * yl; v = top + v;
            on = v > a.thr;
            if (on) {
              row_cnt += 1; sum_x32 += (unsigned)x;
              if (x > 0) { min_x = min(min_x, x); max_x = max(max_x, x); }
            }
          }
        }
        const unsigned word = __ballot_sync(0xffffffffu, on);  // pixels base + 32 j .. + 31 (this row's part)
        if (lane == j) mine |= word;
      }
      cnt32 += row_cnt;
      sum_y32 += row_cnt * (unsigned)y;
      if (row_cnt && y > 0) { min_y = min(min_y, y); max_y = max(max_y, y); }
    }
    if (op) {  // lanes 0..15 hold the 16 ballots: two neighbours make one 64-bit word
      const unsigned hi32 = __shfl_down_sync(0xffffffffu, mine, 1);
      const long long pw = base + 64ll * (lane >> 1);
      if (lane < kRun && (lane & 1) == 0 && pw < p_end) op[pw >> 6] = (u64)mine | ((u64)hi32 << 32);
    }
    if (om) {
      const unsigned src_word = __shfl_sync(0xffffffffu, mine, lane >> 1);
      const unsigned bits = (src_word >> (16 * (lane & 1))) & 0xffffu;
      const long long pr = base + (long long)lane * kRun;  // this lane's 16-pixel run
      if (pr < p_end) {
        const int n = (int)min((long long)kRun, p_end - pr);
        if (vec16 && n == kRun) {
          uint4 v4;
          unsigned* wv = reinterpret_cast<unsigned*>(&v4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const unsigned nib = (bits >> (4 * q)) & 15u;
            wv[q] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
          }
          *reinterpret_cast<uint4*>(om + pr) = v4;
        } else {
          for (int q = 0; q < n; ++q) om[pr + q] = (uint8_t)((bits >> q) & 1u);
        }
      }
    }
  }
  // ---- box statistics: warp reduce, then one set of atomics per warp
  u64 cnt = cnt32, sum_y = sum_y32, sum_x = sum_x32;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sum_y += __shfl_xor_sync(0xffffffffu, sum_y, o);
    sum_x += __shfl_xor_sync(0xffffffffu, sum_x, o);
    min_y = min(min_y, __shfl_xor_sync(0xffffffffu, min_y, o));
    max_y = max(max_y, __shfl_xor_sync(0xffffffffu, max_y, o));
    min_x = min(min_x, __shfl_xor_sync(0xffffffffu, min_x, o));
    max_x = max(max_x, __shfl_xor_sync(0xffffffffu, max_x, o));
  }
  if (lane == 0 && cnt) {
    Stats* s = a.stats + (size_t)b * a.D + d;
    atomicAdd(&s->count, cnt);
    atomicAdd(&s->sum_y, sum_y);
    atomicAdd(&s->sum_x, sum_x);
    if (max_y >= 0) { atomicMin(&s->min_y, min_y); atomicMax(&s->max_y, max_y); }
    if (max_x >= 0) { atomicMin(&s->min_x, min_x); atomicMax(&s->max_x, max_x); }
  }
}

// solo_v2.py:606-625: mean = sum / (count + 1e-5); where(yy > 0, yy, mean); min / max over all pixels
__global__ void up_boxes_kernel(const Stats* stats, int n, float* boxes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Stats s = stats[i];
  const float den = (float)s.count + 1e-5f;
  const float ymean = (float)s.sum_y / den, xmean = (float)s.sum_x / den;
  const float ymin = s.max_y >= 0 ? (float)s.min_y : INFINITY, ymax = s.max_y >= 0 ? (float)s.max_y : -INFINITY;
  const float xmin = s.max_x >= 0 ? (float)s.min_x : INFINITY, xmax = s.max_x >= 0 ? (float)s.max_x : -INFINITY;
  boxes[i * 4 + 0] = fminf(ymin, ymean);
  boxes[i * 4 + 1] = fminf(xmin, xmean);
  boxes[i * 4 + 2] = fmaxf(ymax, ymean);
  boxes[i * 4 + 3] = fmaxf(xmax, xmean);
}

struct UpPlan {
  int iters, rows_cap;
  size_t smem;
  float scale_y, scale_x;
};
int up_plan(const d2b_solo_upsample_params* p, UpPlan& pl) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->batch >= 0 && p->num_dets >= 0, "solo_upsample: negative sizes");
  D2B_REQUIRE(p->batch <= 65535 && p->num_dets <= 65535, "solo_upsample: batch / num_dets too large");
  D2B_REQUIRE(p->mask_h >= 1 && p->mask_w >= 1 && p->image_h >= 1 && p->image_w >= 1, "solo_upsample: empty mask / image");
  D2B_REQUIRE(p->mask_w <= 65535 && p->mask_h <= 65535 && (long long)p->image_h * p->image_w < (1ll << 31),
              "solo_upsample: sizes out of range");
  const int ac = p->align_corners != 0;
  pl.scale_y = (ac && p->image_h > 1) ? (float)(p->mask_h - 1) / (float)(p->image_h - 1) : (float)p->mask_h / (float)p->image_h;
  pl.scale_x = (ac && p->image_w > 1) ? (float)(p->mask_w - 1) / (float)(p->image_w - 1) : (float)p->mask_w / (float)p->image_w;
  for (int iters = 8; iters >= 1; iters >>= 1) {
    const long long chunk = (long long)iters * kUpThreads * kRun;
    const long long out_rows = chunk / p->image_w + 2;
    long long rows = (long long)ceil((double)out_rows * (double)pl.scale_y) + 3;
    if (rows > p->mask_h) rows = p->mask_h;
    const size_t smem = (((size_t)rows * p->mask_w + 1) & ~(size_t)1) * 4 + (size_t)rows * 8 + (size_t)p->image_w * sizeof(Tap) + 16;
    if (smem <= 96 * 1024 || iters == 1) {
      pl.iters = iters;
      pl.rows_cap = (int)rows;
      pl.smem = smem;
      D2B_REQUIRE(smem <= 200 * 1024, "solo_upsample: a %d x %d -> %d x %d resize needs %zu bytes of shared memory per CTA",
                  p->mask_h, p->mask_w, p->image_h, p->image_w, smem);
      return D2B_OK;
    }
  }
  return D2B_EINVAL;
}
}  // namespace
}  // namespace d2b

using namespace d2b;

extern "C" size_t d2b_solo_upsample_workspace_bytes(const d2b_solo_upsample_params* p) {
  UpPlan pl;
  if (up_plan(p, pl) != D2B_OK) return 0;
  return ws_slice(sizeof(Stats) * (size_t)p->batch * (p->num_dets > 0 ? p->num_dets : 1));
}

extern "C" int d2b_solo_upsample(const d2b_solo_upsample_params* p, void* workspace, size_t workspace_bytes,
                                 d2b_stream_t stream) {
  UpPlan pl;
  int rc = up_plan(p, pl);
  if (rc != D2B_OK) return rc;
  const int n = p->batch * p->num_dets;
  if (n == 0) return D2B_OK;
  D2B_REQUIRE(p->packed_masks && p->out_boxes, "solo_upsample: NULL pointer");
  const size_t need = ws_slice(sizeof(Stats) * (size_t)n);
  if (workspace == nullptr || workspace_bytes < need) {
    set_last_error("solo_upsample needs %zu workspace bytes", need);
    return D2B_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UpArgs a;
  a.packed = reinterpret_cast<const u64*>(p->packed_masks);
  a.B = p->batch; a.D = p->num_dets; a.h = p->mask_h; a.w = p->mask_w; a.H = p->image_h; a.W = p->image_w;
  a.Wd_in = (int)(((long long)p->mask_h * p->mask_w + 63) / 64);
  a.Wd_out = (int)(((long long)p->image_h * p->image_w + 63) / 64);
  a.align_corners = p->align_corners != 0;
  a.scale_y = pl.scale_y; a.scale_x = pl.scale_x; a.thr = p->mask_threshold;
  a.iters = pl.iters; a.rows_cap = pl.rows_cap;
  a.out_masks = p->out_masks;
  a.out_packed = reinterpret_cast<u64*>(p->out_packed_masks);
  a.stats = static_cast<Stats*>(workspace);
  up_init_stats<<<(n + 255) / 256, 256, 0, st>>>(a.stats, n);
  D2B_LAUNCH_CHECK();
  if (pl.smem > 48 * 1024) {
    static size_t set_to = 0;  // grow-only opt-in (any device: the attribute is per function per device, re-setting is cheap)
    if (pl.smem > set_to) set_to = pl.smem;
    D2B_CUDA(cudaFuncSetAttribute(solo_upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)set_to));
  }
  const long long chunk = (long long)pl.iters * kUpThreads * kRun;
  const long long hw_out = (long long)p->image_h * p->image_w;
  const dim3 grid((unsigned)((hw_out + chunk - 1) / chunk), (unsigned)p->num_dets, (unsigned)p->batch);
  solo_upsample_kernel<<<grid, kUpThreads, pl.smem, st>>>(a);
  D2B_LAUNCH_CHECK();
  up_boxes_kernel<<<(n + 255) / 256, 256, 0, st>>>(a.stats, n, p->out_boxes);
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}
