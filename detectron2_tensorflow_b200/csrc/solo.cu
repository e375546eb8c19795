// solo.cu -- SOLOv2 neighbours of Matrix-NMS (SURVEY.md 8f "next" #4).
#include "kernels.cuh"

// ------------------------------------------------------------------ point_nms (solo_v2.py:29-40)
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
