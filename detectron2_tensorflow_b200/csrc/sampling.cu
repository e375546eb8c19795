// sampling.cu -- subsample_labels (lib/modeling/sampling.py:6-45) for a batch of label vectors.
//
// The reference draws `tf.random_shuffle(positive)[:num_pos]` and `tf.random_shuffle(negative)[:num_neg]`: a uniform
// random subset of each class, in random order.  TF's shuffle has no defined bit pattern, so the parity contract
// here is distributional, with a documented counter-based generator (restated in numpy by the test oracle):
//     u(seed, image n, element i, stream s) = top 24 bits of splitmix64(seed + n * C1 + i * C2 + s * C3) * 2^-24
// Every eligible element gets such a score; the sample = the elements with the LARGEST scores (ties -> lower
// index), i.e. the prefix of a random permutation -- selected with the library's segmented top-k (the same radix
// select + sort as the proposal stage), so the result is deterministic per seed and independent of launch geometry.
#include "kernels.cuh"

namespace d2b {
namespace {

typedef unsigned long long u64;

__device__ __forceinline__ float sample_score(u64 seed, u64 n, u64 i, u64 stream) {
  u64 z = seed + n * 0x9E3779B97F4A7C15ull + i * 0xBF58476D1CE4E5B9ull + stream * 0x94D049BB133111EBull;
  z += 0x9E3779B97F4A7C15ull;  // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(unsigned)(z >> 40) * 5.9604644775390625e-08f;  // [0, 1), 24 bits: exact in fp32
}

// scores[0] = positives (label != -1 && label != bg), scores[1] = negatives (label == bg); -inf elsewhere.
__global__ void sub_scores_kernel(const long long* labels, int N, long long P, long long bg, u64 seed, float* pos,
                                  float* neg, int32_t* num_pos, int32_t* num_neg) {
  const int n = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float ninf = __int_as_float(0xff800000);
  bool is_pos = false, is_neg = false;
  if (i < P) {
    const long long l = labels[(size_t)n * P + i];
    is_pos = (l != -1) && (l != bg);
    is_neg = (l == bg);
    pos[(size_t)n * P + i] = is_pos ? sample_score(seed, (u64)n, (u64)i, 0) : ninf;
    neg[(size_t)n * P + i] = is_neg ? sample_score(seed, (u64)n, (u64)i, 1) : ninf;
  }
  const unsigned mp = __ballot_sync(0xffffffffu, is_pos), mn = __ballot_sync(0xffffffffu, is_neg);
  if ((threadIdx.x & 31) == 0) {
    if (mp) atomicAdd(num_pos + n, __popc(mp));
    if (mn) atomicAdd(num_neg + n, __popc(mn));
  }
}

// per image: num_pos = min(#pos, int(num_samples * fraction)), num_neg = min(#neg, num_samples - num_pos)
// (sampling.py:37-42); index lists padded with -1; optional resampled label vector (rpn_outputs.py:315-329).
__global__ void sub_emit_kernel(const long long* labels, long long P, int k, int want_pos, const int32_t* cnt_pos,
                                const int32_t* cnt_neg, const int32_t* pos_idx, const int32_t* neg_idx,
                                long long* out_pos, long long* out_neg, int32_t* out_num_pos, int32_t* out_num_neg,
                                long long* out_labels) {
  const int n = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = min(cnt_pos[n], want_pos);
  const int nn = min(cnt_neg[n], k - np);
  if (j == 0) {
    if (out_num_pos) out_num_pos[n] = np;
    if (out_num_neg) out_num_neg[n] = nn;
  }
  if (j >= k) return;
  const long long ip = j < np ? (long long)pos_idx[(size_t)n * k + j] : -1;
  const long long in = j < nn ? (long long)neg_idx[(size_t)n * k + j] : -1;
  if (out_pos) out_pos[(size_t)n * k + j] = ip;
  if (out_neg) out_neg[(size_t)n * k + j] = in;
  if (out_labels) {
    if (ip >= 0) out_labels[(size_t)n * P + ip] = labels[(size_t)n * P + ip];
    if (in >= 0) out_labels[(size_t)n * P + in] = labels[(size_t)n * P + in];
  }
}

__global__ void sub_fill_kernel(long long* out, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = -1;
}

struct SubPlan {
  TopkDesc td;
  int k, P2;
  size_t bytes, o_pos, o_neg, o_cnt, o_keys, o_idx_pos, o_idx_neg, o_topk;
};

int sub_plan(const d2b_subsample_labels_params* p, SubPlan& pl) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->num_images >= 0 && p->num_labels >= 0, "subsample_labels: negative sizes");
  D2B_REQUIRE(p->num_samples >= 1 && p->num_samples <= kTopkMaxK, "subsample_labels: num_samples=%d out of [1,%d]",
              p->num_samples, kTopkMaxK);
  D2B_REQUIRE(p->max_positives >= 0 && p->max_positives <= p->num_samples, "max_positives must be in [0, num_samples]");
  D2B_REQUIRE(p->num_labels < (1ll << 31), "subsample_labels: too many labels per image");
  pl.k = p->num_samples;
  pl.P2 = topk_padded_k(pl.k);
  TopkDesc& td = pl.td;
  for (int g = 0; g < D2B_MAX_LEVELS; ++g) { td.scores[g] = nullptr; td.row_len[g] = 0; td.k_limit[g] = 0; }
  td.G = 1; td.rows_per_group = p->num_images; td.k = pl.k; td.transform = D2B_TOPK_IDENTITY;
  td.row_len[0] = p->num_labels;
  const size_t N = p->num_images, P = p->num_labels;
  size_t o = 0;
  pl.o_pos = o; o += ws_slice(N * P * sizeof(float));
  pl.o_neg = o; o += ws_slice(N * P * sizeof(float));
  pl.o_cnt = o; o += ws_slice(2 * N * sizeof(int32_t));
  pl.o_keys = o; o += ws_slice(N * pl.P2 * sizeof(u64));
  pl.o_idx_pos = o; o += ws_slice(N * pl.k * sizeof(int32_t));
  pl.o_idx_neg = o; o += ws_slice(N * pl.k * sizeof(int32_t));
  pl.o_topk = o; o += topk_workspace_bytes(td);
  pl.bytes = o;
  return D2B_OK;
}

}  // namespace
}  // namespace d2b

using namespace d2b;

extern "C" size_t d2b_subsample_labels_workspace_bytes(const d2b_subsample_labels_params* p) {
  SubPlan pl;
  if (sub_plan(p, pl) != D2B_OK) return 0;
  return pl.bytes;
}

extern "C" int d2b_subsample_labels(const d2b_subsample_labels_params* p, void* workspace, size_t workspace_bytes,
                                    d2b_stream_t stream) {
  SubPlan pl;
  int rc = sub_plan(p, pl);
  if (rc != D2B_OK) return rc;
  if (p->num_images == 0) return D2B_OK;
  D2B_REQUIRE(p->num_labels == 0 || p->labels, "subsample_labels: labels is NULL");
  if (workspace == nullptr || workspace_bytes < pl.bytes) {
    set_last_error("subsample_labels needs %zu workspace bytes", pl.bytes);
    return D2B_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  const int N = p->num_images, k = pl.k;
  const long long P = p->num_labels;
  float* pos = reinterpret_cast<float*>(ws + pl.o_pos);
  float* neg = reinterpret_cast<float*>(ws + pl.o_neg);
  int32_t* cnt = reinterpret_cast<int32_t*>(ws + pl.o_cnt);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + pl.o_keys);
  int32_t* idx_pos = reinterpret_cast<int32_t*>(ws + pl.o_idx_pos);
  int32_t* idx_neg = reinterpret_cast<int32_t*>(ws + pl.o_idx_neg);
  D2B_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(int32_t) * N, st));
  if (P > 0) {
    sub_scores_kernel<<<dim3((unsigned)((P + 255) / 256), N), 256, 0, st>>>(
        reinterpret_cast<const long long*>(p->labels), N, P, p->bg_label, p->seed, pos, neg, cnt, cnt + N);
    D2B_LAUNCH_CHECK();
    // the sampled prefix of each class: top-k of the random scores (-inf = not eligible; trimmed by the counts)
    pl.td.scores[0] = pos;
    rc = topk_run(pl.td, keys, nullptr, idx_pos, nullptr, ws + pl.o_topk, st);
    if (rc != D2B_OK) return rc;
    pl.td.scores[0] = neg;
    rc = topk_run(pl.td, keys, nullptr, idx_neg, nullptr, ws + pl.o_topk, st);
    if (rc != D2B_OK) return rc;
  }
  if (p->out_labels && P > 0) {
    sub_fill_kernel<<<(unsigned)(((long long)N * P + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<long long*>(p->out_labels), (long long)N * P);
    D2B_LAUNCH_CHECK();
  }
  const int want_pos = p->max_positives;  // int(num_samples * positive_fraction), :37
  sub_emit_kernel<<<dim3((k + 255) / 256, N), 256, 0, st>>>(
      reinterpret_cast<const long long*>(p->labels), P, k, want_pos, cnt, cnt + N, idx_pos, idx_neg,
      reinterpret_cast<long long*>(p->out_pos_idx), reinterpret_cast<long long*>(p->out_neg_idx), p->out_num_pos,
      p->out_num_neg, reinterpret_cast<long long*>(p->out_labels));
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}
