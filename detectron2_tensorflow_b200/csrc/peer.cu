// peer.cu -- the final gather to rank 0 over NVLink peer memory: CUDA IPC arenas and ONE
// kernel that waits on flags, copies a table of segments (local -> peer stores on the senders, receive slots ->
// full-batch tensors on rank 0) and raises flags.  Replaces pack copies + an NCCL collective + unpack copies.
#include "common.cuh"

namespace d2b {
namespace {

struct PeerArgs {
  d2b_copy_segment seg[D2B_PEER_MAX_SEGMENTS];
  const unsigned long long* wait[D2B_PEER_MAX_FLAGS];
  unsigned long long* sig[D2B_PEER_MAX_FLAGS];
  unsigned long long* counter;
  unsigned int* ticket;
  int* err;
  unsigned long long timeout_ns;
  int nseg, nwait, nsig, lag;
};

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int kPeerThreads = 512;

__global__ void __launch_bounds__(kPeerThreads) peer_copy_kernel(const __grid_constant__ PeerArgs a) {
  const int tid = threadIdx.x;
  // every CTA reads the epoch before it takes its ticket; the counter only moves after the last ticket
  const unsigned long long epoch = ld_flag(a.counter) + 1ull;
  if (a.nwait > 0) {
    if (tid < a.nwait) {
      const unsigned long long want = epoch - (unsigned long long)a.lag;
      const unsigned long long t0 = now_ns();
      unsigned int spins = 0;
      // once a wait has timed out the ranks are out of step for good: later launches do not wait again (one time-out
      // per plan, not one per step), the host sees the error word
      const bool dead = *reinterpret_cast<volatile int*>(a.err) != 0;
      while (!dead && ld_flag(a.wait[tid]) < want) {
        if ((++spins & 63u) == 0u) {
          if (now_ns() - t0 > a.timeout_ns) {
            *a.err = 1;
            break;
          }
          __nanosleep(64);
        }
      }
      __threadfence_system();  // acquire: the peers' data stores precede their flag stores
    }
    __syncthreads();
  }
  const size_t stride = (size_t)gridDim.x * kPeerThreads;
  const size_t gtid = (size_t)blockIdx.x * kPeerThreads + tid;
  for (int s = 0; s < a.nseg; ++s) {
    const char* src = static_cast<const char*>(a.seg[s].src);
    char* dst = static_cast<char*>(a.seg[s].dst);
    const size_t bytes = a.seg[s].bytes;
    const size_t mis = (reinterpret_cast<size_t>(src) | reinterpret_cast<size_t>(dst));
    size_t done;
    if ((mis & 15) == 0) {  // L1-bypassing loads: a receive slot is written by another GPU between launches
      const size_t nv = bytes >> 4;
      for (size_t i = gtid; i < nv; i += stride)
        reinterpret_cast<uint4*>(dst)[i] = __ldcg(reinterpret_cast<const uint4*>(src) + i);
      done = nv << 4;
    } else if ((mis & 3) == 0) {
      const size_t nv = bytes >> 2;
      for (size_t i = gtid; i < nv; i += stride)
        reinterpret_cast<unsigned int*>(dst)[i] = __ldcg(reinterpret_cast<const unsigned int*>(src) + i);
      done = nv << 2;
    } else {
      done = 0;
    }
    for (size_t i = done + gtid; i < bytes; i += stride)
      dst[i] = static_cast<char>(__ldcg(reinterpret_cast<const unsigned char*>(src) + i));
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();  // this CTA's stores (ordered before by the barrier) are visible before the ticket
    const unsigned int t = atomicAdd(a.ticket, 1u);
    if (t == gridDim.x - 1) {
      *a.ticket = 0;
      __threadfence_system();
      for (int i = 0; i < a.nsig; ++i) st_flag(a.sig[i], epoch);
      st_flag(a.counter, epoch);
    }
  }
}

}  // namespace
}  // namespace d2b

using namespace d2b;

extern "C" int d2b_peer_alloc(size_t bytes, void** ptr) {
  D2B_REQUIRE(ptr != nullptr && bytes > 0, "peer_alloc: ptr is NULL or bytes == 0");
  D2B_CUDA(cudaMalloc(ptr, bytes));
  D2B_CUDA(cudaMemset(*ptr, 0, bytes));
  D2B_CUDA(cudaDeviceSynchronize());
  return D2B_OK;
}

extern "C" int d2b_peer_free(void* ptr) {
  if (ptr) D2B_CUDA(cudaFree(ptr));
  return D2B_OK;
}

static_assert(sizeof(cudaIpcMemHandle_t) == D2B_PEER_HANDLE_BYTES, "IPC handle size");

extern "C" int d2b_peer_export(void* ptr, unsigned char handle[D2B_PEER_HANDLE_BYTES]) {
  D2B_REQUIRE(ptr != nullptr && handle != nullptr, "peer_export: NULL argument");
  cudaIpcMemHandle_t h;
  D2B_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, sizeof(h));
  return D2B_OK;
}

extern "C" int d2b_peer_open(const unsigned char handle[D2B_PEER_HANDLE_BYTES], void** ptr) {
  D2B_REQUIRE(ptr != nullptr && handle != nullptr, "peer_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  D2B_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return D2B_OK;
}

extern "C" int d2b_peer_close(void* ptr) {
  if (ptr) D2B_CUDA(cudaIpcCloseMemHandle(ptr));
  return D2B_OK;
}

extern "C" size_t d2b_peer_copy_workspace_bytes(const d2b_peer_copy_params*) { return 0; }

extern "C" int d2b_peer_copy(const d2b_peer_copy_params* p, void*, size_t, d2b_stream_t stream) {
  D2B_REQUIRE(p != nullptr, "params is NULL");
  D2B_REQUIRE(p->num_segments >= 0 && p->num_segments <= D2B_PEER_MAX_SEGMENTS,
              "peer_copy: num_segments must be in [0, %d]", D2B_PEER_MAX_SEGMENTS);
  D2B_REQUIRE(p->num_wait >= 0 && p->num_wait <= D2B_PEER_MAX_FLAGS && p->num_signal >= 0 &&
                  p->num_signal <= D2B_PEER_MAX_FLAGS,
              "peer_copy: at most %d wait / signal flags", D2B_PEER_MAX_FLAGS);
  D2B_REQUIRE(p->wait_lag == 0 || p->wait_lag == 1, "peer_copy: wait_lag must be 0 or 1");
  D2B_REQUIRE(p->epoch_counter && p->ticket && p->error_flag, "peer_copy: epoch_counter / ticket / error_flag are NULL");
  D2B_REQUIRE(p->num_segments == 0 || p->segments, "peer_copy: segments is NULL");
  D2B_REQUIRE(p->num_wait == 0 || p->wait_flags, "peer_copy: wait_flags is NULL");
  D2B_REQUIRE(p->num_signal == 0 || p->signal_flags, "peer_copy: signal_flags is NULL");
  PeerArgs a;
  memset(&a, 0, sizeof(a));
  size_t total = 0;
  for (int i = 0; i < p->num_segments; ++i) {
    D2B_REQUIRE(p->segments[i].bytes == 0 || (p->segments[i].src && p->segments[i].dst), "peer_copy: segment %d is NULL", i);
    a.seg[i] = p->segments[i];
    total += p->segments[i].bytes;
  }
  for (int i = 0; i < p->num_wait; ++i) {
    D2B_REQUIRE(p->wait_flags[i], "peer_copy: wait flag %d is NULL", i);
    a.wait[i] = reinterpret_cast<const unsigned long long*>(p->wait_flags[i]);
  }
  for (int i = 0; i < p->num_signal; ++i) {
    D2B_REQUIRE(p->signal_flags[i], "peer_copy: signal flag %d is NULL", i);
    a.sig[i] = reinterpret_cast<unsigned long long*>(p->signal_flags[i]);
  }
  a.counter = reinterpret_cast<unsigned long long*>(p->epoch_counter);
  a.ticket = p->ticket;
  a.err = p->error_flag;
  a.timeout_ns = (unsigned long long)(p->timeout_ms ? p->timeout_ms : 2000u) * 1000000ull;
  a.nseg = p->num_segments;
  a.nwait = p->num_wait;
  a.nsig = p->num_signal;
  a.lag = p->wait_lag;
  // small transfers: latency, not bandwidth -- one CTA per 32 KB, at most 32
  int grid = (int)((total + 32767) / 32768);
  grid = grid < 1 ? 1 : (grid > 32 ? 32 : grid);
  peer_copy_kernel<<<grid, kPeerThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  D2B_LAUNCH_CHECK();
  return D2B_OK;
}
