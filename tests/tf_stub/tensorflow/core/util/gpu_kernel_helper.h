// Stub of tensorflow/core/util/gpu_kernel_helper.h: Eigen::GpuDevice::stream() and the CUDA runtime types.
#ifndef D2B_TF_STUB_GPU_KERNEL_HELPER_H_
#define D2B_TF_STUB_GPU_KERNEL_HELPER_H_

#if defined(__has_include)
#if __has_include(<cuda_runtime.h>)
#include <cuda_runtime.h>
#define D2B_TF_STUB_HAVE_CUDA 1
#endif
#endif
#ifndef D2B_TF_STUB_HAVE_CUDA
#include <cstddef>
typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
inline cudaError_t cudaMemsetAsync(void* ptr, int value, size_t count, cudaStream_t stream) {
  (void)ptr; (void)value; (void)count; (void)stream;
  return 0;
}
#endif

namespace Eigen {
struct GpuDevice {
  cudaStream_t stream() const { return nullptr; }
};
}  // namespace Eigen

#endif  // D2B_TF_STUB_GPU_KERNEL_HELPER_H_
