// Minimal DECLARATIONS of the TensorFlow C++ custom-op API used by csrc/tf_ops/d2b_tf_ops.cc -- a stand-in for a
// syntax / type check only (tests/test_tf_shim_syntax.py: g++ -fsyntax-only).  TensorFlow is not installable in the
// build image; these headers restate, from the public TF 1.x API, exactly the names and signatures the shim relies
// on.  Nothing here is linked or run.
#ifndef D2B_TF_STUB_OP_KERNEL_H_
#define D2B_TF_STUB_OP_KERNEL_H_

#include <cstdint>
#include <initializer_list>
#include <sstream>
#include <string>
#include <vector>

namespace tensorflow {

typedef long long int64;
typedef int int32;
typedef unsigned char uint8;

class Status {
 public:
  Status() {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}

 private:
  bool ok_ = true;
  std::string msg_;
};

namespace errors {
namespace internal {
inline void Append(std::ostringstream&) {}
template <typename T, typename... Rest>
void Append(std::ostringstream& os, const T& v, const Rest&... rest) {
  os << v;
  Append(os, rest...);
}
template <typename... Args>
Status Make(const Args&... args) {
  std::ostringstream os;
  Append(os, args...);
  return Status(os.str());
}
}  // namespace internal
template <typename... Args> Status InvalidArgument(const Args&... args) { return internal::Make(args...); }
template <typename... Args> Status Internal(const Args&... args) { return internal::Make(args...); }
template <typename... Args> Status Unimplemented(const Args&... args) { return internal::Make(args...); }
}  // namespace errors

enum DataType { DT_FLOAT = 1, DT_UINT8 = 4, DT_INT32 = 3, DT_INT64 = 9, DT_BOOL = 10 };

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> dims) : dims_(dims) {}
  int dims() const { return static_cast<int>(dims_.size()); }
  int64 dim_size(int i) const { return dims_[i]; }

 private:
  std::vector<int64> dims_;
};

template <typename T>
class FlatView {  // what Tensor::flat<T>() returns: .data() and operator()(i)
 public:
  T* data() const { return ptr_; }
  T& operator()(int64 i) const { return ptr_[i]; }
  T* ptr_ = nullptr;
};

class Tensor {
 public:
  int dims() const { return shape_.dims(); }
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  int64 NumElements() const { return 0; }
  const TensorShape& shape() const { return shape_; }
  template <typename T> FlatView<T> flat() { return FlatView<T>(); }
  template <typename T> FlatView<const T> flat() const { return FlatView<const T>(); }

 private:
  TensorShape shape_;
};

class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char* name, T* value) const { (void)name; (void)value; return Status::OK(); }
  void CtxFailure(const char* file, int line, const Status& s) { (void)file; (void)line; (void)s; }
  void CtxFailureWithWarning(const char* file, int line, const Status& s) { (void)file; (void)line; (void)s; }
};

class OpKernelContext {
 public:
  const Tensor& input(int index) { (void)index; return dummy_; }
  Status allocate_output(int index, const TensorShape& shape, Tensor** out) { (void)index; (void)shape; *out = &dummy_; return Status::OK(); }
  Status allocate_temp(DataType type, const TensorShape& shape, Tensor* out) { (void)type; (void)shape; (void)out; return Status::OK(); }
  template <typename Device> const Device& eigen_device() const { static Device d; return d; }
  void CtxFailure(const char* file, int line, const Status& s) { (void)file; (void)line; (void)s; }
  void CtxFailureWithWarning(const char* file, int line, const Status& s) { (void)file; (void)line; (void)s; }

 private:
  Tensor dummy_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction* context) { (void)context; }
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS)                         \
  do {                                                        \
    if (!(EXP)) {                                             \
      (CTX)->CtxFailure(__FILE__, __LINE__, (STATUS));        \
      return;                                                 \
    }                                                         \
  } while (0)

#define OP_REQUIRES_OK(CTX, ...)                                         \
  do {                                                                   \
    ::tensorflow::Status _s(__VA_ARGS__);                                \
    if (!_s.ok()) {                                                      \
      (CTX)->CtxFailureWithWarning(__FILE__, __LINE__, _s);              \
      return;                                                            \
    }                                                                    \
  } while (0)

#define TF_RETURN_IF_ERROR(...)                         \
  do {                                                  \
    const ::tensorflow::Status _status = (__VA_ARGS__); \
    if (!_status.ok()) return _status;                  \
  } while (0)

extern const char* const DEVICE_CPU;
extern const char* const DEVICE_GPU;

// REGISTER_KERNEL_BUILDER(Name("Op").Device(DEVICE_GPU).HostMemory("x"), KernelClass)
class KernelDefBuilder {
 public:
  explicit KernelDefBuilder(const char* op_name) { (void)op_name; }
  KernelDefBuilder& Device(const char* device_type) { (void)device_type; return *this; }
  KernelDefBuilder& HostMemory(const char* arg_name) { (void)arg_name; return *this; }
  template <typename T> KernelDefBuilder& TypeConstraint(const char* attr_name) { (void)attr_name; return *this; }
};
namespace register_kernel {
typedef KernelDefBuilder Name;
struct Registrar {
  Registrar(const KernelDefBuilder& def, OpKernel* (*factory)(OpKernelConstruction*)) { (void)def; (void)factory; }
};
}  // namespace register_kernel

#define D2B_TF_STUB_CONCAT_(a, b) a##b
#define D2B_TF_STUB_CONCAT(a, b) D2B_TF_STUB_CONCAT_(a, b)
#define REGISTER_KERNEL_BUILDER(kernel_builder, ...)                                                     \
  static ::tensorflow::register_kernel::Registrar D2B_TF_STUB_CONCAT(d2b_stub_kernel_registrar_, __COUNTER__)( \
      ::tensorflow::register_kernel::kernel_builder,                                                     \
      [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })

}  // namespace tensorflow

#endif  // D2B_TF_STUB_OP_KERNEL_H_
