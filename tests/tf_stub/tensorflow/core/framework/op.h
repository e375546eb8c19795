// Stub of tensorflow/core/framework/op.h: REGISTER_OP("Name").Input(..).Output(..).Attr(..).SetShapeFn(fn).
#ifndef D2B_TF_STUB_OP_H_
#define D2B_TF_STUB_OP_H_

#include <functional>

#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

namespace tensorflow {

class OpDefBuilderWrapper {
 public:
  explicit OpDefBuilderWrapper(const char* name) { (void)name; }
  OpDefBuilderWrapper& Input(const char* spec) { (void)spec; return *this; }
  OpDefBuilderWrapper& Output(const char* spec) { (void)spec; return *this; }
  OpDefBuilderWrapper& Attr(const char* spec) { (void)spec; return *this; }
  OpDefBuilderWrapper& Doc(const char* text) { (void)text; return *this; }
  OpDefBuilderWrapper& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)> fn) { (void)fn; return *this; }
};

#define REGISTER_OP(name) \
  static ::tensorflow::OpDefBuilderWrapper D2B_TF_STUB_CONCAT(d2b_stub_op_registrar_, __COUNTER__) = ::tensorflow::OpDefBuilderWrapper(name)

}  // namespace tensorflow

#endif  // D2B_TF_STUB_OP_H_
