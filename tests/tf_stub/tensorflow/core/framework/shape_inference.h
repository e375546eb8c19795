// Stub of tensorflow/core/framework/shape_inference.h (see op_kernel.h in this directory).
#ifndef D2B_TF_STUB_SHAPE_INFERENCE_H_
#define D2B_TF_STUB_SHAPE_INFERENCE_H_

#include <initializer_list>

#include "tensorflow/core/framework/op_kernel.h"

namespace tensorflow {
namespace shape_inference {

class DimensionHandle {};
class ShapeHandle {};

struct DimensionOrConstant {
  DimensionOrConstant(DimensionHandle d) : dim(d), val(-1) {}  // NOLINT: implicit, as in TF
  DimensionOrConstant(int64 v) : val(v) {}                     // NOLINT
  DimensionOrConstant(int v) : val(v) {}                       // NOLINT
  DimensionHandle dim;
  int64 val;
};

class InferenceContext {
 public:
  static constexpr int64 kUnknownDim = -1;
  ShapeHandle input(int idx) { (void)idx; return ShapeHandle(); }
  DimensionHandle Dim(ShapeHandle s, int64 idx) { (void)s; (void)idx; return DimensionHandle(); }
  void set_output(int idx, ShapeHandle shape) { (void)idx; (void)shape; }
  ShapeHandle MakeShape(std::initializer_list<DimensionOrConstant> dims) { (void)dims; return ShapeHandle(); }
  ShapeHandle Matrix(DimensionOrConstant dim1, DimensionOrConstant dim2) { (void)dim1; (void)dim2; return ShapeHandle(); }
  ShapeHandle Vector(DimensionOrConstant dim) { (void)dim; return ShapeHandle(); }
  template <typename T> Status GetAttr(const char* name, T* value) const { (void)name; (void)value; return Status::OK(); }
};

}  // namespace shape_inference
}  // namespace tensorflow

#endif  // D2B_TF_STUB_SHAPE_INFERENCE_H_
