// Minimal stand-in for <torch/extension.h> (see oracle/stubs/README.md).  Test infrastructure only.
//
// The reference's interface loops (src/interface/{distance,rigid,graph,cad}_layer.cc, normalize.cc) use five things
// of libtorch: torch::Tensor::size(i), torch::Tensor::storage().data() (the raw buffer; against the libtorch of 2019
// that returned a mutable void*, today's returns const void* and the sources no longer compile against it),
// torch::TensorOptions().dtype(...), the dtype constants, and torch::full(shape, value, options) for fresh outputs.
// This header declares exactly those on a plain host buffer so that the sources compile UNMODIFIED from where they
// lie under /root/reference into oracle/_ref/libmeshode_ref.so.
#ifndef MESHODE_STUB_TORCH_EXTENSION_H_
#define MESHODE_STUB_TORCH_EXTENSION_H_
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <vector>

namespace torch {

enum ScalarType { kFloat32, kFloat64, kInt32 };
inline size_t element_size(ScalarType t) { return t == kFloat64 ? 8 : 4; }

struct TensorOptions {
  ScalarType type = kFloat32;
  TensorOptions dtype(ScalarType t) const { TensorOptions o; o.type = t; return o; }
};

struct Storage {
  void* ptr;
  void* data() const { return ptr; }
};

class Tensor {
 public:
  Tensor() : ptr_(nullptr), type_(kFloat32) {}
  // a view of caller-owned memory (what pybind11 hands the reference: the caller's tensor)
  static Tensor view(void* p, std::initializer_list<long long> shape, ScalarType t) {
    Tensor x; x.ptr_ = p; x.shape_.assign(shape.begin(), shape.end()); x.type_ = t; return x;
  }
  static Tensor owned(std::initializer_list<long long> shape, ScalarType t) {
    Tensor x; x.shape_.assign(shape.begin(), shape.end()); x.type_ = t;
    size_t n = element_size(t);
    for (long long d : x.shape_) n *= (size_t)d;
    x.own_ = std::make_shared<std::vector<unsigned char>>(n ? n : 1);
    x.ptr_ = x.own_->data();
    return x;
  }
  long long size(int i) const { return shape_[(size_t)i]; }
  Storage storage() const { return Storage{ptr_}; }
  long long numel() const { long long n = 1; for (long long d : shape_) n *= d; return n; }
  ScalarType scalar_type() const { return type_; }
 private:
  std::shared_ptr<std::vector<unsigned char>> own_;
  void* ptr_;
  std::vector<long long> shape_;
  ScalarType type_;
};

// torch::full(shape, value, options): a fresh contiguous tensor filled with `value`
inline Tensor full(std::initializer_list<long long> shape, double value, const TensorOptions& o) {
  Tensor x = Tensor::owned(shape, o.type);
  const long long n = x.numel();
  void* p = x.storage().data();
  for (long long i = 0; i < n; ++i) {
    if (o.type == kFloat32) static_cast<float*>(p)[i] = (float)value;
    else if (o.type == kFloat64) static_cast<double*>(p)[i] = value;
    else static_cast<int*>(p)[i] = (int)value;
  }
  return x;
}

}  // namespace torch
#endif
