// Minimal stand-in for <ATen/ATen.h> -- TEST INFRASTRUCTURE ONLY.
//
// The reference's .cu files were written against PyTorch 1.0 (Tensor::type(),
// Tensor::data<T>(), at::zeros(sizes, Type)) and no longer compile against the
// torch 2.x headers in this image. Their use of ATen is tiny (size/dim/type/
// data/zeros/zeros_like/AT_DISPATCH_FLOATING_TYPES/TORCH_CHECK), so this header
// provides just that on top of the CUDA runtime. With it the reference kernels
// compile UNMODIFIED from /root/reference for sm_100a and serve as the
// "reference itself, run on the GPU box" pin for the oracle (oracle/Makefile).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace at {

enum class ScalarType { Float, Double };

struct Type {
  ScalarType st;
  bool cuda;
  bool is_cuda() const { return cuda; }
  bool operator==(const Type& o) const { return st == o.st && cuda == o.cuda; }
  bool operator!=(const Type& o) const { return !(*this == o); }
};

struct Tensor {
  void* ptr = nullptr;
  std::vector<int64_t> sizes;
  ScalarType st = ScalarType::Float;
  std::shared_ptr<void> owner;  // set when the tensor owns its allocation

  int64_t dim() const { return (int64_t)sizes.size(); }
  int64_t size(int64_t i) const { return sizes.at((size_t)i); }
  Type type() const { return Type{st, true}; }
  bool is_contiguous() const { return true; }
  bool is_cuda() const { return true; }
  int64_t numel() const { int64_t n = 1; for (auto s : sizes) n *= s; return n; }
  size_t itemsize() const { return st == ScalarType::Float ? 4 : 8; }
  template <typename T> T* data() const { return reinterpret_cast<T*>(ptr); }
};

inline Tensor zeros(std::vector<int64_t> sizes, Type t) {
  Tensor r;
  r.sizes = sizes;
  r.st = t.st;
  size_t bytes = (size_t)r.numel() * r.itemsize();
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
  cudaMemset(p, 0, bytes);
  r.ptr = p;
  r.owner = std::shared_ptr<void>(p, [](void* q) { cudaFree(q); });
  return r;
}
inline Tensor zeros_like(const Tensor& x) { return zeros(x.sizes, x.type()); }

}  // namespace at

#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...)            \
  [&] {                                                         \
    if ((TYPE).st == at::ScalarType::Float) {                   \
      using scalar_t = float;                                   \
      return __VA_ARGS__();                                     \
    } else {                                                    \
      using scalar_t = double;                                  \
      return __VA_ARGS__();                                     \
    }                                                           \
  }()

#define TORCH_CHECK(cond, ...) \
  if (!(cond)) { throw std::runtime_error(std::string("TORCH_CHECK failed: " #cond)); }
