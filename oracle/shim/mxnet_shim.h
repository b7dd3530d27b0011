// TEST INFRASTRUCTURE (oracle/): a minimal stand-in for the MXNet / mshadow / nnvm / dmlc headers, just large
// enough that the reference's own operator sources under /root/reference/operator_cxx compile UNMODIFIED,
// where they lie, into oracle/_ref/libref_cxx.so (recipe: oracle/build_ref_cxx.py).  Nothing here restates
// any operator arithmetic: tensors are plain views over host memory, Kernel<OP,xpu>::Launch is a serial
// loop over OP::Map (what mxnet_op::Kernel<OP,cpu> does, minus OpenMP), atomicAdd is `+=` in index order,
// registration macros record the registered functors so the harness can call them.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeindex>
#include <unordered_map>
#include <utility>
#include <vector>

#define DMLC_USE_CXX11 1
#define MSHADOW_XINLINE inline
#define MSHADOW_CINLINE inline
#define MSHADOW_FORCE_INLINE inline
#define MSHADOW_USE_CUDA 0
// CUDA source compiled as plain C++: a __global__ function is an ordinary function run by ONE thread of a 1x1 grid,
// so a grid-stride loop visits every index in order
#ifndef __CUDACC__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
// ---- serial emulation of a kernel launch.  The build recipe rewrites `kernel<<<grid, block, ...>>>(args)` into
// `shim_launch(grid, block, ...).run([&](auto... a) { kernel(a...); }, args)`: every thread of every block runs the
// kernel body in turn with blockIdx / threadIdx set.  __shared__ variables are function statics; a block whose
// threads reached __syncthreads() is simply run a second time - exact for the load -> barrier -> compute kernels
// in these files (the second pass sees the complete shared data and overwrites what the first one wrote).
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}  // NOLINT
};
static dim3 blockIdx(0, 0, 0), threadIdx(0, 0, 0), blockDim(1, 1, 1), gridDim(1, 1, 1);
static bool shim_saw_sync = false;
#define __shared__ static
inline void __syncthreads() { shim_saw_sync = true; }
struct shim_launch {
  dim3 grid, block;
  shim_launch() {}
  template <typename S = void*>
  shim_launch(dim3 g, dim3 b, size_t = 0, S = S()) : grid(g), block(b) {}
  template <typename F, typename... Args>
  void run(F f, Args... args) const {
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          blockIdx = dim3(bx, by, bz);
          for (int pass = 0; pass < 2; ++pass) {
            shim_saw_sync = false;
            for (unsigned tz = 0; tz < block.z; ++tz)
              for (unsigned ty = 0; ty < block.y; ++ty)
                for (unsigned tx = 0; tx < block.x; ++tx) {
                  threadIdx = dim3(tx, ty, tz);
                  f(args...);
                }
            if (!shim_saw_sync) break;
          }
        }
    blockIdx = threadIdx = dim3(0, 0, 0);
    blockDim = gridDim = dim3(1, 1, 1);
  }
};
// the slice of the CUDA runtime these files call: one address space, nothing can fail
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { std::memmove(dst, src, n); return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "no error"; }
// CUDA's global min / max overload set (math_functions.hpp): integers by comparison, floating point through fmin /
// fmax, mixed float/double promoted to double
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline double min(double a, double b) { return ::fmin(a, b); }
inline double max(double a, double b) { return ::fmax(a, b); }
inline double min(float a, double b) { return ::fmin((double)a, b); }
inline double max(float a, double b) { return ::fmax((double)a, b); }
inline double min(double a, float b) { return ::fmin(a, (double)b); }
inline double max(double a, float b) { return ::fmax(a, (double)b); }
#ifdef SHIM_CUDA_DEVICE_MATH
// In device code the unqualified math functions have float overloads (exp(float) IS expf there, while g++ binds a
// host-side `exp(float_value)` to ::exp(double) - which is what decodebbox.cc really does on the CPU).  Only the .cu
// translation units get these.
inline float exp(float x) { return ::expf(x); }
inline float log(float x) { return ::logf(x); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float floor(float x) { return ::floorf(x); }
inline float ceil(float x) { return ::ceilf(x); }
inline float round(float x) { return ::roundf(x); }
inline float fabs(float x) { return ::fabsf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
#endif
#endif
#define MXNET_USE_CUDA 0

// ------------------------------------------------------------------------------------------ dmlc
namespace dmlc {
struct Error : public std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};
struct ParamError : public Error {
  explicit ParamError(const std::string& s) : Error(s) {}
};
class LogMessageFatal {
 public:
  LogMessageFatal(const char* file, int line) { s_ << file << ":" << line << ": "; }
  std::ostringstream& stream() { return s_; }
  ~LogMessageFatal() noexcept(false) { throw Error(s_.str()); }

 private:
  std::ostringstream s_;
};
class LogMessageNull {
 public:
  template <typename T>
  LogMessageNull& operator<<(const T&) { return *this; }
};
}  // namespace dmlc
#define SHIM_CHECK_BINARY(x, y, op) \
  if (!((x)op(y))) ::dmlc::LogMessageFatal(__FILE__, __LINE__).stream() << "Check failed: " #x " " #op " " #y << ": "
#define CHECK(x) \
  if (!(x)) ::dmlc::LogMessageFatal(__FILE__, __LINE__).stream() << "Check failed: " #x << ": "
#define CHECK_EQ(x, y) SHIM_CHECK_BINARY(x, y, ==)
#define CHECK_NE(x, y) SHIM_CHECK_BINARY(x, y, !=)
#define CHECK_LT(x, y) SHIM_CHECK_BINARY(x, y, <)
#define CHECK_GT(x, y) SHIM_CHECK_BINARY(x, y, >)
#define CHECK_LE(x, y) SHIM_CHECK_BINARY(x, y, <=)
#define CHECK_GE(x, y) SHIM_CHECK_BINARY(x, y, >=)
#define LOG_FATAL ::dmlc::LogMessageFatal(__FILE__, __LINE__).stream()
#define LOG_INFO ::dmlc::LogMessageNull()
#define LOG_WARNING ::dmlc::LogMessageNull()
#define LOG(severity) LOG_##severity

// ------------------------------------------------------------------------------------------ mshadow
namespace mshadow {
typedef int32_t index_t;  // mshadow/base.h with MSHADOW_INT64_TENSOR_SIZE == 0 (the 1.6.0 build of docker/Dockerfile)
typedef float real_t;
typedef float default_real_t;
struct cpu {
  static const bool kDevCPU = true;
  static const int kDevMask = 1 << 0;
};
struct gpu {
  static const bool kDevCPU = false;
  static const int kDevMask = 1 << 1;
};
enum TypeFlag { kFloat32 = 0, kFloat64 = 1, kFloat16 = 2, kUint8 = 3, kInt32 = 4, kInt8 = 5, kInt64 = 6 };
template <typename DType> struct DataType;
template <> struct DataType<float> { static const int kFlag = kFloat32; };
template <> struct DataType<double> { static const int kFlag = kFloat64; };
template <> struct DataType<int32_t> { static const int kFlag = kInt32; };
template <> struct DataType<uint8_t> { static const int kFlag = kUint8; };
namespace expr {}
namespace red {
namespace limits {
template <typename DType> MSHADOW_XINLINE DType MinValue();
template <> MSHADOW_XINLINE float MinValue<float>() { return -FLT_MAX; }
template <> MSHADOW_XINLINE double MinValue<double>() { return -DBL_MAX; }
}  // namespace limits
}  // namespace red

}  // namespace mshadow
struct CUstream_st;
typedef CUstream_st* cudaStream_t;  // only the handle type: the shim never touches the CUDA runtime
namespace mshadow {
namespace cuda {
const int kMaxThreadsPerBlock = 1024;
const int kMaxGridDim = 65535;
const int kBaseThreadNum = 256;
inline void CheckLaunchParam(dim3, dim3, const char* = "") {}
}  // namespace cuda
template <typename Device> struct Stream {
  static cudaStream_t GetStream(Stream<Device>*) { return nullptr; }
};

template <int dimension>
struct Shape {
  static const int kDimension = dimension;
  index_t shape_[dimension];
  index_t& operator[](int i) { return shape_[i]; }
  const index_t& operator[](int i) const { return shape_[i]; }
  bool operator==(const Shape& o) const {
    for (int i = 0; i < dimension; ++i)
      if (shape_[i] != o.shape_[i]) return false;
    return true;
  }
  bool operator!=(const Shape& o) const { return !(*this == o); }
  size_t Size() const {
    size_t s = 1;
    for (int i = 0; i < dimension; ++i) s *= (size_t)shape_[i];
    return s;
  }
  index_t ProdShape(int b, int e) const {
    index_t s = 1;
    for (int i = b; i < e; ++i) s *= shape_[i];
    return s;
  }
};
inline Shape<1> Shape1(index_t a) { Shape<1> s; s[0] = a; return s; }
inline Shape<2> Shape2(index_t a, index_t b) { Shape<2> s; s[0] = a; s[1] = b; return s; }
inline Shape<3> Shape3(index_t a, index_t b, index_t c) { Shape<3> s; s[0] = a; s[1] = b; s[2] = c; return s; }
inline Shape<4> Shape4(index_t a, index_t b, index_t c, index_t d) {
  Shape<4> s; s[0] = a; s[1] = b; s[2] = c; s[3] = d; return s;
}
inline Shape<5> Shape5(index_t a, index_t b, index_t c, index_t d, index_t e) {
  Shape<5> s; s[0] = a; s[1] = b; s[2] = c; s[3] = d; s[4] = e; return s;
}

// dense host view; stride_ == shape_[dim-1] always
// ---- a small mshadow::expr: lazily evaluated element-wise expressions over contiguous tensors.  An expression
// answers Eval(i) for the flat index i of the DESTINATION and shape() (empty for scalars).  Operator precedence and
// F<OP> nesting build the same tree mshadow builds, and every node rounds like mshadow's CPU Plan does (one DType
// operation per node), so `dst = expr` reproduces the reference's CPU arithmetic operation for operation.
namespace op {
struct plus { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return a + b; } };
struct minus { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return a - b; } };
struct mul { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return a * b; } };
struct div { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return a / b; } };
struct identity { template <typename DType> MSHADOW_XINLINE static DType Map(DType a) { return a; } };
}  // namespace op
namespace red {
struct sum {
  template <typename DType> MSHADOW_XINLINE static void Reduce(volatile DType& dst, volatile DType src) { dst += src; }  // NOLINT
  template <typename DType> MSHADOW_XINLINE static void SetInitValue(DType& v) { v = 0; }  // NOLINT
};
}  // namespace red
namespace expr {
typedef std::vector<index_t> ShapeVec;
inline size_t ShapeSize(const ShapeVec& v) { size_t n = 1; for (auto d : v) n *= (size_t)d; return n; }
template <typename SubType, typename DType>
struct Exp {
  const SubType& self() const { return *static_cast<const SubType*>(this); }
};
template <typename DType>
struct ScalarExp : public Exp<ScalarExp<DType>, DType> {
  DType scalar_;
  ScalarExp(DType s) : scalar_(s) {}  // NOLINT
  DType Eval(size_t) const { return scalar_; }
  ShapeVec shape() const { return {}; }
};
template <typename OP, typename TA, typename TB, typename DType>
struct BinaryMapExp : public Exp<BinaryMapExp<OP, TA, TB, DType>, DType> {
  TA lhs_;
  TB rhs_;
  BinaryMapExp(const TA& l, const TB& r) : lhs_(l), rhs_(r) {}
  DType Eval(size_t i) const { return OP::Map(lhs_.Eval(i), rhs_.Eval(i)); }
  ShapeVec shape() const { ShapeVec a = lhs_.shape(); return a.empty() ? rhs_.shape() : a; }
};
template <typename OP, typename TA, typename DType>
struct UnaryMapExp : public Exp<UnaryMapExp<OP, TA, DType>, DType> {
  TA src_;
  explicit UnaryMapExp(const TA& s) : src_(s) {}
  DType Eval(size_t i) const { return OP::Map(src_.Eval(i)); }
  ShapeVec shape() const { return src_.shape(); }
};
template <typename OP, typename TA, typename DType>
inline UnaryMapExp<OP, TA, DType> F(const Exp<TA, DType>& a) { return UnaryMapExp<OP, TA, DType>(a.self()); }
template <typename OP, typename TA, typename TB, typename DType>
inline BinaryMapExp<OP, TA, TB, DType> F(const Exp<TA, DType>& a, const Exp<TB, DType>& b) {
  return BinaryMapExp<OP, TA, TB, DType>(a.self(), b.self());
}
#define SHIM_EXPR_BINARY(sym, opname)                                                                      \
  template <typename TA, typename TB, typename DType>                                                      \
  inline BinaryMapExp<op::opname, TA, TB, DType> operator sym(const Exp<TA, DType>& a, const Exp<TB, DType>& b) { \
    return BinaryMapExp<op::opname, TA, TB, DType>(a.self(), b.self());                                    \
  }
SHIM_EXPR_BINARY(+, plus)
SHIM_EXPR_BINARY(-, minus)
SHIM_EXPR_BINARY(*, mul)
SHIM_EXPR_BINARY(/, div)
#undef SHIM_EXPR_BINARY
// broadcast_with_axis(src, axis, size): a new axis of `size` AFTER `axis` (extension/broadcast_with_axis.h)
template <typename TA, typename DType>
struct BroadcastWithAxisExp : public Exp<BroadcastWithAxisExp<TA, DType>, DType> {
  TA src_;
  size_t trailing_, size_;
  ShapeVec shape_;
  BroadcastWithAxisExp(const TA& s, int axis, int size) : src_(s), size_((size_t)size) {
    const ShapeVec ss = s.shape();
    trailing_ = 1;
    for (size_t d = (size_t)axis + 1; d < ss.size(); ++d) trailing_ *= (size_t)ss[d];
    shape_ = ss;
    shape_.insert(shape_.begin() + axis + 1, (index_t)size);
  }
  DType Eval(size_t i) const {
    const size_t inner = i % trailing_, outer = i / (trailing_ * size_);
    return src_.Eval(outer * trailing_ + inner);
  }
  ShapeVec shape() const { return shape_; }
};
template <typename TA, typename DType>
inline BroadcastWithAxisExp<TA, DType> broadcast_with_axis(const Exp<TA, DType>& src, int axis, int size) {
  return BroadcastWithAxisExp<TA, DType>(src.self(), axis, size);
}
// broadcast_scalar(1-element source, shape) and broadcast_keepdim(1-element source, 0, size): every element is src[0]
template <typename TA, typename DType>
struct BroadcastScalarExp : public Exp<BroadcastScalarExp<TA, DType>, DType> {
  TA src_;
  ShapeVec shape_;
  BroadcastScalarExp(const TA& s, const ShapeVec& shp) : src_(s), shape_(shp) {}
  DType Eval(size_t) const { return src_.Eval(0); }
  ShapeVec shape() const { return shape_; }
};
template <typename TA, typename DType, int dim>
inline BroadcastScalarExp<TA, DType> broadcast_scalar(const Exp<TA, DType>& src, Shape<dim> shape) {
  ShapeVec v;
  for (int d = 0; d < dim; ++d) v.push_back(shape[d]);
  return BroadcastScalarExp<TA, DType>(src.self(), v);
}
template <typename TA, typename DType>
inline BroadcastScalarExp<TA, DType> broadcast_keepdim(const Exp<TA, DType>& src, int axis, int size) {
  CHECK(ShapeSize(src.self().shape()) == 1 && axis == 0) << "shim: broadcast_keepdim only of a 1-element vector along axis 0";
  return BroadcastScalarExp<TA, DType>(src.self(), ShapeVec{(index_t)size});
}
// broadcast<dimcast>(1-D source, shape): dst[..., i_dimcast, ...] = src[i_dimcast]
template <typename TA, typename DType>
struct Broadcast1DExp : public Exp<Broadcast1DExp<TA, DType>, DType> {
  TA src_;
  size_t trailing_, extent_;
  ShapeVec shape_;
  Broadcast1DExp(const TA& s, int dimcast, const ShapeVec& shp) : src_(s), shape_(shp) {
    trailing_ = 1;
    for (size_t d = (size_t)dimcast + 1; d < shp.size(); ++d) trailing_ *= (size_t)shp[d];
    extent_ = (size_t)shp[dimcast];
  }
  DType Eval(size_t i) const { return src_.Eval((i / trailing_) % extent_); }
  ShapeVec shape() const { return shape_; }
};
template <int dimcast, typename TA, typename DType, int dimdst>
inline Broadcast1DExp<TA, DType> broadcast(const Exp<TA, DType>& src, Shape<dimdst> shape) {
  ShapeVec v;
  for (int d = 0; d < dimdst; ++d) v.push_back(shape[d]);
  return Broadcast1DExp<TA, DType>(src.self(), dimcast, v);
}
// reduce_keepdim<Reducer, false>(src, axis): reduce along `axis`, keep it with extent 1 (extension/reduce_with_axis.h)
template <typename Reducer, typename TA, typename DType>
struct ReduceKeepdimExp : public Exp<ReduceKeepdimExp<Reducer, TA, DType>, DType> {
  TA src_;
  size_t trailing_, extent_;
  ShapeVec shape_;
  ReduceKeepdimExp(const TA& s, int axis) : src_(s) {
    const ShapeVec ss = s.shape();
    trailing_ = 1;
    for (size_t d = (size_t)axis + 1; d < ss.size(); ++d) trailing_ *= (size_t)ss[d];
    extent_ = (size_t)ss[axis];
    shape_ = ss;
    shape_[axis] = 1;
  }
  DType Eval(size_t i) const {
    const size_t inner = i % trailing_, outer = i / trailing_;
    DType res;
    Reducer::SetInitValue(res);
    for (size_t k = 0; k < extent_; ++k) Reducer::Reduce(res, src_.Eval((outer * extent_ + k) * trailing_ + inner));
    return res;
  }
  ShapeVec shape() const { return shape_; }
};
template <typename Reducer, bool mask, typename TA, typename DType>
inline ReduceKeepdimExp<Reducer, TA, DType> reduce_keepdim(const Exp<TA, DType>& src, int axis) {
  return ReduceKeepdimExp<Reducer, TA, DType>(src.self(), axis);
}
// sumall_except_dim<dimkeep>(src): 1-D of extent shape[dimkeep]; CPU order of MapReduceKeepHighDim: for the kept
// index c, sum over the leading dims, then the trailing ones, innermost last
template <typename TA, typename DType>
struct SumAllExceptDimExp : public Exp<SumAllExceptDimExp<TA, DType>, DType> {
  TA src_;
  size_t leading_, extent_, trailing_;
  SumAllExceptDimExp(const TA& s, int dimkeep) : src_(s) {
    const ShapeVec ss = s.shape();
    leading_ = trailing_ = 1;
    for (int d = 0; d < dimkeep; ++d) leading_ *= (size_t)ss[d];
    for (size_t d = (size_t)dimkeep + 1; d < ss.size(); ++d) trailing_ *= (size_t)ss[d];
    extent_ = (size_t)ss[dimkeep];
  }
  DType Eval(size_t c) const {
    DType res = 0;
    for (size_t n = 0; n < leading_; ++n)
      for (size_t t = 0; t < trailing_; ++t) res += src_.Eval((n * extent_ + c) * trailing_ + t);
    return res;
  }
  ShapeVec shape() const { return ShapeVec{(index_t)extent_}; }
};
template <int dimkeep, typename TA, typename DType>
inline SumAllExceptDimExp<TA, DType> sumall_except_dim(const Exp<TA, DType>& src) {
  return SumAllExceptDimExp<TA, DType>(src.self(), dimkeep);
}
}  // namespace expr

template <typename Device, int dimension, typename DType = float>
struct Tensor : public expr::Exp<Tensor<Device, dimension, DType>, DType> {
  DType* dptr_ = nullptr;
  Shape<dimension> shape_;
  DType Eval(size_t i) const { return dptr_[i]; }
  expr::ShapeVec shape() const { expr::ShapeVec v; for (int d = 0; d < dimension; ++d) v.push_back(shape_[d]); return v; }
  // dst = expr: the destination is written in flat index order; the expression may read the destination itself
  // (grad = where(..., grad)) only at the same index, as in mshadow
  template <typename E>
  Tensor& operator=(const expr::Exp<E, DType>& e) {
    const E& x = e.self();
    const size_t n = MSize();
    for (size_t i = 0; i < n; ++i) dptr_[i] = x.Eval(i);
    return *this;
  }
  template <typename E>
  Tensor& operator*=(const expr::Exp<E, DType>& e) { const E& x = e.self(); for (size_t i = 0; i < MSize(); ++i) dptr_[i] *= x.Eval(i); return *this; }
  template <typename E>
  Tensor& operator/=(const expr::Exp<E, DType>& e) { const E& x = e.self(); for (size_t i = 0; i < MSize(); ++i) dptr_[i] /= x.Eval(i); return *this; }
  template <typename E>
  Tensor& operator+=(const expr::Exp<E, DType>& e) { const E& x = e.self(); for (size_t i = 0; i < MSize(); ++i) dptr_[i] += x.Eval(i); return *this; }
  index_t stride_ = 0;
  Stream<Device>* stream_ = nullptr;
  Tensor() { for (int i = 0; i < dimension; ++i) shape_[i] = 0; }
  Tensor(DType* p, const Shape<dimension>& s) : dptr_(p), shape_(s), stride_(s[dimension - 1]) {}
  Tensor(DType* p, const Shape<dimension>& s, Stream<Device>* st) : dptr_(p), shape_(s), stride_(s[dimension - 1]), stream_(st) {}
  Tensor(DType* p, const Shape<dimension>& s, index_t stride, Stream<Device>* st)
      : dptr_(p), shape_(s), stride_(stride), stream_(st) {}
  index_t size(int i) const { return shape_[i]; }
  size_t MSize() const { return shape_.Size(); }
  bool CheckContiguous() const { return true; }
  Tensor<Device, dimension - 1, DType> operator[](index_t idx) const {
    Shape<dimension - 1> s;
    size_t inner = 1;
    for (int i = 1; i < dimension; ++i) { s[i - 1] = shape_[i]; inner *= (size_t)shape_[i]; }
    return Tensor<Device, dimension - 1, DType>(dptr_ + (size_t)idx * inner, s);
  }
  Tensor Slice(index_t begin, index_t end) const {
    Shape<dimension> s = shape_;
    s[0] = end - begin;
    size_t inner = 1;
    for (int i = 1; i < dimension; ++i) inner *= (size_t)shape_[i];
    return Tensor(dptr_ + (size_t)begin * inner, s);
  }
  Tensor& operator=(const Tensor&) = default;
  Tensor& operator=(DType v) { std::fill(dptr_, dptr_ + MSize(), v); return *this; }
  Tensor& operator*=(DType v) { for (size_t i = 0; i < MSize(); ++i) dptr_[i] *= v; return *this; }
  Tensor& operator+=(DType v) { for (size_t i = 0; i < MSize(); ++i) dptr_[i] += v; return *this; }
};
template <typename Device, typename DType>
struct Tensor<Device, 1, DType> : public expr::Exp<Tensor<Device, 1, DType>, DType> {
  DType* dptr_ = nullptr;
  Shape<1> shape_;
  DType Eval(size_t i) const { return dptr_[i]; }
  expr::ShapeVec shape() const { return expr::ShapeVec{shape_[0]}; }
  template <typename E>
  Tensor& operator=(const expr::Exp<E, DType>& e) {
    const E& x = e.self();
    // (temp = f(temp) reads its own element before writing it: evaluate first)
    std::vector<DType> tmp(MSize());
    for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = x.Eval(i);
    std::copy(tmp.begin(), tmp.end(), dptr_);
    return *this;
  }
  index_t stride_ = 0;
  Stream<Device>* stream_ = nullptr;
  Tensor() { shape_[0] = 0; }
  Tensor(DType* p, const Shape<1>& s) : dptr_(p), shape_(s), stride_(s[0]) {}
  Tensor(DType* p, const Shape<1>& s, Stream<Device>* st) : dptr_(p), shape_(s), stride_(s[0]), stream_(st) {}
  Tensor(DType* p, const Shape<1>& s, index_t stride, Stream<Device>* st) : dptr_(p), shape_(s), stride_(stride), stream_(st) {}
  index_t size(int) const { return shape_[0]; }
  size_t MSize() const { return (size_t)shape_[0]; }
  bool CheckContiguous() const { return true; }
  DType& operator[](index_t i) const { return dptr_[i]; }
  Tensor Slice(index_t begin, index_t end) const { return Tensor(dptr_ + begin, Shape1(end - begin)); }
  Tensor& operator=(const Tensor&) = default;
  Tensor& operator=(DType v) { std::fill(dptr_, dptr_ + MSize(), v); return *this; }
  // the only expression-template uses in the files we build: elementwise -= and /= with an equal-shape vector
  Tensor& operator*=(DType v) { for (size_t i = 0; i < MSize(); ++i) dptr_[i] *= v; return *this; }
  Tensor& operator+=(DType v) { for (size_t i = 0; i < MSize(); ++i) dptr_[i] += v; return *this; }
  Tensor& operator-=(const Tensor& o) { for (index_t i = 0; i < shape_[0]; ++i) dptr_[i] -= o.dptr_[i]; return *this; }
  Tensor& operator/=(const Tensor& o) { for (index_t i = 0; i < shape_[0]; ++i) dptr_[i] /= o.dptr_[i]; return *this; }
};

template <typename Device, int dimension, typename DType = float>
class TensorContainer : public Tensor<Device, dimension, DType> {
 public:
  typedef Tensor<Device, dimension, DType> Base;
  TensorContainer() {}
  explicit TensorContainer(const Shape<dimension>& s) { Resize(s); }
  TensorContainer(const Shape<dimension>& s, DType v) { Resize(s); std::fill(store_.begin(), store_.end(), v); }
  TensorContainer(const TensorContainer& o) : Base(o), store_(o.store_) { this->dptr_ = store_.data(); }
  TensorContainer& operator=(const TensorContainer& o) {
    Base::operator=(o); store_ = o.store_; this->dptr_ = store_.data(); return *this;
  }
  TensorContainer& operator=(DType v) { std::fill(store_.begin(), store_.end(), v); return *this; }
  void Resize(const Shape<dimension>& s) {
    store_.resize(s.Size());
    this->shape_ = s;
    this->stride_ = s[dimension - 1];
    this->dptr_ = store_.data();
  }
  void Resize(const Shape<dimension>& s, DType v) { Resize(s); std::fill(store_.begin(), store_.end(), v); }

 private:
  std::vector<DType> store_;
};

template <typename DevA, typename DevB, int dim, typename DType>
inline void Copy(Tensor<DevA, dim, DType> dst, const Tensor<DevB, dim, DType>& src, Stream<gpu>* = nullptr) {
  CHECK(dst.shape_ == src.shape_) << "Copy: shape mismatch";
  if (dst.MSize()) std::memmove(dst.dptr_, src.dptr_, sizeof(DType) * dst.MSize());
}
template <typename DevA, typename DevB, int dim, typename DType>
inline void Copy(Tensor<DevA, dim, DType> dst, const Tensor<DevB, dim, DType>& src, Stream<cpu>*) {
  Copy(dst, src);
}
}  // namespace mshadow

#define MSHADOW_REAL_TYPE_SWITCH(type, DType, ...)                       \
  switch (type) {                                                        \
    case mshadow::kFloat32: { typedef float DType; {__VA_ARGS__} } break;  \
    case mshadow::kFloat64: { typedef double DType; {__VA_ARGS__} } break; \
    default: LOG(FATAL) << "shim: unsupported type flag " << type;       \
  }

// ------------------------------------------------------------------------------------------ nnvm / mxnet
namespace shim {
class any {  // a very small dmlc::any
 public:
  any() {}
  template <typename T>
  any& operator=(const T& v) { p_ = std::make_shared<T>(v); t_ = std::type_index(typeid(T)); return *this; }
  template <typename T>
  const T& get() const {
    if (!p_ || t_ != std::type_index(typeid(T))) throw dmlc::Error("shim::any: bad type");
    return *static_cast<const T*>(p_.get());
  }
 private:
  std::shared_ptr<void> p_;
  std::type_index t_ = std::type_index(typeid(void));
};
}  // namespace shim

namespace nnvm {
template <typename ValueType>
class Tuple {
 public:
  Tuple() {}
  template <typename It>
  Tuple(It b, It e) : v_(b, e) {}
  Tuple(std::initializer_list<ValueType> l) : v_(l) {}
  uint32_t ndim() const { return (uint32_t)v_.size(); }
  ValueType& operator[](size_t i) { return v_[i]; }
  const ValueType& operator[](size_t i) const { return v_[i]; }
  const ValueType* begin() const { return v_.data(); }
  const ValueType* end() const { return v_.data() + v_.size(); }
  void assign(const std::vector<ValueType>& v) { v_ = v; }
  bool operator==(const Tuple& o) const { return v_ == o.v_; }
 protected:
  std::vector<ValueType> v_;
};

struct NodeAttrs {
  std::string name;
  std::unordered_map<std::string, std::string> dict;
  shim::any parsed;
};
template <typename T>
inline const T& get(const shim::any& a) { return a.get<T>(); }
struct Node;
typedef std::shared_ptr<Node> NodePtr;
struct NodeEntry {
  NodePtr node;
  uint32_t index;
  uint32_t version;
};
struct Node {
  NodeAttrs attrs;
  std::vector<NodeEntry> inputs;
};
typedef std::function<std::vector<std::string>(const NodeAttrs&)> FListInputNames;
typedef std::function<std::vector<std::string>(const NodeAttrs&)> FListOutputNames;
typedef std::function<uint32_t(const NodeAttrs&)> FNumVisibleOutputs;
typedef std::function<bool(const NodeAttrs&, std::vector<int>*, std::vector<int>*)> FInferType;
typedef std::function<std::vector<NodeEntry>(const NodePtr&, const std::vector<NodeEntry>&)> FGradient;
typedef std::function<std::vector<std::pair<int, int>>(const NodeAttrs&)> FInplaceOption;
typedef bool TIsBackward;

class Op {  // records what NNVM_REGISTER_OP(...) chains register, by attribute name
 public:
  std::string name;
  int num_inputs = 1, num_outputs = 1;
  std::function<void(NodeAttrs*)> attr_parser;
  std::map<std::string, shim::any> attrs;
  Op& describe(const std::string&) { return *this; }
  Op& set_num_inputs(int n) { num_inputs = n; return *this; }
  Op& set_num_outputs(int n) { num_outputs = n; return *this; }
  Op& set_attr_parser(std::function<void(NodeAttrs*)> f) { attr_parser = f; return *this; }
  template <typename ValueType>
  Op& set_attr(const std::string& key, const ValueType& v, int = 10) { attrs[key] = v; return *this; }
  Op& add_argument(const std::string&, const std::string&, const std::string&) { return *this; }
  template <typename T>
  Op& add_arguments(const T&) { return *this; }
  Op& add_alias(const std::string&) { return *this; }
  static Op& Get(const std::string& n) {
    static std::map<std::string, Op> reg;
    Op& o = reg[n];
    o.name = n;
    return o;
  }
};
}  // namespace nnvm
#define SHIM_CAT_(a, b) a##b
#define SHIM_CAT(a, b) SHIM_CAT_(a, b)
#define NNVM_REGISTER_OP(OpName) static ::nnvm::Op& SHIM_CAT(__nnvm_op_##OpName, __COUNTER__) = ::nnvm::Op::Get(#OpName)

// dmlc::Parameter: a functional miniature (kwargs parsing, defaults, ranges, enums, required fields)
namespace dmlc {
namespace parameter {
struct Manager {
  const std::map<std::string, std::string>* kwargs = nullptr;
  std::set<std::string> seen;
  std::vector<std::string> errors;
};
inline std::string trim(const std::string& s) {
  size_t b = s.find_first_not_of(" \t\n"), e = s.find_last_not_of(" \t\n");
  return b == std::string::npos ? "" : s.substr(b, e - b + 1);
}
inline std::vector<std::string> split_tuple(const std::string& s0) {
  std::string s = trim(s0);
  if (!s.empty() && (s[0] == '(' || s[0] == '[')) s = s.substr(1, s.size() - 2);
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, ','))
    if (!trim(item).empty()) out.push_back(trim(item));
  return out;
}
template <typename T> inline bool parse(const std::string& s, T* v) {
  std::istringstream is(trim(s));
  is >> *v;
  return !is.fail();
}
template <> inline bool parse<bool>(const std::string& s0, bool* v) {
  std::string s = trim(s0);
  std::transform(s.begin(), s.end(), s.begin(), ::tolower);
  if (s == "true" || s == "1") { *v = true; return true; }
  if (s == "false" || s == "0") { *v = false; return true; }
  return false;
}
template <> inline bool parse<std::string>(const std::string& s, std::string* v) { *v = s; return true; }
template <typename T> inline bool parse(const std::string& s, nnvm::Tuple<T>* v) {
  std::vector<T> vals;
  for (const std::string& it : split_tuple(s)) {
    T x;
    if (!parse(it, &x)) return false;
    vals.push_back(x);
  }
  v->assign(vals);
  return true;
}

template <typename T>
class FieldEntry {
 public:
  FieldEntry(Manager* m, const std::string& name, T* ptr) : m_(m), name_(name), ptr_(ptr) {}
  FieldEntry(FieldEntry&& o)
      : m_(o.m_), name_(o.name_), ptr_(o.ptr_), has_default_(o.has_default_), def_(o.def_), has_lo_(o.has_lo_),
        has_hi_(o.has_hi_), lo_(o.lo_), hi_(o.hi_), enums_(o.enums_), ndim_(o.ndim_), nonzero_(o.nonzero_) {
    o.m_ = nullptr;
  }
  FieldEntry& set_default(const T& v) { has_default_ = true; def_ = v; return *this; }
  FieldEntry& describe(const std::string&) { return *this; }
  FieldEntry& set_range(T lo, T hi) { has_lo_ = has_hi_ = true; lo_ = lo; hi_ = hi; return *this; }
  FieldEntry& set_lower_bound(T lo) { has_lo_ = true; lo_ = lo; return *this; }
  FieldEntry& add_enum(const std::string& k, int v) { enums_[k] = v; return *this; }
  FieldEntry& set_expect_ndim(int n) { ndim_ = n; return *this; }
  FieldEntry& enforce_nonzero() { nonzero_ = true; return *this; }
  ~FieldEntry() { if (m_) apply(); }

 private:
  template <typename U> static bool less(const U& a, const U& b, decltype(std::declval<U>() < std::declval<U>())* = nullptr) { return a < b; }
  static bool less(...) { return false; }
  template <typename U> static int ndim_of(const U& v, decltype(std::declval<U>().ndim())* = nullptr) { return (int)v.ndim(); }
  static int ndim_of(...) { return -1; }
  // tuple-like values (ndim() and operator[]): any zero entry?  Other types (NumericalParam has no operator[]): no
  template <typename U> static auto has_zero_impl(const U& v, int) -> decltype((void)v.ndim(), (void)v[0], bool()) {
    for (uint32_t i = 0; i < v.ndim(); ++i) if (v[i] == 0) return true;
    return false;
  }
  template <typename U> static bool has_zero_impl(const U&, long) { return false; }
  template <typename U> static bool has_zero(const U& v) { return has_zero_impl(v, 0); }
  template <typename U> void assign_enum(U*, int) {}
  void assign_enum(int* p, int v) { *p = v; }
  void apply() {
    m_->seen.insert(name_);
    auto it = m_->kwargs->find(name_);
    if (it == m_->kwargs->end()) {
      if (has_default_) *ptr_ = def_;
      else m_->errors.push_back("Required parameter " + name_ + " is missing");
      return;
    }
    if (!enums_.empty()) {
      auto e = enums_.find(trim(it->second));
      if (e == enums_.end()) { m_->errors.push_back("Invalid value for " + name_ + ": " + it->second); return; }
      assign_enum(ptr_, e->second);
      return;
    }
    if (!parse(it->second, ptr_)) { m_->errors.push_back("Invalid value for " + name_ + ": " + it->second); return; }
    if ((has_lo_ && less(*ptr_, lo_)) || (has_hi_ && less(hi_, *ptr_)))
      m_->errors.push_back("value " + it->second + " for Parameter " + name_ + " exceed bound");
    if (ndim_ >= 0 && ndim_of(*ptr_) != ndim_) m_->errors.push_back("Parameter " + name_ + ": wrong ndim");
    if (nonzero_ && has_zero(*ptr_)) m_->errors.push_back("Parameter " + name_ + ": zero entry");
  }
  Manager* m_;
  std::string name_;
  T* ptr_;
  bool has_default_ = false;
  T def_{};
  bool has_lo_ = false, has_hi_ = false;
  T lo_{}, hi_{};
  std::map<std::string, int> enums_;
  int ndim_ = -1;
  bool nonzero_ = false;
};
template <typename T>
inline FieldEntry<T> Declare(Manager* m, const char* name, T& ref) { return FieldEntry<T>(m, name, &ref); }
struct ParamFieldInfo {};
}  // namespace parameter

template <typename PType>
struct Parameter {
  template <typename Container>
  void Init(const Container& kwargs) {
    std::map<std::string, std::string> kw;
    for (const auto& kv : kwargs) kw[kv.first] = kv.second;
    parameter::Manager m;
    m.kwargs = &kw;
    static_cast<PType*>(this)->__DECLARE__(&m);
    for (const auto& kv : kw)
      if (!m.seen.count(kv.first)) m.errors.push_back("Cannot find argument '" + kv.first + "'");
    if (!m.errors.empty()) {
      std::string all;
      for (const auto& e : m.errors) all += e + "; ";
      throw ParamError(all);
    }
  }
  std::map<std::string, std::string> __DICT__() const { return {}; }
  static std::vector<parameter::ParamFieldInfo> __FIELDS__() { return {}; }
};
}  // namespace dmlc
#define DMLC_DECLARE_PARAMETER(PType) inline void __DECLARE__(::dmlc::parameter::Manager* manager)
#define DMLC_DECLARE_FIELD(FieldName) ::dmlc::parameter::Declare(manager, #FieldName, this->FieldName)
#define DMLC_REGISTER_PARAMETER(PType) struct SHIM_CAT(__shim_param_reg_##PType, __COUNTER__) {}

namespace mxnet {
using mshadow::cpu;
using mshadow::gpu;
using mshadow::index_t;
using nnvm::NodeAttrs;

class TShape : public nnvm::Tuple<int64_t> {
 public:
  TShape() {}
  explicit TShape(int ndim) { v_.assign(ndim, 0); }
  template <typename It>
  TShape(It b, It e) : nnvm::Tuple<int64_t>(b, e) {}
  template <int dim>
  TShape(const mshadow::Shape<dim>& s) { v_.assign(s.shape_, s.shape_ + dim); }  // NOLINT
  size_t Size() const { size_t s = 1; for (auto d : v_) s *= (size_t)d; return s; }
  size_t ProdShape(int b, int e) const { size_t s = 1; for (int i = b; i < e; ++i) s *= (size_t)v_[i]; return s; }
  template <int dim>
  mshadow::Shape<dim> get() const {
    CHECK_EQ((int)v_.size(), dim) << "TShape::get: dimension mismatch";
    mshadow::Shape<dim> s;
    for (int i = 0; i < dim; ++i) s[i] = (index_t)v_[i];
    return s;
  }
  bool operator==(const TShape& o) const { return v_ == o.v_; }
  bool operator!=(const TShape& o) const { return v_ != o.v_; }
};
typedef std::vector<TShape> ShapeVector;
}  // namespace mxnet
namespace dmlc { namespace parameter {
template <> inline bool parse<mxnet::TShape>(const std::string& s, mxnet::TShape* v) {
  std::vector<int64_t> vals;
  for (const std::string& it : split_tuple(s)) {
    int64_t x;
    if (!parse(it, &x)) return false;
    vals.push_back(x);
  }
  v->assign(vals);
  return true;
}
}}  // namespace dmlc::parameter

namespace mxnet {
class TBlob {
 public:
  void* dptr_ = nullptr;
  TShape shape_;
  int type_flag_ = mshadow::kFloat32;
  TBlob() {}
  template <typename DType>
  TBlob(DType* p, const TShape& s, int /*dev_mask*/ = cpu::kDevMask, int /*dev_id*/ = -1)
      : dptr_(p), shape_(s), type_flag_(mshadow::DataType<DType>::kFlag) {}
  size_t Size() const { return shape_.Size(); }
  index_t size(int i) const { return (index_t)shape_[i]; }
  int ndim() const { return (int)shape_.ndim(); }
  template <typename DType>
  DType* dptr() const {
    CHECK_EQ(type_flag_, mshadow::DataType<DType>::kFlag) << "TBlob.dptr(): data type do not match specified type";
    return static_cast<DType*>(dptr_);
  }
  template <typename Device, int dim, typename DType>
  mshadow::Tensor<Device, dim, DType> get(mshadow::Stream<Device>* = nullptr) const {
    return mshadow::Tensor<Device, dim, DType>(dptr<DType>(), shape_.get<dim>());
  }
  template <typename Device, int dim, typename DType>
  mshadow::Tensor<Device, dim, DType> get_with_shape(const mshadow::Shape<dim>& s, mshadow::Stream<Device>* = nullptr) const {
    CHECK_EQ(s.Size(), Size()) << "TBlob.get_with_shape: new and old shape do not match total elements";
    return mshadow::Tensor<Device, dim, DType>(dptr<DType>(), s);
  }
  template <typename Device, typename DType>
  mshadow::Tensor<Device, 1, DType> FlatTo1D(mshadow::Stream<Device>* = nullptr) const {
    return mshadow::Tensor<Device, 1, DType>(dptr<DType>(), mshadow::Shape1((index_t)Size()));
  }
  template <typename Device, typename DType>
  mshadow::Tensor<Device, 2, DType> FlatTo2D(mshadow::Stream<Device>* = nullptr) const {
    const index_t last = (index_t)shape_[shape_.ndim() - 1];
    return mshadow::Tensor<Device, 2, DType>(dptr<DType>(), mshadow::Shape2((index_t)(Size() / last), last));
  }
};

enum OpReqType { kNullOp, kWriteTo, kWriteInplace, kAddTo };

struct Resource {
  mutable std::shared_ptr<std::vector<char>> space = std::make_shared<std::vector<char>>();
  template <typename xpu, int ndim, typename DType>
  mshadow::Tensor<xpu, ndim, DType> get_space_typed(mshadow::Shape<ndim> shape, mshadow::Stream<xpu>*) const {
    space->resize(shape.Size() * sizeof(DType) + 64);
    return mshadow::Tensor<xpu, ndim, DType>(reinterpret_cast<DType*>(space->data()), shape);
  }
  template <typename xpu, int ndim>
  mshadow::Tensor<xpu, ndim, mshadow::real_t> get_space(mshadow::Shape<ndim> shape, mshadow::Stream<xpu>* s) const {
    return get_space_typed<xpu, ndim, mshadow::real_t>(shape, s);
  }
  mutable std::shared_ptr<std::vector<char>> host_space = std::make_shared<std::vector<char>>();
  template <int ndim, typename DType>
  mshadow::Tensor<cpu, ndim, DType> get_host_space_typed(mshadow::Shape<ndim> shape) const {
    host_space->resize(shape.Size() * sizeof(DType) + 64);  // (its own buffer: the device space must stay where it is)
    return mshadow::Tensor<cpu, ndim, DType>(reinterpret_cast<DType*>(host_space->data()), shape);
  }
};
struct ResourceRequest {
  enum Type { kRandom, kTempSpace };
  Type type;
  ResourceRequest(Type t) : type(t) {}  // NOLINT
};
struct Context {
  enum DeviceType { kCPU = cpu::kDevMask, kGPU = gpu::kDevMask };
  DeviceType dev_type = kCPU;
  int dev_id = 0;
  int dev_mask() const { return dev_type; }
  static Context CPU() { return Context(); }
};
struct RunContext {
  Context ctx;
};
struct OpContext {
  int is_train = 0;
  RunContext run_ctx;
  std::vector<Resource> requested;
  template <typename xpu>
  mshadow::Stream<xpu>* get_stream() const { static mshadow::Stream<xpu> s; return &s; }
};

class Operator {
 public:
  virtual ~Operator() {}
  virtual void Forward(const OpContext& ctx, const std::vector<TBlob>& in_data, const std::vector<OpReqType>& req,
                       const std::vector<TBlob>& out_data, const std::vector<TBlob>& aux_states) = 0;
  virtual void Backward(const OpContext&, const std::vector<TBlob>&, const std::vector<TBlob>&,
                        const std::vector<TBlob>&, const std::vector<OpReqType>&, const std::vector<TBlob>&,
                        const std::vector<TBlob>&) {
    LOG(FATAL) << "Backward is not implemented";
  }
};

class OperatorProperty {
 public:
  virtual ~OperatorProperty() {}
  virtual void Init(const std::vector<std::pair<std::string, std::string>>& kwargs) = 0;
  virtual std::map<std::string, std::string> GetParams() const = 0;
  virtual std::vector<std::string> ListArguments() const { return {"data"}; }
  virtual std::vector<std::string> ListOutputs() const { return {"output"}; }
  virtual std::vector<std::string> ListAuxiliaryStates() const { return {}; }
  virtual int NumOutputs() const { return (int)ListOutputs().size(); }
  virtual int NumVisibleOutputs() const { return NumOutputs(); }
  virtual bool InferShape(std::vector<TShape>* in_shape, std::vector<TShape>* out_shape,
                          std::vector<TShape>* aux_shape) const = 0;
  virtual bool InferType(std::vector<int>* in_type, std::vector<int>* out_type, std::vector<int>* aux_type) const {
    CHECK(!in_type->empty());
    const int t = (*in_type)[0];
    for (auto& x : *in_type) if (x == -1) x = t;
    out_type->assign(NumOutputs(), t);
    aux_type->assign(ListAuxiliaryStates().size(), t);
    return true;
  }
  virtual OperatorProperty* Copy() const = 0;
  virtual Operator* CreateOperator(Context) const { LOG(FATAL) << "CreateOperator not implemented"; return nullptr; }
  virtual Operator* CreateOperatorEx(Context ctx, std::vector<TShape>*, std::vector<int>*) const {
    return CreateOperator(ctx);
  }
  virtual std::string TypeString() const = 0;
  virtual std::vector<ResourceRequest> ForwardResource(const std::vector<TShape>&) const { return {}; }
  virtual std::vector<ResourceRequest> BackwardResource(const std::vector<TShape>&) const { return {}; }
  virtual std::vector<int> DeclareBackwardDependency(const std::vector<int>& out_grad, const std::vector<int>& in_data,
                                                     const std::vector<int>& out_data) const {
    std::vector<int> r = out_grad;
    r.insert(r.end(), in_data.begin(), in_data.end());
    r.insert(r.end(), out_data.begin(), out_data.end());
    return r;
  }
  virtual std::vector<std::pair<int, void*>> ForwardInplaceOption(const std::vector<int>&, const std::vector<void*>&) const { return {}; }
  virtual std::vector<std::pair<int, void*>> BackwardInplaceOption(const std::vector<int>&, const std::vector<int>&,
                                                                   const std::vector<int>&, const std::vector<void*>&) const { return {}; }
};

namespace shim_reg {
struct PropEntry {
  std::function<OperatorProperty*()> make;
  PropEntry& describe(const std::string&) { return *this; }
  PropEntry& add_argument(const std::string&, const std::string&, const std::string&) { return *this; }
  template <typename T> PropEntry& add_arguments(const T&) { return *this; }
  PropEntry& set_return_type(const std::string&) { return *this; }
  PropEntry& set_key_var_num_args(const std::string&) { return *this; }
  static std::map<std::string, PropEntry>& All() { static std::map<std::string, PropEntry> r; return r; }
  static PropEntry& Register(const std::string& name, std::function<OperatorProperty*()> f) {
    PropEntry& e = All()[name];
    e.make = f;
    return e;
  }
};
}  // namespace shim_reg

typedef std::function<bool(const nnvm::NodeAttrs&, ShapeVector*, ShapeVector*)> FInferShape;
typedef std::function<void(const nnvm::NodeAttrs&, const OpContext&, const std::vector<TBlob>&,
                           const std::vector<OpReqType>&, const std::vector<TBlob>&)> FCompute;

namespace op {
using nnvm::NodeAttrs;
using mxnet::FCompute;
using mxnet::FInferShape;
template <typename PType>
inline void ParamParser(nnvm::NodeAttrs* attrs) {
  PType param;
  param.Init(attrs->dict);
  attrs->parsed = param;
}
inline std::vector<nnvm::NodeEntry> MakeGradNode(const char*, const nnvm::NodePtr&, const std::vector<nnvm::NodeEntry>& heads,
                                                 const std::unordered_map<std::string, std::string>&) {
  return heads;
}
template <bool is_integer, typename xpu, typename DType>
inline void Fill(mshadow::Stream<xpu>*, const TBlob& b, const OpReqType req, DType v) {
  if (req == kNullOp) return;
  DType* p = b.dptr<DType>();
  std::fill(p, p + b.Size(), v);
}
namespace mxnet_op {
// mxnet_op::Kernel<OP, cpu>::Launch runs OP::Map(i, args...) for i in [0, N) (OpenMP-parallel upstream; the
// gpu specialisation launches one thread per i).  Serial, in index order, for both devices here.
template <typename OP, typename xpu>
struct Kernel {
  template <typename... Args>
  static void Launch(mshadow::Stream<xpu>*, const int N, Args... args) {
    for (int i = 0; i < N; ++i) OP::Map(i, args...);
  }
};
}  // namespace mxnet_op
namespace mshadow_op {
struct minimum { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return a < b ? a : b; } };
struct maximum { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return a > b ? a : b; } };
struct floor { template <typename DType> MSHADOW_XINLINE static DType Map(DType a) { return DType(::floorf(a)); } };
struct ceil { template <typename DType> MSHADOW_XINLINE static DType Map(DType a) { return DType(::ceilf(a)); } };
template <> MSHADOW_XINLINE double floor::Map<double>(double a) { return ::floor(a); }
template <> MSHADOW_XINLINE double ceil::Map<double>(double a) { return ::ceil(a); }
// MXNet's src/operator/mshadow_op.h (not in the reference tree), float instantiations: math::exp / log / pow are
// ::expf / ::logf / ::powf for float operands
struct identity { template <typename DType> MSHADOW_XINLINE static DType Map(DType a) { return a; } };
struct negation { template <typename DType> MSHADOW_XINLINE static DType Map(DType a) { return DType(-a); } };
struct sigmoid { template <typename DType> MSHADOW_XINLINE static DType Map(DType a) { return DType(1.0f / (1.0f + ::expf(-a))); } };
struct log { template <typename DType> MSHADOW_XINLINE static DType Map(DType a) { return DType(::logf(a)); } };
struct power { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return DType(::powf(a, b)); } };
struct plus { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return a + b; } };
struct eq { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return a == b ? DType(1) : DType(0); } };
struct le { template <typename DType> MSHADOW_XINLINE static DType Map(DType a, DType b) { return a <= b ? DType(1) : DType(0); } };
}  // namespace mshadow_op
// src/operator/tensor/indexing_op.h one_hot and control_flow_op.h where (MXNet, not in the reference tree)
#define KERNEL_ASSIGN(out, req, val)       \
  {                                        \
    switch (req) {                         \
      case kNullOp: break;                 \
      case kWriteTo:                       \
      case kWriteInplace: (out) = (val); break; \
      case kAddTo: (out) += (val); break;  \
    }                                      \
  }
template <int req>
struct one_hot {
  template <typename DType, typename IType>
  MSHADOW_XINLINE static void Map(int i, DType* out, const IType* indices, int depth, DType on_value) {
    const int offset = i * depth;
    const int j = static_cast<int>(indices[i]);
    if (j >= 0 && j < depth) KERNEL_ASSIGN(out[offset + j], req, on_value);
  }
};
template <int req>
struct where {
  template <typename DType, typename CType>
  MSHADOW_XINLINE static void Map(int i, DType* out, const CType* cond, const DType* x, const DType* y) {
    KERNEL_ASSIGN(out[i], req, (0 != cond[i] ? x[i] : y[i]));
  }
};
}  // namespace op
}  // namespace mxnet

// CUDA atomicAdd as the serial emulation sees it: a plain add, applied in index order
template <typename T>
inline T atomicAdd(T* addr, T v) { T old = *addr; *addr = old + v; return old; }

#define MXNET_REGISTER_OP_PROPERTY(name, OperatorPropertyType)                                      \
  static ::mxnet::shim_reg::PropEntry& SHIM_CAT(__shim_prop_##name, __COUNTER__) =                  \
      ::mxnet::shim_reg::PropEntry::Register(#name, []() -> ::mxnet::OperatorProperty* { return new OperatorPropertyType(); })
// operator_common.h's form for MXNET_USE_CUDA == 0
#ifdef SHIM_GPU_DISPATCH  // the file's .cu (with CreateOp<gpu>) is part of the build: a GPU context gets the gpu operator
#define DO_BIND_DISPATCH(Method, ...)                                             \
  if (ctx.dev_mask() == ::mshadow::cpu::kDevMask) {                               \
    return Method<::mshadow::cpu>(__VA_ARGS__);                                   \
  } else {                                                                        \
    return Method<::mshadow::gpu>(__VA_ARGS__);                                   \
  }
#else
#define DO_BIND_DISPATCH(Method, ...)                                             \
  if (ctx.dev_mask() == ::mshadow::cpu::kDevMask) {                               \
    return Method<::mshadow::cpu>(__VA_ARGS__);                                   \
  } else {                                                                        \
    LOG(FATAL) << "GPU is not enabled";                                           \
    return nullptr;                                                               \
  }
#endif
#define ADD_FILELINE "\n\nFrom:" __FILE__
// operator_common.h's Assign(out, req, exp) for the plain-value uses in the files we build
#define Assign(out, req, exp)                              \
  {                                                        \
    switch (req) {                                         \
      case ::mxnet::kNullOp: break;                        \
      case ::mxnet::kWriteTo:                              \
      case ::mxnet::kWriteInplace: (out) = (exp); break;   \
      case ::mxnet::kAddTo: LOG(FATAL) << "shim: kAddTo"; break; \
    }                                                      \
  }
// operator_common.h
#define UNIFORM_TYPE_CHECK(type, expected, arg) \
  CHECK_EQ(type, expected) << "This layer requires uniform type. Expected vs given at " << arg
#define SHAPE_ASSIGN_CHECK(shape_array, index, shape)                            \
  {                                                                              \
    if ((shape_array)[index].ndim() == 0) (shape_array)[index] = ::mxnet::TShape(shape); \
  }
