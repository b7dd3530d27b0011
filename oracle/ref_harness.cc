// TEST INFRASTRUCTURE (oracle/): C entry points into the reference's own operator sources, compiled unmodified
// from /root/reference/operator_cxx against oracle/shim/mxnet_shim.h (recipe: oracle/build_ref_cxx.py, output
// oracle/_ref/libref_cxx.so).  This file contains no operator arithmetic: it builds TBlobs over the caller's
// numpy buffers, parses the kwargs with the operator's own dmlc::Parameter declaration, calls the operator's
// own InferShape / Forward (legacy OperatorProperty ops) or FInferShape / FCompute<cpu|gpu> (nnvm ops), and
// owns the `rand()` the reference's std::random_shuffle draws from.
#include <mutex>

#include "shim/mxnet_shim.h"

namespace {
thread_local std::string g_err;
int g_rand_mode = 0;   // 0: constant g_rand_value; 1: values from g_rand_seq (cycled)
int g_rand_value = 0;
std::vector<int> g_rand_seq;
size_t g_rand_pos = 0;
long long g_rand_calls = 0;

std::vector<std::pair<std::string, std::string>> parse_kwargs(const char* s) {
  std::vector<std::pair<std::string, std::string>> out;
  if (!s) return out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, '|')) {
    const size_t eq = item.find('=');
    if (eq == std::string::npos) continue;
    out.emplace_back(item.substr(0, eq), item.substr(eq + 1));
  }
  return out;
}

mxnet::TShape make_shape(int ndim, const int64_t* dims) { return mxnet::TShape(dims, dims + ndim); }

template <typename F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}
}  // namespace

extern "C" {

int ref_shim_rand(void) {
  ++g_rand_calls;
  if (g_rand_mode == 1 && !g_rand_seq.empty()) return g_rand_seq[g_rand_pos++ % g_rand_seq.size()];
  return g_rand_value;
}
void ref_set_rand_const(int v) { g_rand_mode = 0; g_rand_value = v; g_rand_calls = 0; }
void ref_set_rand_seq(const int* v, int n) {
  g_rand_mode = 1;
  g_rand_seq.assign(v, v + n);
  g_rand_pos = 0;
  g_rand_calls = 0;
}
long long ref_rand_calls(void) { return g_rand_calls; }
const char* ref_last_error(void) { return g_err.c_str(); }

// 1 if `op` was registered by MXNET_REGISTER_OP_PROPERTY, 2 if by NNVM_REGISTER_OP (with an FCompute), else 0
int ref_op_kind(const char* op) {
  if (mxnet::shim_reg::PropEntry::All().count(op)) return 1;
  nnvm::Op& o = nnvm::Op::Get(op);
  for (const auto& kv : o.attrs)
    if (kv.first.rfind("FCompute", 0) == 0) return 2;
  return 0;
}

// Shapes are rows of 8 int64.  Returns the operator's own output shapes and visible-output count.
int ref_infer_shape(const char* op, const char* kwargs, int n_in, const int* in_ndims, const int64_t* in_dims,
                    int* n_out, int* out_ndims, int64_t* out_dims, int* n_visible) {
  return guarded([&] {
    std::vector<mxnet::TShape> in_shape, out_shape, aux_shape;
    for (int i = 0; i < n_in; ++i) in_shape.push_back(make_shape(in_ndims[i], in_dims + 8 * i));
    const auto kw = parse_kwargs(kwargs);
    const int kind = ref_op_kind(op);
    if (kind == 1) {
      std::unique_ptr<mxnet::OperatorProperty> prop(mxnet::shim_reg::PropEntry::All()[op].make());
      prop->Init(kw);
      if (!prop->InferShape(&in_shape, &out_shape, &aux_shape)) throw dmlc::Error("InferShape returned false");
      *n_visible = prop->NumVisibleOutputs();
    } else if (kind == 2) {
      nnvm::Op& o = nnvm::Op::Get(op);
      nnvm::NodeAttrs attrs;
      for (const auto& kv : kw) attrs.dict[kv.first] = kv.second;
      if (o.attr_parser) o.attr_parser(&attrs);
      const auto& f = o.attrs.at("FInferShape").get<mxnet::FInferShape>();
      if (!f(attrs, &in_shape, &out_shape)) throw dmlc::Error("FInferShape returned false");
      *n_visible = o.attrs.count("FNumVisibleOutputs")
                       ? (int)o.attrs.at("FNumVisibleOutputs").get<nnvm::FNumVisibleOutputs>()(attrs)
                       : (int)out_shape.size();
    } else {
      throw dmlc::Error(std::string("unknown operator ") + op);
    }
    *n_out = (int)out_shape.size();
    for (size_t i = 0; i < out_shape.size(); ++i) {
      out_ndims[i] = (int)out_shape[i].ndim();
      for (uint32_t d = 0; d < out_shape[i].ndim(); ++d) out_dims[8 * i + d] = out_shape[i][d];
    }
  });
}

// float32 tensors over caller memory.  dev: "cpu" or "gpu" selects FCompute<dev> for nnvm ops (the "gpu"
// functors run on the host through the shim's serial Kernel::Launch / atomicAdd).  reqs: OpReqType per output.
int ref_forward(const char* op, const char* kwargs, const char* dev, int n_in, void** in_ptrs, const int* in_ndims,
                const int64_t* in_dims, int n_out, void** out_ptrs, const int* out_ndims, const int64_t* out_dims,
                const int* reqs) {
  return guarded([&] {
    std::vector<mxnet::TBlob> in_data, out_data, aux;
    std::vector<mxnet::TShape> in_shape;
    std::vector<mxnet::OpReqType> req;
    for (int i = 0; i < n_in; ++i) {
      in_shape.push_back(make_shape(in_ndims[i], in_dims + 8 * i));
      in_data.emplace_back(static_cast<float*>(in_ptrs[i]), in_shape.back());
    }
    for (int i = 0; i < n_out; ++i) {
      out_data.emplace_back(static_cast<float*>(out_ptrs[i]), make_shape(out_ndims[i], out_dims + 8 * i));
      req.push_back((mxnet::OpReqType)reqs[i]);
    }
    const auto kw = parse_kwargs(kwargs);
    mxnet::OpContext ctx;
    const int kind = ref_op_kind(op);
    if (kind == 1) {
      std::unique_ptr<mxnet::OperatorProperty> prop(mxnet::shim_reg::PropEntry::All()[op].make());
      prop->Init(kw);
      std::vector<int> in_type(n_in, mshadow::kFloat32);
      for (size_t i = 0; i < prop->ForwardResource(in_shape).size(); ++i) ctx.requested.emplace_back();
      mxnet::Context dctx = mxnet::Context::CPU();
      if (std::string(dev) == "gpu") dctx.dev_type = mxnet::Context::kGPU;  // operators whose .cu is in the build
      std::unique_ptr<mxnet::Operator> o(prop->CreateOperatorEx(dctx, &in_shape, &in_type));
      if (!o) throw dmlc::Error("CreateOperatorEx returned NULL");
      o->Forward(ctx, in_data, req, out_data, aux);
    } else if (kind == 2) {
      nnvm::Op& o = nnvm::Op::Get(op);
      nnvm::NodeAttrs attrs;
      for (const auto& kv : kw) attrs.dict[kv.first] = kv.second;
      if (o.attr_parser) o.attr_parser(&attrs);
      const std::string key = std::string("FCompute<") + dev + ">";
      if (!o.attrs.count(key)) throw dmlc::Error(key + " is not registered for " + op);
      o.attrs.at(key).get<mxnet::FCompute>()(attrs, ctx, in_data, req, out_data);
    } else {
      throw dmlc::Error(std::string("unknown operator ") + op);
    }
  });
}

// Backward of a legacy OperatorProperty operator: out_grad / in_data / out_data in, in_grad out (float32 over caller
// memory; shapes as rows of 8 int64).  The operator object is created exactly as ref_forward creates it.
int ref_backward(const char* op, const char* kwargs, const char* dev, int n_og, void** og_ptrs, const int* og_ndims, const int64_t* og_dims,
                 int n_in, void** in_ptrs, const int* in_ndims, const int64_t* in_dims, int n_out, void** out_ptrs,
                 const int* out_ndims, const int64_t* out_dims, void** ig_ptrs, const int* reqs) {
  return guarded([&] {
    std::vector<mxnet::TBlob> out_grad, in_data, out_data, in_grad, aux;
    std::vector<mxnet::TShape> in_shape;
    std::vector<mxnet::OpReqType> req;
    for (int i = 0; i < n_og; ++i)
      out_grad.emplace_back(static_cast<float*>(og_ptrs[i]), make_shape(og_ndims[i], og_dims + 8 * i));
    for (int i = 0; i < n_in; ++i) {
      in_shape.push_back(make_shape(in_ndims[i], in_dims + 8 * i));
      in_data.emplace_back(static_cast<float*>(in_ptrs[i]), in_shape.back());
      in_grad.emplace_back(static_cast<float*>(ig_ptrs[i]), in_shape.back());
      req.push_back((mxnet::OpReqType)reqs[i]);
    }
    for (int i = 0; i < n_out; ++i)
      out_data.emplace_back(static_cast<float*>(out_ptrs[i]), make_shape(out_ndims[i], out_dims + 8 * i));
    if (ref_op_kind(op) != 1) throw dmlc::Error(std::string(op) + " is not an OperatorProperty operator");
    std::unique_ptr<mxnet::OperatorProperty> prop(mxnet::shim_reg::PropEntry::All()[op].make());
    prop->Init(parse_kwargs(kwargs));
    mxnet::OpContext ctx;
    ctx.is_train = 1;
    for (int i = 0; i < 4; ++i) ctx.requested.emplace_back();  // BackwardResource: at most one temp space in these operators
    std::vector<int> in_type(n_in, mshadow::kFloat32);
    mxnet::Context dctx = mxnet::Context::CPU();
    if (std::string(dev) == "gpu") dctx.dev_type = mxnet::Context::kGPU;
    std::unique_ptr<mxnet::Operator> o(prop->CreateOperatorEx(dctx, &in_shape, &in_type));
    if (!o) throw dmlc::Error("CreateOperatorEx returned NULL");
    o->Backward(ctx, out_grad, in_data, out_data, req, in_grad, aux);
  });
}

}  // extern "C"
