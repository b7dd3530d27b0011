// Detection losses, fused: _contrib_FocalLoss, _contrib_BBoxNorm, _contrib_SigmoidCrossEntropy.
//
// The reference's FocalLoss backward is a chain of mshadow expressions with 7 full-size
// temporaries and a 1.5 GB workspace (operator_cxx/contrib/focal_loss-inl.h:148-230,
// models/retinanet/builder.py:314): >= 10 HBM passes over a 128 MB tensor.  Here: one small
// reduction over the labels + ONE pass that reads `out`, the label row and writes the gradient
// (2 x sz(data) + sz(label) of traffic — the algorithmic minimum).
//
//   sdet_focal_loss_forward / _backward   focal_loss-inl.h:90-114 / :116-231
//   sdet_bbox_norm_backward               bbox_norm-inl.h:99-129  (forward is identity)
//   sdet_sigmoid_ce_forward / _backward   sigmoid_cross_entropy.cu:45-129
#include "common.cuh"

namespace {

// count[0] += #(label >= 1).  count must be zeroed by the caller (memsetAsync).
__global__ void __launch_bounds__(256)
count_positive_kernel(const float* __restrict__ label, const size_t n, float* __restrict__ count) {
  float c = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    c += (1.f <= __ldg(label + i)) ? 1.f : 0.f;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s[w];
    if (t != 0.f) atomicAdd(count, t);  // integer-valued floats < 2^24: exact in any order
  }
}

__global__ void __launch_bounds__(256)
sigmoid_kernel(const float* __restrict__ x, float* __restrict__ y, const size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x[i])));  // mshadow_op::sigmoid
}

struct FocalParams {
  float alpha, gamma, grad_scale;
  int normalization, K, B;
};

// thread per element, K innermost: the label row is a broadcast within a row's threads
__global__ void __launch_bounds__(256)
focal_backward_kernel(const float* __restrict__ out, const float* __restrict__ label,
                      const float* __restrict__ ograd, const float* __restrict__ count,
                      float* __restrict__ gdata, const size_t n, const FocalParams p) {
  const float temp = __fadd_rn(__ldg(count), 1.f);  // sum(label >= 1) + 1   (:218-220)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / p.K;
    const int k = (int)(i - r * p.K);
    const float l = __ldg(label + r);
    const float pr = __ldg(out + i);
    const float lm1 = __fsub_rn(l, 1.f);
    const int hot = (int)lm1;
    float g;
    if (lm1 >= 0.f && hot == k) {
      const float a = __fmul_rn(p.alpha, powf(__fsub_rn(1.f, pr), p.gamma));
      const float b = __fsub_rn(__fadd_rn(__fmul_rn(__fmul_rn(p.gamma, pr), logf(__fadd_rn(pr, 1e-14f))), pr), 1.f);
      g = __fmul_rn(a, b);
    } else {
      const float q = __fsub_rn(1.f, pr);
      const float a = __fmul_rn(__fsub_rn(1.f, p.alpha), powf(pr, p.gamma));
      const float b = __fsub_rn(__fmul_rn(__fmul_rn(p.gamma, q), logf(__fadd_rn(q, 1e-14f))), pr);
      g = -__fmul_rn(a, b);
    }
    if (l == -1.f) g = 0.f;
    if (ograd) g = __fmul_rn(g, __ldg(ograd + i));
    if (p.normalization == 2) g = __fdiv_rn(__fmul_rn(g, p.grad_scale), temp);
    else if (p.normalization == 1) g = __fmul_rn(g, __fdiv_rn(p.grad_scale, (float)p.B));
    else g = __fmul_rn(g, p.grad_scale);
    gdata[i] = g;
  }
}

__global__ void __launch_bounds__(256)
bbox_norm_backward_kernel(const float* __restrict__ gout, const float* __restrict__ count,
                          float* __restrict__ gdata, const size_t n) {
  float temp = __fadd_rn(__ldg(count), 1.f);
  temp = 1.f > temp ? 1.f : temp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    gdata[i] = __fdiv_rn(__ldg(gout + i), temp);
}

// per row r of (R,D): sums[r*2] += loss, sums[r*2+1] += count.  grid = (blocks, R)
__global__ void __launch_bounds__(256)
sigmoid_ce_reduce_kernel(const float* __restrict__ x, const float* __restrict__ t, const size_t D,
                         float* __restrict__ sums, const int want_loss) {
  const int r = blockIdx.y;
  const float* xr = x + (size_t)r * D;
  const float* tr = t + (size_t)r * D;
  float ls = 0.f, cs = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += (size_t)gridDim.x * blockDim.x) {
    const float ti = __ldg(tr + i);
    if (ti != -1.f) {
      cs += 1.f;
      if (want_loss) {
        const float xi = __ldg(xr + i);
        const int ge = xi >= 0.f;
        // -1. * x * (t - (x>=0)) + logf(1 + expf(x - 2*x*(x>=0)))  with the double promotions
        const double a = __dmul_rn(__dmul_rn(-1.0, (double)xi), (double)__fsub_rn(ti, (float)ge));
        const float e = expf(__fsub_rn(xi, __fmul_rn(__fmul_rn(2.f, xi), (float)ge)));
        ls += (float)__dadd_rn(a, (double)logf(__fadd_rn(1.f, e)));
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    ls += __shfl_xor_sync(0xffffffffu, ls, o);
    cs += __shfl_xor_sync(0xffffffffu, cs, o);
  }
  __shared__ float s[16];
  if ((threadIdx.x & 31) == 0) {
    s[threadIdx.x >> 5] = ls;
    s[8 + (threadIdx.x >> 5)] = cs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += s[w];
      c += s[8 + w];
    }
    atomicAdd(sums + r * 2, a);
    atomicAdd(sums + r * 2 + 1, c);
  }
}

__global__ void sigmoid_ce_finish_kernel(const float* __restrict__ sums, float* __restrict__ out, const int R) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) out[r] = __fdiv_rn(sums[r * 2], __fadd_rn(sums[r * 2 + 1], 1e-5f));
}

__global__ void __launch_bounds__(256)
sigmoid_ce_backward_kernel(const float* __restrict__ x, const float* __restrict__ t,
                           const float* __restrict__ sums, float* __restrict__ dx, const size_t D,
                           const float scale) {
  const int r = blockIdx.y;
  const float cnt = __fadd_rn(sums[r * 2 + 1], 1e-5f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = (size_t)r * D + i;
    const float ti = __ldg(t + e);
    float d = 0.f;
    if (ti != -1.f) {
      const double sg = __ddiv_rn(1.0, __dadd_rn(1.0, (double)expf(-__ldg(x + e))));
      d = (float)__dsub_rn(sg, (double)ti);
    }
    dx[e] = __fmul_rn(__fdiv_rn(d, cnt), scale);
  }
}

unsigned grid_for(size_t n) {
  size_t b = (n + 255) / 256;
  return (unsigned)(b > 148 * 16 ? 148 * 16 : (b ? b : 1));
}

}  // namespace

extern "C" int sdet_focal_loss_forward(const float* data, float* out, size_t n, void* stream) {
  SDET_REQUIRE(data && out && n > 0, "bad argument");
  sigmoid_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(data, out, n);
  SDET_LAUNCH_CHECK("sigmoid_kernel");
  return SDET_OK;
}

extern "C" int sdet_focal_loss_backward(const float* out, const float* label, const float* ograd,
                                        float* gdata, int B, int N, int K, float alpha, float gamma,
                                        float grad_scale, int normalization, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  SDET_REQUIRE(out && label && gdata && workspace, "NULL argument");
  SDET_REQUIRE(B > 0 && N > 0 && K > 0, "bad shape");
  SDET_REQUIRE(normalization >= 0 && normalization <= 2, "normalization must be null(0) batch(1) valid(2)");
  if (workspace_bytes < 4) return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need 4 bytes");
  cudaStream_t st = (cudaStream_t)stream;
  float* count = static_cast<float*>(workspace);
  SDET_CUDA(cudaMemsetAsync(count, 0, 4, st));
  const size_t rows = (size_t)B * N, n = rows * K;
  if (normalization == 2) {
    count_positive_kernel<<<grid_for(rows), 256, 0, st>>>(label, rows, count);
    SDET_LAUNCH_CHECK("count_positive_kernel");
  }
  FocalParams p{alpha, gamma, grad_scale, normalization, K, B};
  focal_backward_kernel<<<grid_for(n), 256, 0, st>>>(out, label, ograd, count, gdata, n, p);
  SDET_LAUNCH_CHECK("focal_backward_kernel");
  return SDET_OK;
}

extern "C" int sdet_bbox_norm_backward(const float* gout, const float* label, float* gdata, size_t n,
                                       size_t n_label, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  SDET_REQUIRE(gout && label && gdata && workspace && n > 0 && n_label > 0, "bad argument");
  if (workspace_bytes < 4) return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need 4 bytes");
  cudaStream_t st = (cudaStream_t)stream;
  float* count = static_cast<float*>(workspace);
  SDET_CUDA(cudaMemsetAsync(count, 0, 4, st));
  count_positive_kernel<<<grid_for(n_label), 256, 0, st>>>(label, n_label, count);
  SDET_LAUNCH_CHECK("count_positive_kernel");
  bbox_norm_backward_kernel<<<grid_for(n), 256, 0, st>>>(gout, count, gdata, n);
  SDET_LAUNCH_CHECK("bbox_norm_backward_kernel");
  return SDET_OK;
}

static int sigmoid_ce_common(const float* data, const float* label, int R, size_t D, float* sums,
                             int want_loss, cudaStream_t st) {
  SDET_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * (size_t)R, st));
  unsigned gx = grid_for(D);
  if (gx > 512) gx = 512;
  dim3 grid(gx, (unsigned)R);
  sigmoid_ce_reduce_kernel<<<grid, 256, 0, st>>>(data, label, D, sums, want_loss);
  SDET_LAUNCH_CHECK("sigmoid_ce_reduce_kernel");
  return SDET_OK;
}

extern "C" int sdet_sigmoid_ce_forward(const float* data, const float* label, float* out, int R, size_t D,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  SDET_REQUIRE(data && label && out && workspace && R > 0 && D > 0, "bad argument");
  if (workspace_bytes < sizeof(float) * 2 * (size_t)R)
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", sizeof(float) * 2 * (size_t)R);
  cudaStream_t st = (cudaStream_t)stream;
  float* sums = static_cast<float*>(workspace);
  if (int rc = sigmoid_ce_common(data, label, R, D, sums, 1, st)) return rc;
  sigmoid_ce_finish_kernel<<<(R + 127) / 128, 128, 0, st>>>(sums, out, R);
  SDET_LAUNCH_CHECK("sigmoid_ce_finish_kernel");
  return SDET_OK;
}

extern "C" int sdet_sigmoid_ce_backward(const float* data, const float* label, float* d_data, int R,
                                        size_t D, float scale, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  SDET_REQUIRE(data && label && d_data && workspace && R > 0 && D > 0, "bad argument");
  if (workspace_bytes < sizeof(float) * 2 * (size_t)R)
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", sizeof(float) * 2 * (size_t)R);
  cudaStream_t st = (cudaStream_t)stream;
  float* sums = static_cast<float*>(workspace);
  if (int rc = sigmoid_ce_common(data, label, R, D, sums, 0, st)) return rc;
  unsigned gx = grid_for(D);
  dim3 grid(gx, (unsigned)R);
  sigmoid_ce_backward_kernel<<<grid, 256, 0, st>>>(data, label, sums, d_data, D, scale);
  SDET_LAUNCH_CHECK("sigmoid_ce_backward_kernel");
  return SDET_OK;
}
