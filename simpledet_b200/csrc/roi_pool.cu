// ROIPooling_v1 forward/backward for sm_100a.
// Reference semantics: operator_cxx/roi_pooling_v1.cu:49-113 (forward), :116-152 (backward),
// op wrapper roi_pooling_v1-inl.h:63-137 (out pre-filled -FLT_MAX / -1, grad zero fill).
//
// Layout: one warp owns one (roi, channel) plane of PH*PW bins; the integer bin geometry is
// computed once per lane-bin and the bin is scanned row-major (the reference's order, so the
// first maximum wins exactly as `>` does there).  Adjacent lanes scan adjacent bins of the same
// feature rows, so the warp's loads fall in the same few 128-byte lines.
#include <cfloat>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
roi_pool_v1_fwd_kernel(const float* __restrict__ data, const float* __restrict__ rois,
                       float* __restrict__ out, float* __restrict__ max_idx, const int B,
                       const int C, const int H, const int W, const int PH, const int PW,
                       const float scale, const size_t count) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < count;
       index += (size_t)gridDim.x * blockDim.x) {
    const int pw = (int)(index % PW);
    const int ph = (int)((index / PW) % PH);
    const int c = (int)((index / PW / PH) % C);
    const int n = (int)(index / PW / PH / C);
    const float* r = rois + (size_t)n * 5;
    const int bi = (int)__ldg(r);
    // round() = half away from zero, as C's round() in roi_pooling_v1.cu:69-72
    const int rsw = (int)roundf(__fmul_rn(__ldg(r + 1), scale));
    const int rsh = (int)roundf(__fmul_rn(__ldg(r + 2), scale));
    const int rew = (int)roundf(__fmul_rn(__ldg(r + 3), scale));
    const int reh = (int)roundf(__fmul_rn(__ldg(r + 4), scale));
    const int rw = max(rew - rsw + 1, 1), rh = max(reh - rsh + 1, 1);
    const float bh = __fdiv_rn((float)rh, (float)PH), bw = __fdiv_rn((float)rw, (float)PW);
    int hs = (int)floorf(__fmul_rn((float)ph, bh)), ws = (int)floorf(__fmul_rn((float)pw, bw));
    int he = (int)ceilf(__fmul_rn((float)(ph + 1), bh)), we = (int)ceilf(__fmul_rn((float)(pw + 1), bw));
    hs = min(max(hs + rsh, 0), H);
    he = min(max(he + rsh, 0), H);
    ws = min(max(ws + rsw, 0), W);
    we = min(max(we + rsw, 0), W);
    const bool empty = (he <= hs) || (we <= ws);
    float best = empty ? 0.f : -FLT_MAX;
    int arg = -1;
    if (bi >= 0 && bi < B) {
      const float* plane = data + ((size_t)bi * C + c) * H * W;
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) {
          const float v = __ldg(plane + h * W + w);
          if (v > best) {
            best = v;
            arg = h * W + w;
          }
        }
    }
    out[index] = best;
    if (max_idx) max_idx[index] = (float)arg;
  }
}

__global__ void __launch_bounds__(256)
roi_pool_v1_bwd_kernel(const float* __restrict__ ograd, const float* __restrict__ max_idx,
                       const float* __restrict__ rois, float* __restrict__ grad, const int B,
                       const int C, const int H, const int W, const size_t PP, const size_t count) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < count;
       index += (size_t)gridDim.x * blockDim.x) {
    const int arg = (int)__ldg(max_idx + index);
    if (arg == -1) continue;
    const size_t nc = index / PP;
    const int c = (int)(nc % C);
    const size_t n = nc / C;
    const int bi = (int)__ldg(rois + n * 5);
    if (bi < 0 || bi >= B) continue;
    atomicAdd(grad + ((size_t)bi * C + c) * H * W + arg, __ldg(ograd + index));
  }
}

}  // namespace

extern "C" int sdet_roi_pooling_v1_forward(const float* data, const float* rois, float* out,
                                           float* max_idx, int B, int R, int C, int H, int W,
                                           int pooled_h, int pooled_w, float spatial_scale,
                                           void* stream) {
  SDET_REQUIRE(data && rois && out, "data, rois and out must be non-NULL");
  SDET_REQUIRE(B > 0 && R > 0 && C > 0 && H > 0 && W > 0, "shape must be positive");
  SDET_REQUIRE(pooled_h > 0 && pooled_w > 0, "pooled_size must be non-zero (enforce_nonzero)");
  // DMLC_DECLARE_FIELD(spatial_scale).set_range(0.0, 1.0)  (roi_pooling_v1-inl.h:57)
  SDET_REQUIRE(spatial_scale >= 0.f && spatial_scale <= 1.f, "spatial_scale must be in [0, 1]");
  const size_t count = (size_t)R * C * pooled_h * pooled_w;
  const int threads = 256;
  size_t blocks = (count + threads - 1) / threads;
  if (blocks > 148 * 64) blocks = 148 * 64;
  roi_pool_v1_fwd_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      data, rois, out, max_idx, B, C, H, W, pooled_h, pooled_w, spatial_scale, count);
  SDET_LAUNCH_CHECK("roi_pool_v1_fwd_kernel");
  return SDET_OK;
}

extern "C" int sdet_roi_pooling_v1_backward(const float* ograd, const float* max_idx,
                                            const float* rois, float* grad_data, float* grad_rois,
                                            int B, int R, int C, int H, int W, int pooled_h,
                                            int pooled_w, int accumulate, void* stream) {
  SDET_REQUIRE(ograd && max_idx && rois && grad_data, "NULL argument");
  SDET_REQUIRE(B > 0 && R > 0 && C > 0 && H > 0 && W > 0 && pooled_h > 0 && pooled_w > 0,
               "shape must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate)  // grad_in = 0.0f on kWriteTo (roi_pooling_v1-inl.h:125-127)
    SDET_CUDA(cudaMemsetAsync(grad_data, 0, sizeof(float) * (size_t)B * C * H * W, st));
  if (grad_rois) SDET_CUDA(cudaMemsetAsync(grad_rois, 0, sizeof(float) * (size_t)R * 5, st));
  const size_t count = (size_t)R * C * pooled_h * pooled_w;
  const int threads = 256;
  size_t blocks = (count + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  roi_pool_v1_bwd_kernel<<<(unsigned)blocks, threads, 0, st>>>(
      ograd, max_idx, rois, grad_data, B, C, H, W, (size_t)pooled_h * pooled_w, count);
  SDET_LAUNCH_CHECK("roi_pool_v1_bwd_kernel");
  return SDET_OK;
}
