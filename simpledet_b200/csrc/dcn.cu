// Deformable convolution v1 sampling (the gather / scatter around the dense GEMM).
//
// The operator itself (`mx.sym.contrib.DeformableConvolution`, call sites models/dcn/builder.py:14-17,
// models/sepc/sepc_dconv.py:12-15, models/tridentnet/resnet_v1.py:85-90) lives in apache/incubator-mxnet
// (src/operator/contrib/deformable_convolution-inl.h + nn/deformable_im2col.cuh, tag 1.6.0 in
// docker/Dockerfile:48) and is NOT in the reference tree: this follows the published DCNv1
// formulation those files implement — PARITY UNPINNED.
//
//   col[(c*KH*KW + tap), h_out, w_out] = bilinear(data[c], p0 + p_tap + offset[g, tap, h_out, w_out])
//   with zero outside the map and the "clamp at the last row/column" rule of deformable_im2col_bilinear.
//
// The GEMM col x weight is a plain dense contraction and goes to cuBLAS through torch (tensor-core
// library work, as BASELINE.json prescribes); these kernels are the HBM-bound part:
// algorithmic bytes = sz(data) + sz(offset) + sz(col).
//
// Thread = (channel, output pixel); it walks the KH*KW taps.  Consecutive threads are consecutive
// output columns: offset reads and col writes are coalesced per tap, the four bilinear corners of
// neighbouring threads fall into the same few lines of the channel plane (L1/L2 hits).
#include "common.cuh"

namespace {

struct DcnShape {
  int C, H, W, KH, KW, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, dg, Ho, Wo;
};

// deformable_im2col_bilinear: data points at (h_in, w_in); height/width are the REMAINING extents
__device__ __forceinline__ float dcn_bilinear(const float* __restrict__ data, const int data_width,
                                              const int height, const int width, float h, float w) {
  int h_low = (int)floorf(h), w_low = (int)floorf(w), h_high, w_high;
  if (h_low >= height - 1) {
    h_high = h_low = height - 1;
    h = (float)h_low;
  } else {
    h_high = h_low + 1;
  }
  if (w_low >= width - 1) {
    w_high = w_low = width - 1;
    w = (float)w_low;
  } else {
    w_high = w_low + 1;
  }
  const float lh = h - (float)h_low, lw = w - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
  const float v1 = __ldg(data + h_low * data_width + w_low), v2 = __ldg(data + h_low * data_width + w_high);
  const float v3 = __ldg(data + h_high * data_width + w_low), v4 = __ldg(data + h_high * data_width + w_high);
  return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4;
}

__global__ void __launch_bounds__(256)
deform_im2col_kernel(const float* __restrict__ data, const float* __restrict__ offset, float* __restrict__ col,
                     const DcnShape s, const int B) {
  const int HWo = s.Ho * s.Wo;
  const size_t total = (size_t)B * s.C * HWo;
  const int cpg = s.C / s.dg;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w_col = (int)(idx % s.Wo), h_col = (int)((idx / s.Wo) % s.Ho);
    const int c = (int)((idx / HWo) % s.C), b = (int)(idx / ((size_t)HWo * s.C));
    const int g = c / cpg;
    const int h_in = h_col * s.stride_h - s.pad_h, w_in = w_col * s.stride_w - s.pad_w;
    const float* im = data + ((size_t)b * s.C + c) * s.H * s.W;
    const float* off = offset + ((size_t)b * s.dg + g) * 2 * s.KH * s.KW * HWo + h_col * s.Wo + w_col;
    float* out = col + (((size_t)b * s.C + c) * s.KH * s.KW) * HWo + h_col * s.Wo + w_col;
    for (int i = 0; i < s.KH; ++i)
      for (int j = 0; j < s.KW; ++j) {
        const int t = i * s.KW + j;
        const float oh = __ldg(off + (size_t)(2 * t) * HWo), ow = __ldg(off + (size_t)(2 * t + 1) * HWo);
        const float h_im = (float)(h_in + i * s.dil_h) + oh, w_im = (float)(w_in + j * s.dil_w) + ow;
        float v = 0.f;
        if (h_im >= 0.f && w_im >= 0.f && h_im < (float)s.H && w_im < (float)s.W)
          v = dcn_bilinear(im, s.W, s.H, s.W, h_im, w_im);
        out[(size_t)t * HWo] = v;
      }
  }
}

// Backward of the gather: thread = one col element; scatters to data grad (4 atomics) and
// accumulates the offset gradient of its (group, tap, pixel) (2 atomics: channels of a group share it).
__global__ void __launch_bounds__(256)
deform_col2im_kernel(const float* __restrict__ gcol, const float* __restrict__ data,
                     const float* __restrict__ offset, float* __restrict__ gdata, float* __restrict__ goffset,
                     const DcnShape s, const int B) {
  const int HWo = s.Ho * s.Wo, T = s.KH * s.KW;
  const size_t total = (size_t)B * s.C * T * HWo;
  const int cpg = s.C / s.dg;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w_col = (int)(idx % s.Wo), h_col = (int)((idx / s.Wo) % s.Ho);
    const int t = (int)((idx / HWo) % T), c = (int)((idx / ((size_t)HWo * T)) % s.C);
    const int b = (int)(idx / ((size_t)HWo * T * s.C));
    const int i = t / s.KW, j = t % s.KW, g = c / cpg;
    const float* off = offset + ((size_t)b * s.dg + g) * 2 * T * HWo + h_col * s.Wo + w_col;
    const float oh = __ldg(off + (size_t)(2 * t) * HWo), ow = __ldg(off + (size_t)(2 * t + 1) * HWo);
    float h = (float)(h_col * s.stride_h - s.pad_h + i * s.dil_h) + oh;
    float w = (float)(w_col * s.stride_w - s.pad_w + j * s.dil_w) + ow;
    if (!(h >= 0.f && w >= 0.f && h < (float)s.H && w < (float)s.W)) continue;
    const float go = __ldg(gcol + idx);
    int h_low = (int)floorf(h), w_low = (int)floorf(w), h_high, w_high;
    bool hc = false, wc = false;  // clamped: the coordinate is a constant there (zero offset gradient)
    if (h_low >= s.H - 1) { h_high = h_low = s.H - 1; h = (float)h_low; hc = true; } else h_high = h_low + 1;
    if (w_low >= s.W - 1) { w_high = w_low = s.W - 1; w = (float)w_low; wc = true; } else w_high = w_low + 1;
    const float lh = h - (float)h_low, lw = w - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
    const size_t plane = ((size_t)b * s.C + c) * s.H * s.W;
    if (gdata) {
      float* gd = gdata + plane;
      atomicAdd(gd + h_low * s.W + w_low, go * hh * hw);
      atomicAdd(gd + h_low * s.W + w_high, go * hh * lw);
      atomicAdd(gd + h_high * s.W + w_low, go * lh * hw);
      atomicAdd(gd + h_high * s.W + w_high, go * lh * lw);
    }
    if (goffset) {
      const float* im = data + plane;
      const float v1 = __ldg(im + h_low * s.W + w_low), v2 = __ldg(im + h_low * s.W + w_high);
      const float v3 = __ldg(im + h_high * s.W + w_low), v4 = __ldg(im + h_high * s.W + w_high);
      float* go_ = goffset + ((size_t)b * s.dg + g) * 2 * T * HWo + h_col * s.Wo + w_col;
      if (!hc) atomicAdd(go_ + (size_t)(2 * t) * HWo, go * (hw * (v3 - v1) + lw * (v4 - v2)));
      if (!wc) atomicAdd(go_ + (size_t)(2 * t + 1) * HWo, go * (hh * (v2 - v1) + lh * (v4 - v3)));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// DCNv2 (modulated) sampling — `mx.sym.contrib.ModulatedDeformableConvolution` of upstream MXNet
// (modulated_deformable_im2col.cuh); BASELINE.json's north star names "v1/v2 sampling", the reference
// tree itself only calls v1.  Restated from the published formulation — PARITY UNPINNED.
// Differences from v1: a sample is taken when -1 < h < H and -1 < w < W, corners outside the map count as
// zeros (no clamping), and every tap is multiplied by mask[g, tap, h_out, w_out].
// ------------------------------------------------------------------------------------------------
struct Corner4 {
  int h_low, w_low, h_high, w_high;
  float lh, lw, hh, hw;
  bool ok1, ok2, ok3, ok4;
};
__device__ __forceinline__ Corner4 dcn2_corners(const int H, const int W, const float h, const float w) {
  Corner4 c;
  c.h_low = (int)floorf(h); c.w_low = (int)floorf(w);
  c.h_high = c.h_low + 1; c.w_high = c.w_low + 1;
  c.lh = h - (float)c.h_low; c.lw = w - (float)c.w_low; c.hh = 1.f - c.lh; c.hw = 1.f - c.lw;
  c.ok1 = c.h_low >= 0 && c.w_low >= 0;
  c.ok2 = c.h_low >= 0 && c.w_high <= W - 1;
  c.ok3 = c.h_high <= H - 1 && c.w_low >= 0;
  c.ok4 = c.h_high <= H - 1 && c.w_high <= W - 1;
  return c;
}

__global__ void __launch_bounds__(256)
mdeform_im2col_kernel(const float* __restrict__ data, const float* __restrict__ offset,
                      const float* __restrict__ mask, float* __restrict__ col, const DcnShape s, const int B) {
  const int HWo = s.Ho * s.Wo, T = s.KH * s.KW;
  const size_t total = (size_t)B * s.C * HWo;
  const int cpg = s.C / s.dg;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w_col = (int)(idx % s.Wo), h_col = (int)((idx / s.Wo) % s.Ho);
    const int c = (int)((idx / HWo) % s.C), b = (int)(idx / ((size_t)HWo * s.C));
    const int g = c / cpg, pix = h_col * s.Wo + w_col;
    const int h_in = h_col * s.stride_h - s.pad_h, w_in = w_col * s.stride_w - s.pad_w;
    const float* im = data + ((size_t)b * s.C + c) * s.H * s.W;
    const float* off = offset + ((size_t)b * s.dg + g) * 2 * T * HWo + pix;
    const float* mk = mask + ((size_t)b * s.dg + g) * T * HWo + pix;
    float* out = col + (((size_t)b * s.C + c) * T) * HWo + pix;
    for (int t = 0; t < T; ++t) {
      const int i = t / s.KW, j = t - i * s.KW;
      const float h_im = (float)(h_in + i * s.dil_h) + __ldg(off + (size_t)(2 * t) * HWo);
      const float w_im = (float)(w_in + j * s.dil_w) + __ldg(off + (size_t)(2 * t + 1) * HWo);
      float v = 0.f;
      if (h_im > -1.f && w_im > -1.f && h_im < (float)s.H && w_im < (float)s.W) {
        const Corner4 k = dcn2_corners(s.H, s.W, h_im, w_im);
        const float v1 = k.ok1 ? __ldg(im + k.h_low * s.W + k.w_low) : 0.f;
        const float v2 = k.ok2 ? __ldg(im + k.h_low * s.W + k.w_high) : 0.f;
        const float v3 = k.ok3 ? __ldg(im + k.h_high * s.W + k.w_low) : 0.f;
        const float v4 = k.ok4 ? __ldg(im + k.h_high * s.W + k.w_high) : 0.f;
        v = k.hh * k.hw * v1 + k.hh * k.lw * v2 + k.lh * k.hw * v3 + k.lh * k.lw * v4;
      }
      out[(size_t)t * HWo] = v * __ldg(mk + (size_t)t * HWo);
    }
  }
}

__global__ void __launch_bounds__(256)
mdeform_col2im_kernel(const float* __restrict__ gcol, const float* __restrict__ data,
                      const float* __restrict__ offset, const float* __restrict__ mask, float* __restrict__ gdata,
                      float* __restrict__ goffset, float* __restrict__ gmask, const DcnShape s, const int B) {
  const int HWo = s.Ho * s.Wo, T = s.KH * s.KW;
  const size_t total = (size_t)B * s.C * T * HWo;
  const int cpg = s.C / s.dg;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w_col = (int)(idx % s.Wo), h_col = (int)((idx / s.Wo) % s.Ho);
    const int t = (int)((idx / HWo) % T), c = (int)((idx / ((size_t)HWo * T)) % s.C);
    const int b = (int)(idx / ((size_t)HWo * T * s.C));
    const int i = t / s.KW, j = t % s.KW, g = c / cpg, pix = h_col * s.Wo + w_col;
    const size_t obase = ((size_t)b * s.dg + g) * 2 * T * HWo + pix;
    const size_t mbase = (((size_t)b * s.dg + g) * T + t) * HWo + pix;
    const float h = (float)(h_col * s.stride_h - s.pad_h + i * s.dil_h) + __ldg(offset + obase + (size_t)(2 * t) * HWo);
    const float w = (float)(w_col * s.stride_w - s.pad_w + j * s.dil_w) + __ldg(offset + obase + (size_t)(2 * t + 1) * HWo);
    if (!(h > -1.f && w > -1.f && h < (float)s.H && w < (float)s.W)) continue;  // sample is the constant 0
    const float go = __ldg(gcol + idx), m = __ldg(mask + mbase);
    const Corner4 k = dcn2_corners(s.H, s.W, h, w);
    const size_t plane = ((size_t)b * s.C + c) * s.H * s.W;
    const float* im = data + plane;
    const float v1 = k.ok1 ? __ldg(im + k.h_low * s.W + k.w_low) : 0.f;
    const float v2 = k.ok2 ? __ldg(im + k.h_low * s.W + k.w_high) : 0.f;
    const float v3 = k.ok3 ? __ldg(im + k.h_high * s.W + k.w_low) : 0.f;
    const float v4 = k.ok4 ? __ldg(im + k.h_high * s.W + k.w_high) : 0.f;
    if (gdata) {
      float* gd = gdata + plane;
      const float gm = go * m;
      if (k.ok1) atomicAdd(gd + k.h_low * s.W + k.w_low, gm * k.hh * k.hw);
      if (k.ok2) atomicAdd(gd + k.h_low * s.W + k.w_high, gm * k.hh * k.lw);
      if (k.ok3) atomicAdd(gd + k.h_high * s.W + k.w_low, gm * k.lh * k.hw);
      if (k.ok4) atomicAdd(gd + k.h_high * s.W + k.w_high, gm * k.lh * k.lw);
    }
    if (goffset) {
      const float gm = go * m;
      atomicAdd(goffset + obase + (size_t)(2 * t) * HWo, gm * (k.hw * (v3 - v1) + k.lw * (v4 - v2)));
      atomicAdd(goffset + obase + (size_t)(2 * t + 1) * HWo, gm * (k.hh * (v2 - v1) + k.lh * (v4 - v3)));
    }
    if (gmask) atomicAdd(gmask + mbase, go * (k.hh * k.hw * v1 + k.hh * k.lw * v2 + k.lh * k.hw * v3 + k.lh * k.lw * v4));
  }
}

int fill_shape(DcnShape& s, int C, int H, int W, int KH, int KW, int pad_h, int pad_w, int stride_h, int stride_w,
               int dil_h, int dil_w, int dg) {
  if (C <= 0 || H <= 0 || W <= 0 || KH <= 0 || KW <= 0 || stride_h <= 0 || stride_w <= 0 || dil_h <= 0 ||
      dil_w <= 0 || dg <= 0 || pad_h < 0 || pad_w < 0)
    return sdet::fail(SDET_ERR_INVALID_ARG, "bad deformable convolution geometry");
  if (C % dg) return sdet::fail(SDET_ERR_INVALID_ARG, "channels must be divisible by num_deformable_group");
  s = DcnShape{C, H, W, KH, KW, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, dg,
               (H + 2 * pad_h - (dil_h * (KH - 1) + 1)) / stride_h + 1,
               (W + 2 * pad_w - (dil_w * (KW - 1) + 1)) / stride_w + 1};
  if (s.Ho <= 0 || s.Wo <= 0) return sdet::fail(SDET_ERR_INVALID_ARG, "empty output");
  return SDET_OK;
}

unsigned grid_for(size_t n) {
  size_t b = (n + 255) / 256;
  return (unsigned)(b > 148 * 32 ? 148 * 32 : (b ? b : 1));
}

}  // namespace

extern "C" int sdet_deformable_im2col(const float* data, const float* offset, float* col, int B, int C, int H,
                                      int W, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h,
                                      int stride_w, int dilate_h, int dilate_w, int num_deformable_group,
                                      void* stream) {
  SDET_REQUIRE(data && offset && col && B > 0, "NULL argument");
  DcnShape s;
  if (int rc = fill_shape(s, C, H, W, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilate_h, dilate_w,
                          num_deformable_group))
    return rc;
  const size_t total = (size_t)B * C * s.Ho * s.Wo;
  deform_im2col_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(data, offset, col, s, B);
  SDET_LAUNCH_CHECK("deform_im2col_kernel");
  return SDET_OK;
}

extern "C" int sdet_deformable_col2im(const float* grad_col, const float* data, const float* offset,
                                      float* grad_data, float* grad_offset, int B, int C, int H, int W,
                                      int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h,
                                      int stride_w, int dilate_h, int dilate_w, int num_deformable_group,
                                      void* stream) {
  SDET_REQUIRE(grad_col && data && offset && (grad_data || grad_offset) && B > 0, "NULL argument");
  DcnShape s;
  if (int rc = fill_shape(s, C, H, W, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilate_h, dilate_w,
                          num_deformable_group))
    return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_data) SDET_CUDA(cudaMemsetAsync(grad_data, 0, sizeof(float) * (size_t)B * C * H * W, st));
  if (grad_offset)
    SDET_CUDA(cudaMemsetAsync(grad_offset, 0,
                              sizeof(float) * (size_t)B * num_deformable_group * 2 * kernel_h * kernel_w * s.Ho * s.Wo, st));
  const size_t total = (size_t)B * C * kernel_h * kernel_w * s.Ho * s.Wo;
  deform_col2im_kernel<<<grid_for(total), 256, 0, st>>>(grad_col, data, offset, grad_data, grad_offset, s, B);
  SDET_LAUNCH_CHECK("deform_col2im_kernel");
  return SDET_OK;
}

extern "C" int sdet_modulated_deformable_im2col(const float* data, const float* offset, const float* mask, float* col,
                                                int B, int C, int H, int W, int kernel_h, int kernel_w, int pad_h,
                                                int pad_w, int stride_h, int stride_w, int dilate_h, int dilate_w,
                                                int num_deformable_group, void* stream) {
  SDET_REQUIRE(data && offset && mask && col && B > 0, "NULL argument");
  DcnShape s;
  if (int rc = fill_shape(s, C, H, W, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilate_h, dilate_w,
                          num_deformable_group))
    return rc;
  const size_t total = (size_t)B * C * s.Ho * s.Wo;
  mdeform_im2col_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(data, offset, mask, col, s, B);
  SDET_LAUNCH_CHECK("mdeform_im2col_kernel");
  return SDET_OK;
}

extern "C" int sdet_modulated_deformable_col2im(const float* grad_col, const float* data, const float* offset,
                                                const float* mask, float* grad_data, float* grad_offset,
                                                float* grad_mask, int B, int C, int H, int W, int kernel_h,
                                                int kernel_w, int pad_h, int pad_w, int stride_h, int stride_w,
                                                int dilate_h, int dilate_w, int num_deformable_group, void* stream) {
  SDET_REQUIRE(grad_col && data && offset && mask && (grad_data || grad_offset || grad_mask) && B > 0, "NULL argument");
  DcnShape s;
  if (int rc = fill_shape(s, C, H, W, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilate_h, dilate_w,
                          num_deformable_group))
    return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t taps = (size_t)B * num_deformable_group * kernel_h * kernel_w * s.Ho * s.Wo;
  if (grad_data) SDET_CUDA(cudaMemsetAsync(grad_data, 0, sizeof(float) * (size_t)B * C * H * W, st));
  if (grad_offset) SDET_CUDA(cudaMemsetAsync(grad_offset, 0, sizeof(float) * 2 * taps, st));
  if (grad_mask) SDET_CUDA(cudaMemsetAsync(grad_mask, 0, sizeof(float) * taps, st));
  const size_t total = (size_t)B * C * kernel_h * kernel_w * s.Ho * s.Wo;
  mdeform_col2im_kernel<<<grid_for(total), 256, 0, st>>>(grad_col, data, offset, mask, grad_data, grad_offset, grad_mask, s, B);
  SDET_LAUNCH_CHECK("mdeform_col2im_kernel");
  return SDET_OK;
}
