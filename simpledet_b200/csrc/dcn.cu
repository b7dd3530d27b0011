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

// ------------------------------------------------------------------------------------------------
// Tiled im2col (3x3-class kernels): CTA = (image, deformable group, run of up to 192 consecutive output pixels,
// channel slice).  The offsets of a deformable group are shared by all its channels, so each thread derives the
// sampling geometry of ITS output pixel once - per tap: patch-relative corner index, corner strides, the four
// bilinear weights - and keeps it in registers; the channel loop then is 4 shared-memory reads, 4 products, 3 sums
// and one coalesced store per col element.  The rows of the input a run can reach (its own rows, the kernel extent
// and a margin of kDcnMargin rows for the offsets) are contiguous per channel plane and are staged with 16-byte
// cp.async into a double-buffered shared-memory patch, channel sub-tile s+1 in flight while sub-tile s is gathered.
// Samples displaced beyond the margin read their four corners from global memory instead (flagged per tap).
// Arithmetic is the generic kernel's, operation for operation.
// ------------------------------------------------------------------------------------------------
constexpr int kDcnThreads = 192;
constexpr int kDcnMaxTaps = 9;
constexpr int kDcnMargin = 6;
constexpr int kDcnStageFloats = 12288;  // 48 KB per stage, two stages

__device__ __forceinline__ void dcn_cp16(unsigned dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void dcn_cp4(unsigned dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

template <int kTaps>
__global__ void __launch_bounds__(kDcnThreads)
deform_im2col_tiled_kernel(const float* __restrict__ data, const float* __restrict__ offset, float* __restrict__ col,
                           const DcnShape s, const int B, const int ct, const int ch_slices, const int runs_per_image) {
  extern __shared__ __align__(16) float s_patch[];  // [2][ct][chan_floats]
  const int tid = threadIdx.x;
  const int HWo = s.Ho * s.Wo, cpg = s.C / s.dg;
  int bid = blockIdx.x;
  const int slice = bid % ch_slices; bid /= ch_slices;
  const int run = bid % runs_per_image; bid /= runs_per_image;
  const int g = bid % s.dg, b = bid / s.dg;
  const int p0 = run * kDcnThreads, npx = min(kDcnThreads, HWo - p0);
  const int c_per = (cpg + ch_slices - 1) / ch_slices;
  const int c_begin = g * cpg + slice * c_per, c_end = min((g + 1) * cpg, c_begin + c_per);
  // rows of the input this run can reach
  const int h_lo = p0 / s.Wo, h_hi = (p0 + npx - 1) / s.Wo;
  const int r_lo = max(0, h_lo * s.stride_h - s.pad_h - kDcnMargin);
  const int r_hi = min(s.H, h_hi * s.stride_h - s.pad_h + (s.KH - 1) * s.dil_h + 2 + kDcnMargin);
  const int patch = max(r_hi - r_lo, 0) * s.W;       // floats per channel
  const int chan_floats = (patch + 3 + 4) & ~3;     // + alignment shift, rounded to 16 bytes

  // ---- per-thread sampling geometry (registers)
  float w1[kTaps], w2[kTaps], w3[kTaps], w4[kTaps];
  int idx[kTaps];   // patch-relative index of the top-left corner; -1: sample is zero; -2: corners come from global
  int dxy[kTaps];   // (w_high - w_low) | ((h_high - h_low) * W) << 1   [patch]   or  absolute top-left index [global]
  const bool active = tid < npx;
  const int p = p0 + (active ? tid : 0);
  const int h_col = p / s.Wo, w_col = p - h_col * s.Wo;
  {
    const float* off = offset + ((size_t)b * s.dg + g) * 2 * kTaps * HWo + p;
#pragma unroll
    for (int t = 0; t < kTaps; ++t) {
      const int i = t / s.KW, j = t - i * s.KW;
      const float oh = __ldg(off + (size_t)(2 * t) * HWo), ow = __ldg(off + (size_t)(2 * t + 1) * HWo);
      float h = (float)(h_col * s.stride_h - s.pad_h + i * s.dil_h) + oh;
      float w = (float)(w_col * s.stride_w - s.pad_w + j * s.dil_w) + ow;
      idx[t] = -1;
      dxy[t] = 0;
      w1[t] = w2[t] = w3[t] = w4[t] = 0.f;
      if (active && h >= 0.f && w >= 0.f && h < (float)s.H && w < (float)s.W) {
        int h_low = (int)floorf(h), w_low = (int)floorf(w), h_high, w_high;
        if (h_low >= s.H - 1) { h_high = h_low = s.H - 1; h = (float)h_low; } else h_high = h_low + 1;
        if (w_low >= s.W - 1) { w_high = w_low = s.W - 1; w = (float)w_low; } else w_high = w_low + 1;
        const float lh = h - (float)h_low, lw = w - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
        w1[t] = hh * hw; w2[t] = hh * lw; w3[t] = lh * hw; w4[t] = lh * lw;
        const int d = (w_high - w_low) | (((h_high - h_low) * s.W) << 1);
        if (h_low >= r_lo && h_high < r_hi) {
          idx[t] = (h_low - r_lo) * s.W + w_low;
          dxy[t] = d;
        } else {
          idx[t] = -2;
          dxy[t] = ((h_low * s.W + w_low) << 1) | (w_high - w_low);  // absolute; the row step is re-derived below
          w4[t] = lh * lw;
          // h_high - h_low in bit 30 (indices stay far below 2^29)
          if (h_high != h_low) dxy[t] |= 1 << 30;
        }
      }
    }
  }

  const size_t plane = (size_t)s.H * s.W;
  const float* tensor_begin = data;
  const float* tensor_end = data + (size_t)B * s.C * plane;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(s_patch);
  __shared__ int s_shift[2][32];

  // stage `nc` channels starting at channel c0 into buffer `buf`: the patch of a channel is one contiguous run
  auto stage = [&](int c0, int nc, int buf) {
    for (int k = 0; k < nc; ++k) {
      const float* src = data + ((size_t)b * s.C + c0 + k) * plane + (size_t)r_lo * s.W;
      const int shift = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);  // floats past a 16-byte boundary
      if (tid == 0) s_shift[buf][k] = shift;
      const float* a0 = src - shift;                                         // 16-byte aligned
      const unsigned d0 = sbase + 4u * (unsigned)((buf * ct + k) * chan_floats);
      const int nchunks = (patch + shift + 3) >> 2;
      for (int q = tid; q < nchunks; q += kDcnThreads) {
        const float* a = a0 + 4 * q;
        if (a >= tensor_begin && a + 4 <= tensor_end) {
          dcn_cp16(d0 + 16u * q, a);
        } else {  // first / last chunk of the whole tensor: element-wise, in range only
          for (int e = 0; e < 4; ++e)
            if (a + e >= tensor_begin && a + e < tensor_end) dcn_cp4(d0 + 16u * q + 4u * e, a + e);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int nct = c_end - c_begin;
  const int nst = (nct + ct - 1) / ct;
  if (nst > 0) stage(c_begin, min(ct, nct), 0);
  for (int st = 0; st < nst; ++st) {
    const int buf = st & 1;
    if (st + 1 < nst) {
      stage(c_begin + (st + 1) * ct, min(ct, nct - (st + 1) * ct), buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int c0 = c_begin + st * ct, nc = min(ct, nct - st * ct);
    if (active) {
      for (int k = 0; k < nc; ++k) {
        const float* sp = s_patch + (size_t)(buf * ct + k) * chan_floats + s_shift[buf][k];
        const float* im = data + ((size_t)b * s.C + c0 + k) * plane;
        float* out = col + (((size_t)b * s.C + c0 + k) * kTaps) * HWo + p;
#pragma unroll
        for (int t = 0; t < kTaps; ++t) {
          float v = 0.f;
          if (idx[t] >= 0) {
            const int dx = dxy[t] & 1, dy = dxy[t] >> 1;
            const float* q = sp + idx[t];
            v = w1[t] * q[0] + w2[t] * q[dx] + w3[t] * q[dy] + w4[t] * q[dy + dx];
          } else if (idx[t] == -2) {
            const int dx = dxy[t] & 1, dy = (dxy[t] >> 30) ? s.W : 0;
            const float* q = im + ((dxy[t] & 0x3FFFFFFF) >> 1);
            v = w1[t] * __ldg(q) + w2[t] * __ldg(q + dx) + w3[t] * __ldg(q + dy) + w4[t] * __ldg(q + dy + dx);
          }
          __stcs(out + (size_t)t * HWo, v);
        }
      }
    }
    __syncthreads();
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
  // tiled path: 3x3-class kernels whose reachable input rows fit the staging buffer with >= 4 channels per stage
  const int HWo = s.Ho * s.Wo;
  const int rows_max = ((kDcnThreads + s.Wo - 1) / s.Wo + 1) * stride_h + (kernel_h - 1) * dilate_h + 2 + 2 * kDcnMargin;
  const int chan_floats_max = ((rows_max < H ? rows_max : H) * W + 7) & ~3;
  int ct = kDcnStageFloats / chan_floats_max;
  if (ct > 8) ct = 8;
  if (kernel_h * kernel_w == kDcnMaxTaps && ct >= 4 && (size_t)H * W < (1u << 28)) {
    const int cpg = C / num_deformable_group;
    const int runs = (HWo + kDcnThreads - 1) / kDcnThreads;
    // enough CTAs for ~3 waves: split the channels of a group when the pixel runs alone are too few
    int slices = 1;
    while (slices < 8 && (long long)B * num_deformable_group * runs * slices < 148 * 3 && cpg / (slices * 2) >= ct) slices *= 2;
    const size_t smem = (size_t)2 * ct * chan_floats_max * sizeof(float);
    SDET_CUDA(cudaFuncSetAttribute(deform_im2col_tiled_kernel<kDcnMaxTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)(B * num_deformable_group * runs * slices);
    deform_im2col_tiled_kernel<kDcnMaxTaps><<<grid, kDcnThreads, smem, (cudaStream_t)stream>>>(data, offset, col, s, B, ct,
                                                                                              slices, runs);
    SDET_LAUNCH_CHECK("deform_im2col_tiled_kernel");
    return SDET_OK;
  }
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
