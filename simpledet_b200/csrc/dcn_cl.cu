// Channels-last sampling for _contrib_DeformableConvolution (DCNv1) on sm_100a.
//
// Semantics: upstream MXNet deformable_im2col (apache/incubator-mxnet src/operator/contrib/nn/deformable_im2col.cuh,
// call sites models/dcn/builder.py:14-17) - the same arithmetic, operation for operation, as deform_im2col_kernel
// in dcn.cu.  What changes is the layout, for the same reason as in roi_align_cl.cu: the sampling geometry of an
// (output pixel, tap) is shared by all C / num_deformable_group channels of a deformable group, so with the input in
// NHWC
//
//   a warp owns 32 consecutive output pixels of one (image, deformable group, tap): every lane derives the geometry
//   of ONE of them (coalesced offset reads, the floor / clamp / weight arithmetic once per sample instead of once
//   per channel), then the warp walks the 32 samples, the geometry broadcast by shuffles, lane = 2 channels:
//   four coalesced 256-byte corner loads, 4 + 3 packed fp32x2 operations, one coalesced 256-byte store.
//
// The columns come out as col_t (B, Ho*Wo, KH*KW, C) - the K index of the following GEMM is (tap, channel) - so the
// dense product is out(B, Ho*Wo, F) = col_t x W', W'[(tap, c)][f] = weight[f][c][tap]: channels-last in, channels-
// last out, no transposition anywhere.  ~0.4 warp instructions per column element instead of 1.9 (profiles/r02_dcn.md).
#include "roi_align_common.cuh"  // packed fp32x2 helpers

using sdet_ra::add2;
using sdet_ra::fma2;
using sdet_ra::pack2;

namespace {

struct DcnClShape {
  int C, H, W, KH, KW, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, dg, Ho, Wo;
};

__device__ __forceinline__ uint64_t ldg2_cl(const float* p) {
  uint64_t v;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(*reinterpret_cast<float*>(&v)),
               "=f"(*(reinterpret_cast<float*>(&v) + 1)) : "l"(p));
  return v;
}

constexpr int kWarpsPerCta = 8;

template <bool kOnePass>  // kOnePass: C / num_deformable_group <= 64, a lane's two channels are the whole job
__global__ void __launch_bounds__(kWarpsPerCta * 32)
deform_im2col_cl_kernel(const float* __restrict__ data, const float* __restrict__ offset, float* __restrict__ col_t,
                        const DcnClShape s, const int B, const int blocks_per_plane, const long long nchunks,
                        const uint64_t nz2) {
  const int lane = threadIdx.x & 31;
  const int HWo = s.Ho * s.Wo, T = s.KH * s.KW, cpg = s.C / s.dg;
  // nz2 = {-0.0f, -0.0f}: fma2(w, v, -0.0) is the correctly rounded product.  It arrives as a kernel argument: as a
  // literal ptxas folds the fma into a multiply and then contracts multiply + add into FFMA2 (different rounding).
  // chunk = ((b * dg + g) * blocks + blk) * T + t: the taps of one block of pixels are neighbours (they read the
  // same input pixels) and land in the same CTA
  for (long long chunk = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); chunk < nchunks;
       chunk += (long long)gridDim.x * kWarpsPerCta) {
    const int t = (int)(chunk % T);
    long long r = chunk / T;
    const int blk = (int)(r % blocks_per_plane);
    r /= blocks_per_plane;
    const int g = (int)(r % s.dg), b = (int)(r / s.dg);
    const int p0 = blk * 32, p = p0 + lane;
    // ---- geometry of this lane's sample (deformable_im2col.cuh: deformable_im2col_gpu_kernel + _bilinear)
    float w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
    int o1 = -1, code = 0, e2 = 0, e3 = 0, e4 = 0;  // o1 < 0: outside the image, the column is zero
    if (p < HWo) {
      const int h_col = p / s.Wo, w_col = p - h_col * s.Wo;
      const int i = t / s.KW, j = t - i * s.KW;
      const float* off = offset + (((size_t)b * s.dg + g) * 2 * T + 2 * t) * HWo + p;
      const float oh = __ldg(off), ow = __ldg(off + HWo);
      float h_im = (float)(h_col * s.stride_h - s.pad_h + i * s.dil_h) + oh;
      float w_im = (float)(w_col * s.stride_w - s.pad_w + j * s.dil_w) + ow;
      if (h_im >= 0.f && w_im >= 0.f && h_im < (float)s.H && w_im < (float)s.W) {
        int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im), dy = 1, dx = 1;
        if (h_low >= s.H - 1) { h_low = s.H - 1; h_im = (float)h_low; dy = 0; }
        if (w_low >= s.W - 1) { w_low = s.W - 1; w_im = (float)w_low; dx = 0; }
        const float lh = h_im - (float)h_low, lw = w_im - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
        w1 = hh * hw; w2 = hh * lw; w3 = lh * hw; w4 = lh * lw;
        o1 = h_low * s.W + w_low;
        code = dx | (dy << 1);
        if (kOnePass) {  // element offsets of the four corners instead of the pixel index
          o1 *= s.C;
          e2 = o1 + dx * s.C;
          e3 = o1 + dy * s.W * s.C;
          e4 = e3 + dx * s.C;
        }
      }
    }
    const float* img = data + (size_t)b * s.H * s.W * s.C + g * cpg + 2 * lane;
    float* o = col_t + (((size_t)b * HWo + p0) * T + t) * s.C + g * cpg + 2 * lane;
    const int ostep = T * s.C;         // elements between consecutive pixels of col_t
    const int C = s.C, rstep = s.W * s.C;
    const int n = min(32, HWo - p0);
    for (int u = 0; u < n; ++u, o += ostep) {
      const int uo = __shfl_sync(0xffffffffu, o1, u), uc = __shfl_sync(0xffffffffu, code, u);
      const float a1 = __shfl_sync(0xffffffffu, w1, u), a2 = __shfl_sync(0xffffffffu, w2, u);
      const float a3 = __shfl_sync(0xffffffffu, w3, u), a4 = __shfl_sync(0xffffffffu, w4, u);
      if (kOnePass) {
        // the four corner offsets (elements, < 2^31: checked by the host) were formed by the lane that owns the sample
        const int f2 = __shfl_sync(0xffffffffu, e2, u), f3 = __shfl_sync(0xffffffffu, e3, u);
        const int f4 = __shfl_sync(0xffffffffu, e4, u);
        if (2 * lane < cpg) {
          uint64_t res = 0;  // outside the image: zeros
          if (uo >= 0) {
            const uint64_t v1 = ldg2_cl(img + uo), v2 = ldg2_cl(img + f2), v3 = ldg2_cl(img + f3), v4 = ldg2_cl(img + f4);
            // hh*hw*v1 + hh*lw*v2 + lh*hw*v3 + lh*lw*v4, left to right, every product and sum rounded on its own
            const uint64_t s12 = add2(fma2(pack2(a1, a1), v1, nz2), fma2(pack2(a2, a2), v2, nz2));
            res = add2(add2(s12, fma2(pack2(a3, a3), v3, nz2)), fma2(pack2(a4, a4), v4, nz2));
          }
          __stcs(reinterpret_cast<float2*>(o), *reinterpret_cast<const float2*>(&res));
        }
        continue;
      }
      if (uo < 0) {  // outside the image: zeros
        for (int c = 2 * lane; c < cpg; c += 64) __stcs(reinterpret_cast<float2*>(o + (c - 2 * lane)), make_float2(0.f, 0.f));
        continue;
      }
      const float* q1 = img + uo * C;
      const int dxo = (uc & 1) ? C : 0, dyo = (uc & 2) ? rstep : 0;
#pragma unroll 1
      for (int c = 2 * lane, k = 0; c < cpg; c += 64, k += 64) {
        const float* q = q1 + k;
        const uint64_t v1 = ldg2_cl(q), v2 = ldg2_cl(q + dxo), v3 = ldg2_cl(q + dyo), v4 = ldg2_cl(q + dyo + dxo);
        const uint64_t s12 = add2(fma2(pack2(a1, a1), v1, nz2), fma2(pack2(a2, a2), v2, nz2));
        const uint64_t s123 = add2(s12, fma2(pack2(a3, a3), v3, nz2));
        const uint64_t res = add2(s123, fma2(pack2(a4, a4), v4, nz2));
        __stcs(reinterpret_cast<float2*>(o + k), *reinterpret_cast<const float2*>(&res));
      }
    }
  }
}

}  // namespace

// data (B,H,W,C) channels-last, offset (B, dg*2*KH*KW, Ho, Wo) as the operator defines it, col_t (B, Ho*Wo, KH*KW, C).
// C / num_deformable_group must be even (a lane carries two channels).
extern "C" int sdet_deformable_im2col_nhwc(const float* data, const float* offset, float* col_t, int B, int C, int H,
                                           int W, int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h,
                                           int stride_w, int dilate_h, int dilate_w, int num_deformable_group,
                                           void* stream) {
  SDET_REQUIRE(data && offset && col_t && B > 0, "NULL argument");
  if (C <= 0 || H <= 0 || W <= 0 || kernel_h <= 0 || kernel_w <= 0 || stride_h <= 0 || stride_w <= 0 || dilate_h <= 0 ||
      dilate_w <= 0 || num_deformable_group <= 0 || pad_h < 0 || pad_w < 0)
    return sdet::fail(SDET_ERR_INVALID_ARG, "bad deformable convolution geometry");
  if (C % num_deformable_group) return sdet::fail(SDET_ERR_INVALID_ARG, "channels must be divisible by num_deformable_group");
  if ((C / num_deformable_group) & 1)
    return sdet::fail(SDET_ERR_UNSUPPORTED, "channels-last sampling needs an even channel count per deformable group");
  DcnClShape s{C, H, W, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, dilate_h, dilate_w, num_deformable_group,
               (H + 2 * pad_h - (dilate_h * (kernel_h - 1) + 1)) / stride_h + 1,
               (W + 2 * pad_w - (dilate_w * (kernel_w - 1) + 1)) / stride_w + 1};
  if (s.Ho <= 0 || s.Wo <= 0) return sdet::fail(SDET_ERR_INVALID_ARG, "empty output");
  if ((size_t)H * W * C > 0x7FFFFFFFull) return sdet::fail(SDET_ERR_UNSUPPORTED, "image plane of more than 2^31 elements");
  const int blocks = (s.Ho * s.Wo + 31) / 32;
  const long long nchunks = (long long)B * num_deformable_group * blocks * kernel_h * kernel_w;
  const long long ctas = (nchunks + kWarpsPerCta - 1) / kWarpsPerCta;
  const unsigned grid = (unsigned)(ctas < 148 * 64 ? ctas : 148 * 64);
  const uint64_t nz2 = 0x8000000080000000ull;
  if (C / num_deformable_group <= 64)
    deform_im2col_cl_kernel<true><<<grid, kWarpsPerCta * 32, 0, (cudaStream_t)stream>>>(data, offset, col_t, s, B, blocks,
                                                                                         nchunks, nz2);
  else
    deform_im2col_cl_kernel<false><<<grid, kWarpsPerCta * 32, 0, (cudaStream_t)stream>>>(data, offset, col_t, s, B, blocks,
                                                                                          nchunks, nz2);
  SDET_LAUNCH_CHECK("deform_im2col_cl_kernel");
  return SDET_OK;
}
