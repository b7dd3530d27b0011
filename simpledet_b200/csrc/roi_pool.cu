// ROIPooling_v1 forward/backward for sm_100a.
// Reference semantics: operator_cxx/roi_pooling_v1.cu:49-113 (forward), :116-152 (backward), CPU twin
// roi_pooling_v1.cc:40-221, op wrapper roi_pooling_v1-inl.h:63-137.  NOT the reference's kernel (one thread
// per output scanning its bin out of global memory, one atomic per output in backward):
//
//   forward   CTA = (roi, channel tile).  The integer bin geometry is computed once per roi into shared tables;
//             the roi's window of the tile's channels is staged row-coalesced in shared memory; the max over a bin
//             is taken SEPARABLY - phase 1: per (channel, window row, pw) the first maximum along w, phase 2: per
//             (channel, ph, pw) the first maximum along h of the phase-1 results - which visits every staged value
//             once per overlapping bin column instead of once per overlapping bin, and picks exactly the reference's
//             winner: its row-major scan with strict `>` keeps the first maximum in (h, w) order, i.e. the first row
//             holding the bin's maximum and the first column inside that row.
//   backward  gather form: thread = (channel, window pixel) sums the gradients of the (at most 2 x 2) bins that
//             contain the pixel and elected it as argmax, in (ph, pw) order, then issues ONE red.global per touched
//             pixel - deterministic inside a roi, no per-bin atomics, untouched pixels cost nothing.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int kRpThreads = 256;
constexpr int kRpMaxP = SDET_MAX_POOLED;
constexpr int kRpSmemFloats = 11264;  // 44 KB of staged window + phase-1 results

struct RoiBins {  // integer geometry of one roi (roi_pooling_v1.cu:66-90)
  int hs[kRpMaxP], he[kRpMaxP], ws[kRpMaxP], we[kRpMaxP];
  int bi, hmin, hmax, wmin, wmax;  // batch index; window = union of the bins, clipped to the map
};

__device__ void roi_bins(const float* __restrict__ r, const float scale, const int H, const int W, const int PH,
                         const int PW, RoiBins& b) {
  // round() = half away from zero, as C's round() in roi_pooling_v1.cu:69-72
  const int rsw = (int)roundf(__fmul_rn(__ldg(r + 1), scale)), rsh = (int)roundf(__fmul_rn(__ldg(r + 2), scale));
  const int rew = (int)roundf(__fmul_rn(__ldg(r + 3), scale)), reh = (int)roundf(__fmul_rn(__ldg(r + 4), scale));
  const int rw = max(rew - rsw + 1, 1), rh = max(reh - rsh + 1, 1);  // malformed rois become 1x1
  const float bh = __fdiv_rn((float)rh, (float)PH), bw = __fdiv_rn((float)rw, (float)PW);
  const int t = threadIdx.x;
  if (t < PH) {
    b.hs[t] = min(max((int)floorf(__fmul_rn((float)t, bh)) + rsh, 0), H);
    b.he[t] = min(max((int)ceilf(__fmul_rn((float)(t + 1), bh)) + rsh, 0), H);
  } else if (t >= 64 && t < 64 + PW) {
    const int p = t - 64;
    b.ws[p] = min(max((int)floorf(__fmul_rn((float)p, bw)) + rsw, 0), W);
    b.we[p] = min(max((int)ceilf(__fmul_rn((float)(p + 1), bw)) + rsw, 0), W);
  }
  __syncthreads();
  if (t == 0) {
    b.bi = (int)__ldg(r);
    // bin starts and ends are monotone in the bin index: the window is [first start, last end)
    b.hmin = b.hs[0]; b.hmax = b.he[PH - 1];
    b.wmin = b.ws[0]; b.wmax = b.we[PW - 1];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kRpThreads)
roi_pool_v1_fwd_kernel(const float* __restrict__ data, const float* __restrict__ rois, float* __restrict__ out,
                       float* __restrict__ max_idx, const int B, const int C, const int H, const int W, const int PH,
                       const int PW, const float scale) {
  __shared__ RoiBins sb;
  __shared__ float s_buf[kRpSmemFloats];
  const int n = blockIdx.x, tid = threadIdx.x;
  roi_bins(rois + (size_t)n * 5, scale, H, W, PH, PW, sb);
  const int PP = PH * PW;
  const int Hw = max(sb.hmax - sb.hmin, 0), Ww = max(sb.wmax - sb.wmin, 0);
  const bool valid_img = sb.bi >= 0 && sb.bi < B;
  const int cells = Hw * Ww, m1 = Hw * PW;           // staged floats / phase-1 entries per channel
  const int per_ch = cells + 2 * m1;                  // + phase-1 value and column
  // channels per pass: as many as fit (at least 1 when the window itself fits)
  int ct = per_ch > 0 ? kRpSmemFloats / per_ch : C;
  ct = min(ct, C);
  const bool direct = ct == 0 || !valid_img || cells == 0;  // huge window / no image / nothing to stage
  for (int c0 = blockIdx.y * (direct ? 8 : ct); c0 < C; c0 += gridDim.y * (direct ? 8 : ct)) {
    const int nc = min(direct ? 8 : ct, C - c0);
    float* o = out + ((size_t)n * C + c0) * PP;
    float* oi = max_idx ? max_idx + ((size_t)n * C + c0) * PP : nullptr;
    if (direct) {
      // window too large for shared memory (or nothing to read): scan from global, same visiting order
      for (int e = tid; e < nc * PP; e += kRpThreads) {
        const int cl = e / PP, ph = (e / PW) % PH, pw = e % PW;
        const int hs = sb.hs[ph], he = sb.he[ph], ws = sb.ws[pw], we = sb.we[pw];
        float best = (he <= hs || we <= ws) ? 0.f : -FLT_MAX;
        int arg = -1;
        if (valid_img) {
          const float* plane = data + ((size_t)sb.bi * C + c0 + cl) * H * W;
          for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w) {
              const float v = __ldg(plane + h * W + w);
              if (v > best) { best = v; arg = h * W + w; }
            }
        }
        o[e] = best;
        if (oi) oi[e] = (float)arg;
      }
      continue;
    }
    float* s_win = s_buf;                       // [nc][Hw][Ww]
    float* s_val = s_buf + (size_t)ct * cells;  // [nc][Hw][PW]
    int* s_col = reinterpret_cast<int*>(s_val + (size_t)ct * m1);
    __syncthreads();  // previous pass is done with the buffers
    // ---- stage: rows of the window are contiguous in global memory
    const float* g0 = data + ((size_t)sb.bi * C + c0) * H * W + (size_t)sb.hmin * W + sb.wmin;
    for (int e = tid; e < nc * cells; e += kRpThreads) {
      const int cl = e / cells, rem = e - cl * cells, y = rem / Ww, x = rem - y * Ww;
      s_win[e] = __ldg(g0 + (size_t)cl * H * W + (size_t)y * W + x);
    }
    __syncthreads();
    // ---- phase 1: first maximum along w of every (channel, window row, pw)
    for (int e = tid; e < nc * m1; e += kRpThreads) {
      const int cl = e / m1, rem = e - cl * m1, y = rem / PW, pw = rem - y * PW;
      const float* row = s_win + (size_t)cl * cells + (size_t)y * Ww - sb.wmin;
      float best = -FLT_MAX;
      int col = -1;
      for (int w = sb.ws[pw]; w < sb.we[pw]; ++w) {
        const float v = row[w];
        if (v > best) { best = v; col = w; }
      }
      s_val[e] = best;
      s_col[e] = col;
    }
    __syncthreads();
    // ---- phase 2: first maximum along h; (channel, ph, pw) is contiguous in the output
    for (int e = tid; e < nc * PP; e += kRpThreads) {
      const int cl = e / PP, ph = (e / PW) % PH, pw = e % PW;
      const int hs = sb.hs[ph], he = sb.he[ph];
      const bool empty = (he <= hs) || (sb.we[pw] <= sb.ws[pw]);
      float best = empty ? 0.f : -FLT_MAX;
      int arg = -1;
      const float* v = s_val + (size_t)cl * m1 + pw;
      const int* cidx = s_col + (size_t)cl * m1 + pw;
      for (int h = hs; h < he; ++h) {
        const float x = v[(h - sb.hmin) * PW];
        if (x > best) { best = x; arg = h * W + cidx[(h - sb.hmin) * PW]; }
      }
      o[e] = best;
      if (oi) oi[e] = (float)arg;
    }
  }
}

__global__ void __launch_bounds__(kRpThreads)
roi_pool_v1_bwd_kernel(const float* __restrict__ ograd, const float* __restrict__ max_idx,
                       const float* __restrict__ rois, float* __restrict__ grad, const int B, const int C,
                       const int H, const int W, const int PH, const int PW, const float scale) {
  __shared__ RoiBins sb;
  const int n = blockIdx.x, tid = threadIdx.x;
  roi_bins(rois + (size_t)n * 5, scale, H, W, PH, PW, sb);
  if (sb.bi < 0 || sb.bi >= B) return;
  const int PP = PH * PW;
  const int Hw = max(sb.hmax - sb.hmin, 0), Ww = max(sb.wmax - sb.wmin, 0);
  const int cells = Hw * Ww;
  if (cells == 0) return;
  const int c_per = (C + gridDim.y - 1) / gridDim.y;
  const int c0 = blockIdx.y * c_per, c1 = min(C, c0 + c_per);
  for (long long e = tid; e < (long long)(c1 - c0) * cells; e += kRpThreads) {
    const int cl = (int)(e / cells), rem = (int)(e - (long long)cl * cells), y = rem / Ww, x = rem - y * Ww;
    const int h = sb.hmin + y, w = sb.wmin + x, c = c0 + cl;
    const float target = (float)(h * W + w);
    const float* mi = max_idx + ((size_t)n * C + c) * PP;
    const float* og = ograd + ((size_t)n * C + c) * PP;
    float acc = 0.f;
    bool any = false;
    // the bins containing (h, w): consecutive bins overlap by at most one cell, so very few pass the range tests
    for (int ph = 0; ph < PH; ++ph) {
      if (h < sb.hs[ph] || h >= sb.he[ph]) continue;
      for (int pw = 0; pw < PW; ++pw) {
        if (w < sb.ws[pw] || w >= sb.we[pw]) continue;
        if (__ldg(mi + ph * PW + pw) == target) {
          acc = __fadd_rn(acc, __ldg(og + ph * PW + pw));
          any = true;
        }
      }
    }
    if (any) atomicAdd(grad + ((size_t)sb.bi * C + c) * H * W + (size_t)h * W + w, acc);
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
  if (pooled_h > kRpMaxP || pooled_w > kRpMaxP)
    return sdet::fail(SDET_ERR_UNSUPPORTED, "pooled_size > %d per axis is not supported", kRpMaxP);
  // channel slices per roi: enough CTAs to fill the machine a few times over
  int ysplit = (int)((148 * 8 + R - 1) / R);
  if (ysplit < 1) ysplit = 1;
  if (ysplit > C) ysplit = C;
  dim3 grid((unsigned)R, (unsigned)ysplit);
  roi_pool_v1_fwd_kernel<<<grid, kRpThreads, 0, (cudaStream_t)stream>>>(data, rois, out, max_idx, B, C, H, W,
                                                                       pooled_h, pooled_w, spatial_scale);
  SDET_LAUNCH_CHECK("roi_pool_v1_fwd_kernel");
  return SDET_OK;
}

extern "C" int sdet_roi_pooling_v1_backward(const float* ograd, const float* max_idx,
                                               const float* rois, float* grad_data, float* grad_rois,
                                               int B, int R, int C, int H, int W, int pooled_h,
                                               int pooled_w, float spatial_scale, int accumulate, void* stream) {
  SDET_REQUIRE(ograd && max_idx && rois && grad_data, "NULL argument");
  SDET_REQUIRE(B > 0 && R > 0 && C > 0 && H > 0 && W > 0 && pooled_h > 0 && pooled_w > 0,
               "shape must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate)  // grad_in = 0.0f on kWriteTo (roi_pooling_v1-inl.h:125-127)
    SDET_CUDA(cudaMemsetAsync(grad_data, 0, sizeof(float) * (size_t)B * C * H * W, st));
  if (grad_rois) SDET_CUDA(cudaMemsetAsync(grad_rois, 0, sizeof(float) * (size_t)R * 5, st));
  if (pooled_h > kRpMaxP || pooled_w > kRpMaxP)
    return sdet::fail(SDET_ERR_UNSUPPORTED, "pooled_size > %d per axis is not supported", kRpMaxP);
  int ysplit = (int)((148 * 8 + R - 1) / R);
  if (ysplit < 1) ysplit = 1;
  if (ysplit > C) ysplit = C;
  dim3 grid((unsigned)R, (unsigned)ysplit);
  roi_pool_v1_bwd_kernel<<<grid, kRpThreads, 0, st>>>(ograd, max_idx, rois, grad_data, B, C, H, W, pooled_h,
                                                      pooled_w, spatial_scale);
  SDET_LAUNCH_CHECK("roi_pool_v1_bwd_kernel");
  return SDET_OK;
}
