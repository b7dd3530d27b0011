// _contrib_ROIAlign_v2 forward/backward and the fused FPN variant, hand-written for sm_100a.
//
// Reference semantics: operator_cxx/contrib/roi_align_v2-inl.h:61-153 (forward functor),
// operator_cxx/contrib/roi_align_v2.cu:17-85 (backward), models/FPN/assign_layer_fpn.py:17-40
// (level assignment).  NOT a port: the reference runs one thread per output element with 16
// scattered global loads; here one CTA owns (roi, channel tile):
//
//   1. preamble  — PH+PW threads restate the reference's float/double sample loop ONCE per roi
//                  (the sample coordinates depend only on (roi, ph) / (roi, pw), not on the
//                  channel) into shared-memory axis tables: coordinate, lo/hi pixel, weights.
//   2. stage     — the roi's feature window [CT, Hwin, Wwin] is copied NCHW-row-coalesced into
//                  shared memory (each global byte of the window is read once per CTA).
//   3. compute   — thread = (pw, channel group, ph chunk) walks down the rows with a 2-row
//                  register cache, so a window element is read from shared memory ~once per
//                  column tap instead of once per sample; bilinear weights are shared across
//                  the CPT channels a thread owns.
//
// Arithmetic is bit-identical to the reference's CPU build: every float op on the coordinate
// and value path is an explicit round-to-nearest intrinsic (no FMA contraction), the
// `(hend-hstart)/3.0` and `<= hend-h_stride+0.01` promotions to double are kept.
#include <cfloat>
#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int kMaxS = 4;                 // samples per bin per axis kept in the tables
constexpr int kMaxP = SDET_MAX_POOLED;   // pooled size limit per axis
constexpr int kTab = kMaxP * kMaxS;
constexpr int kFlagNot2 = 1;             // some non-empty bin does not have exactly 2 samples
constexpr int kFlagOverflow = 2;         // some bin has more than kMaxS samples

struct Level {
  const float* data;
  float* grad;
  int H, W;
  float scale;
  int stride_log2;
};

struct RoiAlignArgs {
  Level lvl[SDET_MAX_LEVELS];
  int num_levels;
  int fpn;  // 0: every roi is sampled on lvl[0] (plain _contrib_ROIAlign_v2)
  float scale0, lvl0, k_min, k_max;
  const float* rois;
  float* out;
  float* argx;
  float* argy;
  int32_t* levels_out;
  int B, N, C, PH, PW;
};

struct AxisTab {
  float coord[kTab];
  float w0[kTab];  // 1 - alpha
  float w1[kTab];  // alpha
  int lo[kTab];
  int hi[kTab];
  int cnt[kMaxP];  // -1: bin empty along this axis (end <= start); else #samples (may be 0)
};

__device__ __forceinline__ float min_ref(float a, float b) { return a < b ? a : b; }  // mshadow_op::minimum
__device__ __forceinline__ float max_ref(float a, float b) { return a > b ? a : b; }  // mshadow_op::maximum

// models/FPN/assign_layer_fpn.py:27-33 in float32.  Returns floor-level or INT_MIN for NaN.
__device__ __forceinline__ int fpn_level(float x1, float y1, float x2, float y2, float scale0,
                                         float lvl0, float k_min, float k_max) {
  float area = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.f), __fadd_rn(__fsub_rn(y2, y1), 1.f));
  float sc = __fsqrt_rn(area);
  float t = floorf(__fadd_rn(lvl0, log2f(__fadd_rn(__fdiv_rn(sc, scale0), 1e-6f))));
  t = min_ref(max_ref(t, k_min), k_max);
  return (t != t) ? INT_MIN : (int)t;
}

// roi_align_v2-inl.h:91-125 for one bin of one axis.  Thread-private; writes the axis table.
__device__ void build_axis_bin(AxisTab& t, int p, int P, float roi_start, float roi_end,
                               int extent, int* s_flags, int* s_min, int* s_max) {
  const float size = __fsub_rn(roi_end, roi_start);
  const float bin = __fdiv_rn(size, (float)P);
  const float lim = (float)(extent - 1);
  float s = __fmul_rn((float)p, bin);
  float e = __fmul_rn((float)(p + 1), bin);
  s = min_ref(max_ref(__fadd_rn(s, roi_start), 0.f), lim);
  e = min_ref(max_ref(__fadd_rn(e, roi_start), 0.f), lim);
  if (e <= s) {
    t.cnt[p] = -1;
    return;
  }
  const float stride = (float)__ddiv_rn((double)__fsub_rn(e, s), 3.0);
  const float step = max_ref(stride, 0.01f);
  const double last = __dadd_rn((double)__fsub_rn(e, stride), 0.01);
  int n = 0, mn = INT_MAX, mx = -1;
  for (float h = __fadd_rn(s, stride); (double)h <= last; h = __fadd_rn(h, step)) {
    if (n < kMaxS) {
      int lo = min(max((int)floorf(h), 0), extent - 1);
      int hi = min(max((int)ceilf(h), 0), extent - 1);
      float alpha = (lo == hi) ? 0.5f : __fdiv_rn(__fsub_rn(h, (float)lo), (float)(hi - lo));
      const int j = p * kMaxS + n;
      t.coord[j] = h;
      t.lo[j] = lo;
      t.hi[j] = hi;
      t.w0[j] = __fsub_rn(1.f, alpha);
      t.w1[j] = alpha;
      mn = min(mn, lo);
      mx = max(mx, hi);
    }
    if (++n > 4096) break;  // cannot happen for extents < 2^17; keeps a corrupt roi from hanging
  }
  t.cnt[p] = n;
  if (n != 2) atomicOr(s_flags, kFlagNot2);
  if (n > kMaxS) atomicOr(s_flags, kFlagOverflow);
  if (n > 0) {
    atomicMin(s_min, mn);
    atomicMax(s_max, mx);
  }
}

__device__ __forceinline__ float bilinear_ref(float wtl, float wbl, float wtr, float wbr, float tl,
                                              float bl, float tr, float br) {
  // roi_align_v2-inl.h:137-140: ((tl + bl) + tr) + br, each product rounded separately
  return __fadd_rn(
      __fadd_rn(__fadd_rn(__fmul_rn(wtl, tl), __fmul_rn(wbl, bl)), __fmul_rn(wtr, tr)),
      __fmul_rn(wbr, br));
}

// One output element by the reference's own loop (no tables).  Only used when a bin has more
// samples than the tables hold — unreachable for finite rois on maps narrower than 2^17 px.
__device__ void element_direct(const float* __restrict__ plane, int H, int W, int PH, int PW, int ph,
                               int pw, float rsw, float rsh, float rew, float reh, float& best,
                               float& bx, float& by) {
  const float bh = __fdiv_rn(__fsub_rn(reh, rsh), (float)PH);
  const float bw = __fdiv_rn(__fsub_rn(rew, rsw), (float)PW);
  float hs = min_ref(max_ref(__fadd_rn(__fmul_rn((float)ph, bh), rsh), 0.f), (float)(H - 1));
  float he = min_ref(max_ref(__fadd_rn(__fmul_rn((float)(ph + 1), bh), rsh), 0.f), (float)(H - 1));
  float ws = min_ref(max_ref(__fadd_rn(__fmul_rn((float)pw, bw), rsw), 0.f), (float)(W - 1));
  float we = min_ref(max_ref(__fadd_rn(__fmul_rn((float)(pw + 1), bw), rsw), 0.f), (float)(W - 1));
  best = 0.f;
  bx = by = -1.f;
  if (he <= hs || we <= ws) return;
  best = -FLT_MAX;
  const float hst = (float)__ddiv_rn((double)__fsub_rn(he, hs), 3.0);
  const float wst = (float)__ddiv_rn((double)__fsub_rn(we, ws), 3.0);
  int guard = 0;
  for (float h = __fadd_rn(hs, hst); (double)h <= __dadd_rn((double)__fsub_rn(he, hst), 0.01);
       h = __fadd_rn(h, max_ref(hst, 0.01f))) {
    for (float w = __fadd_rn(ws, wst); (double)w <= __dadd_rn((double)__fsub_rn(we, wst), 0.01);
         w = __fadd_rn(w, max_ref(wst, 0.01f))) {
      int hl = min(max((int)floorf(h), 0), H - 1), hh = min(max((int)ceilf(h), 0), H - 1);
      int wl = min(max((int)floorf(w), 0), W - 1), wr = min(max((int)ceilf(w), 0), W - 1);
      float al = (hl == hh) ? 0.5f : __fdiv_rn(__fsub_rn(h, (float)hl), (float)(hh - hl));
      float be = (wl == wr) ? 0.5f : __fdiv_rn(__fsub_rn(w, (float)wl), (float)(wr - wl));
      float a0 = __fsub_rn(1.f, al), b0 = __fsub_rn(1.f, be);
      float v = bilinear_ref(__fmul_rn(a0, b0), __fmul_rn(al, b0), __fmul_rn(a0, be),
                             __fmul_rn(al, be), plane[hl * W + wl], plane[hh * W + wl],
                             plane[hl * W + wr], plane[hh * W + wr]);
      if (v > best) {
        best = v;
        bx = w;
        by = h;
      }
      if (++guard > (1 << 20)) return;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward kernel.  grid = (B*N rois, ceil(C/CT) channel tiles).  Dynamic smem = staged window.
// ---------------------------------------------------------------------------------------------
template <int CT, int CPT, bool kArg>
__global__ void __launch_bounds__(256)
roi_align_v2_fwd_kernel(const __grid_constant__ RoiAlignArgs a, const int phs,
                        const int smem_cap_floats) {
  extern __shared__ float s_win[];
  __shared__ AxisTab s_th, s_tw;
  __shared__ int s_flags, s_hmin, s_hmax, s_wmin, s_wmax;

  const int tid = threadIdx.x;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * CT;
  const int C = a.C, PH = a.PH, PW = a.PW;
  const int b = n / a.N;  // roi_align_v2-inl.h:77

  const float x1 = __ldg(a.rois + 4 * (size_t)n + 0), y1 = __ldg(a.rois + 4 * (size_t)n + 1);
  const float x2 = __ldg(a.rois + 4 * (size_t)n + 2), y2 = __ldg(a.rois + 4 * (size_t)n + 3);

  int li = 0;
  if (a.fpn) {
    const int t = fpn_level(x1, y1, x2, y2, a.scale0, a.lvl0, a.k_min, a.k_max);
    li = -1;
    for (int l = 0; l < a.num_levels; ++l)
      if (a.lvl[l].stride_log2 == t) li = l;
    if (a.levels_out != nullptr && blockIdx.y == 0 && tid == 0) a.levels_out[n] = li;
  }

  const int nct = min(CT, C - c0);
  const size_t out_base = ((size_t)n * C + c0) * PH * PW;

  if (li < 0) {  // roi matched no level: the reference zeroes it on every level -> all-empty
    for (int e = tid; e < nct * PH * PW; e += blockDim.x) {
      a.out[out_base + e] = 0.f;
      if (kArg) {
        a.argx[out_base + e] = -1.f;
        a.argy[out_base + e] = -1.f;
      }
    }
    return;
  }

  const Level& L = a.lvl[li];
  const int H = L.H, W = L.W;
  const float scale = L.scale;
  const float rsw = __fmul_rn(x1, scale), rsh = __fmul_rn(y1, scale);
  const float rew = __fmul_rn(x2, scale), reh = __fmul_rn(y2, scale);

  if (tid == 0) {
    s_flags = 0;
    s_hmin = INT_MAX;
    s_hmax = -1;
    s_wmin = INT_MAX;
    s_wmax = -1;
  }
  __syncthreads();
  if (tid < PH)
    build_axis_bin(s_th, tid, PH, rsh, reh, H, &s_flags, &s_hmin, &s_hmax);
  else if (tid < PH + PW)
    build_axis_bin(s_tw, tid - PH, PW, rsw, rew, W, &s_flags, &s_wmin, &s_wmax);
  __syncthreads();

  const int flags = s_flags;
  const int hmin = s_hmin, wmin = s_wmin;
  const int Hwin = s_hmax - hmin + 1, Wwin = s_wmax - wmin + 1;
  const bool any = (s_hmax >= 0) && (s_wmax >= 0);
  const size_t HW = (size_t)H * W;
  const float* gplane0 = L.data + ((size_t)b * C + c0) * HW;

  const bool fast = any && (flags == 0) && ((long long)CT * Hwin * Wwin <= smem_cap_floats);

  if (!fast) {
    // ---- generic path: one thread per output element, taps straight from global/L1 ----
    for (int e = tid; e < nct * PH * PW; e += blockDim.x) {
      const int pw = e % PW, ph = (e / PW) % PH, cl = e / (PW * PH);
      const float* plane = gplane0 + (size_t)cl * HW;
      const int nh = s_th.cnt[ph], nw = s_tw.cnt[pw];
      float best = 0.f, bx = -1.f, by = -1.f;
      if (flags & kFlagOverflow) {
        element_direct(plane, H, W, PH, PW, ph, pw, rsw, rsh, rew, reh, best, bx, by);
      } else if (nh >= 0 && nw >= 0) {
        best = -FLT_MAX;
        for (int i = 0; i < nh; ++i) {
          const int hi_ = ph * kMaxS + i;
          const int hl = s_th.lo[hi_], hh = s_th.hi[hi_];
          const float a0 = s_th.w0[hi_], a1 = s_th.w1[hi_];
          for (int j = 0; j < nw; ++j) {
            const int wj = pw * kMaxS + j;
            const int wl = s_tw.lo[wj], wr = s_tw.hi[wj];
            const float b0 = s_tw.w0[wj], b1 = s_tw.w1[wj];
            const float v = bilinear_ref(__fmul_rn(a0, b0), __fmul_rn(a1, b0), __fmul_rn(a0, b1),
                                         __fmul_rn(a1, b1), __ldg(plane + hl * W + wl),
                                         __ldg(plane + hh * W + wl), __ldg(plane + hl * W + wr),
                                         __ldg(plane + hh * W + wr));
            if (v > best) {
              best = v;
              bx = s_tw.coord[wj];
              by = s_th.coord[hi_];
            }
          }
        }
      }
      a.out[out_base + e] = best;
      if (kArg) {
        a.argx[out_base + e] = bx;
        a.argy[out_base + e] = by;
      }
    }
    return;
  }

  // ---- stage the window: rows are contiguous in NCHW -> lanes walk x, warps walk rows ----
  const int plane_sz = Hwin * Wwin;
  {
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int lpr_log2 = Wwin <= 8 ? 3 : (Wwin <= 16 ? 4 : 5);
    const int lpr = 1 << lpr_log2, rpw = 32 >> lpr_log2;
    const int sub = lane >> lpr_log2, xl = lane & (lpr - 1);
    const float* g0 = gplane0 + (size_t)hmin * W + wmin;
    for (int y = warp * rpw + sub; y < Hwin; y += nwarps * rpw) {
      for (int x = xl; x < Wwin; x += lpr) {
        const float* g = g0 + (size_t)y * W + x;
        float* s = s_win + y * Wwin + x;
        float v[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) v[c] = (c < nct) ? __ldg(g + (size_t)c * HW) : 0.f;
#pragma unroll
        for (int c = 0; c < CT; ++c) s[c * plane_sz] = v[c];
      }
    }
  }
  __syncthreads();

  // ---- compute: thread = (pw, channel group, ph chunk) ----
  constexpr int NCG = CT / CPT;
  if (tid >= PW * NCG * phs) return;
  const int pw = tid % PW;
  const int rest = tid / PW;
  const int cg = rest % NCG, pc = rest / NCG;
  const int chunk = (PH + phs - 1) / phs;
  const int ph_beg = pc * chunk, ph_end = min(PH, ph_beg + chunk);

  const int wcnt = s_tw.cnt[pw];
  int xo[4] = {0, 0, 0, 0};
  float b0[2] = {0.f, 0.f}, b1[2] = {0.f, 0.f}, wc[2] = {-1.f, -1.f};
  if (wcnt == 2) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int j = pw * kMaxS + s;
      xo[2 * s] = s_tw.lo[j] - wmin;
      xo[2 * s + 1] = s_tw.hi[j] - wmin;
      b0[s] = s_tw.w0[j];
      b1[s] = s_tw.w1[j];
      wc[s] = s_tw.coord[j];
    }
  }
  const float* sw = s_win + cg * CPT * plane_sz;
  const int cbase = c0 + cg * CPT;
  const size_t PP = (size_t)PH * PW;
  float* outp = a.out + ((size_t)n * C + cbase) * PP + pw;
  float* axp = kArg ? a.argx + ((size_t)n * C + cbase) * PP + pw : nullptr;
  float* ayp = kArg ? a.argy + ((size_t)n * C + cbase) * PP + pw : nullptr;

  float Lr[CPT][4], Hr[CPT][4];
#pragma unroll
  for (int k = 0; k < CPT; ++k)
#pragma unroll
    for (int t = 0; t < 4; ++t) Lr[k][t] = Hr[k][t] = 0.f;
  int rowL = -1, rowH = -1;

  for (int ph = ph_beg; ph < ph_end; ++ph) {
    const int hcnt = s_th.cnt[ph];
    float best[CPT];
    int bi[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      best[k] = -FLT_MAX;
      bi[k] = -1;
    }
    if (hcnt == 2) {
#pragma unroll
      for (int hs = 0; hs < 2; ++hs) {
        const int i = ph * kMaxS + hs;
        const int lo = s_th.lo[i] - hmin, hi = s_th.hi[i] - hmin;
        if (lo != rowL) {
          if (lo == rowH) {
#pragma unroll
            for (int k = 0; k < CPT; ++k)
#pragma unroll
              for (int t = 0; t < 4; ++t) Lr[k][t] = Hr[k][t];
          } else {
            const float* r = sw + lo * Wwin;
#pragma unroll
            for (int k = 0; k < CPT; ++k)
#pragma unroll
              for (int t = 0; t < 4; ++t) Lr[k][t] = r[k * plane_sz + xo[t]];
          }
          rowL = lo;
        }
        if (hi != rowH) {
          const float* r = sw + hi * Wwin;
#pragma unroll
          for (int k = 0; k < CPT; ++k)
#pragma unroll
            for (int t = 0; t < 4; ++t) Hr[k][t] = r[k * plane_sz + xo[t]];
          rowH = hi;
        }
        const float a0 = s_th.w0[i], a1 = s_th.w1[i];
        const float wtl0 = __fmul_rn(a0, b0[0]), wbl0 = __fmul_rn(a1, b0[0]);
        const float wtr0 = __fmul_rn(a0, b1[0]), wbr0 = __fmul_rn(a1, b1[0]);
        const float wtl1 = __fmul_rn(a0, b0[1]), wbl1 = __fmul_rn(a1, b0[1]);
        const float wtr1 = __fmul_rn(a0, b1[1]), wbr1 = __fmul_rn(a1, b1[1]);
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const float v0 =
              bilinear_ref(wtl0, wbl0, wtr0, wbr0, Lr[k][0], Hr[k][0], Lr[k][1], Hr[k][1]);
          const float v1 =
              bilinear_ref(wtl1, wbl1, wtr1, wbr1, Lr[k][2], Hr[k][2], Lr[k][3], Hr[k][3]);
          if (v0 > best[k]) {
            best[k] = v0;
            bi[k] = 2 * hs;
          }
          if (v1 > best[k]) {
            best[k] = v1;
            bi[k] = 2 * hs + 1;
          }
        }
      }
    }
    const bool empty = (hcnt < 0) || (wcnt < 0);
    const float hc0 = s_th.coord[ph * kMaxS], hc1 = s_th.coord[ph * kMaxS + 1];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      if (cbase + k < C) {
        const size_t o = (size_t)k * PP + (size_t)ph * PW;
        outp[o] = empty ? 0.f : best[k];
        if (kArg) {
          const bool none = empty || bi[k] < 0;
          axp[o] = none ? -1.f : ((bi[k] & 1) ? wc[1] : wc[0]);
          ayp[o] = none ? -1.f : ((bi[k] & 2) ? hc1 : hc0);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward (roi_align_v2.cu:35-84): one thread per output-gradient element, 4 red.global adds.
// ---------------------------------------------------------------------------------------------
struct BwdArgs {
  Level lvl[SDET_MAX_LEVELS];
  const float* ograd;
  const float* argx;
  const float* argy;
  const int32_t* levels;  // nullptr: everything on lvl[0]
  int B, N, C, PH, PW;
};

__global__ void __launch_bounds__(256)
roi_align_v2_bwd_kernel(const __grid_constant__ BwdArgs a, const size_t count) {
  const size_t PP = (size_t)a.PH * a.PW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count;
       idx += (size_t)gridDim.x * blockDim.x) {
    const float ax = __ldg(a.argx + idx), ay = __ldg(a.argy + idx);
    if (ax == -1.f || ay == -1.f) continue;
    const size_t nc = idx / PP;
    const int c = (int)(nc % a.C);
    const int n = (int)(nc / a.C);
    int li = 0;
    if (a.levels != nullptr) {
      li = __ldg(a.levels + n);
      if (li < 0) continue;
    }
    const Level& L = a.lvl[li];
    const int H = L.H, W = L.W;
    const int b = n / a.N;
    float* g = L.grad + ((size_t)b * a.C + c) * H * W;
    const int hl = min(max((int)floorf(ay), 0), H - 1), hh = min(max((int)ceilf(ay), 0), H - 1);
    const int wl = min(max((int)floorf(ax), 0), W - 1), wr = min(max((int)ceilf(ax), 0), W - 1);
    const float al = (hl == hh) ? 0.5f : __fdiv_rn(__fsub_rn(ay, (float)hl), (float)(hh - hl));
    const float be = (wl == wr) ? 0.5f : __fdiv_rn(__fsub_rn(ax, (float)wl), (float)(wr - wl));
    const float d = __ldg(a.ograd + idx);
    const float a0 = __fsub_rn(1.f, al), b0 = __fsub_rn(1.f, be);
    // roi_align_v2.cu:79-82: (top_diff * wy) * wx, rounded product by product
    atomicAdd(g + hl * W + wl, __fmul_rn(__fmul_rn(d, a0), b0));
    atomicAdd(g + hl * W + wr, __fmul_rn(__fmul_rn(d, a0), be));
    atomicAdd(g + hh * W + wl, __fmul_rn(__fmul_rn(d, al), b0));
    atomicAdd(g + hh * W + wr, __fmul_rn(__fmul_rn(d, al), be));
  }
}

int check_common(int B, int N, int C, int ph, int pw) {
  if (B <= 0 || N <= 0 || C <= 0) return sdet::fail(SDET_ERR_INVALID_ARG, "B, N, C must be > 0");
  if (ph <= 0 || pw <= 0)
    return sdet::fail(SDET_ERR_INVALID_ARG, "pooled_size must be non-zero (enforce_nonzero)");
  if (ph > kMaxP || pw > kMaxP)
    return sdet::fail(SDET_ERR_UNSUPPORTED, "pooled_size > %d per axis is not supported", kMaxP);
  return SDET_OK;
}

template <int CT, int CPT>
int launch_fwd_t(const RoiAlignArgs& a, int phs, int threads, cudaStream_t st) {
  static int smem_cap = -1;  // bytes this kernel may use; same for every instantiation
  const bool arg = a.argx != nullptr;
  auto k_inf = roi_align_v2_fwd_kernel<CT, CPT, false>;
  auto k_trn = roi_align_v2_fwd_kernel<CT, CPT, true>;
  const int want = 64 * 1024;
  if (smem_cap < 0) {
    SDET_CUDA(cudaFuncSetAttribute(k_inf, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
    SDET_CUDA(cudaFuncSetAttribute(k_trn, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
    smem_cap = want;
  }
  dim3 grid((unsigned)(a.B * a.N), (unsigned)((a.C + CT - 1) / CT));
  if (arg)
    k_trn<<<grid, threads, smem_cap, st>>>(a, phs, smem_cap / 4);
  else
    k_inf<<<grid, threads, smem_cap, st>>>(a, phs, smem_cap / 4);
  SDET_LAUNCH_CHECK("roi_align_v2_fwd_kernel");
  return SDET_OK;
}

int launch_fwd(const RoiAlignArgs& a, cudaStream_t st) {
  if ((a.argx == nullptr) != (a.argy == nullptr))
    return sdet::fail(SDET_ERR_INVALID_ARG, "argmax_x and argmax_y must both be given or both NULL");
  // thread = (pw, channel group, ph chunk); pick CPT / ph split so a CTA has ~128 compute threads
  constexpr int CT = 16;
  const int PW = a.PW, PH = a.PH;
  int cpt = (PW * (CT / 4) * 2 >= 96) ? 4 : 2;
  int phs = 0;
  // tuning overrides (benchmarks/roi_align_sweep.py); not part of the ABI
  if (const char* e = getenv("SDET_RA_CPT")) cpt = atoi(e) == 2 ? 2 : 4;
  if (const char* e = getenv("SDET_RA_PHS")) phs = atoi(e);
  int ncg = CT / cpt;
  if (phs <= 0) {
    phs = 1;
    while (PW * ncg * phs < 112 && phs < PH) ++phs;
  }
  if (phs > PH) phs = PH;
  int threads = ((PW * ncg * phs + 31) / 32) * 32;
  if (threads < ((PH + PW + 31) / 32) * 32) threads = ((PH + PW + 31) / 32) * 32;
  if (threads > 256) return sdet::fail(SDET_ERR_UNSUPPORTED, "pooled size needs > 256 threads");
  return cpt == 4 ? launch_fwd_t<CT, 4>(a, phs, threads, st) : launch_fwd_t<CT, 2>(a, phs, threads, st);
}

}  // namespace

extern "C" int sdet_roi_align_v2_forward(const float* data, const float* rois, float* out,
                                         float* argmax_x, float* argmax_y, int B, int N, int C,
                                         int H, int W, int pooled_h, int pooled_w,
                                         float spatial_scale, void* stream) {
  if (int rc = check_common(B, N, C, pooled_h, pooled_w)) return rc;
  SDET_REQUIRE(data && rois && out, "data, rois and out must be non-NULL");
  SDET_REQUIRE(H > 0 && W > 0, "H, W must be > 0");
  // DMLC_DECLARE_FIELD(spatial_scale).set_range(0.0, 1.0)  (roi_align_v2-inl.h:34)
  SDET_REQUIRE(spatial_scale >= 0.f && spatial_scale <= 1.f, "spatial_scale must be in [0, 1]");
  RoiAlignArgs a{};
  a.lvl[0] = Level{data, nullptr, H, W, spatial_scale, 0};
  a.num_levels = 1;
  a.fpn = 0;
  a.rois = rois;
  a.out = out;
  a.argx = argmax_x;
  a.argy = argmax_y;
  a.levels_out = nullptr;
  a.B = B; a.N = N; a.C = C; a.PH = pooled_h; a.PW = pooled_w;
  return launch_fwd(a, (cudaStream_t)stream);
}

static int ilog2_exact(int v) {
  if (v <= 0 || (v & (v - 1))) return -1;
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

extern "C" int sdet_fpn_roi_align_v2_forward(const float* const* feats, const int* H, const int* W,
                                             const int* strides, int num_levels, const float* rois,
                                             float* out, float* argmax_x, float* argmax_y,
                                             int32_t* levels_out, int B, int N, int C, int pooled_h,
                                             int pooled_w, int roi_canonical_scale,
                                             int roi_canonical_level, void* stream) {
  if (int rc = check_common(B, N, C, pooled_h, pooled_w)) return rc;
  SDET_REQUIRE(feats && H && W && strides && rois && out, "NULL argument");
  SDET_REQUIRE(num_levels >= 1 && num_levels <= SDET_MAX_LEVELS, "num_levels must be in [1, %d]",
               SDET_MAX_LEVELS);
  SDET_REQUIRE(roi_canonical_scale > 0, "roi_canonical_scale must be > 0");
  RoiAlignArgs a{};
  int smin = INT_MAX, smax = 0;
  for (int l = 0; l < num_levels; ++l) {
    const int lg = ilog2_exact(strides[l]);
    if (lg < 0)
      return sdet::fail(SDET_ERR_UNSUPPORTED, "stride %d is not a power of two", strides[l]);
    SDET_REQUIRE(feats[l] && H[l] > 0 && W[l] > 0, "level %d: bad feature pointer / shape", l);
    a.lvl[l] = Level{feats[l], nullptr, H[l], W[l], 1.0f / (float)strides[l], lg};
    smin = strides[l] < smin ? strides[l] : smin;
    smax = strides[l] > smax ? strides[l] : smax;
  }
  a.num_levels = num_levels;
  a.fpn = 1;
  a.scale0 = (float)roi_canonical_scale;
  a.lvl0 = (float)roi_canonical_level;
  a.k_min = (float)ilog2_exact(smin);  // np.log2(min(rcnn_stride)), assign_layer_fpn.py:24
  a.k_max = (float)ilog2_exact(smax);
  a.rois = rois;
  a.out = out;
  a.argx = argmax_x;
  a.argy = argmax_y;
  a.levels_out = levels_out;
  a.B = B; a.N = N; a.C = C; a.PH = pooled_h; a.PW = pooled_w;
  return launch_fwd(a, (cudaStream_t)stream);
}

static int launch_bwd(BwdArgs& a, int num_levels, const int* H, const int* W, float* const* grads,
                      int accumulate, float* grad_rois, cudaStream_t st) {
  for (int l = 0; l < num_levels; ++l) {
    SDET_REQUIRE(grads[l] && H[l] > 0 && W[l] > 0, "level %d: bad grad pointer / shape", l);
    a.lvl[l] = Level{nullptr, grads[l], H[l], W[l], 0.f, 0};
    if (!accumulate)  // kWriteTo: Fill 0 (roi_align_v2.cu:130-133)
      SDET_CUDA(cudaMemsetAsync(grads[l], 0, sizeof(float) * (size_t)a.B * a.C * H[l] * W[l], st));
  }
  if (grad_rois)  // roi_align_v2.cu:139-141
    SDET_CUDA(cudaMemsetAsync(grad_rois, 0, sizeof(float) * (size_t)a.B * a.N * 4, st));
  const size_t count = (size_t)a.B * a.N * a.C * a.PH * a.PW;
  const int threads = 256;
  size_t blocks = (count + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  roi_align_v2_bwd_kernel<<<(unsigned)blocks, threads, 0, st>>>(a, count);
  SDET_LAUNCH_CHECK("roi_align_v2_bwd_kernel");
  return SDET_OK;
}

extern "C" int sdet_roi_align_v2_backward(const float* ograd, const float* argmax_x,
                                          const float* argmax_y, float* grad_data, float* grad_rois,
                                          int B, int N, int C, int H, int W, int pooled_h,
                                          int pooled_w, int accumulate, void* stream) {
  if (int rc = check_common(B, N, C, pooled_h, pooled_w)) return rc;
  SDET_REQUIRE(ograd && argmax_x && argmax_y && grad_data, "NULL argument");
  BwdArgs a{};
  a.ograd = ograd; a.argx = argmax_x; a.argy = argmax_y; a.levels = nullptr;
  a.B = B; a.N = N; a.C = C; a.PH = pooled_h; a.PW = pooled_w;
  float* g[1] = {grad_data};
  return launch_bwd(a, 1, &H, &W, g, accumulate, grad_rois, (cudaStream_t)stream);
}

extern "C" int sdet_fpn_roi_align_v2_backward(const float* ograd, const float* argmax_x,
                                              const float* argmax_y, const int32_t* levels,
                                              float* const* grad_feats, const int* H, const int* W,
                                              int num_levels, int B, int N, int C, int pooled_h,
                                              int pooled_w, int accumulate, void* stream) {
  if (int rc = check_common(B, N, C, pooled_h, pooled_w)) return rc;
  SDET_REQUIRE(ograd && argmax_x && argmax_y && levels && grad_feats && H && W, "NULL argument");
  SDET_REQUIRE(num_levels >= 1 && num_levels <= SDET_MAX_LEVELS, "num_levels must be in [1, %d]",
               SDET_MAX_LEVELS);
  BwdArgs a{};
  a.ograd = ograd; a.argx = argmax_x; a.argy = argmax_y; a.levels = levels;
  a.B = B; a.N = N; a.C = C; a.PH = pooled_h; a.PW = pooled_w;
  return launch_bwd(a, num_levels, H, W, grad_feats, accumulate, nullptr, (cudaStream_t)stream);
}
