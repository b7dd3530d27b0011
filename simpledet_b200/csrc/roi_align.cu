// _contrib_ROIAlign_v2 forward/backward and the fused FPN variant, hand-written for sm_100a.
//
// Reference semantics: operator_cxx/contrib/roi_align_v2-inl.h:61-153 (forward functor),
// operator_cxx/contrib/roi_align_v2.cu:17-85 (backward), models/FPN/assign_layer_fpn.py:17-40
// (level assignment).  NOT a port: the reference runs one thread per output element with 16
// scattered global loads; here one CTA owns (roi, channel tile):
//
//   0. plan      — (with a workspace) one small kernel restates the reference's float/double sample
//                  loop ONCE per roi (the sample coordinates depend only on (roi, ph) / (roi, pw),
//                  not on the channel) into axis tables — coordinate, lo/hi pixel, weights — and
//                  files the roi into a window-size class; a second one turns the classes into a
//                  largest-window-first order for the main kernel's blockIdx.x.
//   1. preamble  — the CTA loads its roi's tables (or computes them inline without a workspace).
//   2. stage     — ONE producer warp copies the roi's feature window [CT, Hwin, Wwin]
//                  NCHW-row-coalesced into a ring of shared-memory buffers with cp.async and
//                  signals full/empty mbarriers.
//   3. compute   — FOUR consumer warps; thread = (pw, w-sample[, channel half]) x (channel group,
//                  ph chunk) walks down the rows with a 2-row register cache, so a window element
//                  is read from shared memory ~once per column tap instead of once per sample;
//                  bilinear weights are shared across the channels a thread owns.
//
// Arithmetic is bit-identical to the reference's CPU build: every float op on the coordinate
// and value path is an explicit round-to-nearest intrinsic (no FMA contraction), the
// `(hend-hstart)/3.0` and `<= hend-h_stride+0.01` promotions to double are kept.
#include <cfloat>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kMaxS = 4;                 // samples per bin per axis kept in the tables
constexpr int kMaxP = SDET_MAX_POOLED;   // pooled size limit per axis
constexpr int kFlagNot2 = 1;             // some non-empty bin does not have exactly 2 samples
constexpr int kFlagOverflow = 2;         // some bin has more than kMaxS samples
constexpr int kFlagEmpty = 4;            // some bin is empty along an axis (end <= start)

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
  const void* plans;  // per-roi preamble records written by roi_align_plan_kernel, or nullptr
  const int* order;   // CTA x -> roi, largest window first (roi_align_order_kernel), or nullptr = identity
  const int* order_count;  // device count of valid `order` entries (CTAs beyond it exit), or nullptr = all
  uint64_t negzero2;  // {-0.0f,-0.0f}: opaque addend that keeps FFMA2 products exact
};

template <int TP>  // TP = max bins per axis this instantiation handles (16 or kMaxP)
struct AxisTab {
  float coord[TP * kMaxS];
  float w0[TP * kMaxS];  // 1 - alpha
  float w1[TP * kMaxS];  // alpha
  int lo[TP * kMaxS];
  int hi[TP * kMaxS];
  int cnt[TP];  // -1: bin empty along this axis (end <= start); else #samples (may be 0)
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
template <int TP>
__device__ void build_axis_bin(AxisTab<TP>& t, int p, int P, float roi_start, float roi_end,
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
    atomicOr(s_flags, kFlagEmpty);
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
// Forward kernel.  grid = (B*N rois, channel-tile groups).  One CTA = one roi x `tiles` channel
// tiles; the per-roi preamble is paid once per CTA.  Dynamic smem (kCapFloats) holds 1 or 2
// window buffers; the channel stride inside a buffer is a compile-time constant (kCS) so every
// tap load is `LDS [reg + imm]`.
//
//   warp  = (channel group cg of CPT channels, ph chunk pc)         -> everything warp-uniform
//   lane  = 2*pw + s : one of the two w-samples of output column pw -> the 28 lanes of a warp
//           read 28 columns of ONE shared-memory row: bank-conflict free by layout
//   the two lanes of a pw exchange their maxima with one shuffle; lane s=0 stores.
//
// Arithmetic: channel pairs run on the packed fp32x2 pipe (FFMA2 / FADD2).  Each product is
// `fma.rn.f32x2(w, x, -0.0)` with the -0.0 coming from a kernel argument, which rounds exactly
// like a lone multiply and cannot be contracted with the following add (ptxas fuses a bare
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2, which would change the rounding); sums are
// add.rn.f32x2 in the reference's order.  Result: bit-identical to the scalar sequence.
// ---------------------------------------------------------------------------------------------

struct HRow {   // per h-sample, 16 bytes, read with one LDS.128
  int off_lo;   // BYTE offset of pixel (lo, wmin) inside a channel plane, incl. the row's 16B shift
  int off_hi;
  float w0;     // 1 - alpha
  float w1;     // alpha
};

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
// mbarrier helpers (CTA scope) for the producer-warp pipeline
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrives on `bar` once every cp.async this thread has issued so far has landed (count pre-charged)
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra W_%=;\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
template <int IMM>
__device__ __forceinline__ float lds_f32_imm(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
  return v;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;  // 3-input max (sm_100+): NaN operands are dropped like fmaxf does
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int CPT, int kCS, int K = 0>
struct TapLoader {  // R[k][t] = smem[base_t + k*kCS*4], fully unrolled with immediate offsets
  static __device__ __forceinline__ void run(float (&R)[CPT][2], unsigned al, unsigned ar) {
    R[K][0] = lds_f32_imm<K * kCS * 4>(al);
    R[K][1] = lds_f32_imm<K * kCS * 4>(ar);
    TapLoader<CPT, kCS, K + 1>::run(R, al, ar);
  }
};
template <int CPT, int kCS>
struct TapLoader<CPT, kCS, CPT> {
  static __device__ __forceinline__ void run(float (&)[CPT][2], unsigned, unsigned) {}
};

// ---------------------------------------------------------------------------------------------
// Per-roi preamble: FPN level + the reference's sample loop restated into the axis tables.
// It is a ~3 us serial dependency chain (double-precision divide + float/double loop), so the
// forward launch runs it ONCE per roi in `roi_align_plan_kernel` and every (roi, channel group)
// CTA of the main kernel just loads the 2.7 KB record; without a workspace the main kernel runs it
// inline (same code, same results).
// s_scal: {li, flags, hmin, hmax, wmin, wmax}
// ---------------------------------------------------------------------------------------------
template <int TP>
__device__ __forceinline__ void roi_preamble(const RoiAlignArgs& a, const int n, const int PH, const int PW,
                                             AxisTab<TP>& s_th, AxisTab<TP>& s_tw, int* s_scal) {
  const int tid = threadIdx.x;
  const float x1 = __ldg(a.rois + 4 * (size_t)n + 0), y1 = __ldg(a.rois + 4 * (size_t)n + 1);
  const float x2 = __ldg(a.rois + 4 * (size_t)n + 2), y2 = __ldg(a.rois + 4 * (size_t)n + 3);
  int li = 0;
  if (a.fpn) {
    const int t = fpn_level(x1, y1, x2, y2, a.scale0, a.lvl0, a.k_min, a.k_max);
    li = -1;
    for (int l = 0; l < a.num_levels; ++l)
      if (a.lvl[l].stride_log2 == t) li = l;
  }
  if (tid == 0) {
    s_scal[0] = li;
    s_scal[1] = 0;
    s_scal[2] = INT_MAX;
    s_scal[3] = -1;
    s_scal[4] = INT_MAX;
    s_scal[5] = -1;
  }
  __syncthreads();
  if (li >= 0) {
    const Level& L = a.lvl[li];
    const float scale = L.scale;
    if (tid < PH)
      build_axis_bin(s_th, tid, PH, __fmul_rn(y1, scale), __fmul_rn(y2, scale), L.H, &s_scal[1], &s_scal[2], &s_scal[3]);
    else if (tid < PH + PW)
      build_axis_bin(s_tw, tid - PH, PW, __fmul_rn(x1, scale), __fmul_rn(x2, scale), L.W, &s_scal[1], &s_scal[4],
                     &s_scal[5]);
  }
  __syncthreads();
}

struct PlanRecord {  // what a (roi, channel group) CTA needs from the preamble (TP = 16)
  int scal[8];
  AxisTab<16> th, tw;
};
static_assert(sizeof(PlanRecord) % 16 == 0, "PlanRecord is copied with 16-byte accesses");

// Scheduling.  CTA run time grows with the roi's window (bytes staged), and windows span 100..1500
// cells: in launch order the last wave is held up by whichever large rois happen to start late
// (measured: 17 % of the 7x7 kernel, 8 % of the 14x14 one).  The plan kernel therefore also files
// every roi into one of kCostBuckets window-size classes (atomic rank inside the class) and
// roi_align_order_kernel turns that into a largest-first order for the main kernel's blockIdx.x.
constexpr int kCostBuckets = 64;
struct PlanSched {
  int* counts;   // kCostBuckets, zeroed before the plan kernel
  int2* slot;    // (bucket, rank) per roi
  int* order;    // position -> roi
};

__global__ void __launch_bounds__(64) roi_align_plan_kernel(const __grid_constant__ RoiAlignArgs a,
                                                            PlanRecord* __restrict__ plans, const PlanSched sc) {
  __shared__ __align__(16) PlanRecord s_rec;
  const int n = blockIdx.x;
  roi_preamble<16>(a, n, a.PH, a.PW, s_rec.th, s_rec.tw, s_rec.scal);
  if (a.levels_out != nullptr && threadIdx.x == 0) a.levels_out[n] = s_rec.scal[0];
  const int4* src = reinterpret_cast<const int4*>(&s_rec);
  int4* dst = reinterpret_cast<int4*>(plans + n);
  for (int i = threadIdx.x; i < (int)(sizeof(PlanRecord) / 16); i += blockDim.x) dst[i] = src[i];
  if (threadIdx.x == 0) {
    int bucket = 0;  // no level / empty window: cheapest
    if (s_rec.scal[0] >= 0 && s_rec.scal[3] >= 0 && s_rec.scal[5] >= 0) {
      const int hwin = s_rec.scal[3] - s_rec.scal[2] + 1, wwin = s_rec.scal[5] - s_rec.scal[4] + 1;
      const int cells = hwin * ((wwin + 6) & ~3);
      bucket = min(kCostBuckets - 1, 1 + cells / 24);
      if (s_rec.scal[1] & (kFlagNot2 | kFlagOverflow)) bucket = kCostBuckets - 1;  // generic path: slowest
    }
    sc.slot[n] = make_int2(bucket, atomicAdd(&sc.counts[bucket], 1));
  }
}

__global__ void __launch_bounds__(256) roi_align_order_kernel(const PlanSched sc, const int total) {
  __shared__ int s_base[kCostBuckets];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = kCostBuckets - 1; b >= 0; --b) {  // largest class first
      s_base[b] = acc;
      acc += sc.counts[b];
    }
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < total) {
    const int2 s = sc.slot[n];
    sc.order[s_base[s.x] + s.y] = n;
  }
}

// kCapFloats: floats of dynamic shared memory for the window buffers (12288 = 48 KB -> 4 CTAs/SM)
#ifdef SDET_RA_ABLATE  // profiling builds only: bit 0 skips compute, bit 1 skips staging (results are garbage)
__device__ int g_ra_ablate = 0;
// cycle counters: [0] consumer warps waiting for a full buffer, [1] computing, [2] CTA lifetime up to the tile
// loop (preamble, tables, barrier init; thread 0), [3] producer waiting, [4] whole CTA lifetime (thread 0)
__device__ unsigned long long g_ra_prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
// Timeline of ONE CTA (blockIdx.x == g_ra_trace_cta, blockIdx.y == 0), clock64 ticks relative to its start:
// row 0 = consumer warp 0 {after preamble, then per tile: data arrived, compute done}, row 1 = producer warp
// {per tile: buffer free, copies issued}.  Read back with sdet_debug_ra_trace.
__device__ int g_ra_trace_cta = -1;
__device__ long long g_ra_trace[2][520];
#endif

// CTA = 4 consumer warps (the arithmetic) + 1 producer warp (all cp.async staging).  Staging and
// arithmetic were measured to ADD, not overlap, when the same four warps did both (issue slots at
// 16 warps/SM); a dedicated producer takes the LDGSTS + address arithmetic off the consumers' path
// and replaces the two CTA-wide barriers per tile by full/empty mbarriers.
constexpr int kFwdThreads = 160;

template <int CPT, bool kArg, int kPH, int kPW, int kCapFloats>
__global__ void __launch_bounds__(kFwdThreads, (kCapFloats <= 10240 ? 5 : 4))
roi_align_v2_fwd_kernel(const __grid_constant__ RoiAlignArgs a, const int tiles) {
  extern __shared__ __align__(16) float s_win[];
  constexpr int TP = (kPH > 0 && kPH <= 16 && kPW <= 16) ? 16 : kMaxP;
  __shared__ __align__(16) AxisTab<TP> s_th, s_tw;
  __shared__ __align__(16) HRow s_hrow[TP * kMaxS];
  __shared__ __align__(16) int s_scal[8];  // {li, flags, hmin, hmax, wmin, wmax}
  __shared__ __align__(8) unsigned long long s_bar[8];  // full[0..3], empty[0..3]

  constexpr int NW = 4;  // warps per CTA
  static_assert(CPT % 4 == 0, "channels are processed in fp32x2 pairs, per half-warp when PW <= 8");

  const int tid = threadIdx.x;
#ifdef SDET_RA_ABLATE
  const long long prof_t0 = clock64();
#endif
  if (a.order_count != nullptr && (int)blockIdx.x >= __ldg(a.order_count)) return;
  const int n = a.order ? __ldg(a.order + blockIdx.x) : (int)blockIdx.x;
  const int C = a.C;
  const int PH = kPH ? kPH : a.PH, PW = kPW ? kPW : a.PW;
  const int PP = PH * PW;
  const int b = n / a.N;  // roi_align_v2-inl.h:77
  const int cgrp0 = blockIdx.y * tiles * (2 * CPT);            // first channel of this CTA
  const int cgrp1 = min(C, cgrp0 + tiles * (2 * CPT));         // one past the last

  // ---- preamble: load the roi's record written by roi_align_plan_kernel, or compute it here
  if (TP == 16 && a.plans != nullptr) {
    const PlanRecord* rec = static_cast<const PlanRecord*>(a.plans) + n;
    constexpr int kHdr = 32 / 16, kTabV = (int)(sizeof(AxisTab<16>) / 16);
    const int4* src = reinterpret_cast<const int4*>(rec);
    for (int i = tid; i < kHdr + 2 * kTabV; i += blockDim.x) {
      const int4 v = __ldg(src + i);
      if (i < kHdr) reinterpret_cast<int4*>(s_scal)[i] = v;
      else if (i < kHdr + kTabV) reinterpret_cast<int4*>(&s_th)[i - kHdr] = v;
      else reinterpret_cast<int4*>(&s_tw)[i - kHdr - kTabV] = v;
    }
    __syncthreads();
  } else {
    roi_preamble<TP>(a, n, PH, PW, s_th, s_tw, s_scal);
    if (a.levels_out != nullptr && blockIdx.y == 0 && tid == 0) a.levels_out[n] = s_scal[0];
  }
  const int li = s_scal[0];

  const size_t out_base = ((size_t)n * C + cgrp0) * PP;
  const int nelem = (cgrp1 - cgrp0) * PP;

  if (li < 0) {  // roi matched no level: the reference zeroes it on every level -> all-empty
    for (int e = tid; e < nelem; e += blockDim.x) {
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

  const int flags = s_scal[1];
  const int hmin = s_scal[2], wmin = s_scal[4];
  const int s_hmax = s_scal[3], s_wmax = s_scal[5];
  const int Hwin = s_hmax - hmin + 1, Wwin = s_wmax - wmin + 1;
  const bool any = (s_hmax >= 0) && (s_wmax >= 0);
  const int HW = H * W;
  const float* gimg = L.data + (size_t)b * C * HW;  // image b, channel 0

  // 16-byte staging needs every channel plane to start 16B-aligned
  const bool vec = ((HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(L.data) & 15) == 0);
  const int Wp = vec ? ((Wwin + 3 + 3) & ~3) : ((Wwin + 3) & ~3);  // smem row pitch (floats)
  const int plane = Hwin * Wp;

  // Buffering mode by window size (kCS = channel stride in floats, ct = channels per tile):
  //   0: ct=2*CPT, 2 buffers, kCS=cap/(4*CPT)   1: ct=2*CPT, 1 buffer, kCS=cap/(2*CPT)
  //   2: ct=CPT,   2 buffers, kCS=cap/(2*CPT)   3: ct=CPT,   1 buffer, kCS=cap/CPT
  constexpr int CS0 = kCapFloats / (4 * CPT), CS1 = kCapFloats / (2 * CPT), CS3 = kCapFloats / CPT;
  int mode = -1;
  if (plane <= CS0) mode = 0;
  else if (plane <= CS1) mode = 1;
  else if (plane <= CS3) mode = 3;
  const bool fast = any && ((flags & (kFlagNot2 | kFlagOverflow)) == 0) && (PW <= 16) && (Wp <= 64) && (Hwin <= 128) &&
                    (mode >= 0) && ((cgrp1 - cgrp0) % (2 * CPT) == 0);

  if (!fast) {
    // ---- generic path: one thread per output element, taps straight from global/L1 ----
    for (int e = tid; e < nelem; e += blockDim.x) {
      const int pw = e % PW, ph = (e / PW) % PH, cl = e / PP;
      const float* pl = gimg + (size_t)(cgrp0 + cl) * HW;
      const int nh = s_th.cnt[ph], nw = s_tw.cnt[pw];
      float best = 0.f, bx = -1.f, by = -1.f;
      if (flags & kFlagOverflow) {
        const float sc_ = L.scale;
        element_direct(pl, H, W, PH, PW, ph, pw, __fmul_rn(__ldg(a.rois + 4 * (size_t)n), sc_),
                       __fmul_rn(__ldg(a.rois + 4 * (size_t)n + 1), sc_),
                       __fmul_rn(__ldg(a.rois + 4 * (size_t)n + 2), sc_),
                       __fmul_rn(__ldg(a.rois + 4 * (size_t)n + 3), sc_), best, bx, by);
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
                                         __fmul_rn(a1, b1), __ldg(pl + hl * W + wl),
                                         __ldg(pl + hh * W + wl), __ldg(pl + hl * W + wr),
                                         __ldg(pl + hh * W + wr));
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

  // ---- per-roi row table: byte offsets of the lo/hi rows (with the per-row 16B shift) ----
  if (tid < PH * kMaxS) {
    const int ph = tid / kMaxS, s = tid % kMaxS;
    if (s_th.cnt[ph] == 2 && s < 2) {
      const int lo = s_th.lo[tid], hi = s_th.hi[tid];
      const int sh_lo = vec ? ((lo * W + wmin) & 3) : 0;
      const int sh_hi = vec ? ((hi * W + wmin) & 3) : 0;
      s_hrow[tid] = HRow{4 * ((lo - hmin) * Wp + sh_lo), 4 * ((hi - hmin) * Wp + sh_hi),
                         s_th.w0[tid], s_th.w1[tid]};
    }
  }
  if (tid == 0) {
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(s_bar);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar0 + 8u * i, 32);        // full[i]: the 32 producer lanes (cp.async completion arrivals)
      mbar_init(bar0 + 8u * (4 + i), NW);  // empty[i]: one arrival per consumer warp
    }
  }
  __syncthreads();  // s_hrow and the mbarriers are visible to every warp

  const unsigned sbase = (unsigned)__cvta_generic_to_shared(s_win);
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kSub = (kPW > 0 && kPW <= 8) ? 2 : 1;  // half-warps per warp that own their own channels
  const int half = (kSub == 2) ? (lane >> 4) : 0;
  const int pw = (kSub == 2 ? (lane & 15) : lane) >> 1, sx = lane & 1;
  const bool lane_on = pw < PW;
  int wcnt = -1, xl = 0, xr = 0;
  float b0 = 0.f, b1 = 0.f, wc = -1.f;
  if (lane_on) {
    wcnt = s_tw.cnt[pw];
    if (wcnt == 2) {
      const int j = pw * kMaxS + sx;
      xl = s_tw.lo[j] - wmin;
      xr = s_tw.hi[j] - wmin;
      b0 = s_tw.w0[j];
      b1 = s_tw.w1[j];
      wc = s_tw.coord[j];
    }
  }
  const uint64_t nz2 = a.negzero2;  // {-0.0f, -0.0f}; a kernel argument on purpose (see header)
  const bool has_empty = (flags & kFlagEmpty) != 0;

  // Everything below is instantiated per (channel stride kCS, channel groups NCG, buffers NBUF).
  auto run = [&](auto cs_tag, auto ncg_tag, auto nbuf_tag) {
    constexpr int kCS = decltype(cs_tag)::value;
    constexpr int NCG = decltype(ncg_tag)::value;
    constexpr int NBUF = decltype(nbuf_tag)::value;
    constexpr int CTILE = NCG * CPT;     // channels per tile
    constexpr int PHS = NW / NCG;        // ph chunks
    constexpr int BUF_BYTES = CTILE * kCS * 4;
    const int ntiles = (cgrp1 - cgrp0) / CTILE;
    const int cg = warp % NCG, pc = warp / NCG;
    const int chunk = (PH + PHS - 1) / PHS;
    const int ph_beg = pc * chunk, ph_end = min(PH, ph_beg + chunk);

    // ---- stage one channel tile into a window buffer with cp.async ----
    // thread = (row slot, 16B chunk) of the window; it walks the tile's channels with constant
    // strides (global: HW floats, shared: kCS floats = an immediate), so the per-copy cost is
    // one 64-bit add + one LDGSTS.
    // The (row, chunk) items of the window are flattened over the CTA's threads, so a pass of 128
    // threads copies 128 useful chunks whatever the window's aspect (an LDGSTS costs the same LSU
    // time with 1 or 32 active lanes).
    const int nch = vec ? ((Wwin + 3 + 3) >> 2) : (Wp >> 2);    // 16B chunks per row (upper bound)
    const int nitems = Hwin * nch;
    const unsigned nch_magic = 0xFFFFFFFFu / (unsigned)nch + 1u; // idx / nch for idx < 2^16
    auto stage = [&](int tile, unsigned buf) {  // executed by the producer warp only
      const float* g0 = gimg + (size_t)(cgrp0 + tile * CTILE) * HW;
      for (int idx = lane; idx < nitems; idx += 32) {
        const int y = (int)__umulhi((unsigned)idx, nch_magic), jchunk = idx - y * nch;
        const int e0 = (hmin + y) * W + wmin;  // first wanted element inside the plane
        unsigned dst = buf + 4u * (unsigned)(y * Wp + jchunk * 4);
        if (vec) {
          const int sh = e0 & 3;
          if (jchunk * 4 < sh + Wwin) {
            const float* src = g0 + (e0 - sh + jchunk * 4);
#pragma unroll
            for (int c = 0; c < CTILE; ++c) {
              cp_async16(dst + c * (kCS * 4), src);
              src += HW;
            }
          }
        } else {
          const float* src = g0 + (e0 + jchunk * 4);
#pragma unroll 1
          for (int c = 0; c < CTILE; ++c) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (jchunk * 4 + e < Wwin) cp_async4(dst + 4u * e, src + e);
            dst += kCS * 4;
            src += HW;
          }
        }
      }
    };

    auto compute = [&](int tile, unsigned buf) {
      // CL = channels per lane: a 7-wide roi fills only 14 lanes of a warp, so its two half-warps
      // take the two halves of the warp's channels (kSub = 2) instead of idling
      constexpr int CL = CPT / kSub;
      const int cbase = cgrp0 + tile * CTILE + cg * CPT + half * CL;
      const unsigned sl = buf + 4u * (unsigned)((cg * CPT + half * CL) * kCS + xl);   // lane's left-tap column
      const unsigned sr = buf + 4u * (unsigned)((cg * CPT + half * CL) * kCS + xr);   // lane's right-tap column
      float RA[CL][2], RB[CL][2];  // 2-row register cache; filled before first use (rowA/B = -1)
      int rowA = -1, rowB = -1;
      const size_t obase = ((size_t)n * C + cbase) * PP + (size_t)ph_beg * PW + pw;
      float* outp = a.out + obase;
      float* axp = kArg ? a.argx + obase : nullptr;
      float* ayp = kArg ? a.argy + obase : nullptr;
      const bool store = lane_on && sx == 0;
      // inference: after the pair exchange both lanes of a pw hold all CL maxima; lane s stores
      // channels [s*CL/2, (s+1)*CL/2) so every store instruction has 28 active lanes
      float* outh = outp + (size_t)(sx * (CL / 2)) * PP;

      // One h-sample: make the register sets hold rows (lo, hi) — whichever set already holds
      // `lo` plays the low row, so nothing is ever moved — then the bilinear values of the CL
      // channels.  The tests are warp-uniform; routing them through a vote lets ptxas emit plain
      // branches instead of divergence bookkeeping.
      auto sample = [&](const int4 hr, float (&v)[CL]) {
        const int olo = hr.x, ohi = hr.y;
        const float a0 = __int_as_float(hr.z), a1 = __int_as_float(hr.w);
        const float wtl = __fmul_rn(a0, b0), wbl = __fmul_rn(a1, b0);
        const float wtr = __fmul_rn(a0, b1), wbr = __fmul_rn(a1, b1);
        const uint64_t wtl2 = pack2(wtl, wtl), wbl2 = pack2(wbl, wbl);
        const uint64_t wtr2 = pack2(wtr, wtr), wbr2 = pack2(wbr, wbr);
        auto step = [&](const float (&Lo)[CL][2], const float (&Hi)[CL][2]) {
#pragma unroll
          for (int k = 0; k < CL; k += 2) {
            // roi_align_v2-inl.h:137-140 for channels k, k+1: ((tl + bl) + tr) + br
            const uint64_t ptl = fma2(wtl2, pack2(Lo[k][0], Lo[k + 1][0]), nz2);
            const uint64_t pbl = fma2(wbl2, pack2(Hi[k][0], Hi[k + 1][0]), nz2);
            const uint64_t ptr = fma2(wtr2, pack2(Lo[k][1], Lo[k + 1][1]), nz2);
            const uint64_t pbr = fma2(wbr2, pack2(Hi[k][1], Hi[k + 1][1]), nz2);
            unpack2(add2(add2(add2(ptl, pbl), ptr), pbr), v[k], v[k + 1]);
          }
        };
#ifndef SDET_RA_NOCACHE
#define SDET_RA_NOCACHE 0
#endif
        if (SDET_RA_NOCACHE && kSub == 2) {
          // narrow outputs: bins are ~2 rows tall, consecutive samples rarely share a row, so the
          // row cache's votes and branches cost more than the reloads they save
          TapLoader<CL, kCS>::run(RA, sl + olo, sr + olo);
          TapLoader<CL, kCS>::run(RB, sl + ohi, sr + ohi);
          step(RA, RB);
        } else if (__all_sync(0xffffffffu, olo == rowA)) {
          if (__any_sync(0xffffffffu, ohi != rowB)) {
            TapLoader<CL, kCS>::run(RB, sl + ohi, sr + ohi);
            rowB = ohi;
          }
          step(RA, RB);
        } else if (__all_sync(0xffffffffu, olo == rowB)) {
          if (__any_sync(0xffffffffu, ohi != rowA)) {
            TapLoader<CL, kCS>::run(RA, sl + ohi, sr + ohi);
            rowA = ohi;
          }
          step(RB, RA);
        } else {
          TapLoader<CL, kCS>::run(RA, sl + olo, sr + olo);
          rowA = olo;
          if (__any_sync(0xffffffffu, ohi != rowB)) {
            TapLoader<CL, kCS>::run(RB, sl + ohi, sr + ohi);
            rowB = ohi;
          }
          step(RA, RB);
        }
      };

      const int4* tab = reinterpret_cast<const int4*>(s_hrow);
      for (int ph = ph_beg; ph < ph_end; ++ph) {
        float v0[CL], v1[CL];
        const bool hvalid = !has_empty || s_th.cnt[ph] == 2;   // warp-uniform
        if (hvalid) {
          const int4 h0 = tab[ph * kMaxS], h1 = tab[ph * kMaxS + 1];
          sample(h0, v0);
          sample(h1, v1);
        } else {
#pragma unroll
          for (int k = 0; k < CL; ++k) v0[k] = v1[k] = -FLT_MAX;  // bin empty along h: zeroed below
        }
        // bins that are empty along an axis (end <= start) pool to 0 / argmax -1
        // (roi_align_v2-inl.h:111-117); only CTAs whose roi has such bins pay for the selects
        const bool zero_out = has_empty && (!hvalid || wcnt < 0);
        // combine: reference order is (h0,w0),(h0,w1),(h1,w0),(h1,w1) with strict '>' starting from
        // -FLT_MAX => maximum, first index on ties, NaN / -inf never win.
        if (kArg) {
          const float hc0 = s_th.coord[ph * kMaxS], hc1 = s_th.coord[ph * kMaxS + 1];
          const float pwc = __shfl_xor_sync(0xffffffffu, wc, 1);
#pragma unroll
          for (int k = 0; k < CL; ++k) {
            float mk = -FLT_MAX;
            int mi = -1;
            if (v0[k] > mk) {
              mk = v0[k];
              mi = sx;
            }
            if (v1[k] > mk) {
              mk = v1[k];
              mi = 2 + sx;
            }
            const float pm = __shfl_xor_sync(0xffffffffu, mk, 1);
            const int pi = __shfl_xor_sync(0xffffffffu, mi, 1);
            const float best = fmaxf(mk, pm);
            int bi = mi;
            float bxc = wc;
            const bool take = (pi >= 0) && (bi < 0 || pm > mk || (pm == mk && pi < bi));
            if (take) {
              bi = pi;
              bxc = pwc;
            }
            if (store) {
              const bool none = zero_out || bi < 0;
              outp[k * PP] = zero_out ? 0.f : best;
              axp[k * PP] = none ? -1.f : bxc;
              ayp[k * PP] = none ? -1.f : ((bi & 2) ? hc1 : hc0);
            }
          }
          axp += PW;
          ayp += PW;
          outp += PW;
        } else {
          float best[CL];
#pragma unroll
          for (int k = 0; k < CL; ++k) {
            const float mk = fmaxf(v0[k], v1[k]);  // fmaxf drops a NaN operand like `v > m` does
            best[k] = max3f(mk, __shfl_xor_sync(0xffffffffu, mk, 1), -FLT_MAX);
          }
          if (lane_on) {
            if (!has_empty) {
#pragma unroll
              for (int k = 0; k < CL / 2; ++k) outh[k * PP] = sx ? best[CL / 2 + k] : best[k];
            } else {
#pragma unroll
              for (int k = 0; k < CL / 2; ++k)
                outh[k * PP] = zero_out ? 0.f : (sx ? best[CL / 2 + k] : best[k]);
            }
          }
          outh += PW;
        }
      }
    };

    // ---- producer / consumer pipeline over the channel tiles of this roi (ring of NBUF buffers) ----
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(s_bar);
#ifdef SDET_RA_ABLATE
    if (tid == 0) atomicAdd(&g_ra_prof[2], (unsigned long long)(clock64() - prof_t0));
#endif
#ifdef SDET_RA_ABLATE
    long long prof_wait = 0, prof_comp = 0;
    const bool trace = ((int)blockIdx.x == g_ra_trace_cta) && blockIdx.y == 0 && lane == 0;
    if (trace && warp == 0) {
      g_ra_trace[0][0] = clock64() - prof_t0;
      g_ra_trace[0][1] = ntiles;
    }
#endif
    if (warp == NW) {
      for (int t = 0; t < ntiles; ++t) {
        const int b = t % NBUF, k = t / NBUF;
#ifdef SDET_RA_ABLATE
        const long long c0 = clock64();
#endif
        if (k > 0) mbar_wait(bar0 + 8u * (4 + b), (unsigned)((k - 1) & 1));  // consumers released the buffer
#ifdef SDET_RA_ABLATE
        prof_wait += clock64() - c0;
        if (trace && t < 256) g_ra_trace[1][2 + 2 * t] = clock64() - prof_t0;
#endif
        stage(t, sbase + (unsigned)b * BUF_BYTES);
        cp_async_mbar_arrive(bar0 + 8u * b);
#ifdef SDET_RA_ABLATE
        if (trace && t < 256) g_ra_trace[1][3 + 2 * t] = clock64() - prof_t0;
#endif
      }
#ifdef SDET_RA_ABLATE
      if (lane == 0) atomicAdd(&g_ra_prof[3], (unsigned long long)prof_wait);
#endif
      return;
    }
    for (int t = 0; t < ntiles; ++t) {
      const int b = t % NBUF, k = t / NBUF;
#ifdef SDET_RA_ABLATE
      const long long c0 = clock64();
#endif
      mbar_wait(bar0 + 8u * b, (unsigned)(k & 1));   // tile t has landed
#ifdef SDET_RA_ABLATE
      const long long c1 = clock64();
#endif
      compute(t, sbase + (unsigned)b * BUF_BYTES);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8u * (4 + b));
#ifdef SDET_RA_ABLATE
      prof_wait += c1 - c0;
      prof_comp += clock64() - c1;
      if (trace && warp == 0 && t < 256) {
        g_ra_trace[0][2 + 2 * t] = c1 - prof_t0;
        g_ra_trace[0][3 + 2 * t] = clock64() - prof_t0;
      }
#endif
    }
#ifdef SDET_RA_ABLATE
    if (lane == 0) {
      atomicAdd(&g_ra_prof[0], (unsigned long long)prof_wait);
      atomicAdd(&g_ra_prof[1], (unsigned long long)prof_comp);
    }
    if (tid == 0) atomicAdd(&g_ra_prof[4], (unsigned long long)(clock64() - prof_t0));
#endif
  };

  // (channel stride, channel groups per tile, ring depth) by window size.  With the producer warp a
  // deeper ring of one-channel-group tiles is a small win (7x7 bench shape 234.6 -> 228.4 us, 14x14
  // target 115.7 -> 113.7 us; profiles/r01_roi_align_7x7_ablation.txt), so it is the default;
  // -DSDET_RA_DEEP=0 restores two channel groups per tile.
  using std::integral_constant;
#ifndef SDET_RA_DEEP
#define SDET_RA_DEEP 2
#endif
  constexpr bool kDeep = (SDET_RA_DEEP == 2) || (SDET_RA_DEEP == 1 && kPW > 0 && kPW <= 8);
  if (mode == 0) {
    if (kDeep) run(integral_constant<int, CS0>{}, integral_constant<int, 1>{}, integral_constant<int, 4>{});
    else run(integral_constant<int, CS0>{}, integral_constant<int, 2>{}, integral_constant<int, 2>{});
  } else if (mode == 1) {
    if (kDeep) run(integral_constant<int, CS1>{}, integral_constant<int, 1>{}, integral_constant<int, 2>{});
    else run(integral_constant<int, CS1>{}, integral_constant<int, 2>{}, integral_constant<int, 1>{});
  } else {
    run(integral_constant<int, CS3>{}, integral_constant<int, 1>{}, integral_constant<int, 1>{});
  }
}


// =============================================================================================
// Window sharing (opt-in prototype, SDET_RA_SHARE=1): rois of one (image, level) whose windows
// overlap are processed by ONE CTA against their union window, staged once per channel tile.
// 7x7 launches are paced by L2 -> SM delivery of per-roi windows (profiles/r01_roi_align_7x7_ablation.txt);
// benchmarks/window_sharing_sim.py estimates 1.9x (uniform random rois) to 5.5x (proposal-like) fewer
// staged cells.  Inference only (no argmax planes), exactly-2-sample rois, 16-byte-stageable levels;
// everything else stays with roi_align_v2_fwd_kernel through `left_order`.
// =============================================================================================
constexpr int kGMax = 6;            // members per group (6 x 1 KB of tables keeps 4 CTAs/SM next to 48 KB windows)
constexpr int kGroupMaxRois = 4096; // per launch (B*N), bounded by the grouping kernel's shared memory
constexpr int kGroupBuckets = 4096; // (segment = level*B + image) x 16 x 32 spatial cells of 16x16 pixels

struct GroupRec {
  int first, count;              // members = sorted_roi[first .. first+count)
  int li, b;
  int hmin, hmax, wmin, wmax;    // union window
};
struct GroupSched {
  GroupRec* groups;
  int* sorted_roi;
  int* grp_order;    // CTA x -> group, costliest first
  int* left_order;   // rois that stay with the per-roi kernel
  int* counters;     // [0] number of groups, [1] number of left-over rois
};

__device__ __forceinline__ int win_plane(const int4 w) {  // rows x 16-byte padded pitch, as the kernel lays it out
  return (w.y - w.x + 1) * ((w.w - w.z + 1 + 6) & ~3);
}

// One CTA.  Counting sort of the rois by (segment, 16x16-cell bucket of the window origin), greedy merge of
// list neighbours inside a segment while the union window fits `cap_cells`, then a largest-first order.
__global__ void __launch_bounds__(1024)
roi_align_group_kernel(const __grid_constant__ RoiAlignArgs a, const PlanRecord* __restrict__ plans,
                       const GroupSched gs, const int total, const int cap_cells) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  int* s_hist = reinterpret_cast<int*>(s_raw);                          // kGroupBuckets + 1 (+ pad)
  int4* s_w4 = reinterpret_cast<int4*>(s_raw + (kGroupBuckets + 4) * 4);   // window per sorted position
  unsigned short* s_n = reinterpret_cast<unsigned short*>(s_w4 + kGroupMaxRois);  // roi per sorted position
  unsigned short* s_bkt = s_n + kGroupMaxRois;                          // bucket per roi (then: rank)
  unsigned short* s_rank = s_bkt + kGroupMaxRois;
  unsigned char* s_seg = reinterpret_cast<unsigned char*>(s_rank + kGroupMaxRois);  // segment per sorted position
  unsigned char* s_start = s_seg + kGroupMaxRois;                       // member count at a group's first position
  __shared__ int s_wsum[32];
  __shared__ int s_cls[64], s_clsbase[64];
  const int tid = threadIdx.x;
  const int B = a.B, nseg = a.num_levels * B;
  for (int i = tid; i <= kGroupBuckets; i += blockDim.x) s_hist[i] = 0;
  if (tid < 64) s_cls[tid] = 0;
  __syncthreads();
  // ---- 1. bucket + rank
  for (int n = tid; n < total; n += blockDim.x) {
    const int4 h0 = __ldg(reinterpret_cast<const int4*>(plans + n));        // li, flags, hmin, hmax
    const int4 h1 = __ldg(reinterpret_cast<const int4*>(plans + n) + 1);    // wmin, wmax, -, -
    const int li = h0.x, flags = h0.y;
    int bucket = kGroupBuckets;  // not shareable
    if (nseg <= kGroupBuckets / 512 && li >= 0 && h0.w >= 0 && h1.y >= 0 &&
        (flags & (kFlagNot2 | kFlagOverflow | kFlagEmpty)) == 0) {
      const Level& L = a.lvl[li];
      const bool vec = (((L.H * L.W) & 3) == 0) && ((reinterpret_cast<uintptr_t>(L.data) & 15) == 0);
      const int4 w = make_int4(h0.z, h0.w, h1.x, h1.y);
      const int hw_ = w.y - w.x + 1, wp = (w.w - w.z + 1 + 6) & ~3;
      if (vec && win_plane(w) <= cap_cells && wp <= 64 && hw_ <= 128) {
        const int seg = li * B + n / a.N;
        bucket = seg * 512 + min(15, w.x >> 4) * 32 + min(31, w.z >> 4);
      }
    }
    s_bkt[n] = (unsigned short)bucket;
    s_rank[n] = (unsigned short)atomicAdd(&s_hist[bucket], 1);
  }
  __syncthreads();
  // ---- 2. exclusive scan of the histogram (4 bins per thread + the spill bin)
  {
    int v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = s_hist[tid * 4 + k]; sum += v[k]; }
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
      int w = s_wsum[tid];
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (tid >= o) w += t;
      }
      s_wsum[tid] = w;
    }
    __syncthreads();
    int run = inc - sum + ((tid >> 5) ? s_wsum[(tid >> 5) - 1] : 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) { s_hist[tid * 4 + k] = run; run += v[k]; }
    if (tid == 1023) s_hist[kGroupBuckets] = run;  // first position of the non-shareable rois
  }
  __syncthreads();
  const int ngp = s_hist[kGroupBuckets];  // shareable rois come first (none when there are too many segments)
  // ---- 3. scatter
  for (int n = tid; n < total; n += blockDim.x) {
    const int bucket = s_bkt[n];
    const int pos = s_hist[bucket] + s_rank[n];
    const int4 h0 = __ldg(reinterpret_cast<const int4*>(plans + n));
    const int4 h1 = __ldg(reinterpret_cast<const int4*>(plans + n) + 1);
    s_n[pos] = (unsigned short)n;
    s_w4[pos] = make_int4(h0.z, h0.w, h1.x, h1.y);
    s_seg[pos] = (unsigned char)(bucket < kGroupBuckets ? bucket / 512 : 255);
    s_start[pos] = 0;
  }
  __syncthreads();
  // ---- 4. greedy merge, one thread per segment (the list of a segment is contiguous)
  for (int p = tid; p < ngp; p += blockDim.x) {
    if (p > 0 && s_seg[p - 1] == s_seg[p]) continue;
    const int seg = s_seg[p];
    int first = p, cnt = 1;
    int4 cur = s_w4[p];
    for (int q = p + 1; q < ngp && s_seg[q] == seg; ++q) {
      const int4 w = s_w4[q];
      const int4 u = make_int4(min(cur.x, w.x), max(cur.y, w.y), min(cur.z, w.z), max(cur.w, w.w));
      if (cnt < kGMax && win_plane(u) <= cap_cells && ((u.w - u.z + 1 + 6) & ~3) <= 64 && (u.y - u.x + 1) <= 128) {
        cur = u;
        ++cnt;
      } else {
        s_start[first] = (unsigned char)cnt;
        s_w4[first] = cur;
        first = q;
        cur = w;
        cnt = 1;
      }
    }
    s_start[first] = (unsigned char)cnt;
    s_w4[first] = cur;
  }
  __syncthreads();
  // ---- 5. number the groups (scan over the start flags), write records, cost classes
  {
    int f[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int p = tid * 4 + k;
      f[k] = (p < ngp && s_start[p] != 0) ? 1 : 0;
      sum += f[k];
    }
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
      int w = s_wsum[tid];
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (tid >= o) w += t;
      }
      s_wsum[tid] = w;
    }
    __syncthreads();
    int g = inc - sum + ((tid >> 5) ? s_wsum[(tid >> 5) - 1] : 0);
    const int ngroups = s_wsum[31];
    if (tid == 0) {
      gs.counters[0] = ngroups;
      gs.counters[1] = total - ngp;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int p = tid * 4 + k;
      if (!f[k]) continue;
      const int4 u = s_w4[p];
      const int seg = s_seg[p], cnt = s_start[p];
      GroupRec r;
      r.first = p; r.count = cnt; r.li = seg / B; r.b = seg - r.li * B;
      r.hmin = u.x; r.hmax = u.y; r.wmin = u.z; r.wmax = u.w;
      gs.groups[g] = r;
      // cost class for the largest-first order: staged cells + per-member arithmetic
      const int cls = min(63, (win_plane(u) + 160 * cnt) / 48);
      s_bkt[g] = (unsigned short)cls;                       // (s_bkt / s_rank are free again)
      s_rank[g] = (unsigned short)atomicAdd(&s_cls[cls], 1);
      ++g;
    }
    __syncthreads();
    if (tid == 0) {
      int acc = 0;
      for (int c = 63; c >= 0; --c) {
        s_clsbase[c] = acc;
        acc += s_cls[c];
      }
    }
    __syncthreads();
    for (int i = tid; i < ngroups; i += blockDim.x) gs.grp_order[s_clsbase[s_bkt[i]] + s_rank[i]] = i;
  }
  for (int p = tid; p < total; p += blockDim.x) {
    gs.sorted_roi[p] = s_n[p];
    if (p >= ngp) gs.left_order[p - ngp] = s_n[p];
  }
}

template <int CPT, int kPH, int kPW, int kCapFloats>
__global__ void __launch_bounds__(kFwdThreads, 4)
roi_align_v2_fwd_grouped_kernel(const __grid_constant__ RoiAlignArgs a, const GroupSched gs, const int tiles) {
  extern __shared__ __align__(16) float s_win[];
  __shared__ __align__(16) int4 s_hrow_g[kGMax][32];    // per member, per (ph, h-sample): off_lo, off_hi, w0, w1
  __shared__ __align__(16) int4 s_wtab_g[kGMax][32];    // per member, per (pw, w-sample): xl, xr, w0, w1
  __shared__ int s_member[kGMax];
  __shared__ __align__(8) unsigned long long s_bar[8];
  constexpr int NW = 4;
  constexpr int PH = kPH, PW = kPW, PP = kPH * kPW;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= __ldg(gs.counters)) return;
  const GroupRec G = gs.groups[__ldg(gs.grp_order + blockIdx.x)];
  const int C = a.C;
  const int cgrp0 = blockIdx.y * tiles * (2 * CPT), cgrp1 = min(C, cgrp0 + tiles * (2 * CPT));
  const Level& L = a.lvl[G.li];
  const int H = L.H, W = L.W, HW = H * W;
  const int hmin = G.hmin, wmin = G.wmin;
  const int Hwin = G.hmax - hmin + 1, Wwin = G.wmax - wmin + 1;
  const float* gimg = L.data + (size_t)G.b * C * HW;
  const int Wp = (Wwin + 3 + 3) & ~3;   // (the grouping kernel only admits 16-byte-stageable levels)
  const int plane = Hwin * Wp;
  constexpr int CS0 = kCapFloats / (4 * CPT), CS1 = kCapFloats / (2 * CPT), CS3 = kCapFloats / CPT;
  const int mode = plane <= CS0 ? 0 : (plane <= CS1 ? 1 : 3);

  // ---- member tables from the plan records
  if (tid < G.count) s_member[tid] = __ldg(gs.sorted_roi + G.first + tid);
  __syncthreads();
  const PlanRecord* plans = static_cast<const PlanRecord*>(a.plans);
  for (int idx = tid; idx < G.count * 64; idx += blockDim.x) {
    const int m = idx >> 6, e = idx & 63;
    const PlanRecord* rec = plans + s_member[m];
    if (e < 32) {
      const int ph = e >> 1, sidx = e & 1;
      if (ph < PH) {
        const int j = ph * kMaxS + sidx;
        const int lo = __ldg(rec->th.lo + j), hi = __ldg(rec->th.hi + j);
        const int sh_lo = (lo * W + wmin) & 3, sh_hi = (hi * W + wmin) & 3;
        s_hrow_g[m][e] = make_int4(4 * ((lo - hmin) * Wp + sh_lo), 4 * ((hi - hmin) * Wp + sh_hi),
                                   __float_as_int(__ldg(rec->th.w0 + j)), __float_as_int(__ldg(rec->th.w1 + j)));
      }
    } else {
      const int pw_ = (e - 32) >> 1, sidx = e & 1;
      if (pw_ < PW) {
        const int j = pw_ * kMaxS + sidx;
        s_wtab_g[m][e - 32] = make_int4(__ldg(rec->tw.lo + j) - wmin, __ldg(rec->tw.hi + j) - wmin,
                                        __float_as_int(__ldg(rec->tw.w0 + j)), __float_as_int(__ldg(rec->tw.w1 + j)));
      }
    }
  }
  if (tid == 0) {
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(s_bar);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar0 + 8u * i, 32);
      mbar_init(bar0 + 8u * (4 + i), NW);
    }
  }
  __syncthreads();

  const unsigned sbase = (unsigned)__cvta_generic_to_shared(s_win);
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kSub = (kPW <= 8) ? 2 : 1;
  const int half = (kSub == 2) ? (lane >> 4) : 0;
  const int pw = (kSub == 2 ? (lane & 15) : lane) >> 1, sx = lane & 1;
  const bool lane_on = pw < PW;
  const uint64_t nz2 = a.negzero2;

  auto run = [&](auto cs_tag, auto ncg_tag, auto nbuf_tag) {
    constexpr int kCS = decltype(cs_tag)::value;
    constexpr int NCG = decltype(ncg_tag)::value;
    constexpr int NBUF = decltype(nbuf_tag)::value;
    constexpr int CTILE = NCG * CPT;
    constexpr int PHS = NW / NCG;
    constexpr int BUF_BYTES = CTILE * kCS * 4;
    constexpr int CL = CPT / kSub;
    const int ntiles = (cgrp1 - cgrp0) / CTILE;
    const int cg = warp % NCG, pc = warp / NCG;
    const int chunk = (PH + PHS - 1) / PHS;
    const int ph_beg = pc * chunk, ph_end = min(PH, ph_beg + chunk);
    const int nch = (Wwin + 3 + 3) >> 2;
    const int nitems = Hwin * nch;
    const unsigned nch_magic = 0xFFFFFFFFu / (unsigned)nch + 1u;
    auto stage = [&](int tile, unsigned buf) {
      const float* g0 = gimg + (size_t)(cgrp0 + tile * CTILE) * HW;
      for (int idx = lane; idx < nitems; idx += 32) {
        const int y = (int)__umulhi((unsigned)idx, nch_magic), jchunk = idx - y * nch;
        const int e0 = (hmin + y) * W + wmin;
        const unsigned dst = buf + 4u * (unsigned)(y * Wp + jchunk * 4);
        const int sh = e0 & 3;
        if (jchunk * 4 < sh + Wwin) {
          const float* src = g0 + (e0 - sh + jchunk * 4);
#pragma unroll
          for (int c = 0; c < CTILE; ++c) {
            cp_async16(dst + c * (kCS * 4), src);
            src += HW;
          }
        }
      }
    };
    // one member roi of the group against the staged tile
    auto compute = [&](int tile, unsigned buf, int m) {
      const int n = s_member[m];
      int xl = 0, xr = 0;
      float b0 = 0.f, b1 = 0.f;
      if (lane_on) {
        const int4 wt = s_wtab_g[m][pw * 2 + sx];
        xl = wt.x; xr = wt.y; b0 = __int_as_float(wt.z); b1 = __int_as_float(wt.w);
      }
      const int cbase = cgrp0 + tile * CTILE + cg * CPT + half * CL;
      const unsigned sl = buf + 4u * (unsigned)((cg * CPT + half * CL) * kCS + xl);
      const unsigned sr = buf + 4u * (unsigned)((cg * CPT + half * CL) * kCS + xr);
      float RA[CL][2], RB[CL][2];
      int rowA = -1, rowB = -1;
      float* outh = a.out + ((size_t)n * C + cbase) * PP + (size_t)ph_beg * PW + pw + (size_t)(sx * (CL / 2)) * PP;
      auto sample = [&](const int4 hr, float (&v)[CL]) {
        const int olo = hr.x, ohi = hr.y;
        const float a0 = __int_as_float(hr.z), a1 = __int_as_float(hr.w);
        const float wtl = __fmul_rn(a0, b0), wbl = __fmul_rn(a1, b0);
        const float wtr = __fmul_rn(a0, b1), wbr = __fmul_rn(a1, b1);
        const uint64_t wtl2 = pack2(wtl, wtl), wbl2 = pack2(wbl, wbl);
        const uint64_t wtr2 = pack2(wtr, wtr), wbr2 = pack2(wbr, wbr);
        auto step = [&](const float (&Lo)[CL][2], const float (&Hi)[CL][2]) {
#pragma unroll
          for (int k = 0; k < CL; k += 2) {
            const uint64_t ptl = fma2(wtl2, pack2(Lo[k][0], Lo[k + 1][0]), nz2);
            const uint64_t pbl = fma2(wbl2, pack2(Hi[k][0], Hi[k + 1][0]), nz2);
            const uint64_t ptr = fma2(wtr2, pack2(Lo[k][1], Lo[k + 1][1]), nz2);
            const uint64_t pbr = fma2(wbr2, pack2(Hi[k][1], Hi[k + 1][1]), nz2);
            unpack2(add2(add2(add2(ptl, pbl), ptr), pbr), v[k], v[k + 1]);
          }
        };
        if (__all_sync(0xffffffffu, olo == rowA)) {
          if (__any_sync(0xffffffffu, ohi != rowB)) {
            TapLoader<CL, kCS>::run(RB, sl + ohi, sr + ohi);
            rowB = ohi;
          }
          step(RA, RB);
        } else if (__all_sync(0xffffffffu, olo == rowB)) {
          if (__any_sync(0xffffffffu, ohi != rowA)) {
            TapLoader<CL, kCS>::run(RA, sl + ohi, sr + ohi);
            rowA = ohi;
          }
          step(RB, RA);
        } else {
          TapLoader<CL, kCS>::run(RA, sl + olo, sr + olo);
          rowA = olo;
          if (__any_sync(0xffffffffu, ohi != rowB)) {
            TapLoader<CL, kCS>::run(RB, sl + ohi, sr + ohi);
            rowB = ohi;
          }
          step(RA, RB);
        }
      };
      const int4* tab = s_hrow_g[m];
      for (int ph = ph_beg; ph < ph_end; ++ph) {
        float v0[CL], v1[CL];
        sample(tab[ph * 2], v0);
        sample(tab[ph * 2 + 1], v1);
        float best[CL];
#pragma unroll
        for (int k = 0; k < CL; ++k) {
          const float mk = fmaxf(v0[k], v1[k]);
          best[k] = max3f(mk, __shfl_xor_sync(0xffffffffu, mk, 1), -FLT_MAX);
        }
        if (lane_on) {
#pragma unroll
          for (int k = 0; k < CL / 2; ++k) outh[k * PP] = sx ? best[CL / 2 + k] : best[k];
        }
        outh += PW;
      }
    };
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(s_bar);
    if (warp == NW) {
      for (int t = 0; t < ntiles; ++t) {
        const int b = t % NBUF, k = t / NBUF;
        if (k > 0) mbar_wait(bar0 + 8u * (4 + b), (unsigned)((k - 1) & 1));
        stage(t, sbase + (unsigned)b * BUF_BYTES);
        cp_async_mbar_arrive(bar0 + 8u * b);
      }
      return;
    }
    for (int t = 0; t < ntiles; ++t) {
      const int b = t % NBUF, k = t / NBUF;
      mbar_wait(bar0 + 8u * b, (unsigned)(k & 1));
      for (int m = 0; m < G.count; ++m) compute(t, sbase + (unsigned)b * BUF_BYTES, m);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8u * (4 + b));
    }
  };
  using std::integral_constant;
  if (mode == 0) run(integral_constant<int, CS0>{}, integral_constant<int, 1>{}, integral_constant<int, 4>{});
  else if (mode == 1) run(integral_constant<int, CS1>{}, integral_constant<int, 1>{}, integral_constant<int, 2>{});
  else run(integral_constant<int, CS3>{}, integral_constant<int, 1>{}, integral_constant<int, 1>{});
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

template <int CPT, int kPH, int kPW, int kCapFloats>
int launch_fwd_t(const RoiAlignArgs& a, cudaStream_t st) {
  static bool configured = false;
  const bool arg = a.argx != nullptr;
  auto k_inf = roi_align_v2_fwd_kernel<CPT, false, kPH, kPW, kCapFloats>;
  auto k_trn = roi_align_v2_fwd_kernel<CPT, true, kPH, kPW, kCapFloats>;
  constexpr int smem_bytes = kCapFloats * 4;
  if (!configured) {
    SDET_CUDA(cudaFuncSetAttribute(k_inf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    SDET_CUDA(cudaFuncSetAttribute(k_trn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  // channel tiles (of 2*CPT channels) per CTA: amortise the per-roi preamble but keep the grid
  // at >= ~12 CTAs per SM
  constexpr int CT = 2 * CPT;
  const int total_tiles = (a.C + CT - 1) / CT;
  const long long jobs = (long long)a.B * a.N * total_tiles;
  int tpc = (int)(jobs / (148 * 12));
  if (tpc < 1) tpc = 1;
  if (tpc > total_tiles) tpc = total_tiles;
#ifdef SDET_RA_ABLATE
  if (const char* e = getenv("SDET_RA_TILES")) tpc = atoi(e);
  {
    const int v = getenv("SDET_RA_ABLATE") ? atoi(getenv("SDET_RA_ABLATE")) : 0;
    cudaMemcpyToSymbolAsync(g_ra_ablate, &v, sizeof(int), 0, cudaMemcpyHostToDevice, st);
  }
  if (tpc < 1) tpc = 1;
  if (tpc > total_tiles) tpc = total_tiles;
#endif
  dim3 grid((unsigned)(a.B * a.N), (unsigned)((total_tiles + tpc - 1) / tpc));
  if (arg)
    k_trn<<<grid, kFwdThreads, smem_bytes, st>>>(a, tpc);
  else
    k_inf<<<grid, kFwdThreads, smem_bytes, st>>>(a, tpc);
  SDET_LAUNCH_CHECK("roi_align_v2_fwd_kernel");
  return SDET_OK;
}

// PlanRecord[B*N] | counts[64] | (bucket, rank)[B*N] | order[B*N] | GroupRec[B*N] | sorted_roi[B*N] |
// grp_order[B*N] | left_order[B*N] | counters[64]      (the group arrays serve SDET_RA_SHARE=1 only)
size_t ws_align16(size_t v) { return (v + 15) & ~(size_t)15; }
size_t plan_workspace_bytes(int B, int N) {
  const size_t total = (size_t)B * N;
  return sizeof(PlanRecord) * total + 256 + 8 * total + ws_align16(4 * total) + sizeof(GroupRec) * total +
         3 * ws_align16(4 * total) + 256;
}

template <int CPT, int kPH, int kPW, int kCapFloats>
int launch_fwd_t(const RoiAlignArgs& a, cudaStream_t st);

// Window-sharing path: grouping kernel, grouped kernel for the shareable rois, per-roi kernel for the rest.
template <int CPT, int kPH, int kPW, int kCapFloats>
int launch_fwd_grouped_t(RoiAlignArgs a, const GroupSched gs, cudaStream_t st) {
  static bool configured = false;
  auto k_grp = roi_align_v2_fwd_grouped_kernel<CPT, kPH, kPW, kCapFloats>;
  constexpr int smem_bytes = kCapFloats * 4;
  constexpr size_t group_smem = (size_t)(kGroupBuckets + 4) * 4 + (size_t)kGroupMaxRois * (16 + 3 * 2 + 2);
  if (!configured) {
    SDET_CUDA(cudaFuncSetAttribute(k_grp, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    SDET_CUDA(cudaFuncSetAttribute(roi_align_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)group_smem));
    configured = true;
  }
  const int total = a.B * a.N;
  roi_align_group_kernel<<<1, 1024, group_smem, st>>>(a, static_cast<const PlanRecord*>(a.plans), gs, total,
                                                       kCapFloats / CPT);
  SDET_LAUNCH_CHECK("roi_align_group_kernel");
  constexpr int CT = 2 * CPT;
  const int total_tiles = (a.C + CT - 1) / CT;
  // a group carries several rois' arithmetic, and there are several times fewer groups than rois (their
  // number is only known on the device): split the channels finer than the per-roi kernel does
  int tpc = (int)((long long)total * total_tiles / (148 * 12)) / 4;
  if (const char* e = getenv("SDET_RA_SHARE_TPC")) tpc = atoi(e);
  if (tpc < 1) tpc = 1;
  if (tpc > total_tiles) tpc = total_tiles;
  dim3 grid((unsigned)total, (unsigned)((total_tiles + tpc - 1) / tpc));
  k_grp<<<grid, kFwdThreads, smem_bytes, st>>>(a, gs, tpc);
  SDET_LAUNCH_CHECK("roi_align_v2_fwd_grouped_kernel");
  a.order = gs.left_order;
  a.order_count = gs.counters + 1;
  return launch_fwd_t<CPT, kPH, kPW, kCapFloats>(a, st);
}

int launch_fwd(RoiAlignArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if ((a.argx == nullptr) != (a.argy == nullptr))
    return sdet::fail(SDET_ERR_INVALID_ARG, "argmax_x and argmax_y must both be given or both NULL");
  a.negzero2 = 0x8000000080000000ull;  // {-0.0f, -0.0f}, see the forward kernel's header
  a.plans = nullptr;
  a.order = nullptr;
  a.order_count = nullptr;
  if (workspace != nullptr && a.PH <= 16 && a.PW <= 16) {
    const size_t total = (size_t)a.B * a.N;
    const size_t need = plan_workspace_bytes(a.B, a.N);
    if (workspace_bytes < need)
      return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
    if (reinterpret_cast<uintptr_t>(workspace) % 16)
      return sdet::fail(SDET_ERR_INVALID_ARG, "workspace must be 16-byte aligned");
    char* w = static_cast<char*>(workspace);
    PlanSched sc{};
    sc.counts = reinterpret_cast<int*>(w + sizeof(PlanRecord) * total);
    sc.slot = reinterpret_cast<int2*>(w + sizeof(PlanRecord) * total + 256);
    sc.order = reinterpret_cast<int*>(w + sizeof(PlanRecord) * total + 256 + 8 * total);
    SDET_CUDA(cudaMemsetAsync(sc.counts, 0, sizeof(int) * kCostBuckets, st));
    roi_align_plan_kernel<<<(unsigned)total, 64, 0, st>>>(a, static_cast<PlanRecord*>(workspace), sc);
    SDET_LAUNCH_CHECK("roi_align_plan_kernel");
    roi_align_order_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sc, (int)total);
    SDET_LAUNCH_CHECK("roi_align_order_kernel");
    a.plans = workspace;
    a.order = sc.order;
    // opt-in prototype: rois with overlapping windows share one staged union window
    static const bool share = getenv("SDET_RA_SHARE") != nullptr && atoi(getenv("SDET_RA_SHARE")) != 0;
    const bool sq7 = a.PH == 7 && a.PW == 7, sq14 = a.PH == 14 && a.PW == 14;
    if (share && a.argx == nullptr && (sq7 || sq14) && total <= (size_t)kGroupMaxRois && a.C % 16 == 0) {
      char* g = w + sizeof(PlanRecord) * total + 256 + 8 * total + ws_align16(4 * total);
      GroupSched gs{};
      gs.groups = reinterpret_cast<GroupRec*>(g); g += sizeof(GroupRec) * total;
      gs.sorted_roi = reinterpret_cast<int*>(g); g += ws_align16(4 * total);
      gs.grp_order = reinterpret_cast<int*>(g); g += ws_align16(4 * total);
      gs.left_order = reinterpret_cast<int*>(g); g += ws_align16(4 * total);
      gs.counters = reinterpret_cast<int*>(g);
      return sq7 ? launch_fwd_grouped_t<8, 7, 7, 12288>(a, gs, st) : launch_fwd_grouped_t<8, 14, 14, 12288>(a, gs, st);
    }
  }
#ifdef SDET_RA_CAP7
  if (a.PH == 7 && a.PW == 7) return launch_fwd_t<8, 7, 7, SDET_RA_CAP7>(a, st);
#endif
  if (a.PH == 7 && a.PW == 7) return launch_fwd_t<8, 7, 7, 12288>(a, st);
  if (a.PH == 14 && a.PW == 14) return launch_fwd_t<8, 14, 14, 12288>(a, st);
  return launch_fwd_t<8, 0, 0, 12288>(a, st);
}

}  // namespace


#ifdef SDET_RA_ABLATE
extern "C" int sdet_debug_ra_prof(unsigned long long* out4, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out4, g_ra_prof, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_ra_prof, z, sizeof(z));
  }
  return 0;
}
#endif

#ifdef SDET_RA_ABLATE
// cta >= 0: arm the timeline for that CTA (call before the launch); out != NULL: copy the 2 x 520 ticks back.
extern "C" int sdet_debug_ra_trace(int cta, long long* out) {
  cudaDeviceSynchronize();
  if (out) cudaMemcpyFromSymbol(out, g_ra_trace, sizeof(long long) * 2 * 520);
  cudaMemcpyToSymbol(g_ra_trace_cta, &cta, sizeof(int));
  return 0;
}
#endif

extern "C" size_t sdet_roi_align_v2_workspace(int B, int N) {
  return (B > 0 && N > 0) ? plan_workspace_bytes(B, N) : 0;
}

extern "C" int sdet_roi_align_v2_forward(const float* data, const float* rois, float* out,
                                         float* argmax_x, float* argmax_y, int B, int N, int C,
                                         int H, int W, int pooled_h, int pooled_w,
                                         float spatial_scale, void* workspace, size_t workspace_bytes,
                                         void* stream) {
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
  return launch_fwd(a, workspace, workspace_bytes, (cudaStream_t)stream);
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
                                             int roi_canonical_level, void* workspace,
                                             size_t workspace_bytes, void* stream) {
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
  return launch_fwd(a, workspace, workspace_bytes, (cudaStream_t)stream);
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
