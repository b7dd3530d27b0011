// Shared definitions of the RoIAlign_v2 kernels (per-roi kernel in roi_align.cu, band-stationary kernel in
// roi_align_band.cu): argument block, per-roi axis tables restating roi_align_v2-inl.h:91-140, PTX wrappers.
#pragma once
#include <cfloat>
#include <climits>
#include <cstdint>

#include "common.cuh"

namespace sdet_ra {

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
__device__ inline void element_direct_strided(const float* __restrict__ plane, const int pstride, int H, int W, int PH, int PW, int ph,
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
                             __fmul_rn(al, be), plane[(size_t)(hl * W + wl) * pstride], plane[(size_t)(hh * W + wl) * pstride],
                             plane[(size_t)(hl * W + wr) * pstride], plane[(size_t)(hh * W + wr) * pstride]);
      if (v > best) {
        best = v;
        bx = w;
        by = h;
      }
      if (++guard > (1 << 20)) return;
    }
  }
}

__device__ inline void element_direct(const float* __restrict__ plane, int H, int W, int PH, int PW, int ph,
                               int pw, float rsw, float rsh, float rew, float reh, float& best,
                               float& bx, float& by) {
  element_direct_strided(plane, 1, H, W, PH, PW, ph, pw, rsw, rsh, rew, reh, best, bx, by);
}

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"   // HW-suspended wait, woken by the arrival
      "@!p bra W_%=;\n"
      "}\n" ::"r"(bar), "r"(parity), "r"(0x989680)
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

// ---------------------------------------------------------------------------------------------
// Band-stationary forward path (roi_align_band.cu).  A feature level is cut into bands of kBandR
// rows; a band plus a halo (kBandRS rows, full width) of a group of channels is staged ONCE in
// shared memory with 1-D bulk TMA copies (rows of one channel plane are contiguous in NCHW) and
// every roi bin whose first sample row lies in the band is computed against it.  Bins of one roi
// that fall into the same band form an *item* (roi, ph0, nph); a *unit* of work is
// (band, <= cap items, channel chunk).
// ---------------------------------------------------------------------------------------------
constexpr int kBandR = 8;            // rows per band: 8*W*4 bytes is a multiple of 16 for every W
constexpr int kBandRS = 12;          // rows staged per band (band + halo): 12*W*4 is a multiple of 16
constexpr int kBandStageFloats = 16384;  // one pipeline stage: 64 KB = CT channels x CS floats
constexpr int kBandStages = 3;
constexpr int kBandMaxBands = 2048;  // bands over all images and levels (else the per-roi kernel is used)
constexpr int kBandItemsPerRoi = 16; // a roi yields at most min(PH, 16) items
constexpr int kBandMaxChunks = 16;   // channel chunks per band (bounds the unit list)
constexpr int kBandCapInfer = 30;    // items per unit (shared-memory table slots), inference
constexpr int kBandCapTrain = 20;    //   ... with argmax outputs (slots carry the sample coordinates)

struct BandGeom {   // per level, filled by the host
  int ok;           // the level can be band-staged (W small enough, planes 8-byte aligned)
  int cs_log2;      // channel stride class in shared memory: CS = 1 << cs_log2 floats >= 12*W + 4
  int nb;           // bands per image
  int base;         // index of the level's first band inside one image's band range
  int oddshift;     // bytes by which odd channels' planes miss a 16-byte boundary (0 or 8)
  int chunk;        // channels per unit on this level
};

struct RoiTab {     // per roi, written by roi_align_plan_kernel for band-able rois
  uint2 lane[32];       // [2*pw+s]: {xl*4 | xr*4 << 16, beta}; x = 0xFFFFFFFF: bin empty along w
  uint2 row[16][2];     // [ph][s]: {off_lo | off_hi << 16 (bytes from the band's first row), alpha}; x = ~0: empty along h
  float wcoord[32];     // sample x coordinate per lane entry      (argmax outputs only)
  float hcoord[16][2];  // sample y coordinate per (ph, s)
};
static_assert(sizeof(RoiTab) == 768, "RoiTab pieces are moved with 16-byte bulk copies");

struct RoiItems {   // per roi: its items, to be scattered into the band lists once the band offsets are known
  int n;            // number of items; -1: the roi is not band-able (it is in left_order)
  int flags;
  int pad[2];
  int4 it[kBandItemsPerRoi];  // {global band, rank inside the band, ph0, nph}
};

struct BandWork {   // device pointers into the workspace
  int* ctr;         // [0] unit queue head, [1] number of units, [2] leftover rois, [3] total items
  int* band_cnt;    // items per band
  int* band_rows;   // sum of nph per band (cost estimate)
  int* band_off;    // exclusive prefix of band_cnt
  RoiTab* tabs;
  RoiItems* ritems;
  int2* band_items; // {roi, ph0 | nph << 8 | flags << 16}, grouped by band
  int4* units;      // {global band, first item, nitems, c0 | nch << 16}, most expensive first
  int* left_order;  // rois the band kernel does not take
  int max_units;
};

struct BandArgs {
  BandGeom geom[SDET_MAX_LEVELS];
  int bands_per_image;
  int num_bands;    // B * bands_per_image
  int cap;          // items per unit
  BandWork w;
};

size_t band_workspace_bytes(size_t total_rois);
bool band_setup(const RoiAlignArgs& a, void* ws, BandArgs& ba);
int band_launch(const RoiAlignArgs& a, const BandArgs& ba, PlanRecord* plans, cudaStream_t st);
// channels-last path (roi_align_cl.cu)
size_t cl_scratch_bytes(int B, int C, const int* H, const int* W, int num_levels);
int cl_transpose(RoiAlignArgs& a, void* scratch, cudaStream_t st);
int cl_launch(const RoiAlignArgs& a, const PlanRecord* plans, const int* order, cudaStream_t st);

}  // namespace sdet_ra
