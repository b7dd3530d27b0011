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
#include <cstdlib>
#include <type_traits>

#include "roi_align_common.cuh"

using namespace sdet_ra;

namespace {


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
    if (warp == NW) {
      for (int t = 0; t < ntiles; ++t) {
        const int b = t % NBUF, k = t / NBUF;
        if (k > 0) mbar_wait(bar0 + 8u * (4 + b), (unsigned)((k - 1) & 1));  // consumers released the buffer
        stage(t, sbase + (unsigned)b * BUF_BYTES);
        cp_async_mbar_arrive(bar0 + 8u * b);
      }
      return;
    }
    for (int t = 0; t < ntiles; ++t) {
      const int b = t % NBUF, k = t / NBUF;
      mbar_wait(bar0 + 8u * b, (unsigned)(k & 1));   // tile t has landed
      compute(t, sbase + (unsigned)b * BUF_BYTES);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8u * (4 + b));
    }
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
int launch_fwd_t(const RoiAlignArgs& a, int max_rois, cudaStream_t st) {
  const bool arg = a.argx != nullptr;
  auto k_inf = roi_align_v2_fwd_kernel<CPT, false, kPH, kPW, kCapFloats>;
  auto k_trn = roi_align_v2_fwd_kernel<CPT, true, kPH, kPW, kCapFloats>;
  constexpr int smem_bytes = kCapFloats * 4;
  // per device and cheap: set on every launch rather than cached in a process-wide static
  SDET_CUDA(cudaFuncSetAttribute(arg ? k_trn : k_inf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  // channel tiles (of 2*CPT channels) per CTA: amortise the per-roi preamble but keep the grid
  // at >= ~12 CTAs per SM
  constexpr int CT = 2 * CPT;
  const int total_tiles = (a.C + CT - 1) / CT;
  const long long jobs = (long long)a.B * a.N * total_tiles;
  int tpc = (int)(jobs / (148 * 12));
  if (tpc < 1) tpc = 1;
  if (tpc > total_tiles) tpc = total_tiles;
  dim3 grid((unsigned)max_rois, (unsigned)((total_tiles + tpc - 1) / tpc));
  if (arg)
    k_trn<<<grid, kFwdThreads, smem_bytes, st>>>(a, tpc);
  else
    k_inf<<<grid, kFwdThreads, smem_bytes, st>>>(a, tpc);
  SDET_LAUNCH_CHECK("roi_align_v2_fwd_kernel");
  return SDET_OK;
}

int launch_per_roi(const RoiAlignArgs& a, int max_rois, cudaStream_t st) {
  if (a.PH == 7 && a.PW == 7) return launch_fwd_t<8, 7, 7, 12288>(a, max_rois, st);
  if (a.PH == 14 && a.PW == 14) return launch_fwd_t<8, 14, 14, 12288>(a, max_rois, st);
  return launch_fwd_t<8, 0, 0, 12288>(a, max_rois, st);
}

// Workspace: PlanRecord[B*N] | counts[64] | (bucket, rank)[B*N] | order[B*N] | band section (roi_align_band.cu)
size_t ws_align16(size_t v) { return (v + 15) & ~(size_t)15; }
size_t per_roi_workspace_bytes(size_t total) {
  return sizeof(PlanRecord) * total + 256 + 8 * total + ws_align16(4 * total);
}
size_t plan_workspace_bytes(int B, int N) {
  const size_t total = (size_t)B * N;
  return per_roi_workspace_bytes(total) + band_workspace_bytes(total);
}

// mode: 0 = automatic, 1 = per-roi kernel only, 2 = band-stationary kernel (+ per-roi leftovers),
//       3 = channels-last kernel (features re-laid to NHWC in the scratch that follows the plan workspace).
// Automatic = the planned per-roi kernel: with NCHW features it is the fastest of the three on both measured shapes
// (profiles/r02_roialign_paths.md); the channels-last kernel wins when the features ARE channels-last, which is the
// separate entry point sdet_fpn_roi_align_v2_forward_nhwc.  2 and 3 stay selectable for measurement.
// g_last_path reports what ran: 0 inline per-roi, 1 planned per-roi, 2 band-stationary, 3 channels-last.
thread_local int g_last_path = 0;

int plan_and_order(const RoiAlignArgs& a, char* w, size_t total, PlanSched& sc, cudaStream_t st) {
  sc.counts = reinterpret_cast<int*>(w + sizeof(PlanRecord) * total);
  sc.slot = reinterpret_cast<int2*>(w + sizeof(PlanRecord) * total + 256);
  sc.order = reinterpret_cast<int*>(w + sizeof(PlanRecord) * total + 256 + 8 * total);
  SDET_CUDA(cudaMemsetAsync(sc.counts, 0, sizeof(int) * kCostBuckets, st));
  roi_align_plan_kernel<<<(unsigned)total, 64, 0, st>>>(a, reinterpret_cast<PlanRecord*>(w), sc);
  SDET_LAUNCH_CHECK("roi_align_plan_kernel");
  roi_align_order_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sc, (int)total);
  SDET_LAUNCH_CHECK("roi_align_order_kernel");
  return SDET_OK;
}

int launch_fwd(RoiAlignArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st, int mode = 0) {
  if ((a.argx == nullptr) != (a.argy == nullptr))
    return sdet::fail(SDET_ERR_INVALID_ARG, "argmax_x and argmax_y must both be given or both NULL");
  a.negzero2 = 0x8000000080000000ull;  // {-0.0f, -0.0f}, see the forward kernel's header
  a.plans = nullptr;
  a.order = nullptr;
  a.order_count = nullptr;
  g_last_path = 0;
  const int total_i = a.B * a.N;
  if (workspace != nullptr && a.PH <= 16 && a.PW <= 16) {
    const size_t total = (size_t)total_i;
    const size_t need = plan_workspace_bytes(a.B, a.N);
    if (workspace_bytes < need)
      return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
    if (reinterpret_cast<uintptr_t>(workspace) % 16)
      return sdet::fail(SDET_ERR_INVALID_ARG, "workspace must be 16-byte aligned");
    char* w = static_cast<char*>(workspace);
    // channels-last: needs the NHWC scratch behind the plan/band sections, no argmax planes, an even channel count
    {
      int Hs[SDET_MAX_LEVELS], Ws[SDET_MAX_LEVELS];
      for (int l = 0; l < a.num_levels; ++l) { Hs[l] = a.lvl[l].H; Ws[l] = a.lvl[l].W; }
      const size_t scratch = cl_scratch_bytes(a.B, a.C, Hs, Ws, a.num_levels);
      const bool cl_ok = a.argx == nullptr && (a.C & 1) == 0 && workspace_bytes >= need + scratch;
      if (mode == 3 && !cl_ok)
        return sdet::fail(SDET_ERR_WORKSPACE, "channels-last path: needs no argmax planes, even C and a workspace of "
                          "sdet_fpn_roi_align_v2_workspace() = %zu bytes", need + scratch);
      if (cl_ok && mode == 3) {
        PlanSched sc{};
        if (int rc = plan_and_order(a, w, total, sc, st)) return rc;
        RoiAlignArgs t = a;  // levels re-pointed at the NHWC copies
        if (int rc = cl_transpose(t, w + need, st)) return rc;
        g_last_path = 3;
        return cl_launch(t, static_cast<const PlanRecord*>(workspace), sc.order, st);
      }
    }
    BandArgs ba{};
    if (mode == 2 && band_setup(a, w + per_roi_workspace_bytes(total), ba)) {
      if (int rc = band_launch(a, ba, static_cast<PlanRecord*>(workspace), st)) return rc;
      g_last_path = 2;
      a.plans = workspace;
      a.order = ba.w.left_order;
      a.order_count = ba.w.ctr + 2;
      return launch_per_roi(a, total_i, st);  // CTAs beyond the leftover count exit at once
    }
    PlanSched sc{};
    if (int rc = plan_and_order(a, w, total, sc, st)) return rc;
    a.plans = workspace;
    a.order = sc.order;
    g_last_path = 1;
  }
  return launch_per_roi(a, total_i, st);
}

}  // namespace




extern "C" size_t sdet_roi_align_v2_workspace(int B, int N) {
  return (B > 0 && N > 0) ? plan_workspace_bytes(B, N) : 0;
}

extern "C" int sdet_roi_align_v2_forward_ex(const float* data, const float* rois, float* out,
                                            float* argmax_x, float* argmax_y, int B, int N, int C,
                                            int H, int W, int pooled_h, int pooled_w,
                                            float spatial_scale, void* workspace, size_t workspace_bytes,
                                            void* stream, int path, int* path_used) {
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
  SDET_REQUIRE(path >= 0 && path <= 3, "path must be 0 (automatic), 1 (per-roi), 2 (band-stationary) or 3 (channels-last)");
  const int rc = launch_fwd(a, workspace, workspace_bytes, (cudaStream_t)stream, path);
  if (path_used) *path_used = g_last_path;
  return rc;
}

extern "C" int sdet_roi_align_v2_forward(const float* data, const float* rois, float* out,
                                         float* argmax_x, float* argmax_y, int B, int N, int C,
                                         int H, int W, int pooled_h, int pooled_w,
                                         float spatial_scale, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  return sdet_roi_align_v2_forward_ex(data, rois, out, argmax_x, argmax_y, B, N, C, H, W, pooled_h, pooled_w,
                                      spatial_scale, workspace, workspace_bytes, stream, 0, nullptr);
}

static int ilog2_exact(int v) {
  if (v <= 0 || (v & (v - 1))) return -1;
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

extern "C" int sdet_fpn_roi_align_v2_forward_ex(const float* const* feats, const int* H, const int* W,
                                                const int* strides, int num_levels, const float* rois,
                                                float* out, float* argmax_x, float* argmax_y,
                                                int32_t* levels_out, int B, int N, int C, int pooled_h,
                                                int pooled_w, int roi_canonical_scale,
                                                int roi_canonical_level, void* workspace,
                                                size_t workspace_bytes, void* stream, int path,
                                                int* path_used) {
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
  SDET_REQUIRE(path >= 0 && path <= 3, "path must be 0 (automatic), 1 (per-roi), 2 (band-stationary) or 3 (channels-last)");
  const int rc = launch_fwd(a, workspace, workspace_bytes, (cudaStream_t)stream, path);
  if (path_used) *path_used = g_last_path;
  return rc;
}

// Workspace that also holds the NHWC re-layout of the feature maps (enables the channels-last kernel).
extern "C" size_t sdet_fpn_roi_align_v2_workspace(int B, int N, int C, const int* H, const int* W, int num_levels) {
  if (B <= 0 || N <= 0 || C <= 0 || !H || !W || num_levels < 1 || num_levels > SDET_MAX_LEVELS) return 0;
  return plan_workspace_bytes(B, N) + cl_scratch_bytes(B, C, H, W, num_levels);
}

// Features already channels-last (B, H_l, W_l, C): no re-layout pass.  workspace >= sdet_roi_align_v2_workspace(B, N).
extern "C" int sdet_fpn_roi_align_v2_forward_nhwc(const float* const* feats_nhwc, const int* H, const int* W,
                                                  const int* strides, int num_levels, const float* rois, float* out,
                                                  int32_t* levels_out, int B, int N, int C, int pooled_h, int pooled_w,
                                                  int roi_canonical_scale, int roi_canonical_level, void* workspace,
                                                  size_t workspace_bytes, void* stream) {
  if (int rc = check_common(B, N, C, pooled_h, pooled_w)) return rc;
  SDET_REQUIRE(feats_nhwc && H && W && strides && rois && out && workspace, "NULL argument");
  SDET_REQUIRE(num_levels >= 1 && num_levels <= SDET_MAX_LEVELS, "num_levels must be in [1, %d]", SDET_MAX_LEVELS);
  SDET_REQUIRE(roi_canonical_scale > 0, "roi_canonical_scale must be > 0");
  SDET_REQUIRE(pooled_h <= 16 && pooled_w <= 16 && (C & 1) == 0, "channels-last path: pooled_size <= 16, even C");
  RoiAlignArgs a{};
  int smin = INT_MAX, smax = 0;
  for (int l = 0; l < num_levels; ++l) {
    const int lg = ilog2_exact(strides[l]);
    if (lg < 0) return sdet::fail(SDET_ERR_UNSUPPORTED, "stride %d is not a power of two", strides[l]);
    SDET_REQUIRE(feats_nhwc[l] && H[l] > 0 && W[l] > 0, "level %d: bad feature pointer / shape", l);
    SDET_REQUIRE((reinterpret_cast<uintptr_t>(feats_nhwc[l]) & 7) == 0, "level %d: features must be 8-byte aligned", l);
    a.lvl[l] = Level{feats_nhwc[l], nullptr, H[l], W[l], 1.0f / (float)strides[l], lg};
    smin = strides[l] < smin ? strides[l] : smin;
    smax = strides[l] > smax ? strides[l] : smax;
  }
  a.num_levels = num_levels;
  a.fpn = num_levels > 1 ? 1 : 0;
  a.scale0 = (float)roi_canonical_scale;
  a.lvl0 = (float)roi_canonical_level;
  a.k_min = (float)ilog2_exact(smin);
  a.k_max = (float)ilog2_exact(smax);
  a.rois = rois; a.out = out; a.levels_out = levels_out;
  a.B = B; a.N = N; a.C = C; a.PH = pooled_h; a.PW = pooled_w;
  a.negzero2 = 0x8000000080000000ull;
  const size_t total = (size_t)B * N;
  if (workspace_bytes < per_roi_workspace_bytes(total))
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", per_roi_workspace_bytes(total));
  SDET_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "workspace must be 16-byte aligned");
  PlanSched sc{};
  if (int rc = plan_and_order(a, static_cast<char*>(workspace), total, sc, (cudaStream_t)stream)) return rc;
  return cl_launch(a, static_cast<const PlanRecord*>(workspace), sc.order, (cudaStream_t)stream);
}

extern "C" int sdet_fpn_roi_align_v2_forward(const float* const* feats, const int* H, const int* W,
                                             const int* strides, int num_levels, const float* rois,
                                             float* out, float* argmax_x, float* argmax_y,
                                             int32_t* levels_out, int B, int N, int C, int pooled_h,
                                             int pooled_w, int roi_canonical_scale,
                                             int roi_canonical_level, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  return sdet_fpn_roi_align_v2_forward_ex(feats, H, W, strides, num_levels, rois, out, argmax_x, argmax_y,
                                          levels_out, B, N, C, pooled_h, pooled_w, roi_canonical_scale,
                                          roi_canonical_level, workspace, workspace_bytes, stream, 0, nullptr);
}

// assign_layer_fpn (models/FPN/assign_layer_fpn.py:17-40, CustomOp 'assign_layer_fpn') as its own operator: one
// output per level, the roi where it is assigned and zeros elsewhere.
struct FpnAssignArgs {
  float* out[SDET_MAX_LEVELS];
  int stride_log2[SDET_MAX_LEVELS];
  int num_levels;
  float scale0, lvl0, k_min, k_max;
};
__global__ void __launch_bounds__(256)
fpn_assign_kernel(const float4* __restrict__ rois, const int total, const __grid_constant__ FpnAssignArgs a,
                  int32_t* __restrict__ levels) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float4 r = __ldg(rois + i);
  const int t = fpn_level(r.x, r.y, r.z, r.w, a.scale0, a.lvl0, a.k_min, a.k_max);
  int li = -1;
  for (int l = 0; l < a.num_levels; ++l) {
    const bool mine = a.stride_log2[l] == t;
    if (mine) li = l;
    if (a.out[l]) reinterpret_cast<float4*>(a.out[l])[i] = mine ? r : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (levels) levels[i] = li;
}

extern "C" int sdet_fpn_assign(const float* rois, int total_rois, const int* strides, int num_levels,
                               int roi_canonical_scale, int roi_canonical_level, float* const* out_rois,
                               int32_t* levels_out, void* stream) {
  SDET_REQUIRE(rois && strides && total_rois > 0, "NULL argument / no rois");
  SDET_REQUIRE(num_levels >= 1 && num_levels <= SDET_MAX_LEVELS, "num_levels must be in [1, %d]", SDET_MAX_LEVELS);
  SDET_REQUIRE(roi_canonical_scale > 0, "roi_canonical_scale must be > 0");
  SDET_REQUIRE((reinterpret_cast<uintptr_t>(rois) & 15) == 0, "rois must be 16-byte aligned");
  FpnAssignArgs a{};
  int smin = INT_MAX, smax = 0;
  for (int l = 0; l < num_levels; ++l) {
    const int lg = ilog2_exact(strides[l]);
    if (lg < 0) return sdet::fail(SDET_ERR_UNSUPPORTED, "stride %d is not a power of two", strides[l]);
    a.stride_log2[l] = lg;
    a.out[l] = out_rois ? out_rois[l] : nullptr;
    SDET_REQUIRE((reinterpret_cast<uintptr_t>(a.out[l]) & 15) == 0, "out_rois[%d] must be 16-byte aligned", l);
    smin = strides[l] < smin ? strides[l] : smin;
    smax = strides[l] > smax ? strides[l] : smax;
  }
  a.num_levels = num_levels;
  a.scale0 = (float)roi_canonical_scale;
  a.lvl0 = (float)roi_canonical_level;
  a.k_min = (float)ilog2_exact(smin);
  a.k_max = (float)ilog2_exact(smax);
  fpn_assign_kernel<<<(unsigned)((total_rois + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(rois), total_rois, a, levels_out);
  SDET_LAUNCH_CHECK("fpn_assign_kernel");
  return SDET_OK;
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
