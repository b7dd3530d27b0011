// Channels-last _contrib_ROIAlign_v2 forward for sm_100a.
//
// Reference semantics: operator_cxx/contrib/roi_align_v2-inl.h:61-153, models/FPN/assign_layer_fpn.py:17-40.
// In NCHW a bilinear tap of one channel is 4 bytes and the lanes of a warp have to be spread over output columns:
// every lane then carries its own addresses, weights and row bookkeeping, and the instruction stream is 2/3
// overhead (profiles/r02_band_*).  With the feature map in NHWC a tap of 64 channels is 256 contiguous bytes:
//
//   lane = 2 channels (one packed fp32x2 register pair), warp = 64 channels of one output bin;
//   every coordinate, weight and address is warp-uniform (uniform datapath / broadcast loads from the roi's table);
//   the 2x2 samples of a bin read a patch of at most 3x3 distinct pixels (rows and columns of neighbouring samples
//   coincide), each one LDG.64 per lane = one fully coalesced 256-byte request, no shared-memory staging at all;
//   all four samples of a bin live in one lane, so the max needs no shuffle; the PW outputs of a (roi, ph) row are
//   collected in a small shared-memory tile and written out as contiguous runs of the NCHW-ordered output.
//
// The arithmetic per (sample, channel) is the band / per-roi kernels' - products rounded separately through
// fma.rn.f32x2(w, x, -0.0), sums in the reference's order - so `out` is bit-identical.
//
// Features given in NCHW (the operator's contract) are first re-laid into a caller-provided NHWC scratch by
// nchw_to_nhwc_kernel (all levels in one launch); callers whose producer already emits channels-last (tensor-core
// convolutions do) pass NHWC features directly through sdet_fpn_roi_align_v2_forward_nhwc and skip that pass.
#include <algorithm>

#include "roi_align_common.cuh"

using namespace sdet_ra;

namespace sdet_ra {

namespace {

constexpr int kClWarps = 7;       // 7x7 and 14x14 rows divide evenly among 7 warps
constexpr int kClThreads = kClWarps * 32;
#ifndef SDET_CL_MINB
#define SDET_CL_MINB 4
#endif
constexpr int kClGroup = 64;       // channels per warp pass (2 per lane)


// ---------------------------------------------------------------------------------------------
// NCHW -> NHWC, all levels and images in one launch: 32 (pixels) x 32 (channels) tiles through shared memory
// ---------------------------------------------------------------------------------------------
struct TrArgs {
  const float* src[SDET_MAX_LEVELS];
  float* dst[SDET_MAX_LEVELS];
  int hw[SDET_MAX_LEVELS];
  int tile0[SDET_MAX_LEVELS + 1];  // first pixel-tile index of each level
  int num_levels, C, B;
};

__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const __grid_constant__ TrArgs a) {
  __shared__ float tile[32][33];
  int l = 0;
  while (l + 1 < a.num_levels && (int)blockIdx.x >= a.tile0[l + 1]) ++l;
  const int p0 = ((int)blockIdx.x - a.tile0[l]) * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
  const int HW = a.hw[l];
  const float* src = a.src[l] + (size_t)b * a.C * HW;
  float* dst = a.dst[l] + (size_t)b * a.C * HW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, p = p0 + tx;
    tile[ty + 8 * k][tx] = (c < a.C && p < HW) ? __ldg(src + (size_t)c * HW + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = p0 + ty + 8 * k, c = c0 + tx;
    if (p < HW && c < a.C) dst[(size_t)p * a.C + c] = tile[tx][ty + 8 * k];
  }
}

// ---------------------------------------------------------------------------------------------
// gather kernel: CTA = (roi, 64-channel group); warps take ph rows
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t ldg2(const float* p) {
  uint64_t v;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(*reinterpret_cast<float*>(&v)),
               "=f"(*(reinterpret_cast<float*>(&v) + 1)) : "l"(p));
  return v;
}

// one sample: ((wtl*tl + wbl*bl) + wtr*tr) + wbr*br for the lane's two channels (roi_align_v2-inl.h:137-140)
__device__ __forceinline__ uint64_t bilin2(const float a0, const float a1, const float b0, const float b1,
                                           const uint64_t tl, const uint64_t bl, const uint64_t tr, const uint64_t br,
                                           const uint64_t nz2) {
  const float wtl = __fmul_rn(a0, b0), wbl = __fmul_rn(a1, b0), wtr = __fmul_rn(a0, b1), wbr = __fmul_rn(a1, b1);
  const uint64_t ptl = fma2(pack2(wtl, wtl), tl, nz2), pbl = fma2(pack2(wbl, wbl), bl, nz2);
  const uint64_t ptr = fma2(pack2(wtr, wtr), tr, nz2), pbr = fma2(pack2(wbr, wbr), br, nz2);
  return add2(add2(add2(ptl, pbl), ptr), pbr);
}

__device__ __forceinline__ void max_update(float& m0, float& m1, const uint64_t v) {
  float x, y;
  unpack2(v, x, y);
  m0 = fmaxf(m0, x);  // fmaxf drops a NaN operand like `value > maxval` does
  m1 = fmaxf(m1, y);
}

struct ClAux {
  const PlanRecord* plans;
  const int* order;  // CTA x -> roi (largest window first) or nullptr
};

// Per output column / row of the roi, built once per CTA from the plan record: everything the inner loop needs about
// the two samples of that axis in 32 bytes (two broadcast LDS.128).  The four pixels an axis touches - low and
// high neighbour of sample 0 and of sample 1 - fall into one of three patterns:
//   class 0: both samples between the same two adjacent pixels          -> 2 distinct pixels  {0, 1}
//   class 1: sample 1 one pixel further                                 -> 3 distinct pixels  {0, 1, 2}
//   class 2: anything else (further apart, clamped at the border, ...)  -> 4 loads at lo0 + {0, d1, d2, d3}
// so a bin reads a patch of (2 + cy) x (2 + cx) pixels instead of 16.  cls = -1: not exactly two samples (or an
// offset that does not fit the packing): the bin takes the table walk.
struct __align__(16) AxisRec {
  int lo0;      // low pixel of sample 0
  int cls;      // 0, 1, 2 or -1
  int pack;     // class 2: (hi0 - lo0) | (lo1 - lo0) << 10 | (hi1 - lo0) << 20
  int cnt;      // samples in the bin along this axis (-1: empty)
  float w00, w01, w10, w11;  // {1 - frac, frac} of sample 0 and of sample 1
};

__device__ __forceinline__ AxisRec make_axis_rec(const AxisTab<16>& t, const int p) {  // t in global memory
  AxisRec r;
  const int k = p * kMaxS;
  r.cnt = __ldg(&t.cnt[p]);
  const int lo0 = __ldg(&t.lo[k]), hi0 = __ldg(&t.hi[k]), lo1 = __ldg(&t.lo[k + 1]), hi1 = __ldg(&t.hi[k + 1]);
  r.lo0 = lo0;
  r.w00 = __ldg(&t.w0[k]); r.w01 = __ldg(&t.w1[k]); r.w10 = __ldg(&t.w0[k + 1]); r.w11 = __ldg(&t.w1[k + 1]);
  const int d1 = hi0 - lo0, d2 = lo1 - lo0, d3 = hi1 - lo0;
  r.pack = d1 | (d2 << 10) | (d3 << 20);
  if (r.cnt != 2 || ((d1 | d2 | d3) & ~1023)) r.cls = -1;
  else if (d1 == 1 && d2 == 0 && d3 == 1) r.cls = 0;
  else if (d1 == 1 && d2 == 1 && d3 == 2) r.cls = 1;
  else r.cls = 2;
  return r;
}

// One bin, kG channel groups of 64.  Every load is one coalesced 256-byte request of the warp.  The 16 corner
// weights are the same for every channel: formed once per bin, amortised over the groups.
template <int CX, int CY, int kG>
__device__ __forceinline__ void patch_bin(const float* __restrict__ q, const int rstep, const int C, const AxisRec& hr,
                                          const AxisRec& wr, const uint64_t nz2, float (&m)[2 * kG]) {
  float wt[2][2][4];
  const float ha[2][2] = {{hr.w00, hr.w01}, {hr.w10, hr.w11}}, wb[2][2] = {{wr.w00, wr.w01}, {wr.w10, wr.w11}};
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      // (top-left, bottom-left) = (1-fy, fy) * (1-fx) and (top-right, bottom-right) = (1-fy, fy) * fx
      // (roi_align_v2-inl.h:133-136), two products per packed instruction, each rounded like a lone multiply
      const uint64_t hp = pack2(ha[i][0], ha[i][1]);
      unpack2(fma2(hp, pack2(wb[j][0], wb[j][0]), nz2), wt[i][j][0], wt[i][j][1]);
      unpack2(fma2(hp, pack2(wb[j][1], wb[j][1]), nz2), wt[i][j][2], wt[i][j][3]);
    }
  int ro[2 + CY], co[2 + CX];
  ro[0] = 0;
  co[0] = 0;
  if (CY == 2) {
    ro[1] = (hr.pack & 1023) * rstep; ro[2] = ((hr.pack >> 10) & 1023) * rstep; ro[3 % (2 + CY)] = (hr.pack >> 20) * rstep;
  } else {
#pragma unroll
    for (int r = 1; r < 2 + CY; ++r) ro[r] = r * rstep;
  }
  if (CX == 2) {
    co[1] = (wr.pack & 1023) * C; co[2] = ((wr.pack >> 10) & 1023) * C; co[3 % (2 + CX)] = (wr.pack >> 20) * C;
  } else {
#pragma unroll
    for (int c = 1; c < 2 + CX; ++c) co[c] = c * C;
  }
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    uint64_t P[2 + CY][2 + CX];
#pragma unroll
    for (int r = 0; r < 2 + CY; ++r)
#pragma unroll
      for (int c = 0; c < 2 + CX; ++c) P[r][c] = ldg2(q + g * kClGroup + ro[r] + co[c]);
    float m0 = -FLT_MAX, m1 = -FLT_MAX;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = i * CY, c = j * CX;  // index of the sample's low pixel in the patch
        const float* w = wt[i][j];
        // ((wtl*tl + wbl*bl) + wtr*tr) + wbr*br, products rounded on their own (roi_align_v2-inl.h:137-140)
        const uint64_t ptl = fma2(pack2(w[0], w[0]), P[r][c], nz2), pbl = fma2(pack2(w[1], w[1]), P[r + 1][c], nz2);
        const uint64_t ptr = fma2(pack2(w[2], w[2]), P[r][c + 1], nz2);
        const uint64_t pbr = fma2(pack2(w[3], w[3]), P[r + 1][c + 1], nz2);
        max_update(m0, m1, add2(add2(add2(ptl, pbl), ptr), pbr));
      }
    m[2 * g] = m0;
    m[2 * g + 1] = m1;
  }
}

__device__ __forceinline__ void sts_f32(const uint32_t addr, const float v) {
  // no "memory" clobber: the tile is read back only after a __syncthreads(), and the clobber would pin the loop's
  // table loads behind these stores
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v));
}

// Everything the fast path does not take - an axis without samples (0 by roi_align_v2-inl.h:111-117), sample counts
// other than two, tables that overflowed - for one bin and all kG groups, stored straight into the tile.

template <int kG>
__device__ __forceinline__ void slow_bin(const RoiAlignArgs& a, const PlanRecord& rec, const int n, const int li,
                                      const int flags, const int c0, const int C, const int ph, const int pw,
                                      const int nh, const int nw, const uint32_t sa, const uint32_t st4) {
  const int lane = threadIdx.x & 31;
  const Level& L = a.lvl[li];
  const int W = L.W;
  const float* img = L.data + (size_t)(n / a.N) * L.H * W * C;
  for (int g = 0; g < kG; ++g) {
    const int c = c0 + g * kClGroup + 2 * lane;
    const float* base = img + (c < C ? c : 0);  // lanes past the last channel read channel 0; nobody reads their slot
    float m0 = -FLT_MAX, m1 = -FLT_MAX;
    if (nh < 0 || nw < 0) {
      m0 = m1 = 0.f;
    } else if ((flags & kFlagOverflow) == 0) {
      // general table walk: any sample count up to kMaxS, clamped / coincident corners
      const int hb = ph * kMaxS, wb = pw * kMaxS;
      for (int i = 0; i < nh; ++i) {
        const int hl = __ldg(&rec.th.lo[hb + i]), hh = __ldg(&rec.th.hi[hb + i]);
        const float a0 = __ldg(&rec.th.w0[hb + i]), a1 = __ldg(&rec.th.w1[hb + i]);
        for (int j = 0; j < nw; ++j) {
          const int wl = __ldg(&rec.tw.lo[wb + j]), wrr = __ldg(&rec.tw.hi[wb + j]);
          const float* t = base + ((size_t)hl * W + wl) * C;
          const float* u = base + ((size_t)hh * W + wl) * C;
          const size_t dxc = (size_t)(wrr - wl) * C;
          max_update(m0, m1, bilin2(a0, a1, __ldg(&rec.tw.w0[wb + j]), __ldg(&rec.tw.w1[wb + j]), ldg2(t), ldg2(u),
                                    ldg2(t + dxc), ldg2(u + dxc), a.negzero2));
        }
      }
    } else {
      // more samples than the tables hold (unreachable for finite rois on maps narrower than 2^17): the
      // reference's own loop, channel by channel
      const float sc_ = L.scale;
      const float* rp = a.rois + 4 * (size_t)n;
      const float rsw = __fmul_rn(__ldg(rp), sc_), rsh = __fmul_rn(__ldg(rp + 1), sc_);
      const float rew = __fmul_rn(__ldg(rp + 2), sc_), reh = __fmul_rn(__ldg(rp + 3), sc_);
      float bx, by;
      element_direct_strided(base, C, L.H, W, a.PH, a.PW, ph, pw, rsw, rsh, rew, reh, m0, bx, by);
      element_direct_strided(base + 1, C, L.H, W, a.PH, a.PW, ph, pw, rsw, rsh, rew, reh, m1, bx, by);
    }
    const uint32_t s = sa + (uint32_t)(g * kClGroup) * st4;
    sts_f32(s, m0);
    sts_f32(s + st4, m1);
  }
}

// One output row of the roi for kG channel groups; CY = class of the row axis.
template <int CY, int kC, int kG>
__device__ __forceinline__ void patch_row(const RoiAlignArgs& a, const PlanRecord& rec, const AxisRec* __restrict__ s_w,
                                          const AxisRec& hr, const float* __restrict__ r0, const int rstep, const int C,
                                          const int PW, uint32_t sa, const uint32_t st4, const int n, const int li,
                                          const int flags, const int c0, const int ph) {
  const uint64_t nz2 = a.negzero2;
  uint32_t sa1 = sa + st4, sb = sa + kClGroup * st4, sb1 = sb + st4;  // the lane's four channel rows (kG = 2)
  // the row pointer and the row step stay in registers: without the barrier the compiler re-derives them from
  // the kernel parameters (two constant loads and a 64-bit multiply-add chain) in every bin
  uint64_t r0v = reinterpret_cast<uint64_t>(r0);
  int rs = rstep;
  asm volatile("" : "+l"(r0v), "+r"(rs), "+r"(sa), "+r"(sa1), "+r"(sb), "+r"(sb1));
  for (int pw = 0; pw < PW; ++pw) {
    const AxisRec wr = s_w[pw];
    const int cx = wr.cls;
    if (cx >= 0) {
      float m[2 * kG];
      const float* q = reinterpret_cast<const float*>(r0v) + wr.lo0 * C;
      if (cx == 0) patch_bin<0, CY, kG>(q, rs, C, hr, wr, nz2, m);
      else if (cx == 1) patch_bin<1, CY, kG>(q, rs, C, hr, wr, nz2, m);
      else patch_bin<2, CY, kG>(q, rs, C, hr, wr, nz2, m);
      sts_f32(sa, m[0]);
      sts_f32(sa1, m[1]);
      if (kG > 1) {
        sts_f32(sb, m[2 % (2 * kG)]);
        sts_f32(sb1, m[3 % (2 * kG)]);
      }
#pragma unroll
      for (int g = 2; g < kG; ++g) {
        sts_f32(sa + (uint32_t)(g * kClGroup) * st4, m[2 * g]);
        sts_f32(sa1 + (uint32_t)(g * kClGroup) * st4, m[2 * g + 1]);
      }
    } else {
      slow_bin<kG>(a, rec, n, li, flags, c0, C, ph, pw, hr.cnt, wr.cnt, sa, st4);
    }
    sa += 4; sa1 += 4; sb += 4; sb1 += 4;
  }
}

// CTA = (roi, chunk of `rows` output rows, set of kG 64-channel groups), flattened in blockIdx.x; a warp takes whole
// output rows.  Results are collected channel-major in shared memory - the order of the NCHW-shaped output - and
// written out by the whole CTA as contiguous runs.
template <int kC, int kG>  // channel count as a compile-time constant (0 = run-time): pixel steps become immediates
__global__ void __launch_bounds__(kClThreads, SDET_CL_MINB)
roi_align_cl_kernel(const __grid_constant__ RoiAlignArgs a, const ClAux aux, const int rows, const int stride,
                    const int nsets, const int chunks) {
  __shared__ AxisRec s_h[16], s_w[16];
  __shared__ int s_scal[8];
  extern __shared__ __align__(16) float s_tile[];  // [kG * 64 channels][stride >= rows * PW]
  constexpr int kSet = kG * kClGroup;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  // launch order: rois fastest (largest window first), then row chunks, then channel sets - measured better than
  // keeping the CTAs of one roi adjacent (profiles/r02_roialign_cl.md)
  const int nroi = a.B * a.N;
  const int roi_i = blockIdx.x % nroi, roi_t = blockIdx.x / nroi, chunk_i = roi_t % chunks, set_i = roi_t / chunks;
  const int n = aux.order ? __ldg(aux.order + roi_i) : roi_i;
  const int C = kC ? kC : a.C, PH = a.PH, PW = a.PW, PP = PH * PW;
  const PlanRecord& rec = aux.plans[n];
  if (tid < 32) {  // the first warp (a CTA may have no other)
    const int p = tid & 15;
    if (p < (tid < 16 ? PH : PW)) (tid < 16 ? s_h : s_w)[p] = make_axis_rec(tid < 16 ? rec.th : rec.tw, p);
    if (tid < 8) s_scal[tid] = __ldg(&rec.scal[tid]);
  }
  __syncthreads();
  const int li = s_scal[0], flags = s_scal[1];
  const int c0 = set_i * kSet;                      // first channel of the set
  const int ncs = min(kSet, C - c0);                // channels in it
  const int ph0 = chunk_i * rows, nrow = min(rows, PH - ph0), run = nrow * PW;
  float* out_c = a.out + ((size_t)n * C + c0) * PP + (size_t)ph0 * PW;  // + c * PP + e
  const bool any = li >= 0 && s_scal[3] >= 0 && s_scal[5] >= 0;
  if (any) {
    const Level& L = a.lvl[li];
    const int W = L.W;
    const float* img = L.data + (size_t)(n / a.N) * L.H * W * C;  // NHWC: ((b*H + y)*W + x)*C + c
    const int rstep = W * C;  // elements between rows (fits 32 bits: a level has < 2^31 elements per image)
#ifdef SDET_CL_L2PF  // measured: no gain (profiles/r02_roialign_cl.md) - off
    // Ask L2 for the rows of the window this CTA is about to walk, every channel of them (the CTAs of the other
    // channel sets want the rest): one bulk prefetch per row, asynchronous.  The bins then find their pixels in L2
    // instead of paying an HBM round trip each - the loop is latency-bound otherwise (profiles/r02_roialign_cl.md).
    if ((C & 3) == 0) {
      const AxisRec& ha = s_h[ph0];
      const AxisRec& hb = s_h[ph0 + nrow - 1];
      const int y0 = ha.cls >= 0 ? ha.lo0 : s_scal[2];
      const int y1 = hb.cls >= 0 ? min(s_scal[3], hb.lo0 + (hb.pack >> 20)) : s_scal[3];
      const int x0 = s_scal[4];
      const uint32_t bytes = (uint32_t)(s_scal[5] - x0 + 1) * (uint32_t)C * 4u;
      for (int y = y0 + tid; y <= y1; y += blockDim.x) {
        const float* p = img + ((size_t)y * W + x0) * C;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
      }
    }
#endif
    const bool fast_ok = c0 + kSet <= C && (flags & kFlagOverflow) == 0;  // all lanes carry real channels
    const uint32_t st4 = (uint32_t)stride * 4u;
    const uint32_t tile_u32 = (uint32_t)__cvta_generic_to_shared(s_tile);
    for (int pr = warp; pr < nrow; pr += nwarps) {
      const int ph = ph0 + pr;
      const AxisRec hr = s_h[ph];
      const int cy = fast_ok ? hr.cls : -1;
      const float* r0 = img + (size_t)hr.lo0 * rstep + c0 + 2 * lane;  // one 64-bit multiply per output row
      const uint32_t sa = tile_u32 + (uint32_t)(2 * lane) * st4 + (uint32_t)(pr * PW) * 4u;
      if (cy == 0) patch_row<0, kC, kG>(a, rec, s_w, hr, r0, rstep, C, PW, sa, st4, n, li, flags, c0, ph);
      else if (cy == 1) patch_row<1, kC, kG>(a, rec, s_w, hr, r0, rstep, C, PW, sa, st4, n, li, flags, c0, ph);
      else if (cy == 2) patch_row<2, kC, kG>(a, rec, s_w, hr, r0, rstep, C, PW, sa, st4, n, li, flags, c0, ph);
      else
        for (int pw = 0; pw < PW; ++pw)
          slow_bin<kG>(a, rec, n, li, flags, c0, C, ph, pw, hr.cnt, s_w[pw].cnt, sa + 4u * pw, st4);
    }
    __syncthreads();
  }
  // write-out: channel c of the set owns `run` consecutive floats of the output (zeros when the roi has no level
  // or lies outside the map: the reference pools those to zero)
  if (any && stride == run && run == PP && (((size_t)ncs * PP) & 3) == 0 && ((((size_t)n * C + c0) * PP) & 3) == 0) {
    // whole rows of every channel and an unpadded tile: the set is ONE contiguous, 16-byte aligned block
    const float4* src = reinterpret_cast<const float4*>(s_tile);
    float4* dst = reinterpret_cast<float4*>(out_c);
    for (int i = tid; i < ncs * PP / 4; i += blockDim.x) __stcs(dst + i, src[i]);
  } else if (any) {
    // one channel per warp and step; run <= 128 (the host sizes the tile so): lanes take e = lane + 32 k
    const int k1 = lane + 32 < run, k2 = lane + 64 < run, k3 = lane + 96 < run;
    uint32_t ta = (uint32_t)__cvta_generic_to_shared(s_tile) + (uint32_t)(warp * stride + lane) * 4u;
    const uint32_t tstep = (uint32_t)(nwarps * stride) * 4u;
    float* op = out_c + (size_t)warp * PP + lane;
    const size_t ostep = (size_t)nwarps * PP;
    if (lane < run) {
      for (int c = warp; c < ncs; c += nwarps, ta += tstep, op += ostep) {
        float v0, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(ta));
        if (k1) asm volatile("ld.shared.f32 %0, [%1+128];" : "=f"(v1) : "r"(ta));
        if (k2) asm volatile("ld.shared.f32 %0, [%1+256];" : "=f"(v2) : "r"(ta));
        if (k3) asm volatile("ld.shared.f32 %0, [%1+384];" : "=f"(v3) : "r"(ta));
        __stcs(op, v0);
        if (k1) __stcs(op + 32, v1);
        if (k2) __stcs(op + 64, v2);
        if (k3) __stcs(op + 96, v3);
      }
    }
  } else {
    for (int c = warp; c < ncs; c += nwarps)
      for (int e = lane; e < run; e += 32) __stcs(out_c + (size_t)c * PP + e, 0.f);
  }
}

template <int kC, int kG>
int cl_launch_t(const RoiAlignArgs& a, const ClAux& aux, cudaStream_t st) {
  constexpr int kSet = kG * kClGroup;
  const int nsets = (a.C + kSet - 1) / kSet;
  // rows per CTA: as many as keep the tile within ~50 KB and a channel's run within 128 floats; one warp per row
  int rows = std::max(1, std::min(a.PH, std::min(128, 50 * 1024 / 4 / kSet) / a.PW));
  const int chunks = (a.PH + rows - 1) / rows;
  rows = (a.PH + chunks - 1) / chunks;
  const int warps = std::min(rows, kClWarps);
  const int stride = (rows * a.PW) | 1;  // odd: the lanes' channel rows fall into distinct banks (2-way at worst)
  const size_t smem = (size_t)kSet * stride * sizeof(float);
  SDET_CUDA(cudaFuncSetAttribute(roi_align_cl_kernel<kC, kG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)(a.B * a.N) * (unsigned)(nsets * chunks);
  roi_align_cl_kernel<kC, kG><<<grid, warps * 32, smem, st>>>(a, aux, rows, stride, nsets, chunks);
  return SDET_OK;
}

}  // namespace

size_t cl_scratch_bytes(int B, int C, const int* H, const int* W, int num_levels) {
  size_t total = 0;
  for (int l = 0; l < num_levels; ++l) total += (((size_t)B * C * H[l] * W[l] * sizeof(float)) + 255) & ~(size_t)255;
  return total;
}

// Re-lay `a.lvl[*].data` (NCHW) into `scratch` (NHWC) and point the levels at it.
int cl_transpose(RoiAlignArgs& a, void* scratch, cudaStream_t st) {
  TrArgs t{};
  char* w = static_cast<char*>(scratch);
  int tiles = 0;
  for (int l = 0; l < a.num_levels; ++l) {
    t.src[l] = a.lvl[l].data;
    t.dst[l] = reinterpret_cast<float*>(w);
    t.hw[l] = a.lvl[l].H * a.lvl[l].W;
    t.tile0[l] = tiles;
    tiles += (t.hw[l] + 31) / 32;
    w += (((size_t)a.B * a.C * t.hw[l] * sizeof(float)) + 255) & ~(size_t)255;
    a.lvl[l].data = t.dst[l];
  }
  t.tile0[a.num_levels] = tiles;
  t.num_levels = a.num_levels;
  t.C = a.C;
  t.B = a.B;
  dim3 grid((unsigned)tiles, (unsigned)((a.C + 31) / 32), (unsigned)a.B);
  nchw_to_nhwc_kernel<<<grid, 256, 0, st>>>(t);
  SDET_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return SDET_OK;
}

// `a.lvl[*].data` must be NHWC; plans / order as written by roi_align_plan_kernel + roi_align_order_kernel.
int cl_launch(const RoiAlignArgs& a, const PlanRecord* plans, const int* order, cudaStream_t st) {
  if (a.argx != nullptr) return sdet::fail(SDET_ERR_UNSUPPORTED, "channels-last path: argmax planes are not built");
  if (a.PH > 16 || a.PW > 16 || (a.C & 1))
    return sdet::fail(SDET_ERR_UNSUPPORTED, "channels-last path needs pooled_size <= 16 and an even channel count");
  ClAux aux{plans, order};
  int rc;
#ifndef SDET_CL_G
#define SDET_CL_G 2
#endif
  if (a.C == 256) rc = cl_launch_t<256, SDET_CL_G>(a, aux, st);
  else rc = cl_launch_t<0, SDET_CL_G>(a, aux, st);
  if (rc != SDET_OK) return rc;
  SDET_LAUNCH_CHECK("roi_align_cl_kernel");
  return SDET_OK;
}

}  // namespace sdet_ra
