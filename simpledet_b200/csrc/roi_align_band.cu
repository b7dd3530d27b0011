// Band-stationary _contrib_ROIAlign_v2 forward for sm_100a (inference: no argmax planes).
//
// Reference semantics: operator_cxx/contrib/roi_align_v2-inl.h:61-153 (forward functor),
// models/FPN/assign_layer_fpn.py:17-40 (level assignment).  The per-roi kernel (roi_align.cu) stages
// every roi's own window, so a feature byte travels L2 -> SM once per roi that covers it (4-7x on the
// 800x1333 pyramid) in 56-170 byte rows that cannot be described to the TMA.  Here the *feature map*
// is stationary instead:
//
//   plan    one small CTA per roi restates the reference's sample loop into compact tables
//           (RoiTab) and cuts the roi's output rows into items = (roi, ph0, nph): the bins whose first
//           sample row lies in the same 8-row band of the level;
//   layout  prefix sum over the bands, items scattered into per-band lists, units = (band,
//           <= cap items, channel chunk) written most-expensive-first;
//   main    persistent CTAs (one per SM) pull units from a queue.  Warp 0 is the producer: for each
//           stage it issues ONE 1-D bulk TMA copy per channel — 12 full-width rows of a channel plane
//           are contiguous in NCHW, start on a 16-byte boundary (8*W*4 is a multiple of 16) and are
//           a multiple of 16 bytes long — into a 3-deep ring of 64 KB stages, completing on the
//           stage's `full` mbarrier together with the unit's item tables.  15 consumer warps take
//           (item, channel quad) jobs from a per-stage counter and run the bit-exact packed-fp32
//           bilinear/max arithmetic against the staged band; `empty` mbarriers hand the stage back.
//
// Every feature byte is staged 1.5x (halo) instead of once per covering roi; rows are read as whole
// 1.3 KB bursts.  Rois the band path cannot take (bins taller than the halo, sample counts != 2,
// unaligned or very wide levels) go to the per-roi kernel through `left_order`.
#include <type_traits>

#include "roi_align_common.cuh"

using namespace sdet_ra;

namespace sdet_ra {

namespace {

// A/B knobs (measured, profiles/r02_band_ab.txt): 15 consumer warps with 4-channel jobs is the best point -
// 8-channel jobs need 236 registers (spills at any useful warp count), 11 warps lose 16 % to latency.
#ifndef SDET_BAND_CONSUMERS
#define SDET_BAND_CONSUMERS 15
#endif
#ifndef SDET_BAND_CL8
#define SDET_BAND_CL8 0
#endif
constexpr int kBandConsumers = SDET_BAND_CONSUMERS;      // consumer warps; warp 0 is the producer
constexpr int kBandThreads = 32 * (kBandConsumers + 1);
constexpr int kStageBytes = kBandStageFloats * 4;

struct StageDesc {   // 64 bytes, written by the producer before it arms the stage's `full` barrier
  int stop;          // 1: no more work
  int nitems;        // items of the unit
  int tbl;           // which of the two table buffers holds the unit's items
  int c0;            // first channel of this stage
  int nch;           // channels in this stage (multiple of 4)
  int cs_log2;       // channel stride class
  int oddshift;      // byte shift of odd channels inside their slot
  int last;          // last stage of the unit
  int next;          // job counter (consumers atomicAdd)
  int pad[7];
};
static_assert(sizeof(StageDesc) == 64, "StageDesc layout");

__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk TMA copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint2 lds64(unsigned addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds32(unsigned addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(unsigned addr, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// R[k][t] = smem[(k odd ? base_odd_t : base_even_t) + k*kCS*4]: immediates fully unrolled
template <int CL, int kCS, int K = 0>
struct TapLoaderEO {
  static __device__ __forceinline__ void run(float (&R)[CL][2], unsigned ale, unsigned are, unsigned alo,
                                             unsigned aro) {
    R[K][0] = lds_f32_imm<K * kCS * 4>((K & 1) ? alo : ale);
    R[K][1] = lds_f32_imm<K * kCS * 4>((K & 1) ? aro : are);
    TapLoaderEO<CL, kCS, K + 1>::run(R, ale, are, alo, aro);
  }
};
template <int CL, int kCS>
struct TapLoaderEO<CL, kCS, CL> {
  static __device__ __forceinline__ void run(float (&)[CL][2], unsigned, unsigned, unsigned, unsigned) {}
};

__device__ __forceinline__ int band_level(const BandArgs& ba, int num_levels, int rem) {
  int li = 0;
  for (int l = 0; l < num_levels; ++l)
    if (ba.geom[l].ok && rem >= ba.geom[l].base && rem < ba.geom[l].base + ba.geom[l].nb) li = l;
  return li;
}

// ---------------------------------------------------------------------------------------------
// plan: one CTA of 64 threads per roi
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
roi_align_band_plan_kernel(const __grid_constant__ RoiAlignArgs a, const __grid_constant__ BandArgs ba,
                           PlanRecord* __restrict__ plans) {
  __shared__ __align__(16) PlanRecord s_rec;
  __shared__ int s_bd[16];
  __shared__ int s_code[32];
  __shared__ int s_ok;
  const int n = blockIdx.x, tid = threadIdx.x;
  const int PH = a.PH, PW = a.PW;
  roi_preamble<16>(a, n, PH, PW, s_rec.th, s_rec.tw, s_rec.scal);
  const int li = s_rec.scal[0], flags = s_rec.scal[1];
  if (a.levels_out != nullptr && tid == 0) a.levels_out[n] = li;
  if (tid == 0) {
    bool ok = li >= 0 && ba.geom[li >= 0 ? li : 0].ok && (flags & (kFlagNot2 | kFlagOverflow)) == 0 &&
              s_rec.scal[3] >= 0 && s_rec.scal[5] >= 0;
    if (ok) {
      int first = -1;
      for (int ph = 0; ph < PH; ++ph) {
        int bd = -1;
        if (s_rec.th.cnt[ph] == 2) {
          const int lo0 = s_rec.th.lo[ph * kMaxS], hi1 = s_rec.th.hi[ph * kMaxS + 1];
          bd = lo0 / kBandR;
          if (hi1 - bd * kBandR >= kBandRS) ok = false;  // the bin is taller than the halo
          if (first < 0) first = bd;
        }
        s_bd[ph] = bd;
      }
      if (first < 0) ok = false;
      int cur = first;  // bins that are empty along h ride with the previous bin's band
      for (int ph = 0; ph < PH; ++ph) {
        if (s_bd[ph] < 0) s_bd[ph] = cur;
        else cur = s_bd[ph];
      }
      // Row-cache codes: simulate band_compute's two register sets (tags = absolute rows) over each item.
      // bits [1:0]: 0 low row is in set A, 1 in set B, 2 load it into A; bit 16: load the high row into the other set.
      int tagA = -1, tagB = -1, prev_bd = -1;
      for (int ph = 0; ph < PH; ++ph) {
        if (s_bd[ph] != prev_bd) {  // a new item starts with an empty cache
          tagA = tagB = -1;
          prev_bd = s_bd[ph];
        }
        for (int s = 0; s < 2; ++s) {
          int code = 0;
          if (s_rec.th.cnt[ph] == 2) {
            const int lo = s_rec.th.lo[ph * kMaxS + s], hi = s_rec.th.hi[ph * kMaxS + s];
            if (lo == tagA) {
              if (hi != tagB) { code |= 0x10000; tagB = hi; }
            } else if (lo == tagB) {
              code |= 1;
              if (hi != tagA) { code |= 0x10000; tagA = hi; }
            } else {
              code |= 2;
              tagA = lo;
              if (hi != tagB) { code |= 0x10000; tagB = hi; }
            }
          }
          s_code[ph * 2 + s] = code;
        }
      }
    }
    s_ok = ok ? 1 : 0;
  }
  __syncthreads();
  if (!s_ok) {  // the per-roi kernel takes it: it needs the full record
    const int4* src = reinterpret_cast<const int4*>(&s_rec);
    int4* dst = reinterpret_cast<int4*>(plans + n);
    for (int i = tid; i < (int)(sizeof(PlanRecord) / 16); i += blockDim.x) dst[i] = src[i];
    if (tid == 0) {
      ba.w.left_order[atomicAdd(&ba.w.ctr[2], 1)] = n;
      ba.w.ritems[n].n = -1;
    }
    return;
  }
  const Level& L = a.lvl[li];
  const int W4 = L.W * 4;
  RoiTab& T = ba.w.tabs[n];
  if (tid < 32) {
    const int pw = tid >> 1, s = tid & 1;
    uint2 e = make_uint2(0xFFFFFFFFu, 0u);
    float wc = -1.f;
    if (pw < PW && s_rec.tw.cnt[pw] == 2) {
      const int k = pw * kMaxS + s;
      e.x = (unsigned)(s_rec.tw.lo[k] * 4) | ((unsigned)(s_rec.tw.hi[k] * 4) << 16);
      e.y = __float_as_uint(s_rec.tw.w1[k]);
      wc = s_rec.tw.coord[k];
    }
    T.lane[tid] = e;
    T.wcoord[tid] = wc;
  } else {
    const int idx = tid - 32, ph = idx >> 1, s = idx & 1;
    uint2 e = make_uint2(0xFFFFFFFFu, 0u);
    float hc = -1.f;
    if (ph < PH && s_rec.th.cnt[ph] == 2) {
      const int k = ph * kMaxS + s;
      const int r0 = s_bd[ph] * kBandR;
      e.x = (unsigned)((s_rec.th.lo[k] - r0) * W4) | ((unsigned)((s_rec.th.hi[k] - r0) * W4) << 16) |
            (unsigned)s_code[ph * 2 + s];
      e.y = __float_as_uint(s_rec.th.w1[k]);
      hc = s_rec.th.coord[k];
    }
    T.row[ph][s] = e;
    T.hcoord[ph][s] = hc;
  }
  if (tid == 0) {
    const int b = n / a.N;
    RoiItems& R = ba.w.ritems[n];
    int k = 0, ph = 0;
    while (ph < PH) {
      const int bd = s_bd[ph], ph0 = ph;
      while (ph < PH && s_bd[ph] == bd) ++ph;
      const int g = b * ba.bands_per_image + ba.geom[li].base + bd;
      const int rank = atomicAdd(&ba.w.band_cnt[g], 1);
      atomicAdd(&ba.w.band_rows[g], ph - ph0);
      R.it[k++] = make_int4(g, rank, ph0, ph - ph0);
    }
    R.n = k;
    R.flags = flags;
  }
}

// ---------------------------------------------------------------------------------------------
// layout: band offsets (every CTA, redundantly), item scatter (one thread per roi), unit list (CTA 0)
// ---------------------------------------------------------------------------------------------
constexpr int kUnitBuckets = 64;

__global__ void __launch_bounds__(256)
roi_align_band_layout_kernel(const __grid_constant__ BandArgs ba, const int num_levels, const int total_rois,
                             const int C) {
  __shared__ int s_off[kBandMaxBands + 1];
  __shared__ int s_warp[8];
  __shared__ int s_hist[kUnitBuckets], s_base[kUnitBuckets];
  const int tid = threadIdx.x, NB = ba.num_bands;
  const int per = (NB + 255) / 256;
  int local = 0;
  for (int i = 0; i < per; ++i) {
    const int g = tid * per + i;
    if (g < NB) local += ba.w.band_cnt[g];
  }
  int incl = local;
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= d) incl += v;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  if (tid < kUnitBuckets) s_hist[tid] = 0;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < (tid >> 5); ++w) wbase += s_warp[w];
  int run = wbase + incl - local;
  for (int i = 0; i < per; ++i) {
    const int g = tid * per + i;
    if (g < NB) {
      s_off[g] = run;
      run += ba.w.band_cnt[g];
    }
  }
  if (tid == 255) s_off[NB] = run;
  __syncthreads();

  const int n = blockIdx.x * blockDim.x + tid;
  if (n < total_rois) {
    const RoiItems& R = ba.w.ritems[n];
    const int k = R.n, fl = R.flags;
    for (int i = 0; i < k; ++i) {
      const int4 it = R.it[i];
      ba.w.band_items[s_off[it.x] + it.y] = make_int2(n, it.z | (it.w << 8) | (fl << 16));
    }
  }
  if (blockIdx.x != 0) return;

  for (int g = tid; g <= NB; g += blockDim.x) ba.w.band_off[g] = s_off[g];
  // units, most expensive first: pass 1 histogram of cost classes, pass 2 scatter
  auto visit = [&](bool emit) {
    for (int g = tid; g < NB; g += blockDim.x) {
      const int cnt = s_off[g + 1] - s_off[g];
      if (cnt == 0) continue;
      const int rem = g % ba.bands_per_image;
      const int li = band_level(ba, num_levels, rem);
      const int chunk = ba.geom[li].chunk;
      const int nchunks = (C + chunk - 1) / chunk;
      const int rows = ba.w.band_rows[g];
      const int nst = max(1, chunk >> (14 - ba.geom[li].cs_log2));  // 64 KB stages per unit
      for (int i0 = 0; i0 < cnt; i0 += ba.cap) {
        const int nit = min(ba.cap, cnt - i0);
        const int cost = max((int)((long long)rows * nit / cnt) * chunk, 400 * nst);  // arithmetic vs staging
        const int bucket = min(kUnitBuckets - 1, cost >> 9);
        if (!emit) {
          atomicAdd(&s_hist[bucket], nchunks);
        } else {
          const int pos = s_base[bucket] + atomicAdd(&s_hist[bucket], nchunks);
          for (int c = 0; c < nchunks; ++c) {
            const int c0 = c * chunk, nch = min(chunk, C - c0);
            if (pos + c < ba.w.max_units) ba.w.units[pos + c] = make_int4(g, s_off[g] + i0, nit, c0 | (nch << 16));
          }
        }
      }
    }
  };
  visit(false);
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int b = kUnitBuckets - 1; b >= 0; --b) {
      s_base[b] = acc;
      acc += s_hist[b];
      s_hist[b] = 0;
    }
    ba.w.ctr[1] = min(acc, ba.w.max_units);
    ba.w.ctr[3] = s_off[NB];
  }
  __syncthreads();
  visit(true);
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
// One job = one item x CL channels.  The row cache (two register sets RA / RB, each one staged row of the
// lane's two taps for CL channels) is driven by codes the plan kernel precomputed by simulating exactly this
// replacement policy from the item's first bin: bits [1:0] of an entry's low half say where the sample's low
// row lives (0: in RA, 1: in RB, 2: nowhere -> load it into RA), bit 16 says the high row must be loaded into
// the other set.  No run-time tag compares, no votes.
template <int kCS, int CL>
__device__ __forceinline__ void band_compute(const RoiAlignArgs& a, const unsigned chbase, const unsigned slot,
                                             const int2 it, const int c, const unsigned oddshift, const int lane,
                                             const uint64_t nz2) {
  const int n = it.x, ph0 = it.y & 0xFF, nph = (it.y >> 8) & 0xFF;
  const bool has_empty = ((it.y >> 16) & kFlagEmpty) != 0;
  const int PH = a.PH, PW = a.PW, PP = PH * PW;
  const int pw = lane >> 1, sx = lane & 1;
  const bool lane_on = pw < PW;
  const uint2 le = lds64(slot + (unsigned)lane * 8u);
  const bool wvalid = le.x != 0xFFFFFFFFu;
  const unsigned xl = wvalid ? (le.x & 0xFFFFu) : 0u, xr = wvalid ? (le.x >> 16) : 0u;
  const float b1 = __uint_as_float(le.y), b0 = __fsub_rn(1.f, b1);
  const unsigned ale = chbase + xl, are = chbase + xr, alo = ale + oddshift, aro = are + oddshift;
  float RA[CL][2], RB[CL][2];
  // after the pair exchange both lanes of a pw hold all CL maxima; lane s stores channels
  // [s*CL/2, (s+1)*CL/2) so every store instruction has 2*PW active lanes
  float* outh = a.out + ((size_t)n * a.C + c + sx * (CL / 2)) * PP + (size_t)ph0 * PW + pw;

  auto sample = [&](const unsigned offs, const float a1, float (&m)[CL], auto first_tag) {
    constexpr bool kFirst = decltype(first_tag)::value;
    const unsigned olo = offs & 0xFFFCu, ohi = (offs >> 16) & 0xFFFCu;
    const float a0 = __fsub_rn(1.f, a1);
    const float wtl = __fmul_rn(a0, b0), wbl = __fmul_rn(a1, b0);
    const float wtr = __fmul_rn(a0, b1), wbr = __fmul_rn(a1, b1);
    const uint64_t wtl2 = pack2(wtl, wtl), wbl2 = pack2(wbl, wbl);
    const uint64_t wtr2 = pack2(wtr, wtr), wbr2 = pack2(wbr, wbr);
    auto step = [&](const float (&Lo)[CL][2], const float (&Hi)[CL][2]) {
#pragma unroll
      for (int k = 0; k < CL; k += 2) {
        // roi_align_v2-inl.h:137-140 for channels k, k+1: ((tl + bl) + tr) + br, products rounded separately
        const uint64_t ptl = fma2(wtl2, pack2(Lo[k][0], Lo[k + 1][0]), nz2);
        const uint64_t pbl = fma2(wbl2, pack2(Hi[k][0], Hi[k + 1][0]), nz2);
        const uint64_t ptr = fma2(wtr2, pack2(Lo[k][1], Lo[k + 1][1]), nz2);
        const uint64_t pbr = fma2(wbr2, pack2(Hi[k][1], Hi[k + 1][1]), nz2);
        float va, vb;
        unpack2(add2(add2(add2(ptl, pbl), ptr), pbr), va, vb);
        // running maximum over the bin's samples; fmaxf drops a NaN operand like `v > m` does
        m[k] = kFirst ? va : fmaxf(m[k], va);
        m[k + 1] = kFirst ? vb : fmaxf(m[k + 1], vb);
      }
    };
    const unsigned path = offs & 3u;
    const bool ldhi = (offs & 0x10000u) != 0;
    if (path == 1u) {  // low row in RB
      if (ldhi) TapLoaderEO<CL, kCS>::run(RA, ale + ohi, are + ohi, alo + ohi, aro + ohi);
      step(RB, RA);
    } else {
      if (path == 2u) TapLoaderEO<CL, kCS>::run(RA, ale + olo, are + olo, alo + olo, aro + olo);
      if (ldhi) TapLoaderEO<CL, kCS>::run(RB, ale + ohi, are + ohi, alo + ohi, aro + ohi);
      step(RA, RB);
    }
  };

  const unsigned rowtab = slot + 256u + (unsigned)ph0 * 16u;
  for (int i = 0; i < nph; ++i) {
    const uint4 re = lds128(rowtab + (unsigned)i * 16u);  // {offs s0, alpha s0, offs s1, alpha s1}
    if (re.x == 0xFFFFFFFFu) {  // bin empty along h (warp-uniform): pools to 0 (roi_align_v2-inl.h:111-117)
      if (lane_on) {
#pragma unroll
        for (int k = 0; k < CL / 2; ++k) __stcs(outh + k * PP, 0.f);
      }
      outh += PW;
      continue;
    }
    float best[CL];
    sample(re.x, __uint_as_float(re.y), best, std::true_type{});
    sample(re.z, __uint_as_float(re.w), best, std::false_type{});
    const bool zero_out = has_empty && !wvalid;  // bin empty along w
#pragma unroll
    for (int k = 0; k < CL; ++k) best[k] = max3f(best[k], __shfl_xor_sync(0xffffffffu, best[k], 1), -FLT_MAX);
    if (lane_on) {
#pragma unroll
      for (int k = 0; k < CL / 2; ++k) {
        const float r = sx ? best[CL / 2 + k] : best[k];
        __stcs(outh + k * PP, zero_out ? 0.f : r);
      }
    }
    outh += PW;
  }
}

template <bool kArg>
__global__ void __launch_bounds__(kBandThreads, 1)
roi_align_band_kernel(const __grid_constant__ RoiAlignArgs a, const __grid_constant__ BandArgs ba) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kSlot = kArg ? 768 : 512;
  constexpr int kCap = kArg ? kBandCapTrain : kBandCapInfer;
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem_raw);
  const unsigned tbl0 = s0 + kBandStages * kStageBytes;
  const unsigned items0 = tbl0 + 2 * kCap * kSlot;
  const unsigned desc0 = items0 + 2 * 32 * 8;
  const unsigned bar0 = desc0 + kBandStages * 64;
  // full[i] = bar0 + 8i, empty[i] = bar0 + 24 + 8i, tblfree[j] = bar0 + 48 + 8j
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kBandStages; ++i) {
      mbar_init(bar0 + 8u * i, 1);
      mbar_init(bar0 + 24u + 8u * i, kBandConsumers);
    }
    mbar_init(bar0 + 48u, kBandConsumers);
    mbar_init(bar0 + 56u, kBandConsumers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // =============================== producer ===============================
    const int nunits = *reinterpret_cast<volatile const int*>(ba.w.ctr + 1);
    int seq = 0, ucount = 0;
    for (;;) {
      int u = 0;
      if (lane == 0) u = atomicAdd(&ba.w.ctr[0], 1);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= nunits) break;
      const int4 U = __ldg(ba.w.units + u);
      const int g = U.x, item0 = U.y, nit = U.z, c0 = U.w & 0xFFFF, nch = U.w >> 16;
      const int b = g / ba.bands_per_image, rem = g - b * ba.bands_per_image;
      const int li = band_level(ba, a.num_levels, rem);
      const BandGeom& G = ba.geom[li];
      const Level& L = a.lvl[li];
      const int r0 = (rem - G.base) * kBandR;
      const int rows = min(kBandRS, L.H - r0);
      const int CT = kBandStageFloats >> G.cs_log2;
      const unsigned chan_bytes = 4u << G.cs_log2;
      const int tbl = ucount & 1;
      if (ucount >= 2) mbar_wait(bar0 + 48u + 8u * tbl, (unsigned)(((ucount >> 1) - 1) & 1));
      int2 it = make_int2(0, 0);
      if (lane < nit) {
        it = __ldg(ba.w.band_items + item0 + lane);
        asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(items0 + (unsigned)(tbl * 32 + lane) * 8u), "r"(it.x),
                     "r"(it.y)
                     : "memory");
      }
      __syncwarp();
      const int nst = (nch + CT - 1) / CT;
      const size_t plane = (size_t)L.H * L.W;
      const float* gband = L.data + (size_t)b * a.C * plane + (size_t)r0 * L.W;
      const unsigned rowbytes = (unsigned)(rows * L.W * 4);
      for (int st = 0; st < nst; ++st, ++seq) {
        const int slot = seq % kBandStages;
        if (seq >= kBandStages) mbar_wait(bar0 + 24u + 8u * slot, (unsigned)(((seq / kBandStages) - 1) & 1));
        const int cbase = c0 + st * CT, nc = min(CT, nch - st * CT);
        unsigned mybytes = 0;
        const char* src = nullptr;
        if (lane < nc) {
          const int c = cbase + lane;
          const unsigned shift = (c & 1) ? (unsigned)G.oddshift : 0u;
          mybytes = (rowbytes + shift + 15u) & ~15u;
          src = reinterpret_cast<const char*>(gband + (size_t)c * plane) - shift;
        }
        unsigned tx = mybytes;
        for (int d = 16; d > 0; d >>= 1) tx += __shfl_xor_sync(0xffffffffu, tx, d);
        if (st == 0) tx += (unsigned)(nit * kSlot);
        const unsigned full = bar0 + 8u * slot;
        if (lane == 0) {
          const unsigned d = desc0 + 64u * slot;
          sts32(d + 0, 0);
          sts32(d + 4, nit);
          sts32(d + 8, tbl);
          sts32(d + 12, cbase);
          sts32(d + 16, nc);
          sts32(d + 20, G.cs_log2);
          sts32(d + 24, G.oddshift);
          sts32(d + 28, st == nst - 1);
          sts32(d + 32, 0);
          sts32(d + 36, (int)(0xFFFFFFFFu / (unsigned)nit + 1u));  // w / nit == umulhi(w, magic) for w < 2^16
          mbar_expect_tx(full, tx);
        }
        __syncwarp();
        if (lane < nc) bulk_g2s(s0 + (unsigned)slot * kStageBytes + (unsigned)lane * chan_bytes, src, mybytes, full);
        if (st == 0 && lane < nit)
          bulk_g2s(tbl0 + (unsigned)(tbl * kCap + lane) * kSlot, ba.w.tabs + it.x, kSlot, full);
      }
      ++ucount;
    }
    const int slot = seq % kBandStages;
    if (seq >= kBandStages) mbar_wait(bar0 + 24u + 8u * slot, (unsigned)(((seq / kBandStages) - 1) & 1));
    if (lane == 0) {
      sts32(desc0 + 64u * slot, 1);
      mbar_arrive(bar0 + 8u * slot);
    }
    return;
  }

  // =============================== consumers ===============================
  const uint64_t nz2 = a.negzero2;
  for (int seq = 0;; ++seq) {
    const int slot = seq % kBandStages;
    mbar_wait(bar0 + 8u * slot, (unsigned)((seq / kBandStages) & 1));
    const unsigned d = desc0 + 64u * slot;
    const uint4 d0 = lds128(d), d1 = lds128(d + 16);
    if (d0.x) break;
    const int nitems = (int)d0.y, tbl = (int)d0.z, cbase = (int)d0.w;
    const int nchs = (int)d1.x, cs_log2 = (int)d1.y;
    const unsigned oddshift = d1.z;
    const int last = (int)d1.w;
    const uint4 d2 = lds128(d + 32);  // {next, magic = ceil(2^32 / nitems), -, -}
    const unsigned magic = d2.y;
    const unsigned sbase = s0 + (unsigned)slot * kStageBytes;
    // classes with >= 8 channels per stage run 8 channels per lane (weights, table reads and row-cache control
    // are paid once per sample whatever the channel count); the widest class (4 channels per stage) runs 4
    const int cl_log2 = (!SDET_BAND_CL8 || cs_log2 == 12 || (nchs & 7)) ? 2 : 3;
    const int total = nitems * (nchs >> cl_log2);
    for (;;) {
      int w = 0;
      if (lane == 0) {
        asm volatile("atom.shared.add.s32 %0, [%1], 1;" : "=r"(w) : "r"(d + 32) : "memory");
      }
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w >= total) break;
      const int q = nitems == 1 ? w : (int)__umulhi((unsigned)w, magic);
      const int item = w - q * nitems;
      const uint2 itv = lds64(items0 + (unsigned)(tbl * 32 + item) * 8u);
      const int2 it = make_int2((int)itv.x, (int)itv.y);
      const unsigned slot_addr = tbl0 + (unsigned)(tbl * kCap + item) * kSlot;
      const unsigned chbase = sbase + ((unsigned)(q << cl_log2) << (cs_log2 + 2));
      const int c = cbase + (q << cl_log2);
      if (cl_log2 == 2) {
        switch (cs_log2) {
          case 9: band_compute<512, 4>(a, chbase, slot_addr, it, c, oddshift, lane, nz2); break;
          case 10: band_compute<1024, 4>(a, chbase, slot_addr, it, c, oddshift, lane, nz2); break;
          case 11: band_compute<2048, 4>(a, chbase, slot_addr, it, c, oddshift, lane, nz2); break;
          default: band_compute<4096, 4>(a, chbase, slot_addr, it, c, oddshift, lane, nz2); break;
        }
      } else {
        switch (cs_log2) {
          case 9: band_compute<512, 8>(a, chbase, slot_addr, it, c, oddshift, lane, nz2); break;
          case 10: band_compute<1024, 8>(a, chbase, slot_addr, it, c, oddshift, lane, nz2); break;
          default: band_compute<2048, 8>(a, chbase, slot_addr, it, c, oddshift, lane, nz2); break;
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(bar0 + 24u + 8u * slot);
      if (last) mbar_arrive(bar0 + 48u + 8u * tbl);
    }
  }
}

size_t al16(size_t v) { return (v + 15) & ~(size_t)15; }

}  // namespace

int band_max_units(size_t total_rois) {
  return (int)((total_rois * kBandItemsPerRoi / kBandCapTrain + kBandMaxBands) * kBandMaxChunks);
}

size_t band_workspace_bytes(size_t total) {
  return 256 + 3 * al16(4 * (kBandMaxBands + 1)) + sizeof(RoiTab) * total + sizeof(RoiItems) * total +
         al16(8 * total * kBandItemsPerRoi) + 16 * (size_t)band_max_units(total) + al16(4 * total);
}

// Fills `ba` (geometry + workspace carve-up).  Returns false when the band path cannot serve this call at all.
bool band_setup(const RoiAlignArgs& a, void* ws, BandArgs& ba) {
  if (a.argx != nullptr) return false;  // argmax planes: per-roi kernel
  if (a.PH > 16 || a.PW > 16 || (a.C & 3) != 0) return false;
  int base = 0;
  bool any = false;
  for (int l = 0; l < a.num_levels; ++l) {
    BandGeom& G = ba.geom[l];
    const Level& L = a.lvl[l];
    G = BandGeom{};
    const long long hw = (long long)L.H * L.W;
    if ((reinterpret_cast<uintptr_t>(L.data) & 15) != 0 || (hw & 1) != 0 || kBandRS * L.W + 4 > 4096) continue;
    G.ok = 1;
    int lg = 9;
    while ((1 << lg) < kBandRS * L.W + 4) ++lg;
    G.cs_log2 = lg;
    G.nb = (L.H + kBandR - 1) / kBandR;
    G.base = base;
    G.oddshift = (hw & 3) ? 8 : 0;
    const int CT = kBandStageFloats >> lg;
    int chunk = CT > 32 ? CT : 32;
    while ((a.C + chunk - 1) / chunk > kBandMaxChunks) chunk *= 2;
    G.chunk = chunk;
    base += G.nb;
    any = true;
  }
  if (!any) return false;
  ba.bands_per_image = base;
  const long long nb = (long long)base * a.B;
  if (nb > kBandMaxBands) return false;
  ba.num_bands = (int)nb;
  ba.cap = a.argx ? kBandCapTrain : kBandCapInfer;
  const size_t total = (size_t)a.B * a.N;
  char* w = static_cast<char*>(ws);
  ba.w.ctr = reinterpret_cast<int*>(w); w += 256;
  ba.w.band_cnt = reinterpret_cast<int*>(w); w += al16(4 * (kBandMaxBands + 1));
  ba.w.band_rows = reinterpret_cast<int*>(w); w += al16(4 * (kBandMaxBands + 1));
  ba.w.band_off = reinterpret_cast<int*>(w); w += al16(4 * (kBandMaxBands + 1));
  ba.w.tabs = reinterpret_cast<RoiTab*>(w); w += sizeof(RoiTab) * total;
  ba.w.ritems = reinterpret_cast<RoiItems*>(w); w += sizeof(RoiItems) * total;
  ba.w.band_items = reinterpret_cast<int2*>(w); w += al16(8 * total * kBandItemsPerRoi);
  ba.w.max_units = band_max_units(total);
  ba.w.units = reinterpret_cast<int4*>(w); w += 16 * (size_t)ba.w.max_units;
  ba.w.left_order = reinterpret_cast<int*>(w);
  return true;
}

// plan + layout + persistent main kernel.  The caller then runs the per-roi kernel over
// (ba.w.left_order, ba.w.ctr + 2) with `plans`.
int band_launch(const RoiAlignArgs& a, const BandArgs& ba, PlanRecord* plans, cudaStream_t st) {
  const int total = a.B * a.N;
  // ctr + band_cnt + band_rows are contiguous
  SDET_CUDA(cudaMemsetAsync(ba.w.ctr, 0, 256 + 2 * al16(4 * (kBandMaxBands + 1)), st));
  roi_align_band_plan_kernel<<<(unsigned)total, 64, 0, st>>>(a, ba, plans);
  SDET_LAUNCH_CHECK("roi_align_band_plan_kernel");
  roi_align_band_layout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ba, a.num_levels, total, a.C);
  SDET_LAUNCH_CHECK("roi_align_band_layout_kernel");
  int dev = 0, sms = 0;
  SDET_CUDA(cudaGetDevice(&dev));
  SDET_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  constexpr int smem = kBandStages * kStageBytes + 2 * kBandCapInfer * 512 + 2 * 32 * 8 + kBandStages * 64 + 64;
  auto k = roi_align_band_kernel<false>;
  SDET_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k<<<(unsigned)sms, kBandThreads, smem, st>>>(a, ba);
  SDET_LAUNCH_CHECK("roi_align_band_kernel");
  return SDET_OK;
}

}  // namespace sdet_ra
