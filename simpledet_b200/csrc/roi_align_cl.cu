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
#include "roi_align_common.cuh"

using namespace sdet_ra;

namespace sdet_ra {

namespace {

constexpr int kClWarps = 8;
constexpr int kClThreads = kClWarps * 32;
constexpr int kClGroup = 64;       // channels per warp pass (2 per lane)
constexpr int kClRowStride = 66;   // floats between pw slots of the output tile (even: 8-byte stores; 2*pw + c banks)

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

__global__ void __launch_bounds__(kClThreads, 4)
roi_align_cl_kernel(const __grid_constant__ RoiAlignArgs a, const ClAux aux) {
  __shared__ __align__(16) PlanRecord s_rec;
  extern __shared__ __align__(16) float s_tile[];  // [warp][PW][kClRowStride]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = aux.order ? __ldg(aux.order + blockIdx.x) : (int)blockIdx.x;
  const int C = a.C, PH = a.PH, PW = a.PW, PP = PH * PW;
  {
    const int4* src = reinterpret_cast<const int4*>(aux.plans + n);
    int4* dst = reinterpret_cast<int4*>(&s_rec);
    for (int i = tid; i < (int)(sizeof(PlanRecord) / 16); i += kClThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int li = s_rec.scal[0], flags = s_rec.scal[1];
  const int cg0 = blockIdx.y * kClGroup;               // first channel of this CTA's group
  const int ncg = min(kClGroup, C - cg0);              // channels in the group
  float* out_n = a.out + ((size_t)n * C + cg0) * PP;
  const bool any = li >= 0 && s_rec.scal[3] >= 0 && s_rec.scal[5] >= 0;
  if (!any) {  // no level / nothing inside the map: the reference pools to zeros
    for (int e = tid; e < ncg * PP; e += kClThreads) __stcs(out_n + e, 0.f);
    return;
  }
  const Level& L = a.lvl[li];
  const int W = L.W;
  const int b = n / a.N;
  const int c = cg0 + 2 * lane;
  const bool lane_on = c < C;
  const float* base = L.data + (size_t)b * L.H * W * C + (lane_on ? c : cg0);  // NHWC: ((b*H + y)*W + x)*C + c
  const uint64_t nz2 = a.negzero2;
  float* tile = s_tile + (size_t)warp * PW * kClRowStride;
  const bool fast_roi = (flags & (kFlagNot2 | kFlagOverflow)) == 0;

  for (int ph = warp; ph < PH; ph += kClWarps) {
    const int nh = s_rec.th.cnt[ph];
    // the two h-samples of this bin row (valid when nh == 2)
    const int hb = ph * kMaxS;
    const int lo0 = s_rec.th.lo[hb], hi0 = s_rec.th.hi[hb], lo1 = s_rec.th.lo[hb + 1], hi1 = s_rec.th.hi[hb + 1];
    const float a00 = s_rec.th.w0[hb], a01 = s_rec.th.w1[hb], a10 = s_rec.th.w0[hb + 1], a11 = s_rec.th.w1[hb + 1];
    // canonical rows: both samples interpolate between adjacent rows, the second at most one row below the first
    const int dy = lo1 - lo0;
    const bool rows_ok = nh == 2 && hi0 == lo0 + 1 && hi1 == lo1 + 1 && (dy == 0 || dy == 1);
    const float* r0 = base + (size_t)lo0 * W * C;
    const size_t rstep = (size_t)W * C;
    for (int pw = 0; pw < PW; ++pw) {
      const int nw = s_rec.tw.cnt[pw];
      float m0, m1;
      if (nh < 0 || nw < 0) {  // empty along an axis: 0 (roi_align_v2-inl.h:111-117)
        m0 = m1 = 0.f;
      } else {
        m0 = m1 = -FLT_MAX;
        const int wb = pw * kMaxS;
        const int xl0 = s_rec.tw.lo[wb], xr0 = s_rec.tw.hi[wb], xl1 = s_rec.tw.lo[wb + 1], xr1 = s_rec.tw.hi[wb + 1];
        const int dx = xl1 - xl0;
        if (fast_roi && rows_ok && nw == 2 && xr0 == xl0 + 1 && xr1 == xl1 + 1 && (dx == 0 || dx == 1)) {
          const float b00 = s_rec.tw.w0[wb], b01 = s_rec.tw.w1[wb], b10 = s_rec.tw.w0[wb + 1], b11 = s_rec.tw.w1[wb + 1];
          // patch of (2 + dy) x (2 + dx) pixels, every load one coalesced 256-byte request of the warp
          const float* q = r0 + (size_t)xl0 * C;
          const uint64_t p00 = ldg2(q), p01 = ldg2(q + C), p10 = ldg2(q + rstep), p11 = ldg2(q + rstep + C);
          uint64_t p02 = 0, p12 = 0, p20 = 0, p21 = 0, p22 = 0;
          if (dx) {
            p02 = ldg2(q + 2 * C);
            p12 = ldg2(q + rstep + 2 * C);
          }
          if (dy) {
            p20 = ldg2(q + 2 * rstep);
            p21 = ldg2(q + 2 * rstep + C);
            if (dx) p22 = ldg2(q + 2 * rstep + 2 * C);
          }
          // reference order (h0,w0), (h0,w1), (h1,w0), (h1,w1); the max itself is order-free
          max_update(m0, m1, bilin2(a00, a01, b00, b01, p00, p10, p01, p11, nz2));
          if (dx) max_update(m0, m1, bilin2(a00, a01, b10, b11, p01, p11, p02, p12, nz2));
          else max_update(m0, m1, bilin2(a00, a01, b10, b11, p00, p10, p01, p11, nz2));
          if (dy) {
            max_update(m0, m1, bilin2(a10, a11, b00, b01, p10, p20, p11, p21, nz2));
            if (dx) max_update(m0, m1, bilin2(a10, a11, b10, b11, p11, p21, p12, p22, nz2));
            else max_update(m0, m1, bilin2(a10, a11, b10, b11, p10, p20, p11, p21, nz2));
          } else {
            max_update(m0, m1, bilin2(a10, a11, b00, b01, p00, p10, p01, p11, nz2));
            if (dx) max_update(m0, m1, bilin2(a10, a11, b10, b11, p01, p11, p02, p12, nz2));
            else max_update(m0, m1, bilin2(a10, a11, b10, b11, p00, p10, p01, p11, nz2));
          }
        } else if ((flags & kFlagOverflow) == 0) {
          // general table walk: any sample count up to kMaxS, clamped / coincident corners
          for (int i = 0; i < nh; ++i) {
            const int hl = s_rec.th.lo[hb + i], hh = s_rec.th.hi[hb + i];
            const float a0 = s_rec.th.w0[hb + i], a1 = s_rec.th.w1[hb + i];
            for (int j = 0; j < nw; ++j) {
              const int wl = s_rec.tw.lo[wb + j], wr = s_rec.tw.hi[wb + j];
              const float* t = base + ((size_t)hl * W + wl) * C;
              const float* u = base + ((size_t)hh * W + wl) * C;
              const size_t dxc = (size_t)(wr - wl) * C;
              max_update(m0, m1, bilin2(a0, a1, s_rec.tw.w0[wb + j], s_rec.tw.w1[wb + j], ldg2(t), ldg2(u), ldg2(t + dxc),
                                        ldg2(u + dxc), nz2));
            }
          }
        } else {
          // more samples than the tables hold (unreachable for finite rois on maps narrower than 2^17): the
          // reference's own loop, channel by channel
          const float sc_ = L.scale;
          const float rsw = __fmul_rn(__ldg(a.rois + 4 * (size_t)n), sc_), rsh = __fmul_rn(__ldg(a.rois + 4 * (size_t)n + 1), sc_);
          const float rew = __fmul_rn(__ldg(a.rois + 4 * (size_t)n + 2), sc_), reh = __fmul_rn(__ldg(a.rois + 4 * (size_t)n + 3), sc_);
          float bx, by;
          element_direct_strided(base, C, L.H, W, PH, PW, ph, pw, rsw, rsh, rew, reh, m0, bx, by);
          element_direct_strided(base + 1, C, L.H, W, PH, PW, ph, pw, rsw, rsh, rew, reh, m1, bx, by);
        }
      }
      if (lane_on) *reinterpret_cast<float2*>(tile + pw * kClRowStride + 2 * lane) = make_float2(m0, m1);
    }
    __syncwarp();
    // write the (ph) row of the group's channels: PW contiguous floats per channel in the NCHW-ordered output
    float* orow = out_n + (size_t)ph * PW;
    for (int e = lane; e < ncg * PW; e += 32) {
      const int cl = e / PW, pw = e - cl * PW;
      __stcs(orow + (size_t)cl * PP + pw, tile[pw * kClRowStride + cl]);
    }
    __syncwarp();
  }
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
  const size_t smem = (size_t)kClWarps * a.PW * kClRowStride * sizeof(float);
  SDET_CUDA(cudaFuncSetAttribute(roi_align_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(a.B * a.N), (unsigned)((a.C + kClGroup - 1) / kClGroup));
  roi_align_cl_kernel<<<grid, kClThreads, smem, st>>>(a, aux);
  SDET_LAUNCH_CHECK("roi_align_cl_kernel");
  return SDET_OK;
}

}  // namespace sdet_ra
