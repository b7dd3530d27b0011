// RPN anchor targets on the GPU: AnchorTarget2D (core/detection_input.py:353-565) and
// PyramidAnchorTarget2D (models/FPN/input.py:55-148).  The reference runs this per image in loader
// threads (numpy + the Cython IoU); here a batch is four launches:
//   1. anchor_iou_kernel     per anchor: inside test, max / argmax IoU over the gts; per gt: max over
//                            the inside anchors (integer atomicMax on the IoU bits, IoU >= 0)
//   2. anchor_label_kernel   label rules (:455-475) — needs every gt's maximum, hence the second pass
//   3. anchor_quota_kernel   one CTA per image: the priority thresholds that keep `fg_quota` positives
//                            and `image_anchor - kept` negatives (radix select, _sample_anchor :477-494)
//   4. anchor_write_kernel   targets (float64 nonlinear_transform), scatter, (h,w,A) -> (A, h*w) layout
// Sub-sampling: np.random.choice has no portable restatement, so "which surplus anchors are
// disabled" is defined by 32-bit priorities (smallest priority first, ties: larger index first),
// injected by the caller or drawn from Philox4x32-10(seed, image*N + anchor).
#include <curand_kernel.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "topk.cuh"

namespace {

using sdet::kTopkThreads;
constexpr int kMaxA = 16;
constexpr int kMaxGt = 512;

struct ATLevel {
  int stride, dim_short, dim_long;
  int cell_off;    // cells before this level (sum of fh*fw)
};

struct ATParams {
  const float* im_info;    // (B,3)
  const float* gt;         // (B,G,gt_stride)
  const uint32_t* priorities;  // (B,N) or nullptr
  unsigned long long seed;
  ATLevel lvl[SDET_MAX_LEVELS];
  double base[SDET_MAX_LEVELS][kMaxA * 4];
  int num_levels, A, B, G, gt_stride;
  int N, S;                // anchors / cells per image over all levels
  float border, neg_thr, pos_thr, min_pos_thr;
  int image_anchor, fg_quota, k_pow2;
  // workspace
  float* best;             // (B,N) max IoU, -2 = outside the image
  int* which;              // (B,N) argmax gt (index into the compacted gts)
  signed char* label;      // (B,N) -1 / 0 / 1 before sub-sampling
  int* gt_max;             // (B,G) float bits
  unsigned long long* thr; // (B,2) smallest kept key for fg / bg (0 = keep all)
  // outputs
  float* cls_label;        // (B, A*S)
  float* reg_target;       // (B, 4A, S)
  float* reg_weight;
};

// Compacts image b's gt rows with x1 != -1 (order kept) into shared memory; returns their number.
__device__ int load_gts(const ATParams& p, int b, float4* s_gt, int* s_n) {
  if (threadIdx.x == 0) {
    int n = 0;
    for (int g = 0; g < p.G; ++g) {
      const float* r = p.gt + ((size_t)b * p.G + g) * p.gt_stride;
      if (r[0] != -1.0f) s_gt[n++] = make_float4(r[0], r[1], r[2], r[3]);   // detection_input.py:531-535
    }
    *s_n = n;
  }
  __syncthreads();
  return *s_n;
}

// anchor n -> level, cell, a; returns the float64 box (v_all_anchor / h_all_anchor, :405-441)
__device__ __forceinline__ void anchor_box(const ATParams& p, bool vertical, int n, int& l, int& cell, int& a,
                                           double (&e)[4]) {
  l = 0;
  for (int i = 0; i < p.num_levels; ++i) {
    const int c = p.lvl[i].dim_short * p.lvl[i].dim_long;
    if (n < (p.lvl[i].cell_off + c) * p.A) { l = i; break; }
  }
  const ATLevel& L = p.lvl[l];
  const int r = n - L.cell_off * p.A;
  cell = r / p.A;
  a = r - cell * p.A;
  const int fw = vertical ? L.dim_short : L.dim_long;
  const int y = cell / fw, x = cell - y * fw;
  const float sx = __fmul_rn((float)x, (float)L.stride), sy = __fmul_rn((float)y, (float)L.stride);
  e[0] = (double)sx + p.base[l][a * 4 + 0];
  e[1] = (double)sy + p.base[l][a * 4 + 1];
  e[2] = (double)sx + p.base[l][a * 4 + 2];
  e[3] = (double)sy + p.base[l][a * 4 + 3];
}

__device__ __forceinline__ bool inside_image(const ATParams& p, const double (&e)[4], float h, float w) {
  // _gather_valid_anchor :508-516; `w + allowed_border` is a float32 sum (np.float32 + python int)
  const double lo = -(double)p.border;
  return e[0] >= lo && e[1] >= lo && e[2] < (double)__fadd_rn(w, p.border) && e[3] < (double)__fadd_rn(h, p.border);
}

__device__ __forceinline__ float fminr(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float fmaxr(float a, float b) { return a < b ? b : a; }

// bbox_overlaps_cython (operator_py/cython/bbox.pyx:32-73) with the generated C's double promotions.
__device__ __forceinline__ float iou_cython(const float (&b)[4], const float4 q, float q_area) {
  const float iw = (float)((double)__fsub_rn(fminr(b[2], q.z), fmaxr(b[0], q.x)) + 1.0);
  if (!(iw > 0)) return 0.f;
  const float ih = (float)((double)__fsub_rn(fminr(b[3], q.w), fmaxr(b[1], q.y)) + 1.0);
  if (!(ih > 0)) return 0.f;
  const float inter = __fmul_rn(iw, ih);
  const float ua = (float)((((double)__fsub_rn(b[2], b[0]) + 1.0) * ((double)__fsub_rn(b[3], b[1]) + 1.0) +
                            (double)q_area) - (double)inter);
  return __fdiv_rn(inter, ua);
}
__device__ __forceinline__ float gt_area(const float4 q) {
  return (float)(((double)__fsub_rn(q.z, q.x) + 1.0) * ((double)__fsub_rn(q.w, q.y) + 1.0));
}

__global__ void __launch_bounds__(256) anchor_iou_kernel(const __grid_constant__ ATParams p) {
  __shared__ float4 s_gt[kMaxGt];
  __shared__ float s_area[kMaxGt];
  __shared__ int s_max[kMaxGt];
  __shared__ int s_n;
  const int b = blockIdx.y;
  const int ng = load_gts(p, b, s_gt, &s_n);
  for (int g = threadIdx.x; g < ng; g += blockDim.x) {
    s_area[g] = gt_area(s_gt[g]);
    s_max[g] = 0;
  }
  __syncthreads();
  const float h = p.im_info[b * 3], w = p.im_info[b * 3 + 1];
  const bool vertical = h >= w;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < p.N) {
    int l, cell, a;
    double e[4];
    anchor_box(p, vertical, n, l, cell, a, e);
    float best = -2.f;
    int which = 0;
    if (inside_image(p, e, h, w)) {
      const float bf[4] = {(float)e[0], (float)e[1], (float)e[2], (float)e[3]};
      best = 0.f;
      for (int g = 0; g < ng; ++g) {
        const float ov = iou_cython(bf, s_gt[g], s_area[g]);
        if (g == 0 || ov > best) {  // np.argmax: first maximum
          best = ov;
          which = g;
        }
        if (ov > 0.f) atomicMax(&s_max[g], __float_as_int(ov));
      }
    }
    p.best[(size_t)b * p.N + n] = best;
    p.which[(size_t)b * p.N + n] = which;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < ng; g += blockDim.x)
    if (s_max[g] > 0) atomicMax(&p.gt_max[(size_t)b * p.G + g], s_max[g]);
}

__global__ void __launch_bounds__(256) anchor_label_kernel(const __grid_constant__ ATParams p) {
  __shared__ float4 s_gt[kMaxGt];
  __shared__ float s_area[kMaxGt];
  __shared__ float s_max[kMaxGt];
  __shared__ int s_n;
  const int b = blockIdx.y;
  const int ng = load_gts(p, b, s_gt, &s_n);
  for (int g = threadIdx.x; g < ng; g += blockDim.x) {
    s_area[g] = gt_area(s_gt[g]);
    s_max[g] = __int_as_float(p.gt_max[(size_t)b * p.G + g]);
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const float best = p.best[(size_t)b * p.N + n];
  signed char label = -1;
  if (best != -2.f) {
    if (ng == 0) {
      label = 0;  // :472-474
    } else {
      const float h = p.im_info[b * 3], w = p.im_info[b * 3 + 1];
      int l, cell, a;
      double e[4];
      anchor_box(p, h >= w, n, l, cell, a, e);
      const float bf[4] = {(float)e[0], (float)e[1], (float)e[2], (float)e[3]};
      bool hit = false;  // (overlaps == gt_max_overlaps) & (overlaps >= min_pos_thr), any gt (:466-467)
      for (int g = 0; g < ng; ++g) {
        const float ov = iou_cython(bf, s_gt[g], s_area[g]);
        hit |= (ov == s_max[g]) && (ov >= p.min_pos_thr);
      }
      if (best < p.neg_thr) label = 0;
      if (hit) label = 1;
      if (best >= p.pos_thr) label = 1;
    }
  }
  p.label[(size_t)b * p.N + n] = label;
}

__device__ __forceinline__ uint32_t priority_of(const ATParams& p, int b, int n) {
  if (p.priorities) return __ldg(p.priorities + (size_t)b * p.N + n);
  curandStatePhilox4_32_10_t st;
  curand_init(p.seed, (unsigned long long)b * p.N + n, 0ull, &st);
  return curand(&st);
}
__device__ __forceinline__ unsigned long long sample_key(const ATParams& p, int b, int n) {
  return ((unsigned long long)priority_of(p, b, n) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)n);
}

__global__ void __launch_bounds__(kTopkThreads) anchor_quota_kernel(const __grid_constant__ ATParams p) {
  extern __shared__ unsigned long long s_sel[];
  __shared__ uint32_t s_hist[sdet::kRadixBins];
  __shared__ int s_cnt[2];
  const int b = blockIdx.x;
  const signed char* lab = p.label + (size_t)b * p.N;
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int c1 = 0, c0 = 0;
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
    const signed char v = lab[i];
    c1 += v == 1;
    c0 += v == 0;
  }
  for (int o = 16; o; o >>= 1) {
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_cnt[0], c1);
    atomicAdd(&s_cnt[1], c0);
  }
  __syncthreads();
  const int nfg = s_cnt[0], nbg = s_cnt[1];
  unsigned long long thr_fg = 0ull, thr_bg = 0ull;
  auto threshold = [&](signed char value, int count, int quota) -> unsigned long long {
    if (count <= quota) return 0ull;        // nothing to disable
    if (quota <= 0) return ~0ull;           // everything disabled
    auto key_at = [&](int i) -> uint64_t { return lab[i] == value ? sample_key(p, b, i) : 0ull; };
    sdet::block_topk_sorted(p.N, quota, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), p.k_pow2);
    const unsigned long long t = s_sel[quota - 1];
    __syncthreads();
    return t;
  };
  thr_fg = threshold(1, nfg, p.fg_quota);
  const int kept_fg = nfg < p.fg_quota ? nfg : p.fg_quota;
  thr_bg = threshold(0, nbg, p.image_anchor - kept_fg);
  if (threadIdx.x == 0) {
    p.thr[b * 2] = thr_fg;
    p.thr[b * 2 + 1] = thr_bg;
  }
}

// One thread per OUTPUT label element (a, s): coalesced stores of the (A, sum HW) / (4A, sum HW) planes.
__global__ void __launch_bounds__(256) anchor_write_kernel(const __grid_constant__ ATParams p) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.A * p.S) return;
  const int a = t / p.S, s = t - a * p.S;
  const int n = s * p.A + a;  // anchors are ordered (level, cell, a) and levels are contiguous in s
  float label = (float)p.label[(size_t)b * p.N + n];
  if (label >= 0.f) {
    const unsigned long long thr = p.thr[b * 2 + (label == 1.f ? 0 : 1)];
    if (thr != 0ull && sample_key(p, b, n) < thr) label = -1.f;
  }
  p.cls_label[(size_t)b * p.A * p.S + t] = label;
  float tg[4] = {0.f, 0.f, 0.f, 0.f};
  const float wt = label == 1.f ? 1.f : 0.f;
  if (label == 1.f) {
    const float h = p.im_info[b * 3], w = p.im_info[b * 3 + 1];
    int l2, cell, a2;
    double e[4];
    anchor_box(p, h >= w, n, l2, cell, a2, e);
    // the matched gt: which[] indexes the compacted list, so walk the valid rows
    int want = p.which[(size_t)b * p.N + n];
    const float* g = nullptr;
    for (int i = 0; i < p.G; ++i) {
      const float* r = p.gt + ((size_t)b * p.G + i) * p.gt_stride;
      if (r[0] != -1.0f && want-- == 0) { g = r; break; }
    }
    // nonlinear_transform (operator_py/bbox_transform.py:52-78): anchor float64, gt float32
    const double ew = e[2] - e[0] + 1.0, eh = e[3] - e[1] + 1.0;
    const double ecx = e[0] + 0.5 * (ew - 1.0), ecy = e[1] + 0.5 * (eh - 1.0);
    const float gw = __fadd_rn(__fsub_rn(g[2], g[0]), 1.0f), gh = __fadd_rn(__fsub_rn(g[3], g[1]), 1.0f);
    const float gcx = __fadd_rn(g[0], __fmul_rn(0.5f, __fsub_rn(gw, 1.0f)));
    const float gcy = __fadd_rn(g[1], __fmul_rn(0.5f, __fsub_rn(gh, 1.0f)));
    tg[0] = (float)(((double)gcx - ecx) / (ew + 1e-14));
    tg[1] = (float)(((double)gcy - ecy) / (eh + 1e-14));
    tg[2] = (float)log((double)gw / ew);
    tg[3] = (float)log((double)gh / eh);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const size_t o = ((size_t)b * p.A * 4 + (size_t)a * 4 + k) * p.S + s;
    p.reg_target[o] = tg[k];
    p.reg_weight[o] = wt;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" size_t sdet_anchor_target_workspace(int B, int total_anchors, int max_gt) {
  if (B <= 0 || total_anchors <= 0 || max_gt < 0) return 0;
  const size_t bn = (size_t)B * total_anchors;
  return align_up(bn * 4, 256) * 2 + align_up(bn, 256) + align_up((size_t)B * (max_gt > 0 ? max_gt : 1) * 4, 256) +
         align_up((size_t)B * 16, 256);
}

extern "C" int sdet_anchor_target(const float* im_info, const float* gt_bbox, int gt_stride, float* cls_label,
                                  float* reg_target, float* reg_weight, int B, int G, int num_levels,
                                  const int* strides, const int* shorts, const int* longs, const double* scales,
                                  int num_scales, const double* aspects, int num_aspects, float allowed_border,
                                  float neg_thr, float pos_thr, float min_pos_thr, int image_anchor, int fg_quota,
                                  const uint32_t* priorities, unsigned long long seed, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  SDET_REQUIRE(im_info && cls_label && reg_target && reg_weight && strides && shorts && longs && scales && aspects &&
               workspace, "NULL argument");
  SDET_REQUIRE(gt_bbox || G == 0, "gt_bbox is NULL");
  SDET_REQUIRE(B > 0 && G >= 0 && num_levels >= 1 && num_levels <= SDET_MAX_LEVELS, "bad shape");
  SDET_REQUIRE(gt_stride == 4 || gt_stride == 5, "gt_bbox rows must have 4 or 5 columns");
  SDET_REQUIRE(image_anchor > 0 && fg_quota >= 0 && fg_quota <= image_anchor, "bad sampling quota");
  const int A = num_scales * num_aspects;
  if (A <= 0 || A > kMaxA) return sdet::fail(SDET_ERR_UNSUPPORTED, "1..%d anchors per cell supported", kMaxA);
  if (G > kMaxGt) return sdet::fail(SDET_ERR_UNSUPPORTED, "more than %d gt rows per image", kMaxGt);
  if (image_anchor > 16384) return sdet::fail(SDET_ERR_UNSUPPORTED, "image_anchor > 16384");
  ATParams p{};
  int cells = 0;
  for (int l = 0; l < num_levels; ++l) {
    SDET_REQUIRE(strides[l] > 0 && shorts[l] > 0 && longs[l] > 0, "level %d: bad stride / size", l);
    p.lvl[l] = ATLevel{strides[l], shorts[l], longs[l], cells};
    cells += shorts[l] * longs[l];
    // base_anchor (detection_input.py:377-403): float64, aspect-major, np.round = round-half-even
    const double side = (double)strides[l], ctr = 0.5 * (side - 1.0);
    int k = 0;
    for (int i = 0; i < num_aspects; ++i) {
      const double wr = std::nearbyint(std::sqrt(side * side / aspects[i]));
      const double hr = std::nearbyint(wr * aspects[i]);
      for (int j = 0; j < num_scales; ++j, ++k) {
        const double ws = wr * scales[j], hs = hr * scales[j];
        p.base[l][k * 4 + 0] = ctr - 0.5 * (ws - 1.0);
        p.base[l][k * 4 + 1] = ctr - 0.5 * (hs - 1.0);
        p.base[l][k * 4 + 2] = ctr + 0.5 * (ws - 1.0);
        p.base[l][k * 4 + 3] = ctr + 0.5 * (hs - 1.0);
      }
    }
  }
  const long long N64 = (long long)cells * A;
  if (N64 > 0x7FFFFFF0ll) return sdet::fail(SDET_ERR_UNSUPPORTED, "too many anchors");
  const int N = (int)N64;
  if (workspace_bytes < sdet_anchor_target_workspace(B, N, G))
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", sdet_anchor_target_workspace(B, N, G));
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = static_cast<char*>(workspace);
  const size_t bn = (size_t)B * N;
  p.best = reinterpret_cast<float*>(ws); ws += align_up(bn * 4, 256);
  p.which = reinterpret_cast<int*>(ws); ws += align_up(bn * 4, 256);
  p.label = reinterpret_cast<signed char*>(ws); ws += align_up(bn, 256);
  p.gt_max = reinterpret_cast<int*>(ws); ws += align_up((size_t)B * (G > 0 ? G : 1) * 4, 256);
  p.thr = reinterpret_cast<unsigned long long*>(ws);
  p.im_info = im_info; p.gt = gt_bbox; p.priorities = priorities; p.seed = seed;
  p.num_levels = num_levels; p.A = A; p.B = B; p.G = G; p.gt_stride = gt_stride;
  p.N = N; p.S = cells;
  p.border = allowed_border; p.neg_thr = neg_thr; p.pos_thr = pos_thr; p.min_pos_thr = min_pos_thr;
  p.image_anchor = image_anchor; p.fg_quota = fg_quota;
  p.k_pow2 = sdet::next_pow2(image_anchor);
  p.cls_label = cls_label; p.reg_target = reg_target; p.reg_weight = reg_weight;
  if (G > 0) SDET_CUDA(cudaMemsetAsync(p.gt_max, 0, (size_t)B * G * 4, st));
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)B);
  anchor_iou_kernel<<<grid, 256, 0, st>>>(p);
  SDET_LAUNCH_CHECK("anchor_iou_kernel");
  anchor_label_kernel<<<grid, 256, 0, st>>>(p);
  SDET_LAUNCH_CHECK("anchor_label_kernel");
  const size_t smem = (size_t)p.k_pow2 * 8;
  if (smem > 48 * 1024)  // per device and cheap: set on every launch, no process-wide cache
    SDET_CUDA(cudaFuncSetAttribute(anchor_quota_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  anchor_quota_kernel<<<(unsigned)B, kTopkThreads, smem, st>>>(p);
  SDET_LAUNCH_CHECK("anchor_quota_kernel");
  anchor_write_kernel<<<grid, 256, 0, st>>>(p);
  SDET_LAUNCH_CHECK("anchor_write_kernel");
  return SDET_OK;
}
