// RetinaNet test-time ops: _contrib_GenAnchor (operator_cxx/contrib/generate_anchor.cu) and
// _contrib_GenProposalRetina (generate_proposal_retina.cu:307-469), used per FPN level by
// models/retinanet/builder.py:358-386.
//
// The reference decodes every one of the 720*H*W (anchor, class) pairs into a (count,5) buffer and
// thrust-sorts all of them (12 M keys on P3) to keep 1000.  Here one coalesced pass over cls_prob
// histograms the scores of the pairs that survive `score > thresh` and the min-size test (4608 bins of 2^13 ulps
// above the threshold), a second pass writes out as 64-bit keys (score, reference index) only the pairs in the
// bins that can still reach the top `pre` - a few thousand keys whatever the score distribution - and one CTA per
// image radix-selects and sorts the top `pre` of those and decodes just the winners.  Zeroed pairs never need materialising: with thresh >= 0 every survivor scores > 0,
// so they sort after all survivors and contribute all-zero output rows.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "topk.cuh"

namespace {

using sdet::kTopkThreads;
constexpr int kMaxBaseAnchors = 32;

__device__ __forceinline__ float fmin_ref(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float fmax_ref(float a, float b) { return a < b ? b : a; }

struct AnchorGrid {
  double base[kMaxBaseAnchors * 4];
  int A, H, W, stride;
  float* out;
};

__global__ void __launch_bounds__(256) anchor_grid_kernel(const __grid_constant__ AnchorGrid g) {
  const int total = g.H * g.W * g.A * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 3, index = i >> 2;
    const int a = index % g.A, w = (index / g.A) % g.W, h = index / g.A / g.W;
    g.out[i] = (float)(g.base[a * 4 + j] + (double)(((j & 1) ? h : w) * g.stride));  // generate_anchor.cu:72-79
  }
}

struct RetinaParams {
  const float* cls_prob;   // (B, A*K, H, W)
  const float* bbox_pred;  // (B, A*4, H, W)
  const float* im_info;    // (B, 3)
  const float* anchors;    // (H*W*A, 4)
  size_t anchor_batch_stride;
  float mean[4], stdv[4];
  float thresh;
  int A, K, H, W, min_size, pre, pre_pow2, out_channel;
  unsigned long long* cand;  // (B, count) keys
  int* cand_count;           // (B)
  uint32_t* hist;            // (B, kScoreBins) survivors per score bin
  uint32_t thresh_ord;       // score_to_sortable(thresh)
  float* out;                // (B, pre_param, 4)
  float* out_score;          // (B, pre_param, out_channel)
  int out_rows;
};

// Decode of one (cell, anchor): BBoxPredKernel (generate_proposal_retina.cu:96-150).
__device__ __forceinline__ float4 retina_decode(const RetinaParams& p, int b, int a_box, int h, int w, float im_h,
                                                float im_w) {
  const int HW = p.H * p.W, r = h * p.W + w;
  const float* an = p.anchors + (size_t)b * p.anchor_batch_stride + ((size_t)r * p.A + a_box) * 4;
  const float ax1 = __ldg(an), ay1 = __ldg(an + 1), ax2 = __ldg(an + 2), ay2 = __ldg(an + 3);
  const float* dl = p.bbox_pred + ((size_t)b * p.A * 4 + (size_t)a_box * 4) * HW + r;
  const float dx = __fadd_rn(__fmul_rn(__ldg(dl), p.stdv[0]), p.mean[0]);
  const float dy = __fadd_rn(__fmul_rn(__ldg(dl + HW), p.stdv[1]), p.mean[1]);
  const float dw = __fadd_rn(__fmul_rn(__ldg(dl + 2 * HW), p.stdv[2]), p.mean[2]);
  const float dh = __fadd_rn(__fmul_rn(__ldg(dl + 3 * HW), p.stdv[3]), p.mean[3]);
  const float width = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f), height = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
  const float ctr_x = __fadd_rn(ax1, __fmul_rn(0.5f, __fsub_rn(width, 1.0f)));
  const float ctr_y = __fadd_rn(ay1, __fmul_rn(0.5f, __fsub_rn(height, 1.0f)));
  const float pcx = __fadd_rn(__fmul_rn(dx, width), ctr_x), pcy = __fadd_rn(__fmul_rn(dy, height), ctr_y);
  const float pw = __fmul_rn(expf(dw), width), ph = __fmul_rn(expf(dh), height);
  const float hw_ = __fmul_rn(0.5f, __fsub_rn(pw, 1.0f)), hh_ = __fmul_rn(0.5f, __fsub_rn(ph, 1.0f));
  const float mx = __fsub_rn(im_w, 1.0f), my = __fsub_rn(im_h, 1.0f);
  float4 o;
  o.x = fmax_ref(fmin_ref(__fsub_rn(pcx, hw_), mx), 0.0f);
  o.y = fmax_ref(fmin_ref(__fsub_rn(pcy, hh_), my), 0.0f);
  o.z = fmax_ref(fmin_ref(__fadd_rn(pcx, hw_), mx), 0.0f);
  o.w = fmax_ref(fmin_ref(__fadd_rn(pcy, hh_), my), 0.0f);
  return o;
}

constexpr int kScoreBins = 4608;  // bins of 2^13 ulps above the threshold: (0.05, 1] spans 4506 of them
constexpr int kScoreShift = 13;

__device__ __forceinline__ int score_bin(const RetinaParams& p, float s) {  // s > thresh
  const uint32_t d = (sdet::score_to_sortable(s) - p.thresh_ord) >> kScoreShift;
  return d < (uint32_t)kScoreBins ? (int)d : kScoreBins - 1;  // anything further up shares the top bin
}

// Is (image b, flat cls_prob index i) a survivor?  FilterBoxKernel :218 zeroes score <= thresh, :205-216 boxes
// smaller than min_size.
__device__ __forceinline__ bool retina_survivor(const RetinaParams& p, int b, int i, int HW, int AK, float im_h,
                                                float im_w, float min_size, float s, unsigned long long& key) {
  if (!(s > p.thresh)) return false;
  const int a = i / HW, r = i - a * HW, h = r / p.W, w = r - h * p.W;
  if (p.min_size > 0 || min_size > 1.f) {  // a decoded box is at least 1 pixel wide: nothing to test otherwise
    const float4 bx = retina_decode(p, b, a / p.K, h, w, im_h, im_w);
    const float iw = __fadd_rn(__fsub_rn(bx.z, bx.x), 1.0f), ih = __fadd_rn(__fsub_rn(bx.w, bx.y), 1.0f);
    if (iw < min_size || ih < min_size) return false;
  }
  key = sdet::make_key(s, (uint32_t)(r * AK + a));  // reference index (h*W+w)*AK + a
  return true;
}

// Pass 1: stream cls_prob in memory order, histogram the survivors' scores (per-CTA shared histogram, merged once).
__global__ void __launch_bounds__(256) retina_hist_kernel(const __grid_constant__ RetinaParams p) {
  __shared__ uint32_t s_h[kScoreBins];
  const int b = blockIdx.y;
  const int HW = p.H * p.W, AK = p.A * p.K, count = AK * HW;
  const float im_h = __ldg(p.im_info + b * 3), im_w = __ldg(p.im_info + b * 3 + 1);
  const float min_size = __fmul_rn((float)p.min_size, __ldg(p.im_info + b * 3 + 2));
  const float* sc = p.cls_prob + (size_t)b * count;
  for (int i = threadIdx.x; i < kScoreBins; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  const int span = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += span) {
    const float s = __ldg(sc + i);
    unsigned long long key;
    if (retina_survivor(p, b, i, HW, AK, im_h, im_w, min_size, s, key)) atomicAdd(&s_h[score_bin(p, s)], 1u);
  }
  __syncthreads();
  uint32_t* gh = p.hist + (size_t)b * kScoreBins;
  for (int i = threadIdx.x; i < kScoreBins; i += blockDim.x)
    if (s_h[i]) atomicAdd(gh + i, s_h[i]);
}

// Pass 2: find the lowest bin that can still hold one of the top `pre` survivors, stream cls_prob again and keep
// the survivors at or above it as keys (warp-aggregated append).
__global__ void __launch_bounds__(256) retina_candidates_kernel(const __grid_constant__ RetinaParams p) {
  __shared__ uint32_t s_part[256];
  __shared__ int s_cut;
  const int b = blockIdx.y;
  const int HW = p.H * p.W, AK = p.A * p.K, count = AK * HW;
  {  // suffix sums of the histogram, 18 bins per thread: cut = highest bin with (survivors in bins >= cut) >= pre
    constexpr int kPer = kScoreBins / 256;
    const uint32_t* gh = p.hist + (size_t)b * kScoreBins;
    uint32_t mine = 0;
    for (int j = 0; j < kPer; ++j) mine += gh[threadIdx.x * kPer + j];
    s_part[threadIdx.x] = mine;
    if (threadIdx.x == 0) s_cut = 0;
    __syncthreads();
    uint32_t above = 0;  // survivors in the bins of the threads above this one
    for (int t = threadIdx.x + 1; t < 256; ++t) above += s_part[t];
    if (above < (uint32_t)p.pre && above + mine >= (uint32_t)p.pre) {  // the cut lies in this thread's bins
      int j = kPer - 1;
      for (; j > 0; --j) {
        above += gh[threadIdx.x * kPer + j];
        if (above >= (uint32_t)p.pre) break;
      }
      s_cut = threadIdx.x * kPer + j;
    }
    __syncthreads();
  }
  const int cut = s_cut;  // 0 when there are fewer than `pre` survivors: keep them all
  const float im_h = __ldg(p.im_info + b * 3), im_w = __ldg(p.im_info + b * 3 + 1);
  const float min_size = __fmul_rn((float)p.min_size, __ldg(p.im_info + b * 3 + 2));
  const float* sc = p.cls_prob + (size_t)b * count;
  unsigned long long* cand = p.cand + (size_t)b * count;
  const int lane = threadIdx.x & 31;
  const int span = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x; i0 < count; i0 += span) {  // i0 is warp-uniform modulo lane
    const int i = i0 + threadIdx.x;
    bool keep = false;
    unsigned long long key = 0;
    if (i < count) {
      const float s = __ldg(sc + i);
      keep = retina_survivor(p, b, i, HW, AK, im_h, im_w, min_size, s, key) && score_bin(p, s) >= cut;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(p.cand_count + b, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (keep) cand[base + __popc(m & ((1u << lane) - 1))] = key;
    }
  }
}

// Pass 2: one CTA per image selects + sorts the top keys, decodes the winners, writes every output row.
__global__ void __launch_bounds__(kTopkThreads) retina_select_kernel(const __grid_constant__ RetinaParams p) {
  extern __shared__ unsigned long long s_sel[];
  __shared__ uint32_t s_hist[sdet::kRadixBins];
  const int b = blockIdx.x;
  const int HW = p.H * p.W, AK = p.A * p.K, count = AK * HW;
  const int n = p.cand_count[b];
  const int k = n < p.pre ? n : p.pre;
  const unsigned long long* cand = p.cand + (size_t)b * count;
  auto key_at = [&](int i) -> uint64_t { return cand[i]; };
  sdet::block_topk_sorted(n, k, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), p.pre_pow2);
  const float im_h = __ldg(p.im_info + b * 3), im_w = __ldg(p.im_info + b * 3 + 1);
  const int oc = p.out_channel;
  for (int j = threadIdx.x; j < p.out_rows; j += blockDim.x) {
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    float* srow = p.out_score + ((size_t)b * p.out_rows + j) * oc;
    for (int c = 0; c < oc; ++c) srow[c] = 0.f;
    if (j < k) {
      const uint64_t key = reinterpret_cast<const uint64_t*>(s_sel)[j];
      const int idx = (int)sdet::key_index(key);
      const int a = idx % AK, r = idx / AK, h = r / p.W, w = r - h * p.W;
      bx = retina_decode(p, b, a / p.K, h, w, im_h, im_w);
      const int cls = idx % p.K;  // ReorderProposalsKernel :253: order_i % num_class
      const int cid = min(oc - 1, cls + 1);
      srow[cid] = sdet::key_score(key);
    }
    *reinterpret_cast<float4*>(p.out + ((size_t)b * p.out_rows + j) * 4) = bx;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" int sdet_gen_anchor(float* out, int H, int W, int feature_stride, const double* scales, int num_scales,
                               const double* ratios, int num_ratios, void* stream) {
  SDET_REQUIRE(out && scales && ratios, "NULL argument");
  SDET_REQUIRE(H > 0 && W > 0 && num_scales > 0 && num_ratios > 0, "bad shape");
  if (num_scales * num_ratios > kMaxBaseAnchors)
    return sdet::fail(SDET_ERR_UNSUPPORTED, "more than %d base anchors", kMaxBaseAnchors);
  AnchorGrid g{};
  // gen_anchor_utils::GenerateAnchors in double (generate_anchor-inl.h:139-183): ratios outer, scales inner
  const double b2 = feature_stride - 1.0f;
  int k = 0;
  for (int j = 0; j < num_ratios; ++j)
    for (int s = 0; s < num_scales; ++s) {
      const double w = b2 - 0.0 + 1.0f, h = b2 - 0.0 + 1.0f;
      const double x_ctr = 0.0 + 0.5 * (w - 1.0f), y_ctr = 0.0 + 0.5 * (h - 1.0f);
      const double size_ratios = (w * h) / ratios[j];
      const double new_w = std::rint(std::sqrt(size_ratios)) * scales[s];
      const double new_h = std::rint((new_w / scales[s] * ratios[j])) * scales[s];
      g.base[k * 4 + 0] = x_ctr - 0.5f * (new_w - 1.0f);
      g.base[k * 4 + 1] = y_ctr - 0.5f * (new_h - 1.0f);
      g.base[k * 4 + 2] = x_ctr + 0.5f * (new_w - 1.0f);
      g.base[k * 4 + 3] = y_ctr + 0.5f * (new_h - 1.0f);
      ++k;
    }
  g.A = k; g.H = H; g.W = W; g.stride = feature_stride; g.out = out;
  const int total = H * W * k * 4;
  anchor_grid_kernel<<<std::min((total + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(g);
  SDET_LAUNCH_CHECK("anchor_grid_kernel");
  return SDET_OK;
}

extern "C" size_t sdet_gen_proposal_retina_workspace(int B, int AK, int H, int W) {
  if (B <= 0 || AK <= 0 || H <= 0 || W <= 0) return 0;
  return align_up((size_t)B * 4 + (size_t)B * kScoreBins * 4, 256) + (size_t)B * AK * H * W * 8;
}

extern "C" int sdet_gen_proposal_retina(const float* cls_prob, const float* bbox_pred, const float* im_info,
                                        const float* anchors, float* out, float* out_score, int B, int AK, int H,
                                        int W, int num_anchors, int feature_stride, int rpn_pre_nms_top_n,
                                        int rpn_min_size, float thresh, const float* anchor_mean,
                                        const float* anchor_std, int iou_loss, int output_one_hot,
                                        int batch_wise_anchor, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  (void)feature_stride;  // only feeds real_height/real_width, whose mask is commented out upstream (:143-147)
  SDET_REQUIRE(cls_prob && bbox_pred && im_info && anchors && out && out_score && workspace, "NULL argument");
  SDET_REQUIRE(B > 0 && AK > 0 && H > 0 && W > 0 && num_anchors > 0 && rpn_pre_nms_top_n > 0, "bad shape");
  SDET_REQUIRE(AK % num_anchors == 0, "cls_prob channels (%d) not a multiple of num_anchors (%d)", AK, num_anchors);
  const int K = AK / num_anchors;
  if (iou_loss) return sdet::fail(SDET_ERR_UNSUPPORTED, "iou_loss: the reference kernel indexes deltas out of range for num_class > 1");
  if (batch_wise_anchor && K != 1)
    return sdet::fail(SDET_ERR_UNSUPPORTED, "batch_wise_anchor with num_class > 1: reference offset i*count*4 overruns the anchors");
  if (!(thresh >= 0.f)) return sdet::fail(SDET_ERR_UNSUPPORTED, "thresh must be >= 0");
  const size_t count = (size_t)AK * H * W;
  if (count > 0x7FFFFFFFull) return sdet::fail(SDET_ERR_UNSUPPORTED, "more than 2^31 (anchor,class) pairs");
  if (workspace_bytes < sdet_gen_proposal_retina_workspace(B, AK, H, W))
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", sdet_gen_proposal_retina_workspace(B, AK, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  RetinaParams p{};
  p.cls_prob = cls_prob; p.bbox_pred = bbox_pred; p.im_info = im_info; p.anchors = anchors;
  p.anchor_batch_stride = batch_wise_anchor ? count * 4 : 0;
  for (int i = 0; i < 4; ++i) { p.mean[i] = anchor_mean ? anchor_mean[i] : 0.f; p.stdv[i] = anchor_std ? anchor_std[i] : 1.f; }
  p.thresh = thresh;
  p.A = num_anchors; p.K = K; p.H = H; p.W = W; p.min_size = rpn_min_size;
  p.pre = (int)std::min<size_t>((size_t)rpn_pre_nms_top_n, count);
  p.pre_pow2 = sdet::next_pow2(p.pre);
  p.out_channel = output_one_hot ? K + 1 : 1;
  const size_t head = (size_t)B * 4 + (size_t)B * kScoreBins * 4;
  p.cand_count = static_cast<int*>(workspace);
  p.hist = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + (size_t)B * 4);
  p.thresh_ord = 0;  // filled below (host copy of score_to_sortable)
  {
    uint32_t u;
    std::memcpy(&u, &thresh, 4);
    p.thresh_ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  p.cand = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + align_up(head, 256));
  p.out = out; p.out_score = out_score; p.out_rows = rpn_pre_nms_top_n;
  SDET_CUDA(cudaMemsetAsync(workspace, 0, head, st));
  dim3 grid((unsigned)std::min<size_t>((count + 255) / 256, 148 * 8), (unsigned)B);
  retina_hist_kernel<<<grid, 256, 0, st>>>(p);
  SDET_LAUNCH_CHECK("retina_hist_kernel");
  retina_candidates_kernel<<<grid, 256, 0, st>>>(p);
  SDET_LAUNCH_CHECK("retina_candidates_kernel");
  const size_t smem = (size_t)sdet::next_pow2(p.pre) * 8;
  if (smem > 200 * 1024) return sdet::fail(SDET_ERR_UNSUPPORTED, "rpn_pre_nms_top_n too large for shared-memory select");
  if (smem > 48 * 1024)  // per device and cheap: set on every launch, no process-wide cache
    SDET_CUDA(cudaFuncSetAttribute(retina_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  retina_select_kernel<<<(unsigned)B, kTopkThreads, smem, st>>>(p);
  SDET_LAUNCH_CHECK("retina_select_kernel");
  return SDET_OK;
}
