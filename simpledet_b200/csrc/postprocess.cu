// Cross-level top proposals and test-time per-class NMS, on the device.
//
//   sdet_get_top_proposal   models/FPN/get_top_proposal.py:15-40 (CustomOp `get_top_proposal`,
//                           = mxnext.tvm.get_top_proposal at models/FPN/builder.py:319-321)
//   sdet_multiclass_nms     detection_test.py:233-260 `do_nms` (per class: score > min_det_score,
//                           operator_py/nms.py:41-75 `nms`, IoU <= thr kept), all (image, class)
//                           problems of a batch in three launches instead of a Python loop inside
//                           a multiprocessing pool.
#include "common.cuh"
#include "topk.cuh"

namespace {

using sdet::kTopkThreads;

__global__ void __launch_bounds__(kTopkThreads)
top_proposal_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int M,
                    const int top_n, const int k_pow2, float* __restrict__ out_boxes,
                    float* __restrict__ out_scores) {
  extern __shared__ unsigned long long s_sel[];
  __shared__ uint32_t s_hist[sdet::kRadixBins];
  const int b = blockIdx.x;
  const float* sc = scores + (size_t)b * M;
  auto key_at = [&](int i) -> uint64_t { return sdet::make_key(__ldg(sc + i), (uint32_t)i); };
  sdet::block_topk_sorted(M, top_n, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), k_pow2);
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)b * M;
  for (int j = threadIdx.x; j < top_n; j += blockDim.x) {
    const uint64_t key = reinterpret_cast<const uint64_t*>(s_sel)[j];
    const int i = (int)sdet::key_index(key);
    reinterpret_cast<float4*>(out_boxes)[(size_t)b * top_n + j] = __ldg(bx + i);
    out_scores[(size_t)b * top_n + j] = __ldg(sc + i);
  }
}

// One CTA per (image, class) problem, N <= 2048 candidates: threshold, sort descending
// (ties: lower roi index first), write dets (P, n_pad, 5) + counts (P).
__global__ void __launch_bounds__(1024)
class_sort_kernel(const float* __restrict__ cls_score, const float* __restrict__ bbox, const int N,
                  const int K, const int bbox_dim, const int first_class, const float min_score,
                  const int n_pad, float* __restrict__ dets, int* __restrict__ counts,
                  int* __restrict__ src_index) {
  extern __shared__ unsigned long long s_keys[];
  __shared__ int s_valid;
  const int ncls = K - first_class;
  const int p = blockIdx.x, b = p / ncls, cid = first_class + p % ncls;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  // survivors are appended (warp-aggregated) in any order - the sort fixes it - so that only next_pow2(survivors)
  // keys are sorted, not n_pad: most classes of an image keep a handful of the N candidates
  const int lane = threadIdx.x & 31;
  for (int i0 = 0; i0 < N; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    uint64_t key = 0ull;
    bool keep = false;
    if (i < N) {
      const float s = __ldg(cls_score + ((size_t)b * N + i) * K + cid);
      if (s > min_score) {  // detection_test.py:243
        key = sdet::make_key(s, (uint32_t)i);
        keep = true;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_valid, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (keep) s_keys[base + __popc(m & ((1u << lane) - 1))] = key;
    }
  }
  __syncthreads();
  int sort_n = 1;
  while (sort_n < s_valid) sort_n <<= 1;
  for (int i = s_valid + threadIdx.x; i < sort_n; i += blockDim.x) s_keys[i] = 0ull;
  __syncthreads();
  sdet::block_bitonic_sort_desc(reinterpret_cast<uint64_t*>(s_keys), sort_n);
  const int nv = s_valid;
  if (threadIdx.x == 0) counts[p] = nv;
  for (int j = threadIdx.x; j < n_pad; j += blockDim.x) {
    float* o = dets + ((size_t)p * n_pad + j) * 5;
    if (j < nv) {
      const uint64_t key = s_keys[j];
      const int i = (int)sdet::key_index(key);
      const float* bp = bbox + ((size_t)b * N + i) * bbox_dim + (bbox_dim == 4 ? 0 : cid * 4);
      o[0] = __ldg(bp); o[1] = __ldg(bp + 1); o[2] = __ldg(bp + 2); o[3] = __ldg(bp + 3);
      o[4] = sdet::key_score(key);
      if (src_index) src_index[(size_t)p * n_pad + j] = i;
    } else {
      o[0] = o[1] = o[2] = o[3] = o[4] = 0.f;
      if (src_index) src_index[(size_t)p * n_pad + j] = -1;
    }
  }
}

// detection_test.py:268-291: the kept detections of all classes of an image, listed class by class in
// NMS order, `sorted(result, key=score)[-max_det:]` — a stable ascending sort, so among equal scores
// the LATER list entries survive.  Key = (sortable score, list position): the top `max_det` keys in
// descending order are the reference's slice read backwards.  One CTA per image.
__global__ void __launch_bounds__(kTopkThreads)
final_dets_kernel(const float* __restrict__ dets, const int* __restrict__ keep, const int* __restrict__ nkeep,
                  const int ncls, const int n_pad, const int max_det, const int k_pow2,
                  float* __restrict__ out, int* __restrict__ out_count, const int xyxy, const int descending) {
  extern __shared__ unsigned long long s_sel[];
  __shared__ uint32_t s_hist[sdet::kRadixBins];
  __shared__ int s_base[1025];  // list position of every class's first kept detection
  const int b = blockIdx.x;
  const int* nk = nkeep + (size_t)b * ncls;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int c = 0; c < ncls; ++c) {
      s_base[c] = acc;
      acc += nk[c];
    }
    s_base[ncls] = acc;
  }
  __syncthreads();
  const int total = s_base[ncls];
  const int k = min(max_det, total);
  auto key_at = [&](int i) -> uint64_t {
    const int c = i / n_pad, j = i - c * n_pad;
    if (j >= __ldg(nk + c)) return 0ull;  // not a kept slot: below every real key
    const size_t p = (size_t)b * ncls + c;
    const int row = __ldg(keep + p * n_pad + j);
    const float sc = __ldg(dets + (p * n_pad + row) * 5 + 4);
    return ((uint64_t)sdet::score_to_sortable(sc) << 32) | (uint64_t)(uint32_t)(s_base[c] + j + 1);
  };
  if (k > 0) sdet::block_topk_sorted(ncls * n_pad, k, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), k_pow2);
  if (threadIdx.x == 0) out_count[b] = k;
  for (int r = threadIdx.x; r < max_det; r += blockDim.x) {
    float* o = out + ((size_t)b * max_det + r) * 6;
    if (r >= k) {
      o[0] = o[1] = o[2] = o[3] = o[4] = 0.f;
      o[5] = -1.f;
      continue;
    }
    const uint64_t key = s_sel[descending ? r : k - 1 - r];  // ascending score like the reference's slice
    const int pos = (int)(uint32_t)key - 1;
    int c = 0;
    for (int lo = 0, hi = ncls; lo < hi;) {  // largest c with s_base[c] <= pos
      const int mid = (lo + hi) >> 1;
      if (s_base[mid + 1] <= pos) lo = mid + 1; else hi = mid;
      c = lo;
    }
    const size_t p = (size_t)b * ncls + c;
    const float* d = dets + (p * n_pad + __ldg(keep + p * n_pad + (pos - s_base[c]))) * 5;
    // COCO box: x, y, w = x2 - x1 + 1, h = y2 - y1 + 1 (detection_test.py:277-280)
    o[0] = d[0]; o[1] = d[1];
    o[2] = xyxy ? d[2] : __fadd_rn(__fsub_rn(d[2], d[0]), 1.f);
    o[3] = xyxy ? d[3] : __fadd_rn(__fsub_rn(d[3], d[1]), 1.f);
    o[4] = d[4];
    o[5] = (float)c;
  }
}

}  // namespace

extern "C" int sdet_final_detections_ex(const float* dets, const int* keep, const int* nkeep, int B, int num_classes,
                                        int n_pad, int max_det, float* out, int* out_count, int xyxy, int descending,
                                        void* stream) {
  SDET_REQUIRE(dets && keep && nkeep && out && out_count, "NULL argument");
  SDET_REQUIRE(B > 0 && num_classes > 0 && n_pad > 0 && max_det > 0, "bad shape");
  if (num_classes > 1024) return sdet::fail(SDET_ERR_UNSUPPORTED, "more than 1024 classes");
  const int k_pow2 = sdet::next_pow2(max_det);
  const size_t smem = (size_t)k_pow2 * 8;
  if (smem > 160 * 1024) return sdet::fail(SDET_ERR_UNSUPPORTED, "max_det too large");
  // per device and cheap: set on every launch that needs it instead of caching in a process-wide static
  if (smem > 48 * 1024)
    SDET_CUDA(cudaFuncSetAttribute(final_dets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  final_dets_kernel<<<(unsigned)B, kTopkThreads, smem, (cudaStream_t)stream>>>(
      dets, keep, nkeep, num_classes, n_pad, max_det, k_pow2, out, out_count, xyxy != 0, descending != 0);
  SDET_LAUNCH_CHECK("final_dets_kernel");
  return SDET_OK;
}

extern "C" int sdet_final_detections(const float* dets, const int* keep, const int* nkeep, int B, int num_classes,
                                     int n_pad, int max_det, float* out, int* out_count, void* stream) {
  return sdet_final_detections_ex(dets, keep, nkeep, B, num_classes, n_pad, max_det, out, out_count, 0, 0, stream);
}

extern "C" int sdet_get_top_proposal(const float* boxes, const float* scores, float* out_boxes,
                                     float* out_scores, int B, int M, int top_n, void* stream) {
  SDET_REQUIRE(boxes && scores && out_boxes && out_scores, "NULL argument");
  SDET_REQUIRE(B > 0 && M > 0 && top_n > 0 && top_n <= M, "need 0 < top_n <= M");
  SDET_REQUIRE((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(out_boxes)) % 16 == 0,
               "boxes must be 16-byte aligned");
  const int k_pow2 = sdet::next_pow2(top_n);
  const size_t smem = (size_t)k_pow2 * 8;
  if (smem > 200 * 1024) return sdet::fail(SDET_ERR_UNSUPPORTED, "top_n too large");
  if (smem > 48 * 1024)
    SDET_CUDA(cudaFuncSetAttribute(top_proposal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  top_proposal_kernel<<<(unsigned)B, kTopkThreads, smem, (cudaStream_t)stream>>>(
      boxes, scores, M, top_n, k_pow2, out_boxes, out_scores);
  SDET_LAUNCH_CHECK("top_proposal_kernel");
  return SDET_OK;
}

extern "C" size_t sdet_multiclass_nms_workspace(int B, int N, int K, int first_class) {
  if (B <= 0 || N <= 0 || K <= first_class) return 0;
  return sdet_nms_workspace(B * (K - first_class), sdet::next_pow2(N));
}

extern "C" int sdet_multiclass_nms(const float* cls_score, const float* bbox, int B, int N, int K,
                                   int bbox_dim, int first_class, float min_det_score, float nms_thresh,
                                   float* dets, int* counts, int* keep, int* nkeep, int* src_index,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  SDET_REQUIRE(cls_score && bbox && dets && counts && keep && nkeep && workspace, "NULL argument");
  SDET_REQUIRE(B > 0 && N > 0 && K > first_class && first_class >= 0, "bad shape");
  SDET_REQUIRE(bbox_dim == 4 || bbox_dim == 4 * K, "bbox last dim must be 4 or 4*K");
  const int n_pad = sdet::next_pow2(N);
  if (n_pad > 4096) return sdet::fail(SDET_ERR_UNSUPPORTED, "more than 4096 candidates per image");
  const int P = B * (K - first_class);
  class_sort_kernel<<<(unsigned)P, 1024, (size_t)n_pad * 8, (cudaStream_t)stream>>>(
      cls_score, bbox, N, K, bbox_dim, first_class, min_det_score, n_pad, dets, counts, src_index);
  SDET_LAUNCH_CHECK("class_sort_kernel");
  // operator_py/nms.py:72 keeps ovr <= thresh  <=>  suppress ovr > thresh (ge = 0)
  return sdet_nms_sorted(dets, counts, P, n_pad, nms_thresh, 0, keep, nkeep, workspace, workspace_bytes,
                         stream);
}
