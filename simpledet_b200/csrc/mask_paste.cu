// Test-time mask paste of Mask R-CNN (models/maskrcnn/utils.py:26-67 `segm_results`, called per image from
// models/maskrcnn/process_output.py:6-19 under mask_test.py:207): per detection, resize the 28x28 class mask to the
// detection's box, threshold, paste into the image, RLE-encode.  The reference does it with cv2.resize + a fresh
// im_h x im_w uint8 image + pycocotools per detection on the host (about 1 MB written and re-read per detection).
// Here the image is never materialised: the kernels emit the column-major positions at which the pasted mask flips
// ("mask paste").
//
//   pass 1  mask_paste_count_kernel   thread = (detection, image column): number of flips in its column
//   (host-side exclusive scan of the (N, im_w) counts: torch.cumsum in ops.segm_results, or the caller's own)
//   pass 2  mask_paste_write_kernel   same walk, positions written at the scanned offsets
//
// Traffic: 4*M*M bytes of mask per detection (L1-resident; every thread of a detection reads the same 3 KB) in,
// 4 bytes per flip out (a smooth 28x28 mask has 2-4 flips per column).  The walk itself is arithmetic: one double
// coordinate + one fma per pixel, (h_box x w_box) pixels per detection, so this is latency / issue bound, not an
// HBM-roofline kernel (NMS-like: report time, not GB/s).  The per-column body lives in mask_paste_core.cuh so the
// host emulation in tests/ runs the same source.
#include "common.cuh"
#include "mask_paste_core.cuh"

namespace {

constexpr int kPasteThreads = 128;

template <bool kWrite>
__global__ void __launch_bounds__(kPasteThreads)
mask_paste_kernel(const float* __restrict__ boxes, const int* __restrict__ cls, const float* __restrict__ masks, int K,
                  int M, int im_h, int im_w, int* __restrict__ col_counts, const long long* __restrict__ col_offsets,
                  int* __restrict__ positions) {
  sdet_paste::paste_thread(kWrite, blockIdx.y, blockIdx.x * kPasteThreads + threadIdx.x, boxes, cls, masks, K, M, im_h,
                           im_w, col_counts, col_offsets, positions);
}

int check_args(const float* boxes, const int* cls, const float* masks, int N, int K, int M, int im_h, int im_w) {
  SDET_REQUIRE(boxes && cls && masks, "NULL argument");
  SDET_REQUIRE(N > 0 && K > 0 && M > 0 && im_h > 0 && im_w > 0, "bad shape");
  if (M + 2 > sdet_paste::kMaxSide) return sdet::fail(SDET_ERR_UNSUPPORTED, "mask side %d > %d", M, sdet_paste::kMaxSide - 2);
  if ((long long)im_h * im_w >= (1ll << 31)) return sdet::fail(SDET_ERR_UNSUPPORTED, "image of %d x %d pixels", im_h, im_w);
  if (N > 65535) return sdet::fail(SDET_ERR_UNSUPPORTED, "more than 65535 detections per call");
  return SDET_OK;
}

}  // namespace

extern "C" int sdet_mask_paste_count(const float* boxes, const int* cls, const float* masks, int N, int K, int M,
                                     int im_h, int im_w, int* col_counts, void* stream) {
  if (int rc = check_args(boxes, cls, masks, N, K, M, im_h, im_w)) return rc;
  SDET_REQUIRE(col_counts, "NULL argument");
  dim3 grid((unsigned)((im_w + kPasteThreads - 1) / kPasteThreads), (unsigned)N);
  mask_paste_kernel<false><<<grid, kPasteThreads, 0, (cudaStream_t)stream>>>(boxes, cls, masks, K, M, im_h, im_w,
                                                                             col_counts, nullptr, nullptr);
  SDET_LAUNCH_CHECK("mask_paste_count_kernel");
  return SDET_OK;
}

extern "C" int sdet_mask_paste_write(const float* boxes, const int* cls, const float* masks, int N, int K, int M,
                                     int im_h, int im_w, const long long* col_offsets, int* positions, void* stream) {
  if (int rc = check_args(boxes, cls, masks, N, K, M, im_h, im_w)) return rc;
  SDET_REQUIRE(col_offsets && positions, "NULL argument");
  dim3 grid((unsigned)((im_w + kPasteThreads - 1) / kPasteThreads), (unsigned)N);
  mask_paste_kernel<true><<<grid, kPasteThreads, 0, (cudaStream_t)stream>>>(boxes, cls, masks, K, M, im_h, im_w, nullptr,
                                                                            col_offsets, positions);
  SDET_LAUNCH_CHECK("mask_paste_write_kernel");
  return SDET_OK;
}
