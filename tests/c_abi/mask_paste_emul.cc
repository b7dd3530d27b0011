// Host emulation of mask_paste.cu's two kernels (test infrastructure): the kernels' whole thread body lives in
// simpledet_b200/csrc/mask_paste_core.cuh and is written without CUDA built-ins, so this file runs THAT SOURCE for
// every (blockIdx.y, blockIdx.x * 128 + threadIdx.x) the launch would create, one after another.  Compiled by
// tests/test_mask_paste_host.py with g++ -O2 -ffp-contract=off (the library is nvcc -fmad=false: fused only where the
// source says fmaf).  Same argument lists as sdet_mask_paste_count / sdet_mask_paste_write minus the stream.
#include "mask_paste_core.cuh"

extern "C" int emul_mask_paste_count(const float* boxes, const int* cls, const float* masks, int N, int K, int M,
                                     int im_h, int im_w, int* col_counts) {
  const int threads = 128, gx = (im_w + threads - 1) / threads;
  for (int by = 0; by < N; ++by)
    for (int bx = 0; bx < gx; ++bx)
      for (int t = 0; t < threads; ++t)
        sdet_paste::paste_thread(false, by, bx * threads + t, boxes, cls, masks, K, M, im_h, im_w, col_counts, nullptr,
                                 nullptr);
  return 0;
}

extern "C" int emul_mask_paste_write(const float* boxes, const int* cls, const float* masks, int N, int K, int M,
                                     int im_h, int im_w, const long long* col_offsets, int* positions) {
  const int threads = 128, gx = (im_w + threads - 1) / threads;
  for (int by = 0; by < N; ++by)
    for (int bx = gx - 1; bx >= 0; --bx)       // any order must do: threads are independent
      for (int t = threads - 1; t >= 0; --t)
        sdet_paste::paste_thread(true, by, bx * threads + t, boxes, cls, masks, K, M, im_h, im_w, nullptr, col_offsets,
                                 positions);
  return 0;
}
