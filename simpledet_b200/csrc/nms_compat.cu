// `_nms`: the one real C ABI the reference ships (operator_py/cython/gpu_nms.hpp:1-2, implementation
// operator_py/cython/nms_kernel.cu:91-144, caller gpu_nms.pyx:16-31 which pre-sorts by score).  Kept as a
// drop-in compatibility export over the library's own NMS kernels (sdet_nms_sorted: upper-triangular IoU
// bitmask + on-device greedy scan) - same symbol, same argument meaning, host pointers in, blocking, device
// memory allocated and freed inside the call like the reference does.  Suppression is `IoU > thresh`
// (nms_kernel.cu:71), IoU with the +1 pixel convention; keep_out receives positions in the caller's order.
#include <cstdio>
#include <vector>

#include "common.cuh"

extern "C" void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                     float nms_overlap_thresh, int device_id) {
  if (num_out) *num_out = 0;
  if (!keep_out || !num_out || !boxes_host || boxes_num <= 0 || boxes_dim < 4) return;
  auto fail = [](const char* what, cudaError_t e) {  // the reference's CUDA_CHECK only prints (nms_kernel.cu:12-19)
    std::fprintf(stderr, "_nms: %s: %s\n", what, e == cudaSuccess ? sdet_last_error() : cudaGetErrorString(e));
  };
  cudaError_t e = cudaSetDevice(device_id);
  if (e != cudaSuccess) return fail("cudaSetDevice", e);
  // the kernels read (x1, y1, x2, y2, score) rows of 5 floats; other widths are repacked (the score is unused)
  std::vector<float> packed;
  const float* src = boxes_host;
  if (boxes_dim != 5) {
    packed.assign((size_t)boxes_num * 5, 0.f);
    for (int i = 0; i < boxes_num; ++i)
      for (int k = 0; k < (boxes_dim < 5 ? boxes_dim : 5); ++k) packed[(size_t)i * 5 + k] = boxes_host[(size_t)i * boxes_dim + k];
    src = packed.data();
  }
  const size_t ws_bytes = sdet_nms_workspace(1, boxes_num);
  float* dets = nullptr;
  int* keep = nullptr;
  void* ws = nullptr;
  cudaStream_t st = nullptr;
  const size_t det_bytes = sizeof(float) * 5 * (size_t)boxes_num;
  if ((e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaMalloc(&dets, det_bytes)) == cudaSuccess && (e = cudaMalloc(&keep, sizeof(int) * ((size_t)boxes_num + 1))) == cudaSuccess &&
      (e = cudaMalloc(&ws, ws_bytes ? ws_bytes : 16)) == cudaSuccess &&
      (e = cudaMemcpyAsync(dets, src, det_bytes, cudaMemcpyHostToDevice, st)) == cudaSuccess) {
    int* nkeep = keep + boxes_num;
    if (sdet_nms_sorted(dets, nullptr, 1, boxes_num, nms_overlap_thresh, /*ge=*/0, keep, nkeep, ws, ws_bytes, st) != SDET_OK) {
      fail("sdet_nms_sorted", cudaSuccess);
    } else {
      int n = 0;
      if ((e = cudaMemcpyAsync(&n, nkeep, sizeof(int), cudaMemcpyDeviceToHost, st)) == cudaSuccess &&
          (e = cudaStreamSynchronize(st)) == cudaSuccess && n > 0 &&
          (e = cudaMemcpyAsync(keep_out, keep, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st)) == cudaSuccess)
        e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) *num_out = n;
    }
  }
  if (e != cudaSuccess) fail("cuda", e);
  cudaFree(dets);
  cudaFree(keep);
  cudaFree(ws);
  cudaStreamDestroy(st);
}
