/*
 * simpledet_b200 — C ABI of the B200-native (sm_100a) detection hot path.
 *
 * This header is the drop-in boundary.  Every entry point replaces one MXNet
 * operator of tusen-ai/simpledet; the reference interface it replaces is cited as
 * `file:line` relative to the reference checkout.  Conventions (all entry points):
 *
 *   - plain pointers + sizes, no framework types.  Pointers named `d_*` or documented as
 *     "device" are CUDA device pointers on the current device; the caller owns every buffer,
 *     the library allocates nothing on the device and keeps no global state.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     stream-ordered on it; no entry point synchronises the device unless documented.
 *   - return value: SDET_OK (0) or an sdet_status error code; sdet_last_error() returns a
 *     thread-local human-readable message.  The library never aborts (the reference's
 *     CHECK / LOG(FATAL) macros become error codes).
 *   - fp32 tensors, dense row-major, shapes exactly as the reference operators define them.
 */
#ifndef SIMPLEDET_B200_H_
#define SIMPLEDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sdet_status {
  SDET_OK = 0,
  SDET_ERR_INVALID_ARG = 1, /* bad shape / null pointer / parameter out of range */
  SDET_ERR_UNSUPPORTED = 2, /* valid for the reference, outside this build's limits */
  SDET_ERR_CUDA = 3,        /* a CUDA runtime call failed; message has cudaGetErrorString */
  SDET_ERR_WORKSPACE = 4    /* caller-provided workspace too small */
} sdet_status;

#define SDET_MAX_LEVELS 8
#define SDET_MAX_POOLED 32 /* max pooled_size per axis handled by the RoIAlign kernels */

/* ABI version of this header; bumped on any signature change. */
int sdet_abi_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* sdet_last_error(void);
/* Number of kernels this library has launched in this process (all threads); used by bench.py's
 * `gpu_launches` claim. */
uint64_t sdet_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * _contrib_ROIAlign_v2   (operator_cxx/contrib/roi_align_v2.cc:170-228, kernel
 *                         roi_align_v2-inl.h:61-153, driver :157-195)
 *   data  (B,C,H,W) device, rois (B,N,4) device [x1,y1,x2,y2] image px, image = n / N.
 *   out, argmax_x, argmax_y (B,N,C,PH,PW) device.  argmax_x/argmax_y may both be NULL
 *   (inference: they are hidden outputs, roi_align_v2.cc:175-178) — then they are not written.
 * ------------------------------------------------------------------------------------------ */
int sdet_roi_align_v2_forward(const float* data, const float* rois, float* out, float* argmax_x,
                              float* argmax_y, int B, int N, int C, int H, int W, int pooled_h,
                              int pooled_w, float spatial_scale, void* stream);

/* _backward_ROIAlign_v2  (operator_cxx/contrib/roi_align_v2.cu:17-85 kernel, :88-143 driver).
 *   ograd, argmax_x, argmax_y (B,N,C,PH,PW); grad_data (B,C,H,W).
 *   accumulate = 0 -> kWriteTo (grad_data is zero-filled first, :130-133), 1 -> kAddTo.
 *   grad_rois (B,N,4) may be NULL; when given it is zero-filled (:139-141). */
int sdet_roi_align_v2_backward(const float* ograd, const float* argmax_x, const float* argmax_y,
                               float* grad_data, float* grad_rois, int B, int N, int C, int H,
                               int W, int pooled_h, int pooled_w, int accumulate, void* stream);

/* Fused FPN RoIAlign: replaces  assign_layer_fpn (models/FPN/assign_layer_fpn.py:17-40) /
 * mxnext.tvm.fpn_roi_assign  +  num_levels x _contrib_ROIAlign_v2  +  add_n
 * (models/FPN/builder.py:573-605).  Each roi is sampled on its assigned level only; the result
 * equals the reference's sum over levels because a zeroed roi yields an all-zero output
 * (roi_align_v2-inl.h:111-117).
 *   feats[l]  device pointer to level l data (B,C,H[l],W[l]);  strides[l] the level's stride
 *   (spatial_scale = 1/stride, must be a power of two as in the reference's `2**lvl == s` test).
 *   feats/H/W/strides are HOST arrays of length num_levels (<= SDET_MAX_LEVELS).
 *   levels_out (B*N int32, device, may be NULL): assigned level index, -1 if no level matched.
 *   argmax_x/argmax_y as above (may be NULL). */
int sdet_fpn_roi_align_v2_forward(const float* const* feats, const int* H, const int* W,
                                  const int* strides, int num_levels, const float* rois,
                                  float* out, float* argmax_x, float* argmax_y,
                                  int32_t* levels_out, int B, int N, int C, int pooled_h,
                                  int pooled_w, int roi_canonical_scale, int roi_canonical_level,
                                  void* stream);

/* Backward of the fused op: scatters into the grad tensor of each roi's assigned level.
 *   grad_feats[l] device (B,C,H[l],W[l]); levels (B*N int32 device) as written by the forward. */
int sdet_fpn_roi_align_v2_backward(const float* ograd, const float* argmax_x,
                                   const float* argmax_y, const int32_t* levels,
                                   float* const* grad_feats, const int* H, const int* W,
                                   int num_levels, int B, int N, int C, int pooled_h, int pooled_w,
                                   int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------
 * ROIPooling_v1   (operator_cxx/roi_pooling_v1.cc:243, op roi_pooling_v1-inl.h:63-137,
 *                  kernels roi_pooling_v1.cu:49-113 / :116-152)
 *   data (B,C,H,W), rois (R,5) = [batch_index,x1,y1,x2,y2]; out, max_idx (R,C,PH,PW);
 *   max_idx holds the flat h*W+w index as float, -1 for empty bins.
 * ------------------------------------------------------------------------------------------ */
int sdet_roi_pooling_v1_forward(const float* data, const float* rois, float* out, float* max_idx,
                                int B, int R, int C, int H, int W, int pooled_h, int pooled_w,
                                float spatial_scale, void* stream);
int sdet_roi_pooling_v1_backward(const float* ograd, const float* max_idx, const float* rois,
                                 float* grad_data, float* grad_rois, int B, int R, int C, int H,
                                 int W, int pooled_h, int pooled_w, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMPLEDET_B200_H_ */
