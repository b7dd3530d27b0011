/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * CPU restatement (plain C, fp32, same operation order and the same
 * float/double promotions) of the reference's RoI operators.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (simpledet_b200/) never does.
 *
 * Compile with -ffp-contract=off: the reference's CPU build has no FMA
 * contraction (x86-64 SSE), and the CUDA kernels are written with explicit
 * __fmul_rn/__fadd_rn so that both sides round identically.
 *
 * Parity pin: ROIPooling_v1 is pinned by the reference's docstring example
 * (roi_pooling_v1.cc:265-285, see tests/test_oracle_golden.py).  ROIAlign_v2
 * has no fixture anywhere in the reference; both operators
 * are pinned bit for bit against the reference's own sources (roi_align_v2.cc/.cu,
 * roi_pooling_v1.cc) compiled through oracle/shim: tests/test_oracle_ref_cxx.py,
 * vectors in tests/golden/reference_cxx_ops.npz.
 *
 * Functions follow (reference file:line, relative to /root/reference):
 *   oracle_roi_align_v2_forward   operator_cxx/contrib/roi_align_v2-inl.h:61-153
 *   oracle_roi_align_v2_backward  operator_cxx/contrib/roi_align_v2.cu:35-84,130-141
 *                                 (GPU semantics, sequential adds in index order)
 *   oracle_roi_pool_v1_forward    operator_cxx/roi_pooling_v1.cu:49-113 (== .cc:40-126)
 *   oracle_roi_pool_v1_backward   operator_cxx/roi_pooling_v1.cu:116-152
 *   oracle_fpn_assign_levels      models/FPN/assign_layer_fpn.py:17-40
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float fminf_(float a, float b) { return a < b ? a : b; }  /* mshadow_op::minimum */
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }  /* mshadow_op::maximum */
static inline int imin_(int a, int b) { return a < b ? a : b; }
static inline int imax_(int a, int b) { return a > b ? a : b; }

/* roi_align_v2-inl.h:61-153.  data (B,C,H,W), rois (B,N,4) -> out/argmax (B,N,C,PH,PW).
 * argmax_x / argmax_y may be NULL (inference: only `out` is visible, roi_align_v2.cc:175-178). */
void oracle_roi_align_v2_forward(const float* bottom_data, const float* rois, int B, int N, int C,
                                 int height, int width, int pooled_height, int pooled_width,
                                 float spatial_scale, float* top_data, float* argmax_x,
                                 float* argmax_y) {
  const long count = (long)B * N * C * pooled_height * pooled_width;
#pragma omp parallel for schedule(static)
  for (long index = 0; index < count; ++index) {
    int pw = (int)(index % pooled_width);
    int ph = (int)((index / pooled_width) % pooled_height);
    int c = (int)((index / pooled_width / pooled_height) % C);
    int n = (int)(index / pooled_width / pooled_height / C);
    const float* bottom_rois = rois + (long)n * 4;
    int roi_batch_ind = n / N; /* :77 */

    float roi_start_w = bottom_rois[0] * spatial_scale;
    float roi_start_h = bottom_rois[1] * spatial_scale;
    float roi_end_w = bottom_rois[2] * spatial_scale;
    float roi_end_h = bottom_rois[3] * spatial_scale;
    float roi_width = roi_end_w - roi_start_w;
    float roi_height = roi_end_h - roi_start_h;
    float bin_size_h = roi_height / (float)pooled_height;
    float bin_size_w = roi_width / (float)pooled_width;

    float hstart = (float)(ph)*bin_size_h;
    float wstart = (float)(pw)*bin_size_w;
    float hend = (float)(ph + 1) * bin_size_h;
    float wend = (float)(pw + 1) * bin_size_w;
    hstart = fminf_(fmaxf_(hstart + roi_start_h, 0.f), (float)(height - 1));
    hend = fminf_(fmaxf_(hend + roi_start_h, 0.f), (float)(height - 1));
    wstart = fminf_(fmaxf_(wstart + roi_start_w, 0.f), (float)(width - 1));
    wend = fminf_(fmaxf_(wend + roi_start_w, 0.f), (float)(width - 1));
    int is_empty = (hend <= hstart) || (wend <= wstart);

    float maxidx_x = -1.f, maxidx_y = -1.f, maxval = 0.f;
    if (!is_empty) {
      maxval = -FLT_MAX; /* mshadow::red::limits::MinValue<float>() */
      const float* data = bottom_data + ((long)roi_batch_ind * C + c) * height * width;
      /* :120-121 — the literal 3.0 is double: divide in double, narrow to float */
      float h_stride = (float)((double)(hend - hstart) / 3.0);
      float w_stride = (float)((double)(wend - wstart) / 3.0);
      /* :122-125 — `hend-h_stride` is a float expression, `+0.01` promotes to double */
      for (float h = hstart + h_stride; (double)h <= (double)(hend - h_stride) + 0.01;
           h += fmaxf_(h_stride, 0.01f)) {
        for (float w = wstart + w_stride; (double)w <= (double)(wend - w_stride) + 0.01;
             w += fmaxf_(w_stride, 0.01f)) {
          int hlow = imin_(imax_((int)floorf(h), 0), height - 1);
          int hhigh = imin_(imax_((int)ceilf(h), 0), height - 1);
          int wleft = imin_(imax_((int)floorf(w), 0), width - 1);
          int wright = imin_(imax_((int)ceilf(w), 0), width - 1);
          int topleft = hlow * width + wleft;
          int topright = hlow * width + wright;
          int bottomleft = hhigh * width + wleft;
          int bottomright = hhigh * width + wright;
          float alpha = (hlow == hhigh) ? 0.5f : (h - (float)hlow) / (float)(hhigh - hlow);
          float beta = (wleft == wright) ? 0.5f : (w - (float)wleft) / (float)(wright - wleft);
          float value = (1.f - alpha) * (1.f - beta) * data[topleft] +
                        alpha * (1.f - beta) * data[bottomleft] +
                        (1.f - alpha) * beta * data[topright] + alpha * beta * data[bottomright];
          if (value > maxval) {
            maxval = value;
            maxidx_x = w;
            maxidx_y = h;
          }
        }
      }
    }
    top_data[index] = maxval;
    if (argmax_x) argmax_x[index] = maxidx_x;
    if (argmax_y) argmax_y[index] = maxidx_y;
  }
}

/* roi_align_v2.cu:35-84 with the kWriteTo zero fill of :130-133 (accumulate=0) or kAddTo
 * (accumulate=1).  The GPU kernel's atomics are unordered; the oracle adds in index order. */
void oracle_roi_align_v2_backward(const float* top_diff, const float* argmax_x,
                                  const float* argmax_y, int B, int N, int C, int height, int width,
                                  int pooled_height, int pooled_width, int accumulate,
                                  float* bottom_diff) {
  if (!accumulate) memset(bottom_diff, 0, sizeof(float) * (size_t)B * C * height * width);
  const long count = (long)B * N * C * pooled_height * pooled_width;
  for (long index = 0; index < count; ++index) {
    int c = (int)((index / pooled_width / pooled_height) % C);
    int n = (int)(index / pooled_width / pooled_height / C);
    int roi_batch_ind = n / N;
    float* offset_bottom_diff = bottom_diff + ((long)roi_batch_ind * C + c) * height * width;
    float a_x = argmax_x[index];
    float a_y = argmax_y[index];
    if (a_x != -1.f && a_y != -1.f) {
      int hlow = imin_(imax_((int)floorf(a_y), 0), height - 1);
      int hhigh = imin_(imax_((int)ceilf(a_y), 0), height - 1);
      int wleft = imin_(imax_((int)floorf(a_x), 0), width - 1);
      int wright = imin_(imax_((int)ceilf(a_x), 0), width - 1);
      float alpha = (hlow == hhigh) ? 0.5f : (a_y - (float)hlow) / (float)(hhigh - hlow);
      float beta = (wleft == wright) ? 0.5f : (a_x - (float)wleft) / (float)(wright - wleft);
      float g = top_diff[index];
      offset_bottom_diff[hlow * width + wleft] += g * (1.f - alpha) * (1.f - beta);
      offset_bottom_diff[hlow * width + wright] += g * (1.f - alpha) * beta;
      offset_bottom_diff[hhigh * width + wleft] += g * alpha * (1.f - beta);
      offset_bottom_diff[hhigh * width + wright] += g * alpha * beta;
    }
  }
}

/* roi_pooling_v1.cu:49-113.  rois (R,5) = [batch, x1, y1, x2, y2]; out/argmax (R,C,PH,PW). */
void oracle_roi_pool_v1_forward(const float* bottom_data, const float* bottom_rois, int R, int C,
                                int height, int width, int pooled_height, int pooled_width,
                                float spatial_scale, float* top_data, float* argmax_data) {
  const long count = (long)R * C * pooled_height * pooled_width;
#pragma omp parallel for schedule(static)
  for (long index = 0; index < count; ++index) {
    int pw = (int)(index % pooled_width);
    int ph = (int)((index / pooled_width) % pooled_height);
    int c = (int)((index / pooled_width / pooled_height) % C);
    int n = (int)(index / pooled_width / pooled_height / C);
    const float* r = bottom_rois + (long)n * 5;
    int roi_batch_ind = (int)r[0];
    int roi_start_w = (int)round(r[1] * spatial_scale);
    int roi_start_h = (int)round(r[2] * spatial_scale);
    int roi_end_w = (int)round(r[3] * spatial_scale);
    int roi_end_h = (int)round(r[4] * spatial_scale);
    int roi_width = imax_(roi_end_w - roi_start_w + 1, 1);
    int roi_height = imax_(roi_end_h - roi_start_h + 1, 1);
    float bin_size_h = (float)roi_height / (float)pooled_height;
    float bin_size_w = (float)roi_width / (float)pooled_width;
    int hstart = (int)floorf((float)ph * bin_size_h);
    int wstart = (int)floorf((float)pw * bin_size_w);
    int hend = (int)ceilf((float)(ph + 1) * bin_size_h);
    int wend = (int)ceilf((float)(pw + 1) * bin_size_w);
    hstart = imin_(imax_(hstart + roi_start_h, 0), height);
    hend = imin_(imax_(hend + roi_start_h, 0), height);
    wstart = imin_(imax_(wstart + roi_start_w, 0), width);
    wend = imin_(imax_(wend + roi_start_w, 0), width);
    int is_empty = (hend <= hstart) || (wend <= wstart);
    float maxval = is_empty ? 0.f : -FLT_MAX;
    int maxidx = -1;
    const float* data = bottom_data + ((long)roi_batch_ind * C + c) * height * width;
    for (int h = hstart; h < hend; ++h)
      for (int w = wstart; w < wend; ++w) {
        int bi = h * width + w;
        if (data[bi] > maxval) {
          maxval = data[bi];
          maxidx = bi;
        }
      }
    top_data[index] = maxval;
    if (argmax_data) argmax_data[index] = (float)maxidx;
  }
}

/* roi_pooling_v1.cu:116-152 (scatter by stored argmax), zero fill per roi_pooling_v1-inl.h:125-127. */
void oracle_roi_pool_v1_backward(const float* top_diff, const float* argmax_data,
                                 const float* bottom_rois, int R, int B, int C, int height,
                                 int width, int pooled_height, int pooled_width, int accumulate,
                                 float* bottom_diff) {
  if (!accumulate) memset(bottom_diff, 0, sizeof(float) * (size_t)B * C * height * width);
  const long count = (long)R * C * pooled_height * pooled_width;
  for (long index = 0; index < count; ++index) {
    int c = (int)((index / pooled_width / pooled_height) % C);
    int n = (int)(index / pooled_width / pooled_height / C);
    int roi_batch_ind = (int)bottom_rois[(long)n * 5];
    int argmax = (int)argmax_data[index];
    if (argmax != -1)
      bottom_diff[((long)roi_batch_ind * C + c) * height * width + argmax] += top_diff[index];
  }
}

/* models/FPN/assign_layer_fpn.py:17-40.  mx.nd float32 arithmetic:
 *   area = (x2-x1+1)*(y2-y1+1); lvl = clip(floor(lvl0 + log2(sqrt(area)/scale0 + 1e-6)), kmin, kmax)
 * `1e-6` is added to a float32 NDArray => float32 scalar.  Returns the level (log2 stride). */
void oracle_fpn_assign_levels(const float* rois, long n, float scale0, float lvl0, float k_min,
                              float k_max, int32_t* levels) {
  for (long i = 0; i < n; ++i) {
    const float* r = rois + i * 4;
    float area = (r[2] - r[0] + 1.f) * (r[3] - r[1] + 1.f);
    float scale = sqrtf(area);
    float t = floorf(lvl0 + log2f(scale / scale0 + 1e-6f));
    t = fminf_(fmaxf_(t, k_min), k_max);
    /* NaN area (x2<x1-1 & ...) -> sqrt NaN -> clip semantics of mx.nd.clip keep NaN;
     * `2**NaN -> uint8` is 0 in numpy => no level matches => roi zeroed on every level. */
    levels[i] = (t != t) ? -1 : (int32_t)t;
  }
}
