/* ORACLE - test infrastructure only (see oracle/__init__.py).
 *
 * Test-time mask paste of Mask R-CNN, restated the way the reference runs it - materialising everything:
 *   models/maskrcnn/utils.py:7-23   expand_boxes
 *   models/maskrcnn/utils.py:26-67  segm_results: zero ring, cv2.resize, `> 0.5`, paste, mask_util.encode
 * and the two third-party pieces it calls, neither under /root/reference:
 *   cv2.resize(float32, dsize) (opencv-python, INTER_LINEAR): restated from its observable behaviour and PINNED bit
 *     for bit against the cv2 that is installed here (4.13, x86-64 wheel, IPP on - the same kind of wheel the
 *     reference's `pip install opencv-python` gives): tests/test_mask_paste_host.py::test_resize_against_cv2.
 *     With cv2.ipp.setUseIPP(False) OpenCV's own code path differs from this in the last bit (it computes
 *     s0*(1-t) + s1*t without fusing and the coordinate in float32); the IPP path is what a user of the wheel gets.
 *   pycocotools mask.encode (cocoapi common/maskApi.c rleEncode: runs of the column-major image, first run counts
 *     zeros; rleToString lives in np_ops.rle_to_string): restated from the published algorithm, PARITY UNPINNED
 *     (pycocotools is not installed and the reference holds no vector for it).
 * The whole function is additionally compared with the reference's own segm_results run unmodified on top of the
 * installed cv2 and a stand-in `pycocotools.mask.encode` (tests/golden/make_golden_mask_paste.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* one axis of cv2.resize's linear interpolation (IPP path): double coordinate, float weight, borders clamp to a
 * single sample with weight 0 */
static void axis_coord(int d, int dn, int sn, int* s0, int* s1, float* t) {
  double f = ((double)d + 0.5) * ((double)sn / (double)dn) - 0.5;
  double fl = floor(f);
  int s = (int)fl;
  float frac = (float)(f - fl);
  if (s < 0) { s = 0; frac = 0.f; }
  if (s >= sn - 1) { s = sn - 1; frac = 0.f; }
  *s0 = s;
  *s1 = s + 1 < sn ? s + 1 : sn - 1;
  *t = frac;
}

/* dst (dh, dw) = cv2.resize(src (sh, sw) float32, (dw, dh)) */
void oracle_resize_linear_f32(const float* src, int sh, int sw, float* dst, int dh, int dw) {
  for (int y = 0; y < dh; ++y) {
    int r0, r1;
    float ty;
    axis_coord(y, dh, sh, &r0, &r1, &ty);
    for (int x = 0; x < dw; ++x) {
      int c0, c1;
      float tx;
      axis_coord(x, dw, sw, &c0, &c1, &tx);
      float a = src[r0 * sw + c0], b = src[r0 * sw + c1];
      float c = src[r1 * sw + c0], d = src[r1 * sw + c1];
      float top = fmaf(tx, b - a, a);
      float bot = fmaf(tx, d - c, c);
      dst[(long)y * dw + x] = fmaf(ty, bot - top, top);
    }
  }
}

/* utils.py:7-23 + :35-36: the expanded box truncated to int32.  box float32 x1,y1,x2,y2. */
void oracle_expand_box_int(const float* box, int M, int* out) {
  float scale = (float)(((double)M + 2.0) / (double)M);
  float w_half = (box[2] - box[0]) * 0.5f;
  float h_half = (box[3] - box[1]) * 0.5f;
  float x_c = (box[2] + box[0]) * 0.5f;
  float y_c = (box[3] + box[1]) * 0.5f;
  w_half *= scale;
  h_half *= scale;
  out[0] = (int)(double)(x_c - w_half);
  out[2] = (int)(double)(x_c + w_half);
  out[1] = (int)(double)(y_c - h_half);
  out[3] = (int)(double)(y_c + h_half);
}

/* utils.py:39-59 for one detection: im_mask (im_h, im_w) uint8, row-major, zero-filled by the caller; scratch holds
 * (M+2)^2 + w*h floats.  Returns 0, or -1 where the reference's slice assignment cannot work (the box has no pixel
 * inside the image): nothing is pasted then. */
int oracle_segm_paste(const float* box, const float* mask, int M, int im_h, int im_w, uint8_t* im_mask, float* scratch) {
  int rb[4];
  oracle_expand_box_int(box, M, rb);
  int S = M + 2;
  float* padded = scratch;
  memset(padded, 0, sizeof(float) * S * S);
  for (int r = 0; r < M; ++r) memcpy(padded + (r + 1) * S + 1, mask + r * M, sizeof(float) * M);
  int w = rb[2] - rb[0] + 1, h = rb[3] - rb[1] + 1;
  if (w < 1) w = 1;
  if (h < 1) h = 1;
  int x_0 = rb[0] > 0 ? rb[0] : 0, x_1 = rb[2] + 1 < im_w ? rb[2] + 1 : im_w;
  int y_0 = rb[1] > 0 ? rb[1] : 0, y_1 = rb[3] + 1 < im_h ? rb[3] + 1 : im_h;
  if (rb[2] < rb[0] || rb[3] < rb[1]) return 0; /* empty slices on both sides of the assignment */
  if (x_1 <= x_0 || y_1 <= y_0) return -1;
  float* rs = scratch + S * S;
  oracle_resize_linear_f32(padded, S, S, rs, h, w);
  for (int y = y_0; y < y_1; ++y)
    for (int x = x_0; x < x_1; ++x) im_mask[(long)y * im_w + x] = rs[(long)(y - rb[1]) * w + (x - rb[0])] > 0.5f;
  return 0;
}

/* maskApi.c rleEncode for one mask: im (h, w) row-major here, walked in column-major order.  counts has room for
 * h*w + 1 entries; returns how many were written. */
long oracle_rle_encode(const uint8_t* im, int h, int w, uint32_t* counts) {
  long k = 0;
  uint32_t c = 0;
  uint8_t p = 0;
  for (int x = 0; x < w; ++x)
    for (int y = 0; y < h; ++y) {
      uint8_t v = im[(long)y * w + x];
      if (v != p) { counts[k++] = c; c = 0; p = v; }
      ++c;
    }
  counts[k++] = c;
  return k;
}
