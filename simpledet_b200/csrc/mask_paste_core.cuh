// Mask paste (models/maskrcnn/utils.py:26-67 `segm_results`): the per-column body shared by the two kernels of
// mask_paste.cu.  It is written against SDET_HD only (no CUDA built-ins), so tests/c_abi/mask_paste_emul.cc can run
// the very same source on the host, thread by thread, against the oracle (the kernels were added after the round's
// GPU budget was spent; this is how their arithmetic was checked).
//
// What the reference does per detection: expand the box by (M+2)/M (float32), truncate to int32, cv2.resize the
// zero-ringed (M+2)x(M+2) mask of the detection's class to the box's integer (w, h), threshold `> 0.5`, paste the
// part inside the image into an im_h x im_w uint8 image, RLE-encode it in column-major order (pycocotools).
// Here the image is never materialised: a thread owns one image column of one detection, walks the box rows of that
// column and emits the flat column-major positions p = x*im_h + y at which the pasted mask changes value; run
// lengths are differences of consecutive positions.
//
// cv2.resize(float32, INTER_LINEAR) as the opencv-python wheel computes it (its IPP path; pinned bit for bit in
// tests/test_mask_paste_host.py): per axis, in double, f = (d + 0.5) * (src / dst) - 0.5, s = floor(f), t = float(f
// - s), clamped to (s, t) = (0, 0) below the first and (src-1, 0) from the last source sample; value =
// fma(ty, bot - top, top) with top / bot = fma(tx, b - a, a) on the two source rows.  All float32, fused where
// written as fmaf and nowhere else (the library is compiled -fmad=false, the emulation -ffp-contract=off).
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define SDET_HD __host__ __device__ __forceinline__
#else
#define SDET_HD inline
#endif

namespace sdet_paste {

constexpr int kMaxSide = 64;  // M + 2 <= 64

struct Geom {
  int x0i, y0i;            // top-left of the expanded integer box (may be negative)
  int w, h;                // resize target, >= 1
  int x_0, x_1, y_0, y_1;  // the pasted part: columns [x_0, x_1), rows [y_0, y_1) of the image; empty if x_1 <= x_0
};

SDET_HD Geom paste_geom(const float* box, int M, int im_h, int im_w) {
  const float scale = (float)(((double)M + 2.0) / (double)M);  // python float, cast to the array's float32
  float w_half = (box[2] - box[0]) * 0.5f;
  float h_half = (box[3] - box[1]) * 0.5f;
  const float x_c = (box[2] + box[0]) * 0.5f;
  const float y_c = (box[3] + box[1]) * 0.5f;
  w_half = w_half * scale;
  h_half = h_half * scale;
  const int x0 = (int)(x_c - w_half), x1 = (int)(x_c + w_half);  // astype(np.int32): truncation
  const int y0 = (int)(y_c - h_half), y1 = (int)(y_c + h_half);
  Geom g;
  g.x0i = x0;
  g.y0i = y0;
  g.w = x1 - x0 + 1 > 1 ? x1 - x0 + 1 : 1;
  g.h = y1 - y0 + 1 > 1 ? y1 - y0 + 1 : 1;
  g.x_0 = x0 > 0 ? x0 : 0;
  g.x_1 = x1 + 1 < im_w ? x1 + 1 : im_w;
  g.y_0 = y0 > 0 ? y0 : 0;
  g.y_1 = y1 + 1 < im_h ? y1 + 1 : im_h;
  // a box whose integer extent is inverted (w forced to 1) pastes nothing in numpy (empty slices on both sides);
  // a box entirely outside the image makes the reference's slice assignment raise: both are "nothing pasted" here
  if (x1 < x0 || y1 < y0 || g.x_1 <= g.x_0 || g.y_1 <= g.y_0) g.x_1 = g.x_0, g.y_1 = g.y_0;
  return g;
}

// one axis of the resize: destination index d of dn -> source samples s0, s1 and the weight of s1
SDET_HD void axis_coord(int d, int dn, int sn, int* s0, int* s1, float* t) {
  const double f = ((double)d + 0.5) * ((double)sn / (double)dn) - 0.5;
  const double fl = floor(f);
  int s = (int)fl;
  float frac = (float)(f - fl);
  if (s < 0) s = 0, frac = 0.f;
  if (s >= sn - 1) s = sn - 1, frac = 0.f;
  *s0 = s;
  *s1 = s + 1 < sn ? s + 1 : sn - 1;
  *t = frac;
}

// the zero-ringed mask: (r, c) in [0, M+2)
SDET_HD float padded(const float* mask, int M, int r, int c) {
  return (r >= 1 && r <= M && c >= 1 && c <= M) ? mask[(r - 1) * M + (c - 1)] : 0.f;
}

// the horizontally interpolated source column for destination column mx of the resized mask: hcol[r], r in [0, M+2)
SDET_HD void column_profile(const float* mask, int M, const Geom& g, int mx, float* hcol) {
  int c0, c1;
  float tx;
  axis_coord(mx, g.w, M + 2, &c0, &c1, &tx);
  for (int r = 0; r < M + 2; ++r) {
    const float a = padded(mask, M, r, c0), b = padded(mask, M, r, c1);
    hcol[r] = fmaf(tx, b - a, a);
  }
}

SDET_HD int pixel_bit(const float* hcol, int M, const Geom& g, int my) {
  int r0, r1;
  float ty;
  axis_coord(my, g.h, M + 2, &r0, &r1, &ty);
  const float top = hcol[r0], bot = hcol[r1];
  return fmaf(ty, bot - top, top) > 0.5f ? 1 : 0;
}

// Transitions of image column x (x_0 <= x < x_1) of one detection, in flat column-major positions.  Returns their
// number; writes them to `out` when it is not null.  The thread of column x owns the positions x*im_h + y for y in
// [y_0, y_1) plus the position right after its last row, unless that position is the first row of the next box
// column (then the next column's thread owns it, with this column's last bit as its predecessor).
SDET_HD int column_transitions(const float* mask, int M, const Geom& g, int im_h, int im_w, int x, int* out) {
  float hcol[kMaxSide];
  column_profile(mask, M, g, x - g.x0i, hcol);
  const bool full_height = g.y_0 == 0 && g.y_1 == im_h;
  int prev = 0;  // the pixel before (x, y_0) in column-major order: above the box (zero), or the previous column's last row
  if (full_height && x > g.x_0) {
    float hprev[kMaxSide];
    column_profile(mask, M, g, x - 1 - g.x0i, hprev);
    prev = pixel_bit(hprev, M, g, im_h - 1 - g.y0i);
  }
  int n = 0;
  const int base = x * im_h;
  for (int y = g.y_0; y < g.y_1; ++y) {
    const int bit = pixel_bit(hcol, M, g, y - g.y0i);
    if (bit != prev) {
      if (out) out[n] = base + y;
      ++n;
      prev = bit;
    }
  }
  const bool next_is_box = full_height && x + 1 < g.x_1;
  const long long close = (long long)base + g.y_1;
  if (prev && !next_is_box && close < (long long)im_h * im_w) {
    if (out) out[n] = (int)close;
    ++n;
  }
  return n;
}

// The whole thread of both kernels: detection `det`, column slot `j` (image column x_0 + j).  Pass 1 (write false)
// stores the column's flip count in col_counts[det][j]; pass 2 writes the positions at col_offsets[det][j].
SDET_HD void paste_thread(bool write, int det, int j, const float* boxes, const int* cls, const float* masks, int K,
                          int M, int im_h, int im_w, int* col_counts, const long long* col_offsets, int* positions) {
  if (j >= im_w) return;
  const Geom g = paste_geom(boxes + 4 * (size_t)det, M, im_h, im_w);
  const int x = g.x_0 + j;
  const size_t slot = (size_t)det * im_w + j;
  const int c = cls[det];
  if (x >= g.x_1 || c < 0 || c >= K) {  // past the pasted width, or not a class of this mask tensor: nothing here
    if (!write) col_counts[slot] = 0;
    return;
  }
  const float* mask = masks + ((size_t)det * K + (size_t)c) * M * M;
  if (write)
    column_transitions(mask, M, g, im_h, im_w, x, positions + col_offsets[slot]);
  else
    col_counts[slot] = column_transitions(mask, M, g, im_h, im_w, x, nullptr);
}

}  // namespace sdet_paste
