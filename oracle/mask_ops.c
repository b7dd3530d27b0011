/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * Polygon -> roi-normalised binary mask, as ProposalMaskTarget does it
 * (operator_cxx/proposal_mask_target.cc:155-213 convertPoly2Mask).  The rasteriser itself
 * (rleFrPoly / rleDecode) comes from RogerChern/cocoapi common/maskApi.c, which is NOT vendored in
 * the reference and is cloned unpinned (doc/INSTALL.md:90-93): restated here from the published
 * pycocotools algorithm — PARITY UNPINNED.  What IS pinned: everything the reference's
 * operator does around the rasteriser (oracle_poly2mask / oracle_poly2mask_ratio against proposal_mask_target.cc
 * compiled with a stand-in maskApi.h, tests/test_oracle_ref_cxx.py::test_proposal_mask_target_*).
 *
 *   oracle_rle_fr_poly_mask   maskApi.c rleFrPoly + rleDecode: polygon (k vertices, x/y doubles)
 *                             -> h*w bytes, COLUMN-major like the RLE
 *   oracle_poly2mask          proposal_mask_target.cc:155-213 for one (roi, encoded polygon row)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int uint_cmp(const void* a, const void* b) {
  unsigned c = *(const unsigned*)a, d = *(const unsigned*)b;
  return c > d ? 1 : c < d ? -1 : 0;
}

/* mask (h*w bytes, column-major: index x*h + y) of one polygon */
void oracle_rle_fr_poly_mask(const double* xy, long k, long h, long w, unsigned char* mask) {
  const double scale = 5;
  long j, m = 0;
  int* x = (int*)malloc(sizeof(int) * (k + 1));
  int* y = (int*)malloc(sizeof(int) * (k + 1));
  for (j = 0; j < k; j++) x[j] = (int)(scale * xy[j * 2 + 0] + .5);
  x[k] = x[0];
  for (j = 0; j < k; j++) y[j] = (int)(scale * xy[j * 2 + 1] + .5);
  y[k] = y[0];
  for (j = 0; j < k; j++) {
    int ax = abs(x[j] - x[j + 1]), ay = abs(y[j] - y[j + 1]);
    m += (ax > ay ? ax : ay) + 1;
  }
  int* u = (int*)malloc(sizeof(int) * (m + 1));
  int* v = (int*)malloc(sizeof(int) * (m + 1));
  m = 0;
  for (j = 0; j < k; j++) {
    int xs = x[j], xe = x[j + 1], ys = y[j], ye = y[j + 1], dx, dy, t, d, flip;
    double s;
    dx = abs(xe - xs);
    dy = abs(ys - ye);
    flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
    if (flip) { t = xs; xs = xe; xe = t; t = ys; ys = ye; ye = t; }
    s = dx >= dy ? (double)(ye - ys) / dx : (double)(xe - xs) / dy;
    if (dx >= dy) for (d = 0; d <= dx; d++) {
      t = flip ? dx - d : d; u[m] = t + xs; v[m] = (int)(ys + s * t + .5); m++;
    } else for (d = 0; d <= dy; d++) {
      t = flip ? dy - d : d; v[m] = t + ys; u[m] = (int)(xs + s * t + .5); m++;
    }
  }
  /* get points along y-boundary and downsample */
  long kk = m;
  unsigned* a = (unsigned*)malloc(sizeof(unsigned) * (kk + 2));
  m = 0;
  for (j = 1; j < kk; j++) if (u[j] != u[j - 1]) {
    double xd = (double)(u[j] < u[j - 1] ? u[j] : u[j] - 1), yd;
    xd = (xd + .5) / scale - .5;
    if (floor(xd) != xd || xd < 0 || xd > w - 1) continue;
    yd = (double)(v[j] < v[j - 1] ? v[j] : v[j - 1]);
    yd = (yd + .5) / scale - .5;
    if (yd < 0) yd = 0; else if (yd > h) yd = h;
    yd = ceil(yd);
    a[m++] = (unsigned)((int)xd * (int)h + (int)yd);
  }
  /* RLE from the sorted boundary positions, decoded straight away (rleDecode: runs alternate
   * 0,1,0,... in column-major order; zero-length runs merge their neighbours) */
  a[m++] = (unsigned)(h * w);
  qsort(a, (size_t)m, sizeof(unsigned), uint_cmp);
  memset(mask, 0, (size_t)(h * w));
  {
    unsigned p = 0; long q; unsigned char val = 0;
    for (q = 0; q < m; q++) {
      unsigned e = a[q] > (unsigned)(h * w) ? (unsigned)(h * w) : a[q];
      if (val) memset(mask + p, 1, (size_t)(e > p ? e - p : 0));
      if (e > p) p = e;
      val = !val;
    }
  }
  free(x); free(y); free(u); free(v); free(a);
}

/* proposal_mask_target.cc:155-213: poly row = [category, n_seg, len_1..len_n, x y x y ...];
 * mask (M*M floats) in the order the reference flattens it. */
void oracle_poly2mask(const float* roi, const float* poly, int mask_size, float* mask) {
  float w = roi[2] - roi[0], h = roi[3] - roi[1];
  w = 1.f > w ? 1.f : w; /* max((DType)1., w) */
  h = 1.f > h ? 1.f : h;
  const int n_seg = (int)poly[1];
  int offset = 2 + n_seg;
  const int MM = mask_size * mask_size;
  unsigned char* seg = (unsigned char*)malloc((size_t)MM);
  for (int j = 0; j < MM; ++j) mask[j] = 0.f;
  for (int i = 0; i < n_seg; ++i) {
    const int cur_len = (int)poly[i + 2];
    double* xys = (double*)malloc(sizeof(double) * (size_t)(cur_len > 0 ? cur_len : 1));
    for (int j = 0; j < cur_len; ++j) {
      if (j % 2 == 0) xys[j] = (poly[offset + j + 1] - roi[1]) * mask_size / h; /* y' first (:184) */
      else xys[j] = (poly[offset + j - 1] - roi[0]) * mask_size / w;            /* then x' (:186) */
    }
    oracle_rle_fr_poly_mask(xys, cur_len / 2, mask_size, mask_size, seg);
    for (int j = 0; j < MM; ++j) if (seg[j] == 1) mask[j] = 1.f; /* OR over segments (:197-208) */
    free(xys);
    offset += cur_len;
  }
  free(seg);
}

/* proposal_mask_target.cc:20-152 convertPoly2MaskWithRatio (output_ratio=True, Mask Scoring R-CNN): the roi mask
 * with the vertex transform carried out in DOUBLE (`poly_index` is a double there; the plain variant above works in
 * float), plus  ratio = |polygon ∩ roi crop| / (|polygon| + 1e-4)  counted on two integer rasters: the crop
 * (crop_h x crop_w, roi corners truncated to int) and the polygon's own extent joined with the roi (full_h x full_w).
 * Returns the ratio as the double the reference returns; the operator stores it into a float. */
double oracle_poly2mask_ratio(const float* roi, const float* poly, int mask_size, float* mask) {
  float w = roi[2] - roi[0], h = roi[3] - roi[1];
  w = 1.f > w ? 1.f : w;
  h = 1.f > h ? 1.f : h;
  const int n_seg = (int)poly[1];
  const int MM = mask_size * mask_size;
  const int rx1 = (int)roi[0], rx2 = (int)roi[2], ry1 = (int)roi[1], ry2 = (int)roi[3];
  const long crop_w = rx2 - rx1 + 1, crop_h = ry2 - ry1 + 1;
  double ox1 = roi[0], ox2 = roi[2], oy1 = roi[1], oy2 = roi[3];
  unsigned char* seg = (unsigned char*)malloc((size_t)MM);
  unsigned char* crop = (unsigned char*)calloc((size_t)(crop_w * crop_h > 0 ? crop_w * crop_h : 1), 1);
  unsigned char* tmp = (unsigned char*)malloc((size_t)(crop_w * crop_h > 0 ? crop_w * crop_h : 1));
  for (int j = 0; j < MM; ++j) mask[j] = 0.f;
  int offset = 2 + n_seg;
  for (int i = 0; i < n_seg; ++i) {
    const int cur_len = (int)poly[i + 2];
    double* xys = (double*)malloc(sizeof(double) * (size_t)(cur_len > 0 ? cur_len : 1));
    double* xyc = (double*)malloc(sizeof(double) * (size_t)(cur_len > 0 ? cur_len : 1));
    for (int j = 0; j < cur_len; ++j) {
      if (j % 2 == 0) {                       /* a y coordinate (:51-58) */
        const double py = poly[offset + j + 1];
        oy1 = oy1 < py ? oy1 : py;
        oy2 = oy2 < py ? py : oy2;
        xys[j] = (py - roi[1]) * mask_size / h;
        xyc[j + 1] = py - roi[1];
      } else {                                /* an x coordinate (:59-65) */
        const double px = poly[offset + j - 1];
        ox1 = ox1 < px ? ox1 : px;
        ox2 = ox2 < px ? px : ox2;
        xys[j] = (px - roi[0]) * mask_size / w;
        xyc[j - 1] = px - roi[0];
      }
    }
    oracle_rle_fr_poly_mask(xys, cur_len / 2, mask_size, mask_size, seg);
    for (int j = 0; j < MM; ++j) if (seg[j] == 1) mask[j] = 1.f;
    oracle_rle_fr_poly_mask(xyc, cur_len / 2, crop_h, crop_w, tmp);
    for (long j = 0; j < crop_w * crop_h; ++j) crop[j] |= tmp[j];
    free(xys);
    free(xyc);
    offset += cur_len;
  }
  const int ix1 = (int)ox1, ix2 = (int)ox2, iy1 = (int)oy1, iy2 = (int)oy2;
  const long full_w = ix2 - ix1 + 1, full_h = iy2 - iy1 + 1;
  unsigned char* full = (unsigned char*)calloc((size_t)(full_w * full_h > 0 ? full_w * full_h : 1), 1);
  unsigned char* ftmp = (unsigned char*)malloc((size_t)(full_w * full_h > 0 ? full_w * full_h : 1));
  offset = 2 + n_seg;
  for (int i = 0; i < n_seg; ++i) {
    const int cur_len = (int)poly[i + 2];
    double* xyo = (double*)malloc(sizeof(double) * (size_t)(cur_len > 0 ? cur_len : 1));
    for (int j = 0; j < cur_len; ++j) {
      if (j % 2 == 0) xyo[j + 1] = (double)poly[offset + j + 1] - oy1;   /* :89-92 */
      else xyo[j - 1] = (double)poly[offset + j - 1] - ox1;             /* :93-96 */
    }
    oracle_rle_fr_poly_mask(xyo, cur_len / 2, full_h, full_w, ftmp);
    for (long j = 0; j < full_w * full_h; ++j) full[j] |= ftmp[j];
    free(xyo);
    offset += cur_len;
  }
  double origin_sum = 0, crop_sum = 0;
  for (long j = 0; j < full_w * full_h; ++j) origin_sum += full[j];
  for (long j = 0; j < crop_w * crop_h; ++j) crop_sum += crop[j];
  double ratio = crop_sum / (origin_sum + 0.0001);
  ratio = ratio < 1e-10 ? 1e-10 : ratio;
  free(seg); free(crop); free(tmp); free(full); free(ftmp);
  return ratio;
}
