/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * Polygon -> roi-normalised binary mask, as ProposalMaskTarget does it
 * (operator_cxx/proposal_mask_target.cc:155-213 convertPoly2Mask).  The rasteriser itself
 * (rleFrPoly / rleDecode) comes from RogerChern/cocoapi common/maskApi.c, which is NOT vendored in
 * the reference and is cloned unpinned (doc/INSTALL.md:90-93): restated here from the published
 * pycocotools algorithm — PARITY UNPINNED.
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
