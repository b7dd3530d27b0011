// ORACLE - TEST INFRASTRUCTURE ONLY.
//
// Stand-in for `coco_api/common/maskApi.h`, which proposal_mask_target.cc:10 includes from a cocoapi checkout that
// is NOT part of the reference tree (RogerChern/cocoapi, cloned unpinned by doc/INSTALL.md:90-93).  It provides the
// four entry points that file calls.  The rasteriser (rleFrPoly) is restated from the published pycocotools
// algorithm, in its run-length form - so what compiling proposal_mask_target.cc against this header pins is the
// OPERATOR (roi / ground-truth matching, sampling, the polygon transform into roi coordinates, the union over
// segments, the mask ratio), not the rasteriser: that stays "parity unpinned" (DESIGN.md section 2).
#ifndef ORACLE_SHIM_MASKAPI_H_
#define ORACLE_SHIM_MASKAPI_H_
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

typedef unsigned int uint;
typedef unsigned long siz;
typedef unsigned char byte;
typedef struct { siz h, w, m; uint* cnts; } RLE;

inline void rlesInit(RLE** R, siz n) {
  *R = static_cast<RLE*>(std::calloc(n ? n : 1, sizeof(RLE)));
}
inline void rlesFree(RLE** R, siz n) {
  for (siz i = 0; i < n; ++i) std::free((*R)[i].cnts);
  std::free(*R);
  *R = nullptr;
}

// polygon (k vertices, xy interleaved, already in mask coordinates) -> column-major run lengths
inline void rleFrPoly(RLE* R, const double* xy, siz k, siz h, siz w) {
  const double scale = 5;
  std::vector<int> x(k + 1), y(k + 1);
  for (siz j = 0; j < k; ++j) x[j] = static_cast<int>(scale * xy[j * 2 + 0] + .5);
  for (siz j = 0; j < k; ++j) y[j] = static_cast<int>(scale * xy[j * 2 + 1] + .5);
  x[k] = x[0];
  y[k] = y[0];
  // upsampled boundary: every edge walked along its major axis
  std::vector<int> u, v;
  for (siz j = 0; j < k; ++j) {
    int xs = x[j], xe = x[j + 1], ys = y[j], ye = y[j + 1];
    const int dx = std::abs(xe - xs), dy = std::abs(ys - ye);
    const bool flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
    if (flip) { std::swap(xs, xe); std::swap(ys, ye); }
    const double s = dx >= dy ? static_cast<double>(ye - ys) / dx : static_cast<double>(xe - xs) / dy;
    if (dx >= dy) {
      for (int d = 0; d <= dx; ++d) {
        const int t = flip ? dx - d : d;
        u.push_back(t + xs);
        v.push_back(static_cast<int>(ys + s * t + .5));
      }
    } else {
      for (int d = 0; d <= dy; ++d) {
        const int t = flip ? dy - d : d;
        v.push_back(t + ys);
        u.push_back(static_cast<int>(xs + s * t + .5));
      }
    }
  }
  // column crossings, downsampled
  std::vector<uint> a;
  for (size_t j = 1; j < u.size(); ++j) {
    if (u[j] == u[j - 1]) continue;
    double xd = static_cast<double>(u[j] < u[j - 1] ? u[j] : u[j] - 1);
    xd = (xd + .5) / scale - .5;
    if (std::floor(xd) != xd || xd < 0 || xd > w - 1) continue;
    double yd = static_cast<double>(v[j] < v[j - 1] ? v[j] : v[j - 1]);
    yd = (yd + .5) / scale - .5;
    if (yd < 0) yd = 0; else if (yd > h) yd = h;
    yd = std::ceil(yd);
    a.push_back(static_cast<uint>(static_cast<int>(xd) * static_cast<int>(h) + static_cast<int>(yd)));
  }
  a.push_back(static_cast<uint>(h * w));
  std::sort(a.begin(), a.end());
  uint p = 0;
  for (auto& e : a) { const uint t = e; e -= p; p = t; }
  // zero-length runs merge their neighbours
  std::vector<uint> b;
  size_t j = 0;
  b.push_back(a[j++]);
  while (j < a.size()) {
    if (a[j] > 0) b.push_back(a[j++]);
    else { ++j; if (j < a.size()) b.back() += a[j++]; }
  }
  R->h = h; R->w = w; R->m = b.size();
  R->cnts = static_cast<uint*>(std::malloc(sizeof(uint) * b.size()));
  std::copy(b.begin(), b.end(), R->cnts);
}

inline void rleDecode(const RLE* R, byte* M, siz n) {
  for (siz i = 0; i < n; ++i) {
    byte val = 0;
    for (siz j = 0; j < R[i].m; ++j) {
      for (uint c = 0; c < R[i].cnts[j]; ++c) *(M++) = val;
      val = !val;
    }
  }
}
#endif
