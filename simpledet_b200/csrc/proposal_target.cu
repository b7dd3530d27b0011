// ProposalTarget on the device (the reference's "gpu" operator copies everything to the host and
// loops there: operator_cxx/proposal_target-inl.h:146-149,251-255; there is no .cu).
//
// Reference semantics: operator_cxx/proposal_target-inl.h:123-256 (filter padding, append gt),
// operator_cxx/proposal_target.cc:22-163 SampleROI, :165-185 BBoxOverlap, :187-202
// ExpandBboxRegressionTargets, :204-227 NonLinearTransformAndNormalization.
//
// One CTA per image does the whole assignment in shared memory: ordered compaction of valid
// gt / rois, IoU + first-max argmax per roi, fg / bg / neg partition (index order), priority
// sort for each "random_shuffle", target encoding and the class-slot scatter.  A shuffle is
// "order the candidates by a 32-bit priority, ties by index"; priorities come from cuRAND Philox
// (seed, image, draw, candidate) or are injected by the caller (tests inject the same array into
// the oracle, making every output comparable bit for bit — SURVEY.md §0.8).
#include <curand_kernel.h>

#include <cfloat>

#include "common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kMaxT = 4096;  // rois + gt boxes per image

struct PTParams {
  const float* rois;      // (B,R,4)
  const float* gt;        // (B,G,5)
  const uint32_t* prio;   // (B,D,T) injected priorities or nullptr -> Philox
  uint32_t* prio_used;    // (B,D,T) or nullptr: the priorities this call used
  float* rois_out;        // (B,IR,4)
  float* labels;          // (B,IR)
  float* tgt;             // (B,IR,NC4)
  float* wgt;             // (B,IR,NC4)
  float* iou;             // (B,IR)
  int* kept;              // (B,IR) or nullptr: index into the compacted roi list, -1 = empty row
  int* gt_index;          // (B,IR) or nullptr: source row (0..G-1) of the matched gt box, -1 = none
  int* fg_count;          // (B) or nullptr: number of foreground rows (labels may be > 0 only there)
  int B, R, G, NC4, IR, D, fg_per_image;
  float fg_thresh, bg_hi, bg_lo;
  int without_gt, agnostic;
  const float* valid_ranges;  // (B,2) or nullptr (ProposalTarget_v2)
  int filter_scales, no_fg_cap;
  float mean[4], std[4], weight[4];
  unsigned long long seed;
};

__device__ __forceinline__ float fmin_ref(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float fmax_ref(float a, float b) { return a < b ? b : a; }

// Ordered compaction: list <- { i in [0,n) : pred(i) } in increasing i.  Returns the count.
// All threads call; s_warp is 32 ints of scratch.
template <typename Pred>
__device__ int block_compact(int n, Pred pred, int* list, int* s_warp, int* s_total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *s_total = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    const bool f = (i < n) && pred(i);
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      int v = s_warp[lane], inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      s_warp[lane] = inc - v;  // exclusive prefix of the warp counts
      if (lane == 31) s_warp[32] = inc;
    }
    __syncthreads();
    const int off = *s_total + s_warp[warp] + __popc(m & ((1u << lane) - 1u));
    if (f) list[off] = i;
    __syncthreads();
    if (tid == 0) *s_total += s_warp[32];
    __syncthreads();
  }
  const int total = *s_total;
  __syncthreads();  // nobody may re-enter (and reset *s_total) before every thread has read it
  return total;
}

// Sort list[0..n) by (prio[list[i]], list[i]) ascending.  keys: next_pow2(n) u64 of scratch.
__device__ void block_priority_sort(int* list, int n, const uint32_t* prio, unsigned long long* keys) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x)
    keys[i] = i < n ? (((unsigned long long)prio[list[i]] << 32) | (unsigned)list[i]) : ~0ull;
  __syncthreads();
  for (int size = 2; size <= np2; size <<= 1)
    for (int sh = 31 - __clz(size >> 1); sh >= 0; --sh) {  // stride = 1 << sh: shifts, not divisions, per element
      const int stride = 1 << sh;
      for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        const int lo = ((t >> sh) << (sh + 1)) | (t & (stride - 1)), hi = lo + stride;
        const bool asc = ((lo & size) == 0);
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == asc) {
          keys[lo] = c;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += blockDim.x) list[i] = (int)(keys[i] & 0xFFFFFFFFull);
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
proposal_target_kernel(const __grid_constant__ PTParams p) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int R = p.R, G = p.G, T = R + G, IR = p.IR, NC4 = p.NC4;
  int np2 = 1;
  while (np2 < T) np2 <<= 1;
  // shared layout
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(s_raw);  // np2
  float4* s_box = reinterpret_cast<float4*>(s_keys + np2);                    // T compacted rois
  float* s_gt = reinterpret_cast<float*>(s_box + T);                          // G*5 compacted gt
  float* s_maxov = s_gt + G * 5;                                              // T
  int* s_assign = reinterpret_cast<int*>(s_maxov + T);                        // T
  int* s_src = s_assign + T;                                                  // T  (scratch list)
  int* s_fg = s_src + T;                                                      // T
  int* s_bg = s_fg + T;                                                       // T
  int* s_neg = s_bg + T;                                                      // T
  int* s_kept = s_neg + T;                                                    // IR + T
  uint32_t* s_prio = reinterpret_cast<uint32_t*>(s_kept + IR + T);            // T (current draw)
  int* s_gsrc = reinterpret_cast<int*>(s_prio + T);                           // G source rows of valid gt
  __shared__ int s_warp[33];
  __shared__ int s_total, s_last[2];

  const float* rois = p.rois + (size_t)b * R * 4;
  const float* gt = p.gt + (size_t)b * G * 5;

  // ---- valid gt (cls != -1, -inl.h:158) and valid rois (y2 > 0, :174), gt appended (:177-185)
  const int ng = block_compact(G, [&](int j) { return gt[j * 5 + 4] != -1.f; }, s_src, s_warp, &s_total);
  for (int e = tid; e < ng * 5; e += blockDim.x) s_gt[e] = gt[s_src[e / 5] * 5 + e % 5];
  for (int j = tid; j < ng; j += blockDim.x) s_gsrc[j] = s_src[j];
  __syncthreads();
  const int nr = block_compact(R, [&](int j) { return rois[j * 4 + 3] > 0.f; }, s_src, s_warp, &s_total);
  for (int i = tid; i < nr; i += blockDim.x) {
    const float* r = rois + (size_t)s_src[i] * 4;
    s_box[i] = make_float4(r[0], r[1], r[2], r[3]);
  }
  int n = nr;
  if (!p.without_gt) {
    int napp = ng;
    if (p.filter_scales && p.valid_ranges) {  // ProposalTarget_v2: only gt inside the valid scale range
      const float v0 = p.valid_ranges[b * 2], v1 = p.valid_ranges[b * 2 + 1];
      const float vmin = __fmul_rn(v0, v0), vmax = __fmul_rn(v1, v1);
      napp = block_compact(ng, [&](int j) {
        const float gw = (float)__dadd_rn((double)__fsub_rn(s_gt[j * 5 + 2], s_gt[j * 5 + 0]), 1.0);
        const float gh = (float)__dadd_rn((double)__fsub_rn(s_gt[j * 5 + 3], s_gt[j * 5 + 1]), 1.0);
        const float ar = __fmul_rn(gw, gh);
        return !(ar < vmin || ar > vmax);
      }, s_src, s_warp, &s_total);
    } else {
      for (int j = tid; j < ng; j += blockDim.x) s_src[j] = j;
      __syncthreads();
    }
    for (int q = tid; q < napp; q += blockDim.x) {
      const int j = s_src[q];
      s_box[nr + q] = make_float4(s_gt[j * 5], s_gt[j * 5 + 1], s_gt[j * 5 + 2], s_gt[j * 5 + 3]);
    }
    n = nr + napp;
  }
  __syncthreads();

  // ---- BBoxOverlap + first-max argmax (proposal_target.cc:165-185, :51-63)
  for (int i = tid; i < n; i += blockDim.x) {
    const float4 bx = s_box[i];
    float best = 0.f;
    int bi = 0;
    const float ba = __fmul_rn(__fadd_rn(__fsub_rn(bx.z, bx.x), 1.f), __fadd_rn(__fsub_rn(bx.w, bx.y), 1.f));
    for (int j = 0; j < ng; ++j) {
      const float* q = s_gt + j * 5;
      float ov = 0.f;
      const float iw = __fadd_rn(__fsub_rn(fmin_ref(bx.z, q[2]), fmax_ref(bx.x, q[0])), 1.f);
      if (iw > 0.f) {
        const float ih = __fadd_rn(__fsub_rn(fmin_ref(bx.w, q[3]), fmax_ref(bx.y, q[1])), 1.f);
        if (ih > 0.f) {
          const float qa = __fmul_rn(__fadd_rn(__fsub_rn(q[2], q[0]), 1.f), __fadd_rn(__fsub_rn(q[3], q[1]), 1.f));
          const float inter = __fmul_rn(iw, ih);
          ov = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ba, qa), inter));
        }
      }
      if (j == 0) {
        best = ov;
      } else if (best < ov) {
        best = ov;
        bi = j;
      }
    }
    s_maxov[i] = best;
    s_assign[i] = bi;
  }
  __syncthreads();

  // ---- priorities of one draw into s_prio (injected or Philox), optionally exported
  auto load_draw = [&](int d) {
    for (int i = tid; i < T; i += blockDim.x) {
      uint32_t v;
      if (p.prio) {
        v = p.prio[((size_t)b * p.D + d) * T + i];
      } else {
        curandStatePhilox4_32_10_t st;
        curand_init(p.seed, (unsigned long long)b * T + i, (unsigned long long)d, &st);
        v = curand(&st);
      }
      s_prio[i] = v;
      if (p.prio_used) p.prio_used[((size_t)b * p.D + d) * T + i] = v;
    }
    __syncthreads();
  };

  // ---- fg / neg / bg partitions in index order (:72-78, :95-99)
  const float fg_thr = p.fg_thresh, bg_hi = p.bg_hi, bg_lo = p.bg_lo;
  const int nfg = block_compact(n, [&](int i) { return s_maxov[i] >= fg_thr; }, s_fg, s_warp, &s_total);
  const int nneg = block_compact(n, [&](int i) { return !(s_maxov[i] >= fg_thr); }, s_neg, s_warp, &s_total);
  const int nbg = block_compact(n, [&](int i) { return s_maxov[i] >= bg_lo && s_maxov[i] < bg_hi; }, s_bg,
                                s_warp, &s_total);
  const int fg_n = p.no_fg_cap ? nfg : min(p.fg_per_image, nfg);
  // draws are always materialised in the same order: 0 fg, 1 bg, 2+ negative padding
  load_draw(0);
  if (nfg > fg_n) block_priority_sort(s_fg, nfg, s_prio, s_keys);  // :81-85
  load_draw(1);
  const int bg_n = min(IR - fg_n, nbg);
  if (nbg > bg_n) block_priority_sort(s_bg, nbg, s_prio, s_keys);  // :100-104
  for (int i = tid; i < fg_n; i += blockDim.x) s_kept[i] = s_fg[i];
  for (int i = tid; i < bg_n; i += blockDim.x) s_kept[fg_n + i] = s_bg[i];
  int nk = fg_n + bg_n;
  __syncthreads();
  for (int r = 0; nk < IR && nneg > 0; ++r) {  // pad with negatives (:116-122)
    const int gap = IR - nk;
    load_draw(2 + r % (p.D - 2));
    block_priority_sort(s_neg, nneg, s_prio, s_keys);
    const int take = min(gap, nneg);
    for (int i = tid; i < take; i += blockDim.x) s_kept[nk + i] = s_neg[i];
    nk += take;
    __syncthreads();
  }

  // ---- outputs.  Everything is zero-initialised by the reference (-inl.h:188-192): the two (IR, 4*num_classes)
  // planes were zero-filled by the launcher (a wide memset, not 650 KB of stores from this one CTA).
  float* o_tgt = p.tgt + (size_t)b * IR * NC4;
  float* o_wgt = p.wgt + (size_t)b * IR * NC4;
  for (int i = tid; i < IR; i += blockDim.x) {
    const size_t row = (size_t)b * IR + i;
    float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
    float label = 0.f, ov = 0.f;
    int k = -1;
    if (i < nk) {
      k = s_kept[i];
      rb = s_box[k];
      ov = s_maxov[k];
      if (ng > 0) {
        const float* g = s_gt + s_assign[k] * 5;
        if (i < fg_n) label = g[4];  // labels only for the first fg_n rows (:128-131)
        // NonLinearTransformAndNormalization (:204-227); `0.5 * (w - 1.f)` is double arithmetic
        const float ew = __fadd_rn(__fsub_rn(rb.z, rb.x), 1.f), eh = __fadd_rn(__fsub_rn(rb.w, rb.y), 1.f);
        const float ecx = (float)__dadd_rn((double)rb.x, __dmul_rn(0.5, (double)__fsub_rn(ew, 1.f)));
        const float ecy = (float)__dadd_rn((double)rb.y, __dmul_rn(0.5, (double)__fsub_rn(eh, 1.f)));
        const float gw = __fadd_rn(__fsub_rn(g[2], g[0]), 1.f), gh = __fadd_rn(__fsub_rn(g[3], g[1]), 1.f);
        const float gcx = (float)__dadd_rn((double)g[0], __dmul_rn(0.5, (double)__fsub_rn(gw, 1.f)));
        const float gcy = (float)__dadd_rn((double)g[1], __dmul_rn(0.5, (double)__fsub_rn(gh, 1.f)));
        float t[4];
        t[0] = __fdiv_rn(__fsub_rn(gcx, ecx), __fadd_rn(ew, 1e-14f));
        t[1] = __fdiv_rn(__fsub_rn(gcy, ecy), __fadd_rn(eh, 1e-14f));
        t[2] = logf(__fdiv_rn(gw, ew));
        t[3] = logf(__fdiv_rn(gh, eh));
#pragma unroll
        for (int c = 0; c < 4; ++c) t[c] = __fdiv_rn(__fsub_rn(t[c], p.mean[c]), p.std[c]);
        const float cls = p.agnostic ? (label < 1.f ? label : 1.f) : label;  // :151-157
        if (cls > 0.f) {  // ExpandBboxRegressionTargets (:187-202)
          const int start = 4 * (int)cls;
          if (start + 4 <= NC4) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              o_tgt[(size_t)i * NC4 + start + c] = t[c];
              o_wgt[(size_t)i * NC4 + start + c] = p.weight[c];
            }
          }
        }
      }
    }
    p.rois_out[row * 4 + 0] = rb.x;
    p.rois_out[row * 4 + 1] = rb.y;
    p.rois_out[row * 4 + 2] = rb.z;
    p.rois_out[row * 4 + 3] = rb.w;
    p.labels[row] = label;
    p.iou[row] = ov;
    if (p.kept) p.kept[row] = k;
    if (p.gt_index) p.gt_index[row] = (k >= 0 && ng > 0) ? s_gsrc[s_assign[k]] : -1;
  }
  if (p.fg_count && tid == 0) p.fg_count[b] = fg_n;
}

// --------------------------------------------------------------------------------------------
// ProposalMaskTarget's rasteriser: convertPoly2Mask (operator_cxx/proposal_mask_target.cc:155-213)
// on top of cocoapi's rleFrPoly / rleDecode (RogerChern/cocoapi common/maskApi.c — not vendored
// in the reference, restated from the published algorithm; parity unpinned).
//
// One CTA per (image, foreground row).  rleFrPoly is: upsample x5, walk every edge with integer
// DDA, keep the points where the (upsampled) column changes, snap them to pixel-column boundaries
// and SORT them; the mask is the run-length decode of the sorted positions.  The decode of sorted
// toggle positions is "pixel p is set iff an odd number of positions are <= p", so no sort is needed:
// every edge thread toggles a per-position counter in shared memory and a parity prefix scan over
// the M*M + 1 positions produces the mask.  All coordinate arithmetic is the reference's
// float -> double -> int sequence with explicit rounding (no contraction).
// --------------------------------------------------------------------------------------------
struct MaskParams {
  const float* rois_out;   // (B,IR,4)
  const float* gt_polys;   // (B,G,PL)
  const int* gt_index;     // (B,IR)
  const int* fg_count;     // (B)
  float* mask;             // (B,NM,M,M)
  int IR, G, PL, NM, M;
  int double_xy;           // 1: the roi-frame vertex transform in double (convertPoly2MaskWithRatio, .cc:51-65)
  float* ratio;            // (B,NM) mask ratio, poly_ratio_kernel only
};

__device__ __forceinline__ void dda_point(int xs, int ys, int dx, int dy, double sl, bool flip, int d, int& u,
                                          int& v) {
  // maskApi.c rleFrPoly: t = flip ? (len - d) : d; major axis advances by t, minor = (int)(start + s*t + .5)
  if (dx >= dy) {
    const int t = flip ? dx - d : d;
    u = t + xs;
    v = (int)__dadd_rn(__dadd_rn((double)ys, __dmul_rn(sl, (double)t)), .5);
  } else {
    const int t = flip ? dy - d : d;
    v = t + ys;
    u = (int)__dadd_rn(__dadd_rn((double)xs, __dmul_rn(sl, (double)t)), .5);
  }
}

struct EdgeRec {
  double sl;
  int xs, ys, dx, dy, npts, start;
  bool flip;
};
constexpr int kMaxEdges = 512;

__global__ void __launch_bounds__(256) poly_mask_kernel(const MaskParams p) {
  extern __shared__ int s_tog[];  // M*M + 2 toggle counters, then the OR-accumulated mask (M*M bytes as ints)
  __shared__ EdgeRec s_edge[kMaxEdges];
  __shared__ unsigned s_chunk[400];
  __shared__ int s_cpar[400];
  __shared__ int s_total, s_last[2];
  const int row = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int M = p.M, MM = M * M;
  float* out = p.mask + ((size_t)b * p.NM + row) * MM;
  const int nfg = min(p.fg_count[b], p.NM);
  if (row >= nfg) {  // rows beyond the foreground count keep the ignore value (-inl.h:242-243)
    for (int j = tid; j < MM; j += blockDim.x) out[j] = -1.f;
    return;
  }
  int* s_acc = s_tog + MM + 2;
  const float* roi = p.rois_out + ((size_t)b * p.IR + row) * 4;
  const int gi = p.gt_index[(size_t)b * p.IR + row];
  for (int j = tid; j < MM; j += blockDim.x) s_acc[j] = 0;
  __syncthreads();
  if (gi >= 0) {
    const float* poly = p.gt_polys + ((size_t)b * p.G + gi) * p.PL;
    float w = __fsub_rn(roi[2], roi[0]), h = __fsub_rn(roi[3], roi[1]);
    w = 1.f > w ? 1.f : w;
    h = 1.f > h ? 1.f : h;
    const int n_seg = (int)poly[1];
    int offset = 2 + n_seg;
    for (int sg = 0; sg < n_seg; ++sg) {
      const int cur_len = (int)poly[sg + 2];
      const int k = cur_len / 2;
      for (int j = tid; j < MM + 2; j += blockDim.x) s_tog[j] = 0;
      __syncthreads();
      // vertex j in upsampled integer coordinates; note the (y', x') order of :182-187
      auto vert = [&](int j, int& X, int& Y) {
        j = (j == k) ? 0 : j;
        double a, c;
        if (p.double_xy) {  // `poly_index` is a double in the ratio variant: ((py - roi[1]) * mask_size) / h in double
          a = __ddiv_rn(__dmul_rn(__dsub_rn((double)poly[offset + 2 * j + 1], (double)roi[1]), (double)M), (double)h);
          c = __ddiv_rn(__dmul_rn(__dsub_rn((double)poly[offset + 2 * j], (double)roi[0]), (double)M), (double)w);
        } else {
          a = (double)__fdiv_rn(__fmul_rn(__fsub_rn(poly[offset + 2 * j + 1], roi[1]), (float)M), h);
          c = (double)__fdiv_rn(__fmul_rn(__fsub_rn(poly[offset + 2 * j], roi[0]), (float)M), w);
        }
        X = (int)__dadd_rn(__dmul_rn(5.0, a), .5);
        Y = (int)__dadd_rn(__dmul_rn(5.0, c), .5);
      };
      // edges in batches of kMaxEdges (COCO polygons can have more vertices than fit the shared table)
      for (int e0 = 0; e0 < k; e0 += kMaxEdges) {
        const int nb = min(kMaxEdges, k - e0);
        // ---- one thread per edge: the edge's DDA parameters and point count into shared memory
        for (int e = tid; e < nb; e += blockDim.x) {
          int xs, ys, xe, ye;
          vert(e0 + e, xs, ys);
          vert(e0 + e + 1, xe, ye);
          const int dx = abs(xe - xs), dy = abs(ys - ye);
          const bool flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
          if (flip) {
            int t = xs; xs = xe; xe = t;
            t = ys; ys = ye; ye = t;
          }
          EdgeRec er;
          er.xs = xs; er.ys = ys; er.dx = dx; er.dy = dy; er.flip = flip;
          er.sl = dx >= dy ? __ddiv_rn((double)(ye - ys), (double)dx) : __ddiv_rn((double)(xe - xs), (double)dy);
          er.npts = max(dx, dy) + 1;
          er.start = 0;
          s_edge[e] = er;
        }
        __syncthreads();
        if (tid == 0) {  // start of every edge in the concatenated point list
          int acc = 0;
          for (int e = 0; e < nb; ++e) {
            s_edge[e].start = acc;
            acc += s_edge[e].npts;
          }
          s_total = acc;
        }
        __syncthreads();
        // ---- one thread per POINT of the concatenated list (rleFrPoly walks them in order and compares each with
        // its predecessor; both are functions of (edge, d) alone, so every point is independent)
        const int total = s_total;
        for (int i = tid; i < total; i += blockDim.x) {
          int lo = 0, hi = nb - 1;  // last edge whose start <= i
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_edge[mid].start <= i) lo = mid;
            else hi = mid - 1;
          }
          const EdgeRec er = s_edge[lo];
          const int d = i - er.start;
          int u, v, pu, pv;
          dda_point(er.xs, er.ys, er.dx, er.dy, er.sl, er.flip, d, u, v);
          if (d > 0) {
            dda_point(er.xs, er.ys, er.dx, er.dy, er.sl, er.flip, d - 1, pu, pv);
          } else if (lo > 0) {  // the last point of the previous edge
            const EdgeRec pr = s_edge[lo - 1];
            dda_point(pr.xs, pr.ys, pr.dx, pr.dy, pr.sl, pr.flip, pr.npts - 1, pu, pv);
          } else if (e0 > 0) {  // ... which belongs to the previous batch
            pu = s_last[0];
            pv = s_last[1];
          } else {
            continue;  // the very first point has no predecessor
          }
          if (u != pu) {
            double xd = (double)(u < pu ? u : u - 1);
            xd = __dsub_rn(__ddiv_rn(__dadd_rn(xd, .5), 5.0), .5);
            if (!(floor(xd) != xd || xd < 0 || xd > (double)(M - 1))) {
              double yd = (double)(v < pv ? v : pv);
              yd = __dsub_rn(__ddiv_rn(__dadd_rn(yd, .5), 5.0), .5);
              if (yd < 0) yd = 0;
              else if (yd > (double)M) yd = (double)M;
              yd = ceil(yd);
              const int pos = (int)xd * M + (int)yd;  // <= M*M
              atomicAdd(&s_tog[min(pos, MM)], 1);
            }
          }
        }
        __syncthreads();
        if (tid == 0) {  // hand the batch's last point to the next batch
          const EdgeRec pr = s_edge[nb - 1];
          int lu, lv;
          dda_point(pr.xs, pr.ys, pr.dx, pr.dy, pr.sl, pr.flip, pr.npts - 1, lu, lv);
          s_last[0] = lu;
          s_last[1] = lv;
        }
        __syncthreads();
      }
      // ---- parity prefix over the positions: ballot inside 32-position chunks, then the chunks' parities in order
      const int nchunk = (MM + 31) / 32;
      for (int cidx = tid >> 5; cidx < nchunk; cidx += blockDim.x >> 5) {
        const int j = cidx * 32 + (tid & 31);
        const unsigned bits = __ballot_sync(0xffffffffu, j < MM && (s_tog[j] & 1));
        if ((tid & 31) == 0) s_chunk[cidx] = bits;
      }
      __syncthreads();
      if (tid < 32) {  // exclusive parity of the chunks before each chunk (nchunk <= 392 for M <= 112)
        int carry = 0;
        for (int c0_ = 0; c0_ < nchunk; c0_ += 32) {
          const int cidx = c0_ + tid;
          const int par = cidx < nchunk ? (__popc(s_chunk[cidx]) & 1) : 0;
          int inc = par;  // inclusive xor-scan over the warp
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc ^= t;
          }
          if (cidx < nchunk) s_cpar[cidx] = carry ^ inc ^ par;
          carry ^= __shfl_sync(0xffffffffu, inc, 31);
        }
      }
      __syncthreads();
      for (int j = tid; j < MM; j += blockDim.x) {
        const unsigned bits = s_chunk[j >> 5];
        const int par = s_cpar[j >> 5] ^ (__popc(bits & (0xffffffffu >> (31 - (j & 31)))) & 1);
        if (par) s_acc[j] = 1;
      }
      __syncthreads();
      offset += cur_len;
    }
  }
  for (int j = tid; j < MM; j += blockDim.x) out[j] = s_acc[j] ? 1.f : 0.f;
}


// --------------------------------------------------------------------------------------------
// mask_ratio of ProposalMaskTarget(output_ratio=True) - convertPoly2MaskWithRatio, proposal_mask_target.cc:20-152:
//   ratio = |polygon ∩ roi crop| / (|polygon| + 1e-4), both areas COUNTED on integer rasters produced by rleFrPoly:
//   the crop raster (crop_h x crop_w, roi corners truncated to int, polygon shifted by the roi corner) and the raster
//   of the polygon's own extent joined with the roi (full_h x full_w, shifted by the extent's corner).
// Those rasters are image-sized (up to ~1 M pixels per roi), but a count needs no pixels.  rleFrPoly's output is the
// sorted list of toggle positions a_0 <= a_1 <= ... (column-major pixel index): the segment covers [a_0,a_1) u
// [a_2,a_3) u ...  So per segment: generate the toggle positions (the same point-parallel walk as poly_mask_kernel),
// sort them, tag each with on (even rank) / off (odd rank); merge all segments' events by position (a second sort),
// prefix-sum the +1/-1 tags = how many segments cover the interval that starts at each event, and add up the
// intervals with coverage > 0 - the union over segments the reference forms pixel by pixel (:122-141).
// One CTA per (image, foreground row); events live in shared memory.  A roi whose polygon produces more events than
// the tables hold (kSegCap per segment, kEvCap in total - thousands of column crossings) gets ratio = NaN, loudly.
// --------------------------------------------------------------------------------------------
constexpr int kEvCap = 32768, kSegCap = 16384, kRatioThreads = 512;

__device__ void bitonic_sort_u32(unsigned* a, int n2) {  // n2 a power of two, every thread of the CTA calls it
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kRatioThreads) poly_ratio_kernel(const MaskParams p) {
  extern __shared__ unsigned s_ev[];  // kEvCap merged events (pos << 1 | on), then kSegCap positions of one segment
  unsigned* s_seg = s_ev + kEvCap;
  __shared__ EdgeRec s_edge[kMaxEdges];
  __shared__ int s_scan[kRatioThreads];
  __shared__ float s_red[4][kRatioThreads / 32];
  __shared__ int s_total, s_last[2], s_nseg_ev, s_nev, s_over, s_count;
  const int row = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  float* out = p.ratio + (size_t)b * p.NM + row;
  const int nfg = min(p.fg_count[b], p.NM);
  if (row >= nfg) {  // rows beyond the foreground count keep the initial 0 (-inl.h:244)
    if (tid == 0) *out = 0.f;
    return;
  }
  const int gi = p.gt_index[(size_t)b * p.IR + row];
  if (gi < 0) {  // no ground truth: nothing rasterised, 0 / 1e-4 clamped from below (:143-144)
    if (tid == 0) *out = (float)1e-10;
    return;
  }
  const float* roi = p.rois_out + ((size_t)b * p.IR + row) * 4;
  const float* poly = p.gt_polys + ((size_t)b * p.G + gi) * p.PL;
  const int n_seg = (int)poly[1];
  // ---- the polygon's extent joined with the roi (:45, :54-62): min / max over every vertex
  float mnx = roi[0], mxx = roi[2], mny = roi[1], mxy = roi[3];
  {
    int offset = 2 + n_seg;
    for (int sg = 0; sg < n_seg; ++sg) {
      const int k = (int)poly[sg + 2] / 2;
      for (int j = tid; j < k; j += blockDim.x) {
        const float x = poly[offset + 2 * j], y = poly[offset + 2 * j + 1];
        mnx = fmin_ref(mnx, x);
        mxx = fmax_ref(mxx, x);
        mny = fmin_ref(mny, y);
        mxy = fmax_ref(mxy, y);
      }
      offset += (int)poly[sg + 2];
    }
    for (int o = 16; o > 0; o >>= 1) {
      mnx = fmin_ref(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
      mxx = fmax_ref(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
      mny = fmin_ref(mny, __shfl_xor_sync(0xffffffffu, mny, o));
      mxy = fmax_ref(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((tid & 31) == 0) {
      s_red[0][tid >> 5] = mnx;
      s_red[1][tid >> 5] = mxx;
      s_red[2][tid >> 5] = mny;
      s_red[3][tid >> 5] = mxy;
    }
    if (tid == 0) s_over = 0;
    __syncthreads();
    for (int wi = 0; wi < kRatioThreads / 32; ++wi) {
      mnx = fmin_ref(mnx, s_red[0][wi]);
      mxx = fmax_ref(mxx, s_red[1][wi]);
      mny = fmin_ref(mny, s_red[2][wi]);
      mxy = fmax_ref(mxy, s_red[3][wi]);
    }
  }
  int counts[2];
  for (int z = 0; z < 2; ++z) {
    // raster size and the corner the polygon is shifted by
    double bx, by;
    int W, H;
    if (z == 0) {  // the roi crop (:41-46, :57, :64)
      const int x1 = (int)roi[0], x2 = (int)roi[2], y1 = (int)roi[1], y2 = (int)roi[3];
      W = x2 - x1 + 1;
      H = y2 - y1 + 1;
      bx = (double)roi[0];
      by = (double)roi[1];
    } else {       // the extent (:76-82, :89-96)
      W = (int)(double)mxx - (int)(double)mnx + 1;
      H = (int)(double)mxy - (int)(double)mny + 1;
      bx = (double)mnx;
      by = (double)mny;
    }
    if (W <= 0 || H <= 0 || (long long)W * H > (1ll << 30)) {
      if (tid == 0) s_over = 1;
      counts[z] = 0;
      __syncthreads();
      continue;
    }
    const int HW = W * H;
    if (tid == 0) s_nev = 0;
    __syncthreads();
    int offset = 2 + n_seg;
    for (int sg = 0; sg < n_seg; ++sg) {
      const int cur_len = (int)poly[sg + 2];
      const int k = cur_len / 2;
      if (tid == 0) s_nseg_ev = 0;
      __syncthreads();
      auto vert = [&](int j, int& X, int& Y) {  // natural (x, y) order here
        j = (j == k) ? 0 : j;
        const double a = __dsub_rn((double)poly[offset + 2 * j], bx);
        const double c = __dsub_rn((double)poly[offset + 2 * j + 1], by);
        X = (int)__dadd_rn(__dmul_rn(5.0, a), .5);
        Y = (int)__dadd_rn(__dmul_rn(5.0, c), .5);
      };
      for (int e0 = 0; e0 < k; e0 += kMaxEdges) {
        const int nb = min(kMaxEdges, k - e0);
        for (int e = tid; e < nb; e += blockDim.x) {
          int xs, ys, xe, ye;
          vert(e0 + e, xs, ys);
          vert(e0 + e + 1, xe, ye);
          const int dx = abs(xe - xs), dy = abs(ys - ye);
          const bool flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
          if (flip) {
            int t = xs; xs = xe; xe = t;
            t = ys; ys = ye; ye = t;
          }
          EdgeRec er;
          er.xs = xs; er.ys = ys; er.dx = dx; er.dy = dy; er.flip = flip;
          er.sl = dx >= dy ? __ddiv_rn((double)(ye - ys), (double)dx) : __ddiv_rn((double)(xe - xs), (double)dy);
          er.npts = max(dx, dy) + 1;
          er.start = 0;
          s_edge[e] = er;
        }
        __syncthreads();
        if (tid == 0) {
          int acc = 0;
          for (int e = 0; e < nb; ++e) {
            s_edge[e].start = acc;
            acc += s_edge[e].npts;
          }
          s_total = acc;
        }
        __syncthreads();
        const int total = s_total;
        for (int i = tid; i < total; i += blockDim.x) {
          int lo = 0, hi = nb - 1;
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_edge[mid].start <= i) lo = mid;
            else hi = mid - 1;
          }
          const EdgeRec er = s_edge[lo];
          const int d = i - er.start;
          int u, v, pu, pv;
          dda_point(er.xs, er.ys, er.dx, er.dy, er.sl, er.flip, d, u, v);
          if (d > 0) {
            dda_point(er.xs, er.ys, er.dx, er.dy, er.sl, er.flip, d - 1, pu, pv);
          } else if (lo > 0) {
            const EdgeRec pr = s_edge[lo - 1];
            dda_point(pr.xs, pr.ys, pr.dx, pr.dy, pr.sl, pr.flip, pr.npts - 1, pu, pv);
          } else if (e0 > 0) {
            pu = s_last[0];
            pv = s_last[1];
          } else {
            continue;
          }
          if (u != pu) {
            double xd = (double)(u < pu ? u : u - 1);
            xd = __dsub_rn(__ddiv_rn(__dadd_rn(xd, .5), 5.0), .5);
            if (!(floor(xd) != xd || xd < 0 || xd > (double)(W - 1))) {
              double yd = (double)(v < pv ? v : pv);
              yd = __dsub_rn(__ddiv_rn(__dadd_rn(yd, .5), 5.0), .5);
              if (yd < 0) yd = 0;
              else if (yd > (double)H) yd = (double)H;
              yd = ceil(yd);
              const int pos = (int)xd * H + (int)yd;  // <= H*W
              const int slot = atomicAdd(&s_nseg_ev, 1);
              if (slot < kSegCap) s_seg[slot] = (unsigned)pos;
            }
          }
        }
        __syncthreads();
        if (tid == 0) {
          const EdgeRec pr = s_edge[nb - 1];
          int lu, lv;
          dda_point(pr.xs, pr.ys, pr.dx, pr.dy, pr.sl, pr.flip, pr.npts - 1, lu, lv);
          s_last[0] = lu;
          s_last[1] = lv;
        }
        __syncthreads();
      }
      // ---- this segment's positions in order: even rank switches the segment on, odd rank off
      const int nraw = s_nseg_ev, base = s_nev;
      const int n = min(nraw, kSegCap);
      if (nraw > kSegCap || base + n > kEvCap) {
        if (tid == 0) s_over = 1;
      } else if (n > 0) {
        int n2 = 1;
        while (n2 < n) n2 <<= 1;
        for (int i = n + tid; i < n2; i += blockDim.x) s_seg[i] = 0xffffffffu;
        __syncthreads();
        bitonic_sort_u32(s_seg, n2);
        for (int i = tid; i < n; i += blockDim.x) s_ev[base + i] = (s_seg[i] << 1) | ((i & 1) ? 0u : 1u);
      }
      __syncthreads();
      if (tid == 0 && !(nraw > kSegCap || base + n > kEvCap)) s_nev = base + n;
      __syncthreads();
      offset += cur_len;
    }
    // ---- merge by position, coverage = running sum of on / off, add up the covered intervals
    const int n = s_nev;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = n + tid; i < n2; i += blockDim.x) s_ev[i] = 0xffffffffu;
    if (tid == 0) s_count = 0;
    __syncthreads();
    if (n > 1) bitonic_sort_u32(s_ev, n2);
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int i0 = min(n, tid * per), i1 = min(n, i0 + per);
    int local = 0;
    for (int i = i0; i < i1; ++i) local += (s_ev[i] & 1u) ? 1 : -1;
    s_scan[tid] = local;
    __syncthreads();
    if (tid == 0) {  // exclusive prefix of the per-thread sums
      int acc = 0;
      for (int t = 0; t < (int)blockDim.x; ++t) {
        const int v = s_scan[t];
        s_scan[t] = acc;
        acc += v;
      }
    }
    __syncthreads();
    int cover = s_scan[tid], area = 0;
    for (int i = i0; i < i1; ++i) {
      const unsigned e = s_ev[i];
      cover += (e & 1u) ? 1 : -1;
      const int pos = (int)(e >> 1);
      const int nxt = (i + 1 < n) ? (int)(s_ev[i + 1] >> 1) : HW;
      if (cover > 0) area += nxt - pos;
    }
    for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(0xffffffffu, area, o);
    if ((tid & 31) == 0 && area) atomicAdd(&s_count, area);
    __syncthreads();
    counts[z] = s_count;
    __syncthreads();
  }
  if (tid == 0) {
    if (s_over) {
      *out = __int_as_float(0x7fc00000);
    } else {
      double r = __ddiv_rn((double)counts[0], __dadd_rn((double)counts[1], 0.0001));  // :143
      r = r < 1e-10 ? 1e-10 : r;                                                          // :144 max(ratio, 1e-10)
      *out = (float)r;
    }
  }
}

size_t pt_smem_bytes(int T, int G, int IR) {
  int np2 = 1;
  while (np2 < T) np2 <<= 1;
  return (size_t)np2 * 8 + (size_t)T * 16 + (size_t)G * 5 * 4 + (size_t)T * 4 * 7 + (size_t)(IR + T) * 4 +
         (size_t)G * 4 + 64;
}

}  // namespace

static int proposal_target_core(const float* rois, const float* gt_boxes, const float* valid_ranges,
                                float* rois_out, float* labels, float* bbox_targets, float* bbox_weights,
                                float* match_gt_ious, int* kept, int B, int R, int G, int num_classes,
                                int image_rois, float fg_fraction, float fg_thresh, float bg_thresh_hi,
                                float bg_thresh_lo, int proposal_without_gt, int class_agnostic,
                                int filter_scales, const float* bbox_mean, const float* bbox_std,
                                const float* bbox_weight, unsigned long long seed, const uint32_t* priorities,
                                int num_draws, uint32_t* priorities_used, int* gt_index, int* fg_count,
                                void* stream) {
  const int no_fg_cap = (image_rois == -1);  // ProposalTarget_v2: keep all foreground rois, R rows
  if (no_fg_cap) image_rois = R;
  SDET_REQUIRE(rois && gt_boxes && rois_out && labels && bbox_targets && bbox_weights && match_gt_ious &&
               bbox_mean && bbox_std && bbox_weight, "NULL argument");
  SDET_REQUIRE(B > 0 && R > 0 && G >= 0 && num_classes > 0 && image_rois > 0, "bad shape");
  SDET_REQUIRE(num_draws >= 3, "num_draws must be >= 3 (fg, bg, >= 1 negative-padding draw)");
  if (R + G > kMaxT) return sdet::fail(SDET_ERR_UNSUPPORTED, "rois + gt per image > %d", kMaxT);
  PTParams p{};
  p.rois = rois; p.gt = gt_boxes; p.prio = priorities; p.prio_used = priorities_used;
  p.rois_out = rois_out; p.labels = labels; p.tgt = bbox_targets; p.wgt = bbox_weights;
  p.iou = match_gt_ious; p.kept = kept; p.gt_index = gt_index; p.fg_count = fg_count;
  p.B = B; p.R = R; p.G = G; p.NC4 = num_classes * 4; p.IR = image_rois; p.D = num_draws;
  p.fg_per_image = (int)(image_rois * fg_fraction);  // index_t truncation, proposal_target-inl.h:194
  p.fg_thresh = fg_thresh; p.bg_hi = bg_thresh_hi; p.bg_lo = bg_thresh_lo;
  p.without_gt = proposal_without_gt ? 1 : 0;
  p.agnostic = class_agnostic ? 1 : 0;
  p.valid_ranges = valid_ranges;
  p.filter_scales = filter_scales ? 1 : 0;
  p.no_fg_cap = no_fg_cap;
  for (int i = 0; i < 4; ++i) {
    p.mean[i] = bbox_mean[i];
    p.std[i] = bbox_std[i];
    p.weight[i] = bbox_weight[i];
  }
  p.seed = seed;
  const size_t smem = pt_smem_bytes(R + G, G, image_rois);
  if (smem > 220 * 1024) return sdet::fail(SDET_ERR_UNSUPPORTED, "ProposalTarget needs %zu B shared memory", smem);
  if (smem > 48 * 1024)  // per device and cheap: set on every launch, no process-wide cache
    SDET_CUDA(cudaFuncSetAttribute(proposal_target_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SDET_CUDA(cudaMemsetAsync(bbox_targets, 0, sizeof(float) * (size_t)B * image_rois * p.NC4, (cudaStream_t)stream));
  SDET_CUDA(cudaMemsetAsync(bbox_weights, 0, sizeof(float) * (size_t)B * image_rois * p.NC4, (cudaStream_t)stream));
  proposal_target_kernel<<<(unsigned)B, kThreads, smem, (cudaStream_t)stream>>>(p);
  SDET_LAUNCH_CHECK("proposal_target_kernel");
  return SDET_OK;
}

extern "C" int sdet_proposal_target(const float* rois, const float* gt_boxes, float* rois_out,
                                    float* labels, float* bbox_targets, float* bbox_weights,
                                    float* match_gt_ious, int* kept, int B, int R, int G, int num_classes,
                                    int image_rois, float fg_fraction, float fg_thresh, float bg_thresh_hi,
                                    float bg_thresh_lo, int proposal_without_gt, int class_agnostic,
                                    const float* bbox_mean, const float* bbox_std, const float* bbox_weight,
                                    unsigned long long seed, const uint32_t* priorities, int num_draws,
                                    uint32_t* priorities_used, int* gt_index, int* fg_count, void* stream) {
  if (image_rois <= 0) return sdet::fail(SDET_ERR_INVALID_ARG, "image_rois must be > 0 (ProposalTarget_v2 takes -1)");
  return proposal_target_core(rois, gt_boxes, nullptr, rois_out, labels, bbox_targets, bbox_weights, match_gt_ious,
                              kept, B, R, G, num_classes, image_rois, fg_fraction, fg_thresh, bg_thresh_hi,
                              bg_thresh_lo, proposal_without_gt, class_agnostic, 0, bbox_mean, bbox_std, bbox_weight,
                              seed, priorities, num_draws, priorities_used, gt_index, fg_count, stream);
}

extern "C" int sdet_proposal_target_v2(const float* rois, const float* gt_boxes, const float* valid_ranges,
                                       float* rois_out, float* labels, float* bbox_targets,
                                       float* bbox_weights, float* match_gt_ious, int* kept, int B, int R,
                                       int G, int num_classes, int image_rois, float fg_fraction,
                                       float fg_thresh, float bg_thresh_hi, float bg_thresh_lo,
                                       int proposal_without_gt, int class_agnostic, int filter_scales,
                                       const float* bbox_mean, const float* bbox_std, const float* bbox_weight,
                                       unsigned long long seed, const uint32_t* priorities, int num_draws,
                                       uint32_t* priorities_used, int* gt_index, int* fg_count, void* stream) {
  if (image_rois <= 0 && image_rois != -1)
    return sdet::fail(SDET_ERR_INVALID_ARG, "image_rois must be > 0 or -1");
  if (filter_scales && !valid_ranges) return sdet::fail(SDET_ERR_INVALID_ARG, "filter_scales needs valid_ranges");
  return proposal_target_core(rois, gt_boxes, valid_ranges, rois_out, labels, bbox_targets, bbox_weights,
                              match_gt_ious, kept, B, R, G, num_classes, image_rois, fg_fraction, fg_thresh,
                              bg_thresh_hi, bg_thresh_lo, proposal_without_gt, class_agnostic, filter_scales,
                              bbox_mean, bbox_std, bbox_weight, seed, priorities, num_draws, priorities_used, gt_index,
                              fg_count, stream);
}

// mask_ratio == nullptr: convertPoly2Mask; else convertPoly2MaskWithRatio (double vertex transform + the ratio)
static int run_poly_mask(const float* rois_out, const float* gt_polys, const int* gt_index, const int* fg_count,
                         float* mask_target, float* mask_ratio, int B, int image_rois, int G, int poly_len,
                         int num_mask_rows, int mask_size, void* stream) {
  MaskParams p{rois_out, gt_polys, gt_index, fg_count, mask_target, image_rois, G, poly_len, num_mask_rows, mask_size,
               mask_ratio ? 1 : 0, mask_ratio};
  const size_t smem = sizeof(int) * (size_t)(2 * mask_size * mask_size + 2);
  if (smem > 48 * 1024)  // per device and cheap: set on every launch, no process-wide cache
    SDET_CUDA(cudaFuncSetAttribute(poly_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)num_mask_rows, (unsigned)B);
  poly_mask_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  SDET_LAUNCH_CHECK("poly_mask_kernel");
  if (mask_ratio) {
    const size_t rsmem = sizeof(unsigned) * (size_t)(kEvCap + kSegCap);
    SDET_CUDA(cudaFuncSetAttribute(poly_ratio_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
    poly_ratio_kernel<<<grid, kRatioThreads, rsmem, (cudaStream_t)stream>>>(p);
    SDET_LAUNCH_CHECK("poly_ratio_kernel");
  }
  return SDET_OK;
}

extern "C" int sdet_poly_mask_target(const float* rois_out, const float* gt_polys, const int* gt_index,
                                     const int* fg_count, float* mask_target, int B, int image_rois, int G,
                                     int poly_len, int num_mask_rows, int mask_size, void* stream) {
  SDET_REQUIRE(rois_out && gt_polys && gt_index && fg_count && mask_target, "NULL argument");
  SDET_REQUIRE(B > 0 && image_rois > 0 && G > 0 && poly_len > 2 && num_mask_rows > 0 && mask_size > 0, "bad shape");
  SDET_REQUIRE(num_mask_rows <= image_rois, "mask rows exceed image_rois");
  if (mask_size > 112) return sdet::fail(SDET_ERR_UNSUPPORTED, "mask_size > 112");
  return run_poly_mask(rois_out, gt_polys, gt_index, fg_count, mask_target, nullptr, B, image_rois, G, poly_len,
                       num_mask_rows, mask_size, stream);
}

extern "C" int sdet_poly_mask_target_ratio(const float* rois_out, const float* gt_polys, const int* gt_index,
                                           const int* fg_count, float* mask_target, float* mask_ratio, int B,
                                           int image_rois, int G, int poly_len, int num_mask_rows, int mask_size,
                                           void* stream) {
  SDET_REQUIRE(rois_out && gt_polys && gt_index && fg_count && mask_target && mask_ratio, "NULL argument");
  SDET_REQUIRE(B > 0 && image_rois > 0 && G > 0 && poly_len > 2 && num_mask_rows > 0 && mask_size > 0, "bad shape");
  SDET_REQUIRE(num_mask_rows <= image_rois, "mask rows exceed image_rois");
  if (mask_size > 112) return sdet::fail(SDET_ERR_UNSUPPORTED, "mask_size > 112");
  return run_poly_mask(rois_out, gt_polys, gt_index, fg_count, mask_target, mask_ratio, B, image_rois, G, poly_len,
                       num_mask_rows, mask_size, stream);
}
