/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * CPU restatement (plain C, fp32) of the reference's box decode / proposal / NMS operators.
 * Compile with -ffp-contract=off (see roi_ops.c).  Source-level semantics: where the reference
 * only has a .cu (Proposal_v3, _contrib_NMS) the restatement follows the .cu statement by
 * statement WITHOUT FMA contraction; nvcc's contraction of the reference build is a property of
 * its compiler flags, not of the algorithm.
 *
 * Parity pin: the reference holds no fixtures for these ops, so they are pinned against the
 * reference's own code, compiled here into oracle/_ref (test infrastructure, git-ignored):
 *   - greedy_nms / soft_nms / bbox_overlaps: its Cython (operator_py/cython/{cpu_nms,bbox}.pyx), oracle/build_ref.py,
 *     tests/test_oracle_ref.py;
 *   - decode_bbox, gen_anchor, proposal_v3, proposal (v1), proposal_v2, contrib_nms, gen_proposal,
 *     gen_proposal_retina: its operator_cxx sources, unmodified, through oracle/shim (the .cu operators run on
 *     the host: every `<<<>>>` launch becomes a serial loop over all threads), oracle/build_ref_cxx.py,
 *     tests/test_oracle_ref_cxx.py - bit for bit.
 * The same cases are committed as vectors under tests/golden/ (tests/test_oracle_golden*.py run anywhere).
 *
 *   oracle_decode_bbox        operator_cxx/contrib/decodebbox.cc:34-133
 *   oracle_proposal_v3        operator_cxx/contrib/proposal_v3.cu:65-419,463-637 (+ anchors
 *                             proposal_v3-inl.h:280-318)
 *   oracle_contrib_nms        operator_cxx/contrib/nms.cu:102-236,274-364
 *   oracle_bbox_overlaps      operator_py/cython/bbox.pyx:32-73
 *   oracle_greedy_nms         operator_py/cython/cpu_nms.pyx:37-87 (order supplied by caller)
 *   oracle_soft_nms           operator_py/cython/cpu_nms.pyx:98-203
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float fmin_(float a, float b) { return a < b ? a : b; } /* std::min / CUDA min */
static inline float fmax_(float a, float b) { return a < b ? b : a; } /* std::max / CUDA max */

/* ---- _contrib_DecodeBBox (decodebbox.cc:34-133).  rois (B,N,4), deltas (B,N,K4), im_info (B,3).
 * out (B,N,4) if class_agnostic else (B,N,K4).  decode_type 0 = xywh, 1 = xyxy. ---- */
void oracle_decode_bbox(const float* rois, const float* deltas, const float* im_info, int B, int N,
                        int K4, const float* means, const float* stds, int class_agnostic,
                        int decode_type, float* out) {
  const int num_class = class_agnostic ? 1 : K4 / 4;
  const int out_dim = class_agnostic ? 4 : K4;
  for (int n = 0; n < B; ++n)
    for (int i = 0; i < N; ++i)
      for (int cls = 0; cls < num_class; ++cls) {
        const float* b = rois + ((long)n * N + i) * 4;
        const float* d = deltas + ((long)n * N + i) * K4 + (class_agnostic ? 1 : cls) * 4;
        float* o = out + ((long)n * N + i) * out_dim + cls * 4;
        const float im_h = im_info[n * 3 + 0], im_w = im_info[n * 3 + 1];
        float x1, y1, x2, y2;
        if (decode_type == 0) {
          float width = b[2] - b[0] + 1.0f;
          float height = b[3] - b[1] + 1.0f;
          float ctr_x = b[0] + 0.5f * (width - 1.0f);
          float ctr_y = b[1] + 0.5f * (height - 1.0f);
          float dx = d[0] * stds[0] + means[0];
          float dy = d[1] * stds[1] + means[1];
          float dw = d[2] * stds[2] + means[2];
          float dh = d[3] * stds[3] + means[3];
          float pred_ctr_x = dx * width + ctr_x;
          float pred_ctr_y = dy * height + ctr_y;
          /* decodebbox.cc:62-63 `exp(dw) * width` with a float dw: the unqualified call binds to ::exp(double)
           * (the file pulls in <cmath> only, no float overload in the global namespace), the product is formed
           * in double and narrowed once.  Pinned by the compiled reference (tests/test_oracle_ref_cxx.py). */
          float pred_w = (float)(exp((double)dw) * (double)width);
          float pred_h = (float)(exp((double)dh) * (double)height);
          x1 = pred_ctr_x - 0.5f * (pred_w - 1.0f);
          y1 = pred_ctr_y - 0.5f * (pred_h - 1.0f);
          x2 = pred_ctr_x + 0.5f * (pred_w - 1.0f);
          y2 = pred_ctr_y + 0.5f * (pred_h - 1.0f);
        } else {
          float width = b[2] - b[0] + 1.0f;
          float height = b[3] - b[1] + 1.0f;
          x1 = b[0] + (d[0] * stds[0] + means[0]) * width;
          y1 = b[1] + (d[1] * stds[1] + means[1]) * height;
          x2 = b[2] + (d[2] * stds[2] + means[2]) * width;
          y2 = b[3] + (d[3] * stds[3] + means[3]) * height;
        }
        o[0] = fmax_(fmin_(x1, im_w - 1.0f), 0.0f);
        o[1] = fmax_(fmin_(y1, im_h - 1.0f), 0.0f);
        o[2] = fmax_(fmin_(x2, im_w - 1.0f), 0.0f);
        o[3] = fmax_(fmin_(y2, im_h - 1.0f), 0.0f);
      }
}

/* proposal_v3-inl.h:280-318: (ratios outer, scales inner), 5 floats per anchor. */
void oracle_generate_anchors_v3(int feature_stride, const float* ratios, int nr, const float* scales,
                                int ns, float* anchors /* nr*ns*4 */) {
  const float base[4] = {0.f, 0.f, (float)(feature_stride - 1.0), (float)(feature_stride - 1.0)};
  int k = 0;
  for (int j = 0; j < nr; ++j)
    for (int s = 0; s < ns; ++s) {
      float w = base[2] - base[0] + 1.0f;
      float h = base[3] - base[1] + 1.0f;
      float x_ctr = (float)(base[0] + 0.5 * (w - 1.0f)); /* `0.5` is double in the source */
      float y_ctr = (float)(base[1] + 0.5 * (h - 1.0f));
      float size = w * h;
      float size_ratios = floorf(size / ratios[j]);
      float new_w = rintf(sqrtf(size_ratios)) * scales[s];
      float new_h = rintf((new_w / scales[s] * ratios[j])) * scales[s];
      anchors[k * 4 + 0] = x_ctr - 0.5f * (new_w - 1.0f);
      anchors[k * 4 + 1] = y_ctr - 0.5f * (new_h - 1.0f);
      anchors[k * 4 + 2] = x_ctr + 0.5f * (new_w - 1.0f);
      anchors[k * 4 + 3] = y_ctr + 0.5f * (new_h - 1.0f);
      ++k;
    }
}

/* devIoU (proposal_v3.cu:271-279 / nms.cu:91-99) */
static inline float dev_iou(const float* a, const float* b) {
  float left = fmax_(a[0], b[0]), right = fmin_(a[2], b[2]);
  float top = fmax_(a[1], b[1]), bottom = fmin_(a[3], b[3]);
  float width = fmax_(right - left + 1, 0.f), height = fmax_(bottom - top + 1, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
  float Sb = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  return interS / (Sa + Sb - interS);
}

/* nms_kernel + host scan of _nms (proposal_v3.cu:281-380 / nms.cu:102-202): greedy over the
 * given order; box j is removed by kept box i<j when IoU >= thr (ge=1) or > thr (ge=0). */
static int greedy_scan(const float* dets5, int n, float thr, int ge, int* keep) {
  unsigned char* removed = (unsigned char*)calloc((size_t)n, 1);
  int nk = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep[nk++] = i;
    for (int j = i + 1; j < n; ++j) {
      float v = dev_iou(dets5 + (long)i * 5, dets5 + (long)j * 5);
      if (ge ? (v >= thr) : (v > thr)) removed[j] = 1;
    }
  }
  free(removed);
  return nk;
}

/* stable descending argsort (thrust::stable_sort_by_key with greater<float>) */
typedef struct { float s; int i; } si_t;
static int cmp_desc_stable(const void* a, const void* b) {
  const si_t *x = (const si_t*)a, *y = (const si_t*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);
}
void oracle_stable_argsort_desc(const float* score, int n, int* order) {
  si_t* t = (si_t*)malloc(sizeof(si_t) * (size_t)n);
  for (int i = 0; i < n; ++i) { t[i].s = score[i]; t[i].i = i; }
  qsort(t, (size_t)n, sizeof(si_t), cmp_desc_stable);
  for (int i = 0; i < n; ++i) order[i] = t[i].i;
  free(t);
}

/* ---- _contrib_Proposal_v3, GPU semantics (proposal_v3.cu:435-638).
 * cls_prob (B,2A,H,W), bbox_pred (B,4A,H,W), im_info (B,3) -> out (B,post,4), out_score (B,post).
 * dbg_dets (B,pre,5) and dbg_keep (B,pre)+dbg_nkeep (B) are optional stage outputs for tests. */
void oracle_proposal_v3(const float* cls_prob, const float* bbox_pred, const float* im_info, int B,
                        int A, int H, int W, int feature_stride, const float* scales, int ns,
                        const float* ratios, int nr, int pre_nms_top_n, int post_nms_top_n,
                        float threshold, int rpn_min_size, int iou_loss, int is_train, float* out,
                        float* out_score, float* dbg_dets, int* dbg_keep, int* dbg_nkeep) {
  const int count = A * H * W;
  int pre = pre_nms_top_n > 0 ? pre_nms_top_n : count;
  if (pre > count) pre = count;
  int post = post_nms_top_n < pre ? post_nms_top_n : pre;
  if (!is_train) post = post_nms_top_n;
  float* anchors = (float*)malloc(sizeof(float) * 4 * (size_t)A);
  oracle_generate_anchors_v3(feature_stride, ratios, nr, scales, ns, anchors);
  float* prop = (float*)malloc(sizeof(float) * 5 * (size_t)count);
  float* score = (float*)malloc(sizeof(float) * (size_t)count);
  int* order = (int*)malloc(sizeof(int) * (size_t)count);
  float* dets = (float*)malloc(sizeof(float) * 5 * (size_t)pre);
  int* keep = (int*)calloc((size_t)pre, sizeof(int));
  for (int b = 0; b < B; ++b) {
    const float im_h = im_info[b * 3 + 0], im_w = im_info[b * 3 + 1], im_s = im_info[b * 3 + 2];
    const int real_h = (int)(im_h / feature_stride), real_w = (int)(im_w / feature_stride);
    const float* fg = cls_prob + (long)b * 2 * count + count; /* second half = foreground */
    const float* dl = bbox_pred + (long)b * 4 * count;
    for (int index = 0; index < count; ++index) {
      int a = index % A, w = (index / A) % W, h = index / A / W;
      /* ProposalGridKernel :65-85 */
      float bx1 = anchors[a * 4 + 0] + w * feature_stride;
      float by1 = anchors[a * 4 + 1] + h * feature_stride;
      float bx2 = anchors[a * 4 + 2] + w * feature_stride;
      float by2 = anchors[a * 4 + 3] + h * feature_stride;
      float sc = fg[(a * H + h) * W + w];
      float d0 = dl[((a * 4 + 0) * H + h) * W + w], d1 = dl[((a * 4 + 1) * H + h) * W + w];
      float d2 = dl[((a * 4 + 2) * H + h) * W + w], d3 = dl[((a * 4 + 3) * H + h) * W + w];
      float x1, y1, x2, y2;
      if (iou_loss) { /* IoUPredKernel :163-205 */
        x1 = fmax_(fmin_(bx1 + d0, im_w - 1.0f), 0.0f);
        y1 = fmax_(fmin_(by1 + d1, im_h - 1.0f), 0.0f);
        x2 = fmax_(fmin_(bx2 + d2, im_w - 1.0f), 0.0f);
        y2 = fmax_(fmin_(by2 + d3, im_h - 1.0f), 0.0f);
        if (h >= real_h || w >= real_w) sc = -1.0f;
      } else { /* BBoxPredKernel :93-155 */
        float width = bx2 - bx1 + 1.0f, height = by2 - by1 + 1.0f;
        float ctr_x = bx1 + 0.5f * width, ctr_y = by1 + 0.5f * height;
        /* min(float, double literal): compare in double, narrow */
        float dw = (float)((double)d2 < 4.135166556742356 ? (double)d2 : 4.135166556742356);
        float dh = (float)((double)d3 < 4.135166556742356 ? (double)d3 : 4.135166556742356);
        float pcx = d0 * width + ctr_x, pcy = d1 * height + ctr_y;
        float pw = expf(dw) * width, ph = expf(dh) * height;
        x1 = pcx - 0.5f * pw;
        y1 = pcy - 0.5f * ph;
        x2 = pcx + 0.5f * pw - 1.0f;
        y2 = pcy + 0.5f * ph - 1.0f;
        x1 = fmax_(fmin_(x1, im_w - 1.0f), 0.0f);
        y1 = fmax_(fmin_(y1, im_h - 1.0f), 0.0f);
        x2 = fmax_(fmin_(x2, im_w - 1.0f), 0.0f);
        y2 = fmax_(fmin_(y2, im_h - 1.0f), 0.0f);
      }
      float* p = prop + (long)index * 5;
      p[0] = x1; p[1] = y1; p[2] = x2; p[3] = y2; p[4] = sc;
      score[index] = sc;
    }
    oracle_stable_argsort_desc(score, count, order); /* :564-568 */
    for (int i = 0; i < pre; ++i) memcpy(dets + (long)i * 5, prop + (long)order[i] * 5, 5 * sizeof(float));
    for (int i = 0; i < pre; ++i) { /* FilterBoxKernel :211-235 */
      float* d = dets + (long)i * 5;
      float ws_o = (d[2] - d[0]) / im_s + 1.0f, hs_o = (d[3] - d[1]) / im_s + 1.0f;
      float msm = fmax_((float)rpn_min_size, 1.0f);
      float ws = d[2] - d[0] + 1.0f, hs = d[3] - d[1] + 1.0f;
      float x_ctr = d[0] + ws / 2.0f, y_ctr = d[1] + hs / 2.0f;
      if (ws_o < msm || hs_o < msm || x_ctr >= im_w || y_ctr >= im_h) {
        d[0] -= msm / 2; d[1] -= msm / 2; d[2] += msm / 2; d[3] += msm / 2; d[4] = -1.0f;
      }
    }
    int nk = greedy_scan(dets, pre, threshold, /*ge=*/1, keep); /* :319 uses >= */
    for (int i = 0; i < post; ++i) { /* PrepareOutput :387-419 */
      float* o = out + ((long)b * post + i) * 4;
      if (i < nk) {
        memcpy(o, dets + (long)keep[i] * 5, 4 * sizeof(float));
        out_score[(long)b * post + i] = dets[(long)keep[i] * 5 + 4];
      } else if (is_train) {
        int k = keep[i % nk];
        memcpy(o, dets + (long)k * 5, 4 * sizeof(float));
        out_score[(long)b * post + i] = dets[(long)k * 5 + 4];
      } else {
        o[0] = o[1] = o[2] = o[3] = 0.f;
        out_score[(long)b * post + i] = 0.f;
      }
    }
    if (dbg_dets) memcpy(dbg_dets + (long)b * pre * 5, dets, sizeof(float) * 5 * (size_t)pre);
    if (dbg_keep) memcpy(dbg_keep + (long)b * pre, keep, sizeof(int) * (size_t)pre);
    if (dbg_nkeep) dbg_nkeep[b] = nk;
  }
  free(anchors); free(prop); free(score); free(order); free(dets); free(keep);
}


/* ---- _contrib_Proposal (version 1, operator_cxx/contrib/proposal.cu:65-420,430-620) and
 * _contrib_Proposal_v2 (version 2, proposal_v2.cu): legacy pipeline.  Differences from v3:
 * anchors round with floor(x + 0.5) (proposal-inl.h:302-303), legacy decode without the exp clip,
 * padded cells (h >= im_h/stride or w >= im_w/stride) get score -1, the min-size filter
 * (rpn_min_size * im_scale, on the decoded box) runs on ALL anchors BEFORE the sort, NMS removes
 * IoU > thr.  v2 adds the valid-range filter (proposal_v2.cu:217-219), always zero-pads and has
 * post = min(post, pre). ---- */
void oracle_generate_anchors_legacy(int feature_stride, const float* ratios, int nr, const float* scales,
                                    int ns, float* anchors) {
  const float b2 = (float)(feature_stride - 1.0);
  int k = 0;
  for (int j = 0; j < nr; ++j)
    for (int s = 0; s < ns; ++s) {
      float w = b2 - 0.f + 1.0f, h = b2 - 0.f + 1.0f;
      float x_ctr = (float)(0.f + 0.5 * (w - 1.0f)), y_ctr = (float)(0.f + 0.5 * (h - 1.0f));
      float size = w * h;
      float size_ratios = floorf(size / ratios[j]);
      float new_w = floorf(sqrtf(size_ratios) + 0.5f) * scales[s];
      float new_h = floorf((new_w / scales[s] * ratios[j]) + 0.5f) * scales[s];
      anchors[k * 4 + 0] = x_ctr - 0.5f * (new_w - 1.0f);
      anchors[k * 4 + 1] = y_ctr - 0.5f * (new_h - 1.0f);
      anchors[k * 4 + 2] = x_ctr + 0.5f * (new_w - 1.0f);
      anchors[k * 4 + 3] = y_ctr + 0.5f * (new_h - 1.0f);
      ++k;
    }
}

void oracle_proposal_legacy(const float* cls_prob, const float* bbox_pred, const float* im_info,
                            const float* valid_ranges, int version, int B, int A, int H, int W,
                            int feature_stride, const float* scales, int ns, const float* ratios, int nr,
                            int pre_nms_top_n, int post_nms_top_n, float threshold, int rpn_min_size,
                            int iou_loss, int is_train, int filter_scales, float* out, float* out_score) {
  const int count = A * H * W;
  int pre = pre_nms_top_n > 0 ? pre_nms_top_n : count;
  if (pre > count) pre = count;
  int post = post_nms_top_n < pre ? post_nms_top_n : pre;
  if (version == 1 && !is_train) post = post_nms_top_n; /* proposal.cu:453-455 */
  float* anchors = (float*)malloc(sizeof(float) * 4 * (size_t)A);
  oracle_generate_anchors_legacy(feature_stride, ratios, nr, scales, ns, anchors);
  float* prop = (float*)malloc(sizeof(float) * 5 * (size_t)count);
  float* score = (float*)malloc(sizeof(float) * (size_t)count);
  int* order = (int*)malloc(sizeof(int) * (size_t)count);
  float* dets = (float*)malloc(sizeof(float) * 5 * (size_t)pre);
  int* keep = (int*)calloc((size_t)pre, sizeof(int));
  for (int b = 0; b < B; ++b) {
    const float im_h = im_info[b * 3 + 0], im_w = im_info[b * 3 + 1], im_s = im_info[b * 3 + 2];
    const int real_h = (int)(im_h / feature_stride), real_w = (int)(im_w / feature_stride);
    const float* fg = cls_prob + (long)b * 2 * count + count;
    const float* dl = bbox_pred + (long)b * 4 * count;
    const float min_size = rpn_min_size * im_s; /* int * float (proposal.cu:533) */
    for (int index = 0; index < count; ++index) {
      int a = index % A, w = (index / A) % W, h = index / A / W;
      float bx1 = anchors[a * 4 + 0] + w * feature_stride, by1 = anchors[a * 4 + 1] + h * feature_stride;
      float bx2 = anchors[a * 4 + 2] + w * feature_stride, by2 = anchors[a * 4 + 3] + h * feature_stride;
      float sc = fg[(a * H + h) * W + w];
      float d0 = dl[((a * 4 + 0) * H + h) * W + w], d1 = dl[((a * 4 + 1) * H + h) * W + w];
      float d2 = dl[((a * 4 + 2) * H + h) * W + w], d3 = dl[((a * 4 + 3) * H + h) * W + w];
      float x1, y1, x2, y2;
      if (iou_loss) {
        x1 = bx1 + d0; y1 = by1 + d1; x2 = bx2 + d2; y2 = by2 + d3;
      } else { /* BBoxPredKernel proposal.cu:93-145 */
        float width = bx2 - bx1 + 1.0f, height = by2 - by1 + 1.0f;
        float ctr_x = bx1 + 0.5f * (width - 1.0f), ctr_y = by1 + 0.5f * (height - 1.0f);
        float pcx = d0 * width + ctr_x, pcy = d1 * height + ctr_y;
        float pw = expf(d2) * width, ph = expf(d3) * height;
        x1 = pcx - 0.5f * (pw - 1.0f); y1 = pcy - 0.5f * (ph - 1.0f);
        x2 = pcx + 0.5f * (pw - 1.0f); y2 = pcy + 0.5f * (ph - 1.0f);
      }
      x1 = fmax_(fmin_(x1, im_w - 1.0f), 0.0f); y1 = fmax_(fmin_(y1, im_h - 1.0f), 0.0f);
      x2 = fmax_(fmin_(x2, im_w - 1.0f), 0.0f); y2 = fmax_(fmin_(y2, im_h - 1.0f), 0.0f);
      if (h >= real_h || w >= real_w) sc = -1.0f;
      /* FilterBoxKernel (proposal.cu:207-224 / proposal_v2.cu:200-222), before the sort */
      float iw = x2 - x1 + 1.0f, ih = y2 - y1 + 1.0f;
      if (iw < min_size || ih < min_size) {
        x1 -= min_size / 2; y1 -= min_size / 2; x2 += min_size / 2; y2 += min_size / 2; sc = -1.0f;
      } else if (version == 2 && filter_scales) {
        float vmin = valid_ranges[b * 2] * valid_ranges[b * 2], vmax = valid_ranges[b * 2 + 1] * valid_ranges[b * 2 + 1];
        if (iw * ih < vmin || iw * ih > vmax) sc = -1.0f;
      }
      float* p = prop + (long)index * 5;
      p[0] = x1; p[1] = y1; p[2] = x2; p[3] = y2; p[4] = sc;
      score[index] = sc;
    }
    oracle_stable_argsort_desc(score, count, order);
    for (int i = 0; i < pre; ++i) memcpy(dets + (long)i * 5, prop + (long)order[i] * 5, 5 * sizeof(float));
    int nk = greedy_scan(dets, pre, threshold, /*ge=*/0, keep); /* proposal.cu:301 uses > */
    for (int i = 0; i < post; ++i) {
      float* o = out + ((long)b * post + i) * 4;
      int k = -1;
      if (i < nk) k = keep[i];
      else if (version == 1 && is_train) k = keep[i % nk];
      if (k >= 0) { memcpy(o, dets + (long)k * 5, 16); out_score[(long)b * post + i] = dets[(long)k * 5 + 4]; }
      else { o[0] = o[1] = o[2] = o[3] = 0.f; out_score[(long)b * post + i] = 0.f; }
    }
  }
  free(anchors); free(prop); free(score); free(order); free(dets); free(keep);
}


/* ---- _contrib_GenAnchor (generate_anchor-inl.h:139-183 GenerateAnchors in double with rint;
 * generate_anchor.cu:62-81 AnchorGridKernel casts double + int shift to float). ---- */
void oracle_gen_anchor(int H, int W, int feature_stride, const double* scales, int ns, const double* ratios,
                       int nr, float* out) {
  const int A = ns * nr;
  double* base = (double*)malloc(sizeof(double) * 4 * (size_t)A);
  const double b2 = feature_stride - 1.0f;
  int k = 0;
  for (int j = 0; j < nr; ++j)
    for (int s = 0; s < ns; ++s) {
      double w = b2 - 0.0 + 1.0f, h = b2 - 0.0 + 1.0f;
      double x_ctr = 0.0 + 0.5 * (w - 1.0f), y_ctr = 0.0 + 0.5 * (h - 1.0f);
      double size_ratios = (w * h) / ratios[j];
      double new_w = rint(sqrt(size_ratios)) * scales[s];
      double new_h = rint((new_w / scales[s] * ratios[j])) * scales[s];
      base[k * 4 + 0] = x_ctr - 0.5f * (new_w - 1.0f);
      base[k * 4 + 1] = y_ctr - 0.5f * (new_h - 1.0f);
      base[k * 4 + 2] = x_ctr + 0.5f * (new_w - 1.0f);
      base[k * 4 + 3] = y_ctr + 0.5f * (new_h - 1.0f);
      ++k;
    }
  for (int index = 0; index < H * W * A; ++index) {
    int a = index % A, w = (index / A) % W, h = index / A / W;
    out[index * 4 + 0] = (float)(base[a * 4 + 0] + w * feature_stride);
    out[index * 4 + 1] = (float)(base[a * 4 + 1] + h * feature_stride);
    out[index * 4 + 2] = (float)(base[a * 4 + 2] + w * feature_stride);
    out[index * 4 + 3] = (float)(base[a * 4 + 3] + h * feature_stride);
  }
  free(base);
}

/* ---- _contrib_GenProposal (generate_proposal.cu:289-430): legacy decode on supplied anchors, mask,
 * min-size filter, stable sort, no NMS; out (B, pre_param, 5).  Column 0 of padded rows is left
 * untouched by the reference; written as 0 here. ---- */
void oracle_gen_proposal(const float* cls_prob, const float* bbox_pred, const float* im_info,
                         const float* anchors, int B, int A, int H, int W, int feature_stride,
                         int pre_nms_top_n, int rpn_min_size, int iou_loss, float* out) {
  const int count = A * H * W;
  int pre = pre_nms_top_n > 0 ? pre_nms_top_n : count;
  if (pre > count) pre = count;
  float* prop = (float*)malloc(sizeof(float) * 5 * (size_t)count);
  float* score = (float*)malloc(sizeof(float) * (size_t)count);
  int* order = (int*)malloc(sizeof(int) * (size_t)count);
  for (int b = 0; b < B; ++b) {
    const float im_h = im_info[b * 3 + 0], im_w = im_info[b * 3 + 1], im_s = im_info[b * 3 + 2];
    const int real_h = (int)(im_h / feature_stride), real_w = (int)(im_w / feature_stride);
    const float* fg = cls_prob + (long)b * 2 * count + count;
    const float* dl = bbox_pred + (long)b * 4 * count;
    const float min_size = rpn_min_size * im_s;
    for (int index = 0; index < count; ++index) {
      int a = index % A, w = (index / A) % W, h = index / A / W;
      float bx1 = anchors[index * 4 + 0], by1 = anchors[index * 4 + 1];
      float bx2 = anchors[index * 4 + 2], by2 = anchors[index * 4 + 3];
      float sc = fg[(a * H + h) * W + w];
      float d0 = dl[((a * 4 + 0) * H + h) * W + w], d1 = dl[((a * 4 + 1) * H + h) * W + w];
      float d2 = dl[((a * 4 + 2) * H + h) * W + w], d3 = dl[((a * 4 + 3) * H + h) * W + w];
      float x1, y1, x2, y2;
      if (iou_loss) {
        x1 = bx1 + d0; y1 = by1 + d1; x2 = bx2 + d2; y2 = by2 + d3;
      } else {
        float width = bx2 - bx1 + 1.0f, height = by2 - by1 + 1.0f;
        float ctr_x = bx1 + 0.5f * (width - 1.0f), ctr_y = by1 + 0.5f * (height - 1.0f);
        float pcx = d0 * width + ctr_x, pcy = d1 * height + ctr_y;
        float pw = expf(d2) * width, ph = expf(d3) * height;
        x1 = pcx - 0.5f * (pw - 1.0f); y1 = pcy - 0.5f * (ph - 1.0f);
        x2 = pcx + 0.5f * (pw - 1.0f); y2 = pcy + 0.5f * (ph - 1.0f);
      }
      x1 = fmax_(fmin_(x1, im_w - 1.0f), 0.0f); y1 = fmax_(fmin_(y1, im_h - 1.0f), 0.0f);
      x2 = fmax_(fmin_(x2, im_w - 1.0f), 0.0f); y2 = fmax_(fmin_(y2, im_h - 1.0f), 0.0f);
      if (h >= real_h || w >= real_w) sc = -1.0f;
      float iw = x2 - x1 + 1.0f, ih = y2 - y1 + 1.0f;
      if (iw < min_size || ih < min_size) {
        x1 -= min_size / 2; y1 -= min_size / 2; x2 += min_size / 2; y2 += min_size / 2; sc = -1.0f;
      }
      float* p = prop + (long)index * 5;
      p[0] = x1; p[1] = y1; p[2] = x2; p[3] = y2; p[4] = sc;
      score[index] = sc;
    }
    oracle_stable_argsort_desc(score, count, order);
    float* o = out + (long)b * pre_nms_top_n * 5;
    for (int i = 0; i < pre_nms_top_n; ++i) {
      if (i < pre) memcpy(o + (long)i * 5, prop + (long)order[i] * 5, 20);
      else memset(o + (long)i * 5, 0, 20);
    }
  }
  free(prop); free(score); free(order);
}

/* ---- _contrib_GenProposalRetina (generate_proposal_retina.cu:307-469). ---- */
void oracle_gen_proposal_retina(const float* cls_prob, const float* bbox_pred, const float* im_info,
                                const float* anchors, int B, int AK, int H, int W, int num_anchors,
                                int pre_nms_top_n, int rpn_min_size, float thresh, const float* mean,
                                const float* stdv, int output_one_hot, float* out, float* out_score) {
  const int K = AK / num_anchors, count = AK * H * W;
  int pre = pre_nms_top_n > 0 ? pre_nms_top_n : count;
  if (pre > count) pre = count;
  const int oc = output_one_hot ? K + 1 : 1;
  float* prop = (float*)malloc(sizeof(float) * 5 * (size_t)count);
  float* score = (float*)malloc(sizeof(float) * (size_t)count);
  int* order = (int*)malloc(sizeof(int) * (size_t)count);
  memset(out, 0, sizeof(float) * 4 * (size_t)B * pre_nms_top_n);
  memset(out_score, 0, sizeof(float) * (size_t)oc * B * pre_nms_top_n);
  for (int b = 0; b < B; ++b) {
    const float im_h = im_info[b * 3 + 0], im_w = im_info[b * 3 + 1], im_s = im_info[b * 3 + 2];
    const float* sc = cls_prob + (long)b * count;
    const float* dl = bbox_pred + (long)b * 4 * count / K;
    const float min_size = rpn_min_size * im_s;
    for (int index = 0; index < count; ++index) {
      int a = index % AK, w = (index / AK) % W, h = index / AK / W;
      int ai = (h * W + w) * (AK / K) + a / K;
      float bx1 = anchors[ai * 4 + 0], by1 = anchors[ai * 4 + 1], bx2 = anchors[ai * 4 + 2], by2 = anchors[ai * 4 + 3];
      float s = sc[(a * H + h) * W + w];
      float width = bx2 - bx1 + 1.0f, height = by2 - by1 + 1.0f;
      float ctr_x = bx1 + 0.5f * (width - 1.0f), ctr_y = by1 + 0.5f * (height - 1.0f);
      float dx = dl[((a / K * 4) * H + h) * W + w] * stdv[0] + mean[0];
      float dy = dl[((a / K * 4 + 1) * H + h) * W + w] * stdv[1] + mean[1];
      float dw = dl[((a / K * 4 + 2) * H + h) * W + w] * stdv[2] + mean[2];
      float dh = dl[((a / K * 4 + 3) * H + h) * W + w] * stdv[3] + mean[3];
      float pcx = dx * width + ctr_x, pcy = dy * height + ctr_y;
      float pw = expf(dw) * width, ph = expf(dh) * height;
      float x1 = pcx - 0.5f * (pw - 1.0f), y1 = pcy - 0.5f * (ph - 1.0f);
      float x2 = pcx + 0.5f * (pw - 1.0f), y2 = pcy + 0.5f * (ph - 1.0f);
      x1 = fmax_(fmin_(x1, im_w - 1.0f), 0.0f); y1 = fmax_(fmin_(y1, im_h - 1.0f), 0.0f);
      x2 = fmax_(fmin_(x2, im_w - 1.0f), 0.0f); y2 = fmax_(fmin_(y2, im_h - 1.0f), 0.0f);
      float iw = x2 - x1 + 1.0f, ih = y2 - y1 + 1.0f;
      if (iw < min_size || ih < min_size || s <= thresh) { x1 = y1 = x2 = y2 = s = 0.0f; }
      float* p = prop + (long)index * 5;
      p[0] = x1; p[1] = y1; p[2] = x2; p[3] = y2; p[4] = s;
      score[index] = s;
    }
    oracle_stable_argsort_desc(score, count, order);
    for (int i = 0; i < pre && i < pre_nms_top_n; ++i) {
      const float* p = prop + (long)order[i] * 5;
      memcpy(out + ((long)b * pre_nms_top_n + i) * 4, p, 16);
      int cid = order[i] % K + 1;
      if (cid > oc - 1) cid = oc - 1;
      out_score[((long)b * pre_nms_top_n + i) * oc + cid] = p[4];
    }
  }
  free(prop); free(score); free(order);
}

/* ---- _contrib_NMS (nms.cu:274-364): proposals (B,count,5) -> out (B,post,4), score (B,post);
 * IoU > thr, zero padding (nms.cu:208-231). ---- */
void oracle_contrib_nms(const float* proposals, int B, int count, int pre_nms_top_n,
                        int post_nms_top_n, float threshold, int already_sorted, float* out,
                        float* out_score) {
  int pre = pre_nms_top_n > 0 ? pre_nms_top_n : count;
  if (pre > count) pre = count;
  int post = post_nms_top_n < pre ? post_nms_top_n : pre;
  float* score = (float*)malloc(sizeof(float) * (size_t)count);
  int* order = (int*)malloc(sizeof(int) * (size_t)count);
  float* dets = (float*)malloc(sizeof(float) * 5 * (size_t)pre);
  int* keep = (int*)calloc((size_t)pre, sizeof(int));
  for (int b = 0; b < B; ++b) {
    const float* p = proposals + (long)b * count * 5;
    for (int i = 0; i < count; ++i) { score[i] = p[(long)i * 5 + 4]; order[i] = i; }
    if (!already_sorted) oracle_stable_argsort_desc(score, count, order);
    for (int i = 0; i < pre; ++i) memcpy(dets + (long)i * 5, p + (long)order[i] * 5, 5 * sizeof(float));
    int nk = greedy_scan(dets, pre, threshold, /*ge=*/0, keep); /* nms.cu:140 uses > */
    /* NB: the op's declared output has param.rpn_post_nms_top_n rows (nms-inl.h:96-99) but only
     * `post` = min(post, pre) are written (nms.cu:354-358); rows beyond stay untouched. */
    for (int i = 0; i < post; ++i) {
      float* o = out + ((long)b * post_nms_top_n + i) * 4;
      if (i < nk) {
        memcpy(o, dets + (long)keep[i] * 5, 4 * sizeof(float));
        out_score[(long)b * post_nms_top_n + i] = dets[(long)keep[i] * 5 + 4];
      } else {
        o[0] = o[1] = o[2] = o[3] = 0.f;
        out_score[(long)b * post_nms_top_n + i] = 0.f;
      }
    }
  }
  free(score); free(order); free(dets); free(keep);
}

/* ---- bbox_overlaps_cython (bbox.pyx:32-73): boxes (N,4), query (K,4) -> (N,K).
 * Cython coerces the int literal in `x2 - x1 + 1` to the C constant `1.0` (a double), so the
 * generated C adds and multiplies those terms in DOUBLE and narrows on assignment to the
 * float32 variables; `float(...)` around `ua` is a C double cast.  Restated from the generated C
 * and pinned against the compiled reference module (tests/test_oracle_ref.py). ---- */
void oracle_bbox_overlaps(const float* boxes, int N, const float* query, int K, float* overlaps) {
  memset(overlaps, 0, sizeof(float) * (size_t)N * K);
  for (int k = 0; k < K; ++k) {
    const float* q = query + (long)k * 4;
    float box_area = (float)(((double)(q[2] - q[0]) + 1.0) * ((double)(q[3] - q[1]) + 1.0));
    for (int n = 0; n < N; ++n) {
      const float* b = boxes + (long)n * 4;
      float iw = (float)((double)(fmin_(b[2], q[2]) - fmax_(b[0], q[0])) + 1.0);
      if (iw > 0) {
        float ih = (float)((double)(fmin_(b[3], q[3]) - fmax_(b[1], q[1])) + 1.0);
        if (ih > 0) {
          float ua = (float)((((double)(b[2] - b[0]) + 1.0) * ((double)(b[3] - b[1]) + 1.0) +
                              (double)box_area) - (double)(iw * ih));
          overlaps[(long)n * K + k] = iw * ih / ua;
        }
      }
    }
  }
}


/* ---- bbox_selfoverlaps_cython (operator_py/cython/bbox_self.pyx:32-75): intersection over the area
 * of `boxes[n]` (IoA).  Same double promotions as bbox_overlaps above. ---- */
void oracle_bbox_selfoverlaps(const float* boxes, int N, const float* query, int K, float* overlaps) {
  memset(overlaps, 0, sizeof(float) * (size_t)N * K);
  for (int k = 0; k < K; ++k) {
    const float* q = query + (long)k * 4;
    for (int n = 0; n < N; ++n) {
      const float* b = boxes + (long)n * 4;
      float iw = (float)((double)(fmin_(b[2], q[2]) - fmax_(b[0], q[0])) + 1.0);
      if (iw > 0) {
        float ih = (float)((double)(fmin_(b[3], q[3]) - fmax_(b[1], q[1])) + 1.0);
        if (ih > 0) {
          float ub = (float)(((double)(b[2] - b[0]) + 1.0) * ((double)(b[3] - b[1]) + 1.0));
          overlaps[(long)n * K + k] = iw * ih / ub;
        }
      }
    }
  }
}

/* ---- greedy_nms (cpu_nms.pyx:37-87).  `order` = scores.argsort()[::-1] is supplied by the
 * caller (numpy's unstable sort decides ties); suppressed (ndets) is the output mask. ---- */
void oracle_greedy_nms(const float* dets5, int ndets, const long* order, float thresh,
                       unsigned char* suppressed) {
  memset(suppressed, 0, (size_t)ndets);
  for (int _i = 0; _i < ndets; ++_i) {
    long i = order[_i];
    if (suppressed[i]) continue;
    const float* a = dets5 + i * 5;
    float iarea = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
    for (int _j = _i + 1; _j < ndets; ++_j) {
      long j = order[_j];
      if (suppressed[j]) continue;
      const float* c = dets5 + j * 5;
      /* the pyx's own max/min: a if a >= b else b / a if a <= b else b */
      float xx1 = a[0] >= c[0] ? a[0] : c[0], yy1 = a[1] >= c[1] ? a[1] : c[1];
      float xx2 = a[2] <= c[2] ? a[2] : c[2], yy2 = a[3] <= c[3] ? a[3] : c[3];
      float w = 0.0f >= (xx2 - xx1 + 1) ? 0.0f : (xx2 - xx1 + 1);
      float h = 0.0f >= (yy2 - yy1 + 1) ? 0.0f : (yy2 - yy1 + 1);
      float inter = w * h;
      float ovr = inter / (iarea + (c[2] - c[0] + 1) * (c[3] - c[1] + 1) - inter);
      if (ovr >= thresh) suppressed[j] = 1;
    }
  }
}

/* ---- soft_nms (cpu_nms.pyx:98-203).  boxes (N,5) modified in place, inds (N) in/out.
 * Returns the new N.  method 0 hard, 1 linear, 2 gaussian. ---- */
int oracle_soft_nms(float* boxes, long* inds, int N, float sigma, float Nt, float threshold,
                    unsigned int method) {
  for (int i = 0; i < N; ++i) inds[i] = i;
  for (int i = 0; i < N; ++i) {
    float maxscore = boxes[i * 5 + 4];
    int maxpos = i;
    float t[5];
    memcpy(t, boxes + i * 5, sizeof(t));
    long ti = inds[i];
    for (int pos = i + 1; pos < N; ++pos)
      if (maxscore < boxes[pos * 5 + 4]) { maxscore = boxes[pos * 5 + 4]; maxpos = pos; }
    memcpy(boxes + i * 5, boxes + maxpos * 5, sizeof(t));
    inds[i] = inds[maxpos];
    memcpy(boxes + maxpos * 5, t, sizeof(t));
    inds[maxpos] = ti;
    const float tx1 = boxes[i * 5], ty1 = boxes[i * 5 + 1], tx2 = boxes[i * 5 + 2], ty2 = boxes[i * 5 + 3];
    int pos = i + 1;
    while (pos < N) {
      float* p = boxes + pos * 5;
      float x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
      /* `+ 1` is the C double constant 1.0 in the generated code (see oracle_bbox_overlaps) */
      float area = (float)(((double)(x2 - x1) + 1.0) * ((double)(y2 - y1) + 1.0));
      float iw = (float)((double)((tx2 <= x2 ? tx2 : x2) - (tx1 >= x1 ? tx1 : x1)) + 1.0);
      if (iw > 0) {
        float ih = (float)((double)((ty2 <= y2 ? ty2 : y2) - (ty1 >= y1 ? ty1 : y1)) + 1.0);
        if (ih > 0) {
          float ua = (float)((((double)(tx2 - tx1) + 1.0) * ((double)(ty2 - ty1) + 1.0) +
                              (double)area) - (double)(iw * ih));
          float ov = iw * ih / ua, weight;
          if (method == 1) weight = ov > Nt ? (float)(1.0 - (double)ov) : 1;
          else if (method == 2) weight = (float)exp((double)(-(ov * ov) / sigma)); /* np.exp on a Python float */
          else weight = ov > Nt ? 0 : 1;
          p[4] = weight * p[4];
          if (p[4] < threshold) {
            memcpy(p, boxes + (N - 1) * 5, 5 * sizeof(float));
            inds[pos] = inds[N - 1];
            N = N - 1;
            pos = pos - 1;
          }
        }
      }
      pos = pos + 1;
    }
  }
  return N;
}
