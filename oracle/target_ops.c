/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * CPU restatement of ProposalTarget (operator_cxx/proposal_target-inl.h:123-256 op,
 * operator_cxx/proposal_target.cc:22-227 SampleROI / BBoxOverlap / ExpandBboxRegressionTargets /
 * NonLinearTransformAndNormalization).  No fixtures in the reference; pinned bit for bit against the
 * reference's own proposal_target.cc / proposal_target_v2.cc compiled through oracle/shim with a controlled rand()
 * (tests/test_oracle_ref_cxx.py, vectors in tests/golden/reference_cxx_ops.npz).
 *
 * Randomness.  The reference calls std::random_shuffle on the global rand() state shared by all
 * device threads (proposal_target.cc:83,102,118) — not reproducible even against itself.  What IS
 * defined: "a uniformly random order of the candidate list, truncated".  Both this oracle and the
 * CUDA op realise a shuffle as "sort the candidates by an injected 32-bit priority (ties by the
 * candidate's index)", with one priority array per draw: draw 0 = fg shuffle, 1 = bg shuffle,
 * 2+r = r-th shuffle of the negative padding loop (cycled modulo the draws supplied).  With the
 * same priorities injected on both sides every output is comparable bit for bit; the CUDA op's own
 * Philox priorities are checked through invariants.
 *
 * Defined where the reference has UB: no valid gt box -> every overlap is 0 (the reference reads
 * IOUs[i][0] of a 0-column matrix); fewer kept rois than image_rois with an empty negative list ->
 * remaining rows stay zero (the reference reads kept_indexes out of range, :142-145).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float fminf__(float a, float b) { return a < b ? a : b; }
static inline float fmaxf__(float a, float b) { return a < b ? b : a; }

typedef struct { uint32_t pr; int idx; } cand_t;
static int cmp_cand(const void* a, const void* b) {
  const cand_t *x = (const cand_t*)a, *y = (const cand_t*)b;
  if (x->pr != y->pr) return x->pr < y->pr ? -1 : 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}
/* "random_shuffle(list)" := order by (priority[idx], idx) */
static void shuffle_by_priority(int* list, int n, const uint32_t* prio) {
  cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) { c[i].pr = prio[list[i]]; c[i].idx = list[i]; }
  qsort(c, (size_t)n, sizeof(cand_t), cmp_cand);
  for (int i = 0; i < n; ++i) list[i] = c[i].idx;
  free(c);
}

/* rois (B,R,4), gt_boxes (B,G,5).  priorities (B, D, R+G) uint32, D >= 3 draws.
 * Outputs: rois_out (B,IR,4), labels (B,IR), bbox_targets/bbox_weights (B,IR,NC*4),
 * match_gt_ious (B,IR); dbg_kept (B,IR) int: index into the image's compacted roi list or -1. */
void oracle_proposal_target(const float* rois, const float* gt_boxes, int B, int R, int G,
                            int num_classes, int image_rois, float fg_fraction, float fg_thresh,
                            float bg_thresh_hi, float bg_thresh_lo, int proposal_without_gt,
                            int class_agnostic, const float* bbox_mean, const float* bbox_std,
                            const float* bbox_weight, const uint32_t* priorities, int D,
                            float* rois_out, float* labels, float* bbox_targets, float* bbox_weights,
                            float* match_gt_ious, int* dbg_kept, int* gt_index, int* fg_count,
                            const float* valid_ranges, int filter_scales, int no_fg_cap) {
  /* ProposalTarget_v2 deltas (proposal_target_v2-inl.h:145-270, proposal_target_v2.cc:80-85):
   * valid_ranges (B,2) + filter_scales: gt boxes whose area lies outside [min^2, max^2] are not
   * APPENDED to the rois (they still take part in the IoU); no_fg_cap (image_rois == -1): keep
   * every foreground roi, rows per image = R. */
  const int NC4 = num_classes * 4, T = R + G;
  memset(rois_out, 0, sizeof(float) * (size_t)B * image_rois * 4);
  memset(labels, 0, sizeof(float) * (size_t)B * image_rois);
  memset(bbox_targets, 0, sizeof(float) * (size_t)B * image_rois * NC4);
  memset(bbox_weights, 0, sizeof(float) * (size_t)B * image_rois * NC4);
  memset(match_gt_ious, 0, sizeof(float) * (size_t)B * image_rois);
  const int fg_rois_per_image = (int)(image_rois * fg_fraction); /* -inl.h:194 truncation */
  float* all = (float*)malloc(sizeof(float) * 4 * (size_t)T);
  float* gts = (float*)malloc(sizeof(float) * 5 * (size_t)(G > 0 ? G : 1));
  int* gsrc = (int*)malloc(sizeof(int) * (size_t)(G > 0 ? G : 1));
  float* maxov = (float*)malloc(sizeof(float) * (size_t)T);
  float* lab = (float*)malloc(sizeof(float) * (size_t)T);
  int* assign = (int*)malloc(sizeof(int) * (size_t)T);
  int *fg = (int*)malloc(sizeof(int) * (size_t)T), *bg = (int*)malloc(sizeof(int) * (size_t)T);
  int *neg = (int*)malloc(sizeof(int) * (size_t)T), *kept = (int*)malloc(sizeof(int) * (size_t)(image_rois + T));
  for (int b = 0; b < B; ++b) {
    const uint32_t* prio = priorities + (size_t)b * D * T;
    int ng = 0, n = 0;
    for (int j = 0; j < G; ++j) /* -inl.h:155-161: padding gt has cls == -1 */
      if (gt_boxes[((size_t)b * G + j) * 5 + 4] != -1.f) { gsrc[ng] = j; memcpy(gts + 5 * ng++, gt_boxes + ((size_t)b * G + j) * 5, 20); }
    for (int j = 0; j < R; ++j) /* :171-176: y2 == 0 indicates padding */
      if (rois[((size_t)b * R + j) * 4 + 3] > 0) memcpy(all + 4 * n++, rois + ((size_t)b * R + j) * 4, 16);
    if (!proposal_without_gt) /* :177-185: all valid gt boxes appended after the rois */
      for (int j = 0; j < ng; ++j) {
        if (filter_scales && valid_ranges) { /* v2-inl.h:191-200 */
          float vmin = valid_ranges[b * 2] * valid_ranges[b * 2], vmax = valid_ranges[b * 2 + 1] * valid_ranges[b * 2 + 1];
          float gw = (float)((double)(gts[5 * j + 2] - gts[5 * j + 0]) + 1.0);
          float gh = (float)((double)(gts[5 * j + 3] - gts[5 * j + 1]) + 1.0);
          if (gw * gh < vmin || gw * gh > vmax) continue;
        }
        memcpy(all + 4 * n++, gts + 5 * j, 16);
      }
    /* BBoxOverlap (proposal_target.cc:165-185) + row argmax with strict '<' (:51-63) */
    for (int i = 0; i < n; ++i) {
      const float* bx = all + 4 * i;
      float best = 0.f; int bi = 0;
      for (int j = 0; j < ng; ++j) {
        const float* q = gts + 5 * j;
        float qa = (q[2] - q[0] + 1.f) * (q[3] - q[1] + 1.f), ov = 0.f;
        float iw = fminf__(bx[2], q[2]) - fmaxf__(bx[0], q[0]) + 1.f;
        if (iw > 0) {
          float ih = fminf__(bx[3], q[3]) - fmaxf__(bx[1], q[1]) + 1.f;
          if (ih > 0) {
            float ba = (bx[2] - bx[0] + 1.f) * (bx[3] - bx[1] + 1.f);
            ov = iw * ih / (ba + qa - iw * ih);
          }
        }
        if (j == 0) { best = ov; bi = 0; } else if (best < ov) { best = ov; bi = j; }
      }
      maxov[i] = best; assign[i] = bi; lab[i] = ng > 0 ? gts[5 * bi + 4] : 0.f;
    }
    int nfg = 0, nbg = 0, nneg = 0, nk = 0;
    for (int i = 0; i < n; ++i) { if (maxov[i] >= fg_thresh) fg[nfg++] = i; else neg[nneg++] = i; }
    int fg_n = no_fg_cap ? nfg : (fg_rois_per_image < nfg ? fg_rois_per_image : nfg);
    if (nfg > fg_n) shuffle_by_priority(fg, nfg, prio + 0 * (size_t)T); /* :81-85 */
    for (int i = 0; i < n; ++i) if (maxov[i] >= bg_thresh_lo && maxov[i] < bg_thresh_hi) bg[nbg++] = i;
    int bg_n = (image_rois - fg_n) < nbg ? (image_rois - fg_n) : nbg;
    if (nbg > bg_n) shuffle_by_priority(bg, nbg, prio + 1 * (size_t)T); /* :100-104 */
    for (int i = 0; i < fg_n; ++i) kept[nk++] = fg[i];
    for (int i = 0; i < bg_n; ++i) kept[nk++] = bg[i];
    for (int r = 0; nk < image_rois && nneg > 0; ++r) { /* :116-122 */
      int gap = image_rois - nk;
      shuffle_by_priority(neg, nneg, prio + (size_t)(2 + r % (D - 2)) * T);
      for (int i = 0; i < gap && i < nneg; ++i) kept[nk++] = neg[i];
    }
    if (fg_count) fg_count[b] = fg_n;
    for (int i = 0; i < image_rois; ++i) {
      size_t row = (size_t)b * image_rois + i;
      if (dbg_kept) dbg_kept[row] = i < nk ? kept[i] : -1;
      if (gt_index) gt_index[row] = (i < nk && ng > 0) ? gsrc[assign[kept[i]]] : -1;
      if (i >= nk) continue;
      const int k = kept[i];
      float label = i < fg_n ? lab[k] : 0.f; /* :128-131 */
      labels[row] = label;
      memcpy(rois_out + row * 4, all + 4 * k, 16);
      match_gt_ious[row] = maxov[k];
      if (ng == 0) continue;
      /* NonLinearTransformAndNormalization (:204-227); `0.5 *` is a double literal */
      const float* ex = all + 4 * k; const float* gt = gts + 5 * assign[k];
      float ew = ex[2] - ex[0] + 1.f, eh = ex[3] - ex[1] + 1.f;
      float ecx = (float)(ex[0] + 0.5 * (ew - 1.f)), ecy = (float)(ex[1] + 0.5 * (eh - 1.f));
      float gw = gt[2] - gt[0] + 1.f, gh = gt[3] - gt[1] + 1.f;
      float gcx = (float)(gt[0] + 0.5 * (gw - 1.f)), gcy = (float)(gt[1] + 0.5 * (gh - 1.f));
      float t[4] = {(gcx - ecx) / (ew + 1e-14f), (gcy - ecy) / (eh + 1e-14f), logf(gw / ew), logf(gh / eh)};
      for (int c = 0; c < 4; ++c) { t[c] -= bbox_mean[c]; t[c] /= bbox_std[c]; }
      float cls = class_agnostic ? (label < 1.f ? label : 1.f) : label; /* :151-157 */
      if (cls > 0) { /* ExpandBboxRegressionTargets :187-202 */
        int start = 4 * (int)cls;
        memcpy(bbox_targets + row * NC4 + start, t, 16);
        memcpy(bbox_weights + row * NC4 + start, bbox_weight, 16);
      }
    }
  }
  free(all); free(gts); free(gsrc); free(maxov); free(lab); free(assign); free(fg); free(bg); free(neg); free(kept);
}
