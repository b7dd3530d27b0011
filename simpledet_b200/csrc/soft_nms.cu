// Batched Soft-NMS on the device, sequentially-exact.
//
// Reference: operator_py/cython/cpu_nms.pyx:98-203 (`soft_nms`), called per class and image from
// detection_test.py:233-264 through operator_py/nms.py:5-16 (`cython_soft_nms_wrapper`).
// The reference algorithm is an in-place O(m^2) loop: select the max-score box among the
// remaining ones (first position wins ties), swap it to the front, re-weight every remaining box
// by its IoU with it, and delete boxes whose score fell below `threshold` by swapping them with
// the current last box.  One CTA per (image, class) problem keeps the boxes in shared memory and
// reproduces that exactly, including the final ORDER the swaps produce:
//   - argmax: block reduction on (score, -position), folded into the re-weighting pass of the previous
//     iteration; the winner's box travels with the reduction, so an iteration costs ONE barrier
//     (round 1: five or more; profiles/r02_launches_dcn_softnms_summary.txt);
//   - re-weighting: one thread per remaining box (independent);
//   - swap-with-last deletion of a whole pass == "fill the deleted slots below the new length, in
//     increasing order, with the surviving boxes above it, taken from the end": computed with two
//     ordered compactions instead of a serial walk.
// Arithmetic follows the Cython-GENERATED C (the `+ 1` terms are double adds/multiplies, see
// oracle/box_ops.c), so scores, order and indices are bit-identical to the compiled reference.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int kThreads = 512;

__device__ __forceinline__ int block_scan_excl(int v, int* s_warp, int* total) {
  // exclusive prefix sum of v over the CTA; all threads call
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
  for (int w = 0; w < kThreads / 32; ++w) {
    const int c = s_warp[w];
    if (w < warp) base += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

// A box travelling through the arg-max reduction: the winner's values reach every thread together with its
// position, so nobody has to read the winner's slot after the barrier (which is what lets one barrier per
// iteration suffice: the slot is rewritten by the swap right away).
struct Cand {
  float s;    // score
  int pos;    // position in the working arrays; 0x7fffffff: none
  float x1, y1, x2, y2, area;
  int idx;    // original index
};

struct __align__(16) Partial {
  Cand c;
  int ndel;
  int pad;
};
static_assert(sizeof(Partial) == 48, "Partial is read back as three 16-byte words");

__device__ __forceinline__ uint32_t score_bits(const float s) {  // larger float -> larger uint
  const uint32_t u = __float_as_uint(s + 0.f);  // -0.0 and +0.0 compare equal, like `maxscore < score` does
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Block-wide: best candidate over all threads (larger score, ties to the lower position) and the sum of ndel.
// Warp level and block level are each two REDUX instructions (max of the sortable score bits, then min of the
// positions that carry it) instead of a five-level shuffle tree; the winner's lane publishes its whole box.
// One barrier; `buf` must alternate between two buffers from call to call (a fast warp may already fill the next one
// while a slow one still reads this one).
__device__ __forceinline__ Cand reduce_best(const Cand c, const int ndel, Partial* buf, int* ndel_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool have = c.pos != 0x7fffffff;
  const uint32_t mine = have ? score_bits(c.s) : 0u;  // score_bits of a real score is never 0 (NaN-free inputs aside, 0 needs u = 0xFFFFFFFF)
  const uint32_t top = __reduce_max_sync(0xffffffffu, mine);
  const int pos = (int)__reduce_min_sync(0xffffffffu, (have && mine == top) ? (uint32_t)c.pos : 0x7fffffffu);
  const int nd_w = __reduce_add_sync(0xffffffffu, ndel);
  if (pos != 0x7fffffff ? (c.pos == pos) : (lane == 0)) {
    Partial p;
    p.c = c;
    p.c.pos = pos;
    p.ndel = nd_w;
    p.pad = 0;
    buf[warp] = p;
  }
  __syncthreads();
  // second level: lane w looks at warp w's partial
  constexpr int kW = kThreads / 32;
  float ws = -FLT_MAX;
  int wp = 0x7fffffff, wn = 0;
  if (lane < kW) {
    ws = buf[lane].c.s;
    wp = buf[lane].c.pos;
    wn = buf[lane].ndel;
  }
  const bool whave = wp != 0x7fffffff;
  const uint32_t wmine = whave ? score_bits(ws) : 0u;
  const uint32_t wtop = __reduce_max_sync(0xffffffffu, wmine);
  const int bpos = (int)__reduce_min_sync(0xffffffffu, (whave && wmine == wtop) ? (uint32_t)wp : 0x7fffffffu);
  *ndel_total = __reduce_add_sync(0xffffffffu, wn);
  const unsigned who = __ballot_sync(0xffffffffu, whave && wp == bpos);
  Cand best;
  if (who) {
    best = buf[__ffs(who) - 1].c;
  } else {
    best.s = -FLT_MAX; best.pos = 0x7fffffff; best.x1 = best.y1 = best.x2 = best.y2 = best.area = 0.f; best.idx = 0;
  }
  return best;
}

__global__ void __launch_bounds__(kThreads)
soft_nms_kernel(const float* __restrict__ dets, const int* __restrict__ counts, const int m,
                const float sigma, const float Nt, const float threshold, const int method,
                float* __restrict__ out_boxes, int* __restrict__ out_inds, int* __restrict__ out_counts) {
  extern __shared__ float s_mem[];
  float* sb = s_mem;                                   // m x 5
  float* sa = sb + (size_t)m * 5;                      // m areas (float of the double product, as the reference forms it)
  int* si = reinterpret_cast<int*>(sa + m);            // m original indices
  int* sflag = si + m;                                 // m: 1 = deleted in this pass
  int* slistA = sflag + m;                             // m
  int* slistB = slistA + m;                            // m
  __shared__ int s_warp[kThreads / 32];
  __shared__ Partial s_part[2][kThreads / 32];
  const int p = blockIdx.x, tid = threadIdx.x;
  int N = counts ? counts[p] : m;
  const float* d = dets + (size_t)p * m * 5;
  for (int e = tid; e < N * 5; e += kThreads) sb[e] = d[e];
  for (int i = tid; i < N; i += kThreads) si[i] = i;
  __syncthreads();
  for (int i = tid; i < N; i += kThreads) {
    const float* b = sb + i * 5;
    sa[i] = (float)__dmul_rn(__dadd_rn((double)__fsub_rn(b[2], b[0]), 1.0), __dadd_rn((double)__fsub_rn(b[3], b[1]), 1.0));
  }
  __syncthreads();

  // arg-max over [from, N) by a plain scan (start, and after a pass that deleted boxes: positions moved)
  auto scan_best = [&](const int from, Partial* buf) {
    Cand c;
    c.s = -FLT_MAX; c.pos = 0x7fffffff; c.x1 = c.y1 = c.x2 = c.y2 = c.area = 0.f; c.idx = 0;
    for (int q = from + tid; q < N; q += kThreads) {
      const float s = sb[q * 5 + 4];
      if (c.pos == 0x7fffffff || c.s < s) { c.s = s; c.pos = q; }  // a thread's positions ascend: strict < keeps its first
    }
    if (c.pos != 0x7fffffff) {
      const float* b = sb + c.pos * 5;
      c.x1 = b[0]; c.y1 = b[1]; c.x2 = b[2]; c.y2 = b[3]; c.area = sa[c.pos]; c.idx = si[c.pos];
    }
    int nd;
    return reduce_best(c, 0, buf, &nd);
  };

  int par = 0;
  Cand w = scan_best(0, s_part[par]);
  par ^= 1;
  for (int i = 0; i < N; ++i) {
    // ---- 1+2. the winner w (max score over [i, N), first position) goes to slot i, box i to the winner's slot
    // (:126-148).  Nobody reads slot w.pos any more (the values travelled with the reduction), so the thread whose
    // stride visits q == w.pos does the exchange itself, without a barrier.
    const int fp = w.pos == 0x7fffffff ? i : w.pos;
    const float tx1 = w.x1, ty1 = w.y1, tx2 = w.x2, ty2 = w.y2;
    const double ta = __dmul_rn(__dadd_rn((double)__fsub_rn(tx2, tx1), 1.0), __dadd_rn((double)__fsub_rn(ty2, ty1), 1.0));
    // ---- 3. re-weight every remaining box (:158-187); flag the ones that fall below threshold; remember the best
    Cand c;
    c.s = -FLT_MAX; c.pos = 0x7fffffff; c.x1 = c.y1 = c.x2 = c.y2 = c.area = 0.f; c.idx = 0;
    int ndel_local = 0;
    for (int q = i + 1 + tid; q < N; q += kThreads) {
      float x1, y1, x2, y2, sc, area;
      int idx;
      if (q == fp) {  // the exchange: old box i is the one that lives at q from now on
        x1 = sb[i * 5]; y1 = sb[i * 5 + 1]; x2 = sb[i * 5 + 2]; y2 = sb[i * 5 + 3]; sc = sb[i * 5 + 4];
        area = sa[i]; idx = si[i];
        sb[i * 5] = tx1; sb[i * 5 + 1] = ty1; sb[i * 5 + 2] = tx2; sb[i * 5 + 3] = ty2; sb[i * 5 + 4] = w.s;
        sa[i] = w.area; si[i] = w.idx;
        float* b = sb + q * 5;
        b[0] = x1; b[1] = y1; b[2] = x2; b[3] = y2;
        sa[q] = area; si[q] = idx;
      } else {
        const float* b = sb + q * 5;
        x1 = b[0]; y1 = b[1]; x2 = b[2]; y2 = b[3]; sc = b[4];
        area = sa[q]; idx = si[q];
      }
      int del = 0;
      float ns = sc;
      const float iw = (float)__dadd_rn((double)__fsub_rn(tx2 <= x2 ? tx2 : x2, tx1 >= x1 ? tx1 : x1), 1.0);
      if (iw > 0.f) {
        const float ih = (float)__dadd_rn((double)__fsub_rn(ty2 <= y2 ? ty2 : y2, ty1 >= y1 ? ty1 : y1), 1.0);
        if (ih > 0.f) {
          const float inter = __fmul_rn(iw, ih);
          const float ua = (float)__dsub_rn(__dadd_rn(ta, (double)area), (double)inter);
          const float ov = __fdiv_rn(inter, ua);
          float weight;
          if (method == 1) weight = ov > Nt ? (float)__dsub_rn(1.0, (double)ov) : 1.f;
          else if (method == 2) weight = (float)exp((double)__fdiv_rn(-__fmul_rn(ov, ov), sigma));
          else weight = ov > Nt ? 0.f : 1.f;
          ns = __fmul_rn(weight, sc);
          del = ns < threshold;
        }
      }
      sb[q * 5 + 4] = ns;
      sflag[q] = del;
      ndel_local += del;
      if (c.pos == 0x7fffffff || c.s < ns) {  // next iteration's arg-max, for free (valid when nothing is deleted)
        c.s = ns; c.pos = q; c.x1 = x1; c.y1 = y1; c.x2 = x2; c.y2 = y2; c.area = area; c.idx = idx;
      }
    }
    int ndel;
    w = reduce_best(c, ndel_local, s_part[par], &ndel);  // the iteration's one barrier
    par ^= 1;
    // ---- 4. swap-with-last deletion of the whole pass (:191-199)
    if (ndel > 0) {
      const int Nn = N - ndel;
      // A: deleted slots below Nn, ascending.  B: survivors at >= Nn, ascending (used from the end).
      int na = 0, nb = 0;
      for (int base = i + 1; base < N; base += kThreads) {
        const int q = base + tid;
        const int fa = (q < Nn && q < N) ? sflag[q] : 0;
        const int fb = (q >= Nn && q < N) ? !sflag[q] : 0;
        int ta_, tb_;
        const int oa = block_scan_excl(fa, s_warp, &ta_);
        const int ob = block_scan_excl(fb, s_warp, &tb_);
        if (fa) slistA[na + oa] = q;
        if (fb) slistB[nb + ob] = q;
        na += ta_;
        nb += tb_;
      }
      __syncthreads();
      // the j-th deleted slot (ascending) receives the j-th survivor from the END
      for (int j = tid; j < na; j += kThreads) {
        const int dst = slistA[j], src = slistB[nb - 1 - j];
        for (int k = 0; k < 5; ++k) sb[dst * 5 + k] = sb[src * 5 + k];
        sa[dst] = sa[src];
        si[dst] = si[src];
      }
      N = Nn;
      __syncthreads();
      w = scan_best(i + 1, s_part[par]);  // positions moved: find the next winner by a scan
      par ^= 1;
    }
  }
  __syncthreads();
  for (int e = tid; e < m * 5; e += kThreads) out_boxes[(size_t)p * m * 5 + e] = e < N * 5 ? sb[e] : 0.f;
  for (int i = tid; i < m; i += kThreads) out_inds[(size_t)p * m + i] = i < N ? si[i] : -1;
  if (tid == 0) out_counts[p] = N;
}

}  // namespace

extern "C" int sdet_soft_nms(const float* dets, const int* counts, int problems, int m, float sigma,
                             float Nt, float threshold, int method, float* out_boxes, int* out_inds,
                             int* out_counts, void* stream) {
  SDET_REQUIRE(dets && out_boxes && out_inds && out_counts, "NULL argument");
  SDET_REQUIRE(problems > 0 && m > 0, "problems and m must be > 0");
  SDET_REQUIRE(method >= 0 && method <= 2, "method must be 0 (hard), 1 (linear) or 2 (gaussian)");
  const size_t smem = (size_t)m * (5 * 4 + 5 * 4);
  if (smem > 200 * 1024) return sdet::fail(SDET_ERR_UNSUPPORTED, "soft-NMS over %d boxes per problem", m);
  if (smem > 48 * 1024)  // per device and cheap: set on every launch, no process-wide cache
    SDET_CUDA(cudaFuncSetAttribute(soft_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  soft_nms_kernel<<<(unsigned)problems, kThreads, smem, (cudaStream_t)stream>>>(
      dets, counts, m, sigma, Nt, threshold, method, out_boxes, out_inds, out_counts);
  SDET_LAUNCH_CHECK("soft_nms_kernel");
  return SDET_OK;
}
