// Batched Soft-NMS on the device, sequentially-exact.
//
// Reference: operator_py/cython/cpu_nms.pyx:98-203 (`soft_nms`), called per class and image from
// detection_test.py:233-264 through operator_py/nms.py:5-16 (`cython_soft_nms_wrapper`).
// The reference algorithm is an in-place O(m^2) loop: select the max-score box among the
// remaining ones (first position wins ties), swap it to the front, re-weight every remaining box
// by its IoU with it, and delete boxes whose score fell below `threshold` by swapping them with
// the current last box.  One CTA per (image, class) problem keeps the boxes in shared memory and
// reproduces that exactly, including the final ORDER the swaps produce:
//   - argmax: block reduction on (score, -position);
//   - re-weighting: one thread per remaining box (independent);
//   - swap-with-last deletion of a whole pass == "fill the deleted slots below the new length, in
//     increasing order, with the surviving boxes above it, taken from the end": computed with two
//     ordered compactions instead of a serial walk.
// Arithmetic follows the Cython-GENERATED C (the `+ 1` terms are double adds/multiplies, see
// oracle/box_ops.c), so scores, order and indices are bit-identical to the compiled reference.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int block_scan_excl(int v, int* s_warp, int* total) {
  // exclusive prefix sum of v over the CTA (kThreads = 256 -> 8 warps); all threads call
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

__global__ void __launch_bounds__(kThreads)
soft_nms_kernel(const float* __restrict__ dets, const int* __restrict__ counts, const int m,
                const float sigma, const float Nt, const float threshold, const int method,
                float* __restrict__ out_boxes, int* __restrict__ out_inds, int* __restrict__ out_counts) {
  extern __shared__ float s_mem[];
  float* sb = s_mem;                                   // m x 5
  int* si = reinterpret_cast<int*>(sb + (size_t)m * 5);  // m original indices
  int* sflag = si + m;                                 // m: 1 = deleted in this pass
  int* slistA = sflag + m;                             // m
  int* slistB = slistA + m;                            // m
  __shared__ int s_warp[kThreads / 32];
  __shared__ float s_best[kThreads / 32];
  __shared__ int s_bpos[kThreads / 32];
  __shared__ float s_sel[5];
  const int p = blockIdx.x, tid = threadIdx.x;
  int N = counts ? counts[p] : m;
  const float* d = dets + (size_t)p * m * 5;
  for (int e = tid; e < N * 5; e += kThreads) sb[e] = d[e];
  for (int i = tid; i < N; i += kThreads) si[i] = i;
  __syncthreads();

  for (int i = 0; i < N; ++i) {
    // ---- 1. max score over [i, N), first position wins (`if maxscore < boxes[pos,4]`, :126-130)
    float bs = -FLT_MAX;
    int bp = 0x7fffffff;
    bool have = false;
    for (int q = i + tid; q < N; q += kThreads) {
      const float s = sb[q * 5 + 4];
      if (!have || bs < s) {  // a thread's positions ascend, so strict < keeps its first max
        bs = s;
        bp = q;
        have = true;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      const bool oh = op != 0x7fffffff;
      if (oh && (bp == 0x7fffffff || bs < os || (bs == os && op < bp))) {
        bs = os;
        bp = op;
      }
    }
    if ((tid & 31) == 0) {
      s_best[tid >> 5] = bs;
      s_bpos[tid >> 5] = bp;
    }
    __syncthreads();
    if (tid == 0) {
      float fs = s_best[0];
      int fp = s_bpos[0];
      for (int w = 1; w < kThreads / 32; ++w) {
        const float os = s_best[w];
        const int op = s_bpos[w];
        if (op != 0x7fffffff && (fp == 0x7fffffff || fs < os || (fs == os && op < fp))) {
          fs = os;
          fp = op;
        }
      }
      if (fp == 0x7fffffff) fp = i;
      // ---- 2. swap box i <-> box fp (:133-148)
      for (int k = 0; k < 5; ++k) {
        const float t = sb[i * 5 + k];
        sb[i * 5 + k] = sb[fp * 5 + k];
        sb[fp * 5 + k] = t;
        s_sel[k] = sb[i * 5 + k];
      }
      const int ti = si[i];
      si[i] = si[fp];
      si[fp] = ti;
    }
    __syncthreads();
    const float tx1 = s_sel[0], ty1 = s_sel[1], tx2 = s_sel[2], ty2 = s_sel[3];
    // ---- 3. re-weight every remaining box (:158-187); flag the ones that fall below threshold
    int ndel_local = 0;
    for (int q = i + 1 + tid; q < N; q += kThreads) {
      float* b = sb + q * 5;
      const float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
      int del = 0;
      const float area = (float)__dmul_rn(__dadd_rn((double)__fsub_rn(x2, x1), 1.0),
                                          __dadd_rn((double)__fsub_rn(y2, y1), 1.0));
      const float iw = (float)__dadd_rn((double)__fsub_rn(tx2 <= x2 ? tx2 : x2, tx1 >= x1 ? tx1 : x1), 1.0);
      if (iw > 0.f) {
        const float ih = (float)__dadd_rn((double)__fsub_rn(ty2 <= y2 ? ty2 : y2, ty1 >= y1 ? ty1 : y1), 1.0);
        if (ih > 0.f) {
          const double ta = __dmul_rn(__dadd_rn((double)__fsub_rn(tx2, tx1), 1.0),
                                      __dadd_rn((double)__fsub_rn(ty2, ty1), 1.0));
          const float inter = __fmul_rn(iw, ih);
          const float ua = (float)__dsub_rn(__dadd_rn(ta, (double)area), (double)inter);
          const float ov = __fdiv_rn(inter, ua);
          float weight;
          if (method == 1) weight = ov > Nt ? (float)__dsub_rn(1.0, (double)ov) : 1.f;
          else if (method == 2) weight = (float)exp((double)__fdiv_rn(-__fmul_rn(ov, ov), sigma));
          else weight = ov > Nt ? 0.f : 1.f;
          const float ns = __fmul_rn(weight, b[4]);
          b[4] = ns;
          del = ns < threshold;
        }
      }
      sflag[q] = del;
      ndel_local += del;
    }
    // ---- 4. swap-with-last deletion of the whole pass (:191-199)
    int ndel;
    (void)block_scan_excl(ndel_local, s_warp, &ndel);  // also a barrier: sflag / scores visible
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
        si[dst] = si[src];
      }
      N = Nn;
    }
    __syncthreads();
  }
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
  const size_t smem = (size_t)m * (5 * 4 + 4 * 4);
  if (smem > 200 * 1024) return sdet::fail(SDET_ERR_UNSUPPORTED, "soft-NMS over %d boxes per problem", m);
  if (smem > 48 * 1024)  // per device and cheap: set on every launch, no process-wide cache
    SDET_CUDA(cudaFuncSetAttribute(soft_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  soft_nms_kernel<<<(unsigned)problems, kThreads, smem, (cudaStream_t)stream>>>(
      dets, counts, m, sigma, Nt, threshold, method, out_boxes, out_inds, out_counts);
  SDET_LAUNCH_CHECK("soft_nms_kernel");
  return SDET_OK;
}
