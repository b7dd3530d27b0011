// Box decode / proposal generation / NMS for sm_100a, everything on the device (the reference
// copies to the host for DecodeBBox and for the greedy NMS scan — SURVEY.md §0.3).
//
//   sdet_decode_bbox       _contrib_DecodeBBox   operator_cxx/contrib/decodebbox.cc:34-133
//   sdet_proposal_v3       _contrib_Proposal_v3  operator_cxx/contrib/proposal_v3.cu:435-638
//   sdet_contrib_nms       _contrib_NMS          operator_cxx/contrib/nms.cu:274-364
//   sdet_nms_sorted        batched greedy NMS over pre-sorted boxes (building block; also the
//                          device side of the `_nms` compatibility export)
//
// Pipeline of a Proposal call (all (image) problems of the batch in each launch):
//   1. proposal_topk_kernel  — one CTA per image: radix-select the top `pre` fg scores with the
//      reference's stable order (topk.cuh), then decode + clip + min-size-filter ONLY those.
//   2. nms_mask_kernel       — 64x64 IoU tiles -> suppression bitmask, upper triangle only.
//   3. nms_scan_kernel       — one CTA per image walks the bitmask in 64-row blocks (the serial
//      greedy dependency) and writes the padded outputs.  No host round trip, no sync.
//
// Float ops that feed an index decision (IoU vs threshold, min-size filter) are explicit
// round-to-nearest intrinsics in the reference's source order, so keep/suppress decisions are
// bit-identical to the oracle's for identical inputs.
#include <algorithm>
#include <cfloat>
#include <climits>

#include "common.cuh"
#include "topk.cuh"

namespace {

using sdet::kTopkThreads;

__device__ __forceinline__ float fmin_ref(float a, float b) { return a < b ? a : b; }  // CUDA min()
__device__ __forceinline__ float fmax_ref(float a, float b) { return a < b ? b : a; }  // CUDA max()

// --------------------------------------------------------------------------------------------
// DecodeBBox: thread per (roi, class); float4 in / out.
// --------------------------------------------------------------------------------------------
struct DecodeParams {
  float mean[4], std[4];
  int class_agnostic, xyxy;
};

__global__ void __launch_bounds__(256)
decode_bbox_kernel(const float4* __restrict__ rois, const float4* __restrict__ deltas,
                   const float* __restrict__ im_info, float4* __restrict__ out, const int N,
                   const int K, const int total, const DecodeParams p) {
  const int ncls = p.class_agnostic ? 1 : K;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int cls = t % ncls;
    const int ri = t / ncls;  // n*N + i
    const int n = ri / N;
    const float4 b = __ldg(rois + ri);
    const float4 d = __ldg(deltas + (size_t)ri * K + (p.class_agnostic ? 1 : cls));
    const float im_h = __ldg(im_info + n * 3), im_w = __ldg(im_info + n * 3 + 1);
    const float width = __fadd_rn(__fsub_rn(b.z, b.x), 1.0f);
    const float height = __fadd_rn(__fsub_rn(b.w, b.y), 1.0f);
    const float dx = __fadd_rn(__fmul_rn(d.x, p.std[0]), p.mean[0]);
    const float dy = __fadd_rn(__fmul_rn(d.y, p.std[1]), p.mean[1]);
    const float dw = __fadd_rn(__fmul_rn(d.z, p.std[2]), p.mean[2]);
    const float dh = __fadd_rn(__fmul_rn(d.w, p.std[3]), p.mean[3]);
    float x1, y1, x2, y2;
    if (!p.xyxy) {
      const float ctr_x = __fadd_rn(b.x, __fmul_rn(0.5f, __fsub_rn(width, 1.0f)));
      const float ctr_y = __fadd_rn(b.y, __fmul_rn(0.5f, __fsub_rn(height, 1.0f)));
      const float pcx = __fadd_rn(__fmul_rn(dx, width), ctr_x);
      const float pcy = __fadd_rn(__fmul_rn(dy, height), ctr_y);
      // decodebbox.cc:62-63: `exp(dw) * width` binds to ::exp(double) on the host - double exp, double product,
      // one narrowing (pinned by the compiled reference).  No exp clip in DecodeBBox (Appendix A.12).
      const float pw = (float)__dmul_rn(exp((double)dw), (double)width);
      const float ph = (float)__dmul_rn(exp((double)dh), (double)height);
      const float hw = __fmul_rn(0.5f, __fsub_rn(pw, 1.0f)), hh = __fmul_rn(0.5f, __fsub_rn(ph, 1.0f));
      x1 = __fsub_rn(pcx, hw);
      y1 = __fsub_rn(pcy, hh);
      x2 = __fadd_rn(pcx, hw);
      y2 = __fadd_rn(pcy, hh);
    } else {
      x1 = __fadd_rn(b.x, __fmul_rn(dx, width));
      y1 = __fadd_rn(b.y, __fmul_rn(dy, height));
      x2 = __fadd_rn(b.z, __fmul_rn(dw, width));
      y2 = __fadd_rn(b.w, __fmul_rn(dh, height));
    }
    const float mx = __fsub_rn(im_w, 1.0f), my = __fsub_rn(im_h, 1.0f);
    float4 o;
    o.x = fmax_ref(fmin_ref(x1, mx), 0.0f);
    o.y = fmax_ref(fmin_ref(y1, my), 0.0f);
    o.z = fmax_ref(fmin_ref(x2, mx), 0.0f);
    o.w = fmax_ref(fmin_ref(y2, my), 0.0f);
    out[(size_t)ri * ncls + cls] = o;
  }
}

// --------------------------------------------------------------------------------------------
// NMS building blocks.  dets: (P, n, 5) rows [x1,y1,x2,y2,score] already in greedy order.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float dev_iou(const float* a, const float* b) {
  // devIoU, proposal_v3.cu:271-279
  const float left = fmax_ref(a[0], b[0]), right = fmin_ref(a[2], b[2]);
  const float top = fmax_ref(a[1], b[1]), bottom = fmin_ref(a[3], b[3]);
  const float width = fmax_ref(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  const float height = fmax_ref(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float inter = __fmul_rn(width, height);
  const float Sa = __fmul_rn(__fadd_rn(__fsub_rn(a[2], a[0]), 1.f), __fadd_rn(__fsub_rn(a[3], a[1]), 1.f));
  const float Sb = __fmul_rn(__fadd_rn(__fsub_rn(b[2], b[0]), 1.f), __fadd_rn(__fsub_rn(b[3], b[1]), 1.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(Sa, Sb), inter));
}

// grid = (ceil(col_blocks / tpc), row_blocks, P); 64 threads = the 64 rows of a row block; a CTA walks
// `tpc` column tiles (only col >= row tiles do work).  tpc = 1 spreads a few large problems over the
// SMs; many small problems (per-class NMS: 160 x 136 tiles) take tpc = col_blocks so that the launch
// is not 20 000 64-thread CTAs of which most exit at once.
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float* __restrict__ dets, const int* __restrict__ counts, const int n_max,
                const float thr, const int ge, unsigned long long* __restrict__ mask,
                const float* __restrict__ sets,  // (P,n_max) or nullptr: set_nms (nms.py:77-107)
                const int tpc) {
  const int row_b = blockIdx.y, p = blockIdx.z;
  const int n = counts ? counts[p] : n_max;
  const int col_blocks = (n_max + 63) >> 6, nb = (n + 63) >> 6;
  const int c_beg = max((int)blockIdx.x * tpc, row_b), c_end = min((int)(blockIdx.x + 1) * tpc, nb);
  if (row_b >= nb || c_beg >= c_end) return;  // the scan only reads words j >= row block (proposal_v3.cu:373)
  const int row_size = min(n - row_b * 64, 64);
  const float* d = dets + (size_t)p * n_max * 5;
  __shared__ float sb[64 * 5];
  __shared__ float s_set[64];
  const int t = threadIdx.x;
  const int cur = row_b * 64 + t;
  float cb[4] = {0.f, 0.f, 0.f, 0.f};
  float my_set = 0.f;
  if (t < row_size) {
#pragma unroll
    for (int k = 0; k < 4; ++k) cb[k] = d[(size_t)cur * 5 + k];
    if (sets) my_set = sets[(size_t)p * n_max + cur];
  }
  // Disjoint pairs (the vast majority) have IoU = 0/union: never above a positive threshold, and a
  // NaN union compares false as well — decided without the division.  thr <= 0 takes the full path.
  const bool skip_disjoint = thr > 0.f;
  for (int col_b = c_beg; col_b < c_end; ++col_b) {
    const int col_size = min(n - col_b * 64, 64);
    __syncthreads();  // the previous tile has been consumed
    if (t < col_size) {
#pragma unroll
      for (int k = 0; k < 5; ++k) sb[t * 5 + k] = d[(size_t)(col_b * 64 + t) * 5 + k];
      if (sets) s_set[t] = sets[(size_t)p * n_max + col_b * 64 + t];
    }
    __syncthreads();
    if (t < row_size) {
      unsigned long long bits = 0;
      const int start = (row_b == col_b) ? t + 1 : 0;
      for (int i = start; i < col_size; ++i) {
        const float* q = sb + i * 5;
        if (skip_disjoint) {
          const float w = __fadd_rn(__fsub_rn(fmin_ref(cb[2], q[2]), fmax_ref(cb[0], q[0])), 1.f);
          const float h = __fadd_rn(__fsub_rn(fmin_ref(cb[3], q[3]), fmax_ref(cb[1], q[1])), 1.f);
          if (!(w > 0.f && h > 0.f)) continue;
        }
        if (sets && s_set[i] == my_set) continue;  // members of one set never suppress each other
        const float v = dev_iou(cb, q);
        if (ge ? (v >= thr) : (v > thr)) bits |= 1ull << i;
      }
      mask[((size_t)p * n_max + cur) * col_blocks + col_b] = bits;
    }
  }
}

struct ScanOut {
  float* out;        // (P, out_rows, 4) or nullptr
  float* out_score;  // (P, out_rows)
  int* keep;         // (P, n_max) or nullptr: kept positions in greedy order
  int* nkeep;        // (P) or nullptr
  int out_rows;      // rows per problem in out/out_score
  int write_rows;    // rows actually written (PrepareOutput's `count`)
  int clip_to_count; // 1: a problem writes min(write_rows, its box count) rows and zero-fills the rest of out_rows
  int pad_mode;      // 0: zeros, 1: wrap keep[i % nkeep] (Proposal is_train)
  int level_B;       // 0: problem p writes rows [p*out_rows, ...).  >0: p = l*B + b writes image b's
                     //    slice l of a (B, L*out_rows) level-major concat (models/FPN/builder.py:316-317)
  int level_L;
};

// One CTA (256 threads) per problem: greedy scan of the bitmask in 64-row blocks.
// Shared: removed words (col_blocks), `nbuf` blocks of mask rows, kept list (n_max ints).  With
// nbuf = 2 the next block's rows stream in (cp.async) while the current block is resolved.
__global__ void __launch_bounds__(256)
nms_scan_kernel(const float* __restrict__ dets, const int* __restrict__ counts, const int n_max,
                const unsigned long long* __restrict__ mask, const ScanOut o, const int nbuf) {
  extern __shared__ unsigned long long s_dyn[];
  const int p = blockIdx.x;
  const int n = counts ? counts[p] : n_max;
  const int cbs = (n_max + 63) >> 6;
  unsigned long long* s_removed = s_dyn;          // cbs words
  unsigned long long* s_rows0 = s_dyn + cbs;      // nbuf x 64 x cbs words (mask rows of a block)
  int* s_keep = reinterpret_cast<int*>(s_rows0 + (size_t)nbuf * 64 * cbs);  // n_max ints
  __shared__ int s_nkeep;
  __shared__ unsigned long long s_alive;
  const int tid = threadIdx.x;
  for (int i = tid; i < cbs; i += blockDim.x) s_removed[i] = 0ull;
  if (tid == 0) s_nkeep = 0;
  __syncthreads();
  const unsigned long long* m = mask + (size_t)p * n_max * cbs;
  const int nblocks = (n + 63) >> 6;
  // stage a block's mask rows (only words >= rb are defined / needed)
  auto stage_rows = [&](int rb, unsigned long long* dst) {
    const int rows = min(64, n - rb * 64), wpr = cbs - rb;
    for (int e = tid; e < rows * wpr; e += blockDim.x) {
      const int r = e / wpr, w = rb + e - r * wpr;
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + r * cbs + w);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(m + (size_t)(rb * 64 + r) * cbs + w)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (nblocks > 0) stage_rows(0, s_rows0);
  for (int rb = 0; rb < nblocks; ++rb) {
    const int rows = min(64, n - rb * 64);
    unsigned long long* s_rows = s_rows0 + (nbuf == 2 ? (size_t)(rb & 1) * 64 * cbs : 0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (nbuf == 2 && rb + 1 < nblocks) stage_rows(rb + 1, s_rows0 + (size_t)((rb + 1) & 1) * 64 * cbs);
    // The serial greedy dependency lives entirely inside the block's DIAGONAL words.  Warp 0 holds
    // them in registers (lane l: rows l and l+32) and walks the 64 rows with shuffles whose source
    // lane is known in advance — only three ALU operations per row sit on the dependent chain.
    // Row r's word has bits above r only, so rows >= 32 never touch the low half.
    if (tid < 32) {
      unsigned long long alive0 = ~s_removed[rb];
      if (rows < 64) alive0 &= (1ull << rows) - 1ull;   // (their words are never applied below)
      unsigned lo = (unsigned)alive0, hi = (unsigned)(alive0 >> 32);
      const uint2* diag = reinterpret_cast<const uint2*>(s_rows + rb);
      const uint2 wa = diag[(size_t)tid * cbs], wb = diag[(size_t)(tid + 32) * cbs];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const unsigned wx = __shfl_sync(0xffffffffu, wa.x, r), wy = __shfl_sync(0xffffffffu, wa.y, r);
        const unsigned m = (unsigned)((int)(lo << (31 - r)) >> 31);   // all ones if row r is alive
        lo &= ~(wx & m);
        hi &= ~(wy & m);
      }
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const unsigned wy = __shfl_sync(0xffffffffu, wb.y, r);
        const unsigned m = (unsigned)((int)(hi << (31 - r)) >> 31);
        hi &= ~(wy & m);
      }
      if (tid == 0) s_alive = ((unsigned long long)hi << 32) | lo;
    }
    const int nk0 = s_nkeep;
    __syncthreads();  // s_alive is published; s_removed[rb] / s_nkeep have been read before they change
    const unsigned long long alive = s_alive;
    // OR the kept rows' words into `removed`: 16 lanes per word, each takes the rows r = g (mod 16);
    // the loads are independent, the 16 partial words meet in a shuffle tree.
    {
      const int g = tid & 15;
      const unsigned hmask = 0xFFFFu << (tid & 16);  // the two half-warps own different words
      for (int w = rb + 1 + (tid >> 4); w < cbs; w += (int)(blockDim.x >> 4)) {
        unsigned long long acc = 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = g + 16 * j;
          if ((alive >> r) & 1ull) acc |= s_rows[r * cbs + w];
        }
#pragma unroll
        for (int o = 8; o; o >>= 1) acc |= __shfl_xor_sync(hmask, acc, o);
        if (g == 0) s_removed[w] |= acc;
      }
    }
    if (tid < 64 && ((alive >> tid) & 1ull))
      s_keep[nk0 + __popcll(alive & ((1ull << tid) - 1ull))] = rb * 64 + tid;
    if (tid == 0) s_nkeep = nk0 + __popcll(alive);
    __syncthreads();
    if (nbuf == 1 && rb + 1 < nblocks) stage_rows(rb + 1, s_rows0);
  }
  const int nk = s_nkeep;
  if (o.nkeep && tid == 0) o.nkeep[p] = nk;
  if (o.keep)
    for (int i = tid; i < n_max; i += blockDim.x) o.keep[(size_t)p * n_max + i] = i < nk ? s_keep[i] : 0;
  if (o.out) {  // PrepareOutput, proposal_v3.cu:387-419 / nms.cu:208-231
    const float* d = dets + (size_t)p * n_max * 5;
    const int wr = o.clip_to_count ? min(o.write_rows, n) : o.write_rows;
    for (int i = tid; i < (o.clip_to_count ? o.out_rows : o.write_rows); i += blockDim.x) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      float s = 0.f;
      int k = -1;
      if (i >= wr) k = -1;  // beyond what this level can fill: a zero row (ProposalTarget skips y2 == 0 rows)
      else if (i < nk) k = s_keep[i];
      else if (o.pad_mode == 1 && nk > 0) k = s_keep[i % nk];
      if (k >= 0) {
        b = make_float4(d[(size_t)k * 5], d[(size_t)k * 5 + 1], d[(size_t)k * 5 + 2], d[(size_t)k * 5 + 3]);
        s = d[(size_t)k * 5 + 4];
      }
      size_t row = (size_t)p * o.out_rows + i;
      if (o.level_B > 0) {
        const int l = p / o.level_B, bi = p - l * o.level_B;
        row = ((size_t)bi * o.level_L + l) * o.out_rows + i;
      }
      float* op = o.out + row * 4;
      op[0] = b.x; op[1] = b.y; op[2] = b.z; op[3] = b.w;
      o.out_score[row] = s;
    }
  }
}

size_t scan_smem_bytes(int n_max, int nbuf) {
  const size_t cbs = (size_t)(n_max + 63) / 64;
  return cbs * 8 + (size_t)nbuf * 64 * cbs * 8 + (size_t)n_max * 4;
}

// --------------------------------------------------------------------------------------------
// Proposal_v3 stage 1: top-`pre` fg scores of one image -> decoded, clipped, filtered dets.
// --------------------------------------------------------------------------------------------
constexpr int kMaxAnchors = 16;
struct ProposalLevel {
  const float* cls_prob;   // (B, 2A, H, W)
  const float* bbox_pred;  // (B, 4A, H, W)
  float anchors[kMaxAnchors * 4];  // base anchors of this stride, proposal_v3-inl.h:280-318
  int H, W, stride, pre;   // pre = min(rpn_pre_nms_top_n, A*H*W) of this level
  int nchunks;             // > 1: the level is pre-selected chunk by chunk (proposal_chunk_topk_kernel)
  int chunk_base;          // first CTA of this level in the chunk kernel's grid
  size_t cand_off;         // offset (keys) of this level's candidates: [b][chunk][pre]
};
struct ProposalParams {
  ProposalLevel lvl[SDET_MAX_LEVELS];
  const float* im_info;    // (B, 3)
  int B, A, num_levels;
  int pre_max, k_pow2;     // row pitch of dets / smem keys (max over levels)
  int min_size;
  int iou_loss;            // IoUPredKernel (:163-205): additive decode, padded cells score -1 BEFORE the sort
  float* dets;             // (num_levels*B, pre_max, 5), problem p = l*B + b
  int* counts;             // (num_levels*B) = pre of the level
  unsigned long long* cand;  // chunk winners (keys), zero-padded
  int cache_keys;            // u64 slots of shared memory after the selection buffer (0 = none): the keys of
                             // a problem are materialised there once instead of being re-derived from
                             // global memory in each of the 4-5 selection sweeps (those sweeps are
                             // latency-bound: one dependent load per 1024 elements)
};

// Large levels (P2 of an 800x1333 image has 201 600 anchors) would leave one CTA sweeping the whole
// score map while 147 SMs idle.  They are split into chunks of kChunkElems scores: every chunk's
// CTA selects its own top-`pre` keys; the per-problem CTA then selects among nchunks*pre keys.
// Exact: the top-k of a union is contained in the union of the parts' top-k, and keys are unique.
constexpr int kChunkElems = 16384;
inline int level_chunks(int count) { return count > kChunkElems ? (count + kChunkElems - 1) / kChunkElems : 1; }

// (rh, rw) = cells inside the un-padded image, only consulted by the iou_loss path; pass W for rw to disable
__device__ __forceinline__ uint64_t proposal_key(const float* fg, int i, int HW, int A, unsigned magic, int W = 1,
                                                 int rh = 0x7FFFFFFF, int rw = 0x7FFFFFFF) {
  // element i is visited in MEMORY order (a, h, w) for coalescing; its reference index (the
  // stable-sort tie breaker) is (h*W + w)*A + a  (ProposalGridKernel, :73-75)
  int a = (int)__umulhi((unsigned)i, magic), r = i - a * HW;  // a = i / HW without a divide (+ fix-up)
  while (r >= HW) {
    r -= HW;
    ++a;
  }
  float sc = __ldg(fg + i);
  if (rh != 0x7FFFFFFF) {
    const int h = r / W, w = r - h * W;
    if (h >= rh || w >= rw) sc = -1.0f;
  }
  return sdet::make_key(sc, (uint32_t)(r * A + a));
}

__global__ void __launch_bounds__(kTopkThreads)
proposal_chunk_topk_kernel(const __grid_constant__ ProposalParams p) {
  extern __shared__ unsigned long long s_sel[];
  __shared__ uint32_t s_hist[sdet::kRadixBins];
  int l = 0;
  for (int i = 1; i < p.num_levels; ++i)
    if (p.lvl[i].nchunks > 1 && (int)blockIdx.x >= p.lvl[i].chunk_base) l = i;
  while (p.lvl[l].nchunks <= 1) ++l;  // (the first chunked level, when blockIdx.x precedes every later base)
  const ProposalLevel& L = p.lvl[l];
  const int rel = blockIdx.x - L.chunk_base;
  const int b = rel / L.nchunks, chunk = rel - b * L.nchunks;
  const int A = p.A, HW = L.H * L.W, count = A * HW, pre = L.pre;
  const float* fg = L.cls_prob + (size_t)b * 2 * count + count;
  const unsigned magic = 0xFFFFFFFFu / (unsigned)HW;
  const int i0 = chunk * kChunkElems;
  const int n = min(kChunkElems, count - i0);
  int rh = 0x7FFFFFFF, rw = 0x7FFFFFFF;
  if (p.iou_loss) {
    rh = (int)__fdiv_rn(__ldg(p.im_info + b * 3), (float)L.stride);
    rw = (int)__fdiv_rn(__ldg(p.im_info + b * 3 + 1), (float)L.stride);
  }
  const int k = min(pre, n);
  if (p.cache_keys >= kChunkElems) {
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_sel) + p.k_pow2;
#pragma unroll
    for (int it = 0; it < kChunkElems / kTopkThreads; ++it) {  // independent loads, all in flight together
      const int i = threadIdx.x + it * kTopkThreads;
      if (i < n) s_keys[i] = proposal_key(fg, i0 + i, HW, A, magic, L.W, rh, rw);
    }
    __syncthreads();
    auto key_at = [&](int i) -> uint64_t { return s_keys[i]; };
    sdet::block_topk_sorted<false>(n, k, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), p.k_pow2);
  } else {
    auto key_at = [&](int i) -> uint64_t { return proposal_key(fg, i0 + i, HW, A, magic, L.W, rh, rw); };
    sdet::block_topk_sorted<false>(n, k, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), p.k_pow2);
  }
  unsigned long long* dst = p.cand + L.cand_off + ((size_t)b * L.nchunks + chunk) * pre;
  for (int j = threadIdx.x; j < pre; j += blockDim.x) dst[j] = (j < k) ? s_sel[j] : 0ull;  // 0 < every real key
}

__global__ void __launch_bounds__(kTopkThreads)
proposal_topk_kernel(const __grid_constant__ ProposalParams p) {
  extern __shared__ unsigned long long s_sel[];
  __shared__ uint32_t s_hist[sdet::kRadixBins];
  const int prob = blockIdx.x;
  const int l = prob / p.B, b = prob - l * p.B;
  const ProposalLevel& L = p.lvl[l];
  const int A = p.A, H = L.H, W = L.W, HW = H * W;
  const int count = A * HW, pre = L.pre;
  const float* fg = L.cls_prob + (size_t)b * 2 * count + count;  // second half = foreground (:522)
  if (L.nchunks > 1) {  // chunk winners, written by proposal_chunk_topk_kernel
    const unsigned long long* cand = p.cand + L.cand_off + (size_t)b * L.nchunks * pre;
    const int nc = L.nchunks * pre;
    if (nc <= p.cache_keys) {
      uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_sel) + p.k_pow2;
#pragma unroll 4
      for (int i = threadIdx.x; i < nc; i += kTopkThreads) s_keys[i] = cand[i];
      __syncthreads();
      auto key_at = [&](int i) -> uint64_t { return s_keys[i]; };
      sdet::block_topk_sorted(nc, pre, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), p.k_pow2);
    } else {
      auto key_at = [&](int i) -> uint64_t { return cand[i]; };
      sdet::block_topk_sorted(nc, pre, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), p.k_pow2);
    }
  } else {
    const unsigned magic = 0xFFFFFFFFu / (unsigned)HW;
    int rh = 0x7FFFFFFF, rw = 0x7FFFFFFF;
    if (p.iou_loss) {
      rh = (int)__fdiv_rn(__ldg(p.im_info + b * 3), (float)L.stride);
      rw = (int)__fdiv_rn(__ldg(p.im_info + b * 3 + 1), (float)L.stride);
    }
    if (count <= p.cache_keys) {
      uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_sel) + p.k_pow2;
#pragma unroll 4
      for (int i = threadIdx.x; i < count; i += kTopkThreads) s_keys[i] = proposal_key(fg, i, HW, A, magic, W, rh, rw);
      __syncthreads();
      auto key_at = [&](int i) -> uint64_t { return s_keys[i]; };
      sdet::block_topk_sorted(count, pre, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), p.k_pow2);
    } else {
      auto key_at = [&](int i) -> uint64_t { return proposal_key(fg, i, HW, A, magic, W, rh, rw); };
      sdet::block_topk_sorted(count, pre, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), p.k_pow2);
    }
  }
  if (threadIdx.x == 0) p.counts[prob] = pre;
  // decode only the winners
  const float im_h = __ldg(p.im_info + b * 3), im_w = __ldg(p.im_info + b * 3 + 1);
  const float im_s = __ldg(p.im_info + b * 3 + 2);
  const float* dl = L.bbox_pred + (size_t)b * 4 * count;
  const float fs = (float)L.stride;
  float* dets = p.dets + (size_t)prob * p.pre_max * 5;
  for (int j = threadIdx.x; j < p.pre_max; j += blockDim.x) {
    float* o = dets + (size_t)j * 5;
    if (j >= pre) {  // padding rows of a level with fewer than pre_max anchors: never read (counts)
      o[0] = o[1] = o[2] = o[3] = o[4] = 0.f;
      continue;
    }
    const uint64_t key = reinterpret_cast<const uint64_t*>(s_sel)[j];
    const int index = (int)sdet::key_index(key);
    float sc = sdet::key_score(key);
    const int a = index % A, w = (index / A) % W, h = index / A / W;
    const float bx1 = __fadd_rn(L.anchors[a * 4 + 0], __fmul_rn((float)w, fs));  // :77-80
    const float by1 = __fadd_rn(L.anchors[a * 4 + 1], __fmul_rn((float)h, fs));
    const float bx2 = __fadd_rn(L.anchors[a * 4 + 2], __fmul_rn((float)w, fs));
    const float by2 = __fadd_rn(L.anchors[a * 4 + 3], __fmul_rn((float)h, fs));
    const float d0 = __ldg(dl + ((a * 4 + 0) * H + h) * W + w), d1 = __ldg(dl + ((a * 4 + 1) * H + h) * W + w);
    const float d2 = __ldg(dl + ((a * 4 + 2) * H + h) * W + w), d3 = __ldg(dl + ((a * 4 + 3) * H + h) * W + w);
    const float mx = __fsub_rn(im_w, 1.0f), my = __fsub_rn(im_h, 1.0f);
    float x1, y1, x2, y2;
    if (p.iou_loss) {  // IoUPredKernel :163-205
      x1 = __fadd_rn(bx1, d0); y1 = __fadd_rn(by1, d1); x2 = __fadd_rn(bx2, d2); y2 = __fadd_rn(by2, d3);
    } else {           // BBoxPredKernel :93-155
      const float width = __fadd_rn(__fsub_rn(bx2, bx1), 1.0f), height = __fadd_rn(__fsub_rn(by2, by1), 1.0f);
      const float ctr_x = __fadd_rn(bx1, __fmul_rn(0.5f, width)), ctr_y = __fadd_rn(by1, __fmul_rn(0.5f, height));
      const float dw = (float)fmin((double)d2, 4.135166556742356), dh = (float)fmin((double)d3, 4.135166556742356);
      const float pcx = __fadd_rn(__fmul_rn(d0, width), ctr_x), pcy = __fadd_rn(__fmul_rn(d1, height), ctr_y);
      const float pw = __fmul_rn(expf(dw), width), ph = __fmul_rn(expf(dh), height);
      x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
      y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
      x2 = __fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.0f);
      y2 = __fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 1.0f);
    }
    x1 = fmax_ref(fmin_ref(x1, mx), 0.0f);
    y1 = fmax_ref(fmin_ref(y1, my), 0.0f);
    x2 = fmax_ref(fmin_ref(x2, mx), 0.0f);
    y2 = fmax_ref(fmin_ref(y2, my), 0.0f);
    // FilterBoxKernel :211-235 (after top-k, original-image scale)
    const float ws_o = __fadd_rn(__fdiv_rn(__fsub_rn(x2, x1), im_s), 1.0f);
    const float hs_o = __fadd_rn(__fdiv_rn(__fsub_rn(y2, y1), im_s), 1.0f);
    const float msm = fmax_ref((float)p.min_size, 1.0f);
    const float ws = __fadd_rn(__fsub_rn(x2, x1), 1.0f), hs = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
    const float x_ctr = __fadd_rn(x1, __fdiv_rn(ws, 2.0f)), y_ctr = __fadd_rn(y1, __fdiv_rn(hs, 2.0f));
    if (ws_o < msm || hs_o < msm || x_ctr >= im_w || y_ctr >= im_h) {
      const float hm = __fdiv_rn(msm, 2.f);
      x1 = __fsub_rn(x1, hm);
      y1 = __fsub_rn(y1, hm);
      x2 = __fadd_rn(x2, hm);
      y2 = __fadd_rn(y2, hm);
      sc = -1.0f;
    }
    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = sc;
  }
}

// _contrib_NMS stage 1: top-`pre` rows of (count,5) proposals by column 4, gathered in order.
__global__ void __launch_bounds__(kTopkThreads)
rows_topk_kernel(const float* __restrict__ proposals, const int count, const int pre, const int k_pow2,
                 const int already_sorted, float* __restrict__ dets) {
  extern __shared__ unsigned long long s_sel[];
  __shared__ uint32_t s_hist[sdet::kRadixBins];
  const int b = blockIdx.x;
  const float* src = proposals + (size_t)b * count * 5;
  if (!already_sorted) {
    auto key_at = [&](int i) -> uint64_t { return sdet::make_key(__ldg(src + (size_t)i * 5 + 4), (uint32_t)i); };
    sdet::block_topk_sorted(count, pre, key_at, s_hist, reinterpret_cast<uint64_t*>(s_sel), k_pow2);
  }
  for (int j = threadIdx.x; j < pre; j += blockDim.x) {
    const int i = already_sorted ? j : (int)sdet::key_index(reinterpret_cast<const uint64_t*>(s_sel)[j]);
#pragma unroll
    for (int k = 0; k < 5; ++k) dets[((size_t)b * pre + j) * 5 + k] = __ldg(src + (size_t)i * 5 + k);
  }
}

// Base anchors exactly as proposal_v3-inl.h:280-318 (host, float).
void gen_anchors_v3(int stride, const float* ratios, int nr, const float* scales, int ns, float* out) {
  const float base2 = (float)(stride - 1.0);
  int k = 0;
  for (int j = 0; j < nr; ++j)
    for (int s = 0; s < ns; ++s) {
      const float w = base2 - 0.f + 1.0f, h = base2 - 0.f + 1.0f;
      const float x_ctr = (float)(0.f + 0.5 * (w - 1.0f)), y_ctr = (float)(0.f + 0.5 * (h - 1.0f));
      const float size = w * h;
      const float size_ratios = std::floor(size / ratios[j]);
      const float new_w = rintf(std::sqrt(size_ratios)) * scales[s];
      const float new_h = rintf((new_w / scales[s] * ratios[j])) * scales[s];
      out[k * 4 + 0] = x_ctr - 0.5f * (new_w - 1.0f);
      out[k * 4 + 1] = y_ctr - 0.5f * (new_h - 1.0f);
      out[k * 4 + 2] = x_ctr + 0.5f * (new_w - 1.0f);
      out[k * 4 + 3] = y_ctr + 0.5f * (new_h - 1.0f);
      ++k;
    }
}

// --------------------------------------------------------------------------------------------
// Legacy _contrib_Proposal / _contrib_Proposal_v2 (operator_cxx/contrib/proposal.cu, proposal_v2.cu):
// the padded-cell mask and the min-size (and v2 scale) filter run BEFORE the sort, so every anchor
// has to be decoded: one coalesced pass writes (count,5) rows in reference index order, then the
// generic rows_topk -> mask -> scan pipeline takes over.
// --------------------------------------------------------------------------------------------
struct LegacyParams {
  const float* cls_prob;
  const float* bbox_pred;
  const float* im_info;
  const float* valid_ranges;  // (B,2) or nullptr
  const float* grid_anchors;  // (H*W*A,4) or nullptr: _contrib_GenProposal takes the shifted anchors as input
  float anchors[kMaxAnchors * 4];
  int A, H, W, stride, min_size, iou_loss, version, filter_scales;
  float* props;  // (B, count, 5)
};

__global__ void __launch_bounds__(256) proposal_legacy_decode_kernel(const __grid_constant__ LegacyParams p) {
  const int b = blockIdx.y;
  const int A = p.A, H = p.H, W = p.W, HW = H * W, count = A * HW;
  const float im_h = __ldg(p.im_info + b * 3), im_w = __ldg(p.im_info + b * 3 + 1), im_s = __ldg(p.im_info + b * 3 + 2);
  const int real_h = (int)__fdiv_rn(im_h, (float)p.stride), real_w = (int)__fdiv_rn(im_w, (float)p.stride);
  const float min_size = __fmul_rn((float)p.min_size, im_s);
  const float mx = __fsub_rn(im_w, 1.0f), my = __fsub_rn(im_h, 1.0f);
  const float* fg = p.cls_prob + (size_t)b * 2 * count + count;
  const float* dl = p.bbox_pred + (size_t)b * 4 * count;
  const float fs = (float)p.stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const int a = i / HW, r = i - a * HW, h = r / W, w = r - h * W;  // memory order (a,h,w): coalesced reads
    float bx1, by1, bx2, by2;
    if (p.grid_anchors) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.grid_anchors) + (size_t)r * A + a);
      bx1 = g.x; by1 = g.y; bx2 = g.z; by2 = g.w;
    } else {
      bx1 = __fadd_rn(p.anchors[a * 4 + 0], __fmul_rn((float)w, fs));
      by1 = __fadd_rn(p.anchors[a * 4 + 1], __fmul_rn((float)h, fs));
      bx2 = __fadd_rn(p.anchors[a * 4 + 2], __fmul_rn((float)w, fs));
      by2 = __fadd_rn(p.anchors[a * 4 + 3], __fmul_rn((float)h, fs));
    }
    float sc = __ldg(fg + i);
    const float d0 = __ldg(dl + (a * 4 + 0) * HW + r), d1 = __ldg(dl + (a * 4 + 1) * HW + r);
    const float d2 = __ldg(dl + (a * 4 + 2) * HW + r), d3 = __ldg(dl + (a * 4 + 3) * HW + r);
    float x1, y1, x2, y2;
    if (p.iou_loss) {
      x1 = __fadd_rn(bx1, d0); y1 = __fadd_rn(by1, d1); x2 = __fadd_rn(bx2, d2); y2 = __fadd_rn(by2, d3);
    } else {
      const float width = __fadd_rn(__fsub_rn(bx2, bx1), 1.0f), height = __fadd_rn(__fsub_rn(by2, by1), 1.0f);
      const float ctr_x = __fadd_rn(bx1, __fmul_rn(0.5f, __fsub_rn(width, 1.0f)));
      const float ctr_y = __fadd_rn(by1, __fmul_rn(0.5f, __fsub_rn(height, 1.0f)));
      const float pcx = __fadd_rn(__fmul_rn(d0, width), ctr_x), pcy = __fadd_rn(__fmul_rn(d1, height), ctr_y);
      const float pw = __fmul_rn(expf(d2), width), ph = __fmul_rn(expf(d3), height);
      const float hw_ = __fmul_rn(0.5f, __fsub_rn(pw, 1.0f)), hh_ = __fmul_rn(0.5f, __fsub_rn(ph, 1.0f));
      x1 = __fsub_rn(pcx, hw_); y1 = __fsub_rn(pcy, hh_); x2 = __fadd_rn(pcx, hw_); y2 = __fadd_rn(pcy, hh_);
    }
    x1 = fmax_ref(fmin_ref(x1, mx), 0.0f); y1 = fmax_ref(fmin_ref(y1, my), 0.0f);
    x2 = fmax_ref(fmin_ref(x2, mx), 0.0f); y2 = fmax_ref(fmin_ref(y2, my), 0.0f);
    if (h >= real_h || w >= real_w) sc = -1.0f;
    const float iw = __fadd_rn(__fsub_rn(x2, x1), 1.0f), ih = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
    if (iw < min_size || ih < min_size) {
      const float hm = __fdiv_rn(min_size, 2.f);
      x1 = __fsub_rn(x1, hm); y1 = __fsub_rn(y1, hm); x2 = __fadd_rn(x2, hm); y2 = __fadd_rn(y2, hm);
      sc = -1.0f;
    } else if (p.version == 2 && p.filter_scales) {
      const float v0 = __ldg(p.valid_ranges + b * 2), v1 = __ldg(p.valid_ranges + b * 2 + 1);
      const float ar = __fmul_rn(iw, ih);
      if (ar < __fmul_rn(v0, v0) || ar > __fmul_rn(v1, v1)) sc = -1.0f;
    }
    float* o = p.props + ((size_t)b * count + (size_t)r * A + a) * 5;  // reference index (h*W+w)*A + a
    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = sc;
  }
}

// _contrib_GenProposal PrepareOutput (generate_proposal.cu:268-287): rows < out_size are the sorted
// proposals (x1,y1,x2,y2,score); later rows get columns 1..4 zeroed (column 0 is zeroed here too — the
// reference leaves it uninitialised).
__global__ void gen_proposal_out_kernel(const float* __restrict__ dets, const int pre, const int out_rows,
                                        float* __restrict__ out) {
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < out_rows * 5; i += gridDim.x * blockDim.x)
    out[(size_t)b * out_rows * 5 + i] = (i / 5 < pre) ? dets[(size_t)b * pre * 5 + i] : 0.f;
}

void gen_anchors_legacy(int stride, const float* ratios, int nr, const float* scales, int ns, float* out) {
  const float base2 = (float)(stride - 1.0);
  int k = 0;
  for (int j = 0; j < nr; ++j)
    for (int s = 0; s < ns; ++s) {
      const float w = base2 - 0.f + 1.0f, h = base2 - 0.f + 1.0f;
      const float x_ctr = (float)(0.f + 0.5 * (w - 1.0f)), y_ctr = (float)(0.f + 0.5 * (h - 1.0f));
      const float size = w * h;
      const float size_ratios = std::floor(size / ratios[j]);
      const float new_w = std::floor(std::sqrt(size_ratios) + 0.5f) * scales[s];   // proposal-inl.h:302
      const float new_h = std::floor((new_w / scales[s] * ratios[j]) + 0.5f) * scales[s];
      out[k * 4 + 0] = x_ctr - 0.5f * (new_w - 1.0f);
      out[k * 4 + 1] = y_ctr - 0.5f * (new_h - 1.0f);
      out[k * 4 + 2] = x_ctr + 0.5f * (new_w - 1.0f);
      out[k * 4 + 3] = y_ctr + 0.5f * (new_h - 1.0f);
      ++k;
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace layout for P problems of n boxes: dets (P,n,5) f32 | mask (P,n,cbs) u64
size_t nms_ws_bytes(int P, int n) {
  const size_t cbs = (size_t)(n + 63) / 64;
  return align_up((size_t)P * n * 5 * 4, 256) + align_up((size_t)P * n * cbs * 8, 256);
}

int run_mask_and_scan(const float* dets, const int* counts, int P, int n, float thr, int ge,
                      unsigned long long* mask, const ScanOut& so, cudaStream_t st, const float* sets = nullptr) {
  // every limit is checked before anything is enqueued
  const int nbuf = scan_smem_bytes(n, 2) <= 96 * 1024 ? 2 : 1;
  const size_t smem = scan_smem_bytes(n, nbuf);
  if (n > 12288 || smem > 200 * 1024)
    return sdet::fail(SDET_ERR_UNSUPPORTED, "NMS over %d boxes needs %zu B shared memory", n, smem);
  if (P > 65535) return sdet::fail(SDET_ERR_UNSUPPORTED, "more than 65535 NMS problems in one call (grid.z)");
  const int cbs = (n + 63) / 64;
  const long long tiles = (long long)P * cbs * (cbs + 1) / 2;
  const int tpc = (int)std::max<long long>(1, std::min<long long>(cbs, tiles / 1200));
  dim3 grid((unsigned)((cbs + tpc - 1) / tpc), (unsigned)cbs, (unsigned)P);
  nms_mask_kernel<<<grid, 64, 0, st>>>(dets, counts, n, thr, ge, mask, sets, tpc);
  SDET_LAUNCH_CHECK("nms_mask_kernel");
  if (smem > 48 * 1024)  // per device and cheap: set on every launch, no process-wide cache
    SDET_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_scan_kernel<<<(unsigned)P, 256, smem, st>>>(dets, counts, n, mask, so, nbuf);
  SDET_LAUNCH_CHECK("nms_scan_kernel");
  return SDET_OK;
}

// The opt-in for more than 48 KB of dynamic shared memory is per device and per kernel: it is set on every launch
// that needs it (cheap) instead of being cached in process-wide statics, which broke when a second entry point
// asked for less, on a second device, or from a second thread.
template <typename K>
int ensure_smem(K kernel, size_t bytes) {
  if (bytes > 200 * 1024) return sdet::fail(SDET_ERR_UNSUPPORTED, "top-k needs %zu B shared memory", bytes);
  if (bytes > 48 * 1024)
    SDET_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SDET_OK;
}

}  // namespace

extern "C" int sdet_decode_bbox(const float* rois, const float* bbox_pred, const float* im_info,
                                float* out, int B, int N, int K4, const float* bbox_mean,
                                const float* bbox_std, int class_agnostic, int decode_type,
                                void* stream) {
  SDET_REQUIRE(rois && bbox_pred && im_info && out && bbox_mean && bbox_std, "NULL argument");
  SDET_REQUIRE(B > 0 && N > 0 && K4 >= 4 && K4 % 4 == 0, "bad shape (bbox_pred last dim must be 4*K)");
  SDET_REQUIRE(decode_type == 0 || decode_type == 1, "bbox_decode_type must be xywh(0) or xyxy(1)");
  SDET_REQUIRE(!class_agnostic || K4 >= 8, "class_agnostic decode reads class slot 1 (decodebbox.cc:54)");
  SDET_REQUIRE((reinterpret_cast<uintptr_t>(rois) | reinterpret_cast<uintptr_t>(bbox_pred) |
                reinterpret_cast<uintptr_t>(out)) % 16 == 0, "rois / bbox_pred / out must be 16-byte aligned");
  DecodeParams p;
  for (int i = 0; i < 4; ++i) {
    p.mean[i] = bbox_mean[i];
    p.std[i] = bbox_std[i];
  }
  p.class_agnostic = class_agnostic ? 1 : 0;
  p.xyxy = decode_type;
  const int K = K4 / 4;
  const int total = B * N * (class_agnostic ? 1 : K);
  const int threads = 256;
  int blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  decode_bbox_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(rois), reinterpret_cast<const float4*>(bbox_pred), im_info,
      reinterpret_cast<float4*>(out), N, K, total, p);
  SDET_LAUNCH_CHECK("decode_bbox_kernel");
  return SDET_OK;
}

extern "C" size_t sdet_nms_workspace(int problems, int n) {
  if (problems <= 0 || n <= 0) return 0;
  return nms_ws_bytes(problems, n);
}

extern "C" int sdet_nms_sorted(const float* dets, const int* counts, int problems, int n, float thresh,
                               int ge, int* keep, int* nkeep, void* workspace, size_t workspace_bytes,
                               void* stream) {
  SDET_REQUIRE(dets && keep && nkeep && workspace, "NULL argument");
  SDET_REQUIRE(problems > 0 && n > 0, "problems and n must be > 0");
  if (workspace_bytes < nms_ws_bytes(problems, n))
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", nms_ws_bytes(problems, n));
  auto* mask = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) +
                                                     align_up((size_t)problems * n * 5 * 4, 256));
  ScanOut so{};
  so.keep = keep;
  so.nkeep = nkeep;
  return run_mask_and_scan(dets, counts, problems, n, thresh, ge, mask, so, (cudaStream_t)stream);
}


// set_nms (operator_py/nms.py:77-107) over pre-sorted boxes: `sets` (P,n) holds column 5.
extern "C" int sdet_set_nms_sorted(const float* dets, const float* sets, const int* counts, int problems, int n,
                                   float thresh, int* keep, int* nkeep, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  SDET_REQUIRE(dets && sets && keep && nkeep && workspace, "NULL argument");
  SDET_REQUIRE(problems > 0 && n > 0, "problems and n must be > 0");
  if (workspace_bytes < nms_ws_bytes(problems, n))
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", nms_ws_bytes(problems, n));
  auto* mask = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) +
                                                     align_up((size_t)problems * n * 5 * 4, 256));
  ScanOut so{};
  so.keep = keep;
  so.nkeep = nkeep;
  return run_mask_and_scan(dets, counts, problems, n, thresh, /*ge=*/0, mask, so, (cudaStream_t)stream, sets);
}

namespace {
// py_weighted_nms (operator_py/nms.py:110-157) after the greedy scan at thresh_lo.  A box j leaves the
// pool at the first kept box k with IoU(k,j) > lo (itself if it is kept: IoU = 1), so "still in the
// pool when i is on top" is first_suppressor(j) >= i, and every kept i averages the boxes of the pool
// with IoU(i,j) > hi, weighted by score.
__global__ void __launch_bounds__(256)
weighted_vote_kernel(const float* __restrict__ dets, const int* __restrict__ counts, const int n_max,
                     const unsigned long long* __restrict__ mask, const int* __restrict__ keep,
                     const int* __restrict__ nkeep, const float thr_hi, int* __restrict__ first_sup,
                     float* __restrict__ out, int* __restrict__ nout) {
  const int p = blockIdx.x;
  const int n = counts ? counts[p] : n_max;
  const int cbs = (n_max + 63) >> 6;
  const float* d = dets + (size_t)p * n_max * 5;
  const unsigned long long* m = mask + (size_t)p * n_max * cbs;
  const int* kp = keep + (size_t)p * n_max;
  int* fs = first_sup + (size_t)p * n_max;
  float* o = out + (size_t)p * n_max * 5;
  const int nk = thr_hi < 1.0f ? nkeep[p] : 0;  // hi >= 1: the top box does not even vote for itself -> break
  if (threadIdx.x == 0) nout[p] = nk;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int f = n;  // (every box is kept or suppressed by a kept one, so this is always overwritten)
    for (int q = 0; q < nk; ++q) {
      const int k = kp[q];
      if (k > j) break;
      if (k == j || ((m[(size_t)k * cbs + (j >> 6)] >> (j & 63)) & 1ull)) {
        f = k;
        break;
      }
    }
    fs[j] = f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int q = warp; q < nk; q += nwarps) {
    const int k = kp[q];
    const float* bk = d + (size_t)k * 5;
    float sw = 0.f, sx1 = 0.f, sy1 = 0.f, sx2 = 0.f, sy2 = 0.f;
    for (int j = k + lane; j < n; j += 32) {
      if (fs[j] < k) continue;
      const float* bj = d + (size_t)j * 5;
      if (dev_iou(bk, bj) > thr_hi) {
        const float s = bj[4];
        sw += s;
        sx1 += s * bj[0];
        sy1 += s * bj[1];
        sx2 += s * bj[2];
        sy2 += s * bj[3];
      }
    }
    for (int off = 16; off; off >>= 1) {
      sw += __shfl_xor_sync(0xffffffffu, sw, off);
      sx1 += __shfl_xor_sync(0xffffffffu, sx1, off);
      sy1 += __shfl_xor_sync(0xffffffffu, sy1, off);
      sx2 += __shfl_xor_sync(0xffffffffu, sx2, off);
      sy2 += __shfl_xor_sync(0xffffffffu, sy2, off);
    }
    if (lane == 0) {
      float* r = o + (size_t)q * 5;
      r[0] = sx1 / sw; r[1] = sy1 / sw; r[2] = sx2 / sw; r[3] = sy2 / sw; r[4] = bk[4];
    }
  }
}
}  // namespace

extern "C" size_t sdet_weighted_nms_workspace(int problems, int n) {
  if (problems <= 0 || n <= 0) return 0;
  return nms_ws_bytes(problems, n) + align_up((size_t)problems * n * 4, 256) * 2 + align_up((size_t)problems * 4, 256);
}

extern "C" int sdet_weighted_nms_sorted(const float* dets, const int* counts, int problems, int n, float thresh_lo,
                                        float thresh_hi, float* out, int* nout, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  SDET_REQUIRE(dets && out && nout && workspace, "NULL argument");
  SDET_REQUIRE(problems > 0 && n > 0, "problems and n must be > 0");
  SDET_REQUIRE(thresh_lo < 1.0f, "thresh_lo must be < 1 (a box always leaves the pool with itself)");
  if (workspace_bytes < sdet_weighted_nms_workspace(problems, n))
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", sdet_weighted_nms_workspace(problems, n));
  cudaStream_t st = (cudaStream_t)stream;
  char* w = static_cast<char*>(workspace);
  auto* mask = reinterpret_cast<unsigned long long*>(w + align_up((size_t)problems * n * 5 * 4, 256));
  char* x = w + nms_ws_bytes(problems, n);
  int* keep = reinterpret_cast<int*>(x); x += align_up((size_t)problems * n * 4, 256);
  int* fsup = reinterpret_cast<int*>(x); x += align_up((size_t)problems * n * 4, 256);
  int* nkeep = reinterpret_cast<int*>(x);
  ScanOut so{};
  so.keep = keep;
  so.nkeep = nkeep;
  if (int rc = run_mask_and_scan(dets, counts, problems, n, thresh_lo, /*ge=*/0, mask, so, st)) return rc;
  weighted_vote_kernel<<<(unsigned)problems, 256, 0, st>>>(dets, counts, n, mask, keep, nkeep, thresh_hi, fsup, out, nout);
  SDET_LAUNCH_CHECK("weighted_vote_kernel");
  return SDET_OK;
}

// workspace: dets (P,pre_max,5) | mask | counts (P)
static size_t proposal_ws_bytes(int P, int pre_max) { return nms_ws_bytes(P, pre_max) + align_up((size_t)P * 4, 256); }

static int level_pre(int A, int H, int W, int rpn_pre_nms_top_n) {
  const int count = A * H * W;
  int pre = rpn_pre_nms_top_n > 0 ? rpn_pre_nms_top_n : count;  // -1 = all (proposal_v3.cu:470-471)
  return pre > count ? count : pre;
}

static size_t proposal_cand_keys(int B, int A, const int* H, const int* W, int num_levels, int rpn_pre_nms_top_n) {
  size_t keys = 0;
  for (int l = 0; l < num_levels; ++l) {
    const int nch = level_chunks(A * H[l] * W[l]);
    if (nch > 1) keys += (size_t)B * nch * level_pre(A, H[l], W[l], rpn_pre_nms_top_n);
  }
  return keys;
}


extern "C" size_t sdet_proposal_v3_fpn_workspace(int B, int A, const int* H, const int* W, int num_levels,
                                                 int rpn_pre_nms_top_n) {
  if (B <= 0 || A <= 0 || !H || !W || num_levels <= 0) return 0;
  int pre_max = 0;
  for (int l = 0; l < num_levels; ++l) pre_max = std::max(pre_max, level_pre(A, H[l], W[l], rpn_pre_nms_top_n));
  return proposal_ws_bytes(B * num_levels, pre_max) +
         align_up(proposal_cand_keys(B, A, H, W, num_levels, rpn_pre_nms_top_n) * 8, 256);
}

extern "C" int sdet_proposal_v3_fpn(const float* const* cls_prob, const float* const* bbox_pred,
                                    const float* im_info, float* out, float* out_score, int B, int A,
                                    const int* H, const int* W, const int* feature_stride, int num_levels,
                                    const float* scales, int num_scales, const float* ratios,
                                    int num_ratios, int rpn_pre_nms_top_n, int rpn_post_nms_top_n,
                                    float threshold, int rpn_min_size, int iou_loss, int is_train,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  SDET_REQUIRE(cls_prob && bbox_pred && im_info && out && out_score && scales && ratios && workspace && H &&
               W && feature_stride, "NULL argument");
  SDET_REQUIRE(B > 0 && A > 0 && num_levels >= 1 && num_levels <= SDET_MAX_LEVELS, "bad shape");
  // CHECK_EQ(num_anchors, ratios.size() * scales.size())  (proposal_v3.cu:488)
  SDET_REQUIRE(A == num_scales * num_ratios, "num_anchors (%d) != len(ratios)*len(scales) (%d)", A,
               num_scales * num_ratios);
  if (A > kMaxAnchors) return sdet::fail(SDET_ERR_UNSUPPORTED, "more than %d anchors per cell", kMaxAnchors);
  SDET_REQUIRE(rpn_post_nms_top_n > 0, "rpn_post_nms_top_n must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  ProposalParams p{};
  int pre_max = 0, pre_min = INT_MAX, chunk_ctas = 0;
  size_t cand_keys = 0;
  for (int l = 0; l < num_levels; ++l) {
    SDET_REQUIRE(cls_prob[l] && bbox_pred[l] && H[l] > 0 && W[l] > 0, "level %d: bad pointer / shape", l);
    ProposalLevel& L = p.lvl[l];
    L.cls_prob = cls_prob[l];
    L.bbox_pred = bbox_pred[l];
    L.H = H[l]; L.W = W[l]; L.stride = feature_stride[l];
    L.pre = level_pre(A, H[l], W[l], rpn_pre_nms_top_n);
    gen_anchors_v3(feature_stride[l], ratios, num_ratios, scales, num_scales, L.anchors);
    pre_max = std::max(pre_max, L.pre);
    pre_min = std::min(pre_min, L.pre);
    L.nchunks = level_chunks(A * H[l] * W[l]);
    L.chunk_base = chunk_ctas;
    L.cand_off = cand_keys;
    if (L.nchunks > 1) {
      chunk_ctas += B * L.nchunks;
      cand_keys += (size_t)B * L.nchunks * L.pre;
    }
  }
  // rows per level in the output: `post` (:472-475); with is_train a level writes min(post, its pre) rows (wrap-
  // padded).  When the levels differ (P6 of an 800x1333 FPN has 819 anchors, rpn_post_nms_top_n is 2000) every level
  // keeps `post` rows like the reference's per-level outputs and the small level's remaining rows are ZERO.  (The
  // reference leaves them unwritten and, for images after the first, writes the level's rows with the shrunken
  // stride, i.e. into the previous image's block - proposal_v3.cu:471-476,:629-631, tests/test_oracle_ref_cxx.py;
  // that is not reproduced.)
  int post = rpn_post_nms_top_n, clip = 0;
  if (is_train) {
    if (pre_min == pre_max) post = std::min(rpn_post_nms_top_n, pre_min);
    else if (rpn_post_nms_top_n > pre_min) clip = 1;
  }
  const int P = B * num_levels;
  const size_t need = proposal_ws_bytes(P, pre_max) + align_up(cand_keys * 8, 256);
  if (workspace_bytes < need) return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
  char* wsb = static_cast<char*>(workspace);
  float* dets = reinterpret_cast<float*>(wsb);
  auto* mask = reinterpret_cast<unsigned long long*>(wsb + align_up((size_t)P * pre_max * 5 * 4, 256));
  int* counts = reinterpret_cast<int*>(wsb + nms_ws_bytes(P, pre_max));
  p.im_info = im_info;
  p.B = B; p.A = A; p.num_levels = num_levels;
  p.pre_max = pre_max;
  p.k_pow2 = sdet::next_pow2(pre_max);
  p.min_size = rpn_min_size;
  p.iou_loss = iou_loss ? 1 : 0;
  p.dets = dets;
  p.counts = counts;
  p.cand = reinterpret_cast<unsigned long long*>(wsb + proposal_ws_bytes(P, pre_max));
  // key cache: kChunkElems slots if the selection buffer leaves room for them (176 KB budget)
  p.cache_keys = ((size_t)p.k_pow2 * 8 + (size_t)kChunkElems * 8 <= 176 * 1024) ? kChunkElems : 0;
  const size_t smem = (size_t)p.k_pow2 * 8 + (size_t)p.cache_keys * 8;
  if (chunk_ctas > 0) {
    if (int rc = ensure_smem(proposal_chunk_topk_kernel, smem)) return rc;
    proposal_chunk_topk_kernel<<<(unsigned)chunk_ctas, kTopkThreads, smem, st>>>(p);
    SDET_LAUNCH_CHECK("proposal_chunk_topk_kernel");
  }
  if (int rc = ensure_smem(proposal_topk_kernel, smem)) return rc;
  proposal_topk_kernel<<<(unsigned)P, kTopkThreads, smem, st>>>(p);
  SDET_LAUNCH_CHECK("proposal_topk_kernel");
  ScanOut so{};
  so.out = out;
  so.out_score = out_score;
  so.out_rows = post;
  so.write_rows = post;
  so.clip_to_count = clip;
  so.pad_mode = is_train ? 1 : 0;
  so.level_B = num_levels > 1 ? B : 0;
  so.level_L = num_levels;
  return run_mask_and_scan(dets, counts, P, pre_max, threshold, /*ge=*/1, mask, so, st);  // `>=` (:319)
}

extern "C" size_t sdet_proposal_v3_workspace(int B, int A, int H, int W, int rpn_pre_nms_top_n) {
  return sdet_proposal_v3_fpn_workspace(B, A, &H, &W, 1, rpn_pre_nms_top_n);
}

extern "C" int sdet_proposal_v3(const float* cls_prob, const float* bbox_pred, const float* im_info,
                                float* out, float* out_score, int B, int A, int H, int W,
                                int feature_stride, const float* scales, int num_scales,
                                const float* ratios, int num_ratios, int rpn_pre_nms_top_n,
                                int rpn_post_nms_top_n, float threshold, int rpn_min_size,
                                int iou_loss, int is_train, void* workspace, size_t workspace_bytes,
                                void* stream) {
  return sdet_proposal_v3_fpn(&cls_prob, &bbox_pred, im_info, out, out_score, B, A, &H, &W, &feature_stride, 1,
                              scales, num_scales, ratios, num_ratios, rpn_pre_nms_top_n, rpn_post_nms_top_n,
                              threshold, rpn_min_size, iou_loss, is_train, workspace, workspace_bytes, stream);
}

extern "C" size_t sdet_contrib_nms_workspace(int B, int count, int rpn_pre_nms_top_n) {
  if (B <= 0 || count <= 0) return 0;
  int pre = rpn_pre_nms_top_n > 0 ? rpn_pre_nms_top_n : count;
  if (pre > count) pre = count;
  return nms_ws_bytes(B, pre);
}

extern "C" int sdet_contrib_nms(const float* proposals, float* out, float* out_score, int B, int count,
                                int rpn_pre_nms_top_n, int rpn_post_nms_top_n, float threshold,
                                int already_sorted, void* workspace, size_t workspace_bytes,
                                void* stream) {
  SDET_REQUIRE(proposals && out && out_score && workspace, "NULL argument");
  SDET_REQUIRE(B > 0 && count > 0 && rpn_post_nms_top_n > 0, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  int pre = rpn_pre_nms_top_n > 0 ? rpn_pre_nms_top_n : count;
  if (pre > count) pre = count;
  const int post = rpn_post_nms_top_n < pre ? rpn_post_nms_top_n : pre;
  if (workspace_bytes < nms_ws_bytes(B, pre))
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes", nms_ws_bytes(B, pre));
  float* dets = static_cast<float*>(workspace);
  auto* mask = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) +
                                                     align_up((size_t)B * pre * 5 * 4, 256));
  const int k_pow2 = sdet::next_pow2(pre);
  const size_t smem = (size_t)k_pow2 * 8;
  if (int rc = ensure_smem(rows_topk_kernel, smem)) return rc;
  rows_topk_kernel<<<(unsigned)B, kTopkThreads, smem, st>>>(proposals, count, pre, k_pow2, already_sorted, dets);
  SDET_LAUNCH_CHECK("rows_topk_kernel");
  ScanOut so{};
  so.out = out;
  so.out_score = out_score;
  so.out_rows = rpn_post_nms_top_n;  // declared output rows (nms-inl.h:96-99)
  so.write_rows = post;              // rows PrepareOutput touches (nms.cu:354-358)
  so.pad_mode = 0;
  return run_mask_and_scan(dets, nullptr, B, pre, threshold, /*ge=*/0, mask, so, st);  // `>` (nms.cu:140)
}

// workspace: props (B,count,5) | nms workspace (dets, mask) for pre boxes
extern "C" size_t sdet_proposal_legacy_workspace(int B, int A, int H, int W, int rpn_pre_nms_top_n) {
  if (B <= 0 || A <= 0 || H <= 0 || W <= 0) return 0;
  const int count = A * H * W;
  return align_up((size_t)B * count * 5 * 4, 256) + nms_ws_bytes(B, level_pre(A, H, W, rpn_pre_nms_top_n));
}

extern "C" int sdet_proposal_legacy(const float* cls_prob, const float* bbox_pred, const float* im_info,
                                    const float* valid_ranges, int version, float* out, float* out_score,
                                    int B, int A, int H, int W, int feature_stride, const float* scales,
                                    int num_scales, const float* ratios, int num_ratios, int rpn_pre_nms_top_n,
                                    int rpn_post_nms_top_n, float threshold, int rpn_min_size, int iou_loss,
                                    int is_train, int filter_scales, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  SDET_REQUIRE(cls_prob && bbox_pred && im_info && out && out_score && scales && ratios && workspace, "NULL argument");
  SDET_REQUIRE(version == 1 || version == 2, "version must be 1 (_contrib_Proposal) or 2 (_contrib_Proposal_v2)");
  SDET_REQUIRE(B > 0 && A > 0 && H > 0 && W > 0 && rpn_post_nms_top_n > 0, "bad shape");
  SDET_REQUIRE(A == num_scales * num_ratios, "num_anchors (%d) != len(ratios)*len(scales) (%d)", A, num_scales * num_ratios);
  SDET_REQUIRE(!(version == 2 && filter_scales) || valid_ranges, "filter_scales needs valid_ranges");
  if (A > kMaxAnchors) return sdet::fail(SDET_ERR_UNSUPPORTED, "more than %d anchors per cell", kMaxAnchors);
  cudaStream_t st = (cudaStream_t)stream;
  const int count = A * H * W;
  const int pre = level_pre(A, H, W, rpn_pre_nms_top_n);
  int post = std::min(rpn_post_nms_top_n, pre);
  if (version == 1 && !is_train) post = rpn_post_nms_top_n;
  if (workspace_bytes < sdet_proposal_legacy_workspace(B, A, H, W, rpn_pre_nms_top_n))
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes",
                      sdet_proposal_legacy_workspace(B, A, H, W, rpn_pre_nms_top_n));
  char* wsb = static_cast<char*>(workspace);
  float* props = reinterpret_cast<float*>(wsb);
  char* nws = wsb + align_up((size_t)B * count * 5 * 4, 256);
  float* dets = reinterpret_cast<float*>(nws);
  auto* mask = reinterpret_cast<unsigned long long*>(nws + align_up((size_t)B * pre * 5 * 4, 256));
  LegacyParams p{};
  p.cls_prob = cls_prob; p.bbox_pred = bbox_pred; p.im_info = im_info; p.valid_ranges = valid_ranges;
  gen_anchors_legacy(feature_stride, ratios, num_ratios, scales, num_scales, p.anchors);
  p.A = A; p.H = H; p.W = W; p.stride = feature_stride; p.min_size = rpn_min_size; p.iou_loss = iou_loss ? 1 : 0;
  p.version = version; p.filter_scales = filter_scales ? 1 : 0;
  p.props = props;
  dim3 grid((unsigned)std::min((count + 255) / 256, 148 * 8), (unsigned)B);
  proposal_legacy_decode_kernel<<<grid, 256, 0, st>>>(p);
  SDET_LAUNCH_CHECK("proposal_legacy_decode_kernel");
  const int k_pow2 = sdet::next_pow2(pre);
  const size_t smem = (size_t)k_pow2 * 8;
  if (int rc = ensure_smem(rows_topk_kernel, smem)) return rc;
  rows_topk_kernel<<<(unsigned)B, kTopkThreads, smem, st>>>(props, count, pre, k_pow2, 0, dets);
  SDET_LAUNCH_CHECK("rows_topk_kernel");
  ScanOut so{};
  so.out = out;
  so.out_score = out_score;
  so.out_rows = post;
  so.write_rows = post;
  so.pad_mode = (version == 1 && is_train) ? 1 : 0;
  return run_mask_and_scan(dets, nullptr, B, pre, threshold, /*ge=*/0, mask, so, st);  // `>` (proposal.cu:301)
}

extern "C" size_t sdet_gen_proposal_workspace(int B, int A, int H, int W, int rpn_pre_nms_top_n) {
  if (B <= 0 || A <= 0 || H <= 0 || W <= 0) return 0;
  const int count = A * H * W;
  return align_up((size_t)B * count * 5 * 4, 256) + align_up((size_t)B * level_pre(A, H, W, rpn_pre_nms_top_n) * 5 * 4, 256);
}

extern "C" int sdet_gen_proposal(const float* cls_prob, const float* bbox_pred, const float* im_info,
                                 const float* anchors, float* out, int B, int A, int H, int W, int feature_stride,
                                 int rpn_pre_nms_top_n, int rpn_min_size, int iou_loss, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  SDET_REQUIRE(cls_prob && bbox_pred && im_info && anchors && out && workspace, "NULL argument");
  SDET_REQUIRE(B > 0 && A > 0 && H > 0 && W > 0 && rpn_pre_nms_top_n > 0, "bad shape");
  SDET_REQUIRE((reinterpret_cast<uintptr_t>(anchors) & 15) == 0, "anchors must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int count = A * H * W;
  const int pre = level_pre(A, H, W, rpn_pre_nms_top_n);
  if (workspace_bytes < sdet_gen_proposal_workspace(B, A, H, W, rpn_pre_nms_top_n))
    return sdet::fail(SDET_ERR_WORKSPACE, "workspace too small: need %zu bytes",
                      sdet_gen_proposal_workspace(B, A, H, W, rpn_pre_nms_top_n));
  char* wsb = static_cast<char*>(workspace);
  float* props = reinterpret_cast<float*>(wsb);
  float* dets = reinterpret_cast<float*>(wsb + align_up((size_t)B * count * 5 * 4, 256));
  LegacyParams p{};
  p.cls_prob = cls_prob; p.bbox_pred = bbox_pred; p.im_info = im_info; p.grid_anchors = anchors;
  p.A = A; p.H = H; p.W = W; p.stride = feature_stride; p.min_size = rpn_min_size; p.iou_loss = iou_loss ? 1 : 0;
  p.version = 1;
  p.props = props;
  dim3 grid((unsigned)std::min((count + 255) / 256, 148 * 8), (unsigned)B);
  proposal_legacy_decode_kernel<<<grid, 256, 0, st>>>(p);
  SDET_LAUNCH_CHECK("proposal_legacy_decode_kernel");
  const int k_pow2 = sdet::next_pow2(pre);
  const size_t smem = (size_t)k_pow2 * 8;
  if (int rc = ensure_smem(rows_topk_kernel, smem)) return rc;
  rows_topk_kernel<<<(unsigned)B, kTopkThreads, smem, st>>>(props, count, pre, k_pow2, 0, dets);
  SDET_LAUNCH_CHECK("rows_topk_kernel");
  dim3 g2((unsigned)((rpn_pre_nms_top_n * 5 + 255) / 256), (unsigned)B);
  gen_proposal_out_kernel<<<g2, 256, 0, st>>>(dets, pre, rpn_pre_nms_top_n, out);
  SDET_LAUNCH_CHECK("gen_proposal_out_kernel");
  return SDET_OK;
}
