// Block-level exact top-k with the reference's ordering:  descending score, ties by ascending
// index  ==  thrust::stable_sort_by_key(score, order, greater<float>()) truncated to k
// (operator_cxx/contrib/proposal_v3.cu:564-568, nms.cu:307-311, models/FPN/get_top_proposal.py).
//
// One CTA owns one problem.  Keys are made unique by packing (sortable score bits, ~index) into
// 64 bits; an 11-bit-digit radix SELECT finds the k-th key in <= 6 sweeps over the scores (3 when
// the k-th score is untied), one more sweep compacts the k winners into shared memory, and a
// bitonic network sorts them.  Only the k winners are ever decoded / gathered, never all n.
#pragma once
#include <cstdint>

namespace sdet {

constexpr int kTopkThreads = 1024;
constexpr int kRadixBits = 11;
constexpr int kRadixBins = 1 << kRadixBits;

__device__ __forceinline__ uint32_t score_to_sortable(float s) {
  const uint32_t u = __float_as_uint(s + 0.f);  // -0.0 -> +0.0: thrust::greater ties them, the index decides
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // larger float -> larger uint
}
__device__ __forceinline__ uint64_t make_key(float s, uint32_t idx) {
  return ((uint64_t)score_to_sortable(s) << 32) | (uint64_t)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ uint32_t key_index(uint64_t key) { return 0xFFFFFFFFu - (uint32_t)key; }
__device__ __forceinline__ float key_score(uint64_t key) {
  const uint32_t u = (uint32_t)(key >> 32);
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// Sorts s[0..k_pow2) descending (k_pow2 a power of two); all threads of the CTA must call it.
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* s_sel, int k_pow2) {
  const int tid = threadIdx.x;
  if (k_pow2 <= (int)blockDim.x) {
    // one key per thread in a register: compare-exchange partners closer than a warp meet through
    // shuffles, only the strides >= 32 go through shared memory (15 barrier pairs instead of 55 for
    // 1024 keys)
    const bool on = tid < k_pow2;
    uint64_t v = on ? s_sel[tid] : 0ull;
    for (int size = 2; size <= k_pow2; size <<= 1) {
      const bool desc = ((tid & size) == 0);
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        uint64_t o;
        if (stride >= 32) {
          __syncthreads();
          if (on) s_sel[tid] = v;
          __syncthreads();
          o = on ? s_sel[tid ^ stride] : 0ull;
        } else {
          o = __shfl_xor_sync(0xffffffffu, v, stride);
        }
        const bool lower = (tid & stride) == 0;           // this thread keeps the pair's first slot
        const bool take_max = (lower == desc);            // descending run: first slot holds the max
        v = take_max ? (v > o ? v : o) : (v < o ? v : o);
      }
    }
    __syncthreads();
    if (on) s_sel[tid] = v;
    __syncthreads();
    return;
  }
  for (int size = 2; size <= k_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (k_pow2 >> 1); t += blockDim.x) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = s_sel[lo], b = s_sel[hi];
        if ((a < b) == desc) {
          s_sel[lo] = b;
          s_sel[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}

// Selects the k largest keys of {key(i) : i in [0,n)} into s_sel[0..k) (sorted descending; the
// rest of s_sel up to next_pow2(k) is zero).  `KeyAt(i)` returns the 64-bit key of element i and
// must be cheap and side-effect free (it is evaluated once per sweep).  All threads of the CTA
// must call this; blockDim.x == kTopkThreads.  s_hist: kRadixBins uint32; s_sel: next_pow2(k) u64.
// kSort = false leaves the k winners unsorted (callers that only feed a second selection).
template <bool kSort = true, typename KeyAt>
__device__ void block_topk_sorted(int n, int k, KeyAt key_at, uint32_t* s_hist, uint64_t* s_sel,
                                  int k_pow2) {
  __shared__ uint64_t s_prefix;
  __shared__ int s_need, s_done, s_cnt;
  __shared__ int s_wsum[32];
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_prefix = 0;
    s_need = k;
    s_done = (k >= n);  // everything is selected: threshold 0
    s_cnt = 0;
  }
  __syncthreads();
  for (int shift = 64 - kRadixBits; !s_done; shift -= kRadixBits) {
    const int sh = shift < 0 ? 0 : shift;
    const int bits = shift < 0 ? kRadixBits + shift : kRadixBits;  // last digit may be short
    const uint64_t hi_mask = (sh + bits >= 64) ? 0ull : (~0ull << (sh + bits));
    for (int i = tid; i < kRadixBins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint64_t prefix = s_prefix;
    // a warp whose 32 digits agree costs one shared atomic (RPN scores tie massively in degenerate
    // inputs, e.g. the all-zero-weight detection_infer_speed harness); otherwise plain atomics
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
      const int i = i0 + tid;
      unsigned digit = 0xFFFFFFFFu;
      if (i < n) {
        const uint64_t key = key_at(i);
        if ((key & hi_mask) == prefix) digit = (unsigned)(key >> sh) & ((1u << bits) - 1);
      }
      const unsigned lead = __shfl_sync(0xffffffffu, digit, 0);
      if (__all_sync(0xffffffffu, digit == lead)) {  // whole warp in one bin: one atomic
        if ((tid & 31) == 0 && digit != 0xFFFFFFFFu) atomicAdd(&s_hist[digit], 32u);
      } else if (digit != 0xFFFFFFFFu) {
        atomicAdd(&s_hist[digit], 1u);
      }
    }
    __syncthreads();
    {  // find the digit of the k-th key: block-wide scan of the histogram (two bins per thread) instead
       // of one warp walking 64 groups of bins (that serial walk was most of a sweep's latency)
      const int need = s_need;
      const int b0 = 2 * tid, b1 = 2 * tid + 1;
      const int h0 = (int)s_hist[b0], h1 = (int)s_hist[b1];
      int inc = h0 + h1;  // inclusive scan over threads (ascending bins)
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += t;
      }
      if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
      __syncthreads();
      if (tid < 32) {
        int w = s_wsum[tid];
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, w, o);
          if (tid >= o) w += t;
        }
        s_wsum[tid] = w;
      }
      __syncthreads();
      const int total = s_wsum[31];
      const int upto = inc + ((tid >> 5) ? s_wsum[(tid >> 5) - 1] : 0);  // keys in bins <= b1
      const int above1 = total - upto, above0 = above1 + h1;             // keys in bins above b1 / above b0
      int found = -1, found_cum = 0, hb = 0;
      if (above1 < need && need <= above1 + h1) { found = b1; found_cum = above1; hb = h1; }
      else if (above0 < need && need <= above0 + h0) { found = b0; found_cum = above0; hb = h0; }
      if (found >= 0) {  // exactly one thread
        s_prefix = prefix | ((uint64_t)found << sh);
        s_need = need - found_cum;
        if (hb == need - found_cum || sh == 0) s_done = 1;  // whole bin taken / last digit
      }
    }
    __syncthreads();
  }
  const uint64_t thr = s_prefix;  // every key >= thr is a winner (exactly k of them; keys unique)
  for (int i = tid; i < k_pow2; i += blockDim.x) s_sel[i] = 0ull;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const uint64_t key = key_at(i);
    if (key >= thr) {
      const int p = atomicAdd(&s_cnt, 1);
      if (p < k_pow2) s_sel[p] = key;
    }
  }
  __syncthreads();
  if (kSort) block_bitonic_sort_desc(s_sel, k_pow2);
}

inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace sdet
