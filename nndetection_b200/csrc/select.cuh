// Radix selection on unique 64-bit keys (value bits in the high word, index in the low word), used for
//   * ATSS: k nearest anchor centres per (ground-truth box, pyramid level)   (matcher/atss.py:85-86)
//   * hard-negative pool: top-`pool` fg probabilities among background anchors  (sampler.py:89)
//   * inference: top-10000 scores per image                                    (retina.py:356-359)
// Keys are unique (the index is part of the key), so "k smallest" is an exact set with the canonical
// tie-break (value, then ascending index) and no tie bookkeeping is needed.  A selection is described by
// (shift, T): element taken iff (key >> shift) <= T.  Rounds stop early when a bucket is consumed whole.
#pragma once
#include "common.cuh"

struct SelThreshold {
  unsigned long long T;
  int shift;
};

__device__ __forceinline__ unsigned long long sel_hi(unsigned long long key, int shift_plus_8) {
  return shift_plus_8 >= 64 ? 0ull : (key >> shift_plus_8);
}

// warp-aggregated shared-memory histogram increment
__device__ __forceinline__ void hist_add(unsigned int* hist, int digit, bool valid) {
  unsigned int active = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  unsigned int peers = __match_any_sync(active, digit);
  if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[digit], __popc(peers));
}

// Single-CTA selection of the k smallest keys among key(i), i in [0, n).  0 < k < n required.
// `hist` is shared memory [256], `ctl` shared memory [4] ints.  All threads of the CTA must call.
template <class KeyFn>
__device__ SelThreshold block_select_smallest(KeyFn key, int n, int k, unsigned int* hist, int* ctl) {
  unsigned long long prefix = 0ull;
  int need = k;
  SelThreshold r{~0ull, 0};
  const int n_round = (n + blockDim.x - 1) / blockDim.x * blockDim.x;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int d = threadIdx.x; d < 256; d += blockDim.x) hist[d] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_round; i += blockDim.x) {
      bool valid = i < n;
      unsigned long long kk = valid ? key(i) : 0ull;
      valid = valid && (sel_hi(kk, shift + 8) == prefix);
      hist_add(hist, (int)((kk >> shift) & 255ull), valid);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, d = 0;
      for (; d < 256; ++d) {
        int c = (int)hist[d];
        if (cum + c >= need) break;
        cum += c;
      }
      ctl[0] = d;
      ctl[1] = need - cum;                                // still needed inside bucket d
      ctl[2] = ((int)hist[d] == need - cum) ? 1 : 0;      // bucket consumed whole -> done
    }
    __syncthreads();
    prefix = (prefix << 8) | (unsigned long long)ctl[0];
    need = ctl[1];
    const int done = ctl[2];
    __syncthreads();
    if (done || shift == 0) { r.T = prefix; r.shift = shift; break; }
  }
  return r;
}

// ------------------------------------------------------------------ multi-CTA variant (whole-GPU streams)
// k LARGEST values of values[i] (optionally restricted to labels[i] == 0), ties -> ascending index.
// Host drives 8 x (hist, pick) launches; kernels return at once when `done` is set (bucket consumed whole).
struct SelState {
  unsigned long long prefix;
  int need, shift, done, pad;
  unsigned long long T;
};

// counts: [0] #positive  [1] #negative  [2] num_pos  [3] num_neg  [4] pool  [5] pool filled  [6] pos_list overflow

namespace {
__device__ __forceinline__ unsigned long long neg_key(float prob, unsigned int idx) {
  // smallest key == largest probability, then smallest index (canonical tie-break)
  unsigned long long large = ((unsigned long long)__float_as_uint(prob) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
  return ~large;
}

__global__ void __launch_bounds__(512)
pool_hist_kernel(const float* __restrict__ labels, const float* __restrict__ probs, long long n,
                 const SelState* __restrict__ st, unsigned int* __restrict__ ghist) {
  if (st->done) return;
  __shared__ unsigned int hist[256];
  for (int d = threadIdx.x; d < 256; d += blockDim.x) hist[d] = 0;
  __syncthreads();
  const int shift = st->shift;
  const unsigned long long prefix = st->prefix;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_round = (n + stride - 1) / stride * stride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    bool valid = i < n && (labels == nullptr || labels[i] == 0.f);
    unsigned long long k = valid ? neg_key(probs[i], (unsigned int)i) : 0ull;
    valid = valid && sel_hi(k, shift + 8) == prefix;
    hist_add(hist, (int)((k >> shift) & 255ull), valid);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < 256; d += blockDim.x)
    if (hist[d]) atomicAdd(&ghist[d], hist[d]);
}

__global__ void pool_pick_kernel(SelState* __restrict__ st, unsigned int* __restrict__ ghist) {
  if (st->done) return;
  __shared__ unsigned int h[256];
  h[threadIdx.x] = ghist[threadIdx.x];
  ghist[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x != 0) return;
  int need = st->need, cum = 0, d = 0;
  for (; d < 256; ++d) {
    int c = (int)h[d];
    if (cum + c >= need) break;
    cum += c;
  }
  unsigned long long prefix = (st->prefix << 8) | (unsigned long long)d;
  need -= cum;
  const int shift = st->shift;
  st->prefix = prefix; st->need = need;
  if ((int)h[d] == need || shift == 0) { st->done = 1; st->T = prefix; /* st->shift stays = shift */ }
  else st->shift = shift - 8;
}


}  // namespace
