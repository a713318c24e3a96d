// 3-D / 2-D greedy NMS, fully on device (no host round trip).
//
// Replaces nndet/csrc/cuda/nms.cu:148-221 (nms_cuda), :99-145 (nms_kernel_3d), :36-51 (devIoU_3d) of the
// reference.  Semantics kept bit-for-bit:
//   * boxes sorted by descending score (stable: equal scores keep ascending index order),
//   * IoU = inter / (Sa + Sb - inter) in fp32, IEEE division, no eps,
//   * box j is removed by a kept box i (i before j) iff IoU > thr (strict) -> NaN never suppresses,
//   * result = indices into the ORIGINAL order, by descending score.
// Differences in mechanism (B200-first):
//   * only the upper triangle of the N x N/64 bitmask is computed, 4 column tiles per CTA so every thread
//     writes one full 32-byte sector,
//   * pairs with an empty intersection skip the division (exactly equivalent, see pair_suppresses()),
//   * the greedy reduction runs on the device in one persistent CTA (remv[] in shared memory, 64-box blocks
//     resolved by a register chain, kept rows OR-ed in by 1024 threads) instead of a D2H copy of the whole
//     mask followed by a sequential host scan (nms.cu:193-215).
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {

constexpr int TILE = 64;
constexpr int COL_TILES_PER_CTA = 4;

template <int DIM>
struct BoxT;
template <>
struct BoxT<3> { float x1, y1, x2, y2, z1, z2; };
template <>
struct BoxT<2> { float x1, y1, x2, y2; };

__global__ void iota_kernel(int* idx, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = i;
}

// sorted_boxes[i] = boxes[order[i]], vol[i] = volume; one thread per box, float2 vector loads.
template <int DIM>
__global__ void gather_boxes_kernel(const float* __restrict__ boxes, const int* __restrict__ order,
                                    float* __restrict__ sorted, float* __restrict__ vol, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2* src = reinterpret_cast<const float2*>(boxes + (size_t)order[i] * (2 * DIM));
  float2* dst = reinterpret_cast<float2*>(sorted + (size_t)i * (2 * DIM));
  float2 a = src[0], b = src[1];
  dst[0] = a; dst[1] = b;
  float v = (b.x - a.x) * (b.y - a.y);          // (x2-x1)*(y2-y1)
  if (DIM == 3) {
    float2 c = src[2];
    dst[2] = c;
    v = v * (c.y - c.x);                         // *(z2-z1), same association as devIoU_3d
  }
  vol[i] = v;
}

// Exactly `inter / (sa + sb - inter) > thr` of devIoU(_3d) + nms_kernel(_3d) line 138, with the division skipped
// when inter == 0: then the quotient is 0 (or NaN when sa+sb is 0 or NaN) so the test is `0 > thr` resp. false.
__device__ __forceinline__ bool pair_suppresses(float inter, float sa, float sb, float thr, bool zero_gt_thr) {
  if (inter == 0.0f) {
    float s = sa + sb;
    return zero_gt_thr && (s != 0.0f) && (s == s);
  }
  return __fdiv_rn(inter, (sa + sb - inter)) > thr;
}

template <int DIM>
__global__ void __launch_bounds__(TILE)
nms_mask_kernel(const float* __restrict__ sorted, const float* __restrict__ vol, int n, int col_blocks,
                float thr, unsigned long long* __restrict__ mask) {
  // blockIdx.y = row tile, blockIdx.x = group of COL_TILES_PER_CTA column tiles; skip groups entirely left of
  // the diagonal (the reference computes them and never reads them, nms.cu:105).
  const int row_tile = blockIdx.y;
  const int col_tile0 = blockIdx.x * COL_TILES_PER_CTA;
  if (col_tile0 + COL_TILES_PER_CTA - 1 < row_tile) return;

  __shared__ float s_box[COL_TILES_PER_CTA * TILE * 2 * DIM];
  __shared__ float s_vol[COL_TILES_PER_CTA * TILE];
  const int col0 = col_tile0 * TILE;
  const int ncols = min(n - col0, COL_TILES_PER_CTA * TILE);
  for (int i = threadIdx.x; i < ncols * 2 * DIM; i += TILE) s_box[i] = sorted[(size_t)col0 * 2 * DIM + i];
  for (int i = threadIdx.x; i < ncols; i += TILE) s_vol[i] = vol[col0 + i];
  __syncthreads();

  const int row = row_tile * TILE + threadIdx.x;
  if (row >= n) return;
  const float* rb = sorted + (size_t)row * 2 * DIM;
  const float ax1 = rb[0], ay1 = rb[1], ax2 = rb[2], ay2 = rb[3];
  float az1 = 0.f, az2 = 0.f;
  if (DIM == 3) { az1 = rb[4]; az2 = rb[5]; }
  const float sa = vol[row];
  const bool zero_gt = 0.0f > thr;

  unsigned long long words[COL_TILES_PER_CTA];
#pragma unroll
  for (int t = 0; t < COL_TILES_PER_CTA; ++t) {
    unsigned long long w = 0ull;
    const int ct = col_tile0 + t;
    if (ct >= row_tile && ct < col_blocks) {
      const int cnt = min(n - ct * TILE, TILE);
      const int start = (ct == row_tile) ? threadIdx.x + 1 : 0;
      for (int j = start; j < cnt; ++j) {
        const float* cb = s_box + (t * TILE + j) * 2 * DIM;
        float w_ = fmaxf(fminf(ax2, cb[2]) - fmaxf(ax1, cb[0]), 0.f);   // "width"  (x extent)
        float h_ = fmaxf(fminf(ay2, cb[3]) - fmaxf(ay1, cb[1]), 0.f);   // "height" (y extent)
        float inter = w_ * h_;
        if (DIM == 3) {
          float d_ = fmaxf(fminf(az2, cb[5]) - fmaxf(az1, cb[4]), 0.f);
          inter = inter * d_;
        }
        if (pair_suppresses(inter, sa, s_vol[t * TILE + j], thr, zero_gt)) w |= 1ull << j;
      }
    }
    words[t] = w;
  }
  unsigned long long* out = mask + (size_t)row * col_blocks + col_tile0;
#pragma unroll
  for (int t = 0; t < COL_TILES_PER_CTA; ++t)
    if (col_tile0 + t < col_blocks) out[t] = words[t];
}

// One persistent CTA walks the 64-box blocks in score order.  The mask rows of a block ("panel": 64 rows x the columns
// still ahead inside the current column chunk) are prefetched with cp.async two blocks ahead, INDEPENDENT of which rows
// will turn out to be kept, so the only dependent chain per block is the 64-step register resolution -- no global-
// memory latency on the critical path (the previous version paid 3-5 L2 round trips per block: 1.2 ms at N = 10 k).
// Columns are processed in chunks of SCAN_WC blocks; when a new chunk starts, the rows kept so far are OR-ed into its
// remv[] words in one bandwidth-bound sweep (only kept rows are read).
constexpr int SCAN_WC = 128;                 // column blocks per chunk (8192 boxes)
constexpr int SCAN_THREADS = 1024;

__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src));
}

// CLUSTER = true (weighted box clustering): additionally records for every box the kept box that removed it FIRST
// (head_of[row], rows in sorted order; kept boxes head themselves) and the position of every kept box in the keep list.
template <bool CLUSTER>
__global__ void __launch_bounds__(SCAN_THREADS)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order, int n, int col_blocks,
                int* __restrict__ kept_rows, long long* __restrict__ keep_out, long long* __restrict__ n_keep_out,
                int* __restrict__ head_of, int* __restrict__ pos_of, int max_keep) {
  extern __shared__ __align__(16) unsigned long long sm_scan[];
  unsigned long long* remv = sm_scan;                                   // [col_blocks]
  unsigned long long* panel = sm_scan + ((col_blocks + 1) & ~1);        // 2 stages x [64][SCAN_WC]
  __shared__ unsigned long long s_kmask;
  __shared__ unsigned long long s_new[TILE];       // CLUSTER: in-block members newly removed by kept row j
  __shared__ int s_count;
  const int tid = threadIdx.x;
  for (int c = tid; c < col_blocks; c += SCAN_THREADS) remv[c] = 0ull;
  if (tid == 0) s_count = 0;
  __syncthreads();

  auto prefetch = [&](int b, int c1) {
    // rows of block b, columns b .. c1-1, into stage b & 1
    if (b < c1) {
      const int base = b * TILE;
      const int cnt = min(n - base, TILE);
      const int width = c1 - b;
      const unsigned dst0 = (unsigned)__cvta_generic_to_shared(panel + (size_t)(b & 1) * TILE * SCAN_WC);
      // warp w copies rows w, w + 32 (lanes stride over the columns): no integer division on the issue path
      for (int j = tid >> 5; j < cnt; j += SCAN_THREADS / 32) {
        const unsigned long long* src = mask + (size_t)(base + j) * col_blocks + b;
        const unsigned dst = dst0 + (unsigned)(j * SCAN_WC) * 8u;
        for (int c = tid & 31; c < width; c += 32) cp_async8(dst + (unsigned)c * 8u, src + c);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  for (int c0 = 0; c0 < col_blocks; c0 += SCAN_WC) {
    const int c1 = min(c0 + SCAN_WC, col_blocks);
    if (c0 > 0) {
      // rows kept in earlier chunks -> remv[c0 .. c1)
      const int kept = s_count;
      const int width = c1 - c0;
      const int groups = SCAN_THREADS / SCAN_WC;                       // 8 row groups x 128 columns
      const int col = tid % SCAN_WC, grp = tid / SCAN_WC;
      if (CLUSTER) {
        // first remover wins: one thread per column walks the kept rows in keep order (loads batched by 8)
        if (grp == 0 && col < width) {
          unsigned long long acc = 0ull;
          for (int k = 0; k < kept; k += 8) {
            unsigned long long v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = k + u < kept ? __ldg(mask + (size_t)kept_rows[k + u] * col_blocks + c0 + col) : 0ull;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              unsigned long long nw = v[u] & ~acc;
              acc |= v[u];
              while (nw) {
                const int bit = __ffsll((long long)nw) - 1;
                nw &= nw - 1;
                head_of[(c0 + col) * TILE + bit] = kept_rows[k + u];
              }
            }
          }
          remv[c0 + col] = acc;
        }
      } else if (col < width) {
        unsigned long long acc = 0ull;
        for (int k = grp; k < kept; k += groups * 8) {
          unsigned long long v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int kk = k + u * groups;
            v[u] = kk < kept ? __ldg(mask + (size_t)kept_rows[kk] * col_blocks + c0 + col) : 0ull;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) acc |= v[u];
        }
        if (acc) atomicOr(&remv[c0 + col], acc);
      }
      __syncthreads();
    }
    prefetch(c0, c1);
    prefetch(c0 + 1, c1);
    for (int b = c0; b < c1; ++b) {
      const int base = b * TILE;
      const int cnt = min(n - base, TILE);
      asm volatile("cp.async.wait_group 1;\n" ::);
      __syncthreads();
      if (s_count >= max_keep) break;                // the caller only wants the first max_keep survivors (score order)
      const unsigned long long* pn = panel + (size_t)(b & 1) * TILE * SCAN_WC;
      if (tid == 0) {
        // greedy resolution inside the block: only boxes that survive are visited (each visit = one dependent shared-
        // memory read of that box's diagonal word), not all 64 positions
        const unsigned long long valid = cnt == TILE ? ~0ull : ((1ull << cnt) - 1ull);
        unsigned long long rem = remv[b], km = 0ull;
        unsigned long long avail = ~rem & valid;
        while (avail) {
          const int j = __ffsll((long long)avail) - 1;
          km |= 1ull << j;
          const unsigned long long dg = pn[j * SCAN_WC];
          if (CLUSTER) s_new[j] = dg & ~rem & valid;
          rem |= dg | (1ull << j);
          avail = ~rem & valid;
        }
        s_kmask = km;
      }
      __syncthreads();
      const unsigned long long km = s_kmask;
      const int count0 = s_count;
      if (tid < TILE && ((km >> tid) & 1ull)) {
        const int pos = count0 + __popcll(km & ((1ull << tid) - 1ull));
        keep_out[pos] = (long long)order[base + tid];
        kept_rows[pos] = base + tid;
        if (CLUSTER) { head_of[base + tid] = base + tid; pos_of[base + tid] = pos; }
      }
      if (CLUSTER && tid < cnt && !((km >> tid) & 1ull)) {
        unsigned long long m = km & ((1ull << tid) - 1ull);         // kept rows before this box, ascending
        while (m) {
          const int j = __ffsll((long long)m) - 1;
          m &= m - 1;
          if ((s_new[j] >> tid) & 1ull) { head_of[base + tid] = base + j; break; }
        }
      }
      // OR the kept rows of this block into remv[c], b < c < c1: 8 row groups x 128 columns, combined with shared atomics
      {
        const int col = tid % SCAN_WC, grp = tid / SCAN_WC;
        if (CLUSTER) {
          if (grp == 0 && col >= 1 && b + col < c1) {
            unsigned long long acc = remv[b + col], m = km;
            while (m) {
              const int j = __ffsll((long long)m) - 1;
              m &= m - 1;
              const unsigned long long row = pn[j * SCAN_WC + col];
              unsigned long long nw = row & ~acc;
              acc |= row;
              while (nw) {
                const int bit = __ffsll((long long)nw) - 1;
                nw &= nw - 1;
                head_of[(b + col) * TILE + bit] = base + j;
              }
            }
            remv[b + col] = acc;
          }
        } else if (col >= 1 && b + col < c1) {
          unsigned long long acc = 0ull;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = grp * 8 + u;
            if ((km >> j) & 1ull) acc |= pn[j * SCAN_WC + col];
          }
          if (acc) atomicOr(&remv[b + col], acc);
        }
      }
      __syncthreads();
      if (tid == 0) s_count = count0 + __popcll(km);
      prefetch(b + 2, c1);                           // stage (b & 1) is free again
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    if (s_count >= max_keep) break;
  }
  __syncthreads();
  if (tid == 0) *n_keep_out = (long long)s_count;
}

struct NmsWs {
  int* idx_in; int* idx_out; float* keys_out; float* sorted; float* vol; unsigned long long* mask;
  void* cub_tmp; size_t cub_bytes; size_t total;
};

size_t cub_sort_bytes(int n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const float*)nullptr, (float*)nullptr,
                                            (const int*)nullptr, (int*)nullptr, n);
  return bytes;
}

NmsWs carve_ws(void* ws, long long n, int dim) {
  NmsWs w;
  char* p = reinterpret_cast<char*>(ws);
  char* p0 = p;
  const long long cb = (n + TILE - 1) / TILE;
  w.idx_in = nnd_carve<int>(p, n);
  w.idx_out = nnd_carve<int>(p, n);
  w.keys_out = nnd_carve<float>(p, n);
  w.sorted = nnd_carve<float>(p, n * 2 * dim);
  w.vol = nnd_carve<float>(p, n);
  w.mask = nnd_carve<unsigned long long>(p, (size_t)n * cb);
  w.cub_bytes = cub_sort_bytes((int)n);
  w.cub_tmp = p;
  p += nnd_align_up(w.cub_bytes);
  w.total = (size_t)(p - p0);
  return w;
}

struct ClusterOut { int* head_of; int* pos_of; };

template <int DIM>
int nms_impl(const float* boxes, const float* scores, long long n, float thr, long long* keep_out,
             long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream, const ClusterOut* cl = nullptr,
             long long max_keep = -1) {
  if (n < 0 || !n_keep_out) return NND_ERR_ARG;
  if (n == 0) {
    NND_CUDA_TRY(cudaMemsetAsync(n_keep_out, 0, sizeof(long long), stream));
    return NND_OK;
  }
  if (!boxes || !scores || !keep_out || !ws) return NND_ERR_ARG;
  if (n > (1ll << 24)) return NND_ERR_ARG;
  NmsWs w = carve_ws(ws, n, DIM);
  if (w.total > ws_bytes) return NND_ERR_WORKSPACE;
  const int ni = (int)n;
  const int col_blocks = (ni + TILE - 1) / TILE;
  const size_t scan_smem = ((size_t)((col_blocks + 1) & ~1) + (size_t)2 * TILE * SCAN_WC) * sizeof(unsigned long long);
  if (scan_smem > 220 * 1024) return NND_ERR_ARG;

  iota_kernel<<<(ni + 255) / 256, 256, 0, stream>>>(w.idx_in, ni);
  NND_LAUNCH_CHECK("iota_kernel");
  size_t cb = w.cub_bytes;
  NND_CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(w.cub_tmp, cb, scores, w.keys_out, w.idx_in, w.idx_out,
                                                         ni, 0, 32, stream));
  gather_boxes_kernel<DIM><<<(ni + 255) / 256, 256, 0, stream>>>(boxes, w.idx_out, w.sorted, w.vol, ni);
  NND_LAUNCH_CHECK("gather_boxes_kernel");
  dim3 grid((col_blocks + COL_TILES_PER_CTA - 1) / COL_TILES_PER_CTA, col_blocks);
  nms_mask_kernel<DIM><<<grid, TILE, 0, stream>>>(w.sorted, w.vol, ni, col_blocks, thr, w.mask);
  NND_LAUNCH_CHECK("nms_mask_kernel");
  static NndPerDeviceOnce scan_attr;
  if (scan_attr.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(nms_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    NND_CUDA_TRY(cudaFuncSetAttribute(nms_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  }
  // idx_in (the iota input of the sort) is dead by now: reuse it for the sorted-row numbers of the kept boxes
  const int mk = (max_keep < 0 || max_keep > n) ? ni : (int)max_keep;
  if (cl) nms_scan_kernel<true><<<1, SCAN_THREADS, scan_smem, stream>>>(w.mask, w.idx_out, ni, col_blocks, w.idx_in, keep_out, n_keep_out, cl->head_of, cl->pos_of, ni);
  else nms_scan_kernel<false><<<1, SCAN_THREADS, scan_smem, stream>>>(w.mask, w.idx_out, ni, col_blocks, w.idx_in, keep_out, n_keep_out, nullptr, nullptr, mk);
  NND_LAUNCH_CHECK("nms_scan_kernel");
  return NND_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Weighted box clustering (nndet/inference/detection/wbc.py:94-198).  Clusters = the suppression sets of the greedy scan
// above (head = kept box, members = boxes it removed first, + itself when IoU(h, h) > thr).  One thread per box adds its
// terms to its cluster's accumulators; one block finalises scores / boxes and compacts clusters with score > threshold.
constexpr int WBC_ACC = 10;      // sum iou*w | sum iou*w*s | sum iou*w*s*box[0..5] | count | sum n_exp

__global__ void wbc_reduce_kernel(const float* __restrict__ sorted, const float* __restrict__ vol, const int* __restrict__ order,
                                  const int* __restrict__ head_of, const int* __restrict__ pos_of,
                                  const float* __restrict__ scores, const float* __restrict__ weights,
                                  const float* __restrict__ n_exp, int n, float thr, int use_area, float* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int h = head_of[i];
  const float* a = sorted + (size_t)h * 6;
  const float* b = sorted + (size_t)i * 6;
  // box_iou_union_3d (nndet/core/boxes/ops.py:131-159), eps = 0, same association as the mask kernel
  const float w_ = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.f);
  const float h_ = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.f);
  const float d_ = fmaxf(fminf(a[5], b[5]) - fmaxf(a[4], b[4]), 0.f);
  const float inter = w_ * h_ * d_;
  const float iou = __fdiv_rn(inter, vol[h] + vol[i] - inter);
  if (i == h && !(iou > thr)) return;            // a head joins its own cluster only if IoU(h, h) > thr (NaN: zero volume)
  const int o = order[i];
  const float wt = use_area ? weights[o] * vol[i] : weights[o];
  const float msw = iou * wt, ms = msw * scores[o];
  float* ac = acc + (size_t)pos_of[h] * WBC_ACC;
  atomicAdd(ac + 0, msw);
  atomicAdd(ac + 1, ms);
#pragma unroll
  for (int k = 0; k < 6; ++k) atomicAdd(ac + 2 + k, b[k] * ms);
  atomicAdd(ac + 8, 1.f);
  atomicAdd(ac + 9, n_exp[o]);
}

__global__ void __launch_bounds__(1024)
wbc_finalize_kernel(const float* __restrict__ acc, const long long* __restrict__ n_clusters, float score_thresh,
                    float missing_weight, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                    long long* __restrict__ n_out) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int K = (int)*n_clusters;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int p0 = 0; p0 < K; p0 += 1024) {
    const int p = p0 + tid;
    bool keep = false;
    float sc = 0.f, bx[6];
    if (p < K) {
      const float* ac = acc + (size_t)p * WBC_ACC;
      const float cnt = ac[8];
      if (cnt > 0.f) {
        const float n_missing = fmaxf(0.f, ac[9] / cnt - cnt);               // wbc.py:188-189
        const float denom = ac[0] + n_missing * (ac[0] / cnt) * missing_weight;
        sc = ac[1] / denom;
        keep = sc > score_thresh;
#pragma unroll
        for (int k = 0; k < 6; ++k) bx[k] = ac[2 + k] / ac[1];
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (keep) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      out_scores[pos] = sc;
#pragma unroll
      for (int k = 0; k < 6; ++k) out_boxes[(size_t)pos * 6 + k] = bx[k];
    }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
  if (tid == 0) *n_out = (long long)s_base;
}

struct WbcWs { void* nms; size_t nms_bytes; int* head_of; int* pos_of; float* acc; long long* keep; long long* n_keep; size_t total; };

WbcWs carve_wbc(void* ws, long long n) {
  WbcWs w;
  char* p = reinterpret_cast<char*>(ws);
  char* p0 = p;
  w.nms_bytes = carve_ws(nullptr, n, 3).total;
  w.nms = p; p += nnd_align_up(w.nms_bytes);
  w.head_of = nnd_carve<int>(p, n);
  w.pos_of = nnd_carve<int>(p, n);
  w.acc = nnd_carve<float>(p, n * WBC_ACC);
  w.keep = nnd_carve<long long>(p, n);
  w.n_keep = nnd_carve<long long>(p, 1);
  w.total = (size_t)(p - p0);
  return w;
}

}  // namespace

extern "C" {

size_t nnd_wbc_workspace_bytes(long long n) {
  if (n <= 0) return 256;
  return carve_wbc(nullptr, n).total;
}

int nnd_wbc3d_f32(const float* boxes, const float* scores, const float* weights, const float* n_exp, long long n,
                  float iou_thresh, float score_thresh, int use_area, float missing_weight, float* out_boxes,
                  float* out_scores, long long* n_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n < 0 || !n_out) return NND_ERR_ARG;
  if (n == 0) {
    NND_CUDA_TRY(cudaMemsetAsync(n_out, 0, sizeof(long long), stream));
    return NND_OK;
  }
  if (!boxes || !scores || !weights || !n_exp || !out_boxes || !out_scores || !ws) return NND_ERR_ARG;
  WbcWs w = carve_wbc(ws, n);
  if (w.total > ws_bytes) return NND_ERR_WORKSPACE;
  const ClusterOut cl{w.head_of, w.pos_of};
  const int rc = nms_impl<3>(boxes, scores, n, iou_thresh, w.keep, w.n_keep, w.nms, w.nms_bytes, stream, &cl);
  if (rc != NND_OK) return rc;
  NmsWs nw = carve_ws(w.nms, n, 3);
  NND_CUDA_TRY(cudaMemsetAsync(w.acc, 0, sizeof(float) * n * WBC_ACC, stream));
  const int ni = (int)n;
  wbc_reduce_kernel<<<(ni + 255) / 256, 256, 0, stream>>>(nw.sorted, nw.vol, nw.idx_out, w.head_of, w.pos_of, scores, weights, n_exp,
                                                         ni, iou_thresh, use_area, w.acc);
  NND_LAUNCH_CHECK("wbc_reduce_kernel");
  wbc_finalize_kernel<<<1, 1024, 0, stream>>>(w.acc, w.n_keep, score_thresh, missing_weight, out_boxes, out_scores, n_out);
  NND_LAUNCH_CHECK("wbc_finalize_kernel");
  return NND_OK;
}

size_t nnd_nms_workspace_bytes(long long n, int dim) {
  if (n <= 0) return 256;
  return carve_ws(nullptr, n, dim == 2 ? 2 : 3).total;
}

int nnd_nms3d_f32(const float* boxes, const float* scores, long long n, float iou_threshold, long long* keep_out,
                  long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return nms_impl<3>(boxes, scores, n, iou_threshold, keep_out, n_keep_out, ws, ws_bytes, stream);
}

int nnd_nms2d_f32(const float* boxes, const float* scores, long long n, float iou_threshold, long long* keep_out,
                  long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return nms_impl<2>(boxes, scores, n, iou_threshold, keep_out, n_keep_out, ws, ws_bytes, stream);
}

// Same as nnd_nms3d_f32 when only the first `max_keep` survivors (in score order) are consumed -- `keep[:detections_per_img]`
// after batched_nms, nndet/core/retina.py:376-378: the greedy scan stops at the end of the 64-box block in which the
// max_keep-th survivor was found; keep_out[0 .. min(*n_keep_out, max_keep)) is identical to the full result's prefix.
int nnd_nms3d_topk_f32(const float* boxes, const float* scores, long long n, float iou_threshold, long long max_keep,
                       long long* keep_out, long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return nms_impl<3>(boxes, scores, n, iou_threshold, keep_out, n_keep_out, ws, ws_bytes, stream, nullptr, max_keep);
}

}  // extern "C"
