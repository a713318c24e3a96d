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

// One persistent CTA walks the 64-box blocks in score order.
__global__ void __launch_bounds__(1024)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order, int n, int col_blocks,
                long long* __restrict__ keep_out, long long* __restrict__ n_keep_out) {
  extern __shared__ unsigned long long remv[];       // [col_blocks]
  __shared__ unsigned long long s_diag[TILE];
  __shared__ unsigned long long s_kmask;
  __shared__ int s_count;
  for (int c = threadIdx.x; c < col_blocks; c += blockDim.x) remv[c] = 0ull;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();

  for (int b = 0; b < col_blocks; ++b) {
    const int base = b * TILE;
    const int cnt = min(n - base, TILE);
    if (threadIdx.x < TILE)
      s_diag[threadIdx.x] = (threadIdx.x < cnt) ? mask[(size_t)(base + threadIdx.x) * col_blocks + b] : 0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long rem = remv[b], km = 0ull;
#pragma unroll
      for (int j0 = 0; j0 < TILE; j0 += 16) {
        unsigned long long dg[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) dg[u] = s_diag[j0 + u];          // independent loads, then a register-only chain
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const bool take = (j0 + u < cnt) && !((rem >> (j0 + u)) & 1ull);
          km |= take ? (1ull << (j0 + u)) : 0ull;
          rem |= take ? dg[u] : 0ull;
        }
      }
      s_kmask = km;
    }
    __syncthreads();
    const unsigned long long km = s_kmask;
    const int count0 = s_count;
    if (threadIdx.x < TILE && ((km >> threadIdx.x) & 1ull)) {
      int pos = count0 + __popcll(km & ((1ull << threadIdx.x) - 1ull));
      keep_out[pos] = (long long)order[base + threadIdx.x];
    }
    // OR the mask rows of this block's kept boxes into remv[c], c > b.  Loads are independent: issue them in
    // batches of 16 so the L2/HBM latency is paid once per batch, not once per kept row.
    for (int c = b + 1 + threadIdx.x; c < col_blocks; c += blockDim.x) {
      unsigned long long acc = remv[c], m = km;
      const unsigned long long* col = mask + (size_t)base * col_blocks + c;
      while (m) {
        unsigned long long v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = m ? __ffsll((long long)m) - 1 : -1;
          m &= m - 1;                                    // 0 stays 0
          v[u] = j >= 0 ? __ldg(col + (size_t)j * col_blocks) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) acc |= v[u];
      }
      remv[c] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_count = count0 + __popcll(km);
  }
  __syncthreads();
  if (threadIdx.x == 0) *n_keep_out = (long long)s_count;
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

template <int DIM>
int nms_impl(const float* boxes, const float* scores, long long n, float thr, long long* keep_out,
             long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
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
  const size_t scan_smem = (size_t)col_blocks * sizeof(unsigned long long);
  if (scan_smem > 200 * 1024) return NND_ERR_ARG;

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
  if (scan_smem > 48 * 1024)
    NND_CUDA_TRY(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem));
  nms_scan_kernel<<<1, 1024, scan_smem, stream>>>(w.mask, w.idx_out, ni, col_blocks, keep_out, n_keep_out);
  NND_LAUNCH_CHECK("nms_scan_kernel");
  return NND_OK;
}

}  // namespace

extern "C" {

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

}  // extern "C"
