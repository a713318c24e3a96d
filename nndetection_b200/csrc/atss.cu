// ATSS anchor matching for a whole batch in four launches, never materialising the reference's
// [G, A] IoU / -INF matrices on the hot path.
//
// Reference: ATSSMatcher.compute_matches, nndet/core/boxes/matcher/atss.py:48-122 (center_in_gt = False,
// v001.yaml:105-107); Matcher.__call__ no-GT shortcut, matcher/base.py:51-56; labels from
// BaseRetinaNet.assign_targets_to_anchors, nndet/core/retina.py:256-288.
//   1. centre distance of every (gt, anchor) pair                       (atss.py:77)
//   2. per (gt, level): k = min(num_candidates * anchors_per_loc, A_l) nearest anchors, canonical tie-break
//      (distance, then ascending anchor index)                          (atss.py:82-90)
//   3. per gt: IoU of its candidates, mean + unbiased std -> threshold, positives = IoU >= thr (atss.py:94-101);
//      positives race for their anchor with a packed 64-bit atomicMax (highest IoU, then lowest gt index,
//      = overlaps_inf.max(dim=0), atss.py:119)
//   4. winners decode the packed value into the per-image gt index; everything else stays -1.
#include "common.cuh"
#include "select.cuh"

namespace {

struct Box6 { float x1, y1, x2, y2, z1, z2; };
__device__ __forceinline__ Box6 ldbox(const float* p) {
  const float2* q = reinterpret_cast<const float2*>(p);
  float2 a = q[0], b = q[1], c = q[2];
  return {a.x, a.y, b.x, b.y, c.x, c.y};
}

__global__ void fill_i64_kernel(long long* p, long long v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// grid (ceil(A/256), G)
__global__ void atss_dist_kernel(const float* __restrict__ gt, const float* __restrict__ anchors, int A,
                                 float* __restrict__ dist) {
  const int g = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  Box6 b = ldbox(gt + (size_t)g * 6), c = ldbox(anchors + (size_t)a * 6);
  float bx = __fdiv_rn(b.x2 + b.x1, 2.f), by = __fdiv_rn(b.y2 + b.y1, 2.f), bz = __fdiv_rn(b.z2 + b.z1, 2.f);
  float cx = __fdiv_rn(c.x2 + c.x1, 2.f), cy = __fdiv_rn(c.y2 + c.y1, 2.f), cz = __fdiv_rn(c.z2 + c.z1, 2.f);
  float dx = bx - cx, dy = by - cy, dz = bz - cz;
  float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  dist[(size_t)g * A + a] = __fsqrt_rn(s);
}

// grid (L, G), block 1024
__global__ void __launch_bounds__(1024)
atss_topk_kernel(const float* __restrict__ dist, int A, const int* __restrict__ level_off, int L, int kc, int ktot,
                 int* __restrict__ cand) {
  __shared__ unsigned int hist[256];
  __shared__ int ctl[4];
  __shared__ int s_cnt;
  const int l = blockIdx.x, g = blockIdx.y;
  const int start = level_off[l];
  const int n = level_off[l + 1] - start;
  int coff = 0;
  for (int i = 0; i < l; ++i) coff += min(kc, level_off[i + 1] - level_off[i]);
  const int k = min(kc, n);
  const float* d = dist + (size_t)g * A + start;
  int* out = cand + (size_t)g * ktot + coff;
  if (k <= 0) return;
  if (k >= n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = start + i;
    return;
  }
  auto key = [&](int i) -> unsigned long long {
    return ((unsigned long long)__float_as_uint(d[i]) << 32) | (unsigned int)i;     // dist >= 0: bit order == value order
  };
  SelThreshold t = block_select_smallest(key, n, k, hist, ctl);
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if ((key(i) >> t.shift) <= t.T) {
      int p = atomicAdd(&s_cnt, 1);
      if (p < k) out[p] = start + i;
    }
  }
}

__device__ __forceinline__ float iou_noeps(const Box6& a, const Box6& b) {
  float va = (a.x2 - a.x1) * (a.y2 - a.y1) * (a.z2 - a.z1);
  float vb = (b.x2 - b.x1) * (b.y2 - b.y1) * (b.z2 - b.z1);
  float dx = fmaxf(fminf(a.x2, b.x2) - fmaxf(a.x1, b.x1), 0.f);
  float dy = fmaxf(fminf(a.y2, b.y2) - fmaxf(a.y1, b.y1), 0.f);
  float dz = fmaxf(fminf(a.z2, b.z2) - fmaxf(a.z1, b.z1), 0.f);
  float inter = __fadd_rn(__fmul_rn(__fmul_rn(dx, dy), dz), 0.f);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(va, vb), inter));
}

__device__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    r = warp_sum_d(r);
    if (threadIdx.x == 0) sh[0] = r;
  }
  __syncthreads();
  r = sh[0];
  __syncthreads();
  return r;
}

constexpr unsigned long long PACK_FLAG = 1ull << 62;

// grid G, block 256
__global__ void __launch_bounds__(256)
atss_stats_kernel(const float* __restrict__ gt, const int* __restrict__ gt_img, const int* __restrict__ gt_local,
                  const float* __restrict__ anchors, int A, const int* __restrict__ cand, int ktot,
                  float* __restrict__ ciou, float* __restrict__ thr_out, long long* __restrict__ matches) {
  __shared__ double sh[32];
  const int g = blockIdx.x;
  Box6 b = ldbox(gt + (size_t)g * 6);
  const int* c = cand + (size_t)g * ktot;
  float* ci = ciou + (size_t)g * ktot;
  double s = 0.0;
  for (int j = threadIdx.x; j < ktot; j += blockDim.x) {
    float v = iou_noeps(b, ldbox(anchors + (size_t)c[j] * 6));
    ci[j] = v;
    s += (double)v;
  }
  const double mean = block_sum_d(s, sh) / (double)ktot;
  double q = 0.0;
  for (int j = threadIdx.x; j < ktot; j += blockDim.x) {
    double dv = (double)ci[j] - mean;
    q += dv * dv;
  }
  const double var = block_sum_d(q, sh) / (double)(ktot - 1);       // unbiased (atss.py:99); ktot == 1 -> NaN
  const float thr = (float)(mean + sqrt(var));
  if (threadIdx.x == 0) thr_out[g] = thr;
  const long long base = (long long)gt_img[g] * A;
  const unsigned long long lowbits = 0x7FFFFFFFull - (unsigned long long)gt_local[g];
  for (int j = threadIdx.x; j < ktot; j += blockDim.x) {
    float v = ci[j];
    if (v >= thr) {
      unsigned long long packed = PACK_FLAG | ((unsigned long long)__float_as_uint(v) << 31) | lowbits;
      atomicMax(reinterpret_cast<long long*>(matches + base + c[j]), (long long)packed);
    }
  }
}

__global__ void __launch_bounds__(256)
atss_finalize_kernel(const int* __restrict__ gt_img, int A, const int* __restrict__ cand, int ktot,
                     const float* __restrict__ ciou, const float* __restrict__ thr, long long* __restrict__ matches) {
  const int g = blockIdx.x;
  const float t = thr[g];
  const long long base = (long long)gt_img[g] * A;
  for (int j = threadIdx.x; j < ktot; j += blockDim.x) {
    if (ciou[(size_t)g * ktot + j] >= t) {
      long long* slot = matches + base + cand[(size_t)g * ktot + j];
      unsigned long long v = (unsigned long long)*slot;
      if (v & PACK_FLAG) *slot = (long long)(0x7FFFFFFFull - (v & 0x7FFFFFFFull));
    }
  }
}

// labels / matched boxes, retina.py:256-288.  labels: -1 ignore (unused by ATSS), 0 background, c+1 foreground.
__global__ void assign_labels_kernel(const long long* __restrict__ matches, long long n, long long A,
                                     const long long* __restrict__ gt_classes, const int* __restrict__ gt_off,
                                     float* __restrict__ labels) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long m = matches[i];
  float lab = 0.f;
  if (m >= 0) lab = (float)gt_classes[gt_off[i / A] + m] + 1.f;
  else if (m == -2) lab = -1.f;
  labels[i] = lab;
}

int ktot_of(const int* level_off_host, int L, int kc) {
  int k = 0;
  for (int l = 0; l < L; ++l) k += (level_off_host[l + 1] - level_off_host[l] < kc) ? level_off_host[l + 1] - level_off_host[l] : kc;
  return k;
}

}  // namespace

extern "C" {

size_t nnd_atss_workspace_bytes(int G, int A, const int* level_off_host, int L, int kc) {
  if (G <= 0) return 256;
  int kt = ktot_of(level_off_host, L, kc);
  return nnd_align_up((size_t)G * A * 4) + 2 * nnd_align_up((size_t)G * kt * 4) + nnd_align_up((size_t)G * 4) +
         nnd_align_up((size_t)(L + 1) * 4) + 256;
}

// gt [G,6] all images concatenated; gt_img[g] image of box g; gt_local[g] index of g inside its image.
// anchors [A,6] shared by all images.  level_off_dev / level_off_host: [L+1] anchor offsets of the pyramid levels.
// kc = num_candidates * num_anchors_per_loc.  matches_out [B*A] int64: per-image gt index or -1.
int nnd_atss_match(const float* gt, const int* gt_img, const int* gt_local, int G, const float* anchors, int A, int B,
                   const int* level_off_dev, const int* level_off_host, int L, int kc, long long* matches_out,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  if (G < 0 || A <= 0 || B <= 0 || L <= 0 || kc <= 0 || !matches_out) return NND_ERR_ARG;
  const long long n = (long long)A * B;
  fill_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(matches_out, -1ll, n);     // matcher/base.py:13
  NND_LAUNCH_CHECK("fill_i64_kernel");
  if (G == 0) return NND_OK;
  if (!gt || !gt_img || !gt_local || !anchors || !level_off_dev || !ws) return NND_ERR_ARG;
  if (nnd_atss_workspace_bytes(G, A, level_off_host, L, kc) > ws_bytes) return NND_ERR_WORKSPACE;
  const int kt = ktot_of(level_off_host, L, kc);
  char* p = reinterpret_cast<char*>(ws);
  float* dist = nnd_carve<float>(p, (size_t)G * A);
  int* cand = nnd_carve<int>(p, (size_t)G * kt);
  float* ciou = nnd_carve<float>(p, (size_t)G * kt);
  float* thr = nnd_carve<float>(p, G);
  atss_dist_kernel<<<dim3((A + 255) / 256, G), 256, 0, st>>>(gt, anchors, A, dist);
  NND_LAUNCH_CHECK("atss_dist_kernel");
  atss_topk_kernel<<<dim3(L, G), 1024, 0, st>>>(dist, A, level_off_dev, L, kc, kt, cand);
  NND_LAUNCH_CHECK("atss_topk_kernel");
  atss_stats_kernel<<<G, 256, 0, st>>>(gt, gt_img, gt_local, anchors, A, cand, kt, ciou, thr, matches_out);
  NND_LAUNCH_CHECK("atss_stats_kernel");
  atss_finalize_kernel<<<G, 256, 0, st>>>(gt_img, A, cand, kt, ciou, thr, matches_out);
  NND_LAUNCH_CHECK("atss_finalize_kernel");
  return NND_OK;
}

// labels[i] for the concatenated batch (retina.py:256-288); gt_off[b] = first gt of image b in gt_classes.
int nnd_assign_labels(const long long* matches, long long n, long long A, const long long* gt_classes,
                      const int* gt_off, float* labels_out, cudaStream_t st) {
  if (n < 0 || A <= 0) return NND_ERR_ARG;
  if (n == 0) return NND_OK;
  assign_labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(matches, n, A, gt_classes, gt_off, labels_out);
  NND_LAUNCH_CHECK("assign_labels_kernel");
  return NND_OK;
}

}  // extern "C"
