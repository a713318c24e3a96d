// Hard-negative sampling and the detection-head losses, device resident (no torch.where / .item() syncs).
//
// Reference:
//   HardNegativeSamplerBatched.__call__     nndet/core/boxes/sampler.py:212-270
//   get_num_pos / get_num_neg               sampler.py:154-185
//   select_negatives (pool = top fg prob)   sampler.py:67-98
//   DetectionHeadHNMNative.compute_loss     nndet/arch/heads/comb.py:352-405
//   decode_single                           nndet/core/boxes/coder.py:90-155
//   GIoULoss / generalized_box_iou_3d       nndet/losses/regression.py:118-162, nndet/core/boxes/ops.py:162-185
//   BCEWithLogitsLossOneHot                 nndet/losses/classification.py:137-181
// Random draws: the reference uses torch.randperm (device RNG stream, not reproducible here); this path ranks
// candidates by a counter hash of (anchor index, seed) -- see oracle/box_oracle.py:hash_priority.
#include "common.cuh"
#include "select.cuh"

namespace {

constexpr int MAX_SEL = 4096;     // upper bound on num_pos / num_neg handled by the single-CTA pick kernel

__global__ void hnm_count_kernel(const float* __restrict__ labels, long long n, int* __restrict__ counts,
                                 int* __restrict__ pos_list, int pos_cap) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float l = (i < n) ? labels[i] : -1.f;
  bool is_pos = l >= 1.f, is_neg = l == 0.f;
  unsigned pm = __ballot_sync(0xffffffffu, is_pos), nm = __ballot_sync(0xffffffffu, is_neg);
  const int lane = threadIdx.x & 31;
  if (pm) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&counts[0], __popc(pm));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (is_pos) {
      int p = base + __popc(pm & ((1u << lane) - 1u));
      if (p < pos_cap) pos_list[p] = (int)i; else counts[6] = 1;
    }
  }
  if (nm && lane == 0) atomicAdd(&counts[1], __popc(nm));
}

__global__ void hnm_plan_kernel(int* __restrict__ counts, int max_pos, double neg_ratio, int min_neg, double pool_size,
                                SelState* __restrict__ st, unsigned int* __restrict__ hist) {
  if (threadIdx.x < 256) hist[threadIdx.x] = 0;
  if (threadIdx.x != 0) return;
  const int P = counts[0], N = counts[1];
  int num_pos = min(P, max_pos);                                        // sampler.py:164-168
  int num_neg = (int)((double)max(1, num_pos) * neg_ratio);             // sampler.py:182
  num_neg = min(N, max(num_neg, min_neg));                              // sampler.py:184
  int pool = min(N, (int)((double)num_neg * pool_size));                // sampler.py:85-86
  counts[2] = num_pos; counts[3] = num_neg; counts[4] = pool; counts[5] = 0;
  st->prefix = 0ull; st->need = pool; st->shift = 56;
  st->done = (pool <= 0 || pool >= N) ? 1 : 0;
  st->T = (pool <= 0) ? 0ull : ~0ull;       // pool >= N: take every negative; pool == 0: collect kernel skips
}

__global__ void pool_collect_kernel(const float* __restrict__ labels, const float* __restrict__ probs, long long n,
                                    const SelState* __restrict__ st, int* __restrict__ counts,
                                    int* __restrict__ pool_list) {
  const int pool = counts[4];
  if (pool <= 0) return;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool take = false;
  if (i < n && labels[i] == 0.f) take = (neg_key(probs[i], (unsigned int)i) >> st->shift) <= st->T;
  unsigned m = __ballot_sync(0xffffffffu, take);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(&counts[5], __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (take) {
    int p = base + __popc(m & ((1u << lane) - 1u));
    if (p < pool) pool_list[p] = (int)i;
  }
}

// bitonic sort of s[0..m) ascending, m <= MAX_SEL, padded with INT_MAX to a power of two
__device__ void block_sort_int(int* s, int m) {
  int p2 = 1;
  while (p2 < m) p2 <<= 1;
  for (int i = m + threadIdx.x; i < p2; i += blockDim.x) s[i] = 0x7FFFFFFF;
  __syncthreads();
  for (int k = 2; k <= p2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          int a = s[i], b = s[ixj];
          bool up = (i & k) == 0;
          if ((a > b) == up) { s[i] = b; s[ixj] = a; }
        }
      }
      __syncthreads();
    }
}

// choose k of list[0..n) with the smallest (hash, index) keys, write them ascending to out
__device__ void pick_by_hash(const int* __restrict__ list, int n, int k, unsigned int seed, unsigned int stream,
                             long long* __restrict__ out, int* s_sel, unsigned int* hist, int* ctl, int* s_cnt) {
  if (k <= 0) return;
  auto key = [&](int i) -> unsigned long long {
    unsigned int idx = (unsigned int)list[i];
    return ((unsigned long long)nnd_hash_priority(idx, seed, stream) << 32) | idx;
  };
  SelThreshold t{~0ull, 0};
  if (k < n) t = block_select_smallest(key, n, k, hist, ctl);
  if (threadIdx.x == 0) *s_cnt = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if ((key(i) >> t.shift) <= t.T) {
      int p = atomicAdd(s_cnt, 1);
      if (p < MAX_SEL) s_sel[p] = list[i];
    }
  __syncthreads();
  const int m = min(min(k, n), MAX_SEL);
  block_sort_int(s_sel, m);
  for (int i = threadIdx.x; i < m; i += blockDim.x) out[i] = (long long)s_sel[i];
  __syncthreads();
}

__global__ void __launch_bounds__(1024)
hnm_pick_kernel(const int* __restrict__ counts, const int* __restrict__ pos_list, int pos_cap, const int* __restrict__ pool_list,
                unsigned int seed, long long* __restrict__ pos_out, long long* __restrict__ neg_out) {
  __shared__ int s_sel[MAX_SEL];
  __shared__ unsigned int hist[256];
  __shared__ int ctl[4];
  __shared__ int s_cnt;
  // counts[0] keeps counting past the list's capacity (counts[6] flags it): never read beyond what hnm_count_kernel stored
  pick_by_hash(pos_list, min(counts[0], pos_cap), counts[2], seed, 1u, pos_out, s_sel, hist, ctl, &s_cnt);
  pick_by_hash(pool_list, counts[4], counts[3], seed, 2u, neg_out, s_sel, hist, ctl, &s_cnt);
}

// ------------------------------------------------------------------ losses
struct Box6 { float x1, y1, x2, y2, z1, z2; };
__device__ __forceinline__ Box6 ldbox(const float* p) {
  const float2* q = reinterpret_cast<const float2*>(p);
  float2 a = q[0], b = q[1], c = q[2];
  return {a.x, a.y, b.x, b.y, c.x, c.y};
}

__device__ __forceinline__ void axis_terms(float p1, float p2, float t1, float t2, float& ext, float& inter, float& hull,
                                           float& di1, float& di2, float& dh1, float& dh2) {
  ext = p2 - p1;
  float ir = fminf(p2, t2) - fmaxf(p1, t1);
  inter = fmaxf(ir, 0.f);
  float pass = ir >= 0.f ? 1.f : 0.f;
  di2 = pass * (p2 < t2 ? 1.f : (p2 == t2 ? 0.5f : 0.f));
  di1 = -pass * (p1 > t1 ? 1.f : (p1 == t1 ? 0.5f : 0.f));
  float hr = fmaxf(p2, t2) - fminf(p1, t1);
  hull = fmaxf(hr, 0.f);
  float hp = hr >= 0.f ? 1.f : 0.f;
  dh2 = hp * (p2 > t2 ? 1.f : (p2 == t2 ? 0.5f : 0.f));
  dh1 = -hp * (p1 < t1 ? 1.f : (p1 == t1 ? 0.5f : 0.f));
}

// GIoU(pred, target) and its gradient w.r.t. the six pred coordinates (x1, y1, x2, y2, z1, z2)
__device__ float giou_with_grad(const Box6& p, const Box6& t, float eps, float g[6]) {
  float ex, ey, ez, ix, iy, iz, hx, hy, hz, dix1, dix2, dhx1, dhx2, diy1, diy2, dhy1, dhy2, diz1, diz2, dhz1, dhz2;
  axis_terms(p.x1, p.x2, t.x1, t.x2, ex, ix, hx, dix1, dix2, dhx1, dhx2);
  axis_terms(p.y1, p.y2, t.y1, t.y2, ey, iy, hy, diy1, diy2, dhy1, dhy2);
  axis_terms(p.z1, p.z2, t.z1, t.z2, ez, iz, hz, diz1, diz2, dhz1, dhz2);
  const float vp = ex * ey * ez;
  const float vt = (t.x2 - t.x1) * (t.y2 - t.y1) * (t.z2 - t.z1);
  const float inter = ix * iy * iz;
  const float uni = vp + vt - inter;
  const float iou = inter / uni;
  const float hull = hx * hy * hz + eps;
  const float giou = iou - (hull - uni) / hull;
  // per coordinate: dvp, dinter, dhull
  const float dvp[6] = {-ey * ez, -ex * ez, ey * ez, ex * ez, -ex * ey, ex * ey};
  const float din[6] = {dix1 * iy * iz, diy1 * ix * iz, dix2 * iy * iz, diy2 * ix * iz, diz1 * ix * iy, diz2 * ix * iy};
  const float dhu[6] = {dhx1 * hy * hz, dhy1 * hx * hz, dhx2 * hy * hz, dhy2 * hx * hz, dhz1 * hx * hy, dhz2 * hx * hy};
  const float iu2 = 1.f / (uni * uni), ih2 = 1.f / (hull * hull);
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    float duni = dvp[c] - din[c];
    float diou = (din[c] * uni - inter * duni) * iu2;
    g[c] = diou - (uni * dhu[c] - hull * duni) * ih2;
  }
  return giou;
}

__device__ float block_sum_f(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0) sh[0] = r;
  }
  __syncthreads();
  r = sh[0];
  __syncthreads();
  return r;
}

// One CTA.  losses[0] = reg (GIoU, sum / max(1, P)), losses[1] = cls (BCE mean over (P+N)*C).
// Compact gradients (for upstream grad 1): g_deltas[P,6], g_logits[(P+N),C] in sampled order (pos then neg).
__global__ void __launch_bounds__(256)
head_loss_kernel(const float* __restrict__ logits, const float* __restrict__ deltas, const float* __restrict__ anchors,
                 long long A, int C, const long long* __restrict__ matches, const float* __restrict__ gt_boxes,
                 const int* __restrict__ gt_off, const float* __restrict__ labels, const long long* __restrict__ pos_idx,
                 const long long* __restrict__ neg_idx, const int* __restrict__ counts, float xform_clip, float eps,
                 float* __restrict__ losses, float* __restrict__ g_deltas, float* __restrict__ g_logits) {
  __shared__ float sh[32];
  const int P = counts[2], N = counts[3];
  const float inv_p = 1.f / (float)max(1, P);
  float acc = 0.f;
  for (int r = threadIdx.x; r < P; r += blockDim.x) {
    const long long i = pos_idx[r];
    const Box6 a = ldbox(anchors + (i % A) * 6);
    const float* d = deltas + i * 6;
    const float dl[6] = {d[0], d[1], d[2], d[3], d[4], d[5]};       // dx dy dw dh dz dd
    float w = a.x2 - a.x1, h = a.y2 - a.y1, dp = a.z2 - a.z1;
    float cx = a.x1 + 0.5f * w, cy = a.y1 + 0.5f * h, cz = a.z1 + 0.5f * dp;
    float pw = expf(fminf(dl[2], xform_clip)) * w, ph = expf(fminf(dl[3], xform_clip)) * h,
          pd = expf(fminf(dl[5], xform_clip)) * dp;
    float pcx = dl[0] * w + cx, pcy = dl[1] * h + cy, pcz = dl[4] * dp + cz;
    Box6 pb = {pcx - 0.5f * pw, pcy - 0.5f * ph, pcx + 0.5f * pw, pcy + 0.5f * ph, pcz - 0.5f * pd, pcz + 0.5f * pd};
    const Box6 tb = ldbox(gt_boxes + ((size_t)gt_off[i / A] + (size_t)matches[i]) * 6);
    float g[6];
    float gi = giou_with_grad(pb, tb, eps, g);
    acc += gi;
    const float s = -inv_p;                                          // d loss / d giou
    float* o = g_deltas + (size_t)r * 6;
    o[0] = s * (g[0] + g[2]) * w;
    o[1] = s * (g[1] + g[3]) * h;
    o[2] = (dl[2] <= xform_clip) ? s * 0.5f * (g[2] - g[0]) * pw : 0.f;
    o[3] = (dl[3] <= xform_clip) ? s * 0.5f * (g[3] - g[1]) * ph : 0.f;
    o[4] = s * (g[4] + g[5]) * dp;
    o[5] = (dl[5] <= xform_clip) ? s * 0.5f * (g[5] - g[4]) * pd : 0.f;
  }
  const float giou_sum = block_sum_f(acc, sh);
  const int R = P + N;
  const float inv_rc = 1.f / ((float)R * (float)C);
  float bacc = 0.f;
  for (int e = threadIdx.x; e < R * C; e += blockDim.x) {
    const int r = e / C, c = e % C;
    const long long i = r < P ? pos_idx[r] : neg_idx[r - P];
    const float x = logits[i * C + c];
    const float y = (labels[i] == (float)(c + 1)) ? 1.f : 0.f;
    bacc += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
    const float sg = 1.f / (1.f + expf(-x));
    g_logits[e] = (sg - y) * inv_rc;
  }
  const float bce_sum = block_sum_f(bacc, sh);
  if (threadIdx.x == 0) {
    losses[0] = -giou_sum * inv_p;
    losses[1] = bce_sum * inv_rc;            // R == 0 -> NaN like torch's mean of an empty tensor
  }
}

// dense gradient buffers (pre-zeroed) <- compact rows * upstream grads
__global__ void head_loss_scatter_kernel(const float* __restrict__ g_deltas, const float* __restrict__ g_logits, int C,
                                         const long long* __restrict__ pos_idx, const long long* __restrict__ neg_idx,
                                         const int* __restrict__ counts, const float* __restrict__ up_reg,
                                         const float* __restrict__ up_cls, float* __restrict__ d_deltas,
                                         float* __restrict__ d_logits) {
  const int P = counts[2], N = counts[3];
  const float ur = up_reg ? *up_reg : 1.f, uc = up_cls ? *up_cls : 1.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < P * 6; e += gridDim.x * blockDim.x)
    d_deltas[pos_idx[e / 6] * 6 + e % 6] = ur * g_deltas[e];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < (P + N) * C; e += gridDim.x * blockDim.x) {
    const int r = e / C;
    const long long i = r < P ? pos_idx[r] : neg_idx[r - P];
    d_logits[i * C + e % C] = uc * g_logits[e];
  }
}

}  // namespace

extern "C" {

int nnd_hnm_max_select(void) { return MAX_SEL; }

size_t nnd_hnm_workspace_bytes(long long n, int pos_cap, int pool_cap) {
  return nnd_align_up((size_t)pos_cap * 4) + nnd_align_up((size_t)pool_cap * 4) + nnd_align_up(sizeof(SelState)) +
         nnd_align_up(256 * 4) + 256;
}

// labels [n] (concatenated batch), fg_probs [n].  max_pos = int(batch_size_per_image * B * positive_fraction),
// neg_ratio = abs(1 - 1 / positive_fraction) (both evaluated by the host in Python float arithmetic).
// counts_out int32[8] (see hnm_count_kernel); pos_out / neg_out int64, ascending anchor indices, capacity
// >= max_pos / max possible num_neg; pool_out optional int32[pool_cap] (unordered hard-negative pool).
int nnd_hnm_sample(const float* labels, const float* fg_probs, long long n, int max_pos, double neg_ratio, int min_neg,
                   double pool_size, unsigned int seed, int* counts_out, long long* pos_out, long long* neg_out,
                   int pos_cap, int pool_cap, int** pool_list_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n <= 0 || !labels || !fg_probs || !counts_out || !pos_out || !neg_out || !ws) return NND_ERR_ARG;
  if (n > 0xFFFFFFFFll || max_pos > MAX_SEL) return NND_ERR_ARG;
  if (nnd_hnm_workspace_bytes(n, pos_cap, pool_cap) > ws_bytes) return NND_ERR_WORKSPACE;
  char* p = reinterpret_cast<char*>(ws);
  int* pos_list = nnd_carve<int>(p, pos_cap);
  int* pool_list = nnd_carve<int>(p, pool_cap);
  SelState* state = nnd_carve<SelState>(p, 1);
  unsigned int* ghist = nnd_carve<unsigned int>(p, 256);
  if (pool_list_out) *pool_list_out = pool_list;
  NND_CUDA_TRY(cudaMemsetAsync(counts_out, 0, 8 * sizeof(int), st));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  hnm_count_kernel<<<blocks, 256, 0, st>>>(labels, n, counts_out, pos_list, pos_cap);
  NND_LAUNCH_CHECK("hnm_count_kernel");
  hnm_plan_kernel<<<1, 256, 0, st>>>(counts_out, max_pos, neg_ratio, min_neg, pool_size, state, ghist);
  NND_LAUNCH_CHECK("hnm_plan_kernel");
  const int hist_blocks = (int)((n + 511) / 512 < NND_NUM_SMS * 4 ? (n + 511) / 512 : NND_NUM_SMS * 4);
  for (int round = 0; round < 8; ++round) {
    pool_hist_kernel<<<hist_blocks, 512, 0, st>>>(labels, fg_probs, n, state, ghist);
    NND_LAUNCH_CHECK("pool_hist_kernel");
    pool_pick_kernel<<<1, 256, 0, st>>>(state, ghist);
    NND_LAUNCH_CHECK("pool_pick_kernel");
  }
  pool_collect_kernel<<<blocks, 256, 0, st>>>(labels, fg_probs, n, state, counts_out, pool_list);
  NND_LAUNCH_CHECK("pool_collect_kernel");
  hnm_pick_kernel<<<1, 1024, 0, st>>>(counts_out, pos_list, pos_cap, pool_list, seed, pos_out, neg_out);
  NND_LAUNCH_CHECK("hnm_pick_kernel");
  return NND_OK;
}

// comb.py:352-405 on device.  losses_out[2] = {reg, cls}; g_deltas [max_pos,6], g_logits [(max_pos+max_neg),C].
int nnd_head_loss_fwd(const float* logits, const float* deltas, const float* anchors, long long A, int C,
                      const long long* matches, const float* gt_boxes, const int* gt_off, const float* labels,
                      const long long* pos_idx, const long long* neg_idx, const int* counts, float xform_clip,
                      float giou_eps, float* losses_out, float* g_deltas, float* g_logits, cudaStream_t st) {
  if (!logits || !deltas || !anchors || !matches || !labels || !counts || !losses_out) return NND_ERR_ARG;
  head_loss_kernel<<<1, 256, 0, st>>>(logits, deltas, anchors, A, C, matches, gt_boxes, gt_off, labels, pos_idx, neg_idx,
                                      counts, xform_clip, giou_eps, losses_out, g_deltas, g_logits);
  NND_LAUNCH_CHECK("head_loss_kernel");
  return NND_OK;
}

// d_deltas [n,6] / d_logits [n,C] must be zero-filled by the caller; up_reg / up_cls: device scalars or NULL (=1)
int nnd_head_loss_bwd(const float* g_deltas, const float* g_logits, int C, const long long* pos_idx,
                      const long long* neg_idx, const int* counts, const float* up_reg, const float* up_cls,
                      float* d_deltas, float* d_logits, cudaStream_t st) {
  head_loss_scatter_kernel<<<8, 256, 0, st>>>(g_deltas, g_logits, C, pos_idx, neg_idx, counts, up_reg, up_cls,
                                              d_deltas, d_logits);
  NND_LAUNCH_CHECK("head_loss_scatter_kernel");
  return NND_OK;
}

}  // extern "C"
