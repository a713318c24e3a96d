// Box engine, streaming kernels: anchor grid, pairwise IoU / GIoU / centre distance, decode + clip, fg-prob.
// All HBM-bound fp32 / index work: one thread per output element, coalesced stores, no tensor cores.
//
// Reference (paths under /root/reference):
//   anchor grid          nndet/core/boxes/anchors.py:337-377 (grid_anchors), :526-549 (generate_anchors)
//   box_iou / GIoU       nndet/core/boxes/ops.py:131-159, :162-185
//   box_center_dist      nndet/core/boxes/ops.py:262-287, :314-327
//   decode_single        nndet/core/boxes/coder.py:90-155 (weights all 1, clamp log(1000/16))
//   clip_boxes_3d        nndet/core/boxes/clip.py:83-101
//   fg prob              nndet/arch/heads/comb.py:262-263 (sigmoid, max over classes)
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ anchors
// out[(pos * nb + a) * 6 + c]; pos enumerates (i0, i1, i2) with i2 fastest ("ij" meshgrid, anchors.py:360-368)
__global__ void anchor_grid_kernel(float* __restrict__ out, const float* __restrict__ base, int nb,
                                   int s0, int s1, int s2, int st0, int st1, int st2) {
  const long long total = (long long)s0 * s1 * s2 * nb;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int a = (int)(idx % nb);
  long long pos = idx / nb;
  int i2 = (int)(pos % s2); pos /= s2;
  int i1 = (int)(pos % s1);
  int i0 = (int)(pos / s1);
  // arange(float) * stride, then shifts + base: two exact-in-fp32 steps like the reference
  float x = (float)i0 * (float)st0, y = (float)i1 * (float)st1, z = (float)i2 * (float)st2;
  const float* b = base + a * 6;
  float2* o = reinterpret_cast<float2*>(out + idx * 6);
  o[0] = make_float2(x + b[0], y + b[1]);
  o[1] = make_float2(x + b[2], y + b[3]);
  o[2] = make_float2(z + b[4], z + b[5]);
}

// ------------------------------------------------------------------ pairwise metrics
struct Box6 { float x1, y1, x2, y2, z1, z2; };

__device__ __forceinline__ Box6 load_box(const float* p) {
  const float2* q = reinterpret_cast<const float2*>(p);
  float2 a = q[0], b = q[1], c = q[2];
  return {a.x, a.y, b.x, b.y, c.x, c.y};
}
__device__ __forceinline__ float box_vol(const Box6& b) { return (b.x2 - b.x1) * (b.y2 - b.y1) * (b.z2 - b.z1); }

// IoU exactly as box_iou_union_3d: inter = (dx+ * dy+) * dz+ + eps; union = (v1 + v2) - inter; iou = inter / union
__device__ __forceinline__ float iou3d(const Box6& a, float va, const Box6& b, float vb, float eps, float* union_out) {
  float dx = fmaxf(fminf(a.x2, b.x2) - fmaxf(a.x1, b.x1), 0.f);
  float dy = fmaxf(fminf(a.y2, b.y2) - fmaxf(a.y1, b.y1), 0.f);
  float dz = fmaxf(fminf(a.z2, b.z2) - fmaxf(a.z1, b.z1), 0.f);
  float inter = __fadd_rn(__fmul_rn(__fmul_rn(dx, dy), dz), eps);
  float uni = __fsub_rn(__fadd_rn(va, vb), inter);
  if (union_out) *union_out = uni;
  return __fdiv_rn(inter, uni);
}

__device__ __forceinline__ float giou3d(const Box6& a, float va, const Box6& b, float vb, float eps) {
  float uni;
  float iou = iou3d(a, va, b, vb, 0.f, &uni);      // inner IoU without eps (ops.py:175)
  float hx = fmaxf(fmaxf(a.x2, b.x2) - fminf(a.x1, b.x1), 0.f);
  float hy = fmaxf(fmaxf(a.y2, b.y2) - fminf(a.y1, b.y1), 0.f);
  float hz = fmaxf(fmaxf(a.z2, b.z2) - fminf(a.z1, b.z1), 0.f);
  float hull = __fadd_rn(__fmul_rn(__fmul_rn(hx, hy), hz), eps);
  return __fsub_rn(iou, __fdiv_rn(__fsub_rn(hull, uni), hull));
}

__device__ __forceinline__ float center_dist(const Box6& a, const Box6& b) {
  float ax = __fdiv_rn(a.x2 + a.x1, 2.f), ay = __fdiv_rn(a.y2 + a.y1, 2.f), az = __fdiv_rn(a.z2 + a.z1, 2.f);
  float bx = __fdiv_rn(b.x2 + b.x1, 2.f), by = __fdiv_rn(b.y2 + b.y1, 2.f), bz = __fdiv_rn(b.z2 + b.z1, 2.f);
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  // pow(2).sum(-1).sqrt(): ((dx*dx + dy*dy) + dz*dz), no fma contraction
  float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return __fsqrt_rn(s);
}

// mode 0: IoU(eps)  1: GIoU(eps)  2: centre distance.  grid (ceil(M/256), N)
__global__ void pairwise_kernel(const float* __restrict__ b1, const float* __restrict__ b2, int n, int m, float eps,
                                int mode, float* __restrict__ out) {
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  Box6 a = load_box(b1 + (size_t)i * 6), b = load_box(b2 + (size_t)j * 6);
  float r;
  if (mode == 0) r = iou3d(a, box_vol(a), b, box_vol(b), eps, nullptr);
  else if (mode == 1) r = giou3d(a, box_vol(a), b, box_vol(b), eps);
  else r = center_dist(a, b);
  out[(size_t)i * m + j] = r;
}

// ------------------------------------------------------------------ decode (+ optional clip)
__device__ __forceinline__ void decode_axis(float lo, float hi, float dc, float ds, float clipv, float& o_lo, float& o_hi) {
  float w = hi - lo;
  float c = __fadd_rn(lo, __fmul_rn(0.5f, w));
  float ds_c = fminf(ds, clipv);                       // torch.clamp(max=clip)
  float pc = __fadd_rn(__fmul_rn(dc, w), c);
  float pw = __fmul_rn(expf(ds_c), w);
  o_lo = __fsub_rn(pc, __fmul_rn(0.5f, pw));
  o_hi = __fadd_rn(pc, __fmul_rn(0.5f, pw));
}

// boxes_out[i] = decode(deltas[i], anchors[i % A]); optionally clamped to [0, shape]
__global__ void decode_kernel(const float* __restrict__ deltas, const float* __restrict__ anchors, long long n,
                              long long A, float clipv, int do_clip, float s0, float s1, float s2,
                              float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2* d = reinterpret_cast<const float2*>(deltas + i * 6);
  float2 d0 = d[0], d1 = d[1], d2 = d[2];       // (dx, dy) (dw, dh) (dz, dd)
  Box6 a = load_box(anchors + (i % A) * 6);
  float x1, x2, y1, y2, z1, z2;
  decode_axis(a.x1, a.x2, d0.x, d1.x, clipv, x1, x2);
  decode_axis(a.y1, a.y2, d0.y, d1.y, clipv, y1, y2);
  decode_axis(a.z1, a.z2, d2.x, d2.y, clipv, z1, z2);
  if (do_clip) {
    x1 = fminf(fmaxf(x1, 0.f), s0); x2 = fminf(fmaxf(x2, 0.f), s0);
    y1 = fminf(fmaxf(y1, 0.f), s1); y2 = fminf(fmaxf(y2, 0.f), s1);
    z1 = fminf(fmaxf(z1, 0.f), s2); z2 = fminf(fmaxf(z2, 0.f), s2);
  }
  float2* o = reinterpret_cast<float2*>(out + i * 6);
  o[0] = make_float2(x1, y1); o[1] = make_float2(x2, y2); o[2] = make_float2(z1, z2);
}

// ------------------------------------------------------------------ probabilities
__device__ __forceinline__ float sigmoidf(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

// probs_out (optional) = sigmoid(logits) [n, C]; fg_out (optional) = max_c sigmoid [n]
__global__ void sigmoid_kernel(const float* __restrict__ logits, long long n, int C, float* __restrict__ probs_out,
                               float* __restrict__ fg_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float best = -1.f;
  for (int c = 0; c < C; ++c) {
    float p = sigmoidf(logits[i * C + c]);
    if (probs_out) probs_out[i * C + c] = p;
    best = fmaxf(best, p);
  }
  if (fg_out) fg_out[i] = best;
}

}  // namespace

extern "C" {

// anchors.py:337-377 for one pyramid level.  base: device [nb, 6]
int nnd_anchor_grid_f32(float* out, const float* base, int nb, const int* size3, const int* stride3, cudaStream_t st) {
  if (!out || !base || nb <= 0) return NND_ERR_ARG;
  long long total = (long long)size3[0] * size3[1] * size3[2] * nb;
  if (total == 0) return NND_OK;
  anchor_grid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, base, nb, size3[0], size3[1], size3[2],
                                                                      stride3[0], stride3[1], stride3[2]);
  NND_LAUNCH_CHECK("anchor_grid_kernel");
  return NND_OK;
}

// mode 0 box_iou (ops.py:131-159), 1 generalized_box_iou (ops.py:162-185), 2 box_center_dist (ops.py:262-287)
int nnd_box_pairwise_f32(const float* b1, const float* b2, int n, int m, float eps, int mode, float* out, cudaStream_t st) {
  if (n < 0 || m < 0 || mode < 0 || mode > 2) return NND_ERR_ARG;
  if (n == 0 || m == 0) return NND_OK;
  if (!b1 || !b2 || !out || n > 65535) return NND_ERR_ARG;
  dim3 grid((m + 255) / 256, n);
  pairwise_kernel<<<grid, 256, 0, st>>>(b1, b2, n, m, eps, mode, out);
  NND_LAUNCH_CHECK("pairwise_kernel");
  return NND_OK;
}

// coder.py:90-155 over n rows; anchors has A rows and is reused cyclically (n = batch * A); clip_shape NULL = no clip
int nnd_decode_boxes_f32(const float* deltas, const float* anchors, long long n, long long A, float xform_clip,
                         const float* clip_shape3_host, float* out, cudaStream_t st) {
  if (n < 0 || A <= 0) return NND_ERR_ARG;
  if (n == 0) return NND_OK;
  if (!deltas || !anchors || !out) return NND_ERR_ARG;
  float s0 = 0, s1 = 0, s2 = 0;
  if (clip_shape3_host) { s0 = clip_shape3_host[0]; s1 = clip_shape3_host[1]; s2 = clip_shape3_host[2]; }
  decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(deltas, anchors, n, A, xform_clip,
                                                             clip_shape3_host != nullptr, s0, s1, s2, out);
  NND_LAUNCH_CHECK("decode_kernel");
  return NND_OK;
}

int nnd_sigmoid_fg_f32(const float* logits, long long n, int C, float* probs_out, float* fg_out, cudaStream_t st) {
  if (n < 0 || C <= 0) return NND_ERR_ARG;
  if (n == 0) return NND_OK;
  sigmoid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logits, n, C, probs_out, fg_out);
  NND_LAUNCH_CHECK("sigmoid_kernel");
  return NND_OK;
}

}  // extern "C"
