// Segmentation head + loss: 1x1x1 conv (Cin -> 2) on the finest decoder map, 0.5 * CE + 0.5 * soft-Dice.
//
// Reference: DiCESegmenterFgBg (nndet/arch/heads/segmenter.py:223-290; conv_out :121-134, loss :184-203),
// SoftDiceLoss(batch_dice, no background, softmax, smooth 1e-5) (nndet/losses/segmentation.py:84-151),
// torch.nn.CrossEntropyLoss (mean over all voxels).  The reference materialises one-hot targets and tp/fp/fn
// tensors; here forward is one streaming pass over the features (logits written once, four scalars reduced),
// backward one pass producing d_features, dW and db.
#include "common.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) { float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}

// 1x1x1 conv Cin -> 2, one thread per voxel
__global__ void __launch_bounds__(256)
seg_conv_fwd_kernel(const uint4* __restrict__ x, int C8, const float* __restrict__ w, const float* __restrict__ bias,
                    long long total, float* __restrict__ logits) {
  extern __shared__ float sw[];                 // [2][C]
  const int C = C8 * 8;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  float l0 = bias[0], l1 = bias[1];
  for (int c8 = 0; c8 < C8; ++c8) {
    float f[8];
    unpack8(x[v * C8 + c8], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { l0 = fmaf(f[j], sw[c8 * 8 + j], l0); l1 = fmaf(f[j], sw[C + c8 * 8 + j], l1); }
  }
  reinterpret_cast<float2*>(logits)[v] = make_float2(l0, l1);
}

// one thread per voxel.  sums[0] = sum CE, [1] = tp, [2] = fp, [3] = fn (foreground class), double accumulators
__global__ void __launch_bounds__(256)
seg_loss_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target, long long total,
                    double* __restrict__ sums) {
  __shared__ double sred[4][8];
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double ce = 0.0, tp = 0.0, fp = 0.0, fn = 0.0;
  if (v < total) {
    const float2 lg = reinterpret_cast<const float2*>(logits)[v];
    const float l0 = lg.x, l1 = lg.y;
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float lse = mx + logf(e0 + e1);
    const float p1 = e1 / (e0 + e1);
    const bool fg = target[v] > 0.f;             // target[target > 0] = 1 (segmenter.py:288)
    ce = (double)(lse - (fg ? l1 : l0));
    if (fg) { tp = p1; fn = 1.f - p1; } else { fp = p1; }
  }
  double vals[4] = {ce, tp, fp, fn};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double r = warp_sum_d(vals[k]);
    if ((threadIdx.x & 31) == 0) sred[k][threadIdx.x >> 5] = r;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double r = 0.0;
    for (int wv = 0; wv < 8; ++wv) r += sred[threadIdx.x][wv];
    atomicAdd(&sums[threadIdx.x], r);
  }
}

__global__ void seg_loss_kernel(const double* __restrict__ sums, double total, float alpha, float smooth,
                                float* __restrict__ losses) {
  const double tp = sums[1], fp = sums[2], fn = sums[3];
  const double dc = (2.0 * tp + smooth) / (2.0 * tp + fp + fn + smooth);
  losses[0] = (float)(alpha * sums[0] / total);               // seg_ce
  losses[1] = (float)((1.0 - alpha) * (1.0 - dc));            // seg_dice (mean over the single fg class)
}

// dlogits [total, 2] of  up_ce * seg_ce + up_dice * seg_dice
__global__ void __launch_bounds__(256)
seg_loss_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target, const double* __restrict__ sums,
                    long long total, float alpha, float smooth, const float* __restrict__ up_ce,
                    const float* __restrict__ up_dice, float* __restrict__ dlogits) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const double tp = sums[1], fp = sums[2], fn = sums[3];
  const double num = 2.0 * tp + smooth, den = 2.0 * tp + fp + fn + smooth;
  const float uce = (up_ce ? *up_ce : 1.f) * alpha / (float)total;
  const float udc = (up_dice ? *up_dice : 1.f) * (1.f - alpha);
  const float2 lg = reinterpret_cast<const float2*>(logits)[v];
  const float mx = fmaxf(lg.x, lg.y);
  const float e0 = expf(lg.x - mx), e1 = expf(lg.y - mx);
  const float p1 = e1 / (e0 + e1), p0 = 1.f - p1;
  const float yv = target[v] > 0.f ? 1.f : 0.f;
  const float dldp1 = udc * (float)(-(2.0 * yv * den - num) / (den * den));   // d(1 - dc)/dp1
  const float sp = dldp1 * p1 * p0;
  reinterpret_cast<float2*>(dlogits)[v] = make_float2(uce * (p0 - (1.f - yv)) - sp, uce * (p1 - yv) + sp);
}

// thread = (voxel, 8-channel chunk).  dX = dlogits @ W;  dW += dlogits^T @ X;  db += sum dlogits
__global__ void __launch_bounds__(256)
seg_conv_bwd_kernel(const uint4* __restrict__ x, int C8, const float* __restrict__ w, const float* __restrict__ dlogits,
                    long long total, uint4* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                    int vox_per_block) {
  extern __shared__ float sm[];                 // [2][C] weights, then [2][C] dW partials + 2 db
  const int C = C8 * 8;
  float* sw = sm; float* sdw = sm + 2 * C; float* sdb = sdw + 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) { sw[i] = w[i]; sdw[i] = 0.f; }
  if (threadIdx.x < 2) sdb[threadIdx.x] = 0.f;
  __syncthreads();
  const int cc = threadIdx.x % C8, rows = blockDim.x / C8, rr = threadIdx.x / C8;
  float aw0[8], aw1[8], ab0 = 0.f, ab1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { aw0[j] = 0.f; aw1[j] = 0.f; }
  const long long v0 = (long long)blockIdx.x * vox_per_block;
  const long long v1 = min(v0 + (long long)vox_per_block, total);
  if (rr < rows)
    for (long long v = v0 + rr; v < v1; v += rows) {
      const float2 dl = reinterpret_cast<const float2*>(dlogits)[v];
      const float d0 = dl.x, d1 = dl.y;
      float f[8], o[8];
      unpack8(x[v * C8 + cc], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = d0 * sw[cc * 8 + j] + d1 * sw[C + cc * 8 + j];
        aw0[j] = fmaf(d0, f[j], aw0[j]);
        aw1[j] = fmaf(d1, f[j], aw1[j]);
      }
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
      dx[v * C8 + cc] = u;
      if (cc == 0) { ab0 += d0; ab1 += d1; }
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) { atomicAdd(&sdw[cc * 8 + j], aw0[j]); atomicAdd(&sdw[C + cc * 8 + j], aw1[j]); }
  if (cc == 0) { atomicAdd(&sdb[0], ab0); atomicAdd(&sdb[1], ab1); }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&dw[i], sdw[i]);
  if (threadIdx.x < 2) atomicAdd(&db[threadIdx.x], sdb[threadIdx.x]);
}

}  // namespace

extern "C" {

// x bf16 [total voxels, C]; w fp32 [2][C]; bias [2] -> logits fp32 [total, 2] (channels-last)
int nnd_seg_conv_fwd(const void* x, int C, const float* w, const float* bias, long long total, float* logits, cudaStream_t st) {
  if (C % 8 || total <= 0) return NND_ERR_ARG;
  seg_conv_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 2 * C * sizeof(float), st>>>((const uint4*)x, C / 8, w, bias, total, logits);
  NND_LAUNCH_CHECK("seg_conv_fwd_kernel");
  return NND_OK;
}

// target fp32 [total]; sums double[4] (kept for backward); losses_out[2] = {seg_ce, seg_dice}
int nnd_seg_loss_fwd(const float* logits, const float* target, long long total, float alpha, float smooth, double* sums,
                     float* losses_out, cudaStream_t st) {
  if (total <= 0) return NND_ERR_ARG;
  NND_CUDA_TRY(cudaMemsetAsync(sums, 0, 4 * sizeof(double), st));
  seg_loss_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(logits, target, total, sums);
  NND_LAUNCH_CHECK("seg_loss_fwd_kernel");
  seg_loss_kernel<<<1, 1, 0, st>>>(sums, (double)total, alpha, smooth, losses_out);
  NND_LAUNCH_CHECK("seg_loss_kernel");
  return NND_OK;
}

int nnd_seg_loss_bwd(const float* logits, const float* target, const double* sums, long long total, float alpha, float smooth,
                     const float* up_ce, const float* up_dice, float* dlogits, cudaStream_t st) {
  seg_loss_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(logits, target, sums, total, alpha, smooth, up_ce, up_dice, dlogits);
  NND_LAUNCH_CHECK("seg_loss_bwd_kernel");
  return NND_OK;
}

// dx bf16 [total, C]; dw [2][C], db [2] accumulated (caller zero-fills)
int nnd_seg_conv_bwd(const void* x, int C, const float* w, const float* dlogits, long long total, void* dx, float* dw, float* db,
                     cudaStream_t st) {
  if (C % 8 || C > 256) return NND_ERR_ARG;
  const int C8 = C / 8;
  const int threads = (256 / C8) * C8;
  const int vpb = 1024;
  const size_t smem = (size_t)(4 * C + 2) * sizeof(float);
  seg_conv_bwd_kernel<<<(unsigned)((total + vpb - 1) / vpb), threads, smem, st>>>((const uint4*)x, C8, w, dlogits, total, (uint4*)dx, dw, db, vpb);
  NND_LAUNCH_CHECK("seg_conv_bwd_kernel");
  return NND_OK;
}

}  // extern "C"
