// Small streaming kernels around the network: weight packing, SGD step, gradient padding/casts, bias gradients.
#include "common.cuh"

namespace {

// w fp32 (PyTorch conv layout [Cout][Cin][T], or transposed-conv layout [Cin][Cout][T] when `transposed`)
//  -> fwd  bf16 [T][CoutPad][CinPadF]  (rows = output channel, K-major)   : fprop operand
//  -> bwd  bf16 [T][CinPadB][CoutPadK] (rows = input channel)             : dgrad operand
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int T, int transposed,
                                    __nv_bfloat16* __restrict__ fwd, int CoutPadF, int CinPadF,
                                    __nv_bfloat16* __restrict__ bwd, int CinPadB, int CoutPadB) {
  const long long nf = (long long)T * CoutPadF * CinPadF, nb = (long long)T * CinPadB * CoutPadB;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (fwd && i < nf) {
    const int ci = (int)(i % CinPadF); long long r = i / CinPadF; const int co = (int)(r % CoutPadF); const int t = (int)(r / CoutPadF);
    float v = 0.f;
    if (co < Cout && ci < Cin) v = transposed ? w[((size_t)ci * Cout + co) * T + t] : w[((size_t)co * Cin + ci) * T + t];
    fwd[i] = __float2bfloat16(v);
  }
  if (bwd && i < nb) {
    const int co = (int)(i % CoutPadB); long long r = i / CoutPadB; const int ci = (int)(r % CinPadB); const int t = (int)(r / CinPadB);
    float v = 0.f;
    if (co < Cout && ci < Cin) v = transposed ? w[((size_t)ci * Cout + co) * T + t] : w[((size_t)co * Cin + ci) * T + t];
    bwd[i] = __float2bfloat16(v);
  }
}

// K-major pack [T][rows_pad][K] (bf16) -> pipeline-item order [row tile][k chunk][T][k group][n][8]: one tap slice of one
// (8 * kg)-channel chunk and one n_tile-row tile becomes n_tile * kg * 16 contiguous bytes -- the unit conv_tc.cu's BULK variant
// fetches with one cp.async.bulk.  One thread moves one 16-byte group.
__global__ void repack_items_kernel(const uint4* __restrict__ src, int T, int rows_pad, int K, int n_tile, int kg, uint4* __restrict__ dst) {
  const int KC = K / (8 * kg), NT = rows_pad / n_tile;
  const long long total = (long long)T * rows_pad * (K / 8);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;            // destination index in 16-byte groups
  if (i >= total) return;
  const int n = (int)(i % n_tile); long long r = i / n_tile;
  const int g = (int)(r % kg); r /= kg;
  const int t = (int)(r % T); r /= T;
  const int kc = (int)(r % KC); const int nt = (int)(r / KC);
  if (nt >= NT) return;
  dst[i] = src[((long long)t * rows_pad + (nt * n_tile + n)) * (K / 8) + kc * kg + g];
}

// torch.optim.SGD(momentum, nesterov, weight_decay) on a flat fp32 buffer; elements >= n_decay get no weight decay
// (norm parameters, nndet/training/optimizer/utils.py).  first_step: momentum buffer initialised with the gradient.
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, long long n,
                           long long n_decay, float lr, float momentum, float wd, int nesterov, int first_step,
                           float grad_scale, const long long* __restrict__ skip, int n_skip) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // parameters that never receive a gradient (torch.optim.SGD skips `p.grad is None`: no decay, no momentum): [lo, hi) element ranges
  for (int k = 0; k < n_skip; ++k)
    if (i >= skip[2 * k] && i < skip[2 * k + 1]) return;
  float gr = g[i] * grad_scale;
  const float pv = p[i];
  if (i < n_decay) gr = fmaf(wd, pv, gr);
  float b = first_step ? gr : fmaf(momentum, mom[i], gr);
  mom[i] = b;
  const float step = nesterov ? fmaf(momentum, b, gr) : b;
  p[i] = pv - lr * step;
}

// dst bf16 [N][rows][Cpad] <- src fp32 (sample stride src_n_stride, row stride C) * (*mul or 1), zero padded
__global__ void pad_cast_kernel(const float* __restrict__ src, int N, long long rows, int C, long long src_n_stride,
                                const float* __restrict__ mul, __nv_bfloat16* __restrict__ dst, int Cpad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * rows * Cpad) return;
  const int c = (int)(i % Cpad); long long r = i / Cpad; const long long n = r / rows; r = r % rows;
  const float m = mul ? *mul : 1.f;
  dst[i] = __float2bfloat16(c < C ? src[n * src_n_stride + r * C + c] * m : 0.f);
}

// out[c] += sum_rows src[r][c]   (bias gradient).  grid (chunks), block 256; src bf16 or fp32
template <typename T>
__global__ void channel_sum_kernel(const T* __restrict__ src, long long rows, int C, long long stride, int rows_per_block,
                                   float scale, float* __restrict__ out) {
  extern __shared__ float sh[];          // [C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) sh[c] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(r0 + (long long)rows_per_block, rows);
  const int tpr = min(C, (int)blockDim.x);          // threads per row
  const int rpb = blockDim.x / tpr;
  const int c0 = threadIdx.x % tpr, rr = threadIdx.x / tpr;
  if (rr < rpb)
    for (int c = c0; c < C; c += tpr) {
      float acc = 0.f;
      for (long long r = r0 + rr; r < r1; r += rpb) acc += (float)src[r * stride + c];
      atomicAdd(&sh[c], acc);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&out[c], sh[c] * scale);
}

// bf16 fast path (C % 8 == 0, stride % 8 == 0): a thread owns one 8-channel chunk, 16-byte loads, 4 rows in flight
__global__ void __launch_bounds__(256)
channel_sum_bf16x8_kernel(const uint4* __restrict__ src, long long rows, int C8, long long stride8, int rows_per_block, float scale,
                          float* __restrict__ out) {
  extern __shared__ float sh[];          // [C8 * 8]
  for (int c = threadIdx.x; c < C8 * 8; c += blockDim.x) sh[c] = 0.f;
  __syncthreads();
  const int rpb = blockDim.x / C8;
  const int cc = threadIdx.x % C8, rr = threadIdx.x / C8;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(r0 + (long long)rows_per_block, rows);
  if (rr < rpb) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long long r = r0 + rr; r < r1; r += 4 * rpb) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) if (r + u * rpb < r1) v[u] = src[(r + u * rpb) * stride8 + cc];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u * rpb >= r1) break;
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 t = __bfloat1622float2(h[k]); acc[2 * k] += t.x; acc[2 * k + 1] += t.y; }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sh[cc * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C8 * 8; c += blockDim.x) atomicAdd(&out[c], sh[c] * scale);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = __float2bfloat16(s[i]);
}

// dscale += sum g * out / scale over [N][len] fp32 ranges at sample stride n_stride (regressor Scale backward)
__global__ void scale_grad_kernel(const float* __restrict__ g, const float* __restrict__ out, int N, long long len,
                                  long long n_stride, const float* __restrict__ scale, float* __restrict__ dscale) {
  __shared__ float sh[32];
  float acc = 0.f;
  const long long total = (long long)N * len;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long o = (i / len) * n_stride + (i % len);
    const float gv = g[o];
    if (gv != 0.f) acc += gv * out[o];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0 && r != 0.f) atomicAdd(dscale, r / *scale);
  }
}

// same reduction, 16-byte loads, one grid row per sample (no per-element division): the dense gradient it scans is zero except for
// the <= max_pos sampled anchors, so the launch is a pure streaming read of g (85 MB at the 32^3 level of the LUNA plan)
__global__ void __launch_bounds__(256)
scale_grad_vec_kernel(const float4* __restrict__ g, const float4* __restrict__ out, long long len4, long long n_stride4,
                      const float* __restrict__ scale, float* __restrict__ dscale) {
  __shared__ float sh[8];
  const float4* gn = g + (long long)blockIdx.y * n_stride4;
  const float4* on = out + (long long)blockIdx.y * n_stride4;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len4; i += (long long)gridDim.x * blockDim.x) {
    const float4 gv = gn[i];
    if (gv.x != 0.f || gv.y != 0.f || gv.z != 0.f || gv.w != 0.f) {
      const float4 ov = on[i];
      acc += gv.x * ov.x + gv.y * ov.y + gv.z * ov.z + gv.w * ov.w;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0 && r != 0.f) atomicAdd(dscale, r / *scale);
  }
}

}  // namespace

extern "C" {

int nnd_pack_weights(const float* w, int Cout, int Cin, int T, int transposed, void* fwd, int CoutPadF, int CinPadF,
                     void* bwd, int CinPadB, int CoutPadB, cudaStream_t st) {
  const long long nf = fwd ? (long long)T * CoutPadF * CinPadF : 0, nb = bwd ? (long long)T * CinPadB * CoutPadB : 0;
  const long long n = nf > nb ? nf : nb;
  if (n == 0) return NND_OK;
  pack_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, Cout, Cin, T, transposed, (__nv_bfloat16*)fwd, CoutPadF,
                                                                   CinPadF, (__nv_bfloat16*)bwd, CinPadB, CoutPadB);
  NND_LAUNCH_CHECK("pack_weights_kernel");
  return NND_OK;
}

int nnd_repack_items_bf16(const void* src, int T, int rows_pad, int K, int n_tile, int kg, void* dst, cudaStream_t st) {
  if (!src || !dst || T <= 0 || n_tile <= 0 || kg <= 0 || rows_pad % n_tile || K % (8 * kg)) return NND_ERR_ARG;
  const long long total = (long long)T * rows_pad * (K / 8);
  repack_items_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint4*)src, T, rows_pad, K, n_tile, kg, (uint4*)dst);
  NND_LAUNCH_CHECK("repack_items_kernel");
  return NND_OK;
}

// skip: device array of n_skip [lo, hi) element ranges left untouched (parameters without a gradient), or NULL / 0
int nnd_sgd_step_skip(float* p, const float* g, float* mom, long long n, long long n_decay, float lr, float momentum, float wd,
                      int nesterov, int first_step, float grad_scale, const long long* skip, int n_skip, cudaStream_t st) {
  if (n <= 0) return NND_OK;
  if (n_skip < 0 || (n_skip > 0 && !skip)) return NND_ERR_ARG;
  sgd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, mom, n, n_decay, lr, momentum, wd, nesterov, first_step, grad_scale, skip, n_skip);
  NND_LAUNCH_CHECK("sgd_kernel");
  return NND_OK;
}

int nnd_sgd_step(float* p, const float* g, float* mom, long long n, long long n_decay, float lr, float momentum, float wd,
                 int nesterov, int first_step, float grad_scale, cudaStream_t st) {
  return nnd_sgd_step_skip(p, g, mom, n, n_decay, lr, momentum, wd, nesterov, first_step, grad_scale, nullptr, 0, st);
}

int nnd_pad_cast_f32_bf16(const float* src, int N, long long rows, int C, long long src_n_stride, const float* mul, void* dst,
                          int Cpad, cudaStream_t st) {
  const long long n = (long long)N * rows * Cpad;
  if (n <= 0) return NND_OK;
  pad_cast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, N, rows, C, src_n_stride, mul, (__nv_bfloat16*)dst, Cpad);
  NND_LAUNCH_CHECK("pad_cast_kernel");
  return NND_OK;
}

// is_bf16: src element type; out[C] accumulated
int nnd_channel_sum(const void* src, int is_bf16, long long rows, int C, long long stride, float scale, float* out, cudaStream_t st) {
  if (rows <= 0) return NND_OK;
  int rpb = 4096;
  while (rpb > 64 && (rows + rpb - 1) / rpb < NND_NUM_SMS * 2) rpb >>= 1;
  const unsigned blocks = (unsigned)((rows + rpb - 1) / rpb);
  if (is_bf16 && C % 8 == 0 && stride % 8 == 0 && C / 8 <= 256 && ((size_t)src & 15) == 0) {
    const int C8 = C / 8, threads = (256 / C8) * C8;
    channel_sum_bf16x8_kernel<<<blocks, threads, C * sizeof(float), st>>>((const uint4*)src, rows, C8, stride / 8, rpb, scale, out);
  } else if (is_bf16) channel_sum_kernel<__nv_bfloat16><<<blocks, 256, C * sizeof(float), st>>>((const __nv_bfloat16*)src, rows, C, stride, rpb, scale, out);
  else channel_sum_kernel<float><<<blocks, 256, C * sizeof(float), st>>>((const float*)src, rows, C, stride, rpb, scale, out);
  NND_LAUNCH_CHECK("channel_sum_kernel");
  return NND_OK;
}

int nnd_cast_f32_bf16(const float* s, void* d, long long n, cudaStream_t st) {
  if (n <= 0) return NND_OK;
  cast_f32_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, (__nv_bfloat16*)d, n);
  NND_LAUNCH_CHECK("cast_f32_bf16_kernel");
  return NND_OK;
}

int nnd_scale_grad(const float* g, const float* out, int N, long long len, long long n_stride, const float* scale, float* dscale,
                   cudaStream_t st) {
  if ((long long)N * len <= 0) return NND_OK;
  if (len % 4 == 0 && n_stride % 4 == 0 && (((size_t)g | (size_t)out) & 15) == 0) {
    const long long len4 = len / 4;
    long long bx = (len4 + 256 * 8 - 1) / (256 * 8);                 // ~8 float4 per thread
    if (bx > 4 * NND_NUM_SMS) bx = 4 * NND_NUM_SMS;
    if (bx < 1) bx = 1;
    scale_grad_vec_kernel<<<dim3((unsigned)bx, (unsigned)N), 256, 0, st>>>((const float4*)g, (const float4*)out, len4, n_stride / 4, scale, dscale);
  } else {
    scale_grad_kernel<<<NND_NUM_SMS, 256, 0, st>>>(g, out, N, len, n_stride, scale, dscale);
  }
  NND_LAUNCH_CHECK("scale_grad_kernel");
  return NND_OK;
}

}  // extern "C"
