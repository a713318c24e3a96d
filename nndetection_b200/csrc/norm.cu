// Instance / group normalisation (+ affine + ReLU) around the convolutions, NDHWC bf16, statistics in fp32.
//
// Reference: nn.InstanceNorm3d(eps 1e-5, affine) in encoder/decoder blocks (nndet/arch/conv.py:195,428), GroupNorm with
// 16 channels per group in the heads (nndet/arch/layers/norm.py:26-50, conv.py:271), nn.ReLU(inplace) (conv.py:445).
// InstanceNorm == GroupNorm with one channel per group, so one kernel family parameterised by `cpg` serves both.
//   forward : the producing conv's epilogue already accumulated per-(sample, channel) sum / sum-of-squares;
//             norm_finalize turns them into per-(sample, channel) scale/shift a, b; norm_apply streams y -> z once.
//   backward: one reduction pass (S1 = sum g, S2 = sum g * xhat with g = dz * [z > 0]) and one apply pass
//             dy = k1 * g + k2 * y + k3, plus dgamma / dbeta from S1, S2.
#include "common.cuh"

namespace {

// grid N, block C (<= 1024).  a = rstd * gamma, b = beta - mean * rstd * gamma; mean/rstd per (n, channel).
__global__ void norm_finalize_kernel(const float* __restrict__ ssum, const float* __restrict__ ssq,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, int C, int cpg,
                                     float count, float eps, float* __restrict__ a, float* __restrict__ b,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int n = blockIdx.x, c = threadIdx.x;
  if (c >= C) return;
  const int g0 = (c / cpg) * cpg;
  double s = 0.0, q = 0.0;
  for (int j = 0; j < cpg; ++j) { s += (double)ssum[n * C + g0 + j]; q += (double)ssq[n * C + g0 + j]; }
  const double m = (double)count * cpg;
  const double mean = s / m;
  double var = q / m - mean * mean;            // biased variance, as torch's instance/group norm
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  a[n * C + c] = rstd * ga;
  b[n * C + c] = be - (float)mean * rstd * ga;
  mean_out[n * C + c] = (float)mean;
  rstd_out[n * C + c] = rstd;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) { float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return u;
}

// z = relu?(a[n,c] * y + b[n,c]).  grid (chunks, N); a thread owns ONE 8-channel chunk (its 16 scale/shift values stay in
// registers) and walks the voxels of its block's range with 4 independent 16-byte loads in flight -- no per-element
// parameter loads, no 64-bit divisions (this pass is pure HBM streaming: 2 * C bytes read + written per voxel).
__global__ void __launch_bounds__(256)
norm_apply_kernel(const uint4* __restrict__ y, const float* __restrict__ a, const float* __restrict__ b,
                  int V, int C8, int rows, int vox_per_block, int relu, uint4* __restrict__ z) {
  const int n = blockIdx.y;
  const int cc = threadIdx.x % C8, rr = threadIdx.x / C8;
  if (rr >= rows) return;
  const int C = C8 * 8;
  float av[8], bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { av[j] = a[n * C + cc * 8 + j]; bv[j] = b[n * C + cc * 8 + j]; }
  const int v0 = blockIdx.x * vox_per_block, v1 = min(v0 + vox_per_block, V);
  const size_t base = (size_t)n * V * C8 + cc;
  for (int v = v0 + rr; v < v1; v += 4 * rows) {
    uint4 in[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (v + u * rows < v1) in[u] = y[base + (size_t)(v + u * rows) * C8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (v + u * rows >= v1) break;
      float f[8];
      unpack8(in[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(av[j], f[j], bv[j]);
        f[j] = relu ? fmaxf(t, 0.f) : t;
      }
      z[base + (size_t)(v + u * rows) * C8] = pack8(f);
    }
  }
}

// S1[n,c] += sum_v g, S2[n,c] += sum_v g * xhat.   grid (chunks, N), block = C8 * rows
__global__ void norm_bwd_reduce_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ y,
                                       const float* __restrict__ a, const float* __restrict__ b,
                                       const float* __restrict__ mean, const float* __restrict__ rstd, int V, int C8,
                                       int rows, int vox_per_block, int relu, float* __restrict__ S1, float* __restrict__ S2) {
  extern __shared__ float sh[];            // [rows][C8*8][2]
  const int n = blockIdx.y;
  const int cc = threadIdx.x % C8, rr = threadIdx.x / C8;
  const int C = C8 * 8, c = cc * 8;
  float av[8], bv[8], mv[8], rv[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    av[j] = a[n * C + c + j]; bv[j] = b[n * C + c + j]; mv[j] = mean[n * C + c + j]; rv[j] = rstd[n * C + c + j];
    s1[j] = 0.f; s2[j] = 0.f;
  }
  const int v0 = blockIdx.x * vox_per_block, v1 = min(v0 + vox_per_block, V);
  const size_t base = (size_t)n * V * C8 + cc;
  for (int v = v0 + rr; v < v1; v += 4 * rows) {
    uint4 iy[4], id[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u * rows < v1) { iy[u] = y[base + (size_t)(v + u * rows) * C8]; id[u] = dz[base + (size_t)(v + u * rows) * C8]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (v + u * rows >= v1) break;
      float fy[8], fd[8];
      unpack8(iy[u], fy); unpack8(id[u], fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pre = fmaf(av[j], fy[j], bv[j]);
        const float gq = (!relu || pre > 0.f) ? fd[j] : 0.f;
        s1[j] += gq;
        s2[j] += gq * (fy[j] - mv[j]) * rv[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sh[(rr * C + c + j) * 2] = s1[j]; sh[(rr * C + c + j) * 2 + 1] = s2[j]; }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float t1 = 0.f, t2 = 0.f;
    for (int r = 0; r < rows; ++r) { t1 += sh[(r * C + ch) * 2]; t2 += sh[(r * C + ch) * 2 + 1]; }
    atomicAdd(&S1[n * C + ch], t1);
    atomicAdd(&S2[n * C + ch], t2);
  }
}

// grid N, block C: coefficients of dy = k1 * g + k2 * y + k3, and dgamma / dbeta accumulation
__global__ void norm_bwd_finalize_kernel(const float* __restrict__ S1, const float* __restrict__ S2,
                                         const float* __restrict__ gamma, const float* __restrict__ mean,
                                         const float* __restrict__ rstd, int C, int cpg, float count,
                                         float* __restrict__ k1, float* __restrict__ k2, float* __restrict__ k3,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int n = blockIdx.x, c = threadIdx.x;
  if (c >= C) return;
  const int g0 = (c / cpg) * cpg;
  float ds = 0.f, db = 0.f;
  for (int j = 0; j < cpg; ++j) {
    const float ga = gamma ? gamma[g0 + j] : 1.f;
    ds += ga * S2[n * C + g0 + j];
    db += ga * S1[n * C + g0 + j];
  }
  const float m = count * cpg;
  const float r = rstd[n * C + c], mu = mean[n * C + c], ga = gamma ? gamma[c] : 1.f;
  // dy = r*ga*g - r*(xhat*ds + db)/m, xhat = (y - mu) * r
  const float c2 = r * ds / m;
  k1[n * C + c] = r * ga;
  k2[n * C + c] = -c2 * r;
  k3[n * C + c] = c2 * r * mu - r * db / m;
  if (dgamma) atomicAdd(&dgamma[c], S2[n * C + c]);
  if (dbeta) atomicAdd(&dbeta[c], S1[n * C + c]);
}

// dy = k1 * g + k2 * y + k3 with g = dz * [pre-activation > 0]; same thread mapping as norm_apply_kernel (40 per-channel
// coefficients in registers, 2 x 4 independent 16-byte loads in flight)
__global__ void __launch_bounds__(256)
norm_bwd_apply_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ y, const float* __restrict__ a,
                      const float* __restrict__ b, const float* __restrict__ k1, const float* __restrict__ k2,
                      const float* __restrict__ k3, int V, int C8, int rows, int vox_per_block, int relu,
                      uint4* __restrict__ dy) {
  const int n = blockIdx.y;
  const int cc = threadIdx.x % C8, rr = threadIdx.x / C8;
  if (rr >= rows) return;
  const int C = C8 * 8;
  float av[8], bv[8], c1[8], c2[8], c3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int o = n * C + cc * 8 + j;
    av[j] = a[o]; bv[j] = b[o]; c1[j] = k1[o]; c2[j] = k2[o]; c3[j] = k3[o];
  }
  const int v0 = blockIdx.x * vox_per_block, v1 = min(v0 + vox_per_block, V);
  const size_t base = (size_t)n * V * C8 + cc;
  for (int v = v0 + rr; v < v1; v += 4 * rows) {
    uint4 iy[4], id[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u * rows < v1) { iy[u] = y[base + (size_t)(v + u * rows) * C8]; id[u] = dz[base + (size_t)(v + u * rows) * C8]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (v + u * rows >= v1) break;
      float fy[8], fd[8];
      unpack8(iy[u], fy); unpack8(id[u], fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pre = fmaf(av[j], fy[j], bv[j]);
        const float gq = (!relu || pre > 0.f) ? fd[j] : 0.f;
        fd[j] = fmaf(c1[j], gq, fmaf(c2[j], fy[j], c3[j]));
      }
      dy[base + (size_t)(v + u * rows) * C8] = pack8(fd);
    }
  }
}

// ---- narrow variants of the two backward passes (opt-in, nnd_norm_set_bwd_narrow): a thread owns FOUR channels (8-byte loads).
// The 8-channel kernels above need 127 / 112 registers (coefficients of 8 channels + 8 x 16-byte loads in flight) = two 256-thread
// CTAs per SM, and were measured at 38 % / 47 % of the HBM peak where norm_apply (62 registers) reaches 88 %; halving the per-thread
// state doubles the resident warps.  Same per-element arithmetic (dy is bit-identical for identical coefficients); S1 / S2 are summed
// in a different order.
__device__ __forceinline__ void unpack4(const uint2& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 2; ++j) { float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint2 pack4(const float* f) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 2; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return u;
}

// grid (chunks, N), block = C4 * rows
__global__ void __launch_bounds__(256)
norm_bwd_reduce4_kernel(const uint2* __restrict__ dz, const uint2* __restrict__ y, const float* __restrict__ a,
                        const float* __restrict__ b, const float* __restrict__ mean, const float* __restrict__ rstd, int V, int C4,
                        int rows, int vox_per_block, int relu, float* __restrict__ S1, float* __restrict__ S2) {
  extern __shared__ float sh[];            // [rows][C][2]
  const int n = blockIdx.y;
  const int cc = threadIdx.x % C4, rr = threadIdx.x / C4;
  const int C = C4 * 4, c = cc * 4;
  float av[4], bv[4], mv[4], rv[4], s1[4], s2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    av[j] = a[n * C + c + j]; bv[j] = b[n * C + c + j]; mv[j] = mean[n * C + c + j]; rv[j] = rstd[n * C + c + j];
    s1[j] = 0.f; s2[j] = 0.f;
  }
  const int v0 = blockIdx.x * vox_per_block, v1 = min(v0 + vox_per_block, V);
  const size_t base = (size_t)n * V * C4 + cc;
  if (rr < rows) {
    for (int v = v0 + rr; v < v1; v += 4 * rows) {
      uint2 iy[4], id[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v + u * rows < v1) { iy[u] = y[base + (size_t)(v + u * rows) * C4]; id[u] = dz[base + (size_t)(v + u * rows) * C4]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (v + u * rows >= v1) break;
        float fy[4], fd[4];
        unpack4(iy[u], fy); unpack4(id[u], fd);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float pre = fmaf(av[j], fy[j], bv[j]);
          const float gq = (!relu || pre > 0.f) ? fd[j] : 0.f;
          s1[j] += gq;
          s2[j] += gq * (fy[j] - mv[j]) * rv[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { sh[(rr * C + c + j) * 2] = s1[j]; sh[(rr * C + c + j) * 2 + 1] = s2[j]; }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float t1 = 0.f, t2 = 0.f;
    for (int r = 0; r < rows; ++r) { t1 += sh[(r * C + ch) * 2]; t2 += sh[(r * C + ch) * 2 + 1]; }
    atomicAdd(&S1[n * C + ch], t1);
    atomicAdd(&S2[n * C + ch], t2);
  }
}

__global__ void __launch_bounds__(256)
norm_bwd_apply4_kernel(const uint2* __restrict__ dz, const uint2* __restrict__ y, const float* __restrict__ a,
                       const float* __restrict__ b, const float* __restrict__ k1, const float* __restrict__ k2,
                       const float* __restrict__ k3, int V, int C4, int rows, int vox_per_block, int relu, uint2* __restrict__ dy) {
  const int n = blockIdx.y;
  const int cc = threadIdx.x % C4, rr = threadIdx.x / C4;
  if (rr >= rows) return;
  const int C = C4 * 4;
  float av[4], bv[4], c1[4], c2[4], c3[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = n * C + cc * 4 + j;
    av[j] = a[o]; bv[j] = b[o]; c1[j] = k1[o]; c2[j] = k2[o]; c3[j] = k3[o];
  }
  const int v0 = blockIdx.x * vox_per_block, v1 = min(v0 + vox_per_block, V);
  const size_t base = (size_t)n * V * C4 + cc;
  for (int v = v0 + rr; v < v1; v += 4 * rows) {
    uint2 iy[4], id[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u * rows < v1) { iy[u] = y[base + (size_t)(v + u * rows) * C4]; id[u] = dz[base + (size_t)(v + u * rows) * C4]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (v + u * rows >= v1) break;
      float fy[4], fd[4];
      unpack4(iy[u], fy); unpack4(id[u], fd);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pre = fmaf(av[j], fy[j], bv[j]);
        const float gq = (!relu || pre > 0.f) ? fd[j] : 0.f;
        fd[j] = fmaf(c1[j], gq, fmaf(c2[j], fy[j], c3[j]));
      }
      dy[base + (size_t)(v + u * rows) * C4] = pack4(fd);
    }
  }
}

int g_norm_bwd_narrow = 0;

// launch shape shared by the streaming passes: grid (voxel chunks, N), block = rows x C8 threads (<= 256)
struct StreamShape { int rows, threads, vpb; dim3 grid; };
static StreamShape stream_shape(int N, long long V, int C8) {
  StreamShape s;
  s.rows = 256 / C8; if (s.rows < 1) s.rows = 1;
  s.threads = s.rows * C8;
  s.vpb = 2048;
  while (s.vpb > 64 && (V + s.vpb - 1) / s.vpb * N < NND_NUM_SMS * 8) s.vpb >>= 1;
  s.grid = dim3((unsigned)((V + s.vpb - 1) / s.vpb), (unsigned)N);
  return s;
}

}  // namespace

extern "C" {

// 1: four-channels-per-thread backward passes (see above); default 0 until measured on a B200 (written without one)
void nnd_norm_set_bwd_narrow(int enable) { g_norm_bwd_narrow = enable; }

// stats [N,C] from the conv epilogue -> a, b, mean, rstd [N,C].  count = voxels per sample.
int nnd_norm_finalize(const float* ssum, const float* ssq, const float* gamma, const float* beta, int N, int C, int cpg,
                      long long count, float eps, float* a, float* b, float* mean, float* rstd, cudaStream_t st) {
  if (C > 1024 || C % cpg) return NND_ERR_ARG;
  norm_finalize_kernel<<<N, C, 0, st>>>(ssum, ssq, gamma, beta, C, cpg, (float)count, eps, a, b, mean, rstd);
  NND_LAUNCH_CHECK("norm_finalize_kernel");
  return NND_OK;
}

// y, z: bf16 [N, V, C] (C % 8 == 0)
int nnd_norm_apply(const void* y, const float* a, const float* b, int N, long long V, int C, int relu, void* z,
                   cudaStream_t st) {
  if (C % 8) return NND_ERR_ARG;
  if ((long long)N * V == 0) return NND_OK;
  if (C > 2048 || V > 0x7fffffffll) return NND_ERR_ARG;
  const StreamShape sh = stream_shape(N, V, C / 8);
  norm_apply_kernel<<<sh.grid, sh.threads, 0, st>>>((const uint4*)y, a, b, (int)V, C / 8, sh.rows, sh.vpb, relu, (uint4*)z);
  NND_LAUNCH_CHECK("norm_apply_kernel");
  return NND_OK;
}

// dz, y bf16 [N,V,C]; a,b,mean,rstd [N,C] from the forward; gamma [C].  dy bf16 [N,V,C]; dgamma/dbeta [C] accumulated.
// ws: 5 * N * C floats.
int nnd_norm_backward(const void* dz, const void* y, const float* a, const float* b, const float* mean,
                      const float* rstd, const float* gamma, int N, long long V, int C, int cpg, int relu, void* dy,
                      float* dgamma, float* dbeta, float* ws, cudaStream_t st) {
  if (C % 8 || C > 1024 || C % cpg) return NND_ERR_ARG;
  float* S1 = ws; float* S2 = ws + (size_t)N * C; float* k1 = S2 + (size_t)N * C; float* k2 = k1 + (size_t)N * C;
  float* k3 = k2 + (size_t)N * C;
  NND_CUDA_TRY(cudaMemsetAsync(S1, 0, sizeof(float) * 2 * N * C, st));
  if (g_norm_bwd_narrow && C <= 1024) {
    const int C4 = C / 4;
    int rows = 256 / C4; if (rows < 1) rows = 1;
    const int threads = rows * C4 > 256 ? C4 : rows * C4;              // C4 <= 256 here (C <= 1024)
    int vpb = 2048;
    while (vpb > 64 && (V + vpb - 1) / vpb * N < NND_NUM_SMS * 8) vpb >>= 1;
    dim3 grid((unsigned)((V + vpb - 1) / vpb), N);
    const size_t smem = (size_t)rows * C * 2 * sizeof(float);
    norm_bwd_reduce4_kernel<<<grid, threads, smem, st>>>((const uint2*)dz, (const uint2*)y, a, b, mean, rstd, (int)V, C4, rows, vpb,
                                                         relu, S1, S2);
    NND_LAUNCH_CHECK("norm_bwd_reduce4_kernel");
    norm_bwd_finalize_kernel<<<N, C, 0, st>>>(S1, S2, gamma, mean, rstd, C, cpg, (float)V, k1, k2, k3, dgamma, dbeta);
    NND_LAUNCH_CHECK("norm_bwd_finalize_kernel");
    norm_bwd_apply4_kernel<<<grid, threads, 0, st>>>((const uint2*)dz, (const uint2*)y, a, b, k1, k2, k3, (int)V, C4, rows, vpb, relu,
                                                     (uint2*)dy);
    NND_LAUNCH_CHECK("norm_bwd_apply4_kernel");
    return NND_OK;
  }
  const int C8 = C / 8;
  int rows = 256 / C8; if (rows < 1) rows = 1;
  const int threads = rows * C8;
  int vpb = 2048;
  while (vpb > 64 && (V + vpb - 1) / vpb * N < NND_NUM_SMS * 4) vpb >>= 1;
  dim3 grid((unsigned)((V + vpb - 1) / vpb), N);
  const size_t smem = (size_t)rows * C * 2 * sizeof(float);
  norm_bwd_reduce_kernel<<<grid, threads, smem, st>>>((const uint4*)dz, (const uint4*)y, a, b, mean, rstd, (int)V, C8, rows,
                                                      vpb, relu, S1, S2);
  NND_LAUNCH_CHECK("norm_bwd_reduce_kernel");
  norm_bwd_finalize_kernel<<<N, C, 0, st>>>(S1, S2, gamma, mean, rstd, C, cpg, (float)V, k1, k2, k3, dgamma, dbeta);
  NND_LAUNCH_CHECK("norm_bwd_finalize_kernel");
  const StreamShape sh = stream_shape(N, V, C8);
  norm_bwd_apply_kernel<<<sh.grid, sh.threads, 0, st>>>((const uint4*)dz, (const uint4*)y, a, b, k1, k2, k3, (int)V, C8, sh.rows,
                                                        sh.vpb, relu, (uint4*)dy);
  NND_LAUNCH_CHECK("norm_bwd_apply_kernel");
  return NND_OK;
}

}  // extern "C"
