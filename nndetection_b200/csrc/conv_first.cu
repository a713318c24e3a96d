// First encoder convolution: Cin in {1..4} image channels -> Cout (32) features, stride 1.
// K = taps * Cin is 27..108: far too thin for tensor-core tiles and only 0.5 % of the network's FLOPs, so this is
// a CUDA-core direct convolution, bandwidth-bound on the bf16 output write.
// Input is the fp32 NCDHW image batch exactly as the data loader hands it over; weights are the fp32 master copy
// in PyTorch layout [Cout][Cin][taps].  Output NDHWC bf16 + per-(sample, channel) statistics for the InstanceNorm.
#include "conv_common.cuh"

namespace {

constexpr int FIRST_MAX_CIN = 4;

template <int COUT>
__global__ void __launch_bounds__(128)
conv_first_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w, const ConvGeom g,
                        __nv_bfloat16* __restrict__ out, float* __restrict__ stat_sum, float* __restrict__ stat_sq) {
  extern __shared__ float sw[];                         // [tap][ci][COUT]
  __shared__ float s_red[2][COUT];
  const int Cin = g.Cin, T = g.T;
  for (int i = threadIdx.x; i < T * Cin * COUT; i += blockDim.x) {
    const int co = i % COUT, r = i / COUT, ci = r % Cin, t = r / Cin;
    sw[i] = w[((size_t)co * Cin + ci) * T + g.tap_w[t]];
  }
  if (threadIdx.x < COUT) { s_red[0][threadIdx.x] = 0.f; s_red[1][threadIdx.x] = 0.f; }
  __syncthreads();
  const int V = g.Ld * g.Lh * g.Lw;
  const int n = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = v < V;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
  if (ok) {
    const int lw = v % g.Lw; const int r = v / g.Lw; const int lh = r % g.Lh; const int ld = r / g.Lh;
    const size_t plane = (size_t)g.Di * g.Hi * g.Wi;
    const float* xn = x + (size_t)n * Cin * plane;
    for (int t = 0; t < T; ++t) {
      const int id = ld + g.off_d[t], ih = lh + g.off_h[t], iw = lw + g.off_w[t];
      if ((unsigned)id >= (unsigned)g.Di || (unsigned)ih >= (unsigned)g.Hi || (unsigned)iw >= (unsigned)g.Wi) continue;
      const size_t pos = ((size_t)id * g.Hi + ih) * g.Wi + iw;
      for (int ci = 0; ci < Cin; ++ci) {
        const float xv = __ldg(xn + ci * plane + pos);
        const float* wr = sw + (t * Cin + ci) * COUT;
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = fmaf(xv, wr[c], acc[c]);
      }
    }
    __nv_bfloat16* o = out + ((size_t)n * V + v) * COUT;
#pragma unroll
    for (int c = 0; c < COUT; c += 8) {
      __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pk[j] = __floats2bfloat162_rn(acc[c + 2 * j], acc[c + 2 * j + 1]);
        acc[c + 2 * j] = __bfloat162float(pk[j].x);
        acc[c + 2 * j + 1] = __bfloat162float(pk[j].y);
      }
      *reinterpret_cast<uint4*>(o + c) = *reinterpret_cast<const uint4*>(pk);
    }
  }
  if (stat_sum) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      float s = ok ? acc[c] : 0.f, q = s * s;
      s = warp_sum(s); q = warp_sum(q);
      if ((threadIdx.x & 31) == 0) { atomicAdd(&s_red[0][c], s); atomicAdd(&s_red[1][c], q); }
    }
    __syncthreads();
    if (threadIdx.x < COUT) {
      atomicAdd(&stat_sum[(size_t)n * COUT + threadIdx.x], s_red[0][threadIdx.x]);
      atomicAdd(&stat_sq[(size_t)n * COUT + threadIdx.x], s_red[1][threadIdx.x]);
    }
  }
}

// dW[co][ci][tap] += sum_v dy[v][co] * x[v + off_tap][ci].  Warps split the taps, lanes are output channels.
__global__ void __launch_bounds__(256)
conv_first_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const ConvGeom g, int Cout,
                        int chunk, float* __restrict__ dw) {
  constexpr int MAXT = 4;                       // taps per warp (8 warps * 4 >= 27)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int V = g.Ld * g.Lh * g.Lw;
  const int n = blockIdx.y;
  const int v0 = blockIdx.x * chunk, v1 = min(v0 + chunk, V);
  const size_t plane = (size_t)g.Di * g.Hi * g.Wi;
  const float* xn = x + (size_t)n * g.Cin * plane;
  for (int cb = 0; cb < Cout; cb += 32) {
    const int co = cb + lane;
    float acc[MAXT][FIRST_MAX_CIN];
#pragma unroll
    for (int i = 0; i < MAXT; ++i)
#pragma unroll
      for (int j = 0; j < FIRST_MAX_CIN; ++j) acc[i][j] = 0.f;
    for (int v = v0; v < v1; ++v) {
      const float d = co < Cout ? __bfloat162float(dy[((size_t)n * V + v) * Cout + co]) : 0.f;
      const int lw = v % g.Lw; const int r = v / g.Lw; const int lh = r % g.Lh; const int ld = r / g.Lh;
#pragma unroll
      for (int i = 0; i < MAXT; ++i) {
        const int t = warp + i * 8;
        if (t >= g.T) break;
        const int id = ld + g.off_d[t], ih = lh + g.off_h[t], iw = lw + g.off_w[t];
        if ((unsigned)id >= (unsigned)g.Di || (unsigned)ih >= (unsigned)g.Hi || (unsigned)iw >= (unsigned)g.Wi) continue;
        const size_t pos = ((size_t)id * g.Hi + ih) * g.Wi + iw;
#pragma unroll
        for (int ci = 0; ci < FIRST_MAX_CIN; ++ci)
          if (ci < g.Cin) acc[i][ci] = fmaf(d, __ldg(xn + ci * plane + pos), acc[i][ci]);
      }
    }
    if (co < Cout) {
#pragma unroll
      for (int i = 0; i < MAXT; ++i) {
        const int t = warp + i * 8;
        if (t >= g.T) break;
#pragma unroll
        for (int ci = 0; ci < FIRST_MAX_CIN; ++ci)
          if (ci < g.Cin) atomicAdd(&dw[((size_t)co * g.Cin + ci) * g.T + g.tap_w[t]], acc[i][ci]);
      }
    }
  }
}

}  // namespace

// x fp32 [N, Cin, D, H, W] (NCDHW), w fp32 [Cout][Cin][T]; geometry: stride 1, Ld/Lh/Lw == Di/Hi/Wi.
int nnd_conv_first_fprop(const float* x, const float* w, const ConvGeom& g, int Cout, __nv_bfloat16* out,
                         float* stat_sum, float* stat_sq, cudaStream_t st) {
  if (!x || !w || !out || g.Cin < 1 || g.Cin > FIRST_MAX_CIN) return NND_ERR_ARG;
  const int V = g.Ld * g.Lh * g.Lw;
  dim3 grid((V + 127) / 128, g.N);
  const size_t smem = (size_t)g.T * g.Cin * Cout * sizeof(float);
  if (Cout == 32) conv_first_fprop_kernel<32><<<grid, 128, smem, st>>>(x, w, g, out, stat_sum, stat_sq);
  else if (Cout == 16) conv_first_fprop_kernel<16><<<grid, 128, smem, st>>>(x, w, g, out, stat_sum, stat_sq);
  else if (Cout == 48) conv_first_fprop_kernel<48><<<grid, 128, smem, st>>>(x, w, g, out, stat_sum, stat_sq);
  else if (Cout == 64) conv_first_fprop_kernel<64><<<grid, 128, smem, st>>>(x, w, g, out, stat_sum, stat_sq);
  else return NND_ERR_ARG;
  NND_LAUNCH_CHECK("conv_first_fprop_kernel");
  return NND_OK;
}

int nnd_conv_first_wgrad(const float* x, const __nv_bfloat16* dy, const ConvGeom& g, int Cout, float* dw, cudaStream_t st) {
  if (!x || !dy || !dw || g.Cin < 1 || g.Cin > FIRST_MAX_CIN || g.T > 32) return NND_ERR_ARG;
  const int V = g.Ld * g.Lh * g.Lw;
  int chunk = 2048;
  while (chunk > 128 && (long long)((V + chunk - 1) / chunk) * g.N < NND_NUM_SMS * 4) chunk >>= 1;
  dim3 grid((V + chunk - 1) / chunk, g.N);
  conv_first_wgrad_kernel<<<grid, 256, 0, st>>>(x, dy, g, Cout, chunk, dw);
  NND_LAUNCH_CHECK("conv_first_wgrad_kernel");
  return NND_OK;
}
