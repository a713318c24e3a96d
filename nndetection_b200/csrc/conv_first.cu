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

// dW[co][ci][tap] += sum_v dy[v][co] * x[v + off_tap][ci].
// One CTA = ROWS consecutive output rows (fixed n, d; h0..h0+ROWS-1).  Warp = one (dz, dy) pair of the filter, lane =
// output channel; the three dx taps slide along w in registers, so a voxel costs 2 shared loads + 3 FMAs per thread.
// dy row tile (bf16) and the 9 input rows it needs (fp32, w-halo included) are staged in shared memory per row.
constexpr int FW_ROWS = 8;
__global__ void __launch_bounds__(288)
conv_first_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const ConvGeom g, int Cout,
                        float* __restrict__ dw) {
  extern __shared__ unsigned char fsm[];
  const int W = g.Lw, H = g.Lh, D = g.Ld;
  float* sx = reinterpret_cast<float*>(fsm);                           // [9][W + 2]
  __nv_bfloat16* sdy = reinterpret_cast<__nv_bfloat16*>(sx + 9 * (W + 2));   // [W][32]
  __shared__ int s_tap[27];                                            // weight tap index of (dz,dy,dx) or -1
  if (threadIdx.x < 27) s_tap[threadIdx.x] = -1;
  __syncthreads();
  if (threadIdx.x < g.T) {
    const int t = threadIdx.x;
    s_tap[((g.off_d[t] + 1) * 3 + (g.off_h[t] + 1)) * 3 + (g.off_w[t] + 1)] = g.tap_w[t];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;         // warp 0..8 = (dz+1)*3 + (dy+1)
  const int dz = warp / 3 - 1, dyo = warp % 3 - 1;
  const int t0 = s_tap[warp * 3 + 0], t1 = s_tap[warp * 3 + 1], t2 = s_tap[warp * 3 + 2];
  const int rows_per_d = (H + FW_ROWS - 1) / FW_ROWS;
  const int n = blockIdx.x / (D * rows_per_d);
  const int rem = blockIdx.x % (D * rows_per_d);
  const int d = rem / rows_per_d, h0 = (rem % rows_per_d) * FW_ROWS;
  const size_t plane = (size_t)D * H * W;
  for (int cb = 0; cb < Cout; cb += 32) {
    for (int ci = 0; ci < g.Cin; ++ci) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      const float* xp = x + ((size_t)n * g.Cin + ci) * plane;
      for (int hr = 0; hr < FW_ROWS && h0 + hr < H; ++hr) {
        const int h = h0 + hr;
        __syncthreads();
        for (int i = threadIdx.x; i < 9 * (W + 2); i += blockDim.x) {
          const int r = i / (W + 2), wx = i % (W + 2) - 1;
          const int zz = d + r / 3 - 1, yy = h + r % 3 - 1;
          float v = 0.f;
          if ((unsigned)zz < (unsigned)D && (unsigned)yy < (unsigned)H && (unsigned)wx < (unsigned)W)
            v = __ldg(xp + ((size_t)zz * H + yy) * W + wx);
          sx[i] = v;
        }
        const __nv_bfloat16* dyr = dy + (((size_t)n * D + d) * H + h) * W * Cout + cb;
        for (int i = threadIdx.x; i < W * 32; i += blockDim.x) {
          const int wv = i >> 5, c = i & 31;
          sdy[i] = (cb + c < Cout) ? dyr[(size_t)wv * Cout + c] : __float2bfloat16(0.f);
        }
        __syncthreads();
        const float* xr = sx + warp * (W + 2);       // row (dz, dy), index w + 1 + dx
        float xm = xr[0], xc = xr[1];
        for (int wv = 0; wv < W; ++wv) {
          const float xn = xr[wv + 2];
          const float dv = __bfloat162float(sdy[wv * 32 + lane]);
          a0 = fmaf(dv, xm, a0); a1 = fmaf(dv, xc, a1); a2 = fmaf(dv, xn, a2);
          xm = xc; xc = xn;
        }
      }
      const int co = cb + lane;
      if (co < Cout) {
        float* base = dw + ((size_t)co * g.Cin + ci) * g.T;
        if (t0 >= 0) atomicAdd(base + t0, a0);
        if (t1 >= 0) atomicAdd(base + t1, a1);
        if (t2 >= 0) atomicAdd(base + t2, a2);
      }
    }
  }
  (void)dz; (void)dyo;
}

}  // namespace

int nnd_conv_first_mma_supported(const ConvGeom& g, int Cout);
int nnd_conv_first_fprop_mma(const float* x, const float* w, const ConvGeom& g, __nv_bfloat16* out, float* stat_sum,
                             float* stat_sq, cudaStream_t st);
int nnd_conv_first_wgrad_mma(const float* x, const __nv_bfloat16* dy, const ConvGeom& g, float* dw, cudaStream_t st);
static int g_first_mma = 1;
extern "C" void nnd_conv_set_first_layer_mma(int enable) { g_first_mma = enable; }   // A/B: 0 = scalar CUDA-core kernels

// x fp32 [N, Cin, D, H, W] (NCDHW), w fp32 [Cout][Cin][T]; geometry: stride 1, Ld/Lh/Lw == Di/Hi/Wi.
int nnd_conv_first_fprop(const float* x, const float* w, const ConvGeom& g, int Cout, __nv_bfloat16* out,
                         float* stat_sum, float* stat_sq, cudaStream_t st) {
  if (!x || !w || !out || g.Cin < 1 || g.Cin > FIRST_MAX_CIN) return NND_ERR_ARG;
  if (g_first_mma && nnd_conv_first_mma_supported(g, Cout)) return nnd_conv_first_fprop_mma(x, w, g, out, stat_sum, stat_sq, st);
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
  if (!x || !dy || !dw || g.Cin < 1 || g.Cin > FIRST_MAX_CIN || g.T > 27) return NND_ERR_ARG;
  for (int t = 0; t < g.T; ++t)
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return NND_ERR_ARG;
  if (g_first_mma && nnd_conv_first_mma_supported(g, Cout)) return nnd_conv_first_wgrad_mma(x, dy, g, dw, st);
  const int rows_per_d = (g.Lh + FW_ROWS - 1) / FW_ROWS;
  const unsigned blocks = (unsigned)(g.N * g.Ld * rows_per_d);
  const size_t smem = (size_t)9 * (g.Lw + 2) * sizeof(float) + (size_t)g.Lw * 32 * sizeof(__nv_bfloat16);
  if (smem > 200 * 1024) return NND_ERR_ARG;
  if (smem > 48 * 1024) NND_CUDA_TRY(cudaFuncSetAttribute(conv_first_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_first_wgrad_kernel<<<blocks, 288, smem, st>>>(x, dy, g, Cout, dw);
  NND_LAUNCH_CHECK("conv_first_wgrad_kernel");
  return NND_OK;
}
