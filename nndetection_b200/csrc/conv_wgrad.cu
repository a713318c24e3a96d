// Weight gradients of the gather convolution (conv_common.cuh), mma.sync bf16 / fp32 accumulate, split-K.
//
//   dW[co][ci][tap] = sum_{n, lo}  dy[n, lo*om + oo][co] * x[n, lo*s + off_tap][ci]
//
// i.e. a GEMM with M = Cout, N = Cin (per tap) and K = all output voxels; both operands are read "K-row-major"
// ([voxel][channel], exactly how the activations live in HBM) and transposed for free by ldmatrix.trans.
// Grid = (K splits, taps, co-tiles * ci-tiles); every CTA adds its partial tile to fp32 dW with atomics.
// Replaces cuDNN wgrad as invoked by autograd for torch.nn.Conv3d / ConvTranspose3d (nndet/arch/conv.py:344-348).
#include "conv_common.cuh"

namespace {

constexpr int BKV = 32, WSTAGES = 4, WTHREADS = 256;

template <int ROWB>
__device__ __forceinline__ int swz(int row, int chunk) {
  return ROWB == 128 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3));
}

__device__ __forceinline__ void ldmatrix_x2_trans(unsigned addr, unsigned& r0, unsigned& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}

struct WgradArgs {
  const __nv_bfloat16* dy; int Cdy;     // dy [N, Do, Ho, Wo, Cdy]
  const __nv_bfloat16* x;  int Cx;      // x  [N, Di, Hi, Wi, Cx]
  float* dw; long long s_co, s_ci, s_tap;
  int Cout, Cin;                        // real channel counts (store mask)
  long long total;                      // N * Ld*Lh*Lw
  long long chunk;                      // voxels per K split (multiple of 32)
  int ci_tiles;
};

template <int BMC, int BNC>
__global__ void __launch_bounds__(WTHREADS)
conv_wgrad_kernel(const ConvGeom g, const WgradArgs a) {
  constexpr int MI = BMC / 32, NI = BNC / 32;             // per-warp m16 / n8 tiles (2 x 4 warps)
  constexpr int A_ROWB = BMC * 2, B_ROWB = BNC * 2;
  constexpr int A_BYTES = BKV * A_ROWB, B_BYTES = BKV * B_ROWB;
  constexpr int A_CH = BMC / 8, B_CH = BNC / 8;           // 16-byte chunks per row
  __shared__ __align__(128) unsigned char sA[WSTAGES * A_BYTES];
  __shared__ __align__(128) unsigned char sB[WSTAGES * B_BYTES];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp & 1, warp_n = warp >> 1;
  const int tap = blockIdx.y;
  const int co0 = (blockIdx.z / a.ci_tiles) * BMC, ci0 = (blockIdx.z % a.ci_tiles) * BNC;
  const long long kb = (long long)blockIdx.x * a.chunk;
  const long long ke = min(kb + a.chunk, a.total);
  if (kb >= ke) return;
  const int iters = (int)((ke - kb + BKV - 1) / BKV);
  const int Lvox = g.Ld * g.Lh * g.Lw;
  const int od = g.off_d[tap], oh = g.off_h[tap], ow = g.off_w[tap];

  auto load_stage = [&](int stage, int it) {
    const long long v0 = kb + (long long)it * BKV;
    const unsigned a_base = smem_u32(sA + stage * A_BYTES), b_base = smem_u32(sB + stage * B_BYTES);
    for (int i = tid; i < BKV * (A_CH + B_CH); i += WTHREADS) {
      const bool isA = i < BKV * A_CH;
      const int j = isA ? i : i - BKV * A_CH;
      const int row = isA ? j / A_CH : j / B_CH;
      const int ch = isA ? j % A_CH : j % B_CH;
      const long long m = v0 + row;
      bool ok = m < ke;
      const __nv_bfloat16* src = a.dy;
      if (ok) {
        const int n = (int)(m / Lvox);
        const int lo = (int)(m % Lvox);
        const int lw = lo % g.Lw; const int r = lo / g.Lw; const int lh = r % g.Lh; const int ld = r / g.Lh;
        if (isA) {
          const long long pv = ((long long)(ld * g.omd + g.ood) * g.Ho + (lh * g.omh + g.ooh)) * g.Wo + (lw * g.omw + g.oow);
          src = a.dy + ((long long)n * g.Do * g.Ho * g.Wo + pv) * a.Cdy + co0 + ch * 8;
        } else {
          const int id = ld * g.sd + od, ih = lh * g.sh + oh, iw = lw * g.sw + ow;
          ok = (unsigned)id < (unsigned)g.Di && (unsigned)ih < (unsigned)g.Hi && (unsigned)iw < (unsigned)g.Wi;
          if (ok) src = a.x + (((long long)(n * g.Di + id) * g.Hi + ih) * g.Wi + iw) * a.Cx + ci0 + ch * 8;
        }
      }
      const unsigned dst = isA ? a_base + row * A_ROWB + swz<A_ROWB>(row, ch) * 16
                               : b_base + row * B_ROWB + swz<B_ROWB>(row, ch) * 16;
      cp_async16(dst, src, ok);
    }
  };

  float acc[MI][NI][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
  for (int s = 0; s < WSTAGES - 1; ++s) {
    if (s < iters) load_stage(s, s);
    cp_async_commit();
  }
  for (int it = 0; it < iters; ++it) {
    cp_async_wait<WSTAGES - 2>();
    __syncthreads();
    if (it + WSTAGES - 1 < iters) load_stage((it + WSTAGES - 1) % WSTAGES, it + WSTAGES - 1);
    cp_async_commit();
    const int stage = it % WSTAGES;
    const unsigned a_st = smem_u32(sA + stage * A_BYTES), b_st = smem_u32(sB + stage * B_BYTES);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      unsigned af[MI][4];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        const int krow = kk * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int mcol = warp_m * (BMC / 2) + mi * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(a_st + krow * A_ROWB + swz<A_ROWB>(krow, mcol >> 3) * 16, af[mi][0], af[mi][1], af[mi][2], af[mi][3]);
      }
      unsigned bf[NI][2];
      if (NI == 2) {
        const int krow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ncol = warp_n * (BNC / 4) + (lane >> 4) * 8;
        ldmatrix_x4_trans(b_st + krow * B_ROWB + swz<B_ROWB>(krow, ncol >> 3) * 16, bf[0][0], bf[0][1], bf[NI - 1][0], bf[NI - 1][1]);
      } else {
        const int krow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ncol = warp_n * (BNC / 4);
        ldmatrix_x2_trans(b_st + krow * B_ROWB + swz<B_ROWB>(krow, ncol >> 3) * 16, bf[0][0], bf[0][1]);
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
          mma_bf16_16816(acc[mi][ni], af[mi][0], af[mi][1], af[mi][2], af[mi][3], bf[ni][0], bf[ni][1]);
    }
  }
  cp_async_wait<0>();

  const int gq = lane >> 2, tq = lane & 3;
  float* dwt = a.dw + (long long)g.tap_w[tap] * a.s_tap;
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < NI; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int co = co0 + warp_m * (BMC / 2) + mi * 16 + gq + (e >> 1) * 8;
        const int ci = ci0 + warp_n * (BNC / 4) + ni * 8 + tq * 2 + (e & 1);
        if (co < a.Cout && ci < a.Cin) atomicAdd(dwt + co * a.s_co + ci * a.s_ci, acc[mi][ni][e]);
      }
}

template <int BMC, int BNC>
int launch_wgrad(const ConvGeom& g, WgradArgs a, int co_pad, int ci_pad, cudaStream_t st) {
  const int co_tiles = co_pad / BMC;
  a.ci_tiles = ci_pad / BNC;
  const long long tiles = (long long)g.T * co_tiles * a.ci_tiles;
  long long splits = (NND_NUM_SMS * 8 + tiles - 1) / tiles;
  const long long max_splits = (a.total + 1023) / 1024;          // at least 32 K-iterations per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long chunk = (a.total + splits - 1) / splits;
  chunk = (chunk + BKV - 1) / BKV * BKV;
  splits = (a.total + chunk - 1) / chunk;
  a.chunk = chunk;
  dim3 grid((unsigned)splits, (unsigned)g.T, (unsigned)(co_tiles * a.ci_tiles));
  conv_wgrad_kernel<BMC, BNC><<<grid, WTHREADS, 0, st>>>(g, a);
  NND_LAUNCH_CHECK("conv_wgrad_kernel");
  return NND_OK;
}

}  // namespace

// dy [N,Do,Ho,Wo,Cdy] bf16 (Cdy % 32 == 0, channels >= Cout are padding), x [N,Di,Hi,Wi,Cx] bf16 (Cx % 32 == 0).
// dw fp32, accumulated (caller zero-fills): dw[co*s_co + ci*s_ci + tap_w*s_tap].
int nnd_conv_wgrad(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                   long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st) {
  if (!dy || !x || !dw || Cdy % 32 || Cx % 32) return NND_ERR_ARG;
  WgradArgs a;
  a.dy = dy; a.Cdy = Cdy; a.x = x; a.Cx = Cx; a.dw = dw; a.s_co = s_co; a.s_ci = s_ci; a.s_tap = s_tap;
  a.Cout = Cout; a.Cin = Cin;
  a.total = (long long)g.N * g.Ld * g.Lh * g.Lw;
  if (a.total <= 0) return NND_OK;
  const bool m64 = Cdy % 64 == 0, n64 = Cx % 64 == 0;
  if (m64 && n64) return launch_wgrad<64, 64>(g, a, Cdy, Cx, st);
  if (m64) return launch_wgrad<64, 32>(g, a, Cdy, Cx, st);
  if (n64) return launch_wgrad<32, 64>(g, a, Cdy, Cx, st);
  return launch_wgrad<32, 32>(g, a, Cdy, Cx, st);
}
