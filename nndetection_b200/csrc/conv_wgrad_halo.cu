// Weight gradient of stride-1 convolutions with tap offsets in [-1, 1] (3x3x3, 1x3x3, ...): "halo" formulation.
//
//   dW[co][ci][tap] = sum_voxels dy[v][co] * x[v + off_tap][ci]
//
// The per-tap kernel of conv_wgrad.cu re-reads dy and x once per tap (27x): at the 32-channel full-resolution layers
// that is 29 GB of L2 traffic per layer and ~50 TFLOP/s.  Here a CTA stages ONE voxel block (1 x 8 x 16 voxels) of dy
// and the matching input block WITH its halo in shared memory and runs every tap of its tap group against them:
// ldmatrix takes per-lane row addresses, so a tap shift is just a different set of row pointers into the halo tile
// (XOR-swizzled rows keep any 8 consecutive voxels conflict-free).  Tap group = all taps (32x32 tiles: 108
// accumulator registers) or the <= 9 taps of one depth offset (wider tiles).  Split-K over voxel blocks, fp32 atomics.
#include "conv_common.cuh"

namespace {

constexpr int VB_H = 8, VB_W = 16;                 // voxel block 1 x 8 x 16 (K = 128 voxels = 8 k16 steps)
// halo extents depend on the stride: rows (VB_H - 1) * sh + 3, columns (VB_W - 1) * sw + 3  (10 x 18 at stride 1, 17 x 33 at 2)

template <int ROWB>
__device__ __forceinline__ int swzh(int row, int chunk) {
  return ROWB == 128 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3));
}
__device__ __forceinline__ void ldsm_x2_trans(unsigned addr, unsigned& r0, unsigned& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}

struct WhArgs {
  const __nv_bfloat16* dy; int Cdy;
  const __nv_bfloat16* x; int Cx;
  float* dw; long long s_co, s_ci, s_tap;
  int Cout, Cin;
  int N, D, H, W;             // OUTPUT (dy) grid
  int Di, Hi, Wi;             // input (x) grid
  int sd, sh, sw;             // strides: x position = v * s + off
  int hbh, hbw;               // halo rows / columns per depth slice
  int hb, wb;                 // blocks along h, w
  long long total_blocks;     // N * D * hb * wb
  long long blocks_per_split;
  int ci_tiles;
  int n_groups;               // tap groups (grid.y)
  int grp_z[3];               // depth offset of each group (ALLZ: unused)
  int grp_cnt[3];             // taps in each group
  signed char tdz[27], tdy[27], tdx[27];   // taps ordered group by group
  unsigned char tw[27];
  int grp_start[3];
};

// BM x BN output tile, WM x WN warps, MAXT accumulated taps per CTA, ZS = halo depth slices held (1 or 3)
template <int BM, int BN, int WM, int WN, int MAXT, int ZS>
__global__ void __launch_bounds__(WM * WN * 32)
conv_wgrad_halo_kernel(const WhArgs a) {
  constexpr int NI = BN / (8 * WN);
  static_assert(BM == 16 * WM && (NI == 1 || NI == 2), "tile/warp mismatch");
  constexpr int THREADS = WM * WN * 32;
  constexpr int A_ROWB = BM * 2, B_ROWB = BN * 2;
  constexpr int A_ROWS = VB_H * VB_W;
  constexpr int A_BYTES = A_ROWS * A_ROWB;
  const int HB_H = a.hbh, HB_W = a.hbw;
  const int B_ROWS = ZS * HB_H * HB_W, B_BYTES = B_ROWS * B_ROWB;
  constexpr int A_CH = BM / 8, B_CH = BN / 8;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                       // 2 stages
  unsigned char* sB = smem + 2 * A_BYTES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp % WM, warp_n = warp / WM;
  const int grp = blockIdx.y;
  const int co0 = (blockIdx.z / a.ci_tiles) * BM, ci0 = (blockIdx.z % a.ci_tiles) * BN;
  const long long b0 = (long long)blockIdx.x * a.blocks_per_split;
  const long long b1 = min(b0 + a.blocks_per_split, a.total_blocks);
  if (b0 >= b1) return;
  const int ntaps = a.grp_cnt[grp], tap0 = a.grp_start[grp];
  const int zbase = (ZS == 3) ? -1 : a.grp_z[grp];       // halo slice 0 holds depth d + zbase

  auto load_block = [&](int stage, long long blk) {
    int r = (int)(blk % a.wb); long long q = blk / a.wb;
    const int wblk = r; r = (int)(q % a.hb); q /= a.hb;
    const int hblk = r; const int d = (int)(q % a.D); const int n = (int)(q / a.D);
    const int h0 = hblk * VB_H, w0 = wblk * VB_W;
    const unsigned a_base = smem_u32(sA + stage * A_BYTES), b_base = smem_u32(sB + stage * B_BYTES);
    for (int i = tid; i < A_ROWS * A_CH; i += THREADS) {
      const int row = i / A_CH, ch = i % A_CH;
      const int h = h0 + row / VB_W, w = w0 + row % VB_W;
      const bool ok = h < a.H && w < a.W;
      const __nv_bfloat16* src = ok ? a.dy + ((((long long)n * a.D + d) * a.H + h) * a.W + w) * a.Cdy + co0 + ch * 8 : a.dy;
      cp_async16(a_base + row * A_ROWB + swzh<A_ROWB>(row, ch) * 16, src, ok);
    }
    for (int i = tid; i < B_ROWS * B_CH; i += THREADS) {
      const int row = i / B_CH, ch = i % B_CH;
      const int zz = row / (HB_H * HB_W); const int r2 = row % (HB_H * HB_W);
      const int dd = d * a.sd + zbase + zz, h = h0 * a.sh - 1 + r2 / HB_W, w = w0 * a.sw - 1 + r2 % HB_W;
      const bool ok = (unsigned)dd < (unsigned)a.Di && (unsigned)h < (unsigned)a.Hi && (unsigned)w < (unsigned)a.Wi;
      const __nv_bfloat16* src = ok ? a.x + ((((long long)n * a.Di + dd) * a.Hi + h) * a.Wi + w) * a.Cx + ci0 + ch * 8 : a.x;
      cp_async16(b_base + row * B_ROWB + swzh<B_ROWB>(row, ch) * 16, src, ok);
    }
  };

  float acc[MAXT][NI][4];
#pragma unroll
  for (int t = 0; t < MAXT; ++t)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[t][j][k] = 0.f;

  // row offset of tap t inside the halo tile, relative to (block row 0, k 0); taps beyond ntaps alias tap 0 (their
  // products are computed and dropped) so that the unrolled tap loop carries no branches and the compiler can
  // interleave the ldmatrix / mma chains of different taps.
  int toff[MAXT];
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    const int tt = tap0 + (t < ntaps ? t : 0);
    const int tz = (ZS == 3) ? a.tdz[tt] + 1 : 0;
    toff[t] = (tz * HB_H + 1 + a.tdy[tt]) * HB_W + 1 + a.tdx[tt];
  }

  load_block(0, b0);
  cp_async_commit();
  int stage = 0;
  for (long long blk = b0; blk < b1; ++blk) {
    if (blk + 1 < b1) load_block(stage ^ 1, blk + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const unsigned a_st = smem_u32(sA + stage * A_BYTES), b_st = smem_u32(sB + stage * B_BYTES);
#pragma unroll 1
    for (int hr = 0; hr < VB_H; ++hr) {
      // A fragment: dy^T, k = the 16 voxels of block row hr, m = this warp's 16 output channels
      unsigned af[4];
      {
        const int krow = hr * VB_W + (lane & 7) + ((lane >> 4) << 3);
        const int mcol = warp_m * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(a_st + krow * A_ROWB + swzh<A_ROWB>(krow, mcol >> 3) * 16, af[0], af[1], af[2], af[3]);
      }
      const int kk = ((lane & 7) + ((lane >> 3) & 1) * 8) * a.sw;   // halo column of this lane's k index (w inside the block row)
      const int ncol = warp_n * (BN / WN) + (NI == 2 ? (lane >> 4) * 8 : 0);
      const int row0 = hr * a.sh * HB_W + kk;
#pragma unroll
      for (int t = 0; t < MAXT; ++t) {
        const int hv = toff[t] + row0;
        unsigned bf[NI][2];
        if (NI == 2) ldmatrix_x4_trans(b_st + hv * B_ROWB + swzh<B_ROWB>(hv, ncol >> 3) * 16, bf[0][0], bf[0][1], bf[NI - 1][0], bf[NI - 1][1]);
        else ldsm_x2_trans(b_st + hv * B_ROWB + swzh<B_ROWB>(hv, ncol >> 3) * 16, bf[0][0], bf[0][1]);
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) mma_bf16_16816(acc[t][ni], af[0], af[1], af[2], af[3], bf[ni][0], bf[ni][1]);
      }
    }
    __syncthreads();
    stage ^= 1;
  }
  cp_async_wait<0>();

  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    if (t < ntaps) {
      float* dwt = a.dw + (long long)a.tw[tap0 + t] * a.s_tap;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = co0 + warp_m * 16 + gq + (e >> 1) * 8;
          const int ci = ci0 + warp_n * (BN / WN) + ni * 8 + tq * 2 + (e & 1);
          if (co < a.Cout && ci < a.Cin) atomicAdd(dwt + co * a.s_co + ci * a.s_ci, acc[t][ni][e]);
        }
    }
  }
}

template <int BM, int BN, int WM, int WN, int MAXT, int ZS>
int launch_halo(WhArgs a, int co_pad, int ci_pad, cudaStream_t st) {
  constexpr int THREADS = WM * WN * 32;
  const size_t SMEM = 2 * ((size_t)VB_H * VB_W * BM * 2 + (size_t)ZS * a.hbh * a.hbw * BN * 2);
  if (SMEM > 227 * 1024) return NND_ERR_ARG;
  const int co_tiles = co_pad / BM;
  a.ci_tiles = ci_pad / BN;
  const long long tiles = (long long)a.n_groups * co_tiles * a.ci_tiles;
  long long splits = (NND_NUM_SMS * 2 + tiles - 1) / tiles;
  if (splits > a.total_blocks) splits = a.total_blocks;
  if (splits < 1) splits = 1;
  a.blocks_per_split = (a.total_blocks + splits - 1) / splits;
  splits = (a.total_blocks + a.blocks_per_split - 1) / a.blocks_per_split;
  static size_t attr_smem = 0;
  if (SMEM > attr_smem) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_halo_kernel<BM, BN, WM, WN, MAXT, ZS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_smem = SMEM;
  }
  dim3 grid((unsigned)splits, (unsigned)a.n_groups, (unsigned)(co_tiles * a.ci_tiles));
  conv_wgrad_halo_kernel<BM, BN, WM, WN, MAXT, ZS><<<grid, THREADS, SMEM, st>>>(a);
  NND_LAUNCH_CHECK("conv_wgrad_halo_kernel");
  return NND_OK;
}

}  // namespace

int nnd_conv_wgrad_halo_supported(const ConvGeom& g, int Cdy, int Cx) {
  if (g.sd < 1 || g.sd > 2 || g.sh < 1 || g.sh > 2 || g.sw < 1 || g.sw > 2) return 0;
  if (g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return 0;
  if (g.Ld != g.Do || g.Lh != g.Ho || g.Lw != g.Wo) return 0;
  if (g.T < 9 || Cdy % 32 || Cx % 32) return 0;
  for (int t = 0; t < g.T; ++t)
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
  return 1;
}

int nnd_conv_wgrad_halo(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                        long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st) {
  WhArgs a;
  a.dy = dy; a.Cdy = Cdy; a.x = x; a.Cx = Cx; a.dw = dw; a.s_co = s_co; a.s_ci = s_ci; a.s_tap = s_tap;
  a.Cout = Cout; a.Cin = Cin; a.N = g.N; a.D = g.Do; a.H = g.Ho; a.W = g.Wo;
  a.Di = g.Di; a.Hi = g.Hi; a.Wi = g.Wi; a.sd = g.sd; a.sh = g.sh; a.sw = g.sw;
  a.hbh = (VB_H - 1) * g.sh + 3; a.hbw = (VB_W - 1) * g.sw + 3;
  a.hb = (g.Ho + VB_H - 1) / VB_H; a.wb = (g.Wo + VB_W - 1) / VB_W;
  a.total_blocks = (long long)g.N * g.Do * a.hb * a.wb;
  if (a.total_blocks <= 0) return NND_OK;
  const bool strided = g.sd != 1 || g.sh != 1 || g.sw != 1;
  const bool small = (Cdy % 64 != 0) && (Cx % 64 != 0) && !strided;       // 32 x 32 tiles, stride 1: all taps in one CTA
  // order taps group by group (group = depth offset), or one group with everything
  int cnt = 0;
  a.n_groups = 0;
  for (int z = -1; z <= 1; ++z) {
    int c0 = cnt;
    for (int t = 0; t < g.T; ++t)
      if (g.off_d[t] == z) { a.tdz[cnt] = g.off_d[t]; a.tdy[cnt] = g.off_h[t]; a.tdx[cnt] = g.off_w[t]; a.tw[cnt] = g.tap_w[t]; ++cnt; }
    if (cnt > c0 && !small) { a.grp_z[a.n_groups] = z; a.grp_start[a.n_groups] = c0; a.grp_cnt[a.n_groups] = cnt - c0; ++a.n_groups; }
  }
  if (small) { a.n_groups = 1; a.grp_z[0] = 0; a.grp_start[0] = 0; a.grp_cnt[0] = cnt; }
  for (int i = 0; i < a.n_groups; ++i) if (a.grp_cnt[i] > (small ? 27 : 9)) return NND_ERR_ARG;
  const bool m64 = Cdy % 64 == 0, n64 = Cx % 64 == 0;
  if (m64 && n64) return launch_halo<64, 64, 4, 4, 9, 1>(a, Cdy, Cx, st);
  if (m64) return launch_halo<64, 32, 4, 2, 9, 1>(a, Cdy, Cx, st);
  if (n64) return launch_halo<32, 64, 2, 4, 9, 1>(a, Cdy, Cx, st);
  if (small) return launch_halo<32, 32, 2, 4, 27, 3>(a, Cdy, Cx, st);
  return launch_halo<32, 32, 2, 4, 9, 1>(a, Cdy, Cx, st);
}
