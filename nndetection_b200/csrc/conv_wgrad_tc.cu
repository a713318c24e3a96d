// Weight gradient of 3x3x3 / 1x3x3 stride-1 convolutions on the 5th-generation tensor cores.
//
//   dW[tap][co][ci] = sum_voxels dy[v][co] * x[v + off_tap][ci]          (nndet/arch/conv.py:344-348 via autograd)
//
// As a UMMA: D[M = co][N = ci] += A[M][K] * B[N][K] with K = 16 consecutive voxels of one w-row.  Both operands live in
// HBM as [voxel][channel] (NDHWC), i.e. they are "MN-major": staged in shared memory as [channel group of 8][row][voxel]
// [8 ch] one voxel is again a 16-byte core-matrix row (canonical no-swizzle MN-major layout: 8 k-rows x 16 B, LBO =
// 128 B between the two k-halves, SBO = channel-group pitch), and the three dx-taps of a filter row are the SAME x row
// read at start offsets 0 / 16 / 32 bytes -- no im2col, no per-tap reload.
// One CTA = one (dz, dy) filter row (3 taps -> 3 accumulators of N_T columns in TMEM), one 128-channel co tile, one
// N_T-channel ci tile, one contiguous range of voxel rows (split-K).  4 producer warps (cp.async, zero fill = padding)
// -> 4-stage ring -> 3 issuer warps (one per dx tap, independent accumulators) -> 4 epilogue warps (tcgen05.ld ->
// fp32 atomicAdd into dW, PyTorch weight layout).
// (Round-2 A/B: a producer mapping with consecutive lanes on consecutive channel groups of one voxel + a 16-byte padded group pitch
// ran 2-4x SLOWER on the B200 -- 128->128 @32^3: 2.91 vs 1.35 ms -- and was reverted; the mapping below already covers 128 contiguous
// bytes per row with lanes 0, 4, 8 ... of each instruction.)
//
// SW = 2 (opt-in until validated on the device, nnd_conv_set_wgrad_strided_tc): the same kernel for stride-2 convolutions,
//   dW[tap][co][ci] = sum_o dy[o][co] * x[o * s + off_tap][ci].
// dy rows are dense as before; an x row is loaded DE-INTERLEAVED: the 33 input voxels 2*w0 - 1 ... 2*w0 + 31 behind 16 outputs go
// to two planes, "odd" (w = 2p - 1, p = 0..16, slots 0..16) and "even" (w = 2p, p = 0..15, slots 17..32), so that the operand
// of tap dx is again 16 CONSECUTIVE 16-byte rows: dx = -1 -> odd plane from slot 0, dx = 0 -> even plane (slot 17), dx = +1 ->
// odd plane from slot 1.  d / h strides only change which input row is fetched.
#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int RW = 16;                 // voxels per row segment (K of one MMA)
constexpr int ROWS = 4;                // row segments per pipeline stage
constexpr int STAGES = 4;
constexpr int M_T = 128;               // co tile (UMMA M); channels beyond Cdy are zero rows
constexpr int WG_THREADS = (4 + 3 + 4) * 32;
constexpr int NPROD = 128;


struct WtArgs {
  const __nv_bfloat16* dy; int Cdy;
  const __nv_bfloat16* x; int Cx;
  float* dw; long long s_co, s_ci, s_tap;
  int Cout, Cin;
  int N, D, H, W, wsegs;            // dy grid and row segments per h-row (ceil(W / 16))
  int Di, Hi, Wi, sd, sh;           // SW = 2 only: x grid and the d / h strides (w stride = SW)
  long long total_rows;             // N * D * H * wsegs
  long long rows_per_split;
  int n_groups;                     // filter rows (dz, dy)
  signed char gdz[9], gdy[9];
  unsigned char gtw[9][3];          // weight tap index of dx = -1, 0, +1 (255 = tap absent)
  int ci_tiles;
};

template <int N_T, int SW>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc_kernel(const WtArgs a) {
  constexpr int XW = SW == 1 ? RW + 2 : 2 * RW + 1;      // x row slots: 18 (halo row) or 17 + 16 (odd / even plane)
  constexpr int A_GROUPS = M_T / 8, B_GROUPS = N_T / 8;
  constexpr int A_GPITCH = ROWS * RW * 16;               // bytes between co groups inside a stage
  constexpr int B_GPITCH = ROWS * XW * 16;               // bytes between ci groups
  constexpr int A_BYTES = A_GROUPS * A_GPITCH, B_BYTES = B_GROUPS * B_GPITCH;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 3 * N_T > 256 ? 512 : (3 * N_T > 128 ? 256 : 128);
  // kind::f16, D fp32, A/B bf16, both MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
  constexpr unsigned IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(N_T >> 3) << 17) | ((128u >> 4) << 24);

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + STAGES * STAGE_BYTES);
  __shared__ unsigned s_tmem_base;
  const unsigned bar0 = smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (STAGES + i); };
  const unsigned DONE = bar0 + 8u * (2 * STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = blockIdx.y;
  const int co0 = (blockIdx.z / a.ci_tiles) * M_T, ci0 = (blockIdx.z % a.ci_tiles) * N_T;
  const long long r0 = (long long)blockIdx.x * a.rows_per_split;
  const long long r1 = min(r0 + a.rows_per_split, a.total_rows);
  const int n_stages = (int)((r1 - r0 + ROWS - 1) / ROWS);
  const int dz = a.gdz[grp], dyo = a.gdy[grp];
  const int co_groups = min(A_GROUPS, (a.Cdy - co0) / 8);           // real channel groups of this co tile

  // zero the A region once: channel groups beyond Cdy stay zero for the whole kernel (zero MMA rows)
  for (int i = tid; i < STAGES * STAGE_BYTES / 16; i += WG_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(FULL(i), NPROD / 32); mbar_init(EMPTY(i), 3); }
    mbar_init(DONE, 3);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = s_tmem_base;
  if (n_stages > 0) {
  if (warp < 4) {
    // ================================================================ producers
    unsigned stage = 0, phase = 0;
    int pending = 0; unsigned done_stage = 0;
    constexpr int LAGW = 2;
    for (int s = 0; s < n_stages; ++s) {
      if (lane == 0) mbar_wait(EMPTY(stage), phase ^ 1);
      __syncwarp();
      const unsigned sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
      // A: dy rows  [co group][row][16 voxels]   B: x rows [ci group][row][18 voxels]
      const int a_items = co_groups * ROWS, b_items = B_GROUPS * ROWS;
      for (int it = tid; it < a_items + b_items; it += NPROD) {
        const bool isA = it < a_items;
        const int j = isA ? it : it - a_items;
        const int gidx = j / ROWS, rr = j % ROWS;
        const long long row = r0 + (long long)s * ROWS + rr;
        bool row_ok = row < r1;
        int ws = 0, h = 0, d = 0, n = 0;
        if (row_ok) { ws = (int)(row % a.wsegs); long long q = row / a.wsegs; h = (int)(q % a.H); q /= a.H; d = (int)(q % a.D); n = (int)(q / a.D); }
        if (isA) {
          const __nv_bfloat16* src = a.dy + ((((long long)n * a.D + d) * a.H + h) * a.W + ws * RW) * a.Cdy + co0 + gidx * 8;
          unsigned dst = sa + gidx * A_GPITCH + rr * (RW * 16);
#pragma unroll
          for (int v = 0; v < RW; ++v) {
            const bool ok = row_ok && ws * RW + v < a.W;
            cp_async16(dst, ok ? src : a.dy, ok);
            dst += 16; src += a.Cdy;
          }
        } else if (SW == 1) {
          const int dd = d + dz, hh = h + dyo;
          row_ok = row_ok && (unsigned)dd < (unsigned)a.D && (unsigned)hh < (unsigned)a.H;
          const __nv_bfloat16* src = a.x + ((((long long)n * a.D + dd) * a.H + hh) * a.W + (ws * RW - 1)) * a.Cx + ci0 + gidx * 8;
          unsigned dst = sb + gidx * B_GPITCH + rr * (XW * 16);
#pragma unroll
          for (int v = 0; v < XW; ++v) {
            const bool ok = row_ok && (unsigned)(ws * RW - 1 + v) < (unsigned)a.W;
            cp_async16(dst, ok ? src : a.x, ok);
            dst += 16; src += a.Cx;
          }
        } else {
          const int dd = d * a.sd + dz, hh = h * a.sh + dyo;
          row_ok = row_ok && (unsigned)dd < (unsigned)a.Di && (unsigned)hh < (unsigned)a.Hi;
          const int w_in0 = ws * RW * 2 - 1;                       // input voxel of slot v = 0 (odd plane, p = 0)
          const __nv_bfloat16* src = a.x + ((((long long)n * a.Di + dd) * a.Hi + hh) * a.Wi + w_in0) * a.Cx + ci0 + gidx * 8;
          const unsigned dst0 = sb + gidx * B_GPITCH + rr * (XW * 16);
#pragma unroll
          for (int v = 0; v < XW; ++v) {                           // v-th input voxel of the row: even v -> odd plane, odd v -> even plane
            const bool ok = row_ok && (unsigned)(w_in0 + v) < (unsigned)a.Wi;
            const int slot = (v & 1) ? (RW + 1) + (v >> 1) : (v >> 1);
            cp_async16(dst0 + slot * 16, ok ? src : a.x, ok);
            src += a.Cx;
          }
        }
      }
      cp_async_commit();
      ++pending;
      if (pending > LAGW) {
        cp_async_wait<LAGW>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(FULL(done_stage));
        done_stage = (done_stage + 1 == STAGES) ? 0 : done_stage + 1;
        --pending;
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    while (pending > 0) {
      if (lane == 0) mbar_arrive(FULL(done_stage));
      done_stage = (done_stage + 1 == STAGES) ? 0 : done_stage + 1;
      --pending;
    }
  } else if (warp < 7) {
    // ================================================================ MMA issuers: one per dx tap
    if (lane == 0) {
      const int tap = warp - 4;                       // dx = tap - 1 -> x start offset tap * 16 bytes
      const bool present = a.gtw[grp][tap] != 255;
      unsigned stage = 0, phase = 0;
      const unsigned d_tmem = tmem_base + tap * N_T;
      for (int s = 0; s < n_stages; ++s) {
        mbar_wait(FULL(stage), phase);
        tc_fence_after();
        if (present) {
          const unsigned sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
          const unsigned long long a0 = make_desc(sa, 128, A_GPITCH);
          // start of the 16 voxel rows tap dx = tap - 1 reads: stride 1 -> the halo row shifted by tap; stride 2 -> odd plane
          // (slot 0), even plane (slot 17), odd plane shifted by one (slot 1)
          const unsigned tap_off = SW == 1 ? tap * 16 : (tap == 1 ? (RW + 1) * 16 : (tap >> 1) * 16);
          const unsigned long long b0 = make_desc(sb + tap_off, 128, B_GPITCH);
#pragma unroll
          for (int rr = 0; rr < ROWS; ++rr)
            tc_mma(d_tmem, a0 + (unsigned long long)((rr * RW * 16) >> 4), b0 + (unsigned long long)((rr * XW * 16) >> 4), IDESC,
                   (s | rr) != 0 ? 1u : 0u);
        }
        tc_commit(EMPTY(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit(DONE);
    }
  } else {
    // ================================================================ epilogue: TMEM -> fp32 atomics into dW
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait_warp_backoff(DONE, 0, lane, 1000);   // the wait lasts the whole kernel
    tc_fence_after();
#pragma unroll 1
    for (int tap = 0; tap < 3; ++tap) {
      const int tw = a.gtw[grp][tap];
      if (tw == 255) continue;
      float* dwt = a.dw + (long long)tw * a.s_tap + (long long)co * a.s_co;
#pragma unroll 1
      for (int c = 0; c < N_T / 32; ++c) {
        unsigned v[32];
        tmem_ld32(tmem_base + ((unsigned)(q * 32) << 16) + tap * N_T + c * 32, v);
        if (co < a.Cout) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int ci = ci0 + c * 32 + j;
            if (ci < a.Cin) atomicAdd(dwt + (long long)ci * a.s_ci, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

template <int N_T, int SW>
int launch_wt(WtArgs a, int co_pad, int ci_pad, cudaStream_t st) {
  constexpr int XW = SW == 1 ? RW + 2 : 2 * RW + 1;
  const int co_tiles = (co_pad + M_T - 1) / M_T;
  a.ci_tiles = ci_pad / N_T;
  const long long tiles = (long long)a.n_groups * co_tiles * a.ci_tiles;
  long long splits = NND_NUM_SMS / tiles;            // one wave of 1-CTA-per-SM blocks
  const long long max_splits = (a.total_rows + 4 * ROWS - 1) / (4 * ROWS);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.rows_per_split = (a.total_rows + splits - 1) / splits;
  a.rows_per_split = (a.rows_per_split + ROWS - 1) / ROWS * ROWS;
  splits = (a.total_rows + a.rows_per_split - 1) / a.rows_per_split;
  constexpr size_t SMEM = (size_t)STAGES * ((M_T / 8) * ROWS * RW * 16 + (N_T / 8) * ROWS * XW * 16) + 8 * (2 * STAGES + 1);
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_tc_kernel<N_T, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  dim3 grid((unsigned)splits, (unsigned)a.n_groups, (unsigned)(co_tiles * a.ci_tiles));
  conv_wgrad_tc_kernel<N_T, SW><<<grid, WG_THREADS, SMEM, st>>>(a);
  NND_LAUNCH_CHECK("conv_wgrad_tc_kernel");
  return NND_OK;
}

}  // namespace

int nnd_conv_wgrad_tc_supported(const ConvGeom& g, int Cdy, int Cx) {
  if (g.sd != 1 || g.sh != 1 || g.sw != 1 || g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return 0;
  if (g.Ld != g.Di || g.Lh != g.Hi || g.Lw != g.Wi || g.Do != g.Di || g.Ho != g.Hi || g.Wo != g.Wi) return 0;
  // 3x3x3 / 1x3x3 filters, or the single tap of a 1x1x1 convolution with >= 64 output channels (laterals: one filter-row group, only
  // the centre issuer works)
  const bool pointwise = g.T == 1 && g.off_d[0] == 0 && g.off_h[0] == 0 && g.off_w[0] == 0;
  if ((g.T < 9 && !pointwise) || Cdy % 32 || Cx % 32) return 0;
  // 32-channel dy fills a quarter of the 128-row MMA; for the pointwise form its 64-voxel pipeline stages are also far too small
  // (measured: 32->32 @128^3 0.87 ms here against 0.47 ms on the mma.sync kernel, 64->64 @64^3 0.13 against 0.15)
  if (Cdy < 64) return 0;
  for (int t = 0; t < g.T; ++t)
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
  return 1;
}

// Stride-2 (in w; 1 or 2 in d / h) convolutions with taps in [-1, 1]: the de-interleaved variant (SW = 2).  Also serves the
// weight gradient of kernel == stride transposed convolutions with the operands' roles swapped by the caller (dense operand =
// the layer input, strided operand = dy read at 2i + {0, 1}: taps (0 / 1, 0 / 1, 0 / 1), arch/conv.py).
int nnd_conv_wgrad_tc_strided_supported(const ConvGeom& g, int Cdy, int Cx) {
  if (g.sw != 2 || g.sd < 1 || g.sd > 2 || g.sh < 1 || g.sh > 2) return 0;
  if (g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return 0;
  if (g.Do != g.Ld || g.Ho != g.Lh || g.Wo != g.Lw) return 0;
  if (g.T < 4 || Cdy % 32 || Cx % 32 || Cdy < 64) return 0;
  for (int t = 0; t < g.T; ++t)
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
  return 1;
}

int nnd_conv_wgrad_tc(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                      long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st) {
  const bool strided = g.sw == 2;
  WtArgs a;
  a.dy = dy; a.Cdy = Cdy; a.x = x; a.Cx = Cx; a.dw = dw; a.s_co = s_co; a.s_ci = s_ci; a.s_tap = s_tap;
  a.Cout = Cout; a.Cin = Cin; a.N = g.N; a.D = g.Ld; a.H = g.Lh; a.W = g.Lw;      // dy grid (= the x grid at stride 1)
  a.Di = g.Di; a.Hi = g.Hi; a.Wi = g.Wi; a.sd = g.sd; a.sh = g.sh;
  a.wsegs = (g.Lw + RW - 1) / RW;
  a.total_rows = (long long)g.N * g.Ld * g.Lh * a.wsegs;
  if (a.total_rows <= 0) return NND_OK;
  a.n_groups = 0;
  for (int z = -1; z <= 1; ++z)
    for (int y = -1; y <= 1; ++y) {
      int found = 0;
      unsigned char tw[3] = {255, 255, 255};
      for (int t = 0; t < g.T; ++t)
        if (g.off_d[t] == z && g.off_h[t] == y) { tw[g.off_w[t] + 1] = g.tap_w[t]; found = 1; }
      if (found) {
        a.gdz[a.n_groups] = (signed char)z; a.gdy[a.n_groups] = (signed char)y;
        for (int k = 0; k < 3; ++k) a.gtw[a.n_groups][k] = tw[k];
        ++a.n_groups;
      }
    }
  if (strided) {
    if (Cx % 128 == 0) return launch_wt<128, 2>(a, Cdy, Cx, st);
    if (Cx % 64 == 0) return launch_wt<64, 2>(a, Cdy, Cx, st);
    return launch_wt<32, 2>(a, Cdy, Cx, st);
  }
  if (Cx % 128 == 0) return launch_wt<128, 1>(a, Cdy, Cx, st);
  if (Cx % 64 == 0) return launch_wt<64, 1>(a, Cdy, Cx, st);
  return launch_wt<32, 1>(a, Cdy, Cx, st);
}
