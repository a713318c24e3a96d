// Weight gradient of the 32 -> 32 channel 3x3x3 stride-1 convolutions (the two full-resolution layers) on tcgen05.
//
//   dW[tz,ty,tx][co][ci] = sum_v dy[v][co] * x[v + (tz,ty,tx)][ci]         (autograd of nndet/arch/conv.py:344-348)
//
// A 32 x 32 output per tap would fill a quarter of a 128-row MMA and, with both operands in shared memory, an
// 128 x 32 x 16 MMA costs 40 cycles instead of 16 (scripts/mma_rate.cu).  So the taps are stacked into the MMA shape:
// substituting z' = z + tz moves the z shift onto dy, and
//     D[(tz, co)][(ty, ci)] += sum_k  dy[z'-tz, y, x0+k][co] * x[z', y+ty, x0+k+tx][ci]
// is ONE 128 x 96 x 16 MMA per dx tap (rows: dy slices z'-1, z', z'+1 (+1 unused), columns: x rows y-1, y, y+1; the dx
// tap is a 16-byte start offset into the x row): 3 MMAs per 16 voxels instead of 27, 75 % of the rows useful.
// Both operands are MN-major (HBM layout [voxel][channel]): staged as [..][channel group][voxel][8 ch], so the row
// groups (slice, channel group) / (y row, channel group) have ONE constant pitch and a single descriptor covers them.
// CTA = persistent over tiles of 2 (z') x 8 (y) x 16 (x) voxels, 3-stage ring (4 x 8 tiles with 2 stages were slower: 0.87 vs 0.58 ms); 4 producer warps (cp.async, zero fill = padding), 3
// issuer warps (one per dx tap = independent accumulators), 4 epilogue warps (TMEM -> fp32 atomics into dW once at the end).
#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int C = 32, KG = 4;             // channels (both operands), 8-channel groups
constexpr int ZT = 2, YT = 8, RW = 16, XW = RW + 2;
constexpr int A_ROW = RW * 16;            // 256 B: one channel group of one dy row
constexpr int A_SLICE = KG * A_ROW;       // 1024 B
constexpr int A_Y = (ZT + 2) * A_SLICE;   // 4096 B: slices z0-1 .. z0+2 of one y
constexpr int A_BYTES = YT * A_Y;         // 32768
constexpr int B_ROW = XW * 16;            // 288 B: one channel group of one x row (with halo)
constexpr int B_Y = KG * B_ROW;           // 1152 B
constexpr int B_Z = (YT + 2) * B_Y;       // 11520 B
constexpr int B_BYTES = ZT * B_Z;         // 23040
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 55808 (multiple of 128)
constexpr int STAGES = 3;
constexpr int NCOL = 3 * C;               // 96 accumulator columns per dx tap
constexpr int THREADS = (4 + 3 + 4) * 32;


struct W32Args {
  const __nv_bfloat16* dy; const __nv_bfloat16* x;
  float* dw; long long s_co, s_ci, s_tap;
  int Cout, Cin;
  int N, D, H, W;
  int ZB, YB, XB;                    // tiles along z (2 slices), y (8 rows), x (16 voxels)
  int total;                         // N * ZB * YB * XB
  unsigned char tw[27];              // weight tap of offsets (tz, ty, tx), index (tz+1)*9 + (ty+1)*3 + (tx+1); 255 = absent
};

__global__ void __launch_bounds__(THREADS, 1) conv_wgrad_tc32_kernel(const W32Args a) {
  // kind::f16, D fp32, A/B bf16, both MN-major (bits 15, 16), N = 96, M = 128
  constexpr unsigned IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(NCOL >> 3) << 17) | ((128u >> 4) << 24);
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + STAGES * STAGE_BYTES + 1024);   // +1 KB: the unused 4th
  __shared__ unsigned s_tmem_base;                                                                      // slice of the last y row
  const unsigned bar0 = smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (STAGES + i); };
  const unsigned DONE = bar0 + 8u * (2 * STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int my_tiles = a.total > (int)blockIdx.x ? (a.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  for (int i = tid; i < (STAGES * STAGE_BYTES + 1024) / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(FULL(i), 4); mbar_init(EMPTY(i), 3); }
    mbar_init(DONE, 3);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = s_tmem_base;

  if (my_tiles > 0) {
    if (warp < 4) {
      // ================================================================ producers: one stage = one tile
      unsigned stage = 0, phase = 0, done_stage = 0;
      int pending = 0;
      constexpr int LAGW = 1;           // must stay <= STAGES - 2: a stage is published one stage late, and the next
                                        // load waits for the stage STAGES back to be consumed
      for (int t = 0; t < my_tiles; ++t) {
        int r = blockIdx.x + t * gridDim.x;
        const int xb = r % a.XB; r /= a.XB;
        const int yb = r % a.YB; r /= a.YB;
        const int zb = r % a.ZB; const int n = r / a.ZB;
        const int z0 = zb * ZT, y0 = yb * YT, x0 = xb * RW;
        mbar_wait_warp(EMPTY(stage), phase ^ 1, lane);
        const unsigned sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
        // dy rows: (yi, slice, channel group), 16 voxels each
        for (int rr = tid; rr < YT * (ZT + 2) * KG; rr += 128) {
          const int g = rr & 3; const int r2 = rr >> 2;
          const int si = r2 % (ZT + 2), yi = r2 / (ZT + 2);
          const int z = z0 - 1 + si, y = y0 + yi;
          const bool row_ok = (unsigned)z < (unsigned)a.D && y < a.H;
          const __nv_bfloat16* src = a.dy + ((((long long)n * a.D + z) * a.H + y) * a.W + x0) * C + g * 8;
          unsigned dst = sa + yi * A_Y + si * A_SLICE + g * A_ROW;
#pragma unroll
          for (int v = 0; v < RW; ++v) {
            const bool ok = row_ok && x0 + v < a.W;
            cp_async16(dst, ok ? src : a.dy, ok);
            dst += 16; src += C;
          }
        }
        // x rows: (z' slice, y row with halo, channel group), 18 voxels each
        for (int rr = tid; rr < ZT * (YT + 2) * KG; rr += 128) {
          const int g = rr & 3; const int r2 = rr >> 2;
          const int yr = r2 % (YT + 2), zi = r2 / (YT + 2);
          const int z = z0 + zi, y = y0 - 1 + yr;
          const bool row_ok = z < a.D && (unsigned)y < (unsigned)a.H;
          const __nv_bfloat16* src = a.x + ((((long long)n * a.D + z) * a.H + y) * a.W + (x0 - 1)) * C + g * 8;
          unsigned dst = sb + zi * B_Z + yr * B_Y + g * B_ROW;
#pragma unroll
          for (int v = 0; v < XW; ++v) {
            const bool ok = row_ok && (unsigned)(x0 - 1 + v) < (unsigned)a.W;
            cp_async16(dst, ok ? src : a.x, ok);
            dst += 16; src += C;
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
      // ================================================================ MMA issuers: one per dx tap (whole warp runs
      // the control flow, the elected lane issues)
      constexpr unsigned A_HI = (unsigned)((A_ROW >> 4) & 0x3FFF) | (1u << 14), B_HI = (unsigned)((B_ROW >> 4) & 0x3FFF) | (1u << 14);
      constexpr unsigned LO_HI = (unsigned)((128 >> 4) & 0x3FFF) << 16;
      const int tx = __shfl_sync(0xffffffffu, warp - 4, 0);
      unsigned stage = 0, phase = 0;
      const unsigned d_tmem = __shfl_sync(0xffffffffu, tmem_base, 0) + tx * NCOL;
      for (int t = 0; t < my_tiles; ++t) {
        mbar_wait_warp(FULL(stage), phase, lane);
        tc_fence_after();
        const unsigned sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
        const unsigned a_lo0 = ((sa >> 4) & 0x3FFF) | LO_HI, b_lo0 = (((sb + tx * 16) >> 4) & 0x3FFF) | LO_HI;
        if (elect_one()) {
#pragma unroll
          for (int zi = 0; zi < ZT; ++zi)
#pragma unroll
            for (int yi = 0; yi < YT; ++yi)
              tc_mma2(d_tmem, a_lo0 + ((yi * A_Y + zi * A_SLICE) >> 4), A_HI, b_lo0 + ((zi * B_Z + yi * B_Y) >> 4), B_HI, IDESC,
                      (t | zi | yi) != 0 ? 1u : 0u);
          tc_commit(EMPTY(stage));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(DONE);
      __syncwarp();
    } else {
      // ================================================================ epilogue: TMEM -> fp32 atomics into dW
      const int q = warp & 3;                      // TMEM lanes 32q..: row block b = q <-> tz = 1 - q (q = 3: unused rows)
      mbar_wait_warp_backoff(DONE, 0, lane, 2000);   // the wait lasts the whole kernel
      tc_fence_after();
      if (q < 3) {
        const int tz = 1 - q;
#pragma unroll 1
        for (int tx = 0; tx < 3; ++tx) {
#pragma unroll 1
          for (int ty = 0; ty < 3; ++ty) {
            unsigned v[32];
            tmem_ld32(tmem_base + ((unsigned)(q * 32) << 16) + tx * NCOL + ty * C, v);
            const int tw = a.tw[(tz + 1) * 9 + ty * 3 + tx];
            if (tw != 255 && lane < a.Cout) {
              float* dwt = a.dw + (long long)tw * a.s_tap + (long long)lane * a.s_co;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < a.Cin) atomicAdd(dwt + (long long)j * a.s_ci, __uint_as_float(v[j]));
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

int nnd_conv_wgrad_tc32_supported(const ConvGeom& g, int Cdy, int Cx) {
  if (g.sd != 1 || g.sh != 1 || g.sw != 1 || g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return 0;
  if (g.Ld != g.Di || g.Lh != g.Hi || g.Lw != g.Wi || g.Do != g.Di || g.Ho != g.Hi || g.Wo != g.Wi) return 0;
  if (Cdy != 32 || Cx != 32 || g.T < 9) return 0;
  for (int t = 0; t < g.T; ++t)
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
  return 1;
}

// worth it only when the volume fills the persistent grid with several tiles per SM
int nnd_conv_wgrad_tc32_profitable(const ConvGeom& g) {
  const long long tiles = (long long)g.N * ((g.Di + ZT - 1) / ZT) * ((g.Hi + YT - 1) / YT) * ((g.Wi + RW - 1) / RW);
  return tiles >= 4 * NND_NUM_SMS;
}

int nnd_conv_wgrad_tc32(const __nv_bfloat16* dy, const __nv_bfloat16* x, const ConvGeom& g, float* dw, long long s_co,
                        long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st) {
  W32Args a;
  a.dy = dy; a.x = x; a.dw = dw; a.s_co = s_co; a.s_ci = s_ci; a.s_tap = s_tap; a.Cout = Cout; a.Cin = Cin;
  a.N = g.N; a.D = g.Di; a.H = g.Hi; a.W = g.Wi;
  a.ZB = (g.Di + ZT - 1) / ZT; a.YB = (g.Hi + YT - 1) / YT; a.XB = (g.Wi + RW - 1) / RW;
  const long long total = (long long)a.N * a.ZB * a.YB * a.XB;
  if (total <= 0) return NND_OK;
  if (total > 0x7fffffffll) return NND_ERR_ARG;
  a.total = (int)total;
  for (int i = 0; i < 27; ++i) a.tw[i] = 255;
  for (int t = 0; t < g.T; ++t) a.tw[(g.off_d[t] + 1) * 9 + (g.off_h[t] + 1) * 3 + (g.off_w[t] + 1)] = g.tap_w[t];
  constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 8 * (2 * STAGES + 1);
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_tc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  const int grid = a.total < NND_NUM_SMS ? a.total : NND_NUM_SMS;
  conv_wgrad_tc32_kernel<<<grid, THREADS, SMEM, st>>>(a);
  NND_LAUNCH_CHECK("conv_wgrad_tc32_kernel");
  return NND_OK;
}
