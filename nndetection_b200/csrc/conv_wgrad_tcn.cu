// Weight gradient of the 3x3x3 stride-1 convolutions with 128 output channels (encoder stage 2, decoder P2..P5 outputs, the
// shared detection-head convolutions: 26 launches per train step) on tcgen05 -- "all taps, thin input-channel class" form.
//
//   dW[tz,ty,tx][co][ci] = sum_v dy[v][co] * x[v + (tz,ty,tx)][ci]         (autograd of nndet/arch/conv.py:344-348)
//
// conv_wgrad_tc.cu gives every CTA one filter row (dz, dy): dy and x are then re-read from L2 once per filter row -- ncu
// (profiles/r01_ncu_conv_wgrad_tc128_summary.txt): 1.22 GB of L2 -> SM traffic for 67 MB of operands, 59 % of the L2
// throughput, tcgen05 pipe 26 % busy.  Here a CTA keeps ALL 27 taps of a 16-input-channel class in TMEM (9 accumulators
// [128 co x (3 dy x 16 ci)] = 432 columns): with z' = z + tz the z shift sits on dy (three A slices), the dy taps are stacked
// along N (x rows y-1, y, y+1 of the 16-channel class have one group pitch), the dx taps are 16-byte start offsets.
// Per 16 voxels: 9 MMAs of 128 x 48 x 16; dy is read once per class (8 classes for 128 input channels) instead of once per
// filter row and tap column, x once (+ halo).  MEASURED (B200, 128 -> 128 @ 32^3 x 4): 0.383 ms vs 0.227 ms for the filter-row
// kernel -- the thin N = 48 MMAs hold the tcgen05 pipe 44 cycles for 24 cycles of math and 8 classes x 9 MMAs per 16 voxels
// is 1.8x more pipe time than 27 N = 128 MMAs; the kernel is therefore OPT-IN (nnd_conv_set_wgrad_tc(4)), kept with its parity
// test as the record of the experiment.  Roles as in conv_wgrad_tc32.cu: 4 producer warps (cp.async, zero fill =
// padding), 3 issuer warps (one per dx tap), 4 epilogue warps (TMEM -> fp32 atomics into dW at the end).
#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int CO = 128, COG = 16;          // output channels (MMA M) and their 8-channel groups
constexpr int CI = 16, CIG = 2;            // input-channel class
constexpr int ZT = 2, YT = 4, RW = 16, XW = RW + 2;
constexpr int A_ROW = RW * 16;             // 256 B: one channel group of one dy row
constexpr int A_SLICE = COG * A_ROW;       // 4096 B
constexpr int A_Y = (ZT + 2) * A_SLICE;    // 16384 B: dy slices z0-1 .. z0+2 of one y
constexpr int A_BYTES = YT * A_Y;          // 65536
constexpr int B_ROW = XW * 16;             // 288 B: one channel group of one x row (with halo)
constexpr int B_Y = CIG * B_ROW;           // 576 B
constexpr int B_Z = (YT + 2) * B_Y;        // 3456 B
constexpr int B_BYTES = ZT * B_Z;          // 6912
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 72448
constexpr int STAGES = 3;
constexpr int NCOL = 3 * CI;               // 48 accumulator columns per (tz, tx)
constexpr int THREADS = (4 + 3 + 4) * 32;


struct WnArgs {
  const __nv_bfloat16* dy; const __nv_bfloat16* x;
  float* dw; long long s_co, s_ci, s_tap;
  int Cout, Cin, Cx;                 // real channel counts (store mask) and the channel stride of x
  int N, D, H, W;
  int ZB, YB, XB;                    // tiles along z (2 slices), y (4 rows), x (16 voxels)
  int total;                         // N * ZB * YB * XB
  unsigned char tw[27];              // weight tap of offsets (tz, ty, tx), index (tz+1)*9 + (ty+1)*3 + (tx+1); 255 = absent
};

__global__ void __launch_bounds__(THREADS, 1) conv_wgrad_tcn_kernel(const WnArgs a) {
  // kind::f16, D fp32, A/B bf16, both MN-major (bits 15, 16), N = 48, M = 128
  constexpr unsigned IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(NCOL >> 3) << 17) | ((128u >> 4) << 24);
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + STAGES * STAGE_BYTES);
  __shared__ unsigned s_tmem_base;
  const unsigned bar0 = smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (STAGES + i); };
  const unsigned DONE = bar0 + 8u * (2 * STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ci0 = blockIdx.y * CI;                 // input-channel class of this CTA
  const int my_tiles = a.total > (int)blockIdx.x ? (a.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(FULL(i), 4); mbar_init(EMPTY(i), 3); }
    mbar_init(DONE, 3);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = s_tmem_base;

  if (my_tiles > 0) {
    if (warp < 4) {
      // ================================================================ producers: one stage = one tile
      unsigned stage = 0, phase = 0, done_stage = 0;
      int pending = 0;
      constexpr int LAGW = 1;           // <= STAGES - 2 (see conv_wgrad_tc32.cu)
      for (int t = 0; t < my_tiles; ++t) {
        int r = blockIdx.x + t * gridDim.x;
        const int xb = r % a.XB; r /= a.XB;
        const int yb = r % a.YB; r /= a.YB;
        const int zb = r % a.ZB; const int n = r / a.ZB;
        const int z0 = zb * ZT, y0 = yb * YT, x0 = xb * RW;
        mbar_wait_warp(EMPTY(stage), phase ^ 1, lane);
        const unsigned sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
        // dy rows: (yi, slice, channel group), 16 voxels each
        for (int rr = tid; rr < YT * (ZT + 2) * COG; rr += 128) {
          const int g = rr % COG; const int r2 = rr / COG;
          const int si = r2 % (ZT + 2), yi = r2 / (ZT + 2);
          const int z = z0 - 1 + si, y = y0 + yi;
          const bool row_ok = (unsigned)z < (unsigned)a.D && y < a.H;
          const __nv_bfloat16* src = a.dy + ((((long long)n * a.D + z) * a.H + y) * a.W + x0) * CO + g * 8;
          unsigned dst = sa + yi * A_Y + si * A_SLICE + g * A_ROW;
#pragma unroll
          for (int v = 0; v < RW; ++v) {
            const bool ok = row_ok && x0 + v < a.W;
            cp_async16(dst, ok ? src : a.dy, ok);
            dst += 16; src += CO;
          }
        }
        // x rows of the 16-channel class: (z' slice, y row with halo, channel group), 18 voxels each
        for (int rr = tid; rr < ZT * (YT + 2) * CIG; rr += 128) {
          const int g = rr % CIG; const int r2 = rr / CIG;
          const int yr = r2 % (YT + 2), zi = r2 / (YT + 2);
          const int z = z0 + zi, y = y0 - 1 + yr;
          const bool row_ok = z < a.D && (unsigned)y < (unsigned)a.H;
          const __nv_bfloat16* src = a.x + ((((long long)n * a.D + z) * a.H + y) * a.W + (x0 - 1)) * a.Cx + ci0 + g * 8;
          unsigned dst = sb + zi * B_Z + yr * B_Y + g * B_ROW;
#pragma unroll
          for (int v = 0; v < XW; ++v) {
            const bool ok = row_ok && (unsigned)(x0 - 1 + v) < (unsigned)a.W;
            cp_async16(dst, ok ? src : a.x, ok);
            dst += 16; src += a.Cx;
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
      // ================================================================ MMA issuers: one per dx tap, three accumulators
      // (tz = +1, 0, -1 <-> dy slice z'-1, z', z'+1) each
      constexpr unsigned A_HI = (unsigned)((A_ROW >> 4) & 0x3FFF) | (1u << 14), B_HI = (unsigned)((B_ROW >> 4) & 0x3FFF) | (1u << 14);
      constexpr unsigned LO_HI = (unsigned)((128 >> 4) & 0x3FFF) << 16;
      const int tx = __shfl_sync(0xffffffffu, warp - 4, 0);
      unsigned stage = 0, phase = 0;
      const unsigned d_tmem = __shfl_sync(0xffffffffu, tmem_base, 0) + tx * NCOL;       // + b * 3 * NCOL per dy slice b
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
#pragma unroll
              for (int b = 0; b < 3; ++b)
                tc_mma2(d_tmem + b * 3 * NCOL, a_lo0 + ((yi * A_Y + (zi + b) * A_SLICE) >> 4), A_HI,
                        b_lo0 + ((zi * B_Z + yi * B_Y) >> 4), B_HI, IDESC, (t | zi | yi) != 0 ? 1u : 0u);
          tc_commit(EMPTY(stage));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(DONE);
      __syncwarp();
    } else {
      // ================================================================ epilogue: TMEM -> fp32 atomics into dW
      const int q = warp & 3;
      const int co = q * 32 + lane;                // TMEM lane = output channel
      mbar_wait_warp_backoff(DONE, 0, lane, 2000);   // the wait lasts the whole kernel
      tc_fence_after();
#pragma unroll 1
      for (int b = 0; b < 3; ++b) {
#pragma unroll 1
        for (int tx = 0; tx < 3; ++tx) {
          // 48 columns = (ty, 16 ci): two 32-column loads, the upper half of the second one is not used (<= column 448)
          unsigned v[64];
          const unsigned col0 = tmem_base + ((unsigned)(q * 32) << 16) + (b * 3 + tx) * NCOL;
          tmem_ld32(col0, v);
          tmem_ld32(col0 + 32, v + 32);
          if (co < a.Cout) {
#pragma unroll
            for (int ty = 0; ty < 3; ++ty) {
              const int tw = a.tw[(2 - b) * 9 + ty * 3 + tx];           // slice b <-> tz = 1 - b
              if (tw == 255) continue;
              float* dwt = a.dw + (long long)tw * a.s_tap + (long long)co * a.s_co;
#pragma unroll
              for (int j = 0; j < CI; ++j) {
                const int c = ty * CI + j;                                // column inside the 48-wide accumulator
                const float val = __uint_as_float(v[c]);
                if (ci0 + j < a.Cin) atomicAdd(dwt + (long long)(ci0 + j) * a.s_ci, val);
              }
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

int nnd_conv_wgrad_tcn_supported(const ConvGeom& g, int Cdy, int Cx) {
  if (g.sd != 1 || g.sh != 1 || g.sw != 1 || g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return 0;
  if (g.Ld != g.Di || g.Lh != g.Hi || g.Lw != g.Wi || g.Do != g.Di || g.Ho != g.Hi || g.Wo != g.Wi) return 0;
  if (Cdy != CO || Cx % CI || Cx > 512 || g.T < 9) return 0;
  for (int t = 0; t < g.T; ++t)
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
  return 1;
}

int nnd_conv_wgrad_tcn(const __nv_bfloat16* dy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw, long long s_co,
                       long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st) {
  WnArgs a;
  a.dy = dy; a.x = x; a.dw = dw; a.s_co = s_co; a.s_ci = s_ci; a.s_tap = s_tap; a.Cout = Cout; a.Cin = Cin; a.Cx = Cx;
  a.N = g.N; a.D = g.Di; a.H = g.Hi; a.W = g.Wi;
  a.ZB = (g.Di + ZT - 1) / ZT; a.YB = (g.Hi + YT - 1) / YT; a.XB = (g.Wi + RW - 1) / RW;
  const long long total = (long long)a.N * a.ZB * a.YB * a.XB;
  if (total <= 0) return NND_OK;
  if (total > 0x7fffffffll) return NND_ERR_ARG;
  a.total = (int)total;
  for (int i = 0; i < 27; ++i) a.tw[i] = 255;
  for (int t = 0; t < g.T; ++t) a.tw[(g.off_d[t] + 1) * 9 + (g.off_h[t] + 1) * 3 + (g.off_w[t] + 1)] = g.tap_w[t];
  constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1);
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_tcn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  const int classes = Cx / CI;
  int per_class = NND_NUM_SMS / classes; if (per_class < 1) per_class = 1;
  if (per_class > a.total) per_class = a.total;
  dim3 grid((unsigned)per_class, (unsigned)classes);
  conv_wgrad_tcn_kernel<<<grid, THREADS, SMEM, st>>>(a);
  NND_LAUNCH_CHECK("conv_wgrad_tcn_kernel");
  return NND_OK;
}
