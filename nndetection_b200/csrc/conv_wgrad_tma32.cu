// Weight gradient of the 32 -> 32 channel 3x3x3 stride-1 convolutions (the two full-resolution layers): the stacked-tap scheme of
// conv_wgrad_tc32.cu with TMA-fed operands.
//
//   dW[tz,ty,tx][co][ci] = sum_v dy[v][co] * x[v + (tz,ty,tx)][ci]         (autograd of nndet/arch/conv.py:344-348)
//
// Substituting z' = z + tz moves the z shift onto dy:   D[(tz, co)][(ty, ci)] += sum_k dy[z'-tz, y, x0+k][co] * x[z', y+ty, x0+k+tx][ci]
// = ONE 128 x 96 x 16 MMA per dx tap (rows: four dy slices z'-1 .. z'+2, the last one unused; columns: x rows y-1, y, y+1): 3 MMAs per 16
// voxels for all 27 taps.  conv_wgrad_tc32.cu stages the operands with one 16-byte cp.async per thread (LSU-bound: 0.60 ms = 776
// TFLOP/s against an MMA bound of 0.31 ms); here a tile of 2 (z') x 8 (y) x 16 (x) voxels is TWO tensor-map boxes,
//   A = dy [32 ch x 16 w x 8 h x 4 d]   (slices z0-1 .. z0+2),      B = x [32 ch x 18 w x 10 h x 2 d]   (halo in w and h),
// in the MN-major SWIZZLE_64B layout (a voxel = a 64-byte K row); the stacks are descriptor strides: M blocks = dy slices (LBO = one
// slice of the box), N blocks = x rows (LBO = one 18-voxel row of the box), the dx tap = the start row -- the row-shifted-start property
// verified on the device for conv_wgrad_tma.cu.  Out-of-bounds zero fill = padding and ragged edges.
// One CTA owns ALL 27 taps (3 x 96 accumulator columns) for a contiguous range of tiles: every voxel is read once; split-K partials per
// CTA in the caller's workspace + finishing pass (or fp32 atomics without one).  Roles: warp 0 TMA producer | 2 issuer warps on
// alternate stages | 4 epilogue warps.
#include <cuda.h>

#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int Q_NI = 2;
constexpr int Q_THREADS = (1 + Q_NI + 4) * 32;
constexpr int Q_ZT = 2, Q_YT = 8, Q_RW = 16, Q_XW = Q_RW + 2;
constexpr int Q_ROWB = 64;                                        // 32 channels x 2 bytes
constexpr int Q_A_BYTES = (Q_ZT + 2) * Q_YT * Q_RW * Q_ROWB;      // 32768
constexpr int Q_B_BYTES = Q_ZT * (Q_YT + 2) * Q_XW * Q_ROWB;      // 23040
constexpr int Q_B_PAD = (Q_B_BYTES + 1023) / 1024 * 1024;         // 23552
constexpr int Q_STAGE = Q_A_BYTES + Q_B_PAD;                      // 56320
constexpr int Q_STAGES = 3;
constexpr int Q_TAIL = 8192;                                      // the unused fourth M block of the last K-steps reads up to one slice further
constexpr int Q_NCOL = 96;

struct QArgs {
  float* dw; long long s_co, s_ci, s_tap;
  float* part; int T;
  int Cout, Cin;
  int N, D, H, W;
  int ZB, YB, XB;
  int total, tiles_per_cta;
  unsigned char tw[27];              // weight tap of offsets (tz, ty, tx), index (tz+1)*9 + (ty+1)*3 + (tx+1); 255 = absent
};

struct QMaps { CUtensorMap dy, x; };

__device__ __forceinline__ void q_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void q_tma_5d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
// MN-major SWIZZLE_64B descriptor words (layout type 4): lo = start >> 4 | (LBO >> 4) << 16, hi = SBO >> 4 | version 1 << 14 | 4 << 29
__device__ __forceinline__ unsigned q_lo(unsigned start, unsigned lbo) { return ((start >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }

struct QCursor {
  int xb, yb, zb, n;
  __device__ __forceinline__ void init(int tile, const QArgs& a) {
    unsigned v = (unsigned)tile;
    xb = (int)(v % (unsigned)a.XB); v /= (unsigned)a.XB;
    yb = (int)(v % (unsigned)a.YB); v /= (unsigned)a.YB;
    zb = (int)(v % (unsigned)a.ZB); n = (int)(v / (unsigned)a.ZB);
  }
  __device__ __forceinline__ void next(const QArgs& a) {
    if (++xb == a.XB) { xb = 0; if (++yb == a.YB) { yb = 0; if (++zb == a.ZB) { zb = 0; ++n; } } }
  }
};

__global__ void __launch_bounds__(Q_THREADS, 1) conv_wgrad_tma32_kernel(const __grid_constant__ QMaps maps, const QArgs a) {
  // kind::f16, D fp32, A/B bf16, both MN-major (bits 15, 16), N = 96, M = 128
  constexpr unsigned IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(Q_NCOL >> 3) << 17) | ((128u >> 4) << 24);
  constexpr unsigned HI = (unsigned)((8 * Q_ROWB) >> 4) | (1u << 14) | (4u << 29);           // SBO = 8 voxels, SWIZZLE_64B
  constexpr unsigned LBO_A = Q_YT * Q_RW * Q_ROWB;                                            // one dy slice of the box: 8192 B
  constexpr unsigned LBO_B = Q_XW * Q_ROWB;                                                   // one x row of the box: 1152 B
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long bars[2 * Q_STAGES + 1];
  __shared__ unsigned s_tmem_base;
  const unsigned bar0 = smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (Q_STAGES + i); };
  const unsigned DONE = bar0 + 8u * (2 * Q_STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = (int)blockIdx.x * a.tiles_per_cta;
  const int t1 = min(t0 + a.tiles_per_cta, a.total);
  const int my_tiles = t1 > t0 ? t1 - t0 : 0;

  // the region behind the last stage is read (never used) by the fourth M block: keep it finite
  for (int i = tid; i < Q_TAIL / 16; i += Q_THREADS) reinterpret_cast<uint4*>(smem + Q_STAGES * Q_STAGE)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < Q_STAGES; ++i) { mbar_init(FULL(i), 1); mbar_init(EMPTY(i), 1); }
    mbar_init(DONE, Q_NI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = s_tmem_base;
  if (warp >= 1 + Q_NI) {                          // accumulators start at zero: every MMA accumulates, an idle CTA adds zeros
    const int q = warp & 3;
    for (int c = 0; c < 3 * Q_NCOL; c += 32) tmem_zero32(tmem_base + ((unsigned)(q * 32) << 16) + c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ================================================================ TMA producer: two boxes per tile
    if (lane == 0) {
      unsigned stage = 0, phase = 0;
      QCursor it;
      it.init(t0, a);
      for (int i = 0; i < my_tiles; ++i, it.next(a)) {
        const int x0 = it.xb * Q_RW, y0 = it.yb * Q_YT, z0 = it.zb * Q_ZT;
        mbar_wait(EMPTY(stage), phase ^ 1);
        const unsigned sa = smem_u32(smem + (size_t)stage * Q_STAGE);
        q_expect_tx(FULL(stage), Q_A_BYTES + Q_B_BYTES);
        q_tma_5d(sa, &maps.dy, 0, x0, y0, z0 - 1, it.n, FULL(stage));
        q_tma_5d(sa + Q_A_BYTES, &maps.x, 0, x0 - 1, y0 - 1, z0, it.n, FULL(stage));
        if (++stage == Q_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp <= Q_NI) {
    // ================================================================ MMA issuers (alternate stages; every MMA accumulates)
    const int me = warp - 1;
    unsigned stage = 0, phase = 0;
    const unsigned tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const unsigned smem0 = smem_u32(smem);
    for (int i = 0; i < my_tiles; ++i) {
      if ((i % Q_NI) == me) {
        mbar_wait_warp(FULL(stage), phase, lane);
        tc_fence_after();
        if (elect_one()) {
          const unsigned sa = smem0 + stage * Q_STAGE, sb = sa + Q_A_BYTES;
#pragma unroll
          for (int zl = 0; zl < Q_ZT; ++zl)
#pragma unroll
            for (int y = 0; y < Q_YT; ++y) {
              // K-step = the 16 voxels (z0 + zl, y0 + y, x0 .. x0+15): dy slices zl .. zl+3 of the box, x rows y .. y+2 of slice zl
              const unsigned a_lo = q_lo(sa + (zl * Q_YT + y) * (Q_RW * Q_ROWB), LBO_A);
              const unsigned b0 = sb + ((zl * (Q_YT + 2) + y) * Q_XW) * Q_ROWB;
#pragma unroll
              for (int t = 0; t < 3; ++t) tc_mma_acc2(tm + t * Q_NCOL, a_lo, HI, q_lo(b0 + t * Q_ROWB, LBO_B), HI, IDESC);
            }
          tc_commit(EMPTY(stage));
        }
        __syncwarp();
      }
      if (++stage == Q_STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) tc_commit(DONE);
    __syncwarp();
  } else {
    // ================================================================ epilogue: TMEM lanes 32 m .. 32 m + 31 = dy slice block m <-> tz = 1 - m;
    // accumulator t = dx tap t - 1, its column block n = x row block <-> ty = n - 1
    const int q = warp & 3;
    mbar_wait_warp_backoff(DONE, 0, lane, 1000);
    tc_fence_after();
    if (q < 3) {
      const int tz = 1 - q, co = lane;
#pragma unroll 1
      for (int t = 0; t < 3; ++t) {
#pragma unroll 1
        for (int n = 0; n < 3; ++n) {
          const int tw = a.tw[(tz + 1) * 9 + n * 3 + t];
          if (tw == 255) continue;
          unsigned v[32];
          tmem_ld32(tmem_base + ((unsigned)(q * 32) << 16) + t * Q_NCOL + n * 32, v);
          if (co < a.Cout) {
            if (a.part) {
              float* pt = a.part + (((long long)blockIdx.x * a.T + tw) * a.Cout + co) * a.Cin;
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(pt + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              float* dwt = a.dw + (long long)tw * a.s_tap + (long long)co * a.s_co;
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
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// dW[tap][co][ci] += sum over the CTAs' partials (see wgrad_finish_kernel in conv_wgrad_tma.cu)
__global__ void __launch_bounds__(256)
wgrad32_finish_kernel(const float* __restrict__ part, int splits, long long block, int Cout, int Cin, float* __restrict__ dw,
                      long long s_co, long long s_ci, long long s_tap) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= block) return;
  float4 acc = *reinterpret_cast<const float4*>(part + i4 * 4);
  for (int s = 1; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(part + (long long)s * block + i4 * 4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  long long r = i4 * 4;
  const int ci = (int)(r % Cin); r /= Cin;
  const int co = (int)(r % Cout); const int tap = (int)(r / Cout);
  float* d = dw + tap * s_tap + co * s_co + ci * s_ci;
  atomicAdd(d, acc.x); atomicAdd(d + s_ci, acc.y); atomicAdd(d + 2 * s_ci, acc.z); atomicAdd(d + 3 * s_ci, acc.w);
}

typedef CUresult (*QEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
QEncodeTiledFn q_encode_fn() {
  static QEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<QEncodeTiledFn>(p);
  }
  return fn;
}

int q_make_map(CUtensorMap* map, const void* base, int N, int D, int H, int W, int bw, int bh, int bd) {
  const QEncodeTiledFn enc = q_encode_fn();
  if (!enc) return NND_ERR_CUDA;
  const cuuint64_t gdim[5] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t gstride[4] = {64, (cuuint64_t)W * 64, (cuuint64_t)H * W * 64, (cuuint64_t)D * H * W * 64};
  const cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NND_OK : NND_ERR_CUDA;
}

void q_plan(const ConvGeom& g, QArgs& a, int& grid) {
  a.N = g.N; a.D = g.Di; a.H = g.Hi; a.W = g.Wi;
  a.ZB = (g.Di + Q_ZT - 1) / Q_ZT; a.YB = (g.Hi + Q_YT - 1) / Q_YT; a.XB = (g.Wi + Q_RW - 1) / Q_RW;
  const long long total = (long long)a.N * a.ZB * a.YB * a.XB;
  a.total = (int)total;
  grid = total < NND_NUM_SMS ? (int)total : NND_NUM_SMS;
  a.tiles_per_cta = grid > 0 ? (int)((total + grid - 1) / grid) : 1;
  grid = grid > 0 ? (int)((total + a.tiles_per_cta - 1) / a.tiles_per_cta) : 0;
}

}  // namespace

int nnd_conv_wgrad_tc32_supported(const ConvGeom& g, int Cdy, int Cx);

int nnd_conv_wgrad_tma32_supported(const ConvGeom& g, int Cdy, int Cx) {
  if (!nnd_conv_wgrad_tc32_supported(g, Cdy, Cx)) return 0;
  return (long long)g.N * ((g.Di + Q_ZT - 1) / Q_ZT) * ((g.Hi + Q_YT - 1) / Q_YT) * ((g.Wi + Q_RW - 1) / Q_RW) < (1ll << 31);
}

long long nnd_conv_wgrad_tma32_workspace(const ConvGeom& g, int Cout, int Cin) {
  if (Cout != 32 || Cin != 32) return 0;
  QArgs a; int grid;
  q_plan(g, a, grid);
  return (long long)grid * g.T * 32 * 32 * 4;
}

int nnd_conv_wgrad_tma32(const __nv_bfloat16* dy, const __nv_bfloat16* x, const ConvGeom& g, float* dw, long long s_co, long long s_ci,
                         long long s_tap, int Cout, int Cin, void* ws, long long ws_bytes, cudaStream_t st) {
  if (((size_t)dy & 15) || ((size_t)x & 15)) return NND_ERR_ARG;
  QArgs a; int grid;
  q_plan(g, a, grid);
  if (grid <= 0) return NND_OK;
  a.dw = dw; a.s_co = s_co; a.s_ci = s_ci; a.s_tap = s_tap; a.Cout = Cout; a.Cin = Cin; a.T = g.T; a.part = nullptr;
  for (int i = 0; i < 27; ++i) a.tw[i] = 255;
  unsigned seen = 0;
  for (int t = 0; t < g.T; ++t) {
    a.tw[(g.off_d[t] + 1) * 9 + (g.off_h[t] + 1) * 3 + (g.off_w[t] + 1)] = g.tap_w[t];
    if (g.tap_w[t] < 32) seen |= 1u << g.tap_w[t];
  }
  const long long block = (long long)g.T * Cout * Cin;
  if (ws && Cout == 32 && Cin == 32 && seen == (g.T >= 32 ? 0xffffffffu : (1u << g.T) - 1u) && ws_bytes >= (long long)grid * block * 4 &&
      !((size_t)ws & 15))
    a.part = reinterpret_cast<float*>(ws);
  QMaps maps;
  if (q_make_map(&maps.dy, dy, g.N, g.Di, g.Hi, g.Wi, Q_RW, Q_YT, Q_ZT + 2) != NND_OK ||
      q_make_map(&maps.x, x, g.N, g.Di, g.Hi, g.Wi, Q_XW, Q_YT + 2, Q_ZT) != NND_OK)
    return NND_ERR_ARG;                                                    // the caller falls back to the cp.async kernel
  constexpr size_t SMEM = (size_t)Q_STAGES * Q_STAGE + Q_TAIL + 1024;
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_tma32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  conv_wgrad_tma32_kernel<<<grid, Q_THREADS, SMEM, st>>>(maps, a);
  NND_LAUNCH_CHECK("conv_wgrad_tma32_kernel");
  if (a.part) {
    const long long n4 = block / 4;
    wgrad32_finish_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(a.part, grid, block, Cout, Cin, dw, s_co, s_ci, s_tap);
    NND_LAUNCH_CHECK("wgrad32_finish_kernel");
  }
  return NND_OK;
}
