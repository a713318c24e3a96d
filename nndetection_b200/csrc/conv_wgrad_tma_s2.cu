// Weight gradient of the stride-2 convolutions and (operands swapped by the caller) of the kernel == stride up-convolutions: the TMA-fed
// tcgen05 kernel of conv_wgrad_tma.cu for a strided second operand.
//
//   dW[tap][co][ci] = sum_o dy[o][co] * x[s * o + off_tap][ci]           (nndet/arch/conv.py:344-348 via autograd; off in [-1, 1])
//
// dy is dense as in the stride-1 kernel (A operand: 64-channel MN-major SWIZZLE_128B boxes, pair mode for 64 output channels).  x is
// read through tensor maps with ELEMENT STRIDE 2 along w (and s_h along h): per 64-voxel unit and channel block two boxes, the "odd"
// plane (positions 2 (w0 + k) - 1, 17 / 9 entries: dx = -1 at row 0, dx = +1 at row 1) and the "even" plane (2 (w0 + k): dx = 0) -- the
// TMA unit de-interleaves, every tap is again 16 consecutive K rows of a swizzled MN-major plane (conv_wgrad_tc.cu, SW = 2, builds the
// same planes with one cp.async per voxel and channel group: 949 us for the 32 -> 64 layer @64^3 = 129 TFLOP/s, LSU-bound).
// Taps dx = -1 / +1 are ONE MMA (N = 2 channel blocks, LBO = one row), dx = 0 a second one.  Channel blocks are 64 wide (SWIZZLE_128B)
// or, for the 32-channel operands of the first encoder stage / last up-convolution, 32 wide (SWIZZLE_64B).
// Pair mode with a strided h axis: MMA rows 64..127 (dy one h row further) meet x rows s_h (h + 1) + (ty - s_h): the filter row ty - s_h,
// i.e. rows +1 and -1 of a stride-2 filter share a CTA.
// Split-K partials + finishing pass, two issuer warps, incremental unit cursor: as in conv_wgrad_tma.cu.
#include <cuda.h>

#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int WS_NI = 2;
constexpr int WS_THREADS = (1 + WS_NI + 4) * 32;
constexpr int WS_KSTEPS = 4;
constexpr int WS_ABLK = 64 * 128;
constexpr int WS_SMEM_MAX = 232448 - 2048;

__host__ __device__ constexpr int ws_al(int v) { return (v + 1023) / 1024 * 1024; }
__host__ __device__ constexpr int ws_pwo(int narrow) { return narrow ? 9 : 17; }
__host__ __device__ constexpr int ws_pwe(int narrow) { return narrow ? 8 : 16; }
__host__ __device__ constexpr int ws_bh(int narrow) { return narrow ? 8 : 4; }
__host__ __device__ constexpr int ws_odd_bytes(int narrow, int cb) { return ws_al(ws_pwo(narrow) * ws_bh(narrow) * cb * 2); }
__host__ __device__ constexpr int ws_even_bytes(int narrow, int cb) { return ws_al(ws_pwe(narrow) * ws_bh(narrow) * cb * 2); }
__host__ __device__ constexpr int ws_stage_bytes(int nb, int narrow, int cb) {
  return 2 * WS_ABLK + nb * (ws_odd_bytes(narrow, cb) + ws_even_bytes(narrow, cb));
}
__host__ __device__ constexpr int ws_stages(int nb, int narrow, int cb) {
  return WS_SMEM_MAX / ws_stage_bytes(nb, narrow, cb) > 8 ? 8 : WS_SMEM_MAX / ws_stage_bytes(nb, narrow, cb);
}

struct WsArgs {
  float* dw; long long s_co, s_ci, s_tap;
  float* part; int T;
  int Cout, Cin, Cdy;
  int D, H, W;                      // dy grid (logical outputs)
  int Di, sd, sh;                   // x depth extent and the d / h strides (w stride = 2)
  int HB, WS;
  long long total_units, units_per_split;
  int n_groups;
  signed char gdz[9], gdy[9];
  unsigned char gtw[9][3], gtw2[9][3];
  int pair, ci_tiles;
  int dxmask;                       // bit 0: dx = -1 present, bit 1: dx = 0, bit 2: dx = +1 (the same for every filter row)
  int mode;
};

struct WsMaps { CUtensorMap dy, xo, xe; };

__device__ __forceinline__ void ws_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ws_tma_5d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned ws_lo(unsigned start, unsigned lbo) { return ((start >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ unsigned ws_hi(unsigned sbo, unsigned layout) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29); }

struct WsCursor {
  int ws, hb, d, n;
  __device__ __forceinline__ void init(long long u, const WsArgs& a) {
    unsigned v = (unsigned)u;
    ws = (int)(v % (unsigned)a.WS); v /= (unsigned)a.WS;
    hb = (int)(v % (unsigned)a.HB); v /= (unsigned)a.HB;
    d = (int)(v % (unsigned)a.D); n = (int)(v / (unsigned)a.D);
  }
  __device__ __forceinline__ void next(const WsArgs& a) {
    if (++ws == a.WS) { ws = 0; if (++hb == a.HB) { hb = 0; if (++d == a.D) { d = 0; ++n; } } }
  }
};

template <int NB, int NARROW, int CB>
__global__ void __launch_bounds__(WS_THREADS, 1)
conv_wgrad_tma_s2_kernel(const __grid_constant__ WsMaps maps, const WsArgs a) {
  constexpr int ROWB = CB * 2;                              // bytes per x row (voxel) = swizzle span
  constexpr unsigned LAYOUT_B = CB == 64 ? 2u : 4u;         // SWIZZLE_128B : SWIZZLE_64B
  constexpr int PWO = ws_pwo(NARROW), PWE = ws_pwe(NARROW), BH_ = ws_bh(NARROW), BW_ = NARROW ? 8 : 16;
  constexpr int ODD_BYTES = ws_odd_bytes(NARROW, CB), EVEN_BYTES = ws_even_bytes(NARROW, CB);
  constexpr int BBLK = ODD_BYTES + EVEN_BYTES;
  constexpr int STAGE_BYTES = ws_stage_bytes(NB, NARROW, CB);
  constexpr int STAGES = ws_stages(NB, NARROW, CB);
  // rows of a plane between two K-steps; second 8-voxel group of a K-step: the next 8 w (wide) or the next h row (narrow)
  constexpr int OSTEP = NARROW ? 2 * PWO : PWO, ESTEP = NARROW ? 2 * PWE : PWE;
  constexpr unsigned SBO_O = NARROW ? PWO * ROWB : 8 * ROWB, SBO_E = 8 * ROWB;
  constexpr int COLS = 3 * CB;                              // accumulator columns of a channel block: [dx = -1 | dx = +1 | dx = 0]
  constexpr int TMEM_COLS = NB * COLS > 256 ? 512 : (NB * COLS > 128 ? 256 : 128);
  constexpr unsigned IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 4) << 24);
  constexpr unsigned IDESC2 = IDESC_BASE | ((unsigned)((2 * CB) >> 3) << 17), IDESC1 = IDESC_BASE | ((unsigned)(CB >> 3) << 17);
  static_assert(STAGE_BYTES % 1024 == 0, "swizzled boxes need 1024-byte alignment");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long bars[2 * STAGES + 1];
  __shared__ unsigned s_tmem_base;
  const unsigned bar0 = smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (STAGES + i); };
  const unsigned DONE = bar0 + 8u * (2 * STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = blockIdx.y;
  const int co0 = (blockIdx.z / a.ci_tiles) * 128, ci0 = (blockIdx.z % a.ci_tiles) * (CB * NB);
  const long long u0 = (long long)blockIdx.x * a.units_per_split;
  const long long u1 = min(u0 + a.units_per_split, a.total_units);
  const int n_units = u1 > u0 ? (int)(u1 - u0) : 0;
  const int dz = a.gdz[grp], ty = a.gdy[grp];
  const int a_blocks = (a.Cdy - co0) >= 128 ? 2 : 1;
  const int pair = a.pair;
  const unsigned lbo_a = pair ? (unsigned)(BW_ * 128) : (unsigned)WS_ABLK;

  if (a_blocks == 1 && !pair)
    for (int i = tid; i < STAGES * (WS_ABLK / 16); i += WS_THREADS)
      reinterpret_cast<uint4*>(smem + (size_t)(i / (WS_ABLK / 16)) * STAGE_BYTES + WS_ABLK)[i % (WS_ABLK / 16)] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(FULL(i), 1); mbar_init(EMPTY(i), 1); }
    mbar_init(DONE, WS_NI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = s_tmem_base;
  if (warp >= 1 + WS_NI) {
    const int q = warp & 3;
    for (int c = 0; c < NB * COLS; c += 32) tmem_zero32(tmem_base + ((unsigned)(q * 32) << 16) + c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      unsigned stage = 0, phase = 0;
      const unsigned tx = (unsigned)((pair ? (BH_ + 1) * BW_ * 128 : a_blocks * WS_ABLK) + NB * (PWO + PWE) * BH_ * ROWB);
      WsCursor it;
      it.init(u0, a);
      for (int i = 0; i < n_units; ++i, it.next(a)) {
        const int xd = it.d * a.sd + dz;
        if ((unsigned)xd >= (unsigned)a.Di) continue;                     // the whole x slice is padding
        const int w0 = it.ws * BW_, h0 = it.hb * BH_ - pair;
        mbar_wait(EMPTY(stage), phase ^ 1);
        const unsigned sa = smem_u32(smem + (size_t)stage * STAGE_BYTES), sb = sa + 2 * WS_ABLK;
        ws_expect_tx(FULL(stage), tx);
        ws_tma_5d(sa, &maps.dy, co0, w0, h0, it.d, it.n, FULL(stage));
        if (a_blocks == 2) ws_tma_5d(sa + WS_ABLK, &maps.dy, co0 + 64, w0, h0, it.d, it.n, FULL(stage));
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          ws_tma_5d(sb + b * BBLK, &maps.xo, ci0 + b * CB, 2 * w0 - 1, h0 * a.sh + ty, xd, it.n, FULL(stage));
          ws_tma_5d(sb + b * BBLK + ODD_BYTES, &maps.xe, ci0 + b * CB, 2 * w0, h0 * a.sh + ty, xd, it.n, FULL(stage));
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp <= WS_NI) {
    // ================================================================ MMA issuers (alternate stages; every MMA accumulates)
    const int me = warp - 1;
    unsigned stage = 0, phase = 0, turn = 0;
    const unsigned tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const unsigned a_hi = ws_hi(1024, 2u), bo_hi = ws_hi(SBO_O, LAYOUT_B), be_hi = ws_hi(SBO_E, LAYOUT_B);
    const unsigned smem0 = smem_u32(smem);
    const int both = (a.dxmask & 5) == 5, minus = (a.dxmask & 1) != 0, plus = (a.dxmask & 4) != 0, centre = (a.dxmask & 2) != 0;
    WsCursor it;
    it.init(u0, a);
    for (int i = 0; i < n_units; ++i, it.next(a)) {
      if ((unsigned)(it.d * a.sd + dz) >= (unsigned)a.Di) continue;
      if ((int)turn == me) {
        mbar_wait_warp(FULL(stage), phase, lane);
        tc_fence_after();
        if (elect_one()) {
          const unsigned sa = smem0 + stage * STAGE_BYTES, sb = sa + 2 * WS_ABLK;
          const unsigned a_lo0 = ws_lo(sa, lbo_a);
#pragma unroll
          for (int j = 0; j < WS_KSTEPS; ++j) {
            const unsigned a_lo = a_lo0 + ((j * 2048) >> 4);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const unsigned odd = sb + b * BBLK + j * OSTEP * ROWB, even = sb + b * BBLK + ODD_BYTES + j * ESTEP * ROWB;
              const unsigned d0 = tm + b * COLS;
              if (both) tc_mma_acc2(d0, a_lo, a_hi, ws_lo(odd, ROWB), bo_hi, IDESC2);                       // dx = -1 | dx = +1
              else if (plus) tc_mma_acc2(d0 + CB, a_lo, a_hi, ws_lo(odd + ROWB, ROWB), bo_hi, IDESC1);
              else if (minus) tc_mma_acc2(d0, a_lo, a_hi, ws_lo(odd, ROWB), bo_hi, IDESC1);
              if (centre) tc_mma_acc2(d0 + 2 * CB, a_lo, a_hi, ws_lo(even, ROWB), be_hi, IDESC1);           // dx = 0
            }
          }
          tc_commit(EMPTY(stage));
        }
        __syncwarp();
      }
      turn = (turn + 1 == WS_NI) ? 0 : turn + 1;
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) tc_commit(DONE);
    __syncwarp();
  } else {
    // ================================================================ epilogue: split-K partials (plain 16-byte stores) or atomics
    const int q = warp & 3;
    const int upper = pair && q >= 2;
    const int co = co0 + (pair ? (q & 1) : q) * 32 + lane;
    mbar_wait_warp_backoff(DONE, 0, lane, 1000);
    tc_fence_after();
#pragma unroll 1
    for (int b = 0; b < NB; ++b) {
#pragma unroll 1
      for (int t = 0; t < 3; ++t) {
        const int tw = upper ? a.gtw2[grp][t] : a.gtw[grp][t];
        if (tw == 255) continue;
        const int col0 = b * COLS + (t == 0 ? 0 : (t == 2 ? CB : 2 * CB));
        float* dwt = a.dw + (long long)tw * a.s_tap + (long long)co * a.s_co;
        float* pt = a.part ? a.part + (((long long)blockIdx.x * a.T + tw) * a.Cout + co) * a.Cin : nullptr;
#pragma unroll 1
        for (int c = 0; c < CB / 32; ++c) {
          unsigned v[32];
          tmem_ld32(tmem_base + ((unsigned)(q * 32) << 16) + col0 + c * 32, v);
          const int cib = ci0 + b * CB + c * 32;
          if (co < a.Cout && cib < a.Cin) {
            if (pt) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(pt + cib + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) atomicAdd(dwt + (long long)(cib + j) * a.s_ci, __uint_as_float(v[j]));
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// dW[tap][co][ci] += sum over the splits (see wgrad_finish_kernel in conv_wgrad_tma.cu)
__global__ void __launch_bounds__(256)
wgrad_s2_finish_kernel(const float* __restrict__ part, int splits, long long block, int Cout, int Cin, float* __restrict__ dw,
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

typedef CUresult (*WsEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
WsEncodeTiledFn ws_encode_fn() {
  static WsEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WsEncodeTiledFn>(p);
  }
  return fn;
}

// [N, D, H, W, C] bf16 tensor; box = cb channels x pw x ph entries taken every (sw, sh)-th voxel of one slice
int ws_make_map(CUtensorMap* map, const void* base, int N, int D, int H, int W, int C, int cb, int pw, int ph, int sw, int sh) {
  const WsEncodeTiledFn enc = ws_encode_fn();
  if (!enc) return NND_ERR_CUDA;
  const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t gstride[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  const cuuint32_t box[5] = {(cuuint32_t)cb, (cuuint32_t)((pw - 1) * sw + 1), (cuuint32_t)((ph - 1) * sh + 1), 1, 1};
  const cuuint32_t estr[5] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NND_OK : NND_ERR_CUDA;
}

struct WsPlan { int narrow, bw, bh, cb, nb; WsArgs a; long long splits; };

void ws_plan(const ConvGeom& g, int Cdy, int Cx, WsPlan& p) {
  const int W = g.Lw, H = g.Lh, D = g.Ld;
  p.narrow = ((W + 7) / 8) * 8 < ((W + 15) / 16) * 16 ? 1 : 0;
  p.bw = p.narrow ? 8 : 16; p.bh = p.narrow ? 8 : 4;
  p.cb = Cx % 64 == 0 ? 64 : 32;
  p.nb = (Cx / p.cb) % 2 == 0 ? 2 : 1;
  WsArgs& a = p.a;
  a.pair = Cdy == 64 ? 1 : 0;
  a.D = D; a.H = H; a.W = W; a.Di = g.Di; a.sd = g.sd; a.sh = g.sh;
  a.HB = (H + a.pair + p.bh - 1) / p.bh;
  a.WS = (W + p.bw - 1) / p.bw;
  a.total_units = (long long)g.N * D * a.HB * a.WS;
  unsigned char rows[3][3][3];
  int present[3][3] = {};
  a.dxmask = 0;
  for (int z = 0; z < 3; ++z) for (int y = 0; y < 3; ++y) for (int k = 0; k < 3; ++k) rows[z][y][k] = 255;
  for (int t = 0; t < g.T; ++t) {
    rows[g.off_d[t] + 1][g.off_h[t] + 1][g.off_w[t] + 1] = g.tap_w[t];
    present[g.off_d[t] + 1][g.off_h[t] + 1] = 1;
    a.dxmask |= 1 << (g.off_w[t] + 1);
  }
  a.n_groups = 0;
  for (int z = 0; z < 3; ++z) {
    int served[3] = {0, 0, 0};
    for (int y = 2; y >= 0; --y) {
      if (!present[z][y] || served[y]) continue;
      const int n = a.n_groups++;
      a.gdz[n] = (signed char)(z - 1); a.gdy[n] = (signed char)(y - 1);
      for (int k = 0; k < 3; ++k) { a.gtw[n][k] = rows[z][y][k]; a.gtw2[n][k] = 255; }
      const int yu = y - g.sh;                             // pair mode: the upper MMA half serves the filter row s_h below this one
      if (a.pair && yu >= 0 && present[z][yu] && !served[yu]) {
        for (int k = 0; k < 3; ++k) a.gtw2[n][k] = rows[z][yu][k];
        served[yu] = 1;
      }
    }
  }
  a.ci_tiles = Cx / (p.cb * p.nb);
  const long long tiles = (long long)a.n_groups * ((Cdy + 127) / 128) * a.ci_tiles;
  long long splits = NND_NUM_SMS / tiles;
  const long long max_splits = (a.total_units + 1) / 2;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.units_per_split = a.total_units > 0 ? (a.total_units + splits - 1) / splits : 1;
  p.splits = a.total_units > 0 ? (a.total_units + a.units_per_split - 1) / a.units_per_split : 0;
}

template <int NB, int NARROW, int CB>
int launch_ws(const WsMaps& maps, WsArgs a, long long splits, void* ws, long long ws_bytes, cudaStream_t st) {
  const long long block = (long long)a.T * a.Cout * a.Cin;
  a.part = (ws && ws_bytes >= splits * block * 4 && !((size_t)ws & 15)) ? reinterpret_cast<float*>(ws) : nullptr;
  constexpr size_t SMEM = (size_t)ws_stages(NB, NARROW, CB) * ws_stage_bytes(NB, NARROW, CB) + 1024;
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_tma_s2_kernel<NB, NARROW, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  dim3 grid((unsigned)splits, (unsigned)a.n_groups, (unsigned)(((a.Cdy + 127) / 128) * a.ci_tiles));
  conv_wgrad_tma_s2_kernel<NB, NARROW, CB><<<grid, WS_THREADS, SMEM, st>>>(maps, a);
  NND_LAUNCH_CHECK("conv_wgrad_tma_s2_kernel");
  if (a.part) {
    const long long n4 = block / 4;
    wgrad_s2_finish_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(a.part, (int)splits, block, a.Cout, a.Cin, a.dw, a.s_co, a.s_ci, a.s_tap);
    NND_LAUNCH_CHECK("wgrad_s2_finish_kernel");
  }
  return NND_OK;
}

}  // namespace

// Stride 2 along w (1 or 2 along d / h), taps within [-1, 1], the same dx taps in every filter row, dy channels in multiples of 64,
// x channels in multiples of 32.
int nnd_conv_wgrad_tma_s2_supported(const ConvGeom& g, int Cdy, int Cx) {
  if (g.sw != 2 || g.sd < 1 || g.sd > 2 || g.sh < 1 || g.sh > 2) return 0;
  if (g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return 0;
  if (g.Do != g.Ld || g.Ho != g.Lh || g.Wo != g.Lw) return 0;
  if (g.T < 4 || Cdy % 64 || Cx % 32) return 0;
  if ((long long)g.N * g.Di * g.Hi * g.Wi >= (1ll << 31)) return 0;
  int rows[3][3] = {}, mask = 0;
  for (int t = 0; t < g.T; ++t) {
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
    rows[g.off_d[t] + 1][g.off_h[t] + 1] |= 1 << (g.off_w[t] + 1);
    mask |= 1 << (g.off_w[t] + 1);
  }
  for (int z = 0; z < 3; ++z)
    for (int y = 0; y < 3; ++y)
      if (rows[z][y] != 0 && rows[z][y] != mask) return 0;
  return 1;
}

long long nnd_conv_wgrad_tma_s2_workspace(const ConvGeom& g, int Cdy, int Cx, int Cout, int Cin) {
  WsPlan p;
  ws_plan(g, Cdy, Cx, p);
  return p.splits * g.T * (long long)Cout * Cin * 4;
}

int nnd_conv_wgrad_tma_s2(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                          long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, int mode, void* ws, long long ws_bytes,
                          cudaStream_t st) {
  if (((size_t)dy & 15) || ((size_t)x & 15)) return NND_ERR_ARG;
  WsPlan p;
  ws_plan(g, Cdy, Cx, p);
  WsArgs& a = p.a;
  if (a.total_units <= 0) return NND_OK;
  a.dw = dw; a.s_co = s_co; a.s_ci = s_ci; a.s_tap = s_tap; a.Cout = Cout; a.Cin = Cin; a.Cdy = Cdy;
  a.T = g.T; a.part = nullptr; a.mode = mode;
  {
    unsigned seen = 0;
    for (int t = 0; t < g.T; ++t) if (g.tap_w[t] < 32) seen |= 1u << g.tap_w[t];
    if (seen != (g.T >= 32 ? 0xffffffffu : (1u << g.T) - 1u) || Cin % 32) { ws = nullptr; ws_bytes = 0; }
  }
  WsMaps maps;
  if (ws_make_map(&maps.dy, dy, g.N, a.D, a.H, a.W, Cdy, 64, p.bw, p.bh + a.pair, 1, 1) != NND_OK ||
      ws_make_map(&maps.xo, x, g.N, g.Di, g.Hi, g.Wi, Cx, p.cb, p.bw + 1, p.bh, 2, g.sh) != NND_OK ||
      ws_make_map(&maps.xe, x, g.N, g.Di, g.Hi, g.Wi, Cx, p.cb, p.bw, p.bh, 2, g.sh) != NND_OK)
    return NND_ERR_ARG;
  if (p.cb == 64) {
    if (p.nb == 2) return p.narrow ? launch_ws<2, 1, 64>(maps, a, p.splits, ws, ws_bytes, st) : launch_ws<2, 0, 64>(maps, a, p.splits, ws, ws_bytes, st);
    return p.narrow ? launch_ws<1, 1, 64>(maps, a, p.splits, ws, ws_bytes, st) : launch_ws<1, 0, 64>(maps, a, p.splits, ws, ws_bytes, st);
  }
  if (p.nb == 2) return p.narrow ? launch_ws<2, 1, 32>(maps, a, p.splits, ws, ws_bytes, st) : launch_ws<2, 0, 32>(maps, a, p.splits, ws, ws_bytes, st);
  return p.narrow ? launch_ws<1, 1, 32>(maps, a, p.splits, ws, ws_bytes, st) : launch_ws<1, 0, 32>(maps, a, p.splits, ws, ws_bytes, st);
}

// Host-only: the plan of a strided launch as flat ints (tests/test_tma_plan_cpu.py replays it in numpy).
//   out = [narrow, bw, bh, pair, nb, cb, HB, WS, n_groups, ci_tiles, splits, units_per_split, total_units, dxmask,  per group: dz, ty, gtw[3], gtw2[3]]
extern "C" int nnd_conv_wgrad_tma_s2_plan_debug(const int* geom, int Cdy, int Cx, int* out, int cap) {
  if (!geom || !out) return -1;
  ConvGeom g;
  g.N = geom[0]; g.Di = geom[1]; g.Hi = geom[2]; g.Wi = geom[3]; g.Cin = geom[4];
  g.Ld = geom[5]; g.Lh = geom[6]; g.Lw = geom[7]; g.sd = geom[8]; g.sh = geom[9]; g.sw = geom[10];
  g.Do = geom[11]; g.Ho = geom[12]; g.Wo = geom[13];
  g.omd = geom[14]; g.omh = geom[15]; g.omw = geom[16]; g.ood = geom[17]; g.ooh = geom[18]; g.oow = geom[19];
  g.T = geom[20];
  if (g.T < 1 || g.T > NND_MAX_TAPS) return -1;
  for (int t = 0; t < g.T; ++t) {
    g.off_d[t] = (signed char)geom[21 + 4 * t]; g.off_h[t] = (signed char)geom[22 + 4 * t]; g.off_w[t] = (signed char)geom[23 + 4 * t];
    g.tap_w[t] = (unsigned char)geom[24 + 4 * t];
  }
  if (!nnd_conv_wgrad_tma_s2_supported(g, Cdy, Cx)) return -1;
  WsPlan p;
  ws_plan(g, Cdy, Cx, p);
  const int need = 14 + p.a.n_groups * 8;
  if (cap < need) return -1;
  int* o = out;
  *o++ = p.narrow; *o++ = p.bw; *o++ = p.bh; *o++ = p.a.pair; *o++ = p.nb; *o++ = p.cb; *o++ = p.a.HB; *o++ = p.a.WS; *o++ = p.a.n_groups;
  *o++ = p.a.ci_tiles; *o++ = (int)p.splits; *o++ = (int)p.a.units_per_split; *o++ = (int)p.a.total_units; *o++ = p.a.dxmask;
  for (int n = 0; n < p.a.n_groups; ++n) {
    *o++ = p.a.gdz[n]; *o++ = p.a.gdy[n];
    for (int k = 0; k < 3; ++k) *o++ = p.a.gtw[n][k];
    for (int k = 0; k < 3; ++k) *o++ = p.a.gtw2[n][k];
  }
  return need;
}
