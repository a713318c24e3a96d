// Weight gradient of the >= 64-channel 3x3x3 / 1x3x3 stride-1 convolutions: TMA-fed tcgen05 kernel.
//
//   dW[tap][co][ci] = sum_voxels dy[v][co] * x[v + off_tap][ci]          (nndet/arch/conv.py:344-348 via autograd)
//
// Same contraction as conv_wgrad_tc.cu (D[M = co][N = ci] += A[M][K] * B[N][K], K = 16 voxels, both operands "MN-major": a voxel is a
// row of channels in HBM), but the operands arrive by TMA instead of per-thread cp.async:
//   * ncu of the cp.async kernel (profiles/r02_ncu_wgrad_tc128_summary.txt): 1.22 GB of L2->SM reads for 67 MB of operands and a
//     tensor pipe at 28-30 % -- every 16-byte cp.async of a thread is its own LSU wavefront (544 per 16-voxel K-step against 192 cycles of
//     MMAs) and moves a 32-byte sector for 16 useful bytes.
//   * here ONE elected lane issues `cp.async.bulk.tensor.5d` boxes of [64 channels x BW voxels x BH rows] straight out of the NDHWC
//     tensors (5-D tensor maps, SWIZZLE_128B, out-of-bounds = zero fill = the convolution's padding AND the ragged edges): whole 128-byte
//     lines, no address arithmetic, no LSU traffic, 4 instructions per 64-voxel pipeline stage.
// Shared-memory image of a box = the canonical MN-major SWIZZLE_128B UMMA layout: one K row (voxel) = 128 bytes = 64 channels, 8 rows
// = one 1024-byte swizzle atom; LBO = distance between 64-channel blocks, SBO = distance between 8-voxel groups.  The three dx taps of
// a filter row read the SAME x box at start addresses shifted by one row (128 bytes); stacked along N with LBO = 128 bytes they are ONE
// M = 128 x N = 192 x K = 16 MMA per 64 input channels (96 cycles = the full tensor rate; the N = 64 / 128 MMAs of the cp.async kernel
// pay the shared-memory re-read of A: max(N/2, 32 + N/4) cycles, scripts/mma_rate.cu).
// Row-shifted starts are not 1024-byte aligned: the descriptor's base_offset field carries (start >> 7) & 7 (mode bit 1; the PTX rule
// for patterns that do not start at the alignment boundary) -- the A/B switch `nnd_conv_set_wgrad_tma` keeps the variants selectable.
//
// Work split as before: one CTA = one (dz, dy) filter row x one 128-channel co tile x NB 64-channel ci blocks x one contiguous range
// of 64-voxel units (split-K, fp32 atomics into dW).  Two unit shapes: "wide" 16 w x 4 h (K-step = 16 consecutive w) and "narrow"
// 8 w x 8 h (K-step = 8 w of two consecutive h rows: SBO = the box's row pitch) for W <= 8 and widths like 24 / 40.
// Warp roles: 0 TMA producer | 1 MMA issuer | 2-5 epilogue (TMEM -> atomics).
#include <cuda.h>

#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int WM_NI = 2;                      // MMA issuer warps (alternate pipeline stages)
constexpr int WM_THREADS = (1 + WM_NI + 4) * 32;
constexpr int WM_KSTEPS = 4;                   // 16-voxel K-steps per stage (64 voxels)
constexpr int WM_ABLK = 64 * 128;              // one 64-channel block of dy: 64 K rows x 128 B
constexpr int WM_SMEM_MAX = 232448 - 2048;     // dynamic shared memory minus alignment slack

__host__ __device__ constexpr int wm_bblk(int narrow) { return (narrow ? 80 : 72) * 128; }       // x box: 10 x 8 or 18 x 4 rows
__host__ __device__ constexpr int wm_stage_bytes(int nb, int narrow) { return 2 * WM_ABLK + nb * wm_bblk(narrow); }
__host__ __device__ constexpr int wm_stages(int nb, int narrow) {
  return WM_SMEM_MAX / wm_stage_bytes(nb, narrow) > 8 ? 8 : WM_SMEM_MAX / wm_stage_bytes(nb, narrow);
}

struct WmArgs {
  float* dw; long long s_co, s_ci, s_tap;
  float* part;                      // split-K partials [split][tap][Cout][Cin] fp32 (caller's workspace) or null: atomics straight into dW
  int T;                            // taps of the filter (first extent of a split's partial block)
  int Cout, Cin, Cdy;
  int D, H, W;                      // grid of dy == grid of x (stride 1)
  int HB, WS;                       // units along h / w
  long long total_units;            // N * D * HB * WS
  long long units_per_split;
  int n_groups;
  signed char gdz[9], gdy[9];
  unsigned char gtw[9][3];          // weight tap index of dx = -1, 0, +1
  unsigned char gtw2[9][3];         // pair mode: taps of filter row (dz, dy - 1) accumulated in MMA rows 64..127 (255 = row absent)
  int pair;                         // Cdy == 64: the second 64-row half of the MMA takes dy shifted by one h row = the filter row above
  int ci_tiles;
  int mode;                         // bit 1: base_offset = (start >> 7) & 7, bit 2: one N = 64 MMA per tap instead of the N = 192 stack
                                    // timing experiments (wrong results): bit 3 no epilogue atomics, bit 4 no MMAs, bit 5 no TMA loads
};

__device__ __forceinline__ void wm_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void wm_tma_5d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}

// MN-major SWIZZLE_128B descriptor words: lo = start >> 4 | (LBO >> 4) << 16; hi = SBO >> 4 | version 1 << 14 | base_offset << 17 |
// layout 2 << 29
__device__ __forceinline__ unsigned wm_lo(unsigned start, unsigned lbo) { return ((start >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ unsigned wm_hi(unsigned start, unsigned sbo, int use_bo) {
  return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (use_bo ? ((start >> 7) & 7u) << 17 : 0u) | (2u << 29);
}

// Position of a 64-voxel unit, advanced incrementally (w segment fastest, then h block, slice, sample): a 64-bit div / mod decode per unit
// cost ~850 cycles in BOTH the producer and the issuer -- more than the unit's MMAs (first device run: 1680 instead of 768 cycles per stage).
struct WmCursor {
  int ws, hb, d, n;
  __device__ __forceinline__ void init(long long u, const WmArgs& a) {
    unsigned v = (unsigned)u;                          // total_units < 2^31 (host check)
    ws = (int)(v % (unsigned)a.WS); v /= (unsigned)a.WS;
    hb = (int)(v % (unsigned)a.HB); v /= (unsigned)a.HB;
    d = (int)(v % (unsigned)a.D); n = (int)(v / (unsigned)a.D);
  }
  __device__ __forceinline__ void next(const WmArgs& a) {
    if (++ws == a.WS) { ws = 0; if (++hb == a.HB) { hb = 0; if (++d == a.D) { d = 0; ++n; } } }
  }
};

template <int NB, int NARROW>
__global__ void __launch_bounds__(WM_THREADS, 1)
conv_wgrad_tma_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WmArgs a) {
  constexpr int BBLK = wm_bblk(NARROW);
  constexpr int STAGE_BYTES = wm_stage_bytes(NB, NARROW);
  constexpr int STAGES = wm_stages(NB, NARROW);
  constexpr int RSTEP = NARROW ? 20 : 18;                 // x rows between two K-steps
  constexpr unsigned SBO_B = NARROW ? 10 * 128 : 1024;    // second 8-voxel group of a K-step: next h row (narrow) or next 8 w (wide)
  constexpr int TMEM_COLS = NB * 192 > 256 ? 512 : 256;
  // kind::f16, D fp32, A/B bf16, both MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
  constexpr unsigned IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 4) << 24);
  constexpr unsigned IDESC192 = IDESC_BASE | ((192u >> 3) << 17), IDESC64 = IDESC_BASE | ((64u >> 3) << 17);
  static_assert(STAGE_BYTES % 1024 == 0 && BBLK % 1024 == 0, "swizzled boxes need 1024-byte alignment");

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
  const int co0 = (blockIdx.z / a.ci_tiles) * 128, ci0 = (blockIdx.z % a.ci_tiles) * (64 * NB);
  const long long u0 = (long long)blockIdx.x * a.units_per_split;
  const long long u1 = min(u0 + a.units_per_split, a.total_units);
  const int n_units = u1 > u0 ? (int)(u1 - u0) : 0;
  const int dz = a.gdz[grp], ty = a.gdy[grp];
  const int a_blocks = (a.Cdy - co0) >= 128 ? 2 : 1;          // 64-channel blocks of dy this co tile really has
  // Pair mode (64 output channels): ONE dy box with an extra h row; MMA rows 64..127 read it one h row further (LBO = one row of the
  // box), i.e. D[64 + co][dx, ci] = sum dy[h + 1][co] * x[h + ty][ci] = the filter row (dz, ty - 1).  Units start at h = -1 so that
  // dy row 0 meets x row ty - 1 as well.  Six CTA groups instead of nine for a 3x3x3 filter, every MMA row but one group's half used.
  const int pair = a.pair;
  constexpr int BH_ = NARROW ? 8 : 4, BW_ = NARROW ? 8 : 16;
  const unsigned lbo_a = pair ? (unsigned)(BW_ * 128) : (unsigned)WM_ABLK;

  // a co tile with one real block: rows 64..127 of every stage stay zero for the whole kernel (zero MMA rows)
  if (a_blocks == 1 && !pair)
    for (int i = tid; i < STAGES * (WM_ABLK / 16); i += WM_THREADS)
      reinterpret_cast<uint4*>(smem + (size_t)(i / (WM_ABLK / 16)) * STAGE_BYTES + WM_ABLK)[i % (WM_ABLK / 16)] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(FULL(i), 1); mbar_init(EMPTY(i), 1); }
    mbar_init(DONE, WM_NI);
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
  if (warp >= 1 + WM_NI) {                         // accumulators start at zero: every MMA accumulates, an idle CTA adds zeros
    const int q = warp & 3;
    for (int c = 0; c < NB * 192; c += 32) tmem_zero32(tmem_base + ((unsigned)(q * 32) << 16) + c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      unsigned stage = 0, phase = 0;
      const unsigned tx = (unsigned)((pair ? (BH_ + 1) * BW_ * 128 : a_blocks * WM_ABLK) + NB * BBLK);
      WmCursor it;
      it.init(u0, a);
      for (int i = 0; i < n_units; ++i, it.next(a)) {
        if ((unsigned)(it.d + dz) >= (unsigned)a.D) continue;            // the whole x box is padding: nothing to add
        const int w0 = it.ws * BW_, h0 = it.hb * BH_ - pair;
        mbar_wait(EMPTY(stage), phase ^ 1);
        const unsigned sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        if (a.mode & 32) { mbar_arrive(FULL(stage)); if (++stage == STAGES) { stage = 0; phase ^= 1; } continue; }
        wm_expect_tx(FULL(stage), tx);
        wm_tma_5d(sa, &map_dy, co0, w0, h0, it.d, it.n, FULL(stage));
        if (a_blocks == 2) wm_tma_5d(sa + WM_ABLK, &map_dy, co0 + 64, w0, h0, it.d, it.n, FULL(stage));
#pragma unroll
        for (int b = 0; b < NB; ++b)
          wm_tma_5d(sa + 2 * WM_ABLK + b * BBLK, &map_x, ci0 + b * 64, w0 - 1, h0 + ty, it.d + dz, it.n, FULL(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp <= WM_NI) {
    // ================================================================ MMA issuers (the whole warp runs the control flow, the elected
    // lane issues: scripts/mma_rate.cu -- a divergent single-lane loop costs 81 instead of 48 issue cycles per MMA).  Two warps take
    // alternate stages: with one, ~370 cycles of per-stage bookkeeping (barrier poll, fence, descriptor set-up, commit) were exposed
    // between the 8 x ~105-cycle MMAs of a stage (second device run: 1260 cycles per stage without any TMA traffic).  Every MMA
    // accumulates onto zero-initialised TMEM, so the order between the two issue streams does not matter.
    const int me = warp - 1;
    unsigned stage = 0, phase = 0, turn = 0;
    const unsigned tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int use_bo = (a.mode >> 1) & 1, unstacked = (a.mode >> 2) & 1, slow = (a.mode & 6) != 0, no_mma = (a.mode & 16) != 0;
    // loop-invariant descriptor halves (base_offset 0): only the 14-bit start field moves
    const unsigned a_hi = wm_hi(0, 1024, 0), b_hi = wm_hi(0, SBO_B, 0);
    const unsigned smem0 = smem_u32(smem);
    WmCursor it;
    it.init(u0, a);
    for (int i = 0; i < n_units; ++i, it.next(a)) {
      if ((unsigned)(it.d + dz) >= (unsigned)a.D) continue;
      if ((int)turn == me) {
        mbar_wait_warp(FULL(stage), phase, lane);
        tc_fence_after();
        if (elect_one()) {
          const unsigned sa = smem0 + stage * STAGE_BYTES, sb = sa + 2 * WM_ABLK;
          if (!slow) {
            const unsigned a_lo0 = wm_lo(sa, lbo_a), b_lo0 = wm_lo(sb, 128);
            if (!no_mma) {
#pragma unroll
              for (int j = 0; j < WM_KSTEPS; ++j)
#pragma unroll
                for (int b = 0; b < NB; ++b)
                  tc_mma_acc2(tm + b * 192, a_lo0 + ((j * 2048) >> 4), a_hi, b_lo0 + ((b * BBLK + j * RSTEP * 128) >> 4), b_hi, IDESC192);
            }
          } else {
            // A/B variants of the descriptor model (tests/test_wgrad_tma_gpu.py): base_offset = (start >> 7) & 7 and / or one N = 64
            // MMA per dx tap instead of the N = 192 stack
#pragma unroll 1
            for (int j = 0; j < WM_KSTEPS; ++j) {
              const unsigned astart = sa + j * 2048;
#pragma unroll 1
              for (int b = 0; b < NB; ++b) {
                const unsigned bstart = sb + b * BBLK + j * RSTEP * 128;
                if (!unstacked) {
                  tc_mma2(tm + b * 192, wm_lo(astart, lbo_a), wm_hi(astart, 1024, use_bo), wm_lo(bstart, 128), wm_hi(bstart, SBO_B, use_bo), IDESC192, 1u);
                } else {
                  for (int t = 0; t < 3; ++t)
                    tc_mma2(tm + b * 192 + t * 64, wm_lo(astart, lbo_a), wm_hi(astart, 1024, use_bo), wm_lo(bstart + t * 128, 128),
                            wm_hi(bstart + t * 128, SBO_B, use_bo), IDESC64, 1u);
                }
              }
            }
          }
          tc_commit(EMPTY(stage));
        }
        __syncwarp();
      }
      turn = (turn + 1 == WM_NI) ? 0 : turn + 1;
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) tc_commit(DONE);
    __syncwarp();
  } else {
    // ================================================================ epilogue.  With a workspace: plain 16-byte stores of this CTA's
    // 128 x (NB x 192) accumulators into its own block of the split-K partials (every (split, tap, co, ci) is written by exactly one
    // CTA, idle ones store zeros), wgrad_finish_kernel sums the splits.  Without: fp32 atomics straight into dW (first device run: 42 of
    // the 155 us of a 128 -> 128 @32^3 launch -- 49 152 REDG lane-operations per CTA at ~1.3 cycles each, 7 M atomics per launch).
    const int q = warp & 3;
    const int upper = pair && q >= 2;               // pair mode: TMEM lanes 64..127 hold the filter row above for co = lane - 64
    const int co = co0 + (pair ? (q & 1) : q) * 32 + lane;
    mbar_wait_warp_backoff(DONE, 0, lane, 1000);   // the wait lasts the whole kernel
    tc_fence_after();
#pragma unroll 1
    for (int b = 0; b < NB; ++b) {
#pragma unroll 1
      for (int t = 0; t < 3; ++t) {
        const int tw = upper ? a.gtw2[grp][t] : a.gtw[grp][t];
        if (tw == 255) continue;                       // pair mode: the filter has no row above this group's
        float* dwt = a.dw + (long long)tw * a.s_tap + (long long)co * a.s_co;
        float* pt = a.part ? a.part + (((long long)blockIdx.x * a.T + tw) * a.Cout + co) * a.Cin : nullptr;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          unsigned v[32];
          tmem_ld32(tmem_base + ((unsigned)(q * 32) << 16) + b * 192 + t * 64 + c * 32, v);
          const int cib = ci0 + b * 64 + c * 32;
          if (co < a.Cout && cib < a.Cin && !(a.mode & 8)) {
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

// dW[tap][co][ci] += sum over the splits of the partials.  A thread owns four consecutive ci of one (tap, co); the additions into dW are
// atomic because launches for the SAME weight tensor (the head's convolutions are shared between pyramid levels that run on different
// streams) may finish concurrently: T * Cout * Cin atomics per launch instead of splits * that.
__global__ void __launch_bounds__(256)
wgrad_finish_kernel(const float* __restrict__ part, int splits, long long block, int Cout, int Cin, float* __restrict__ dw,
                    long long s_co, long long s_ci, long long s_tap) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // index of a float4 inside one split's block
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

typedef CUresult (*WmEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
WmEncodeTiledFn wm_encode_fn() {
  static WmEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WmEncodeTiledFn>(p);
  }
  return fn;
}

// [N, D, H, W, C] bf16 tensor, box = 64 channels x bw x bh voxels of one slice
int wm_make_map(CUtensorMap* map, const void* base, int N, int D, int H, int W, int C, int bw, int bh) {
  const WmEncodeTiledFn enc = wm_encode_fn();
  if (!enc) return NND_ERR_CUDA;
  const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t gstride[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  const cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NND_OK : NND_ERR_CUDA;
}

// split-K plan shared by the launcher and the workspace query: one wave of 1-CTA-per-SM blocks
long long wm_plan_splits(long long total_units, long long tiles, long long* units_per_split) {
  long long splits = NND_NUM_SMS / tiles;
  const long long max_splits = (total_units + 1) / 2;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  *units_per_split = (total_units + splits - 1) / splits;
  return (total_units + *units_per_split - 1) / *units_per_split;
}

template <int NB, int NARROW>
int launch_wm(const CUtensorMap& map_dy, const CUtensorMap& map_x, WmArgs a, long long splits, void* ws, long long ws_bytes, cudaStream_t st) {
  const long long block = (long long)a.T * a.Cout * a.Cin;          // floats per split
  a.part = (ws && ws_bytes >= splits * block * 4 && !((size_t)ws & 15)) ? reinterpret_cast<float*>(ws) : nullptr;
  constexpr size_t SMEM = (size_t)wm_stages(NB, NARROW) * wm_stage_bytes(NB, NARROW) + 1024;
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_tma_kernel<NB, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  dim3 grid((unsigned)splits, (unsigned)a.n_groups, (unsigned)(((a.Cdy + 127) / 128) * a.ci_tiles));
  conv_wgrad_tma_kernel<NB, NARROW><<<grid, WM_THREADS, SMEM, st>>>(map_dy, map_x, a);
  NND_LAUNCH_CHECK("conv_wgrad_tma_kernel");
  if (a.part) {
    const long long n4 = block / 4;
    wgrad_finish_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(a.part, (int)splits, block, a.Cout, a.Cin, a.dw, a.s_co, a.s_ci, a.s_tap);
    NND_LAUNCH_CHECK("wgrad_finish_kernel");
  }
  return NND_OK;
}

}  // namespace

// Stride-1 3x3x3 / 1x3x3 convolutions (every (dz, dy) filter row with its three dx taps), channel counts in multiples of 64.
int nnd_conv_wgrad_tma_supported(const ConvGeom& g, int Cdy, int Cx) {
  if (g.sd != 1 || g.sh != 1 || g.sw != 1 || g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return 0;
  if (g.Ld != g.Di || g.Lh != g.Hi || g.Lw != g.Wi || g.Do != g.Di || g.Ho != g.Hi || g.Wo != g.Wi) return 0;
  if ((g.T != 9 && g.T != 27) || Cdy % 64 || Cx % 64) return 0;
  if (g.Wi < 4 || g.Hi < 2) return 0;
  if ((long long)g.N * g.Di * g.Hi * g.Wi >= (1ll << 31)) return 0;        // unit counts fit 32 bits
  int rows[3][3] = {};
  for (int t = 0; t < g.T; ++t) {
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
    rows[g.off_d[t] + 1][g.off_h[t] + 1] |= 1 << (g.off_w[t] + 1);
  }
  for (int z = 0; z < 3; ++z)
    for (int y = 0; y < 3; ++y)
      if (rows[z][y] != 0 && rows[z][y] != 7) return 0;                   // a filter row is complete or absent
  return 1;
}

namespace {
// Launch plan shared by the launcher and the workspace query.
struct WmPlan {
  int narrow, bw, bh, pair, nb, co_tiles;
  WmArgs a;                       // geometry part filled in (units, groups, ci_tiles, units_per_split)
  long long splits;
};

void wm_plan(const ConvGeom& g, int Cdy, int Cx, WmPlan& p) {
  const int W = g.Lw, H = g.Lh, D = g.Ld;
  // narrow units (8 w x 8 h) when they pad the width less than 16-voxel row segments do
  p.narrow = ((W + 7) / 8) * 8 < ((W + 15) / 16) * 16 ? 1 : 0;
  p.bw = p.narrow ? 8 : 16; p.bh = p.narrow ? 8 : 4;
  p.pair = Cdy == 64 ? 1 : 0;
  WmArgs& a = p.a;
  a.pair = p.pair;
  a.D = D; a.H = H; a.W = W;
  a.HB = (H + p.pair + p.bh - 1) / p.bh;                 // pair mode: units cover h = -1 .. H - 1
  a.WS = (W + p.bw - 1) / p.bw;
  a.total_units = (long long)g.N * D * a.HB * a.WS;
  // CTA groups: one per (dz, dy) filter row; pair mode: one per two rows (dy, dy - 1), walking down from dy = +1
  unsigned char rows[3][3][3];
  int present[3][3] = {};
  for (int t = 0; t < g.T; ++t) {
    rows[g.off_d[t] + 1][g.off_h[t] + 1][g.off_w[t] + 1] = g.tap_w[t];
    present[g.off_d[t] + 1][g.off_h[t] + 1] = 1;
  }
  a.n_groups = 0;
  for (int z = 0; z < 3; ++z)
    for (int y = 2; y >= 0; --y) {
      if (!present[z][y]) continue;
      const int n = a.n_groups++;
      a.gdz[n] = (signed char)(z - 1); a.gdy[n] = (signed char)(y - 1);
      for (int k = 0; k < 3; ++k) { a.gtw[n][k] = rows[z][y][k]; a.gtw2[n][k] = 255; }
      if (p.pair && y > 0 && present[z][y - 1]) {
        for (int k = 0; k < 3; ++k) a.gtw2[n][k] = rows[z][y - 1][k];
        --y;                                              // the row below is served by this group's upper half
      }
    }
  p.co_tiles = (Cdy + 127) / 128;
  p.nb = Cx % 128 == 0 ? 2 : 1;
  a.ci_tiles = Cx / (64 * p.nb);
  const long long tiles = (long long)a.n_groups * p.co_tiles * a.ci_tiles;
  p.splits = a.total_units > 0 ? wm_plan_splits(a.total_units, tiles, &a.units_per_split) : 0;
}
}  // namespace

// Bytes of split-K partials nnd_conv_wgrad_tma wants as its workspace for this launch (0: none needed).  Host only.
long long nnd_conv_wgrad_tma_workspace(const ConvGeom& g, int Cdy, int Cx, int Cout, int Cin) {
  WmPlan p;
  wm_plan(g, Cdy, Cx, p);
  return p.splits * g.T * (long long)Cout * Cin * 4;
}

int nnd_conv_wgrad_tma(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                       long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, int mode, void* ws, long long ws_bytes,
                       cudaStream_t st) {
  if (((size_t)dy & 15) || ((size_t)x & 15)) return NND_ERR_ARG;
  WmPlan p;
  wm_plan(g, Cdy, Cx, p);
  WmArgs& a = p.a;
  if (a.total_units <= 0) return NND_OK;
  a.dw = dw; a.s_co = s_co; a.s_ci = s_ci; a.s_tap = s_tap; a.Cout = Cout; a.Cin = Cin; a.Cdy = Cdy;
  a.T = g.T; a.part = nullptr; a.mode = mode;
  {
    // the partials are indexed by weight tap: usable when the taps are exactly the slices 0 .. T-1 and whole 32-channel chunks are real
    unsigned seen = 0;
    for (int t = 0; t < g.T; ++t) if (g.tap_w[t] < 32) seen |= 1u << g.tap_w[t];
    if (seen != (g.T >= 32 ? 0xffffffffu : (1u << g.T) - 1u) || Cin % 32) { ws = nullptr; ws_bytes = 0; }
  }
  CUtensorMap map_dy, map_x;
  if (wm_make_map(&map_dy, dy, g.N, a.D, a.H, a.W, Cdy, p.bw, p.bh + p.pair) != NND_OK ||
      wm_make_map(&map_x, x, g.N, a.D, a.H, a.W, Cx, p.bw + 2, p.bh) != NND_OK)
    return NND_ERR_ARG;                                                    // the caller falls back to the cp.async kernel
  if (p.nb == 2) return p.narrow ? launch_wm<2, 1>(map_dy, map_x, a, p.splits, ws, ws_bytes, st) : launch_wm<2, 0>(map_dy, map_x, a, p.splits, ws, ws_bytes, st);
  return p.narrow ? launch_wm<1, 1>(map_dy, map_x, a, p.splits, ws, ws_bytes, st) : launch_wm<1, 0>(map_dy, map_x, a, p.splits, ws, ws_bytes, st);
}

// Host-only: the split-K / CTA-group plan of a launch as flat ints (tests/test_tma_plan_cpu.py replays it in numpy).
//   out = [narrow, bw, bh, pair, nb, HB, WS, n_groups, ci_tiles, splits, units_per_split, total_units,  per group: dz, ty, gtw[3], gtw2[3]]
extern "C" int nnd_conv_wgrad_tma_plan_debug(const int* geom, int Cdy, int Cx, int* out, int cap) {
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
  if (!nnd_conv_wgrad_tma_supported(g, Cdy, Cx)) return -1;
  WmPlan p;
  wm_plan(g, Cdy, Cx, p);
  const int need = 12 + p.a.n_groups * 8;
  if (cap < need) return -1;
  int* o = out;
  *o++ = p.narrow; *o++ = p.bw; *o++ = p.bh; *o++ = p.pair; *o++ = p.nb; *o++ = p.a.HB; *o++ = p.a.WS; *o++ = p.a.n_groups;
  *o++ = p.a.ci_tiles; *o++ = (int)p.splits; *o++ = (int)p.a.units_per_split; *o++ = (int)p.a.total_units;
  for (int n = 0; n < p.a.n_groups; ++n) {
    *o++ = p.a.gdz[n]; *o++ = p.a.gdy[n];
    for (int k = 0; k < 3; ++k) *o++ = p.a.gtw[n][k];
    for (int k = 0; k < 3; ++k) *o++ = p.a.gtw2[n][k];
  }
  return need;
}
