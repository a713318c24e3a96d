// TMA-fed variant of the tcgen05 tile kernel (conv_tc.cu): the same im2col-free gather convolution -- 3x3x3 / 1x3x3 stride-1 layers, the
// >= 4-tap parity classes of a stride-2 dgrad, and (S2) stride-2 convolutions / the dgrad of up-convolutions -- with BOTH operand streams
// brought in by the tensor-memory accelerator.  Replaces cuDNN implicit GEMM behind torch.nn.Conv3d (nndet/arch/conv.py:344-348).
//
// Why: the cp.async producers of conv_tc.cu issue one 16-byte copy per thread and every one of them is its own LSU wavefront (weights:
// rows 256 bytes apart; halo: voxels Cin*2 bytes apart) -- ~73 k wavefronts per 128-channel tile against ~55 k cycles of MMAs, ~18 k against
// 5 k for the stride-2 form: the kernels are LSU-bound (128 -> 128 @32^3: 730 TFLOP/s, stride-2 32 -> 64: 180 TFLOP/s).  Here ONE lane
// posts, per CK-channel chunk of a tile, one `cp.async.bulk.tensor.5d` per halo plane and TG 2-D boxes per weight item.
//
// Shared-memory images (the row-shifted-start property they rest on was verified on the device for conv_wgrad_tma.cu: the swizzle is a
// function of the absolute shared-memory address, base_offset stays 0):
//   * halo plane = box [CK channels x PW x PH x PD voxels] of the NDHWC input, K-major rows of ROWB = CK*2 bytes (64: SWIZZLE_64B,
//     stride 1; 32: SWIZZLE_32B, stride 2), out-of-bounds = zero fill = padding.  The A operand of tap (dz, dy, dx) for depth slice mt is
//     the plane at start row ((mt + dz') * PH + dy') * PW + dx': 16 groups of 8 consecutive w rows, SBO = PW * ROWB.
//   * stride 1: one plane (MT+2) x 18 x 10.  Stride 2: input position 2 o + off per strided axis -> per axis an ODD plane (2 (o0 + k) - 1,
//     n + 1 entries; off = -1 reads it at shift 0, off = +1 at shift 1) and an EVEN plane (2 (o0 + k), n entries; off = 0): up to 8 boxes
//     with elementStrides 2 -- the TMA unit de-interleaves, the kernel sees dense planes again.
//   * weight item = TG boxes [CK x N_TILE rows] of the K-major pack [tap][CoutPad][Cin], canonical K-major swizzled rows (SBO = 8 rows).
// Roles: warp 0 producer | MT issuer warps (one per depth slice, whole warp runs the control flow, elected lane issues) | 4 epilogue
// warps (the epilogue of conv_tc.cu: bias / residual / scale, bf16 rows or strided fp32 head outputs, norm statistics).
#include <cuda.h>

#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int TBH = 16, TBW = 8;                 // tile = MT x 16 x 8 voxels (one 128-row MMA per depth slice)
constexpr int T_A_STAGES = 2;
constexpr int T_MAX_PLANES = 8;
constexpr int tct_threads(int mt) { return (1 + mt + 4) * 32; }

struct TctPlane {
  int pd, ph, pw;               // entries per axis
  int cd, ch, cw;               // coordinate of entry 0 relative to (s * o0) per axis: -1 (halo / odd plane) or 0 (even plane)
  int base;                     // byte offset inside a halo stage (1024-aligned)
};

struct TctArgs {
  int DB, HB, WB, NT, total;
  int n_planes;
  TctPlane pl[T_MAX_PLANES];
  int a_bytes;                  // bytes one halo stage receives (sum of the planes)
  int tap_off[NND_MAX_TAPS];    // byte offset of the tap's first row inside a halo stage (plane base + shift rows)
  int tap_sbo[NND_MAX_TAPS];    // PW * ROWB of the tap's plane
  int slice_step[NND_MAX_TAPS]; // bytes between depth slices of the tap's plane (PH * PW * ROWB)
};

__device__ __forceinline__ void tct_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tct_tma_5d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ void tct_tma_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

struct TctMaps { CUtensorMap pl[T_MAX_PLANES]; CUtensorMap w; };

template <int N_TILE, int MT, int TG, int ACC, int B_SLOTS, int S2>
__global__ void __launch_bounds__(tct_threads(MT), 1)
conv_tct_kernel(const __grid_constant__ TctMaps maps, const ConvGeom g, const ConvEpilogue ep, const TctArgs tl, const int a_stage_bytes) {
  constexpr int CK = S2 ? 16 : 32;                           // channels per chunk
  constexpr int ROWB = CK * 2;                               // bytes per operand row = swizzle span
  constexpr int KSTEPS = CK / 16;
  constexpr unsigned LAYOUT = S2 ? 6u : 4u;                  // SWIZZLE_32B : SWIZZLE_64B
  constexpr int B_TAP_BYTES = N_TILE * ROWB;
  constexpr int B_BYTES = TG * B_TAP_BYTES;
  constexpr int ACC_COLS = MT * N_TILE;
  constexpr int TMEM_COLS = ACC * ACC_COLS >= 512 ? 512 : (ACC * ACC_COLS >= 256 ? 256 : (ACC * ACC_COLS >= 128 ? 128 : 64));
  static_assert(ACC * ACC_COLS <= 512, "TMEM overflow");
  static_assert(B_TAP_BYTES % 1024 == 0, "weight slices keep the swizzle phase");
  constexpr unsigned IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N_TILE >> 3) << 17) | ((128u >> 4) << 24);
  constexpr unsigned B_HI = (unsigned)((8 * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                  // T_A_STAGES x a_stage_bytes
  unsigned char* sB = smem + T_A_STAGES * a_stage_bytes;     // B_SLOTS x B_BYTES
  __shared__ unsigned long long bars[2 * B_SLOTS + 2 * T_A_STAGES + 4];
  __shared__ unsigned s_tmem_base;
  __shared__ float s_stat[2][4][N_TILE];
  const unsigned bar0 = smem_u32(bars);
  auto FULLB = [&](int i) { return bar0 + 8u * i; };
  auto EMPTYB = [&](int i) { return bar0 + 8u * (B_SLOTS + i); };
  auto FULLA = [&](int i) { return bar0 + 8u * (2 * B_SLOTS + i); };
  auto EMPTYA = [&](int i) { return bar0 + 8u * (2 * B_SLOTS + T_A_STAGES + i); };
  auto TFULL = [&](int i) { return bar0 + 8u * (2 * B_SLOTS + 2 * T_A_STAGES + i); };
  auto TEMPTY = [&](int i) { return bar0 + 8u * (2 * B_SLOTS + 2 * T_A_STAGES + 2 + i); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = g.Cin / CK, T = g.T;
  const int NI_ITEMS = T / TG;

  if (tid == 0) {
    for (int i = 0; i < B_SLOTS; ++i) { mbar_init(FULLB(i), 1); mbar_init(EMPTYB(i), MT); }
    for (int i = 0; i < T_A_STAGES; ++i) { mbar_init(FULLA(i), 1); mbar_init(EMPTYA(i), MT); }
    for (int i = 0; i < 2; ++i) { mbar_init(TFULL(i), MT); mbar_init(TEMPTY(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = s_tmem_base;

  auto decode_tile = [&](int tile, int& n, int& d0, int& h0, int& w0, int& nt) {
    nt = tile % tl.NT; int r = tile / tl.NT;
    const int wb = r % tl.WB; r /= tl.WB;
    const int hb = r % tl.HB; r /= tl.HB;
    const int db = r % tl.DB; n = r / tl.DB;
    d0 = db * MT; h0 = hb * TBH; w0 = wb * TBW;
  };

  if (warp == 0) {
    // ================================================================ producer: one lane, a flat sequence of chunks = (tile, CK-channel
    // block); a chunk = its halo planes + NI_ITEMS weight items.  The halo of the NEXT chunk is posted right behind the first weight
    // item of the current one (its stage was released by the chunk before): a whole chunk of MMAs hides the load -- for the stride-2
    // form (90 KB of planes per 16-channel chunk against ~2.6 k cycles of MMAs) half a chunk did not.
    if (lane == 0) {
      unsigned slot = 0, slot_phase = 0, astage = 0, a_phase = 0;
      const int my_tiles = (tl.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int n_chunks = my_tiles * KC;
      auto load_halo = [&](int chunk) {
        const int tile = blockIdx.x + (chunk / KC) * gridDim.x;
        const int kc = chunk % KC;
        int n, d0, h0, w0, nt;
        decode_tile(tile, n, d0, h0, w0, nt);
        mbar_wait(EMPTYA(astage), a_phase ^ 1);
        const unsigned a_base = smem_u32(sA + (size_t)astage * a_stage_bytes);
        tct_expect_tx(FULLA(astage), (unsigned)tl.a_bytes);
        for (int p = 0; p < tl.n_planes; ++p) {
          const TctPlane& pl = tl.pl[p];
          tct_tma_5d(a_base + pl.base, &maps.pl[p], kc * CK, w0 * g.sw + pl.cw, h0 * g.sh + pl.ch, d0 * g.sd + pl.cd, n, FULLA(astage));
        }
        astage ^= 1; if (astage == 0) a_phase ^= 1;
      };
      for (int chunk = 0; chunk < n_chunks; ++chunk) {
        const int tile = blockIdx.x + (chunk / KC) * gridDim.x;
        const int kc = chunk % KC;
        const int nt = tile % tl.NT;
        if (chunk == 0) load_halo(0);
        for (int it = 0; it < NI_ITEMS; ++it) {
          if (it == (NI_ITEMS > 1 ? 1 : 0) && chunk + 1 < n_chunks) load_halo(chunk + 1);
          mbar_wait(EMPTYB(slot), slot_phase ^ 1);
          const unsigned b_base = smem_u32(sB + (size_t)slot * B_BYTES);
          tct_expect_tx(FULLB(slot), B_BYTES);
#pragma unroll
          for (int tt = 0; tt < TG; ++tt)
            tct_tma_2d(b_base + tt * B_TAP_BYTES, &maps.w, kc * CK, (int)g.tap_w[it * TG + tt] * ep.CoutPad + nt * N_TILE, FULLB(slot));
          if (++slot == B_SLOTS) { slot = 0; slot_phase ^= 1; }
        }
      }
    }
  } else if (warp <= MT) {
    // ================================================================ MMA issuers: one warp per depth slice (disjoint TMEM columns ->
    // independent MMA streams sharing the weight items)
    const int mt = warp - 1;
    const unsigned tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    unsigned slot = 0, slot_phase = 0, astage = 0, afull_phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < tl.total; tile += gridDim.x) {
      mbar_wait_warp(TEMPTY(acc), acc_phase ^ 1, lane);
      tc_fence_after();
      const unsigned d_tmem = tm + acc * ACC_COLS + mt * N_TILE;
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait_warp(FULLA(astage), afull_phase, lane);
        tc_fence_after();
        const unsigned a_stage = smem_u32(sA + (size_t)astage * a_stage_bytes);
        for (int it = 0; it < NI_ITEMS; ++it) {
          mbar_wait_warp(FULLB(slot), slot_phase, lane);
          tc_fence_after();
          if (elect_one()) {
            const unsigned b_item = smem_u32(sB + (size_t)slot * B_BYTES);
#pragma unroll
            for (int tt = 0; tt < TG; ++tt) {
              const int t = it * TG + tt;
              const unsigned a_start = a_stage + tl.tap_off[t] + mt * tl.slice_step[t];
              const unsigned a_hi = (unsigned)((tl.tap_sbo[t] >> 4) & 0x3FFF) | (1u << 14) | (LAYOUT << 29);
              const unsigned b_start = b_item + tt * B_TAP_BYTES;
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                tc_mma2(d_tmem, ((a_start + k * 32) >> 4) & 0x3FFF, a_hi, ((b_start + k * 32) >> 4) & 0x3FFF, B_HI, IDESC,
                        (kc | it | tt | k) != 0 ? 1u : 0u);
            }
            tc_commit(EMPTYB(slot));
          }
          __syncwarp();
          if (++slot == B_SLOTS) { slot = 0; slot_phase ^= 1; }
        }
        if (elect_one()) tc_commit(EMPTYA(astage));
        __syncwarp();
        astage ^= 1; if (astage == 0) afull_phase ^= 1;
      }
      if (elect_one()) tc_commit(TFULL(acc));
      __syncwarp();
      if (ACC == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else acc_phase ^= 1;
    }
  } else {
    // ================================================================ epilogue (4 warps -> TMEM lane quarter warp % 4); as in conv_tc.cu
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // MMA row = TMEM lane = voxel (hy, wx) of a slice
    const int hy = row >> 3, wx = row & 7;
    const float scale = ep.scale ? *ep.scale : 1.f;
    const bool do_stats = ep.stat_sum != nullptr;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(ep.out);
    unsigned acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < tl.total; tile += gridDim.x) {
      int n, d0, h0, w0, nt;
      decode_tile(tile, n, d0, h0, w0, nt);
      const int h = h0 + hy, w = w0 + wx;                    // logical output coordinates
      const bool hw_ok = h < g.Lh && w < g.Lw;
      mbar_wait_warp(TFULL(acc), acc_phase, lane);
      tc_fence_after();
      if (do_stats) {
        for (int ch = lane; ch < N_TILE; ch += 32) { s_stat[0][q][ch] = 0.f; s_stat[1][q][ch] = 0.f; }
      }
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int d = d0 + mt;
        const bool ok = hw_ok && d < g.Ld;
        const long long vox = ((long long)(n * g.Do + d * g.omd + g.ood) * g.Ho + (h * g.omh + g.ooh)) * g.Wo + (w * g.omw + g.oow);
#pragma unroll 1
        for (int c = 0; c < N_TILE / 32; ++c) {
          unsigned v[32];
          tmem_ld32(tmem_base + ((unsigned)(q * 32) << 16) + acc * ACC_COLS + mt * N_TILE + c * 32, v);
          const int co0 = nt * N_TILE + c * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (co0 + j < ep.Cout) f[j] += ep.bias[co0 + j];
          }
          if (ep.residual && ok) {
            const uint4* rp = reinterpret_cast<const uint4*>(ep.residual + vox * ep.Cout + co0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 rv = rp[u];
              const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 t2 = __bfloat1622float2(hp[k]);
                f[u * 8 + 2 * k] += t2.x; f[u * 8 + 2 * k + 1] += t2.y;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= scale;
          __align__(16) __nv_bfloat162 pk[16];
          if (!ep.out_fp32) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              pk[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              const float2 r2 = __bfloat1622float2(pk[j]);
              f[2 * j] = ok ? r2.x : 0.f; f[2 * j + 1] = ok ? r2.y : 0.f;
            }
          }
          if (ok) {
            if (ep.out_fp32) {
              const long long pv = vox - (long long)n * g.Do * g.Ho * g.Wo;
              float* of = reinterpret_cast<float*>(ep.out) + (long long)n * ep.out_n_stride + pv * ep.out_v_stride + co0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (co0 + j < ep.Cout) of[j] = f[j];
            } else {
              uint4* op = reinterpret_cast<uint4*>(outp + vox * ep.Cout + co0);
              const uint4* sp = reinterpret_cast<const uint4*>(pk);
#pragma unroll
              for (int u = 0; u < 4; ++u) op[u] = sp[u];
            }
          }
          if (do_stats) {
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = f[j] * f[j];
            int col;
            const float cs = warp_transpose_reduce32(f, lane, col);
            const float cq = warp_transpose_reduce32(sq, lane, col);
            s_stat[0][q][c * 32 + col] += cs;          // col is a permutation of the lanes: no two lanes share a slot
            s_stat[1][q][c * 32 + col] += cq;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY(acc));
      if (ACC == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else acc_phase ^= 1;
      if (do_stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - (1 + MT) * 32;        // 0..127
        for (int ch = et; ch < N_TILE; ch += 128) {
          const float s = s_stat[0][0][ch] + s_stat[0][1][ch] + s_stat[0][2][ch] + s_stat[0][3][ch];
          const float qq = s_stat[1][0][ch] + s_stat[1][1][ch] + s_stat[1][2][ch] + s_stat[1][3][ch];
          atomicAdd(&ep.stat_sum[(size_t)n * ep.Cout + nt * N_TILE + ch], s);
          atomicAdd(&ep.stat_sq[(size_t)n * ep.Cout + nt * N_TILE + ch], qq);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
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

typedef CUresult (*TctEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TctEncodeTiledFn tct_encode_fn() {
  static TctEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TctEncodeTiledFn>(p);
  }
  return fn;
}

// S2: planes of a strided axis with n outputs per tile: odd (n + 1 entries from 2 o0 - 1) then even (n entries from 2 o0)
struct AxisPlanes { int count; int entries[2]; int c0[2]; };
AxisPlanes axis_planes(int stride, int n) {
  AxisPlanes a;
  if (stride == 2) { a.count = 2; a.entries[0] = n + 1; a.c0[0] = -1; a.entries[1] = n; a.c0[1] = 0; }
  else { a.count = 1; a.entries[0] = n + 2; a.c0[0] = -1; a.entries[1] = 0; a.c0[1] = 0; }
  return a;
}

// Host plan of a launch: the halo planes of one channel chunk (entries, first coordinate, byte offset inside a stage, tensor-map box
// extents) and, per tap, where its 128 operand rows start (byte offset inside a stage), the pitch of its 8-row groups and of its depth
// slices.  Pure arithmetic: shared by the launcher and by nnd_conv_tct_plan_debug, which tests/test_tma_plan_cpu.py replays in numpy.
struct TctHostPlan {
  TctArgs tl;
  int a_stage_bytes;
  int box[T_MAX_PLANES][3];          // tensor-map box extents (w, h, d) in tensor elements: (entries - 1) * stride + 1
  int max_tw;
};

void tct_host_plan(const ConvGeom& g, int MT, int ROWB, TctHostPlan& hp) {
  TctArgs& tl = hp.tl;
  const AxisPlanes ad = axis_planes(g.sd, MT), ah = axis_planes(g.sh, TBH), aw = axis_planes(g.sw, TBW);
  int plane_of[2][2][2];
  tl.n_planes = 0; tl.a_bytes = 0;
  int off = 0;
  for (int i = 0; i < ad.count; ++i)
    for (int j = 0; j < ah.count; ++j)
      for (int k = 0; k < aw.count; ++k) {
        const int p = tl.n_planes++;
        plane_of[i][j][k] = p;
        TctPlane& pl = tl.pl[p];
        pl.pd = ad.entries[i]; pl.ph = ah.entries[j]; pl.pw = aw.entries[k];
        pl.cd = ad.c0[i]; pl.ch = ah.c0[j]; pl.cw = aw.c0[k];
        pl.base = off;
        const int bytes = pl.pd * pl.ph * pl.pw * ROWB;
        tl.a_bytes += bytes;
        off += (bytes + 1023) / 1024 * 1024;
        // with an element stride s the box spans (entries - 1) * s + 1 tensor elements of which every s-th is copied
        hp.box[p][0] = (pl.pw - 1) * g.sw + 1; hp.box[p][1] = (pl.ph - 1) * g.sh + 1; hp.box[p][2] = (pl.pd - 1) * g.sd + 1;
      }
  hp.a_stage_bytes = off;
  hp.max_tw = 0;
  for (int t = 0; t < g.T; ++t) {
    // axis offset -> (plane, shift): unstrided axis: the halo plane at off + 1; strided: off = 0 even plane, off = -1 / +1 odd plane at 0 / 1
    auto pick = [](int stride, int o, int& plane, int& shift) {
      if (stride == 2) { plane = o == 0 ? 1 : 0; shift = o > 0 ? 1 : 0; } else { plane = 0; shift = o + 1; }
    };
    int pi, pj, pk, sz, sy, sx;
    pick(g.sd, g.off_d[t], pi, sz); pick(g.sh, g.off_h[t], pj, sy); pick(g.sw, g.off_w[t], pk, sx);
    const TctPlane& pl = tl.pl[plane_of[pi][pj][pk]];
    tl.tap_off[t] = pl.base + ((sz * pl.ph + sy) * pl.pw + sx) * ROWB;
    tl.tap_sbo[t] = pl.pw * ROWB;
    tl.slice_step[t] = pl.ph * pl.pw * ROWB;
    if (g.tap_w[t] > hp.max_tw) hp.max_tw = g.tap_w[t];
  }
  for (int t = g.T; t < NND_MAX_TAPS; ++t) { tl.tap_off[t] = 0; tl.tap_sbo[t] = 0; tl.slice_step[t] = 0; }
}

template <int N_TILE, int MT, int TG, int ACC, int B_SLOTS, int S2>
int launch_tct(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st) {
  constexpr int CK = S2 ? 16 : 32, ROWB = CK * 2;
  const TctEncodeTiledFn enc = tct_encode_fn();
  if (!enc) return NND_ERR_ARG;
  if (((size_t)in & 15) || ((size_t)w & 15) || g.T % TG) return NND_ERR_ARG;
  TctHostPlan hp;
  TctArgs& tl = hp.tl;
  tl.DB = (g.Ld + MT - 1) / MT; tl.HB = (g.Lh + TBH - 1) / TBH; tl.WB = (g.Lw + TBW - 1) / TBW; tl.NT = ep.CoutPad / N_TILE;
  tl.total = g.N * tl.DB * tl.HB * tl.WB * tl.NT;
  if (tl.total <= 0) return NND_OK;
  tct_host_plan(g, MT, ROWB, hp);
  TctMaps maps;
  const CUtensorMapSwizzle swz = S2 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
  for (int p = 0; p < tl.n_planes; ++p) {
    const cuuint64_t gdim[5] = {(cuuint64_t)g.Cin, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Di, (cuuint64_t)g.N};
    const cuuint64_t gstride[4] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.Wi * g.Cin * 2, (cuuint64_t)g.Hi * g.Wi * g.Cin * 2,
                                   (cuuint64_t)g.Di * g.Hi * g.Wi * g.Cin * 2};
    const cuuint32_t box[5] = {(cuuint32_t)CK, (cuuint32_t)hp.box[p][0], (cuuint32_t)hp.box[p][1], (cuuint32_t)hp.box[p][2], 1};
    const cuuint32_t estr[5] = {1, (cuuint32_t)g.sw, (cuuint32_t)g.sh, (cuuint32_t)g.sd, 1};
    if (enc(&maps.pl[p], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16*>(in), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NND_ERR_ARG;
  }
  for (int p = tl.n_planes; p < T_MAX_PLANES; ++p) maps.pl[p] = maps.pl[0];
  const int a_stage_bytes = hp.a_stage_bytes;
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)g.Cin, (cuuint64_t)(hp.max_tw + 1) * ep.CoutPad};
    const cuuint64_t gstride[1] = {(cuuint64_t)g.Cin * 2};
    const cuuint32_t box[2] = {(cuuint32_t)CK, (cuuint32_t)N_TILE};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(w), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NND_ERR_ARG;
  }
  const size_t smem = (size_t)T_A_STAGES * a_stage_bytes + (size_t)B_SLOTS * TG * N_TILE * ROWB + 1024;
  if (smem + 6 * 1024 > 232448) return NND_ERR_ARG;          // dynamic + static (statistics rows, barriers) must fit the 227 KB opt-in limit
  static size_t attr_smem[64] = {};                          // per device: the opt-in size set so far
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (smem > attr_smem[dev]) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_tct_kernel<N_TILE, MT, TG, ACC, B_SLOTS, S2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[dev] = smem;
  }
  const int grid = tl.total < NND_NUM_SMS ? tl.total : NND_NUM_SMS;
  conv_tct_kernel<N_TILE, MT, TG, ACC, B_SLOTS, S2><<<grid, tct_threads(MT), smem, st>>>(maps, g, ep, tl, a_stage_bytes);
  NND_LAUNCH_CHECK("conv_tct_kernel");
  return NND_OK;
}

}  // namespace

int nnd_conv_tc_supported(const ConvGeom& g, const ConvEpilogue& ep);
int nnd_conv_tc_s2_supported(const ConvGeom& g, const ConvEpilogue& ep);

// The shapes of the cp.async tile kernel (conv_tc.cu) plus the 2-tap parity classes of a stride-2 dgrad (with one producer lane the
// per-item bookkeeping that kept them on the mma.sync kernel is gone).  NND_ERR_ARG from the launcher (no tensor map / too much shared
// memory) sends the caller back to the cp.async kernels.
int nnd_conv_tct_supported(const ConvGeom& g, const ConvEpilogue& ep) {
  if ((long long)g.N * g.Di * g.Hi * g.Wi >= (1ll << 31) || g.Cin % 32) return 0;
  if (nnd_conv_tc_supported(g, ep)) return 1;
  if (g.T < 2 || g.T >= 4) return 0;
  ConvGeom g4 = g;                       // same predicate with the tap-count floor lifted: a 2- / 3-tap class of a strided dgrad
  g4.T = 4;
  for (int t = g.T; t < 4; ++t) { g4.off_d[t] = g.off_d[0]; g4.off_h[t] = g.off_h[0]; g4.off_w[t] = g.off_w[0]; g4.tap_w[t] = g.tap_w[0]; }
  const bool identity = g.omd == 1 && g.omh == 1 && g.omw == 1 && !g.ood && !g.ooh && !g.oow;
  return !identity && nnd_conv_tc_supported(g4, ep);
}
int nnd_conv_tct_s2_supported(const ConvGeom& g, const ConvEpilogue& ep) {
  if ((long long)g.N * g.Di * g.Hi * g.Wi >= (1ll << 31)) return 0;
  return nnd_conv_tc_s2_supported(g, ep) && g.Cin % 16 == 0;
}

int nnd_conv_tct(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st) {
  if (!nnd_conv_tct_supported(g, ep)) return NND_ERR_ARG;
  const bool g3 = g.T % 3 == 0;
  if (ep.CoutPad % 128 == 0) {
    const long long tiles4 = (long long)g.N * ((g.Ld + 3) / 4) * ((g.Lh + TBH - 1) / TBH) * ((g.Lw + TBW - 1) / TBW) * (ep.CoutPad / 128);
    if (tiles4 >= NND_NUM_SMS) return g3 ? launch_tct<128, 4, 3, 1, 3, 0>(in, w, g, ep, st) : launch_tct<128, 4, 1, 1, 8, 0>(in, w, g, ep, st);
    return g3 ? launch_tct<128, 2, 3, 2, 4, 0>(in, w, g, ep, st) : launch_tct<128, 2, 1, 2, 12, 0>(in, w, g, ep, st);
  }
  if (ep.CoutPad % 64 == 0) return g3 ? launch_tct<64, 4, 3, 2, 6, 0>(in, w, g, ep, st) : launch_tct<64, 4, 1, 2, 12, 0>(in, w, g, ep, st);
  return g3 ? launch_tct<32, 4, 3, 2, 6, 0>(in, w, g, ep, st) : launch_tct<32, 4, 1, 2, 12, 0>(in, w, g, ep, st);
}

int nnd_conv_tct_s2(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st) {
  if (!nnd_conv_tct_s2_supported(g, ep)) return NND_ERR_ARG;
  const bool g3 = g.T % 3 == 0;
  if (ep.CoutPad % 128 == 0) return g3 ? launch_tct<128, 2, 3, 2, 3, 1>(in, w, g, ep, st) : launch_tct<128, 2, 1, 2, 8, 1>(in, w, g, ep, st);
  if (ep.CoutPad % 64 == 0) return g3 ? launch_tct<64, 2, 3, 2, 4, 1>(in, w, g, ep, st) : launch_tct<64, 2, 1, 2, 8, 1>(in, w, g, ep, st);
  return g3 ? launch_tct<32, 2, 3, 2, 4, 1>(in, w, g, ep, st) : launch_tct<32, 2, 1, 2, 8, 1>(in, w, g, ep, st);
}

// Host-only: the plan of a launch as flat ints (tests/test_tma_plan_cpu.py replays it in numpy against a direct gather).
//   out = [n_planes, a_stage_bytes, a_bytes, ROWB,  per plane: pd, ph, pw, cd, ch, cw, base, box_w, box_h, box_d,  per tap: tap_off, tap_sbo, slice_step]
// Returns the number of ints written, or -1 (bad geometry / buffer too small).  No CUDA call.
extern "C" int nnd_conv_tct_plan_debug(const int* geom, int mt, int s2, int* out, int cap) {
  if (!geom || !out || mt < 1 || mt > 4) return -1;
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
  const int rowb = s2 ? 32 : 64;
  TctHostPlan hp;
  tct_host_plan(g, mt, rowb, hp);
  const int need = 4 + hp.tl.n_planes * 10 + g.T * 3;
  if (cap < need) return -1;
  int* o = out;
  *o++ = hp.tl.n_planes; *o++ = hp.a_stage_bytes; *o++ = hp.tl.a_bytes; *o++ = rowb;
  for (int p = 0; p < hp.tl.n_planes; ++p) {
    const TctPlane& pl = hp.tl.pl[p];
    *o++ = pl.pd; *o++ = pl.ph; *o++ = pl.pw; *o++ = pl.cd; *o++ = pl.ch; *o++ = pl.cw; *o++ = pl.base;
    *o++ = hp.box[p][0]; *o++ = hp.box[p][1]; *o++ = hp.box[p][2];
  }
  for (int t = 0; t < g.T; ++t) { *o++ = hp.tl.tap_off[t]; *o++ = hp.tl.tap_sbo[t]; *o++ = hp.tl.slice_step[t]; }
  return need;
}
