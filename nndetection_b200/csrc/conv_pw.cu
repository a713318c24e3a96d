// Pointwise (single-tap) convolutions as a TMA-fed tcgen05 GEMM: the 1x1x1 laterals of the U-FPN and their input gradients
// (nndet/arch/decoder/base.py:216-241), the parity classes of the kernel == stride up-convolutions (:272-304) and -- stacked along N
// -- a whole up-convolution in ONE launch.  Replaces cuDNN's 1x1x1 / ConvTranspose3d behind nndet/arch/conv.py:344-348 (and the
// mma.sync gather kernel conv_igemm.cu that served these forms in round 1).
//
//     OUT[m, n] = sum_k A[m, k] * W[n, k]  (+ bias[n mod Cout])  (+ residual)          m = input voxel (NDHWC row), k = Cin
//
// A is the activation tensor itself viewed as a [M = N*D*H*W, K = Cin] row-major matrix, W the K-major weight pack
// [taps * CoutPad, Cin] (nnd_pack_weights).  These launches move 2 * (Cin + Cout) bytes per voxel for 2 * Cin * Cout flop: HBM-bound
// by two orders of magnitude, so the design goal is a deep, cheap copy pipeline, not MMA efficiency:
//   * warp 0, one lane: TMA producer.  `cp.async.bulk.tensor.2d` (128B / 64B swizzle) brings a [128 rows x BK] box of A and a
//     [BN x BK] box of W into a ring of STAGES slots; rows beyond M are zero-filled by the TMA unit (no per-thread address
//     arithmetic, no predication, one instruction per 8-16 KB), completion lands on the slot's mbarrier (`complete_tx`).
//   * warp 1, one lane: tcgen05.mma issuer, M = 128 x N = BN x K = 16 per instruction straight from the swizzled slots (K-major
//     shared-memory descriptors, SBO = 8 rows * row bytes; the K advance inside the swizzle atom is +32 bytes on the start address),
//     accumulators double-buffered in TMEM; `tcgen05.commit` releases the slot / publishes the accumulator.
//   * warps 2-5: epilogue, TMEM -> registers -> bias / residual -> bf16 -> 64-byte row stores at (voxel * om + oo[tap]) -- the output
//     mapping of the gather-convolution form (conv_common.cuh), so a parity class of an up-convolution or (stacked mode) all of its
//     taps scatter into the fine grid directly, the lateral added on the way (decoder/base.py:405).
// Persistent grid (one CTA per SM), tiles = (m block, n block) with n fastest so that the A box of an m block is re-read from L2.
#include <cuda.h>

#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int PW_BM = 128;
constexpr int PW_THREADS = 10 * 32;            // TMA producer, MMA issuer, 2 x 4 epilogue warps

struct PwArgs {
  long long M;                     // rows of A (voxels)
  int K, Ntot, Cout;               // Ntot = taps * Cout columns; tap of column n = n / Cout
  int BN_tiles;                    // Ntot / BN
  int m_tiles;
  // output mapping: row m = ((n * D + z) * H + y) * W + x  ->  voxel ((n * Do + z*omd + od) * Ho + y*omh + oh) * Wo + x*omw + ow
  int D, H, W, Do, Ho, Wo, omd, omh, omw;
  int identity;                    // output voxel == m (no decode needed)
  signed char od[8], oh[8], ow[8]; // per tap
  __nv_bfloat16* out;
  const float* bias;               // [Cout] or null
  const __nv_bfloat16* residual;   // same layout as out, or null
};

__device__ __forceinline__ void pw_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void pw_tma_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// K-major operand in a swizzled slot: rows of ROWB = BK * 2 bytes (= the swizzle span), 8-row groups ROWB * 8 bytes apart
template <int ROWB>
__device__ __forceinline__ unsigned long long pw_desc(unsigned addr) {
  constexpr unsigned long long LAYOUT = ROWB == 128 ? 2ull : 4ull;          // SWIZZLE_128B : SWIZZLE_64B
  return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((ROWB * 8) >> 4) << 32) | (1ull << 46) | (LAYOUT << 61);
}

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(PW_THREADS, 1)
conv_pw_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const PwArgs a) {
  constexpr int ROWB = BK * 2;                              // bytes per operand row in a slot = swizzle span
  constexpr int A_BYTES = PW_BM * ROWB, B_BYTES = BN * ROWB;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // accumulator stages: the chain MMA -> commit -> epilogue wake-up -> tcgen05.ld -> stores -> TEMPTY -> next MMA costs ~2000 cycles per
  // stage (ncu of the 2-stage version: the epilogue warps polled TFULL 4x per tile), HBM needs ~700 per 128 x 32 tile: 8 stages hide it
  constexpr int ACC = 512 / BN > 8 ? 8 : 512 / BN;
  constexpr int TMEM_COLS = ACC * BN >= 512 ? 512 : (ACC * BN >= 256 ? 256 : (ACC * BN >= 128 ? 128 : 64));
  constexpr unsigned IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(BN >> 3) << 17) | ((128u >> 4) << 24);
  static_assert(STAGE_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "swizzled slots need 1024-byte alignment");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the dynamic segment is only guaranteed 16-byte alignment: round up by hand (the launcher allocates 1 KB of slack)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long bars[2 * STAGES + 2 * ACC];
  __shared__ unsigned s_tmem_base;
  __shared__ float s_bias[1024];                            // [Cout] (zeros without a bias)
  __shared__ int s_tapoff[8];                               // output-voxel offset of a tap: (od * Ho + oh) * Wo + ow
  __shared__ uint4 s_stage[8][128];                         // per epilogue warp: 32 rows x 64 bytes, transposition buffer
  const unsigned bar0 = smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (STAGES + i); };
  auto TFULL = [&](int i) { return bar0 + 8u * (2 * STAGES + i); };
  auto TEMPTY = [&](int i) { return bar0 + 8u * (2 * STAGES + ACC + i); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = a.K / BK;
  const int total = a.m_tiles * a.BN_tiles;                 // < 2^31 (host check)

  for (int i = tid; i < a.Cout; i += PW_THREADS) s_bias[i] = a.bias ? a.bias[i] : 0.f;
  if (tid < 8) s_tapoff[tid] = (a.od[tid] * a.Ho + a.oh[tid]) * a.Wo + a.ow[tid];
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(FULL(i), 1); mbar_init(EMPTY(i), 1); }
    for (int i = 0; i < ACC; ++i) { mbar_init(TFULL(i), 1); mbar_init(TEMPTY(i), BN == 32 ? 4 : 8); }
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

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      unsigned stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int mb = tile / a.BN_tiles, nb = tile - mb * a.BN_tiles;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(EMPTY(stage), phase ^ 1);
          const unsigned sa = smem_u32(smem + stage * STAGE_BYTES);
          pw_expect_tx(FULL(stage), STAGE_BYTES);
          pw_tma_2d(sa, &map_a, kb * BK, mb * PW_BM, FULL(stage));
          pw_tma_2d(sa + A_BYTES, &map_w, kb * BK, nb * BN, FULL(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      unsigned stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        mbar_wait(TEMPTY(acc), acc_phase ^ 1);
        tc_fence_after();
        const unsigned d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(FULL(stage), phase);
          tc_fence_after();
          const unsigned sa = smem_u32(smem + stage * STAGE_BYTES);
          const unsigned long long da = pw_desc<ROWB>(sa), db = pw_desc<ROWB>(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma(d_tmem, da + (unsigned long long)(k * 2), db + (unsigned long long)(k * 2), IDESC, (kb | k) != 0 ? 1u : 0u);
          tc_commit(EMPTY(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(TFULL(acc));
        if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================================================ epilogue: two groups of four warps (TMEM lane quarter = warp % 4).
    // The per-tile chain TFULL wait -> tcgen05.ld -> convert -> stores is latency-, not throughput-bound (ncu of the first version: one
    // group needed ~1200 cycles per 128 x 32 tile against ~690 cycles of HBM time), so two groups work concurrently: single-chunk tiles
    // (BN = 32) alternate between the groups (group g owns accumulator stage g), wider tiles split their 32-column chunks by parity.
    constexpr int CH = BN / 32;
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const unsigned cout = (unsigned)a.Cout;
    int local = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++local) {
      if (CH == 1 && (local & 1) != grp) continue;
      const unsigned acc = (unsigned)(local % ACC), acc_phase = (unsigned)((local / ACC) & 1);
      const int mb = tile / a.BN_tiles, nb = tile - mb * a.BN_tiles;
      const long long m = (long long)mb * PW_BM + row;
      const bool ok = m < a.M;
      unsigned vbase0 = (unsigned)m;                           // output voxel of tap offset (0, 0, 0); voxel counts fit 32 bits (host check)
      if (!a.identity && ok) {
        unsigned t = (unsigned)m;
        const unsigned xc = t % (unsigned)a.W; t /= (unsigned)a.W;
        const unsigned yc = t % (unsigned)a.H; t /= (unsigned)a.H;
        const unsigned zc = t % (unsigned)a.D; const unsigned nn = t / (unsigned)a.D;
        vbase0 = ((nn * a.Do + zc * a.omd) * a.Ho + yc * a.omh) * a.Wo + xc * a.omw;
      }
      auto out_offset = [&](int c, unsigned& co0) -> long long {
        const unsigned n0 = (unsigned)(nb * BN + c * 32);
        const unsigned tap = a.identity ? 0u : n0 / cout;
        co0 = n0 - tap * cout;
        return (long long)(vbase0 + (unsigned)s_tapoff[tap]) * cout + co0;
      };
      const int c_first = CH == 1 ? 0 : grp;
      mbar_wait_warp(TFULL(acc), acc_phase, lane);
      tc_fence_after();
      // Global accesses of a chunk (64 bytes per row) go through a per-warp 2 KB transposition buffer: a lane OWNS one row (its
      // TMEM lane), but instruction u of a warp-wide 16-byte access covers rows 8u .. 8u+7 with FOUR LANES PER ROW (lane l -> row
      // 8u + l/4, piece l%4), i.e. whole 32-byte sectors / 64-byte chunks per row.  (The first version let every lane store its own row:
      // 32 half-written sectors per instruction -- ncu: 33.5 M sector writes for 16.8 M sectors of output, the L1 store path at ~0.5
      // sectors/clk was the kernel's limiter at 50 % of the HBM peak.)  Buffer slot of (row r, piece p): r*4 + (p ^ ((r >> 1) & 3)).
      uint4* stg = s_stage[warp - 2];
      const int sub_row = lane >> 2, piece = lane & 3;
#pragma unroll 1
      for (int c = c_first; c < CH; c += (CH == 1 ? 1 : 2)) {
        unsigned v[32];
        tmem_ld32_issue(tmem_base + ((unsigned)(q * 32) << 16) + acc * BN + c * 32, v);
        unsigned co0;
        const long long o = out_offset(c, co0);                 // this lane's row
        long long o_r[4]; bool ok_r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          o_r[u] = __shfl_sync(0xffffffffu, o, u * 8 + sub_row);
          ok_r[u] = __shfl_sync(0xffffffffu, (int)ok, u * 8 + sub_row) != 0;
        }
        uint4 rr[4];
        if (a.residual) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            rr[u] = ok_r[u] ? *reinterpret_cast<const uint4*>(a.residual + o_r[u] + piece * 8) : make_uint4(0, 0, 0, 0);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + s_bias[co0 + j];
        if (a.residual) {
#pragma unroll
          for (int u = 0; u < 4; ++u) { const int r = u * 8 + sub_row; stg[r * 4 + (piece ^ ((r >> 1) & 3))] = rr[u]; }
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint4 rv = stg[lane * 4 + (u ^ ((lane >> 1) & 3))];
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 t2 = __bfloat1622float2(hp[k]);
              f[u * 8 + 2 * k] += t2.x; f[u * 8 + 2 * k + 1] += t2.y;
            }
          }
          __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 pv;
          __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pv);
#pragma unroll
          for (int k = 0; k < 4; ++k) hp[k] = __floats2bfloat162_rn(f[u * 8 + 2 * k], f[u * 8 + 2 * k + 1]);
          stg[lane * 4 + (u ^ ((lane >> 1) & 3))] = pv;
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = u * 8 + sub_row;
          if (ok_r[u]) *reinterpret_cast<uint4*>(a.out + o_r[u] + piece * 8) = stg[r * 4 + (piece ^ ((r >> 1) & 3))];
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point query: the library must not carry a link-time dependency on
// libcuda.so.1 (it is loaded on GPU-less build hosts for the ABI checks)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map(CUtensorMap* map, const void* base, long long rows, int K, int box_rows, int BK) {
  const EncodeTiledFn cuTensorMapEncodeTiled = encode_tiled_fn();
  if (!cuTensorMapEncodeTiled) return NND_ERR_CUDA;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NND_OK : NND_ERR_CUDA;
}

template <int BN, int BK, int STAGES>
int launch_pw(const __nv_bfloat16* x, const __nv_bfloat16* w, PwArgs a, cudaStream_t st) {
  CUtensorMap map_a, map_w;
  if (make_map(&map_a, x, a.M, a.K, PW_BM, BK) != NND_OK || make_map(&map_w, w, a.Ntot, a.K, BN, BK) != NND_OK)
    return nnd_set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
  a.BN_tiles = a.Ntot / BN;
  a.m_tiles = (int)((a.M + PW_BM - 1) / PW_BM);
  if ((long long)a.m_tiles * a.BN_tiles >= (1ll << 31) || a.Cout > 1024) return NND_ERR_ARG;
  constexpr size_t SMEM = (size_t)STAGES * (PW_BM + BN) * BK * 2 + 1024;
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_pw_kernel<BN, BK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  const long long total = (long long)a.m_tiles * a.BN_tiles;
  const int grid = total < NND_NUM_SMS ? (int)total : NND_NUM_SMS;
  conv_pw_kernel<BN, BK, STAGES><<<grid, PW_THREADS, SMEM, st>>>(map_a, map_w, a);
  NND_LAUNCH_CHECK("conv_pw_kernel");
  return NND_OK;
}

int dispatch_pw(const __nv_bfloat16* x, const __nv_bfloat16* w, const PwArgs& a, cudaStream_t st) {
  const int bn = a.Ntot % 256 == 0 ? 256 : (a.Ntot % 128 == 0 ? 128 : (a.Ntot % 64 == 0 ? 64 : 32));
  if (a.K % 64 == 0) {
    switch (bn) {
      case 256: return launch_pw<256, 64, 4>(x, w, a, st);
      case 128: return launch_pw<128, 64, 6>(x, w, a, st);
      case 64: return launch_pw<64, 64, 8>(x, w, a, st);
      default: return launch_pw<32, 64, 8>(x, w, a, st);
    }
  }
  switch (bn) {
    case 256: return launch_pw<256, 32, 8>(x, w, a, st);
    case 128: return launch_pw<128, 32, 10>(x, w, a, st);
    case 64: return launch_pw<64, 32, 12>(x, w, a, st);
    default: return launch_pw<32, 32, 12>(x, w, a, st);
  }
}

}  // namespace

// Single-tap gathers with stride-1 input addressing: 1x1x1 convolutions (identity output mapping) and the parity classes of a
// kernel == stride transposed convolution (output voxel = lo * om + oo).  bf16 dense output, no norm statistics.
int nnd_conv_pw_supported(const ConvGeom& g, const ConvEpilogue& ep) {
  if (g.T != 1 || g.off_d[0] || g.off_h[0] || g.off_w[0]) return 0;
  if (g.sd != 1 || g.sh != 1 || g.sw != 1) return 0;
  if (g.Ld != g.Di || g.Lh != g.Hi || g.Lw != g.Wi) return 0;              // every input voxel is one row of A
  if (g.Cin % 32 || ep.Cout % 32 || ep.CoutPad != ep.Cout) return 0;
  if (ep.out_fp32 || ep.scale || ep.stat_sum || ep.stat_sq) return 0;
  if (ep.out_v_stride != ep.Cout || ep.out_n_stride != (long long)g.Do * g.Ho * g.Wo * ep.Cout) return 0;
  if ((long long)g.N * g.Ld * g.Lh * g.Lw < PW_BM) return 0;              // tiny volumes: a single partial tile is not worth a TMA set-up
  return 1;
}

int nnd_conv_pw(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st) {
  if (!nnd_conv_pw_supported(g, ep)) return NND_ERR_ARG;
  if (((size_t)in & 15) || ((size_t)w & 15)) return NND_ERR_ARG;
  PwArgs a;
  a.M = (long long)g.N * g.Ld * g.Lh * g.Lw; a.K = g.Cin; a.Ntot = ep.Cout; a.Cout = ep.Cout;
  a.D = g.Ld; a.H = g.Lh; a.W = g.Lw; a.Do = g.Do; a.Ho = g.Ho; a.Wo = g.Wo; a.omd = g.omd; a.omh = g.omh; a.omw = g.omw;
  for (int t = 0; t < 8; ++t) { a.od[t] = 0; a.oh[t] = 0; a.ow[t] = 0; }
  a.od[0] = (signed char)g.ood; a.oh[0] = (signed char)g.ooh; a.ow[0] = (signed char)g.oow;
  a.identity = g.omd == 1 && g.omh == 1 && g.omw == 1 && !g.ood && !g.ooh && !g.oow && g.Do == g.Ld && g.Ho == g.Lh && g.Wo == g.Lw;
  a.out = reinterpret_cast<__nv_bfloat16*>(ep.out); a.bias = ep.bias; a.residual = ep.residual;
  return dispatch_pw(in, w + (size_t)g.tap_w[0] * ep.CoutPad * g.Cin, a, st);
}

// A whole kernel == stride transposed convolution (nn.ConvTranspose3d, decoder/base.py:272-304) in one launch: the taps are stacked
// along N (column n = tap * Cout + co of the fprop weight pack [taps][Cout][Cin]), every input voxel is read ONCE and each 32-channel
// column block is scattered to output voxel (2z + a, 2y + b, 2x + c) of its tap, `residual` (the lateral, same layout as out) added.
// x bf16 [N, D, H, W, Cin], out / residual bf16 [N, D*sd, H*sh, W*sw, Cout], s in {1, 2} per axis, taps ordered (a, b, c) row-major.
int nnd_conv_upconv(const void* x, const void* w_packed, int N, int D, int H, int W, int Cin, int Cout, int sd, int sh, int sw,
                    void* out, const float* bias, const void* residual, cudaStream_t st) {
  if (!x || !w_packed || !out || N <= 0 || D <= 0 || H <= 0 || W <= 0) return NND_ERR_ARG;
  if (Cin % 32 || Cout % 32 || sd < 1 || sd > 2 || sh < 1 || sh > 2 || sw < 1 || sw > 2) return NND_ERR_ARG;
  if (((size_t)x & 15) || ((size_t)w_packed & 15) || ((size_t)out & 15) || ((size_t)residual & 15)) return NND_ERR_ARG;
  PwArgs a;
  a.M = (long long)N * D * H * W; a.K = Cin; a.Cout = Cout; a.Ntot = sd * sh * sw * Cout;
  a.D = D; a.H = H; a.W = W; a.Do = D * sd; a.Ho = H * sh; a.Wo = W * sw; a.omd = sd; a.omh = sh; a.omw = sw;
  int t = 0;
  for (int i = 0; i < sd; ++i) for (int j = 0; j < sh; ++j) for (int k = 0; k < sw; ++k, ++t) {
    a.od[t] = (signed char)i; a.oh[t] = (signed char)j; a.ow[t] = (signed char)k;
  }
  for (; t < 8; ++t) { a.od[t] = 0; a.oh[t] = 0; a.ow[t] = 0; }
  a.identity = 0;
  a.out = reinterpret_cast<__nv_bfloat16*>(out); a.bias = bias; a.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  if (a.M < PW_BM) return NND_ERR_ARG;
  return dispatch_pw(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w_packed), a, st);
}
