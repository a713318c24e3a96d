// Generic gather convolution on the legacy tensor path (mma.sync m16n8k16 bf16, fp32 accumulate).
//
// This is the always-available kernel: it serves every layer form of conv_common.cuh (strided conv, its two
// dgrad forms, transposed conv and its dgrad, 1x1x1, padded head outputs) and is the numerical cross-check for the
// tcgen05 kernel in conv_tc.cu, which takes over the 3x3x3 stride-1 layers that carry ~90 % of the FLOPs.
// Replaces cuDNN's implicit-GEMM fprop/dgrad as invoked by torch.nn.Conv3d / ConvTranspose3d in the reference
// (nndet/arch/conv.py:344-348) with fused epilogues: bias, residual add (decoder top-down sum,
// nndet/arch/decoder/base.py:405), the regressor's learnable scale (nndet/arch/heads/regressor.py:164-165),
// per-(sample, channel) norm statistics, and direct writes into the [N, anchors, C] head layout
// (the permute/contiguous/cat of nndet/arch/heads/classifier.py:176-180 disappears).
//
// CTA tile 128 voxels x BN channels, K step 32 channels of one tap, 4-stage cp.async pipeline, 8 warps (4 x 2),
// XOR-swizzled 64-byte smem rows (conflict-free ldmatrix), zero-fill cp.async for the padding halo.
#include "conv_common.cuh"

namespace {

constexpr int BM = 128, BK = 32, STAGES = 4, THREADS = 256;

__device__ __forceinline__ int swz64(int row, int chunk) { return chunk ^ ((row >> 1) & 3); }

template <int BN>
__global__ void __launch_bounds__(THREADS)
conv_igemm_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ wgt, const ConvGeom g,
                  const ConvEpilogue ep, int tiles_per_sample) {
  constexpr int NI = BN / 16;                // n8 tiles per warp (warp covers BN/2 columns)
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;
  unsigned char* sB = smem + STAGES * A_BYTES;
  __shared__ float s_stat[2][4][BN];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp & 3, warp_n = warp >> 2;
  const int n = blockIdx.x / tiles_per_sample;
  const int tile = blockIdx.x % tiles_per_sample;
  const int n0 = blockIdx.y * BN;
  const int Lvox = g.Ld * g.Lh * g.Lw;
  const int KC = g.Cin / BK;
  const int KI = g.T * KC;

  // ---- producer bookkeeping: this thread's A row
  const int a_row = tid >> 1;
  const int a_c0 = (tid & 1) * 2;
  const int m_load = tile * BM + a_row;
  const bool row_ok = m_load < Lvox;
  int ld = 0, lh = 0, lw = 0;
  if (row_ok) { lw = m_load % g.Lw; int r = m_load / g.Lw; lh = r % g.Lh; ld = r / g.Lh; }
  const int id0 = ld * g.sd, ih0 = lh * g.sh, iw0 = lw * g.sw;
  const __nv_bfloat16* in_n = in + (size_t)n * g.Di * g.Hi * g.Wi * g.Cin;

  int p_tap = 0, p_kc = 0;                     // next (tap, channel chunk) to load
  auto load_stage = [&](int stage) {
    // A: gathered input rows (zero-filled outside the volume)
    const int id = id0 + g.off_d[p_tap], ih = ih0 + g.off_h[p_tap], iw = iw0 + g.off_w[p_tap];
    const bool ok = row_ok && (unsigned)id < (unsigned)g.Di && (unsigned)ih < (unsigned)g.Hi && (unsigned)iw < (unsigned)g.Wi;
    const __nv_bfloat16* src = ok ? in_n + ((size_t)(id * g.Hi + ih) * g.Wi + iw) * g.Cin + p_kc * BK : in;
    const unsigned a_base = smem_u32(sA + stage * A_BYTES) + a_row * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c)
      cp_async16(a_base + swz64(a_row, a_c0 + c) * 16, src + (a_c0 + c) * 8, ok);
    // B: weight slice [BN rows (co)] x [32 ci]
    const __nv_bfloat16* wsrc = wgt + ((size_t)g.tap_w[p_tap] * ep.CoutPad + n0) * g.Cin + p_kc * BK;
    const unsigned b_base = smem_u32(sB + stage * B_BYTES);
#pragma unroll
    for (int i = tid; i < BN * 4; i += THREADS) {
      const int r = i >> 2, c = i & 3;
      cp_async16(b_base + r * 64 + swz64(r, c) * 16, wsrc + (size_t)r * g.Cin + c * 8, true);
    }
    if (++p_kc == KC) { p_kc = 0; ++p_tap; }
  };

  float acc[2][NI][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KI) load_stage(s);
    cp_async_commit();
  }

  for (int ki = 0; ki < KI; ++ki) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (ki + STAGES - 1 < KI) load_stage((ki + STAGES - 1) % STAGES);
    cp_async_commit();

    const int stage = ki % STAGES;
    const unsigned a_st = smem_u32(sA + stage * A_BYTES);
    const unsigned b_st = smem_u32(sB + stage * B_BYTES);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      unsigned af[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r = warp_m * 32 + mi * 16 + (lane & 15);
        const int c = kk * 2 + (lane >> 4);
        ldmatrix_x4(a_st + r * 64 + swz64(r, c) * 16, af[mi][0], af[mi][1], af[mi][2], af[mi][3]);
      }
#pragma unroll
      for (int nj = 0; nj < NI / 2; ++nj) {
        unsigned b0, b1, b2, b3;
        const int r = warp_n * (BN / 2) + nj * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(b_st + r * 64 + swz64(r, c) * 16, b0, b1, b2, b3);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          mma_bf16_16816(acc[mi][nj * 2 + 0], af[mi][0], af[mi][1], af[mi][2], af[mi][3], b0, b1);
          mma_bf16_16816(acc[mi][nj * 2 + 1], af[mi][0], af[mi][1], af[mi][2], af[mi][3], b2, b3);
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---------------------------------------------------------------- epilogue
  const int gq = lane >> 2, tq = lane & 3;
  const float scale = ep.scale ? *ep.scale : 1.f;
  const bool do_stats = ep.stat_sum != nullptr;

  if (!ep.out_fp32 && (ep.Cout & 7) == 0 && (ep.out_v_stride & 7) == 0 && (ep.out_n_stride & 7) == 0) {
    // Vector path (every bf16 NDHWC output): the fp32 tile goes through shared memory so that each thread then owns 8
    // consecutive channels of a voxel -- one 16-byte residual load and one 16-byte store instead of 2-byte / 4-byte
    // accesses.  Most of these launches (1x1x1 laterals, transposed convs, stride-2 layers at 128^3 / 64^3) are bound by
    // exactly this traffic.
    constexpr int OP = BN + 4;                               // padded fp32 row (bank spread)
    float* so = reinterpret_cast<float*>(smem);
    __syncthreads();                                         // all warps are done reading the last stage
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = warp_m * 32 + mi * 16 + gq + half * 8;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          const int c = warp_n * (BN / 2) + ni * 8 + tq * 2;
          *reinterpret_cast<float2*>(&so[r * OP + c]) = make_float2(acc[mi][ni][half * 2], acc[mi][ni][half * 2 + 1]);
        }
      }
    __syncthreads();
    constexpr int CH = BN / 8;                               // 8-channel chunks per row
    const int ch = tid % CH;
    const int co = n0 + ch * 8;
    float bsum[8], bsq[8], bias8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { bsum[j] = 0.f; bsq[j] = 0.f; bias8[j] = (ep.bias && co + j < ep.Cout) ? ep.bias[co + j] : 0.f; }
    if (co < ep.Cout) {
      for (int r = tid / CH; r < BM; r += THREADS / CH) {
        const int m = tile * BM + r;
        if (m >= Lvox) break;
        const int w_ = m % g.Lw; const int r_ = m / g.Lw; const int h_ = r_ % g.Lh; const int d_ = r_ / g.Lh;
        const long long pvox = ((long long)(d_ * g.omd + g.ood) * g.Ho + (h_ * g.omh + g.ooh)) * g.Wo + (w_ * g.omw + g.oow);
        const float4 p0 = *reinterpret_cast<const float4*>(&so[r * OP + ch * 8]);
        const float4 p1 = *reinterpret_cast<const float4*>(&so[r * OP + ch * 8 + 4]);
        float v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += bias8[j];
        if (ep.residual) {
          const uint4 rv = *reinterpret_cast<const uint4*>(ep.residual + ((long long)n * g.Do * g.Ho * g.Wo + pvox) * ep.Cout + co);
          const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float2 t2 = __bfloat1622float2(hp[k]); v[2 * k] += t2.x; v[2 * k + 1] += t2.y; }
        }
        __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          pk[k] = __floats2bfloat162_rn(v[2 * k] * scale, v[2 * k + 1] * scale);
          const float2 r2 = __bfloat1622float2(pk[k]);       // statistics of the stored tensor
          bsum[2 * k] += r2.x; bsum[2 * k + 1] += r2.y; bsq[2 * k] += r2.x * r2.x; bsq[2 * k + 1] += r2.y * r2.y;
        }
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)n * ep.out_n_stride + pvox * ep.out_v_stride + co;
        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(pk);
      }
    }
    if (do_stats) {
      __syncthreads();                                       // the tile in shared memory is consumed: reuse it for the sums
      float* ss = so;                                        // [2][BN]
      if (tid < 2 * BN) ss[tid] = 0.f;
      __syncthreads();
      if (co < ep.Cout) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&ss[ch * 8 + j], bsum[j]); atomicAdd(&ss[BN + ch * 8 + j], bsq[j]); }
      }
      __syncthreads();
      if (tid < BN && n0 + tid < ep.Cout) {
        atomicAdd(&ep.stat_sum[(size_t)n * ep.Cout + n0 + tid], ss[tid]);
        atomicAdd(&ep.stat_sq[(size_t)n * ep.Cout + n0 + tid], ss[BN + tid]);
      }
    }
    return;
  }

  float csum[NI][2], csq[NI][2];
#pragma unroll
  for (int j = 0; j < NI; ++j) { csum[j][0] = csum[j][1] = csq[j][0] = csq[j][1] = 0.f; }

#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = tile * BM + warp_m * 32 + mi * 16 + gq + half * 8;
      if (m >= Lvox) continue;
      const int w_ = m % g.Lw; const int r_ = m / g.Lw; const int h_ = r_ % g.Lh; const int d_ = r_ / g.Lh;
      const long long pvox = ((long long)(d_ * g.omd + g.ood) * g.Ho + (h_ * g.omh + g.ooh)) * g.Wo + (w_ * g.omw + g.oow);
      const long long obase = (long long)n * ep.out_n_stride + pvox * ep.out_v_stride;
      const long long rbase = ((long long)n * g.Do * g.Ho * g.Wo + pvox) * ep.Cout;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int co = n0 + warp_n * (BN / 2) + ni * 8 + tq * 2;
        float v0 = acc[mi][ni][half * 2 + 0], v1 = acc[mi][ni][half * 2 + 1];
        const bool ok0 = co < ep.Cout, ok1 = co + 1 < ep.Cout;
        if (ep.bias) { if (ok0) v0 += ep.bias[co]; if (ok1) v1 += ep.bias[co + 1]; }
        if (ep.residual) {
          if (ok0) v0 += __bfloat162float(ep.residual[rbase + co]);
          if (ok1) v1 += __bfloat162float(ep.residual[rbase + co + 1]);
        }
        v0 *= scale; v1 *= scale;
        if (ep.out_fp32) {
          float* o = reinterpret_cast<float*>(ep.out) + obase + co;
          if (ok0) o[0] = v0;
          if (ok1) o[1] = v1;
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + obase + co;
          const __nv_bfloat16 h0 = __float2bfloat16(v0), h1 = __float2bfloat16(v1);
          if (ok1 && (((obase + co) & 1) == 0)) {
            *reinterpret_cast<__nv_bfloat162*>(o) = __halves2bfloat162(h0, h1);
          } else {
            if (ok0) o[0] = h0;
            if (ok1) o[1] = h1;
          }
          v0 = __bfloat162float(h0); v1 = __bfloat162float(h1);     // statistics of the stored tensor
        }
        if (do_stats) {
          if (ok0) { csum[ni][0] += v0; csq[ni][0] += v0 * v0; }
          if (ok1) { csum[ni][1] += v1; csq[ni][1] += v1 * v1; }
        }
      }
    }

  if (do_stats) {
#pragma unroll
    for (int ni = 0; ni < NI; ++ni)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float s = csum[ni][j], q = csq[ni][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (gq == 0) {
          const int col = warp_n * (BN / 2) + ni * 8 + tq * 2 + j;
          s_stat[0][warp_m][col] = s; s_stat[1][warp_m][col] = q;
        }
      }
    __syncthreads();
    if (tid < BN) {
      const int co = n0 + tid;
      if (co < ep.Cout) {
        float s = s_stat[0][0][tid] + s_stat[0][1][tid] + s_stat[0][2][tid] + s_stat[0][3][tid];
        float q = s_stat[1][0][tid] + s_stat[1][1][tid] + s_stat[1][2][tid] + s_stat[1][3][tid];
        atomicAdd(&ep.stat_sum[(size_t)n * ep.Cout + co], s);
        atomicAdd(&ep.stat_sq[(size_t)n * ep.Cout + co], q);
      }
    }
  }
}

template <int BN>
int launch_igemm(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st) {
  const int Lvox = g.Ld * g.Lh * g.Lw;
  const int tiles = (Lvox + BM - 1) / BM;
  size_t smem = (size_t)STAGES * (BM * BK * 2 + BN * BK * 2);
  if (smem < (size_t)BM * (BN + 4) * 4) smem = (size_t)BM * (BN + 4) * 4;      // fp32 output tile of the vector epilogue
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((unsigned)(tiles * g.N), (unsigned)(ep.CoutPad / BN));
  conv_igemm_kernel<BN><<<grid, THREADS, smem, st>>>(in, w, g, ep, tiles);
  NND_LAUNCH_CHECK("conv_igemm_kernel");
  return NND_OK;
}

}  // namespace

// in [N,Di,Hi,Wi,Cin] bf16 with Cin % 32 == 0; w [taps][CoutPad][Cin] bf16 with CoutPad % 32 == 0.
int nnd_conv_igemm(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep,
                   cudaStream_t st) {
  if (!in || !w || !ep.out) return NND_ERR_ARG;
  if (g.Cin % 32 != 0 || ep.CoutPad % 32 != 0 || g.T < 1 || g.T > NND_MAX_TAPS) return NND_ERR_ARG;
  if (g.N <= 0 || g.Ld * g.Lh * g.Lw <= 0) return NND_OK;
  if (ep.CoutPad % 128 == 0) return launch_igemm<128>(in, w, g, ep, st);
  if (ep.CoutPad % 64 == 0) return launch_igemm<64>(in, w, g, ep, st);
  return launch_igemm<32>(in, w, g, ep, st);
}
