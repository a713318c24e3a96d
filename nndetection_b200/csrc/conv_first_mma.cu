// Image-input layer (Cin in {1, 2} modalities -> 32 features, 3x3x3, stride 1): forward and weight gradient.
//
// K = 27 * Cin is too thin for a tcgen05 tile (and the layer is 0.5 % of the network's FLOPs): what matters is that the
// kernel streams its 64 B/voxel bf16 output (forward) resp. input gradient (wgrad) at HBM speed.  The scalar version
// (conv_first.cu) spent its time on 864 shared-memory weight reads per voxel; here the im2col operand is gathered
// straight from an fp32 halo tile in shared memory into mma.sync fragments (8 gathers per 16 voxels x 16 taps), weights
// live in registers as B fragments, fp32 accumulate.  Operands are rounded to bf16 like every other convolution of the
// network (the reference runs this layer in fp16 under autocast, nndet/arch/conv.py:344-348 + ptmodule AMP).
// Input: fp32 NCDHW image batch exactly as the data loader hands it over; weights fp32 [32][Cin][27] (PyTorch layout).
#include "conv_common.cuh"

namespace {

constexpr int TZ = 4, TY = 8, TX = 32;               // forward tile (1024 voxels, 8 warps x 8 m-tiles of 16 voxels)
constexpr int HZ = TZ + 2, HYY = TY + 2, HXX = TX + 2;
constexpr int CO = 32;

__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<unsigned*>(&p);
}

// halo offset of im2col column k = ci * 27 + tap (tap = (dz+1)*9 + (dy+1)*3 + (dx+1)); columns >= 27 * CIN read offset 0
// (finite image values) against zero weights
template <int CIN, int PZ, int PY, int PX>
__device__ __forceinline__ int col_offset(int k) {
  if (k >= 27 * CIN) return 0;
  const int ci = k / 27, tap = k % 27;
  return ((ci * PZ + tap / 9) * PY + (tap / 3) % 3) * PX + tap % 3;
}

template <int CIN>
__global__ void __launch_bounds__(256)
conv_first_fprop_mma_kernel(const float* __restrict__ x, const float* __restrict__ w, const ConvGeom g,
                            __nv_bfloat16* __restrict__ out, float* __restrict__ stat_sum, float* __restrict__ stat_sq) {
  constexpr int KP = (27 * CIN + 15) / 16 * 16, KSTEPS = KP / 16;
  __shared__ float xs[CIN * HZ * HYY * HXX];
  __shared__ __align__(16) __nv_bfloat16 so[8][16][40];       // per-warp output staging (row pitch 80 B: conflict-free)
  __shared__ float s_red[2][CO];
  __shared__ int s_wt[27];                                      // weight index of tap (dz, dy, dx)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int D = g.Ld, H = g.Lh, W = g.Lw;
  const int XB = (W + TX - 1) / TX, YB = (H + TY - 1) / TY;
  int r = blockIdx.x;
  const int xb = r % XB; r /= XB;
  const int yb = r % YB; const int zb = r / YB;
  const int n = blockIdx.y;
  const int z0 = zb * TZ, y0 = yb * TY, x0 = xb * TX;

  if (tid < 27) s_wt[tid] = -1;
  if (tid < CO) { s_red[0][tid] = 0.f; s_red[1][tid] = 0.f; }
  __syncthreads();
  if (tid < g.T) s_wt[(g.off_d[tid] + 1) * 9 + (g.off_h[tid] + 1) * 3 + (g.off_w[tid] + 1)] = g.tap_w[tid];
  const size_t plane = (size_t)D * H * W;
  const float* xn = x + (size_t)n * CIN * plane;
  for (int i = tid; i < CIN * HZ * HYY * HXX; i += 256) {
    const int xx = i % HXX; int r2 = i / HXX;
    const int yy = r2 % HYY; r2 /= HYY;
    const int zz = r2 % HZ; const int ci = r2 / HZ;
    const int z = z0 - 1 + zz, y = y0 - 1 + yy, xg = x0 - 1 + xx;
    float v = 0.f;
    if ((unsigned)z < (unsigned)D && (unsigned)y < (unsigned)H && (unsigned)xg < (unsigned)W)
      v = __ldg(xn + ci * plane + ((size_t)z * H + y) * W + xg);
    xs[i] = v;
  }
  __syncthreads();

  // B fragments (weights) of this thread: column n = j*8 + gq, rows k = 16s + 2tq (+1, +8, +9)
  unsigned bfr[KSTEPS][4][2];
  int koff[KSTEPS][4];
#pragma unroll
  for (int s = 0; s < KSTEPS; ++s) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 16 * s + 2 * tq + (e & 1) + (e >> 1) * 8;
      koff[s][e] = col_offset<CIN, HZ, HYY, HXX>(k);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float wv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 16 * s + 2 * tq + (e & 1) + (e >> 1) * 8;
        float v = 0.f;
        if (k < 27 * CIN) {
          const int wt = s_wt[k % 27];
          if (wt >= 0) v = w[((size_t)(j * 8 + gq) * CIN + k / 27) * g.T + wt];
        }
        wv[e] = v;
      }
      bfr[s][j][0] = pack_bf16(wv[0], wv[1]);
      bfr[s][j][1] = pack_bf16(wv[2], wv[3]);
    }
  }

  float ssum[4][2], ssq[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { ssum[j][0] = ssum[j][1] = 0.f; ssq[j][0] = ssq[j][1] = 0.f; }
  const int wz = warp >> 1, wy0 = (warp & 1) * 4;
#pragma unroll 1
  for (int mt = 0; mt < 8; ++mt) {
    const int yl = wy0 + (mt >> 1), xl = (mt & 1) * 16;
    const int base = (wz * HYY + yl) * HXX + xl + gq;            // halo index of (row gq, tap (-1,-1,-1))
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) {
      const unsigned a0 = pack_bf16(xs[base + koff[s][0]], xs[base + koff[s][1]]);
      const unsigned a1 = pack_bf16(xs[base + 8 + koff[s][0]], xs[base + 8 + koff[s][1]]);
      const unsigned a2 = pack_bf16(xs[base + koff[s][2]], xs[base + koff[s][3]]);
      const unsigned a3 = pack_bf16(xs[base + 8 + koff[s][2]], xs[base + 8 + koff[s][3]]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_bf16_16816(acc[j], a0, a1, a2, a3, bfr[s][j][0], bfr[s][j][1]);
    }
    const int z = z0 + wz, y = y0 + yl;
    const bool row_ok = z < D && y < H;
    const bool ok0 = row_ok && x0 + xl + gq < W, ok1 = row_ok && x0 + xl + gq + 8 < W;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *reinterpret_cast<unsigned*>(&so[warp][gq][j * 8 + 2 * tq]) = pack_bf16(acc[j][0], acc[j][1]);
      *reinterpret_cast<unsigned*>(&so[warp][gq + 8][j * 8 + 2 * tq]) = pack_bf16(acc[j][2], acc[j][3]);
      if (ok0) { ssum[j][0] += acc[j][0]; ssum[j][1] += acc[j][1]; ssq[j][0] = fmaf(acc[j][0], acc[j][0], ssq[j][0]); ssq[j][1] = fmaf(acc[j][1], acc[j][1], ssq[j][1]); }
      if (ok1) { ssum[j][0] += acc[j][2]; ssum[j][1] += acc[j][3]; ssq[j][0] = fmaf(acc[j][2], acc[j][2], ssq[j][0]); ssq[j][1] = fmaf(acc[j][3], acc[j][3], ssq[j][1]); }
    }
    __syncwarp();
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int rowl = (lane >> 2) + h2 * 8, ch = lane & 3;
      if (row_ok && x0 + xl + rowl < W) {
        const uint4 v4 = *reinterpret_cast<const uint4*>(&so[warp][rowl][ch * 8]);
        const size_t vox = (((size_t)n * D + z) * H + y) * W + x0 + xl + rowl;
        *reinterpret_cast<uint4*>(out + vox * CO + ch * 8) = v4;
      }
    }
    __syncwarp();
  }
  if (stat_sum) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float s = ssum[j][e], q2 = ssq[j][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o); }
        if (gq == 0) { atomicAdd(&s_red[0][j * 8 + 2 * tq + e], s); atomicAdd(&s_red[1][j * 8 + 2 * tq + e], q2); }
      }
    __syncthreads();
    if (tid < CO) {
      atomicAdd(&stat_sum[(size_t)n * CO + tid], s_red[0][tid]);
      atomicAdd(&stat_sq[(size_t)n * CO + tid], s_red[1][tid]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dW[co][ci][tap] = sum_v dy[v][co] * x[ci][v + tap]  as  D[co (2 m-tiles)][n = ci*27 + tap] += A[co][v] * B[v][n], K = voxels.
// A comes from a bf16 dy tile in shared memory through ldmatrix.trans, B is gathered from the fp32 halo tile.
constexpr int WZ = 2, WY = 8, WX = 32;                 // wgrad tile: 512 voxels, 8 warps x 4 k-steps of 16 voxels
constexpr int WHZ = WZ + 2, WHY = WY + 2, WHX = WX + 2;

template <int CIN>
__global__ void __launch_bounds__(256, 2)
conv_first_wgrad_mma_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const ConvGeom g,
                            float* __restrict__ dw, int tiles_per_sample, int n_tiles) {
  constexpr int NCOLS = 27 * CIN, NT = (NCOLS + 7) / 8;
  constexpr int DY_BYTES = WZ * WY * WX * 64, XS_FLOATS = CIN * WHZ * WHY * WHX;
  extern __shared__ __align__(128) unsigned char dsm[];
  unsigned char (*sdy)[DY_BYTES] = reinterpret_cast<unsigned char (*)[DY_BYTES]>(dsm);   // 2 x [voxel][32 co] bf16, chunks swizzled
  float (*xs)[XS_FLOATS] = reinterpret_cast<float (*)[XS_FLOATS]>(dsm + 2 * DY_BYTES);      // 2 x fp32 halo tile
  float (*s_acc)[NT * 8] = reinterpret_cast<float (*)[NT * 8]>(dsm + 2 * DY_BYTES + 2 * XS_FLOATS * 4);
  __shared__ int s_wt[27];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int D = g.Ld, H = g.Lh, W = g.Lw;
  const int XB = (W + WX - 1) / WX, YB = (H + WY - 1) / WY;
  const size_t plane = (size_t)D * H * W;

  if (tid < 27) s_wt[tid] = -1;
  for (int i = tid; i < CO * NT * 8; i += 256) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  if (tid < g.T) s_wt[(g.off_d[tid] + 1) * 9 + (g.off_h[tid] + 1) * 3 + (g.off_w[tid] + 1)] = g.tap_w[tid];

  int noff[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) noff[j] = col_offset<CIN, WHZ, WHY, WHX>(j * 8 + gq);

  auto load_tile = [&](int buf, int tile) {
    int r = tile % tiles_per_sample; const int n = tile / tiles_per_sample;
    const int xb = r % XB; r /= XB;
    const int yb = r % YB; const int zb = r / YB;
    const int z0 = zb * WZ, y0 = yb * WY, x0 = xb * WX;
    const unsigned base = smem_u32(&sdy[buf][0]);
    for (int i = tid; i < WZ * WY * WX * 4; i += 256) {
      const int ch = i & 3, v = i >> 2;
      const int xl = v % WX, yl = (v / WX) % WY, zl = v / (WX * WY);
      const int z = z0 + zl, y = y0 + yl, xg = x0 + xl;
      const bool ok = z < D && y < H && xg < W;
      const __nv_bfloat16* src = dy + ((((size_t)n * D + z) * H + y) * W + xg) * CO + ch * 8;
      cp_async16(base + v * 64 + ((ch ^ ((v >> 1) & 3)) * 16), ok ? src : dy, ok);
    }
    const float* xn = x + (size_t)n * CIN * plane;
    for (int i = tid; i < CIN * WHZ * WHY * WHX; i += 256) {
      const int xx = i % WHX; int r2 = i / WHX;
      const int yy = r2 % WHY; r2 /= WHY;
      const int zz = r2 % WHZ; const int ci = r2 / WHZ;
      const int z = z0 - 1 + zz, y = y0 - 1 + yy, xg = x0 - 1 + xx;
      float v = 0.f;
      if ((unsigned)z < (unsigned)D && (unsigned)y < (unsigned)H && (unsigned)xg < (unsigned)W)
        v = __ldg(xn + ci * plane + ((size_t)z * H + y) * W + xg);
      xs[buf][i] = v;
    }
  };

  float acc[2][NT][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f; }

  int buf = 0;
  if ((int)blockIdx.x < n_tiles) load_tile(0, blockIdx.x);
  cp_async_commit();
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < n_tiles) load_tile(buf ^ 1, next);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const unsigned a_st = smem_u32(&sdy[buf][0]);
    const float* xh = xs[buf];
    const int zl = warp >> 2, yl0 = (warp & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int yl = yl0 + (ks >> 1), xl = (ks & 1) * 16;
      const int v0 = (zl * WY + yl) * WX + xl;                      // first voxel (k = 0) of this k-step inside the tile
      unsigned af[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int krow = v0 + (lane & 7) + ((lane >> 4) << 3);
        const int chunk = mi * 2 + ((lane >> 3) & 1);
        ldmatrix_x4_trans(a_st + krow * 64 + ((chunk ^ ((krow >> 1) & 3)) * 16), af[mi][0], af[mi][1], af[mi][2], af[mi][3]);
      }
      const int hb = (zl * WHY + yl) * WHX + xl + 2 * tq;           // halo index of voxel k = 2tq, tap (-1,-1,-1)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float* p = xh + hb + noff[j];
        const unsigned b0 = pack_bf16(p[0], p[1]);
        const unsigned b1 = pack_bf16(p[8], p[9]);
        mma_bf16_16816(acc[0][j], af[0][0], af[0][1], af[0][2], af[0][3], b0, b1);
        mma_bf16_16816(acc[1][j], af[1][0], af[1][1], af[1][2], af[1][3], b0, b1);
      }
    }
    __syncthreads();
    buf ^= 1;
  }
  cp_async_wait<0>();
  // CTA reduction in shared memory, then one atomic per weight
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        atomicAdd(&s_acc[mi * 16 + gq + (e >> 1) * 8][j * 8 + 2 * tq + (e & 1)], acc[mi][j][e]);
  __syncthreads();
  for (int i = tid; i < CO * NCOLS; i += 256) {
    const int co = i / NCOLS, c = i % NCOLS;
    const int wt = s_wt[c % 27];
    if (wt >= 0) atomicAdd(dw + ((size_t)co * CIN + c / 27) * g.T + wt, s_acc[co][c]);
  }
}

bool full_unit_stride(const ConvGeom& g) {
  if (g.sd != 1 || g.sh != 1 || g.sw != 1 || g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return false;
  if (g.Ld != g.Di || g.Lh != g.Hi || g.Lw != g.Wi || g.Do != g.Di || g.Ho != g.Hi || g.Wo != g.Wi) return false;
  if (g.T > 27) return false;
  for (int t = 0; t < g.T; ++t)
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return false;
  return true;
}

}  // namespace

int nnd_conv_first_mma_supported(const ConvGeom& g, int Cout) {
  return Cout == CO && (g.Cin == 1 || g.Cin == 2) && full_unit_stride(g);
}

int nnd_conv_first_fprop_mma(const float* x, const float* w, const ConvGeom& g, __nv_bfloat16* out, float* stat_sum,
                             float* stat_sq, cudaStream_t st) {
  const int tiles = ((g.Ld + TZ - 1) / TZ) * ((g.Lh + TY - 1) / TY) * ((g.Lw + TX - 1) / TX);
  dim3 grid(tiles, g.N);
  if (g.Cin == 1) conv_first_fprop_mma_kernel<1><<<grid, 256, 0, st>>>(x, w, g, out, stat_sum, stat_sq);
  else conv_first_fprop_mma_kernel<2><<<grid, 256, 0, st>>>(x, w, g, out, stat_sum, stat_sq);
  NND_LAUNCH_CHECK("conv_first_fprop_mma_kernel");
  return NND_OK;
}

int nnd_conv_first_wgrad_mma(const float* x, const __nv_bfloat16* dy, const ConvGeom& g, float* dw, cudaStream_t st) {
  const int tps = ((g.Ld + WZ - 1) / WZ) * ((g.Lh + WY - 1) / WY) * ((g.Lw + WX - 1) / WX);
  const long long n_tiles = (long long)tps * g.N;
  if (n_tiles <= 0) return NND_OK;
  if (n_tiles > 0x7fffffffll) return NND_ERR_ARG;
  const int grid = (int)(n_tiles < 2 * NND_NUM_SMS ? n_tiles : 2 * NND_NUM_SMS);
  const int nt = (27 * g.Cin + 7) / 8;
  const size_t smem = (size_t)2 * WZ * WY * WX * 64 + (size_t)2 * g.Cin * WHZ * WHY * WHX * 4 + (size_t)CO * nt * 8 * 4;
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_first_wgrad_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_first_wgrad_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  if (g.Cin == 1) conv_first_wgrad_mma_kernel<1><<<grid, 256, smem, st>>>(x, dy, g, dw, tps, (int)n_tiles);
  else conv_first_wgrad_mma_kernel<2><<<grid, 256, smem, st>>>(x, dy, g, dw, tps, (int)n_tiles);
  NND_LAUNCH_CHECK("conv_first_wgrad_mma_kernel");
  return NND_OK;
}
