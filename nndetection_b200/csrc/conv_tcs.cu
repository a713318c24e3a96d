// Streaming 3x3x3 stride-1 convolution for the 32- and 64-channel full-resolution layers (tcgen05, "z-window" form).
//
// Why a second tcgen05 kernel: with both operands in shared memory one 128 x N x 16 MMA costs max(N/2, 32 + N/4) cycles
// (scripts/mma_rate.cu, measured on B200: the A tile is re-read at 128 B/clk for every MMA), so the N = 32 MMAs of
// conv_tc.cu cannot exceed 40 % of the tensor peak.  Here the three dz taps of a filter column are stacked along N:
//     input slice s, tap (dy, dx):   D[:, out slices s-1 | s | s+1] += A_s(dy, dx) * [ W(dz=+1) | W(dz=0) | W(dz=-1) ]
// i.e. ONE N = 96 MMA per (dy, dx, 16 channels) accumulates into the three neighbouring output-slice accumulators, which
// sit side by side in a TMEM ring of 16 slots x 32 columns.  9 x Cin/16 MMAs per input slice instead of 27 x Cin/16 per
// output slice, each 2.1x cheaper per column.  The kernel streams along z: input slices pass once through a shared-
// memory ring (halo 18 x 10 voxels x Cin), the whole weight tensor of the 32-channel output tile stays resident in
// shared memory, output slices leave the TMEM ring as soon as their third input slice has been multiplied.
// Replaces cuDNN implicit GEMM behind torch.nn.Conv3d for these layers (nndet/arch/conv.py:344-348); fprop and, with
// flipped taps + transposed weights, dgrad.
//
// Roles: 4 producer warps (cp.async gathers, zero fill = padding) | NI issuer warps (elected lane issues tcgen05.mma;
// all MMAs accumulate -- slots are zeroed by the epilogue -- so the issue order between issuers does not matter) |
// 4 epilogue warps (tcgen05.ld -> zero the slot (tcgen05.st) -> release it -> bias / residual / scale -> bf16 ->
// 64-byte row stores, norm statistics).  Persistent grid, static round-robin over (output tile, sample, z segment, h, w).
#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int BH = 16, BW = 8;
constexpr int HY = BH + 2, HX = BW + 2;
constexpr int ROW_PITCH = HX * 16;               // 160 B
constexpr int GP = HY * ROW_PITCH;               // 2880 B: one 8-channel group of one halo slice
constexpr int NB = 32;                           // output channels per pass
constexpr int ACC_SLOTS = 16;                    // TMEM ring: 16 x 32 columns
constexpr int NPROD = 128;
constexpr int WROW = 3 * NB * 16;                // 1536 B: one k-group of a (dy, dx) weight panel [96 rows][8 ch]



struct TcsTiles {
  int ZS, NZ, HB, WB, NT;      // z-segment length, segments per column, tiles along h / w, 32-channel output tiles
  int per_nt;                  // work items per output tile = N * NZ * HB * WB
  int total;                   // per_nt * NT
  int map_mode;                // halo copy lane mapping: 0 one thread per (group, y) row, 1 lanes along the channel groups of a voxel
};

struct Item { int n, z0, z1, h0, w0, nt; };

__device__ __forceinline__ Item decode_item(int idx, const TcsTiles& tl, int D) {
  Item it;
  it.nt = idx / tl.per_nt; int r = idx - it.nt * tl.per_nt;
  const int wb = r % tl.WB; r /= tl.WB;
  const int hb = r % tl.HB; r /= tl.HB;
  const int zs = r % tl.NZ; it.n = r / tl.NZ;
  it.z0 = zs * tl.ZS; it.z1 = min(it.z0 + tl.ZS, D);
  it.h0 = hb * BH; it.w0 = wb * BW;
  return it;
}

// A_SLOTS input slices in the ring, of which LAG + 1 may still be in flight (published LAG slices late)
// (scripts/mma_rate.cu modes 2 / 3: the 160-byte row pitch of the halo and the 16-byte dx shifts, which put most 8 x 16 B core
// matrices across two 128-byte shared-memory lines, cost nothing per MMA -- 56.0 cycles at N = 96 either way -- so there is
// no aligned-copy variant of the halo.)
template <int CIN, int NI, int A_SLOTS, int LAG, bool STATS>
__global__ void __launch_bounds__((4 + NI + 4) * 32, 1)
conv_tcs_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ wgt, const ConvGeom g,
                const ConvEpilogue ep, const TcsTiles tl) {
  constexpr int KG = CIN / 8;                      // 8-channel groups
  constexpr int KS = CIN / 16;                     // MMA k-steps per tap
  constexpr int SLOT_BYTES = KG * GP;              // one input slice with halo
  constexpr int W_BYTES = 9 * KG * WROW;           // resident weights of one 32-channel output tile
  constexpr int NQ = 9 * KS;                       // MMAs per input slice (full window)
  // K-major no-swizzle descriptors: lo = start >> 4 | (LBO >> 4) << 16, hi = SBO >> 4 | version 1 << 14
  constexpr unsigned A_LO_HI = (unsigned)((GP >> 4) & 0x3FFF) << 16, A_HI = (unsigned)((ROW_PITCH >> 4) & 0x3FFF) | (1u << 14);
  constexpr unsigned B_LO_HI = (unsigned)((WROW >> 4) & 0x3FFF) << 16, B_HI = (unsigned)((128 >> 4) & 0x3FFF) | (1u << 14);

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sW = smem;
  unsigned char* sA = smem + W_BYTES;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sA + A_SLOTS * SLOT_BYTES);
  __shared__ unsigned s_tmem_base;
  __shared__ int s_wtap[27];                       // weight slice of the tap with offsets (dz, dy, dx), index (dz+1)*9+(dy+1)*3+(dx+1)
  __shared__ float s_stat[2][4][NB];
  __shared__ __align__(16) float s_bias[NB];
  const unsigned bar0 = smem_u32(bars);
  auto HFULL = [&](int i) { return bar0 + 8u * i; };
  auto HEMPTY = [&](int i) { return bar0 + 8u * (A_SLOTS + i); };
  auto AFULL = [&](int i) { return bar0 + 8u * (2 * A_SLOTS + i); };
  auto AEMPTY = [&](int i) { return bar0 + 8u * (2 * A_SLOTS + ACC_SLOTS + i); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = g.Ld;

  if (tid == 0) {
    for (int i = 0; i < A_SLOTS; ++i) { mbar_init(HFULL(i), 4); mbar_init(HEMPTY(i), NI); }
    for (int i = 0; i < ACC_SLOTS; ++i) { mbar_init(AFULL(i), NI); mbar_init(AEMPTY(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < g.T) s_wtap[(g.off_d[tid] + 1) * 9 + (g.off_h[tid] + 1) * 3 + (g.off_w[tid] + 1)] = g.tap_w[tid];
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = s_tmem_base;
  if (warp >= 4 + NI) {                            // all accumulator slots start at zero: every MMA accumulates
    const int q = warp & 3;
    for (int s = 0; s < ACC_SLOTS; ++s) tmem_zero32(tmem_base + ((unsigned)(q * 32) << 16) + s * NB);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < 4) {
    // ================================================================ producers
    unsigned slot = 0, phase = 0, done_slot = 0;
    int pending = 0, cur_nt = -1;
    bool issued_any = false;
    unsigned last_slot = 0, last_phase = 0;
    auto publish_all = [&]() {
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      while (pending > 0) {
        if (lane == 0) mbar_arrive(HFULL(done_slot));
        done_slot = (done_slot + 1 == A_SLOTS) ? 0 : done_slot + 1;
        --pending;
      }
    };
    for (int idx = blockIdx.x; idx < tl.total; idx += gridDim.x) {
      const Item it = decode_item(idx, tl, D);
      if (it.nt != cur_nt) {
        // New output tile: the resident weights change.  Everything issued so far must have been consumed.
        if (issued_any) { publish_all(); mbar_wait_warp(HEMPTY(last_slot), last_phase, lane); }
        const unsigned w_base = smem_u32(sW);
        const __nv_bfloat16* wsrc0 = wgt + (size_t)(it.nt * NB) * g.Cin;
        for (int i = tid; i < 9 * KG * 3 * NB; i += NPROD) {
          const int row = i % (3 * NB); const int r2 = i / (3 * NB);
          const int kg = r2 % KG, p = r2 / KG;                  // p = (dy+1)*3 + (dx+1)
          const int j = row / NB, co = row - j * NB;            // column block j <-> dz = 1 - j
          const int tw = s_wtap[(2 - j) * 9 + p];
          cp_async16(w_base + (p * KG + kg) * WROW + row * 16, wsrc0 + (size_t)tw * ep.CoutPad * g.Cin + (size_t)co * g.Cin + kg * 8, true);
        }
        cur_nt = it.nt;                                          // lands with the next slice's commit group
      }
      const int s_lo = max(it.z0 - 1, 0), s_hi = min(it.z1, D - 1);
      const __nv_bfloat16* in_n = in + (size_t)it.n * g.Di * g.Hi * g.Wi * g.Cin;
      for (int s = s_lo; s <= s_hi; ++s) {
        mbar_wait_warp(HEMPTY(slot), phase ^ 1, lane);
        const unsigned a_base = smem_u32(sA + slot * SLOT_BYTES);
        if (tl.map_mode == 1) {
          // lanes = the KG consecutive 16-byte channel groups of one voxel, then the next voxel of the row: a warp instruction reads
          // ~512 contiguous bytes (4-5 lines) instead of 32 different lines -- the LSU processes one line per cycle, and the round-1
          // mapping (one thread per (group, y) row walking x) spent ~720 of the ~1000 cycles an input slice's MMAs take on them
          const __nv_bfloat16* slice = in_n + (long long)s * g.Hi * g.Wi * g.Cin;
          for (int c = tid; c < HY * HX * KG; c += NPROD) {
            const int kg = c % KG; const int t = c / KG;
            const int x = t % HX, y = t / HX;
            const int h = it.h0 - 1 + y, w = it.w0 - 1 + x;
            const bool ok = (unsigned)h < (unsigned)g.Hi && (unsigned)w < (unsigned)g.Wi;
            cp_async16(a_base + kg * GP + y * ROW_PITCH + x * 16, ok ? slice + ((long long)h * g.Wi + w) * g.Cin + kg * 8 : in, ok);
          }
        } else {
        for (int rr = tid; rr < KG * HY; rr += NPROD) {
          const int kg = rr / HY, y = rr - kg * HY;
          const int h = it.h0 - 1 + y;
          const bool row_ok = (unsigned)h < (unsigned)g.Hi;
          const __nv_bfloat16* src = in_n + ((long long)(s * g.Hi + h) * g.Wi + (it.w0 - 1)) * g.Cin + kg * 8;
          unsigned dst = a_base + kg * GP + y * ROW_PITCH;
#pragma unroll
          for (int x = 0; x < HX; ++x) {
            const bool ok = row_ok && (unsigned)(it.w0 - 1 + x) < (unsigned)g.Wi;
            cp_async16(dst, ok ? src : in, ok);
            dst += 16; src += g.Cin;
          }
        }
        }
        cp_async_commit();
        ++pending;
        issued_any = true; last_slot = slot; last_phase = phase;
        if (pending > LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(HFULL(done_slot));
          done_slot = (done_slot + 1 == A_SLOTS) ? 0 : done_slot + 1;
          --pending;
        }
        if (++slot == A_SLOTS) { slot = 0; phase ^= 1; }
      }
    }
    publish_all();
  } else if (warp < 4 + NI) {
    // ================================================================ MMA issuers (whole warp runs the control flow,
    // the elected lane issues): issuer i takes the (tap, k-step) pairs q with q % NI == i of every input slice.
    const int me = __shfl_sync(0xffffffffu, warp - 4, 0);
    const unsigned tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    unsigned hs = 0, hphase = 0;
    int g_base = 0, acquired = 0;
    const unsigned a0 = smem_u32(sA), w0s = smem_u32(sW);
    for (int idx = blockIdx.x; idx < tl.total; idx += gridDim.x) {
      const Item it = decode_item(idx, tl, D);
      const int s_lo = max(it.z0 - 1, 0), s_hi = min(it.z1, D - 1);
      int completed_next = it.z0;
      for (int s = s_lo; s <= s_hi; ++s) {
        const int mlo = max(s - 1, it.z0), mhi = min(s + 1, it.z1 - 1);
        const int glo = g_base + (mlo - it.z0), ghi = g_base + (mhi - it.z0);
        while (acquired <= ghi) {
          mbar_wait_warp(AEMPTY(acquired % ACC_SLOTS), ((unsigned)(acquired / ACC_SLOTS) & 1u) ^ 1u, lane);
          ++acquired;
        }
        mbar_wait_warp(HFULL(hs), hphase, lane);
        tc_fence_after();
        const int jlo = mlo - (s - 1), cnt = mhi - mlo + 1;
        const int c0 = glo % ACC_SLOTS;
        const int n1 = min(cnt, ACC_SLOTS - c0), n2 = cnt - n1;          // the window may wrap around the TMEM ring
        const unsigned idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)((n1 * NB) >> 3) << 17) | ((128u >> 4) << 24);
        const unsigned idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)((n2 * NB) >> 3) << 17) | ((128u >> 4) << 24);
        const unsigned a_lo0 = ((a0 + hs * SLOT_BYTES) >> 4) & 0x3FFF, b_lo0 = ((w0s + jlo * NB * 16) >> 4) & 0x3FFF;
        const unsigned d1 = tm + c0 * NB, d2 = tm;
        const int mdone_hi = (s == s_hi) ? it.z1 - 1 : s - 1;
        if (elect_one()) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            if (q % NI != me) continue;
            const int p = q / KS, ks = q % KS;
            const int dy = p / 3, dx = p % 3;                               // halo offsets (dy, dx) in 0..2
            const unsigned a_lo = (a_lo0 + ((ks * 2 * GP + (dy * HX + dx) * 16) >> 4)) | A_LO_HI;
            const unsigned b_lo = (b_lo0 + (((p * KG + ks * 2) * WROW) >> 4)) | B_LO_HI;
            tc_mma_acc2(d1, a_lo, A_HI, b_lo, B_HI, idesc1);
            if (n2 > 0) tc_mma_acc2(d2, a_lo, A_HI, b_lo + ((n1 * NB * 16) >> 4), B_HI, idesc2);
          }
          tc_commit(HEMPTY(hs));
          for (int m = completed_next; m <= mdone_hi; ++m) tc_commit(AFULL((g_base + m - it.z0) % ACC_SLOTS));
        }
        __syncwarp();
        if (mdone_hi >= completed_next) completed_next = mdone_hi + 1;
        if (++hs == A_SLOTS) { hs = 0; hphase ^= 1; }
      }
      g_base += it.z1 - it.z0;
    }
  } else {
    // ================================================================ epilogue (warp q owns TMEM lanes 32q .. 32q+31)
    // Per output slice a warp only does: TMEM load, zero + release the slot, (bias / residual / scale), bf16 pack, 64-byte
    // row store, and 64 FMAs into per-thread statistics; the cross-lane reduction of the statistics happens once per
    // work item (the epilogue warps are the critical path of this kernel: ~150 instead of ~700 instructions per slice).
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int hy = row >> 3, wx = row & 7;
    const float scale = ep.scale ? *ep.scale : 1.f;
    const bool has_scale = ep.scale != nullptr;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(ep.out);
    int g_base = 0;
    for (int idx = blockIdx.x; idx < tl.total; idx += gridDim.x) {
      const Item it = decode_item(idx, tl, D);
      const int h = it.h0 + hy, w = it.w0 + wx;
      const bool hw_ok = h < g.Lh && w < g.Lw;
      const int co0 = it.nt * NB;
      if (ep.bias) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - (4 + NI) * 32;
        if (et < NB) s_bias[et] = ep.bias[co0 + et];
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      float ssum[STATS ? 32 : 1], ssq[STATS ? 32 : 1];
      if (STATS) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
      }
      long long vox = ((long long)(it.n * g.Do + it.z0) * g.Ho + h) * g.Wo + w;
      const long long vox_step = (long long)g.Ho * g.Wo;
#pragma unroll 1
      for (int m = it.z0; m < it.z1; ++m, vox += vox_step) {
        const int gi = g_base + (m - it.z0);
        const int slot = gi % ACC_SLOTS;
        mbar_wait_warp(AFULL(slot), (unsigned)(gi / ACC_SLOTS) & 1u, lane);
        tc_fence_after();
        unsigned v[32];
        const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + slot * NB;
        tmem_ld32(taddr, v);
        tmem_zero32(taddr);                        // the next output slice using this slot accumulates onto zeros
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(AEMPTY(slot));
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[j]);
            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
          }
        }
        if (hw_ok) {
          if (ep.residual) {
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
          if (has_scale) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= scale;
          }
          __align__(16) __nv_bfloat162 pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          uint4* op = reinterpret_cast<uint4*>(outp + vox * ep.Cout + co0);
          const uint4* sp = reinterpret_cast<const uint4*>(pk);
#pragma unroll
          for (int u = 0; u < 4; ++u) op[u] = sp[u];
          if (STATS) {                              // statistics of the fp32 values (the stored bf16 differs by < 2^-9 relative)
#pragma unroll
            for (int j = 0; j < 32; ++j) { ssum[j] += f[j]; ssq[j] = fmaf(f[j], f[j], ssq[j]); }
          }
        }
      }
      g_base += it.z1 - it.z0;
      if (STATS) {
        int col;
        const float cs = warp_transpose_reduce32(ssum, lane, col);
        const float cq = warp_transpose_reduce32(ssq, lane, col);
        s_stat[0][q][col] = cs;                    // col is a permutation of the lanes
        s_stat[1][q][col] = cq;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - (4 + NI) * 32;          // 0..127
        if (et < NB) {
          const float s = s_stat[0][0][et] + s_stat[0][1][et] + s_stat[0][2][et] + s_stat[0][3][et];
          const float qq = s_stat[1][0][et] + s_stat[1][1][et] + s_stat[1][2][et] + s_stat[1][3][et];
          atomicAdd(&ep.stat_sum[(size_t)it.n * ep.Cout + co0 + et], s);
          atomicAdd(&ep.stat_sq[(size_t)it.n * ep.Cout + co0 + et], qq);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
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

int g_tcs_issuers = 2;
int g_tcs_map = 0;             // halo copy lane mapping (TcsTiles::map_mode); nnd_conv_set_tcs_map

template <int CIN, int NI, int A_SLOTS, int LAG, bool STATS>
int launch_tcs(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st) {
  TcsTiles tl;
  tl.HB = (g.Lh + BH - 1) / BH; tl.WB = (g.Lw + BW - 1) / BW; tl.NT = ep.Cout / NB;
  // z-segment length: minimise rounds x (segment cost + 2 edge slices) over the persistent grid
  const long long cols = (long long)g.N * tl.HB * tl.WB * tl.NT;
  int best_zs = g.Ld; double best = 1e30;
  for (int nz = 1; nz <= g.Ld; ++nz) {
    const int zs = (g.Ld + nz - 1) / nz;
    if (zs < 4 && nz > 1) break;
    const long long items = cols * ((g.Ld + zs - 1) / zs);
    const long long rounds = (items + NND_NUM_SMS - 1) / NND_NUM_SMS;
    const double cost = (double)rounds * (zs * 56.0 + 2 * 40.0 + 8.0);
    if (cost < best) { best = cost; best_zs = zs; }
  }
  tl.ZS = best_zs; tl.NZ = (g.Ld + tl.ZS - 1) / tl.ZS;
  tl.per_nt = g.N * tl.NZ * tl.HB * tl.WB;
  tl.total = tl.per_nt * tl.NT;
  tl.map_mode = g_tcs_map;
  constexpr size_t SMEM = (size_t)9 * (CIN / 8) * WROW + (size_t)A_SLOTS * (CIN / 8) * GP + 8 * (2 * A_SLOTS + 2 * ACC_SLOTS);
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_tcs_kernel<CIN, NI, A_SLOTS, LAG, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  const int grid = tl.total < NND_NUM_SMS ? tl.total : NND_NUM_SMS;
  conv_tcs_kernel<CIN, NI, A_SLOTS, LAG, STATS><<<grid, (4 + NI + 4) * 32, SMEM, st>>>(in, w, g, ep, tl);
  NND_LAUNCH_CHECK("conv_tcs_kernel");
  return NND_OK;
}

}  // namespace

void nnd_conv_tcs_set_issuers(int n) { g_tcs_issuers = n == 1 ? 1 : 2; }
// halo copy lane mapping of the streaming kernel: 0 = one thread per (channel group, y) row, 1 = lanes along the channel groups of a voxel
extern "C" void nnd_conv_set_tcs_map(int mode) { g_tcs_map = mode ? 1 : 0; }

int nnd_conv_tcs_supported(const ConvGeom& g, const ConvEpilogue& ep) {
  if (g.sd != 1 || g.sh != 1 || g.sw != 1) return 0;
  if (g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return 0;
  if (g.Ld != g.Do || g.Lh != g.Ho || g.Lw != g.Wo || g.Ld != g.Di || g.Lh != g.Hi || g.Lw != g.Wi) return 0;
  if (g.T != 27 || (g.Cin != 32 && g.Cin != 64)) return 0;
  unsigned seen = 0;
  for (int t = 0; t < 27; ++t) {
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
    seen |= 1u << ((g.off_d[t] + 1) * 9 + (g.off_h[t] + 1) * 3 + (g.off_w[t] + 1));
  }
  if (seen != (1u << 27) - 1) return 0;
  if (ep.out_fp32 || ep.Cout % 32 || ep.CoutPad != ep.Cout || ep.Cout > 128) return 0;
  if (ep.out_v_stride != ep.Cout || ep.out_n_stride != (long long)g.Do * g.Ho * g.Wo * ep.Cout) return 0;
  if (g.Lh < 8 || g.Lw < 8 || g.Ld < 2) return 0;
  return 1;
}

// the streaming form pays off when the volume feeds the persistent grid; tiny volumes stay on the tile kernel
int nnd_conv_tcs_profitable(const ConvGeom& g, const ConvEpilogue& ep) {
  const long long cols = (long long)g.N * ((g.Lh + BH - 1) / BH) * ((g.Lw + BW - 1) / BW) * (ep.Cout / 32);
  return cols * ((g.Ld + 3) / 4) >= NND_NUM_SMS;
}

int nnd_conv_tcs(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st) {
  if (!nnd_conv_tcs_supported(g, ep)) return NND_ERR_ARG;
  const bool st2 = ep.stat_sum != nullptr;
  if (g.Cin == 32) {
    if (g_tcs_issuers == 1) return st2 ? launch_tcs<32, 1, 12, 5, true>(in, w, g, ep, st) : launch_tcs<32, 1, 12, 5, false>(in, w, g, ep, st);
    return st2 ? launch_tcs<32, 2, 12, 5, true>(in, w, g, ep, st) : launch_tcs<32, 2, 12, 5, false>(in, w, g, ep, st);
  }
  if (g_tcs_issuers == 1) return st2 ? launch_tcs<64, 1, 4, 2, true>(in, w, g, ep, st) : launch_tcs<64, 1, 4, 2, false>(in, w, g, ep, st);
  return st2 ? launch_tcs<64, 2, 4, 2, true>(in, w, g, ep, st) : launch_tcs<64, 2, 4, 2, false>(in, w, g, ep, st);
}
