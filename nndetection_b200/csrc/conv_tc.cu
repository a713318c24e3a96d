// 3x3x3 (or any taps with offsets in [-1, 1]) stride-1 gather convolution on the 5th-generation tensor cores.
//
// This is the im2col-free kernel for the layers that carry ~90 % of the network's convolution FLOPs (fprop of every
// stride-1 3x3x3 conv and, with flipped taps + transposed weights, their dgrad).  It replaces cuDNN implicit GEMM as
// invoked by torch.nn.Conv3d in the reference (nndet/arch/conv.py:344-348).
//
// Data movement (per CTA tile = MT depth slices x 16 x 8 voxels x N_TILE output channels):
//   * the input HALO ((MT+2) x 18 x 10 voxels x 32 channels) is staged ONCE per 32-channel chunk and serves all 27
//     taps: smem layout [channel group of 8][z][y][x][8 ch] -- one voxel = one 16-byte core-matrix row, so the UMMA
//     shared-memory descriptor of tap (dz,dy,dx) is just the halo base + ((dz+1)*18 + (dy+1))*10 + (dx+1) rows:
//     K-major, no swizzle, SBO = 160 B (halo row pitch), LBO = channel-group pitch.  No im2col, no per-tap reload.
//   * the weight slices of TG taps of one 32-channel chunk (TG x [N_TILE x 32]) stream through a ring of pipeline items in
//     the canonical K-major no-swizzle layout [k group][n][8] (SBO 128 B, LBO N_TILE*16 B) and are shared by the MT
//     slice-MMAs; three taps per item (24 KB at N_TILE = 128) amortise the producers' per-item bookkeeping.
//   * accumulators live in TMEM: MT x N_TILE fp32 columns, double buffered so the epilogue of tile i overlaps the
//     MMAs of tile i+1.
// Roles: warps 0-3 producers (cp.async 16-byte gathers with zero fill = padding; completion ->
// fence.proxy.async -> mbarrier), MT issuer warps (lane 0 issues tcgen05.mma / tcgen05.commit for its depth slice), 4 epilogue warps
// (tcgen05.ld -> bias / residual / scale -> bf16 -> 64-byte row stores, per-(sample, channel) norm statistics via a
// warp transpose-reduce).  Persistent grid = #SMs, static round-robin tiles.
//
// S2 = 1 (opt-in until validated on the device, nnd_conv_set_gather_strided_tc): the same kernel for stride-2 gathers (3x3x3
// stride-2 convolutions, and the 2x2x2 stride-2 convolution that is the dgrad of an up-convolution).  Input position = 2 * lo +
// off with off in [-1, 1]; the halo of a tile is staged DE-INTERLEAVED per axis -- "odd" plane (positions 2p - 1, p = 0..n) in slots
// 0..n, "even" plane (2p, p = 0..n-1) in slots n+1..2n -- so that the 8 (w) x 16 (h) x MT (d) voxels a tap multiplies are again
// consecutive slots: off = -1 -> slot 0, off = 0 -> slot n + 1, off = +1 -> slot 1 of that axis.  16 channels per chunk
// (the halo is 8x the tile instead of 1.7x), one K = 16 MMA per tap and depth slice.
//
// BULK = 1 (opt-in until validated on the device, nnd_conv_set_tc_bulk): the weight stream through the bulk-copy engine.  At
// ~700 TFLOP/s the 128-channel layers move only ~6 B/clk/SM of operands: they are bound by the producers' per-item work (hundreds of
// 16-byte cp.async + wait_group / fence / arrive bookkeeping per 8 KB slice), not by bandwidth.  With the weights re-packed in item
// order ONE thread posts `mbarrier.arrive.expect_tx` and TG `cp.async.bulk` copies per item (the barrier completes on the bytes,
// written through the async proxy the MMA reads through: no fence); the halo gets its own FULLA barrier (the producers block on
// `cp.async.wait_group 0` right after issuing it -- they have nothing else to do any more).
#include "conv_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int BH = 16, BW = 8;                   // tile = MT x 16 x 8 voxels (one 128-row MMA per depth slice)
constexpr int HY = BH + 2, HX = BW + 2;          // halo extents in h, w
constexpr int ROW_PITCH = HX * 16;               // 160 B
constexpr int SLICE_PITCH = HY * ROW_PITCH;      // 2880 B
constexpr int KG = 4;                            // channel groups (of 8) per stage = 32 channels
constexpr int A_STAGES = 2;
constexpr int NUM_PROD = 128;
constexpr int tc_threads(int mt) { return (4 + mt + 4) * 32; }   // 4 producer + MT issuer + 4 epilogue warps


struct TcTiles {
  int DB, HB, WB, NT;       // tile counts along d, h, w and output-channel tiles
  int total;
  // BULK variant only: the weights again, re-packed in pipeline-item order [co tile][ci chunk][weight slice][k group][n][8]
  // (one tap slice of one chunk = B_TAP_BYTES contiguous bytes) and the number of weight slices of the pack
  const __nv_bfloat16* items;
  int items_T;
};


// TG = taps per pipeline item (weight slices loaded / consumed together), ACC = TMEM accumulator stages
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int N_TILE, int MT, int TG, int ACC, int B_SLOTS, int S2, int BULK>
__global__ void __launch_bounds__(tc_threads(MT), 1)
conv_tc_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ wgt, const ConvGeom g,
               const ConvEpilogue ep, const TcTiles tl) {
  // stride-1 names of the namespace constants are shadowed by their per-variant values
  constexpr int KG = S2 ? 2 : ::KG;                          // channel groups of 8 per chunk
  constexpr int HY = S2 ? 2 * BH + 1 : ::HY, HX = S2 ? 2 * BW + 1 : ::HX;
  constexpr int ROW_PITCH = HX * 16, SLICE_PITCH = HY * ROW_PITCH;
  constexpr int HALO_SLICES = S2 ? 2 * MT + 1 : MT + 2;
  constexpr int HV = HALO_SLICES * HY * HX;                 // halo voxels per channel group
  constexpr int A_BYTES = KG * HV * 16;
  constexpr int B_TAP_BYTES = N_TILE * KG * 16;            // one tap slice [kg][n][8]
  constexpr int B_BYTES = TG * B_TAP_BYTES;                 // one pipeline item
  constexpr int LAG = B_SLOTS / 2;
  constexpr unsigned LBO_A = HV * 16, SBO_A = ROW_PITCH;
  constexpr unsigned LBO_B = N_TILE * 16, SBO_B = 128;
  constexpr int ACC_COLS = MT * N_TILE;                     // one accumulator stage
  constexpr int TMEM_COLS = ACC * ACC_COLS >= 512 ? 512 : (ACC * ACC_COLS >= 256 ? 256 : (ACC * ACC_COLS >= 128 ? 128 : 64));
  static_assert(ACC * ACC_COLS <= 512, "TMEM overflow");
  constexpr unsigned IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N_TILE >> 3) << 17) | ((128u >> 4) << 24);

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                                  // A_STAGES x A_BYTES
  unsigned char* sB = smem + A_STAGES * A_BYTES;             // B_SLOTS x B_BYTES
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sB + B_SLOTS * B_BYTES);
  // barrier map: [0,B_SLOTS) fullB | [B_SLOTS, 2B) emptyB | +A_STAGES emptyA | +2 tmem_full | +2 tmem_empty
  __shared__ unsigned s_tmem_base;
  __shared__ int s_tapoff[NND_MAX_TAPS];
  __shared__ float s_stat[2][4][N_TILE];
  const unsigned bar0 = smem_u32(bars);
  auto FULLB = [&](int i) { return bar0 + 8u * i; };
  auto EMPTYB = [&](int i) { return bar0 + 8u * (B_SLOTS + i); };
  auto EMPTYA = [&](int i) { return bar0 + 8u * (2 * B_SLOTS + i); };
  auto TFULL = [&](int i) { return bar0 + 8u * (2 * B_SLOTS + A_STAGES + i); };
  auto TEMPTY = [&](int i) { return bar0 + 8u * (2 * B_SLOTS + A_STAGES + 2 + i); };
  auto FULLA = [&](int i) { return bar0 + 8u * (2 * B_SLOTS + A_STAGES + 4 + i); };      // BULK only

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = g.Cin / (8 * KG), T = g.T;

  if (tid == 0) {
    for (int i = 0; i < B_SLOTS; ++i) { mbar_init(FULLB(i), BULK ? 1 : NUM_PROD / 32); mbar_init(EMPTYB(i), MT); }
    for (int i = 0; i < A_STAGES; ++i) mbar_init(EMPTYA(i), MT);
    if (BULK) for (int i = 0; i < A_STAGES; ++i) mbar_init(FULLA(i), NUM_PROD / 32);
    for (int i = 0; i < 2; ++i) { mbar_init(TFULL(i), MT); mbar_init(TEMPTY(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (S2) {
    // slot of output 0 on an axis with n outputs per tile: odd plane first (n + 1 slots), then the even plane
    auto sl = [](int off, int n) { return off == 0 ? n + 1 : (off > 0 ? 1 : 0); };
    // the depth axis may be unstrided (first stride (1, 2, 2) of anisotropic plans): ordinary halo slots d0 - 1 + z there
    if (tid < T) s_tapoff[tid] = (((g.sd == 2 ? sl(g.off_d[tid], MT) : g.off_d[tid] + 1) * HY + sl(g.off_h[tid], BH)) * HX + sl(g.off_w[tid], BW)) * 16;
  } else {
    if (tid < T) s_tapoff[tid] = (((g.off_d[tid] + 1) * HY + (g.off_h[tid] + 1)) * HX + (g.off_w[tid] + 1)) * 16;
  }
  if (warp == 4) {
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
    d0 = db * MT; h0 = hb * BH; w0 = wb * BW;
  };

  if (warp < 4) {
    // ================================================================ producers
    // Work is a flat sequence of chunks = (tile, 32-channel block); each chunk has T items (taps).  Item t of a chunk
    // carries the weight slice of tap t; the HALO of the NEXT chunk is prefetched in the middle of the current one
    // (item T/2) so it has landed long before its first tap is due.  Halo loads are issued row-wise: one thread owns
    // a (channel group, z, y) row of HX voxels = HX consecutive 16-byte chunks in smem, constant stride in HBM.
    unsigned slot = 0, slot_phase = 0;            // B ring cursor (issue side)
    unsigned astage = 0, a_phase = 0;             // A double buffer: stage/phase of the NEXT halo to be loaded
    unsigned done_slot = 0;                       // completion cursor (LAG items behind)
    int pending = 0;
    const int my_tiles = (tl.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_chunks = my_tiles * KC;
    constexpr int HROWS = KG * HALO_SLICES * HY;
    const int NI_ITEMS = T / TG;

    auto load_halo = [&](int chunk) {
      const int tile = blockIdx.x + (chunk / KC) * gridDim.x;
      const int kc = chunk % KC;
      int n, d0, h0, w0, nt;
      decode_tile(tile, n, d0, h0, w0, nt);
      const bool a_free = BULK ? false : __shfl_sync(0xffffffffu, lane == 0 ? (int)mbar_test(EMPTYA(astage), a_phase ^ 1) : 0, 0) != 0;
      if (BULK) {
        mbar_wait_warp(EMPTYA(astage), a_phase ^ 1, lane);      // no arrivals are ever owed in this variant: just wait
      } else if (!a_free) {
        // The halo buffer is still being read.  Its release needs the MMAs of the previous chunk, which may be
        // waiting for weight items whose (deferred) arrival this warp still owes: publish them before blocking.
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        while (pending > 0) {
          if (lane == 0) mbar_arrive(FULLB(done_slot));
          done_slot = (done_slot + 1 == B_SLOTS) ? 0 : done_slot + 1;
          --pending;
        }
        mbar_wait_warp(EMPTYA(astage), a_phase ^ 1, lane);
      }
      const unsigned a_base = smem_u32(sA + astage * A_BYTES);
      const __nv_bfloat16* in_n = in + (size_t)n * g.Di * g.Hi * g.Wi * g.Cin + kc * (8 * KG);
      for (int rr = tid; rr < HROWS; rr += NUM_PROD) {
        const int gidx = rr / (HALO_SLICES * HY); const int r2 = rr - gidx * (HALO_SLICES * HY);
        const int z = r2 / HY, y = r2 - z * HY;
        if (!S2) {
        const int d = d0 - 1 + z, h = h0 - 1 + y;
        const bool row_ok = (unsigned)d < (unsigned)g.Di && (unsigned)h < (unsigned)g.Hi;
        const __nv_bfloat16* src = in_n + ((long long)(d * g.Hi + h) * g.Wi + (w0 - 1)) * g.Cin + gidx * 8;
        unsigned dst = a_base + rr * (HX * 16);
#pragma unroll
        for (int x = 0; x < HX; ++x) {
          const bool ok = row_ok && (unsigned)(w0 - 1 + x) < (unsigned)g.Wi;
          cp_async16(dst, ok ? src : in, ok);
          dst += 16; src += g.Cin;
        }
        } else {
          // slot -> input position: odd plane slot p -> 2 (o0 + p) - 1, even plane slot n + 1 + p -> 2 (o0 + p)
          const int d = g.sd == 2 ? (z <= MT ? 2 * (d0 + z) - 1 : 2 * (d0 + z - (MT + 1))) : d0 - 1 + z;
          const int h = y <= BH ? 2 * (h0 + y) - 1 : 2 * (h0 + y - (BH + 1));
          const bool row_ok = (g.sd == 2 || z < MT + 2) && (unsigned)d < (unsigned)g.Di && (unsigned)h < (unsigned)g.Hi;
          const int w_in0 = 2 * w0 - 1;
          const __nv_bfloat16* src = in_n + ((long long)(d * g.Hi + h) * g.Wi + w_in0) * g.Cin + gidx * 8;
          const unsigned dst0 = a_base + rr * (HX * 16);
#pragma unroll
          for (int v = 0; v < HX; ++v) {                 // v-th input voxel of the row: even v -> odd plane, odd v -> even plane
            const bool ok = row_ok && (unsigned)(w_in0 + v) < (unsigned)g.Wi;
            const int slot = (v & 1) ? (BW + 1) + (v >> 1) : (v >> 1);
            cp_async16(dst0 + slot * 16, ok ? src : in, ok);
            src += g.Cin;
          }
        }
      }
      if (BULK) {                                   // publish the halo on its own barrier as soon as it has landed
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(FULLA(astage));
      }
      astage ^= 1; if (astage == 0) a_phase ^= 1;
    };

    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const int tile = blockIdx.x + (chunk / KC) * gridDim.x;
      const int kc = chunk % KC;
      const int nt = tile % tl.NT;
      const __nv_bfloat16* wbase = wgt + (size_t)(nt * N_TILE) * g.Cin + kc * (8 * KG);
      for (int it = 0; it < NI_ITEMS; ++it) {
        mbar_wait_warp(EMPTYB(slot), slot_phase ^ 1, lane);
        if (BULK) {
          // weights first (one thread, TG bulk copies onto the slot's barrier), then -- if one is due -- the halo, on which all
          // producer warps block until it has landed
          if (tid == 0) {
            const unsigned b_base = smem_u32(sB + slot * B_BYTES);
            mbar_expect_tx(FULLB(slot), B_BYTES);
#pragma unroll
            for (int tt = 0; tt < TG; ++tt) {
              const size_t slice = ((size_t)(nt * KC + kc) * tl.items_T + g.tap_w[it * TG + tt]) * (B_TAP_BYTES / 2);
              bulk_g2s(b_base + tt * B_TAP_BYTES, tl.items + slice, B_TAP_BYTES, FULLB(slot));
            }
          }
          if (chunk == 0 && it == 0) load_halo(0);
          if (it == NI_ITEMS / 2 && chunk + 1 < n_chunks) load_halo(chunk + 1);
          if (++slot == B_SLOTS) { slot = 0; slot_phase ^= 1; }
          continue;
        }
        if (chunk == 0 && it == 0) load_halo(0);
        if (it == NI_ITEMS / 2 && chunk + 1 < n_chunks) load_halo(chunk + 1);
        {
          const unsigned b_base = smem_u32(sB + slot * B_BYTES);
          for (int i = tid; i < TG * N_TILE * KG; i += NUM_PROD) {
            const int tt = i / (N_TILE * KG); const int r = i - tt * (N_TILE * KG);
            const int kg = r / N_TILE, nr = r - kg * N_TILE;
            const __nv_bfloat16* wsrc = wbase + (size_t)g.tap_w[it * TG + tt] * ep.CoutPad * g.Cin;
            cp_async16(b_base + i * 16, wsrc + (size_t)nr * g.Cin + kg * 8, true);
          }
        }
        cp_async_commit();
        ++pending;
        if (pending > LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(FULLB(done_slot));
          done_slot = (done_slot + 1 == B_SLOTS) ? 0 : done_slot + 1;
          --pending;
        }
        if (++slot == B_SLOTS) { slot = 0; slot_phase ^= 1; }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    while (pending > 0) {
      if (lane == 0) mbar_arrive(FULLB(done_slot));
      done_slot = (done_slot + 1 == B_SLOTS) ? 0 : done_slot + 1;
      --pending;
    }
  } else if (warp < 4 + MT) {
    // ================================================================ MMA issuers: one warp (lane 0) per depth slice.
    // The MT slices accumulate into disjoint TMEM columns, so their MMA streams are independent; a single thread
    // cannot issue the 16..64-cycle MMAs of this kernel fast enough (~80 issue cycles each), MT threads can.
    if (lane == 0) {
      const int mt = warp - 4;
      unsigned slot = 0, slot_phase = 0, astage = 0, acc = 0, acc_phase = 0;
      unsigned afull_phase = 0;                      // BULK: phase of FULLA[astage]
      for (int tile = blockIdx.x; tile < tl.total; tile += gridDim.x) {
        mbar_wait(TEMPTY(acc), acc_phase ^ 1);
        tc_fence_after();
        const unsigned d_tmem = tmem_base + acc * ACC_COLS + mt * N_TILE;
        for (int kc = 0; kc < KC; ++kc) {
          if (BULK) { mbar_wait(FULLA(astage), afull_phase); tc_fence_after(); }
          // only the 14-bit start-address field changes between MMAs: add (byte offset >> 4) to a base descriptor
          const unsigned long long a_desc0 = make_desc(smem_u32(sA + astage * A_BYTES) + mt * SLICE_PITCH, LBO_A, SBO_A);
          for (int it = 0; it < T / TG; ++it) {
            mbar_wait(FULLB(slot), slot_phase);
            tc_fence_after();
            const unsigned long long b_item = make_desc(smem_u32(sB + slot * B_BYTES), LBO_B, SBO_B);
#pragma unroll
            for (int tt = 0; tt < TG; ++tt) {
              const unsigned long long b_desc0 = b_item + (unsigned long long)((tt * B_TAP_BYTES) >> 4);
              const unsigned long long a_tap = a_desc0 + (unsigned long long)(s_tapoff[it * TG + tt] >> 4);
              tc_mma(d_tmem, a_tap, b_desc0, IDESC, (kc | it | tt) != 0 ? 1u : 0u);
              if (KG == 4)
                tc_mma(d_tmem, a_tap + (unsigned long long)((2 * LBO_A) >> 4), b_desc0 + (unsigned long long)((2 * LBO_B) >> 4), IDESC, 1u);
            }
            tc_commit(EMPTYB(slot));
            if (++slot == B_SLOTS) { slot = 0; slot_phase ^= 1; }
          }
          tc_commit(EMPTYA(astage));
          astage ^= 1;
          if (BULK && astage == 0) afull_phase ^= 1;
        }
        tc_commit(TFULL(acc));
        if (ACC == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else acc_phase ^= 1;
      }
    }
  } else {
    // ================================================================ epilogue (4 warps -> TMEM lane quarter warp % 4)
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
      // per-warp rows of s_stat accumulate this tile's column sums (lane l owns column c*32 + l of its warp's row)
      if (do_stats) {
        for (int ch = lane; ch < N_TILE; ch += 32) { s_stat[0][q][ch] = 0.f; s_stat[1][q][ch] = 0.f; }
      }
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int d = d0 + mt;
        const bool ok = hw_ok && d < g.Ld;
        // physical voxel: lo * om + oo per axis (stride-2 dgrad writes one parity class of dx)
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
              // head outputs: fp32, written straight into the [N, anchors, C] layout (sample / voxel strides), channels
              // beyond Cout are padding of the weight pack
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
      // accumulator drained -> hand the TMEM stage back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY(acc));
      if (ACC == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else acc_phase ^= 1;
      if (do_stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = tid - (4 + MT) * 32;        // 0..127
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
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

template <int N_TILE, int MT, int TG, int ACC, int B_SLOTS, int S2 = 0, int BULK = 0>
int launch_tc(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st,
              const __nv_bfloat16* items = nullptr, int items_T = 0) {
  constexpr int KG = S2 ? 2 : ::KG;
  constexpr int HY = S2 ? 2 * BH + 1 : ::HY, HX = S2 ? 2 * BW + 1 : ::HX;
  TcTiles tl;
  tl.DB = (g.Ld + MT - 1) / MT; tl.HB = (g.Lh + BH - 1) / BH; tl.WB = (g.Lw + BW - 1) / BW; tl.NT = ep.CoutPad / N_TILE;
  tl.total = g.N * tl.DB * tl.HB * tl.WB * tl.NT;
  tl.items = items; tl.items_T = items_T;
  constexpr int HV = (S2 ? 2 * MT + 1 : MT + 2) * HY * HX;
  const size_t smem = (size_t)A_STAGES * KG * HV * 16 + (size_t)B_SLOTS * TG * N_TILE * KG * 16 + 8 * (2 * B_SLOTS + A_STAGES + 4 + (BULK ? A_STAGES : 0));
  static NndPerDeviceOnce attr_set;
  if (attr_set.need()) {
    NND_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<N_TILE, MT, TG, ACC, B_SLOTS, S2, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int grid = tl.total < NND_NUM_SMS ? tl.total : NND_NUM_SMS;
  conv_tc_kernel<N_TILE, MT, TG, ACC, B_SLOTS, S2, BULK><<<grid, tc_threads(MT), smem, st>>>(in, w, g, ep, tl);
  NND_LAUNCH_CHECK("conv_tc_kernel");
  return NND_OK;
}

}  // namespace

static int g_tc_deep_ring = 2;      // 0: 6 / 12 slots of one tap, 1: 10 / 16 slots, 2 (default): three taps per item
extern "C" void nnd_conv_set_tc_ring(int deep) { g_tc_deep_ring = deep; }

int nnd_conv_tc_supported(const ConvGeom& g, const ConvEpilogue& ep) {
  if (g.sd != 1 || g.sh != 1 || g.sw != 1) return 0;                       // input position = logical position + offset
  const bool identity = g.omd == 1 && g.omh == 1 && g.omw == 1 && !g.ood && !g.ooh && !g.oow;
  if (identity && (g.Ld != g.Do || g.Lh != g.Ho || g.Lw != g.Wo)) return 0;
  // 3x3x3 / 1x3x3 layers, or the >= 4-tap parity classes of a stride-2 dgrad (output = lo * 2 + parity)
  if (!(g.T >= 9 || (!identity && g.T >= 4)) || g.Cin % 32) return 0;
  if (ep.stat_sum && !identity) return 0;
  for (int t = 0; t < g.T; ++t)
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
  if (ep.out_fp32) {                       // strided fp32 head outputs: no residual / statistics, identity mapping
    if (!identity || ep.residual || ep.stat_sum || ep.CoutPad % 32) return 0;
  } else {
    if (ep.Cout % 32 || ep.CoutPad != ep.Cout) return 0;
    if (ep.out_v_stride != ep.Cout || ep.out_n_stride != (long long)g.Do * g.Ho * g.Wo * ep.Cout) return 0;
  }
  if (g.Lh < 8 || g.Lw < 8) return 0;
  return 1;
}

// Stride-2 gathers (S2 variant): 3x3x3 stride-2 convolutions and the 2x2x2 stride-2 convolution behind an up-convolution's dgrad.
int nnd_conv_tc_s2_supported(const ConvGeom& g, const ConvEpilogue& ep) {
  if (g.sd < 1 || g.sd > 2 || g.sh != 2 || g.sw != 2) return 0;
  if (g.omd != 1 || g.omh != 1 || g.omw != 1 || g.ood || g.ooh || g.oow) return 0;
  if (g.Ld != g.Do || g.Lh != g.Ho || g.Lw != g.Wo) return 0;
  if (g.T < 4 || g.Cin % 16) return 0;
  for (int t = 0; t < g.T; ++t)
    if (g.off_d[t] < -1 || g.off_d[t] > 1 || g.off_h[t] < -1 || g.off_h[t] > 1 || g.off_w[t] < -1 || g.off_w[t] > 1) return 0;
  if (ep.out_fp32 || ep.Cout % 32 || ep.CoutPad != ep.Cout) return 0;
  if (ep.out_v_stride != ep.Cout || ep.out_n_stride != (long long)g.Do * g.Ho * g.Wo * ep.Cout) return 0;
  if (g.Lh < 8 || g.Lw < 8) return 0;
  return 1;
}

int nnd_conv_tc_s2(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st) {
  if (!nnd_conv_tc_s2_supported(g, ep)) return NND_ERR_ARG;
  const bool g3 = g.T % 3 == 0;
  // two depth slices per tile: the de-interleaved halo of 16 channels (5 x 33 x 17 voxels) is 88 KB per stage
  if (ep.CoutPad % 128 == 0) return g3 ? launch_tc<128, 2, 3, 2, 3, 1>(in, w, g, ep, st) : launch_tc<128, 2, 1, 2, 8, 1>(in, w, g, ep, st);
  if (ep.CoutPad % 64 == 0) return g3 ? launch_tc<64, 2, 3, 2, 4, 1>(in, w, g, ep, st) : launch_tc<64, 2, 1, 2, 8, 1>(in, w, g, ep, st);
  return g3 ? launch_tc<32, 2, 3, 2, 4, 1>(in, w, g, ep, st) : launch_tc<32, 2, 1, 2, 8, 1>(in, w, g, ep, st);
}

// BULK variant: the 3-taps-per-item instantiations of the stride-1 kernel with the item-order weight pack (see the header comment).
// items_n_tile: the N_TILE the pack was made for (must be the one the dispatch below picks), items_T: weight slices in the pack.
int nnd_conv_tc_bulk_supported(const ConvGeom& g, const ConvEpilogue& ep, const void* items, int items_n_tile, int items_T) {
  if (!items || !nnd_conv_tc_supported(g, ep) || g.T % 3 != 0 || g_tc_deep_ring != 2) return 0;
  const int n_tile = ep.CoutPad % 128 == 0 ? 128 : (ep.CoutPad % 64 == 0 ? 64 : 32);
  if (items_n_tile != n_tile || ((size_t)items & 15)) return 0;
  for (int t = 0; t < g.T; ++t) if (g.tap_w[t] >= items_T) return 0;
  return 1;
}

int nnd_conv_tc_bulk(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st,
                     const __nv_bfloat16* items, int items_n_tile, int items_T) {
  if (!nnd_conv_tc_bulk_supported(g, ep, items, items_n_tile, items_T)) return NND_ERR_ARG;
  if (ep.CoutPad % 128 == 0) {
    const long long tiles4 = (long long)g.N * ((g.Ld + 3) / 4) * ((g.Lh + BH - 1) / BH) * ((g.Lw + BW - 1) / BW) * (ep.CoutPad / 128);
    if (tiles4 >= NND_NUM_SMS) return launch_tc<128, 4, 3, 1, 3, 0, 1>(in, w, g, ep, st, items, items_T);
    return launch_tc<128, 2, 3, 2, 4, 0, 1>(in, w, g, ep, st, items, items_T);
  }
  if (ep.CoutPad % 64 == 0) return launch_tc<64, 4, 3, 2, 6, 0, 1>(in, w, g, ep, st, items, items_T);
  return launch_tc<32, 4, 3, 2, 6, 0, 1>(in, w, g, ep, st, items, items_T);
}

int nnd_conv_tc(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st) {
  if (!nnd_conv_tc_supported(g, ep)) return NND_ERR_ARG;
  const bool g3 = g.T % 3 == 0;          // taps come in (dz, dy) rows of three -> 3 taps per pipeline item
  if (ep.CoutPad % 128 == 0) {
    // 4 depth slices share each weight slice (halves the L2 weight traffic); small volumes keep 2-slice tiles so that
    // the persistent grid still has >= one tile per SM
    const long long tiles4 = (long long)g.N * ((g.Ld + 3) / 4) * ((g.Lh + BH - 1) / BH) * ((g.Lw + BW - 1) / BW) * (ep.CoutPad / 128);
    // weight ring depth: the 8 KB weight slices are the latency-bound stream of these layers (L2 -> smem, ~32 KB in flight
    // per SM with 6 slots); deeper rings keep more bytes in flight (A/B: nnd_conv_set_tc_ring(0) restores 6 / 12)
    // mode 2: three taps per pipeline item (24 KB): the producers' per-item bookkeeping (barrier wait, commit, wait_group,
    // fence, arrive: ~400 cycles) bounds the weight stream at ~40 GB/s per SM with 8 KB items
    if (g_tc_deep_ring == 2 && g3) {
      if (tiles4 >= NND_NUM_SMS) return launch_tc<128, 4, 3, 1, 3>(in, w, g, ep, st);
      return launch_tc<128, 2, 3, 2, 4>(in, w, g, ep, st);
    }
    if (tiles4 >= NND_NUM_SMS) return g_tc_deep_ring ? launch_tc<128, 4, 1, 1, 10>(in, w, g, ep, st) : launch_tc<128, 4, 1, 1, 6>(in, w, g, ep, st);
    return g_tc_deep_ring ? launch_tc<128, 2, 1, 2, 16>(in, w, g, ep, st) : launch_tc<128, 2, 1, 2, 12>(in, w, g, ep, st);
  }
  if (ep.CoutPad % 64 == 0) return g3 ? launch_tc<64, 4, 3, 2, 6>(in, w, g, ep, st) : launch_tc<64, 4, 1, 2, 12>(in, w, g, ep, st);
  return g3 ? launch_tc<32, 4, 3, 2, 6>(in, w, g, ep, st) : launch_tc<32, 4, 1, 2, 12>(in, w, g, ep, st);
}
