// CTA-pair probe for the round-2 128-channel kernels: does `tcgen05.mma.cta_group::2` behave the way
// the planned kernels assume, and what does it cost?  WRITTEN WITHOUT A GPU -- a probe, not product code: run it under `timeout`.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nndetection_b200/csrc -o scripts/pair_probe scripts/pair_probe.cu
//   timeout 60 scripts/pair_probe
//
// One cluster of two CTAs.  Both stage a 128 x 16 bf16 A tile (their own M half) and a (N/2) x 16 B tile in the canonical K-major
// no-swizzle layout at the SAME shared-memory offsets, both allocate TMEM with cta_group::2, the leader issues ONE M = 256 MMA and
// commits to a barrier in both CTAs (multicast), each CTA reads its 128 x N accumulator back.
//   mode 0: A[r][k] = (k == r % 16), B[n][k] = 64 * rank + n  ->  D[r][c] = id of the (rank, n) row of B that fed column c:
//           prints which CTA's B rows end up in which accumulator columns (the assumption: columns 0..N/2-1 <- CTA 0, rest <- CTA 1).
//   mode 1: B[n][k] = k -> D[r][c] must be r % 16 in both CTAs (A rows of CTA 1 really are M rows 128..255).
//   mode 2: issue-rate loop: cycles per M = 256, N = {128, 256}, K = 16 MMA from one elected thread (compare scripts/mma_rate.cu:
//           64 cycles for the single-CTA N = 128 form).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include "tcgen05.cuh"

namespace {

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-CTA MMA: disable-output-lane mask of 8 words (cutlass: SM100_MMA_F16BF16_2x1SM_SS)
__device__ __forceinline__ void tc_mma_pair(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                            unsigned accumulate) {
  const unsigned z = 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(unsigned bar, unsigned short cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_probe_kernel(float* __restrict__ out, int mode, int iters, long long* __restrict__ cycles) {
  constexpr int NH = N / 2;                                   // B rows held by each CTA
  constexpr int A_BYTES = 2 * 128 * 16, B_BYTES = 2 * NH * 16;
  constexpr unsigned IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((256u >> 4) << 24);
  constexpr int TMEM_COLS = N <= 128 ? 128 : 256;
  __shared__ __align__(1024) unsigned char sA[A_BYTES];
  __shared__ __align__(1024) unsigned char sB[B_BYTES];
  __shared__ __align__(8) unsigned long long s_done;
  __shared__ unsigned s_tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned rank = cluster_ctarank();

  // canonical K-major, no swizzle: [k group of 8][row][8 elements]; LBO = rows * 16 B between k groups, SBO = 128 B between 8-row groups
  __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(sA);
  __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(sB);
  for (int i = tid; i < 128 * 16; i += 128) {
    const int r = i / 16, k = i % 16;
    a[((k / 8) * 128 + r) * 8 + (k % 8)] = __float2bfloat16(k == r % 16 ? 1.f : 0.f);
  }
  for (int i = tid; i < NH * 16; i += 128) {
    const int n = i / 16, k = i % 16;
    const float v = mode == 0 ? (float)(64 * (int)rank + (n % 64)) : (mode == 1 ? (float)k : 0.f);
    b[((k / 8) * NH + n) * 8 + (k % 8)] = __float2bfloat16(v);
  }
  const unsigned done = smem_u32(&s_done);
  if (tid == 0) {
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {                                            // the same warp in BOTH CTAs (cute::TMEM::Allocator2Sm)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const unsigned tmem_base = s_tmem_base;

  if (rank == 0 && warp == 1) {                               // leader CTA issues for the pair
    const unsigned long long ad = make_desc(smem_u32(sA), 128 * 16, 128);
    const unsigned long long bd = make_desc(smem_u32(sB), NH * 16, 128);
    long long t0 = 0, t1 = 0;
    if (mode == 2) t0 = clock64();
    for (int it = 0; it < iters; ++it)
      if (elect_one()) tc_mma_pair(tmem_base, ad, bd, IDESC, it != 0 ? 1u : 0u);
    if (elect_one()) tc_commit_pair(done, (unsigned short)0x3);
    __syncwarp();
    if (mode == 2) {
      mbar_wait_warp(done, 0, lane);
      t1 = clock64();
      if (lane == 0) *cycles = t1 - t0;
    }
  }
  mbar_wait_warp(done, 0, lane);                              // every warp of both CTAs: the MMAs of the pair have completed
  tc_fence_after();
  if (mode != 2) {
    float* o = out + ((size_t)rank * 128 + warp * 32 + lane) * N;
#pragma unroll 1
    for (int c = 0; c < N / 32; ++c) {
      unsigned v[32];
      tmem_ld32(tmem_base + ((unsigned)(warp * 32) << 16) + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) o[c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int N>
void run() {
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(float) * 2 * 128 * N));
  CK(cudaMalloc(&d_cyc, sizeof(long long)));
  std::vector<float> h(2 * 128 * N);
  for (int mode = 0; mode < 2; ++mode) {
    CK(cudaMemset(d_out, 0xff, sizeof(float) * 2 * 128 * N));
    pair_probe_kernel<N><<<2, 128>>>(d_out, mode, 1, d_cyc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, sizeof(float) * h.size(), cudaMemcpyDeviceToHost));
    if (mode == 0) {
      printf("N=%d mode 0: source (rank, n) of accumulator column c, read from row 0 of each CTA (expected: c < N/2 -> rank 0, else rank 1)\n", N);
      for (int cta = 0; cta < 2; ++cta) {
        int ok = 1;
        for (int c = 0; c < N; ++c) {
          const int id = (int)h[((size_t)cta * 128 + 0) * N + c];
          const int expect = 64 * (c / (N / 2)) + (c % (N / 2)) % 64;
          if (id != expect) ok = 0;
        }
        printf("  CTA %d: columns 0,1,2 .. = %g %g %g ... N/2-1, N/2 = %g %g ... last = %g   -> %s\n", cta, h[(size_t)cta * 128 * N], h[(size_t)cta * 128 * N + 1],
               h[(size_t)cta * 128 * N + 2], h[(size_t)cta * 128 * N + N / 2 - 1], h[(size_t)cta * 128 * N + N / 2], h[(size_t)cta * 128 * N + N - 1],
               ok ? "as assumed" : "DIFFERENT (dump below)");
        if (!ok) { for (int c = 0; c < N; ++c) printf("%g ", h[(size_t)cta * 128 * N + c]); printf("\n"); }
      }
      // rows: every row r of both CTAs must show the same ids (A is one-hot on k = r % 16 and B does not depend on k)
      int rows_ok = 1;
      for (int cta = 0; cta < 2; ++cta) for (int r = 1; r < 128; ++r) for (int c = 0; c < N; ++c)
        if (h[((size_t)cta * 128 + r) * N + c] != h[(size_t)cta * 128 * N + c]) rows_ok = 0;
      printf("  all 128 rows of both CTAs identical: %s\n", rows_ok ? "yes" : "NO");
    } else {
      int bad = 0;
      for (int cta = 0; cta < 2; ++cta) for (int r = 0; r < 128; ++r) for (int c = 0; c < N; ++c)
        if (h[((size_t)cta * 128 + r) * N + c] != (float)(r % 16)) ++bad;
      printf("N=%d mode 1: D[r][c] == r %% 16 in both CTAs: %s (%d mismatches)\n", N, bad ? "NO" : "yes", bad);
    }
  }
  const int iters = 4000;
  pair_probe_kernel<N><<<2, 128>>>(d_out, 2, iters, d_cyc);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
  printf("N=%d mode 2: %.1f cycles per M=256 x N=%d x K=16 MMA (one elected issuer; 2 x 128 x %d x 16 x 2 FLOP each)\n", N, (double)cyc / iters, N, N);
  CK(cudaFree(d_out)); CK(cudaFree(d_cyc));
}

}  // namespace

int main() {
  run<128>();
  run<256>();
  return 0;
}
