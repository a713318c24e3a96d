// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, both operands in shared memory, K-major no swizzle)
// as a function of N and of the number of issuing threads.  Decides the tile shapes of csrc/conv_tc.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_rate scripts/mma_rate.cu && ./scripts/mma_rate
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long make_desc(unsigned addr, unsigned lbo, unsigned sbo) {
  return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) |
         ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                       unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}

// mode 0: every MMA accumulates into the same N columns; mode 1: sliding windows (column base advances by N/3 per MMA,
// wrapping) -- the access pattern of the z-window convolution kernel.
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(160, 1) rate_kernel(int N, int issuers, int iters, int mode, int style, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned s_tmem;
  __shared__ unsigned long long bar[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = s_tmem;
  const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (style == 1 && warp < issuers) {
    // convergent warp: descriptors stay warp-uniform, only the MMA itself is predicated on the elected lane
    const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
    const unsigned a_base = smem_u32(smem) + uwarp * 8192;
    const unsigned b_base = smem_u32(smem) + 40960 + uwarp * 8192;
    // mode 2 / 3: A laid out like the convolution halo (8-row groups at a 160-byte pitch, k-groups 2880 B apart); mode 2 keeps
    // every core matrix inside one 128-byte line (start offsets multiples of 640 B), mode 3 adds the dx tap shifts (16 / 32 B)
    const unsigned long long a0 = mode >= 2 ? make_desc(a_base, 2880, 160) : make_desc(a_base, 2048, 128);
    const unsigned long long b0 = make_desc(b_base, N * 16, 128);
    const int width = mode == 1 ? N / 3 : N;
    const int span = 512 / issuers;
    const unsigned cbase = __shfl_sync(0xffffffffu, tmem, 0) + uwarp * span;
    t0 = clock64();
    int col = 0;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      const unsigned long long ashift = mode == 3 ? (unsigned long long)(i % 3) : (mode == 2 ? 0ull : (unsigned long long)((i & 7) * 16));
      if (elect_one()) tc_mma(cbase + col, a0 + ashift, b0, idesc, 1u);
      __syncwarp();
      if (mode == 1) { col += width; if (col + N > span) col = 0; }
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0);
    t1 = clock64();
    if (lane == 0) out[blockIdx.x * 4 + warp] = t1 - t0;
  } else if (style == 0 && warp < issuers && lane == 0) {
    // A: 128 rows x 16 k  (2 k-groups: LBO 2048+..), B: N rows x 16 k
    const unsigned a_base = smem_u32(smem) + warp * 8192;
    const unsigned b_base = smem_u32(smem) + 40960 + warp * 8192;
    const unsigned long long a0 = make_desc(a_base, 2048, 128);
    const unsigned long long b0 = make_desc(b_base, N * 16, 128);
    const int width = mode == 1 ? N / 3 : N;
    const int span = 512 / issuers;                    // TMEM columns owned by this issuer
    const unsigned cbase = tmem + warp * span;
    t0 = clock64();
    int col = 0;
    for (int i = 0; i < iters; ++i) {
      tc_mma(cbase + col, a0 + (unsigned long long)((i & 7) * 16), b0, idesc, 1u);
      if (mode == 1) { col += width; if (col + N > span) col = 0; }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    mbar_wait(smem_u32(&bar[warp]), 0);
    t1 = clock64();
    out[blockIdx.x * 4 + warp] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 4 * sizeof(long long));
  long long h[148 * 4];
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int iters = 2048;
  const int Ns[] = {32, 48, 64, 96, 128, 192, 256};
  printf("style mode N issuers cycles_per_mma_per_SM  (ideal 128*N/256 = N/2)\n");
  for (int style = 1; style >= 0; --style)
  for (int mode = 0; mode < 4; ++mode)
    for (int N : Ns)
      for (int issuers : {1, 2, 4}) {
        if (mode == 1 && (N % 48 != 0)) continue;
        if (mode >= 2 && (style == 0 || (N != 96 && N != 128 && N != 32))) continue;
        if (N > 512 / issuers) continue;
        cudaMemset(d, 0, sizeof(h));
        rate_kernel<<<148, 160, 96 * 1024>>>(N, issuers, iters, mode, style, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148 * 4; ++i) if (h[i] > mx) mx = h[i];
        printf("%d %d %3d %d %.1f\n", style, mode, N, issuers, (double)mx / (iters * issuers));
      }
  return 0;
}
