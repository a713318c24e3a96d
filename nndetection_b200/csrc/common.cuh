// Shared helpers for the nndet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

// Status codes of the C ABI (include/nndet_b200.h).
#define NND_OK 0
#define NND_ERR_ARG 1          // bad argument (null pointer, negative size, unsupported shape)
#define NND_ERR_WORKSPACE 2    // caller-provided workspace too small
#define NND_ERR_CUDA 3         // a CUDA runtime call / launch failed (see nnd_last_cuda_error)

extern "C" int nnd_set_cuda_error(cudaError_t e, const char* where);
extern unsigned long long g_nnd_launches;       // kernels launched through the C ABI (bench.py: gpu_launches)

#define NND_CUDA_TRY(expr)                                              \
  do {                                                                  \
    cudaError_t _e = (expr);                                            \
    if (_e != cudaSuccess) return nnd_set_cuda_error(_e, #expr);        \
  } while (0)

#define NND_LAUNCH_CHECK(name)                                          \
  do {                                                                  \
    ++g_nnd_launches;                                                   \
    cudaError_t _e = cudaGetLastError();                                \
    if (_e != cudaSuccess) return nnd_set_cuda_error(_e, name);         \
  } while (0)

// Function attributes (opt-in shared-memory size) are per DEVICE: a launcher's "already set" flag must be too, or a process that
// uses a second GPU launches with the default 48 KB limit there.  `need()` is true once per device.
struct NndPerDeviceOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

static inline size_t nnd_align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

template <typename T>
static inline T* nnd_carve(char*& p, size_t count) {
  T* r = reinterpret_cast<T*>(p);
  p += nnd_align_up(count * sizeof(T));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Counter hash used to draw sampling priorities (mirrors oracle/box_oracle.py:mix32).
__device__ __host__ __forceinline__ uint32_t nnd_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
__device__ __host__ __forceinline__ uint32_t nnd_hash_priority(uint32_t idx, uint32_t seed, uint32_t stream) {
  uint32_t s = seed * 0x9E3779B1u + stream * 0x85EBCA77u;
  return nnd_mix32(idx ^ s);
}

constexpr int NND_NUM_SMS = 148;   // B200
