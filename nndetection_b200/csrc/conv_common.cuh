// Shared definitions of the convolution kernels.
//
// Activations are NDHWC bf16 (torch channels_last_3d), weights are packed per filter tap as [tap][CoutPad][Cin]
// (K-major rows).  One "gather convolution" form covers every layer type of the Retina U-Net
// (nndet/arch/conv.py:297-348 builds torch.nn.Conv3d / ConvTranspose3d for them):
//     out[n, lo*om + oo, co] = sum_taps sum_ci  in[n, lo*s + off_tap, ci] * W[tap_w][co][ci]
// with `lo` running over a logical output grid:
//   * conv fprop (k in {1,3}, stride in {1,2}, pad (k-1)/2): s = stride, off = tap - pad, om = 1, oo = 0
//   * conv dgrad, stride 1: s = 1, off = pad - tap, transposed weights
//   * conv dgrad, stride 2: one launch per output parity class p (om = 2, oo = p), taps with (p + pad - k) even
//   * transposed conv (k = stride) fprop: one launch per parity class, single tap; its dgrad is a k = s conv
#pragma once
#include "common.cuh"

constexpr int NND_MAX_TAPS = 27;

struct ConvGeom {
  int N, Di, Hi, Wi, Cin;          // input tensor [N, Di, Hi, Wi, Cin]
  int Ld, Lh, Lw;                  // logical output grid per sample
  int sd, sh, sw;                  // input position = lo * s + off
  int Do, Ho, Wo;                  // physical output grid (addressing)
  int omd, omh, omw, ood, ooh, oow;  // physical output position = lo * om + oo
  int T;                           // number of taps
  signed char off_d[NND_MAX_TAPS], off_h[NND_MAX_TAPS], off_w[NND_MAX_TAPS];
  unsigned char tap_w[NND_MAX_TAPS];   // which packed weight slice a tap uses
};

struct ConvEpilogue {
  void* out;                       // bf16 or fp32
  long long out_n_stride;          // elements between samples
  long long out_v_stride;          // elements between voxels (>= Cout)
  int out_fp32;
  int Cout, CoutPad;
  const float* bias;               // [Cout] or null
  const float* scale;              // device scalar multiplier (arch/layers/scale.py) or null
  const __nv_bfloat16* residual;   // same geometry as out (bf16, dense NDHWC with Cout channels) or null
  float* stat_sum;                 // [N, Cout] per-(sample, channel) sum of the stored values, or null
  float* stat_sq;                  // [N, Cout] sum of squares
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(unsigned addr, unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(unsigned addr, unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, unsigned a0, unsigned a1, unsigned a2, unsigned a3,
                                               unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
