// C-ABI entry points of the convolution family (include/nndet_b200.h).  Geometry travels as a flat int array:
//   geom[0..20] = N, Di, Hi, Wi, Cin,  Ld, Lh, Lw,  sd, sh, sw,  Do, Ho, Wo,  omd, omh, omw,  ood, ooh, oow,  T
//   geom[21 + 4*t ..] = off_d, off_h, off_w, weight_tap   for each of the T taps
// (see conv_common.cuh for the gather-convolution form these numbers describe).
#include "conv_common.cuh"

#include <cstdio>
#include <mutex>
#include <vector>

int nnd_conv_igemm(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st);
int nnd_conv_tc(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st);
int nnd_conv_tc_supported(const ConvGeom& g, const ConvEpilogue& ep);
int nnd_conv_tct(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st);
int nnd_conv_tct_supported(const ConvGeom& g, const ConvEpilogue& ep);
int nnd_conv_tct_s2(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st);
int nnd_conv_tct_s2_supported(const ConvGeom& g, const ConvEpilogue& ep);
int nnd_conv_tc_s2(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st);
int nnd_conv_tc_s2_supported(const ConvGeom& g, const ConvEpilogue& ep);
int nnd_conv_tc_bulk_supported(const ConvGeom& g, const ConvEpilogue& ep, const void* items, int items_n_tile, int items_T);
int nnd_conv_tc_bulk(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st,
                     const __nv_bfloat16* items, int items_n_tile, int items_T);
int nnd_conv_pw(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st);
int nnd_conv_pw_supported(const ConvGeom& g, const ConvEpilogue& ep);
int nnd_conv_upconv(const void* x, const void* w_packed, int N, int D, int H, int W, int Cin, int Cout, int sd, int sh, int sw,
                    void* out, const float* bias, const void* residual, cudaStream_t st);
int nnd_conv_tcs(const __nv_bfloat16* in, const __nv_bfloat16* w, const ConvGeom& g, const ConvEpilogue& ep, cudaStream_t st);
int nnd_conv_tcs_supported(const ConvGeom& g, const ConvEpilogue& ep);
int nnd_conv_tcs_profitable(const ConvGeom& g, const ConvEpilogue& ep);
void nnd_conv_tcs_set_issuers(int n);
int nnd_conv_wgrad(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                   long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st);
int nnd_conv_wgrad_tc_supported(const ConvGeom& g, int Cdy, int Cx);
int nnd_conv_wgrad_tc_strided_supported(const ConvGeom& g, int Cdy, int Cx);
int nnd_conv_wgrad_tc(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                      long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st);
int nnd_conv_wgrad_tma_supported(const ConvGeom& g, int Cdy, int Cx);
int nnd_conv_wgrad_tma(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                       long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, int mode, void* ws, long long ws_bytes,
                       cudaStream_t st);
long long nnd_conv_wgrad_tma_workspace(const ConvGeom& g, int Cdy, int Cx, int Cout, int Cin);
int nnd_conv_wgrad_tma32_supported(const ConvGeom& g, int Cdy, int Cx);
long long nnd_conv_wgrad_tma32_workspace(const ConvGeom& g, int Cout, int Cin);
int nnd_conv_wgrad_tma32(const __nv_bfloat16* dy, const __nv_bfloat16* x, const ConvGeom& g, float* dw, long long s_co, long long s_ci,
                         long long s_tap, int Cout, int Cin, void* ws, long long ws_bytes, cudaStream_t st);
int nnd_conv_wgrad_tma_s2_supported(const ConvGeom& g, int Cdy, int Cx);
int nnd_conv_wgrad_tma_s2(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                          long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, int mode, void* ws, long long ws_bytes,
                          cudaStream_t st);
long long nnd_conv_wgrad_tma_s2_workspace(const ConvGeom& g, int Cdy, int Cx, int Cout, int Cin);
int nnd_conv_wgrad_tc32_supported(const ConvGeom& g, int Cdy, int Cx);
int nnd_conv_wgrad_tc32_profitable(const ConvGeom& g);
int nnd_conv_wgrad_tc32(const __nv_bfloat16* dy, const __nv_bfloat16* x, const ConvGeom& g, float* dw, long long s_co, long long s_ci,
                        long long s_tap, int Cout, int Cin, cudaStream_t st);
int nnd_conv_wgrad_tcn_supported(const ConvGeom& g, int Cdy, int Cx);
int nnd_conv_wgrad_tcn(const __nv_bfloat16* dy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw, long long s_co,
                       long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st);
int nnd_conv_wgrad_halo_supported(const ConvGeom& g, int Cdy, int Cx);
int nnd_conv_wgrad_halo(const __nv_bfloat16* dy, int Cdy, const __nv_bfloat16* x, int Cx, const ConvGeom& g, float* dw,
                        long long s_co, long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st);
int nnd_conv_first_fprop(const float* x, const float* w, const ConvGeom& g, int Cout, __nv_bfloat16* out, float* stat_sum,
                         float* stat_sq, cudaStream_t st);
int nnd_conv_first_wgrad(const float* x, const __nv_bfloat16* dy, const ConvGeom& g, int Cout, float* dw, cudaStream_t st);

namespace {
int parse_geom(const int* a, ConvGeom& g) {
  if (!a) return NND_ERR_ARG;
  g.N = a[0]; g.Di = a[1]; g.Hi = a[2]; g.Wi = a[3]; g.Cin = a[4];
  g.Ld = a[5]; g.Lh = a[6]; g.Lw = a[7];
  g.sd = a[8]; g.sh = a[9]; g.sw = a[10];
  g.Do = a[11]; g.Ho = a[12]; g.Wo = a[13];
  g.omd = a[14]; g.omh = a[15]; g.omw = a[16]; g.ood = a[17]; g.ooh = a[18]; g.oow = a[19];
  g.T = a[20];
  if (g.T < 1 || g.T > NND_MAX_TAPS) return NND_ERR_ARG;
  for (int t = 0; t < g.T; ++t) {
    g.off_d[t] = (signed char)a[21 + 4 * t]; g.off_h[t] = (signed char)a[22 + 4 * t]; g.off_w[t] = (signed char)a[23 + 4 * t];
    g.tap_w[t] = (unsigned char)a[24 + 4 * t];
  }
  return NND_OK;
}
int g_force_igemm = 0;
int g_wgrad_tc = 1;
int g_stream = 1;
int g_wgrad_strided = 1;      // validated on a B200 in round 2 (tests/test_strided_tcgen05_gpu.py), default since
int g_gather_strided = 1;
int g_tc_bulk = 0;
int g_gather_tma = 3;         // TMA-fed tile kernel (conv_tct.cu; validated on a B200 in round 2): bit 0 stride-1 forms, bit 1 stride-2 forms
int g_wgrad_tma = 1;          // TMA-fed tcgen05 wgrad (conv_wgrad_tma.cu; validated on a B200 in round 2): bit 0 on, bits 1-5 A/B and timing variants
int g_pw = 1;                 // TMA-fed pointwise GEMM (conv_pw.cu) for single-tap gathers (validated on a B200 in round 2; 0 = A/B)

// ---- per-launch trace (profiling aid, off by default): which kernel served which layer shape and how long it ran.
// ncu names kernels, not layers; this table is what maps the step time onto the network.
struct TraceRec {
  const char* kind;      // fprop | wgrad | first_fprop | first_wgrad   (dgrad launches are gather convolutions too: see T / strides)
  const char* kernel;    // dispatch decision
  int N, Di, Hi, Wi, Cin, Cout, Ld, Lh, Lw, sd, sh, sw, T;
  cudaEvent_t e0, e1;
  cudaStream_t st;
};
int g_trace_on = 0;
std::vector<TraceRec> g_trace;
std::mutex g_trace_mu;       // forward runs on the caller's thread, backward on autograd's worker thread

struct TraceScope {
  int idx = -1;
  cudaStream_t st;
  TraceScope(const char* kind, const char* kernel, const ConvGeom& g, int Cin, int Cout, cudaStream_t s) : st(s) {
    if (!g_trace_on) return;
    TraceRec r{kind, kernel, g.N, g.Di, g.Hi, g.Wi, Cin, Cout, g.Ld, g.Lh, g.Lw, g.sd, g.sh, g.sw, g.T, nullptr, nullptr, s};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    std::lock_guard<std::mutex> lk(g_trace_mu);
    g_trace.push_back(r);
    idx = (int)g_trace.size() - 1;
  }
  ~TraceScope() {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(g_trace_mu);
    cudaEventRecord(g_trace[idx].e1, st);
  }
};
}  // namespace

extern "C" {

// 1: route eligible layers to the tcgen05 kernel (default), 0: always use the mma.sync kernel (cross-check / debugging)
void nnd_conv_set_tensor_path(int enable_tcgen05) { g_force_igemm = !enable_tcgen05; }
// 1 (default): tcgen05 wgrad where eligible; 0: mma.sync halo wgrad (A/B); 2: also the stacked-tap 32-channel kernel on
// volumes too small to fill the grid (tests)
void nnd_conv_set_wgrad_tc(int enable) { g_wgrad_tc = enable; }
// 1: stride-2 convolutions take the de-interleaved tcgen05 wgrad (conv_wgrad_tc.cu, SW = 2) instead of the mma.sync kernels.
// Default 1 since round 2 (validated on a B200: tests/test_strided_tcgen05_gpu.py; luna step 32.6 -> 30.8 ms with both switches); 0 = A/B.
void nnd_conv_set_wgrad_strided_tc(int enable) { g_wgrad_strided = enable; }
// 1: stride-2 gathers (3x3x3 stride-2 convolutions, dgrad of up-convolutions) take the de-interleaved-halo tcgen05 tile kernel
// (conv_tc.cu, S2 = 1) instead of the mma.sync kernel.  Default 1 since round 2 (see above).
void nnd_conv_set_gather_strided_tc(int enable) { g_gather_strided = enable; }
// 1: launches that bring an item-order weight pack (nnd_conv_gather_bf16_items) stream it with cp.async.bulk (conv_tc.cu, BULK = 1).
// Default 0: the variant DEADLOCKS on the device (round-2 run: its test hit the 200 s timeout) -- kept only as a record, do not enable.
void nnd_conv_set_tc_bulk(int enable) { g_tc_bulk = enable; }
// 1: single-tap gathers (1x1x1 convolutions and their dgrad, parity classes of up-convolutions) take the TMA-fed tcgen05 GEMM of
// conv_pw.cu instead of the mma.sync gather kernel.
void nnd_conv_set_pointwise_tma(int enable) { g_pw = enable; }
// bit 0: stride-1 3x3x3 / 1x3x3 weight gradients with channel counts in multiples of 64 take the TMA-fed kernel of conv_wgrad_tma.cu
// instead of the cp.async one (conv_wgrad_tc.cu); bit 1: descriptors carry base_offset = (start >> 7) & 7 for row-shifted starts;
// bit 2: one N = 64 MMA per dx tap instead of the N = 192 stack (A/B of the descriptor model); bits 3-5 timing experiments; bit 6: ignore
// the workspace (atomics straight into dW); bit 7: the stride-2 / transposed forms stay on the cp.async kernel (conv_wgrad_tc.cu, SW = 2);
// bit 8: the 32 -> 32 layers stay on the cp.async stacked-tap kernel (conv_wgrad_tc32.cu).
void nnd_conv_set_wgrad_tma(int mode) { g_wgrad_tma = mode; }
// bit 0: launches the tcgen05 tile kernel serves (conv_tc.cu: 128-channel stride-1 layers, small volumes, >= 4-tap stride-2 dgrad classes)
// take its TMA-fed variant (conv_tct.cu); bit 1: the same for the stride-2 forms (conv_tc.cu S2 = 1)
void nnd_conv_set_gather_tma(int mode) { g_gather_tma = mode; }
// 1 (default): streaming z-window tcgen05 kernel (conv_tcs.cu) for the 32/64-channel 3x3x3 stride-1 layers when the volume
// is large enough to feed the persistent grid; 2: whenever the shape is supported (tests); 0: tile kernel.
// issuers: 1 or 2 MMA-issuing warps in that kernel (2 = default; 1 = fixed accumulation order)
void nnd_conv_set_stream_path(int enable, int issuers) { g_stream = enable; nnd_conv_tcs_set_issuers(issuers); }

// Profiling aid: nnd_conv_trace(1) clears the table and starts recording one row per convolution-family launch (two CUDA events
// on the launch stream around it); nnd_conv_trace(0) stops.  nnd_conv_trace_dump synchronises the device and writes the rows as
// CSV (idx,kind,kernel,N,Di,Hi,Wi,Cin,Cout,Ld,Lh,Lw,sd,sh,sw,T,ms,gflop); returns NND_OK or NND_ERR_ARG / NND_ERR_CUDA.
void nnd_conv_trace(int enable) {
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (enable) {
    for (auto& r : g_trace) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_trace.clear();
  }
  g_trace_on = enable ? 1 : 0;
}
long long nnd_conv_trace_count() { return (long long)g_trace.size(); }
int nnd_conv_trace_dump(const char* path) {
  if (!path) return NND_ERR_ARG;
  NND_CUDA_TRY(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_trace_mu);
  FILE* f = fopen(path, "w");
  if (!f) return NND_ERR_ARG;
  fprintf(f, "idx,kind,kernel,N,Di,Hi,Wi,Cin,Cout,Ld,Lh,Lw,sd,sh,sw,T,ms,gflop,t0_ms,stream\n");
  for (size_t i = 0; i < g_trace.size(); ++i) {
    const TraceRec& r = g_trace[i];
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { ms = -1.f; (void)cudaGetLastError(); }
    const double gf = 2.0 * r.N * (double)r.Ld * r.Lh * r.Lw * r.T * r.Cin * r.Cout * 1e-9;
    float t0 = -1.f;                                  // start of the launch relative to the first traced launch: a per-stream timeline
    if (cudaEventElapsedTime(&t0, g_trace[0].e0, r.e0) != cudaSuccess) { t0 = -1.f; (void)cudaGetLastError(); }
    fprintf(f, "%zu,%s,%s,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%.6f,%.4f,%.6f,%llx\n", i, r.kind, r.kernel, r.N, r.Di, r.Hi, r.Wi, r.Cin,
            r.Cout, r.Ld, r.Lh, r.Lw, r.sd, r.sh, r.sw, r.T, ms, gf, t0, (unsigned long long)(size_t)r.st);
  }
  fclose(f);
  return NND_OK;
}

// Dry-run dispatch queries (host only, no CUDA call -- usable without a GPU): which kernel would serve this launch under the current
// switches.  gather: 0 conv_igemm (mma.sync), 1 conv_tc, 2 conv_tcs, 3 conv_tc S2; wgrad: 0 generic, 1 halo (mma.sync), 2 conv_wgrad_tc,
// 3 conv_wgrad_tc32, 4 conv_wgrad_tcn, 5 conv_wgrad_tc SW=2, 6 conv_wgrad_tma, 8 conv_wgrad_tma_s2, 9 conv_wgrad_tma32.  Negative: bad geometry.
int nnd_conv_gather_dispatch(const int* geom, long long out_n_stride, long long out_v_stride, int out_fp32, int Cout, int CoutPad,
                             int has_bias, int has_residual, int has_stats) {
  ConvGeom g;
  if (parse_geom(geom, g) != NND_OK) return -1;
  static float dummy;
  ConvEpilogue ep;
  ep.out = &dummy; ep.out_n_stride = out_n_stride; ep.out_v_stride = out_v_stride; ep.out_fp32 = out_fp32;
  ep.Cout = Cout; ep.CoutPad = CoutPad; ep.bias = has_bias ? &dummy : nullptr; ep.scale = nullptr;
  ep.residual = has_residual ? reinterpret_cast<const __nv_bfloat16*>(&dummy) : nullptr;
  ep.stat_sum = has_stats ? &dummy : nullptr; ep.stat_sq = has_stats ? &dummy : nullptr;
  if (g_force_igemm) return 0;
  if (g_pw && nnd_conv_pw_supported(g, ep)) return 4;
  if (g_stream && nnd_conv_tcs_supported(g, ep) && (g_stream == 2 || nnd_conv_tcs_profitable(g, ep))) return 2;
  if (g_gather_strided && (g_gather_tma & 2) && nnd_conv_tct_s2_supported(g, ep)) return 7;
  if (g_gather_strided && nnd_conv_tc_s2_supported(g, ep)) return 3;
  if ((g_gather_tma & 1) && nnd_conv_tct_supported(g, ep)) return 6;
  return nnd_conv_tc_supported(g, ep) ? 1 : 0;
}

int nnd_conv_wgrad_dispatch(const int* geom, int Cdy, int Cx) {
  ConvGeom g;
  if (parse_geom(geom, g) != NND_OK) return -1;
  if (g_force_igemm) return 0;
  if (g_wgrad_tc && nnd_conv_wgrad_tc32_supported(g, Cdy, Cx) && (g_wgrad_tc == 2 || nnd_conv_wgrad_tc32_profitable(g)))
    return ((g_wgrad_tma & 1) && !(g_wgrad_tma & 256) && nnd_conv_wgrad_tma32_supported(g, Cdy, Cx)) ? 9 : 3;
  if (g_wgrad_tc == 4 && nnd_conv_wgrad_tcn_supported(g, Cdy, Cx)) return 4;
  if (g_wgrad_tc && (g_wgrad_tma & 1) && nnd_conv_wgrad_tma_supported(g, Cdy, Cx)) return 6;
  if (g_wgrad_tc && nnd_conv_wgrad_tc_supported(g, Cdy, Cx)) return 2;
  if (g_wgrad_tc && g_wgrad_strided && (g_wgrad_tma & 1) && !(g_wgrad_tma & 128) && nnd_conv_wgrad_tma_s2_supported(g, Cdy, Cx)) return 8;
  if (g_wgrad_tc && g_wgrad_strided && nnd_conv_wgrad_tc_strided_supported(g, Cdy, Cx)) return 5;
  return nnd_conv_wgrad_halo_supported(g, Cdy, Cx) ? 1 : 0;
}

int nnd_conv_gather_bf16_items(const void* in, const void* w, const int* geom, void* out, long long out_n_stride,
                               long long out_v_stride, int out_fp32, int Cout, int CoutPad, const float* bias, const float* scale,
                               const void* residual, float* stat_sum, float* stat_sq, int* used_tc, cudaStream_t st,
                               const void* w_items, int items_n_tile, int items_T);

int nnd_conv_gather_bf16(const void* in, const void* w, const int* geom, void* out, long long out_n_stride,
                         long long out_v_stride, int out_fp32, int Cout, int CoutPad, const float* bias, const float* scale,
                         const void* residual, float* stat_sum, float* stat_sq, int* used_tc, cudaStream_t st) {
  return nnd_conv_gather_bf16_items(in, w, geom, out, out_n_stride, out_v_stride, out_fp32, Cout, CoutPad, bias, scale, residual,
                                    stat_sum, stat_sq, used_tc, st, nullptr, 0, 0);
}

// Same launch; `w_items` optionally carries the weights a second time, re-packed in pipeline-item order by nnd_repack_items_bf16 for
// tiles of `items_n_tile` rows (`items_T` weight slices): used by the opt-in bulk-copy variant of the tile kernel, ignored otherwise.
int nnd_conv_gather_bf16_items(const void* in, const void* w, const int* geom, void* out, long long out_n_stride,
                               long long out_v_stride, int out_fp32, int Cout, int CoutPad, const float* bias, const float* scale,
                               const void* residual, float* stat_sum, float* stat_sq, int* used_tc, cudaStream_t st,
                               const void* w_items, int items_n_tile, int items_T) {
  ConvGeom g;
  if (parse_geom(geom, g) != NND_OK) return NND_ERR_ARG;
  ConvEpilogue ep;
  ep.out = out; ep.out_n_stride = out_n_stride; ep.out_v_stride = out_v_stride; ep.out_fp32 = out_fp32;
  ep.Cout = Cout; ep.CoutPad = CoutPad; ep.bias = bias; ep.scale = scale;
  ep.residual = (const __nv_bfloat16*)residual; ep.stat_sum = stat_sum; ep.stat_sq = stat_sq;
  if (!g_force_igemm && g_pw && nnd_conv_pw_supported(g, ep)) {
    if (used_tc) *used_tc = 5;
    TraceScope ts("fprop", "conv_pw", g, g.Cin, Cout, st);
    return nnd_conv_pw((const __nv_bfloat16*)in, (const __nv_bfloat16*)w, g, ep, st);
  }
  if (!g_force_igemm && g_stream && nnd_conv_tcs_supported(g, ep) && (g_stream == 2 || nnd_conv_tcs_profitable(g, ep))) {
    if (used_tc) *used_tc = 2;
    TraceScope ts("fprop", "conv_tcs", g, g.Cin, Cout, st);
    return nnd_conv_tcs((const __nv_bfloat16*)in, (const __nv_bfloat16*)w, g, ep, st);
  }
  if (!g_force_igemm && g_gather_strided && (g_gather_tma & 2) && nnd_conv_tct_s2_supported(g, ep)) {
    int r;
    {
      TraceScope ts("fprop", "conv_tct_s2", g, g.Cin, Cout, st);
      r = nnd_conv_tct_s2((const __nv_bfloat16*)in, (const __nv_bfloat16*)w, g, ep, st);
    }
    if (r != NND_ERR_ARG) { if (used_tc) *used_tc = 7; return r; }       // NND_ERR_ARG: no tensor map / shared memory -> cp.async kernel
  }
  if (!g_force_igemm && g_gather_strided && nnd_conv_tc_s2_supported(g, ep)) {
    if (used_tc) *used_tc = 3;
    TraceScope ts("fprop", "conv_tc_s2", g, g.Cin, Cout, st);
    return nnd_conv_tc_s2((const __nv_bfloat16*)in, (const __nv_bfloat16*)w, g, ep, st);
  }
  if (!g_force_igemm && g_tc_bulk && nnd_conv_tc_bulk_supported(g, ep, w_items, items_n_tile, items_T)) {
    if (used_tc) *used_tc = 4;
    TraceScope ts("fprop", "conv_tc_bulk", g, g.Cin, Cout, st);
    return nnd_conv_tc_bulk((const __nv_bfloat16*)in, (const __nv_bfloat16*)w, g, ep, st, (const __nv_bfloat16*)w_items, items_n_tile, items_T);
  }
  if (!g_force_igemm && (g_gather_tma & 1) && nnd_conv_tct_supported(g, ep)) {
    int r;
    {
      TraceScope ts("fprop", "conv_tct", g, g.Cin, Cout, st);
      r = nnd_conv_tct((const __nv_bfloat16*)in, (const __nv_bfloat16*)w, g, ep, st);
    }
    if (r != NND_ERR_ARG) { if (used_tc) *used_tc = 6; return r; }
  }
  const bool tc = !g_force_igemm && nnd_conv_tc_supported(g, ep);
  if (used_tc) *used_tc = tc ? 1 : 0;
  TraceScope ts("fprop", tc ? "conv_tc" : "conv_igemm", g, g.Cin, Cout, st);
  if (tc) return nnd_conv_tc((const __nv_bfloat16*)in, (const __nv_bfloat16*)w, g, ep, st);
  return nnd_conv_igemm((const __nv_bfloat16*)in, (const __nv_bfloat16*)w, g, ep, st);
}

// conv_pw.cu: a whole kernel == stride transposed convolution in one launch (see include/nndet_b200.h)
int nnd_conv_upconv_bf16(const void* x, const void* w_packed, int N, int D, int H, int W, int Cin, int Cout, int sd, int sh, int sw,
                         void* out, const float* bias, const void* residual, cudaStream_t st) {
  ConvGeom g = {};
  g.N = N; g.Di = D; g.Hi = H; g.Wi = W; g.Cin = Cin; g.Ld = D; g.Lh = H; g.Lw = W; g.sd = sd; g.sh = sh; g.sw = sw;
  g.T = sd * sh * sw;                                   // trace row: taps x input voxels (the dump's FLOP formula)
  TraceScope ts("fprop", "conv_pw_up", g, Cin, Cout, st);
  return nnd_conv_upconv(x, w_packed, N, D, H, W, Cin, Cout, sd, sh, sw, out, bias, residual, st);
}

// Workspace the weight-gradient launch can use (split-K partials of the TMA-fed kernel: plain stores + one finishing pass instead of
// splits x as many atomics into dW); 0 when the dispatch takes a kernel that needs none.  Host only, no CUDA call.
long long nnd_conv_wgrad_workspace_bytes(const int* geom, int Cdy, int Cx, int Cout, int Cin) {
  ConvGeom g;
  if (parse_geom(geom, g) != NND_OK) return 0;
  const int code = nnd_conv_wgrad_dispatch(geom, Cdy, Cx);
  if (code == 8) return nnd_conv_wgrad_tma_s2_workspace(g, Cdy, Cx, Cout, Cin);
  if (code == 9) return nnd_conv_wgrad_tma32_workspace(g, Cout, Cin);
  if (code != 6) return 0;
  return nnd_conv_wgrad_tma_workspace(g, Cdy, Cx, Cout, Cin);
}

int nnd_conv_wgrad_bf16_ws(const void* dy, int Cdy, const void* x, int Cx, const int* geom, float* dw, long long s_co,
                           long long s_ci, long long s_tap, int Cout, int Cin, void* ws, long long ws_bytes, cudaStream_t st);

int nnd_conv_wgrad_bf16(const void* dy, int Cdy, const void* x, int Cx, const int* geom, float* dw, long long s_co,
                        long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t st) {
  return nnd_conv_wgrad_bf16_ws(dy, Cdy, x, Cx, geom, dw, s_co, s_ci, s_tap, Cout, Cin, nullptr, 0, st);
}

// Same launch with an optional caller-provided workspace of nnd_conv_wgrad_workspace_bytes (smaller / null: the atomics path).
int nnd_conv_wgrad_bf16_ws(const void* dy, int Cdy, const void* x, int Cx, const int* geom, float* dw, long long s_co,
                           long long s_ci, long long s_tap, int Cout, int Cin, void* ws, long long ws_bytes, cudaStream_t st) {
  ConvGeom g;
  if (parse_geom(geom, g) != NND_OK) return NND_ERR_ARG;
  if (!g_force_igemm && g_wgrad_tc && nnd_conv_wgrad_tc32_supported(g, Cdy, Cx) && (g_wgrad_tc == 2 || nnd_conv_wgrad_tc32_profitable(g))) {
    if ((g_wgrad_tma & 1) && !(g_wgrad_tma & 256) && nnd_conv_wgrad_tma32_supported(g, Cdy, Cx)) {
      int r;
      {
        TraceScope ts("wgrad", "wgrad_tma32", g, Cx, Cdy, st);
        r = nnd_conv_wgrad_tma32((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, g, dw, s_co, s_ci, s_tap, Cout, Cin,
                                 (g_wgrad_tma & 64) ? nullptr : ws, ws_bytes, st);
      }
      if (r != NND_ERR_ARG) return r;          // NND_ERR_ARG: no tensor map -> the cp.async kernel below
    }
    TraceScope ts("wgrad", "wgrad_tc32", g, Cx, Cdy, st);
    return nnd_conv_wgrad_tc32((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, g, dw, s_co, s_ci, s_tap, Cout, Cin, st);
  }
  // g_wgrad_tc: 1 = tcgen05 wgrad kernels (default), 2 = also the stacked 32-channel one on small volumes (tests), 4 = the
  // all-taps 128-output-channel kernel of conv_wgrad_tcn.cu (opt-in: measured SLOWER than the filter-row kernel, 0.38 vs
  // 0.23 ms at 32^3 x 4 -- its N = 48 MMAs cost 44 cycles for 24 cycles of math; kept as the A/B record of that experiment)
  if (!g_force_igemm && g_wgrad_tc && g_wgrad_tc != 4 && (g_wgrad_tma & 1) && nnd_conv_wgrad_tma_supported(g, Cdy, Cx)) {
    int r;
    {
      TraceScope ts("wgrad", "wgrad_tma", g, Cx, Cdy, st);
      r = nnd_conv_wgrad_tma((const __nv_bfloat16*)dy, Cdy, (const __nv_bfloat16*)x, Cx, g, dw, s_co, s_ci, s_tap, Cout, Cin, g_wgrad_tma,
                             (g_wgrad_tma & 64) ? nullptr : ws, ws_bytes, st);
    }
    if (r != NND_ERR_ARG) return r;          // NND_ERR_ARG: no tensor map for this shape -> the cp.async kernels below
  }
  if (!g_force_igemm && g_wgrad_tc && g_wgrad_tc != 4 && g_wgrad_strided && (g_wgrad_tma & 1) && !(g_wgrad_tma & 128) &&
      !nnd_conv_wgrad_tc_supported(g, Cdy, Cx) && nnd_conv_wgrad_tma_s2_supported(g, Cdy, Cx)) {
    int r;
    {
      TraceScope ts("wgrad", "wgrad_tma_s2", g, Cx, Cdy, st);
      r = nnd_conv_wgrad_tma_s2((const __nv_bfloat16*)dy, Cdy, (const __nv_bfloat16*)x, Cx, g, dw, s_co, s_ci, s_tap, Cout, Cin, g_wgrad_tma,
                                (g_wgrad_tma & 64) ? nullptr : ws, ws_bytes, st);
    }
    if (r != NND_ERR_ARG) return r;
  }
  const char* wk = "wgrad_generic";
  if (!g_force_igemm && g_wgrad_tc == 4 && nnd_conv_wgrad_tcn_supported(g, Cdy, Cx)) wk = "wgrad_tcn";
  else if (!g_force_igemm && g_wgrad_tc && nnd_conv_wgrad_tc_supported(g, Cdy, Cx)) wk = "wgrad_tc";
  else if (!g_force_igemm && g_wgrad_tc && g_wgrad_strided && nnd_conv_wgrad_tc_strided_supported(g, Cdy, Cx)) wk = "wgrad_tc_s2";
  else if (!g_force_igemm && nnd_conv_wgrad_halo_supported(g, Cdy, Cx)) wk = "wgrad_halo";
  TraceScope ts("wgrad", wk, g, Cx, Cdy, st);
  if (!g_force_igemm && g_wgrad_tc == 4 && nnd_conv_wgrad_tcn_supported(g, Cdy, Cx))
    return nnd_conv_wgrad_tcn((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, Cx, g, dw, s_co, s_ci, s_tap, Cout, Cin, st);
  if (!g_force_igemm && g_wgrad_tc && nnd_conv_wgrad_tc_supported(g, Cdy, Cx))
    return nnd_conv_wgrad_tc((const __nv_bfloat16*)dy, Cdy, (const __nv_bfloat16*)x, Cx, g, dw, s_co, s_ci, s_tap, Cout, Cin, st);
  if (!g_force_igemm && g_wgrad_tc && g_wgrad_strided && nnd_conv_wgrad_tc_strided_supported(g, Cdy, Cx))
    return nnd_conv_wgrad_tc((const __nv_bfloat16*)dy, Cdy, (const __nv_bfloat16*)x, Cx, g, dw, s_co, s_ci, s_tap, Cout, Cin, st);
  if (!g_force_igemm && nnd_conv_wgrad_halo_supported(g, Cdy, Cx))
    return nnd_conv_wgrad_halo((const __nv_bfloat16*)dy, Cdy, (const __nv_bfloat16*)x, Cx, g, dw, s_co, s_ci, s_tap, Cout, Cin, st);
  return nnd_conv_wgrad((const __nv_bfloat16*)dy, Cdy, (const __nv_bfloat16*)x, Cx, g, dw, s_co, s_ci, s_tap, Cout, Cin, st);
}

int nnd_conv_first_fprop_f32(const float* x, const float* w, const int* geom, int Cout, void* out, float* stat_sum,
                             float* stat_sq, cudaStream_t st) {
  ConvGeom g;
  if (parse_geom(geom, g) != NND_OK) return NND_ERR_ARG;
  TraceScope ts("first_fprop", "conv_first", g, g.Cin, Cout, st);
  return nnd_conv_first_fprop(x, w, g, Cout, (__nv_bfloat16*)out, stat_sum, stat_sq, st);
}

int nnd_conv_first_wgrad_f32(const float* x, const void* dy, const int* geom, int Cout, float* dw, cudaStream_t st) {
  ConvGeom g;
  if (parse_geom(geom, g) != NND_OK) return NND_ERR_ARG;
  TraceScope ts("first_wgrad", "conv_first", g, g.Cin, Cout, st);
  return nnd_conv_first_wgrad(x, (const __nv_bfloat16*)dy, g, Cout, dw, st);
}

}  // extern "C"
