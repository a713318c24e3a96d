/* nndet_b200.h -- C ABI of libnndet_b200.so: the B200 (sm_100a) hot path of nnDetection.
 *
 * Every entry point: plain pointers + sizes, explicit cudaStream_t, caller-provided workspace (so it can come from
 * the host framework's allocator), returns an int status (0 ok / 1 bad argument / 2 workspace too small / 3 CUDA
 * error -> nnd_last_error()), never throws, never synchronises with the host.  All pointers are DEVICE pointers
 * unless named *_host.  Reference citations are file:line under MIC-DKFZ/nnDetection.
 *
 * The reference's only native interface on this path is the pybind symbol
 *     nndet._C.nms(dets, scores, iou_threshold) -> int64 indices      (nndet/csrc/ops.cpp:13-15)
 * which nnd_nms3d_f32 / nnd_nms2d_f32 replace; the other families replace the torch/ATen/cuDNN calls the
 * reference's Python issues for the same step (cited per function).  INTEGRATION.md shows the bindings.
 */
#ifndef NNDET_B200_H
#define NNDET_B200_H
#include <stddef.h>
#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NND_OK 0
#define NND_ERR_ARG 1
#define NND_ERR_WORKSPACE 2
#define NND_ERR_CUDA 3

int nnd_abi_version(void);
const char* nnd_last_error(void);              /* text of the last NND_ERR_CUDA on this thread */
const char* nnd_build_arch(void);              /* "sm_100a" */
unsigned long long nnd_launch_count(void);     /* kernels launched through this library so far */

/* ---- 3-D / 2-D greedy NMS.  Replaces nms_cuda + nms_kernel(_3d) + devIoU(_3d), nndet/csrc/cuda/nms.cu:148-221,54-145,22-51,
 *      and the dispatch in nndet/csrc/cpu/nms.cpp:18-34.  boxes [n, 6|4] (x1,y1,x2,y2[,z1,z2]) fp32, scores [n] fp32.
 *      keep_out [n] int64: indices into `boxes` by descending score; n_keep_out: device scalar. */
size_t nnd_nms_workspace_bytes(long long n, int dim /* 2 or 3 */);
int nnd_nms3d_f32(const float* boxes, const float* scores, long long n, float iou_threshold, long long* keep_out,
                  long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream);
int nnd_nms2d_f32(const float* boxes, const float* scores, long long n, float iou_threshold, long long* keep_out,
                  long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream);
/* prefix variant for `keep[:detections_per_img]` (nndet/core/retina.py:376-378): the scan stops once max_keep boxes survived;
 * keep_out[0 .. min(*n_keep_out, max_keep)) equals the prefix of the full result. */
int nnd_nms3d_topk_f32(const float* boxes, const float* scores, long long n, float iou_threshold, long long max_keep,
                       long long* keep_out, long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream);

/* ---- weighted box clustering.  Replaces wbc + compute_cluster_consolidation, nndet/inference/detection/wbc.py:94-198
 *      (N x N IoU matrix + python while loop with torch.where per cluster).  boxes [n, 6], scores / weights / n_exp_preds [n];
 *      clusters in descending head-score order; out_boxes [n, 6], out_scores [n], n_out device scalar (clusters with
 *      score > score_thresh). */
size_t nnd_wbc_workspace_bytes(long long n);
int nnd_wbc3d_f32(const float* boxes, const float* scores, const float* weights, const float* n_exp_preds, long long n,
                  float iou_thresh, float score_thresh, int use_area, float missing_weight, float* out_boxes,
                  float* out_scores, long long* n_out, void* ws, size_t ws_bytes, cudaStream_t stream);

/* ---- instance segmentation -> detection targets.  Replaces the per-step `pre_trafo` chain FindInstances /
 *      Instances2Boxes / Instances2Segmentation, nndet/io/transforms/instances.py:25-39,42-136,211-301 (invoked at
 *      nndet/ptmodule/retinaunet/base.py:140-141).  target / sem_out fp32 [B, D, H, W]; lut_sem / lut_class int32 [B, max_id]
 *      (label written to the semantic map resp. class of the instance, -1 = id not in the mapping); outputs padded to `cap`
 *      instances per sample, ids ascending; err_out bit 0: id >= max_id or non-integer, bit 1: id missing in the mapping. */
size_t nnd_instances_workspace_bytes(int B, int max_id);
int nnd_instances_to_targets(const float* target, int B, int D, int H, int W, const int* lut_sem, const int* lut_class,
                             int max_id, int cap, float* sem_out, int* out_ids, float* out_boxes, long long* out_classes,
                             int* counts, int* err_out, void* ws, size_t ws_bytes, cudaStream_t stream);

/* ---- anchors: one pyramid level of AnchorGenerator3D.grid_anchors, nndet/core/boxes/anchors.py:337-377.
 *      out [s0*s1*s2*nb, 6]; base [nb, 6] (generate_anchors, anchors.py:526-549); size3/stride3 host ints. */
int nnd_anchor_grid_f32(float* out, const float* base, int nb, const int* size3_host, const int* stride3_host, cudaStream_t stream);

/* ---- pairwise metrics [n, m]: mode 0 box_iou (nndet/core/boxes/ops.py:131-159), 1 generalized_box_iou (ops.py:162-185),
 *      2 box_center_dist (ops.py:262-287). */
int nnd_box_pairwise_f32(const float* b1, const float* b2, int n, int m, float eps, int mode, float* out, cudaStream_t stream);

/* ---- decode_single (nndet/core/boxes/coder.py:90-155, weights 1) [+ clip_boxes_to_image_3d_ (clip.py:83-101)].
 *      deltas/out [n, 6]; anchors [A, 6] reused cyclically (n = batch * A); clip_shape3_host NULL = no clip. */
int nnd_decode_boxes_f32(const float* deltas, const float* anchors, long long n, long long A, float xform_clip,
                         const float* clip_shape3_host, float* out, cudaStream_t stream);
/* ---- sigmoid of logits [n, C] -> probs_out [n, C] (optional) and max over classes fg_out [n] (optional); comb.py:262-263 */
int nnd_sigmoid_fg_f32(const float* logits, long long n, int C, float* probs_out, float* fg_out, cudaStream_t stream);

/* ---- ATSS matching of a whole batch.  Replaces ATSSMatcher.compute_matches (nndet/core/boxes/matcher/atss.py:48-122) and the
 *      per-image loop of assign_targets_to_anchors (nndet/core/retina.py:256-262).  gt [G, 6] (all images), gt_img[g] image of
 *      box g, gt_local[g] its index inside that image; anchors [A, 6]; level_off [L+1] anchor offsets of the pyramid levels
 *      (device + host copy); kc = num_candidates * anchors_per_location; matches_out [B*A] int64 (-1 = background). */
size_t nnd_atss_workspace_bytes(int G, int A, const int* level_off_host, int L, int kc);
int nnd_atss_match(const float* gt, const int* gt_img, const int* gt_local, int G, const float* anchors, int A, int B,
                   const int* level_off_dev, const int* level_off_host, int L, int kc, long long* matches_out, void* ws,
                   size_t ws_bytes, cudaStream_t stream);
/* labels [n] fp32: 0 background, class+1 foreground, -1 ignore (retina.py:263-288); gt_off[b] first gt of image b */
int nnd_assign_labels(const long long* matches, long long n, long long A, const long long* gt_classes, const int* gt_off,
                      float* labels_out, cudaStream_t stream);

/* ---- hard-negative sampler.  Replaces HardNegativeSamplerBatched.__call__ (nndet/core/boxes/sampler.py:212-270).
 *      counts_out int32[8]: #positive, #negative, num_pos, num_neg, pool, pool filled, pos-list overflow, -.
 *      pos_out / neg_out: ascending anchor indices (int64); *pool_list_out points into ws (int32[pool], unordered). */
int nnd_hnm_max_select(void);
size_t nnd_hnm_workspace_bytes(long long n, int pos_cap, int pool_cap);
int nnd_hnm_sample(const float* labels, const float* fg_probs, long long n, int max_pos, double neg_ratio, int min_neg,
                   double pool_size, unsigned int seed, int* counts_out, long long* pos_out, long long* neg_out, int pos_cap,
                   int pool_cap, int** pool_list_out, void* ws, size_t ws_bytes, cudaStream_t stream);

/* ---- head losses.  Replaces DetectionHeadHNMNative.compute_loss (nndet/arch/heads/comb.py:383-405): decode sampled positives,
 *      GIoU loss (nndet/losses/regression.py:118-162) / max(1, P), BCE-with-logits one-hot mean (classification.py:137-181).
 *      losses_out[2] = {reg, cls}; compact gradients g_deltas [max_pos, 6], g_logits [max_pos + max_neg, C]. */
int nnd_head_loss_fwd(const float* logits, const float* deltas, const float* anchors, long long A, int C, const long long* matches,
                      const float* gt_boxes, const int* gt_off, const float* labels, const long long* pos_idx,
                      const long long* neg_idx, const int* counts, float xform_clip, float giou_eps, float* losses_out,
                      float* g_deltas, float* g_logits, cudaStream_t stream);
int nnd_head_loss_bwd(const float* g_deltas, const float* g_logits, int C, const long long* pos_idx, const long long* neg_idx,
                      const int* counts, const float* up_reg, const float* up_cls, float* d_deltas /* zeroed [n,6] */,
                      float* d_logits /* zeroed [n,C] */, cudaStream_t stream);

/* ---- detection post-processing per image.  Replaces postprocess_detections_single_image (nndet/core/retina.py:332-379) +
 *      batched_nms (nndet/core/boxes/nms.py:81-106).  boxes [B*A, 6] decoded+clipped, probs [B*A*C]; out_* [B, det, ...]. */
size_t nnd_detect_postprocess_workspace_bytes(long long A, int C, int topk);
int nnd_detect_postprocess(const float* boxes, const float* probs, int B, long long A, int C, int topk, float score_thresh,
                           int use_score_thresh, float min_size, int use_min_size, float nms_thresh, int det_per_img,
                           float* out_boxes, float* out_scores, long long* out_labels, int* out_counts, void* ws,
                           size_t ws_bytes, cudaStream_t stream);

/* ---- convolutions.  Replace torch.nn.Conv3d / ConvTranspose3d (cuDNN) as built by nd_conv, nndet/arch/conv.py:297-348, and their
 *      autograd.  Activations NDHWC bf16; geom_host: int[21 + 4*T] describing one "gather convolution" launch
 *      (csrc/conv_common.cuh): N,Di,Hi,Wi,Cin, Ld,Lh,Lw, sd,sh,sw, Do,Ho,Wo, omd,omh,omw, ood,ooh,oow, T, then per tap
 *      (off_d, off_h, off_w, weight_tap).  w: bf16 [T][CoutPad][Cin] from nnd_pack_weights.  Epilogue: +bias, +residual,
 *      *scale, optional fp32 output with sample / voxel strides (writes the [N, anchors, C] head layout directly),
 *      per-(sample, channel) sum / sum-of-squares for the following norm.  used_tc_host: 0 mma.sync kernel, 1 tcgen05 tile kernel, 2 tcgen05 streaming kernel, 3 tcgen05 stride-2 tile kernel, 4 tcgen05 tile kernel with bulk-copied weights (do not use), 5 TMA-fed pointwise tcgen05 GEMM. */
void nnd_conv_set_tensor_path(int enable_tcgen05);
void nnd_conv_set_wgrad_tc(int mode);                /* A/B switch: 0 mma.sync wgrad, 1 tcgen05 (default), 2 + stacked 32-ch kernel on small volumes, 4 + all-taps 128-co kernel */
void nnd_conv_set_wgrad_strided_tc(int enable);       /* default 1 (validated on B200, round 2): de-interleaved tcgen05 wgrad for stride-2 / transposed convolutions; 0 = mma.sync (A/B) */
void nnd_conv_set_wgrad_tma(int mode);                /* TMA-fed tcgen05 wgrad (csrc/conv_wgrad_tma.cu) for stride-1 3x3x3 / 1x3x3 layers with channels in multiples of 64: bit 0 on, bit 1 base_offset descriptors (WRONG results: the device-verified model is base_offset 0), bit 2 unstacked taps, bits 3-5 timing experiments, bit 6 ignore the workspace (A/B) */
void nnd_conv_set_gather_tma(int mode);               /* TMA-fed tcgen05 tile kernel (csrc/conv_tct.cu): bit 0 the stride-1 forms of the tile kernel, bit 1 the stride-2 forms (A/B; used_tc 6 / 7) */
void nnd_conv_set_gather_strided_tc(int enable);      /* default 1 (validated on B200, round 2): de-interleaved-halo tcgen05 kernel for stride-2 gathers; 0 = mma.sync (A/B) */
void nnd_conv_set_tcs_map(int mode);                  /* A/B switch: halo copy lane mapping of the streaming kernel (0 row-walking threads, 1 lanes along a voxel's channel groups) */
void nnd_conv_set_stream_path(int enable, int issuers); /* A/B switch: streaming z-window tcgen05 kernel (default on, 2 issuers) */
/* Profiling aid (off by default): one row per convolution-family launch -- kind (fprop | wgrad | first_*), the kernel the dispatch
 * chose, the geometry, and the launch's duration from two CUDA events on its stream.  trace(1) clears + starts, trace(0) stops;
 * trace_dump synchronises the device and writes CSV (idx,kind,kernel,N,Di,Hi,Wi,Cin,Cout,Ld,Lh,Lw,sd,sh,sw,T,ms,gflop,t0_ms,stream: t0_ms = start relative to the first traced launch, i.e. a per-stream timeline). */
void nnd_conv_trace(int enable);
long long nnd_conv_trace_count(void);
int nnd_conv_trace_dump(const char* path);
/* Dry-run dispatch queries (host only, no CUDA call): the kernel that would serve a launch under the current switches.
 * gather: 0 conv_igemm (mma.sync), 1 conv_tc, 2 conv_tcs, 3 conv_tc S2, 4 conv_pw (TMA); wgrad: 0 generic, 1 halo (mma.sync), 2 conv_wgrad_tc,
 * 3 conv_wgrad_tc32, 4 conv_wgrad_tcn, 5 conv_wgrad_tc SW=2; negative: bad geometry. */
int nnd_conv_gather_dispatch(const int* geom_host, long long out_n_stride, long long out_v_stride, int out_fp32, int Cout, int CoutPad,
                             int has_bias, int has_residual, int has_stats);
int nnd_conv_wgrad_dispatch(const int* geom_host, int Cdy, int Cx);
int nnd_conv_gather_bf16(const void* in, const void* w, const int* geom_host, void* out, long long out_n_stride,
                         long long out_v_stride, int out_fp32, int Cout, int CoutPad, const float* bias, const float* scale,
                         const void* residual, float* stat_sum, float* stat_sq, int* used_tc_host, cudaStream_t stream);
/* DO NOT ENABLE (deadlocks on the device; kept as the record of the experiment): the tile kernel's weight stream through cp.async.bulk.  The caller re-packs a
 * K-major weight pack [T][rows_pad][K] into pipeline-item order [row tile][k chunk][T][k group][n][8] (kg = 4: 32-channel chunks)
 * and passes it along with the ordinary pack; used only while nnd_conv_set_tc_bulk(1). */
void nnd_conv_set_tc_bulk(int enable);
int nnd_repack_items_bf16(const void* src, int T, int rows_pad, int K, int n_tile, int kg, void* dst, cudaStream_t stream);
int nnd_conv_gather_bf16_items(const void* in, const void* w, const int* geom_host, void* out, long long out_n_stride,
                               long long out_v_stride, int out_fp32, int Cout, int CoutPad, const float* bias, const float* scale,
                               const void* residual, float* stat_sum, float* stat_sq, int* used_tc_host, cudaStream_t stream,
                               const void* w_items, int items_n_tile, int items_T);
/* dW[co*s_co + ci*s_ci + tap*s_tap] += sum_voxels dy[.., co] * x[.., ci]   (fp32, caller zero-fills) */
int nnd_conv_wgrad_bf16(const void* dy, int Cdy, const void* x, int Cx, const int* geom_host, float* dw, long long s_co,
                        long long s_ci, long long s_tap, int Cout, int Cin, cudaStream_t stream);
/* Same launch with an optional workspace (nnd_conv_wgrad_workspace_bytes, host-only query; 0 = none needed): the TMA-fed kernel then
 * stores its split-K partials with plain 16-byte stores and a finishing pass adds their sum into dw -- taps x Cout x Cin atomics per
 * launch instead of splits x as many.  ws null / too small: the atomics path. */
long long nnd_conv_wgrad_workspace_bytes(const int* geom_host, int Cdy, int Cx, int Cout, int Cin);
int nnd_conv_wgrad_bf16_ws(const void* dy, int Cdy, const void* x, int Cx, const int* geom_host, float* dw, long long s_co,
                           long long s_ci, long long s_tap, int Cout, int Cin, void* ws, long long ws_bytes, cudaStream_t stream);
/* image-input layer (Cin <= 4): x fp32 NCDHW, w fp32 [Cout][Cin][T] */
int nnd_conv_first_fprop_f32(const float* x, const float* w, const int* geom_host, int Cout, void* out, float* stat_sum,
                             float* stat_sq, cudaStream_t stream);
int nnd_conv_first_wgrad_f32(const float* x, const void* dy, const int* geom_host, int Cout, float* dw, cudaStream_t stream);
void nnd_conv_set_tc_ring(int deep);                 /* A/B switch: deeper weight ring in the 128-channel tile kernel (default 1) */
void nnd_conv_set_first_layer_mma(int enable);       /* A/B switch: tensor-core image layer (default) vs scalar kernels */
/* fp32 master weight ([Cout][Cin][T], or [Cin][Cout][T] if transposed) -> bf16 fprop / dgrad operands */
int nnd_pack_weights(const float* w, int Cout, int Cin, int T, int transposed, void* fwd, int CoutPadF, int CinPadF, void* bwd,
                     int CinPadB, int CoutPadB, cudaStream_t stream);

/* ---- instance / group norm (+affine, +ReLU).  Replace nn.InstanceNorm3d / GroupNorm / ReLU (nndet/arch/conv.py:388-446,
 *      nndet/arch/layers/norm.py:26-50).  stats from the conv epilogue; a, b, mean, rstd: [N, C] fp32. */
int nnd_norm_finalize(const float* ssum, const float* ssq, const float* gamma, const float* beta, int N, int C, int cpg,
                      long long count, float eps, float* a, float* b, float* mean, float* rstd, cudaStream_t stream);
int nnd_norm_apply(const void* y, const float* a, const float* b, int N, long long V, int C, int relu, void* z, cudaStream_t stream);
void nnd_norm_set_bwd_narrow(int enable);             /* opt-in (default 0): 4-channels-per-thread backward passes, not yet measured on a device */
int nnd_norm_backward(const void* dz, const void* y, const float* a, const float* b, const float* mean, const float* rstd,
                      const float* gamma, int N, long long V, int C, int cpg, int relu, void* dy, float* dgamma, float* dbeta,
                      float* ws /* 5*N*C floats */, cudaStream_t stream);

/* ---- segmentation head.  Replace DiCESegmenterFgBg (nndet/arch/heads/segmenter.py:223-290) + SoftDiceLoss / CE
 *      (nndet/losses/segmentation.py:84-151).  x bf16 [total, C]; logits fp32 [total, 2]; target fp32 [total]. */
int nnd_seg_conv_fwd(const void* x, int C, const float* w, const float* bias, long long total, float* logits, cudaStream_t stream);
int nnd_seg_loss_fwd(const float* logits, const float* target, long long total, float alpha, float smooth, double* sums4,
                     float* losses_out2, cudaStream_t stream);
int nnd_seg_loss_bwd(const float* logits, const float* target, const double* sums4, long long total, float alpha, float smooth,
                     const float* up_ce, const float* up_dice, float* dlogits, cudaStream_t stream);
int nnd_seg_conv_bwd(const void* x, int C, const float* w, const float* dlogits, long long total, void* dx, float* dw, float* db,
                     cudaStream_t stream);

/* ---- host-only plan queries of the TMA-fed kernels (no CUDA call; tests/test_tma_plan_cpu.py replays them in numpy): flat int tables,
 *      return = ints written or -1.  conv_tct: [n_planes, a_stage_bytes, a_bytes, ROWB, per plane pd,ph,pw,cd,ch,cw,base,box_w,box_h,box_d,
 *      per tap tap_off,tap_sbo,slice_step]; wgrad: [narrow,bw,bh,pair,nb,(cb,)HB,WS,n_groups,ci_tiles,splits,units_per_split,total_units,
 *      (dxmask,) per group dz,ty,gtw[3],gtw2[3]]. */
int nnd_conv_tct_plan_debug(const int* geom_host, int mt, int s2, int* out, int cap);
int nnd_conv_wgrad_tma_plan_debug(const int* geom_host, int Cdy, int Cx, int* out, int cap);
int nnd_conv_wgrad_tma_s2_plan_debug(const int* geom_host, int Cdy, int Cx, int* out, int cap);

/* ---- TMA-fed pointwise GEMM (csrc/conv_pw.cu).  nnd_conv_set_pointwise_tma(1): single-tap launches of nnd_conv_gather_bf16 (1x1x1
 *      convolutions = U-FPN laterals, nndet/arch/decoder/base.py:216-241, their dgrad, parity classes of up-convolutions) take it.
 *      nnd_conv_upconv_bf16: a whole kernel == stride nn.ConvTranspose3d (decoder/base.py:272-304) in one launch; x bf16 [N,D,H,W,Cin],
 *      w_packed = fprop pack of nnd_pack_weights ([taps][Cout][Cin], taps (a,b,c) row-major), out / residual bf16
 *      [N, D*sd, H*sh, W*sw, Cout]; Cin, Cout multiples of 32, strides in {1, 2}. */
void nnd_conv_set_pointwise_tma(int enable);
int nnd_conv_upconv_bf16(const void* x, const void* w_packed, int N, int D, int H, int W, int Cin, int Cout, int sd, int sh, int sw,
                         void* out, const float* bias, const void* residual, cudaStream_t stream);

/* ---- optimizer + small streaming helpers.  nnd_sgd_step == torch.optim.SGD(momentum, nesterov, weight_decay) as configured in
 *      nndet/ptmodule/retinaunet/base.py:300-336; elements >= n_decay get no weight decay (norm params). */
int nnd_sgd_step(float* p, const float* g, float* mom, long long n, long long n_decay, float lr, float momentum, float wd,
                 int nesterov, int first_step, float grad_scale, cudaStream_t stream);
/* same, leaving n_skip [lo, hi) element ranges (device int64 pairs) untouched: parameters that never receive a gradient, which
 * torch.optim.SGD skips (`p.grad is None`) -- no weight decay, no momentum for them. */
int nnd_sgd_step_skip(float* p, const float* g, float* mom, long long n, long long n_decay, float lr, float momentum, float wd,
                      int nesterov, int first_step, float grad_scale, const long long* skip, int n_skip, cudaStream_t stream);
int nnd_pad_cast_f32_bf16(const float* src, int N, long long rows, int C, long long src_n_stride, const float* mul, void* dst,
                          int Cpad, cudaStream_t stream);
int nnd_channel_sum(const void* src, int is_bf16, long long rows, int C, long long stride, float scale, float* out, cudaStream_t stream);
int nnd_cast_f32_bf16(const float* s, void* d, long long n, cudaStream_t stream);
int nnd_scale_grad(const float* g, const float* out, int N, long long len, long long n_stride, const float* scale, float* dscale,
                   cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NNDET_B200_H */
