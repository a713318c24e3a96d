// Detection post-processing on device: top-k candidate selection, score / size filters, per-class NMS, top-N.
//
// Reference: BaseRetinaNet.postprocess_detections_single_image, nndet/core/retina.py:332-379:
//   clip (done by nnd_decode_boxes_f32) -> FULL descending sort of A*C probabilities (:358) -> top `topk_candidates`
//   -> prob > score_thresh -> anchor = idx // C, label = idx % C -> remove_small_boxes (ops.py:241-259)
//   -> batched_nms with the coordinate-offset trick (nms.py:81-106) -> first `detections_per_img`.
// Here the full sort of 1-3 M scores is replaced by a radix *selection* of the k largest unique
// (probability, index) keys followed by a sort of only those k; ties -> ascending flat index (canonical).
// No host synchronisation: padding entries (score -inf, NaN boxes) keep every buffer a fixed size.
#include "common.cuh"
#include "select.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <math_constants.h>

extern "C" size_t nnd_nms_workspace_bytes(long long n, int dim);
extern "C" int nnd_nms3d_topk_f32(const float* boxes, const float* scores, long long n, float iou_threshold, long long max_keep,
                                  long long* keep_out, long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream);
extern "C" int nnd_nms3d_f32(const float* boxes, const float* scores, long long n, float iou_threshold,
                             long long* keep_out, long long* n_keep_out, void* ws, size_t ws_bytes, cudaStream_t stream);

namespace {

__global__ void topk_init_kernel(SelState* st, unsigned int* ghist, int* counter, int k, long long n) {
  if (threadIdx.x < 256) ghist[threadIdx.x] = 0;
  if (threadIdx.x != 0) return;
  *counter = 0;
  st->prefix = 0ull; st->need = k; st->shift = 56;
  st->done = (k >= n) ? 1 : 0;
  st->T = ~0ull;
}

__global__ void topk_collect_kernel(const float* __restrict__ vals, long long n, const SelState* __restrict__ st, int k,
                                    unsigned long long* __restrict__ keys_out, int* __restrict__ counter) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool take = false;
  unsigned long long key = 0ull;
  if (i < n) {
    key = neg_key(vals[i], (unsigned int)i);
    take = (key >> st->shift) <= st->T;
  }
  unsigned m = __ballot_sync(0xffffffffu, take);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (take) {
    int p = base + __popc(m & ((1u << lane) - 1u));
    if (p < k) keys_out[p] = ~key;          // (prob bits << 32) | (0xFFFFFFFF - flat index): descending sort order
  }
}

// Single CTA: ordered filter + compaction of the sorted candidates of one image.
__global__ void __launch_bounds__(1024)
post_filter_kernel(const unsigned long long* __restrict__ keys, int k, const float* __restrict__ boxes_img, int C,
                   float score_thresh, int use_thresh, float min_size, int use_min, float* __restrict__ cand_boxes,
                   float* __restrict__ nms_boxes, float* __restrict__ scores, long long* __restrict__ labels,
                   int* __restrict__ m_out) {
  __shared__ int warp_cnt[32];
  __shared__ float warp_max[32];
  __shared__ int s_m;
  __shared__ float s_maxc;
  if (threadIdx.x == 0) s_m = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float tmax = -CUDART_INF_F;
  for (int base = 0; base < k; base += 1024) {
    const int j = base + threadIdx.x;
    bool valid = false;
    float b[6], prob = 0.f;
    long long lab = 0;
    if (j < k) {
      const unsigned long long key = keys[j];
      prob = __uint_as_float((unsigned int)(key >> 32));
      const unsigned int flat = 0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull);
      const unsigned int a = flat / (unsigned int)C;
      lab = (long long)(flat % (unsigned int)C);
      const float* src = boxes_img + (size_t)a * 6;
#pragma unroll
      for (int c = 0; c < 6; ++c) b[c] = src[c];
      valid = !use_thresh || prob > score_thresh;
      if (use_min) valid = valid && (b[2] - b[0] >= min_size) && (b[3] - b[1] >= min_size) && (b[5] - b[4] >= min_size);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; ++w) { int c = warp_cnt[w]; if (w < wid) before += c; total += c; }
    const int m0 = s_m;
    if (valid) {
      const int pos = m0 + before + __popc(bal & ((1u << lane) - 1u));
#pragma unroll
      for (int c = 0; c < 6; ++c) { cand_boxes[(size_t)pos * 6 + c] = b[c]; tmax = fmaxf(tmax, b[c]); }
      scores[pos] = prob;
      labels[pos] = lab;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_m = m0 + total;
    __syncthreads();
  }
  tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 16));
  tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 8));
  tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 4));
  tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
  tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
  if (lane == 0) warp_max[wid] = tmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = -CUDART_INF_F;
    for (int w = 0; w < 32; ++w) mx = fmaxf(mx, warp_max[w]);
    s_maxc = mx;
    *m_out = s_m;
  }
  __syncthreads();
  const int m = s_m;
  const float off1 = s_maxc + 1.0f;                      // (max_coordinate + 1), nms.py:104
  for (int j = threadIdx.x; j < k; j += 1024) {
    if (j < m) {
      const float off = (float)labels[j] * off1;         // idxs.to(boxes) * (max_coordinate + 1)
#pragma unroll
      for (int c = 0; c < 6; ++c) nms_boxes[(size_t)j * 6 + c] = cand_boxes[(size_t)j * 6 + c] + off;
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) nms_boxes[(size_t)j * 6 + c] = CUDART_NAN_F;   // never suppress / never suppressed
      scores[j] = -CUDART_INF_F;                                                 // sorted behind every real candidate
    }
  }
}

__global__ void post_gather_kernel(const long long* __restrict__ keep, const long long* __restrict__ n_keep,
                                   const int* __restrict__ m_ptr, const float* __restrict__ cand_boxes,
                                   const float* __restrict__ scores, const long long* __restrict__ labels, int det,
                                   float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                   long long* __restrict__ out_labels, int* __restrict__ out_count) {
  const int m = *m_ptr;
  const long long nk = *n_keep;
  int cnt = 0;
  for (int j = threadIdx.x; j < det; j += blockDim.x) {
    bool ok = j < nk && keep[j] < m;
    if (ok) {
      const long long s = keep[j];
#pragma unroll
      for (int c = 0; c < 6; ++c) out_boxes[(size_t)j * 6 + c] = cand_boxes[(size_t)s * 6 + c];
      out_scores[j] = scores[s];
      out_labels[j] = labels[s];
      ++cnt;
    }
  }
  cnt = warp_sum_i(cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out_count, cnt);
}

struct PostWs {
  SelState* st; unsigned int* ghist; int* counter; int* m;
  unsigned long long* keys; unsigned long long* keys_sorted;
  float* cand_boxes; float* nms_boxes; float* scores; long long* labels; long long* keep; long long* n_keep;
  void* cub_tmp; size_t cub_bytes; void* nms_ws; size_t nms_bytes; size_t total;
};

constexpr int MAX_POST_STREAMS = 8;
cudaStream_t g_streams[MAX_POST_STREAMS];
cudaEvent_t g_fork, g_join[MAX_POST_STREAMS];
bool g_streams_ready = false;

int ensure_streams() {
  if (g_streams_ready) return NND_OK;
  for (int i = 0; i < MAX_POST_STREAMS; ++i) {
    NND_CUDA_TRY(cudaStreamCreateWithFlags(&g_streams[i], cudaStreamNonBlocking));
    NND_CUDA_TRY(cudaEventCreateWithFlags(&g_join[i], cudaEventDisableTiming));
  }
  NND_CUDA_TRY(cudaEventCreateWithFlags(&g_fork, cudaEventDisableTiming));
  g_streams_ready = true;
  return NND_OK;
}

PostWs carve(void* ws, int k) {
  PostWs w;
  char* p = reinterpret_cast<char*>(ws);
  char* p0 = p;
  w.st = nnd_carve<SelState>(p, 1);
  w.ghist = nnd_carve<unsigned int>(p, 256);
  w.counter = nnd_carve<int>(p, 1);
  w.m = nnd_carve<int>(p, 1);
  w.keys = nnd_carve<unsigned long long>(p, k);
  w.keys_sorted = nnd_carve<unsigned long long>(p, k);
  w.cand_boxes = nnd_carve<float>(p, (size_t)k * 6);
  w.nms_boxes = nnd_carve<float>(p, (size_t)k * 6);
  w.scores = nnd_carve<float>(p, k);
  w.labels = nnd_carve<long long>(p, k);
  w.keep = nnd_carve<long long>(p, k);
  w.n_keep = nnd_carve<long long>(p, 1);
  w.cub_bytes = 0;
  cub::DeviceRadixSort::SortKeysDescending(nullptr, w.cub_bytes, (const unsigned long long*)nullptr,
                                           (unsigned long long*)nullptr, k);
  w.cub_tmp = p; p += nnd_align_up(w.cub_bytes);
  w.nms_bytes = nnd_nms_workspace_bytes(k, 3);
  w.nms_ws = p; p += nnd_align_up(w.nms_bytes);
  w.total = (size_t)(p - p0);
  return w;
}

}  // namespace

extern "C" {

size_t nnd_detect_postprocess_workspace_bytes(long long A, int C, int topk) {
  long long k = topk < A ? topk : A;
  if (k <= 0) return 256;
  return carve(nullptr, (int)k).total * MAX_POST_STREAMS;     // one private workspace per concurrent image
}

// boxes [B*A,6] decoded + clipped; probs [B*A*C]; out_* sized [B, det_per_img, ...]; out_counts [B] (device).
int nnd_detect_postprocess(const float* boxes, const float* probs, int B, long long A, int C, int topk,
                           float score_thresh, int use_score_thresh, float min_size, int use_min_size, float nms_thresh,
                           int det_per_img, float* out_boxes, float* out_scores, long long* out_labels, int* out_counts,
                           void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B <= 0 || A <= 0 || C <= 0 || topk <= 0 || det_per_img <= 0) return NND_ERR_ARG;
  if (!boxes || !probs || !out_boxes || !out_scores || !out_labels || !out_counts || !ws) return NND_ERR_ARG;
  if (A * C > 0xFFFFFFFFll) return NND_ERR_ARG;
  const int k = (int)(topk < A ? topk : A);
  const size_t per = carve(nullptr, k).total;
  if (per * MAX_POST_STREAMS > ws_bytes) return NND_ERR_WORKSPACE;
  const long long n = A * C;
  NND_CUDA_TRY(cudaMemsetAsync(out_counts, 0, sizeof(int) * B, st));
  // The per-image chains are dominated by single-CTA kernels (ordered filter, NMS scan): run the images of a batch
  // concurrently on forked streams (each with its own workspace slice), then join back into the caller's stream.
  int rc = ensure_streams();
  if (rc != NND_OK) return rc;
  NND_CUDA_TRY(cudaEventRecord(g_fork, st));
  const int hist_blocks = (int)((n + 511) / 512 < NND_NUM_SMS * 4 ? (n + 511) / 512 : NND_NUM_SMS * 4);
  for (int b = 0; b < B; ++b) {
    const int si = b % MAX_POST_STREAMS;
    cudaStream_t s = g_streams[si];
    if (b < MAX_POST_STREAMS) NND_CUDA_TRY(cudaStreamWaitEvent(s, g_fork, 0));
    PostWs w = carve(reinterpret_cast<char*>(ws) + per * si, k);
    const float* pv = probs + (size_t)b * n;
    topk_init_kernel<<<1, 256, 0, s>>>(w.st, w.ghist, w.counter, k, n);
    NND_LAUNCH_CHECK("topk_init_kernel");
    for (int round = 0; round < 8; ++round) {
      pool_hist_kernel<<<hist_blocks, 512, 0, s>>>(nullptr, pv, n, w.st, w.ghist);
      pool_pick_kernel<<<1, 256, 0, s>>>(w.st, w.ghist);
    }
    NND_LAUNCH_CHECK("topk select");
    topk_collect_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pv, n, w.st, k, w.keys, w.counter);
    NND_LAUNCH_CHECK("topk_collect_kernel");
    size_t cb = w.cub_bytes;
    NND_CUDA_TRY(cub::DeviceRadixSort::SortKeysDescending(w.cub_tmp, cb, w.keys, w.keys_sorted, k, 0, 64, s));
    post_filter_kernel<<<1, 1024, 0, s>>>(w.keys_sorted, k, boxes + (size_t)b * A * 6, C, score_thresh, use_score_thresh,
                                          min_size, use_min_size, w.cand_boxes, w.nms_boxes, w.scores, w.labels, w.m);
    NND_LAUNCH_CHECK("post_filter_kernel");
    rc = nnd_nms3d_topk_f32(w.nms_boxes, w.scores, k, nms_thresh, det_per_img, w.keep, w.n_keep, w.nms_ws, w.nms_bytes, s);
    if (rc != NND_OK) return rc;
    post_gather_kernel<<<1, 128, 0, s>>>(w.keep, w.n_keep, w.m, w.cand_boxes, w.scores, w.labels, det_per_img,
                                         out_boxes + (size_t)b * det_per_img * 6, out_scores + (size_t)b * det_per_img,
                                         out_labels + (size_t)b * det_per_img, out_counts + b);
    NND_LAUNCH_CHECK("post_gather_kernel");
  }
  for (int i = 0; i < (B < MAX_POST_STREAMS ? B : MAX_POST_STREAMS); ++i) {
    NND_CUDA_TRY(cudaEventRecord(g_join[i], g_streams[i]));
    NND_CUDA_TRY(cudaStreamWaitEvent(st, g_join[i], 0));
  }
  return NND_OK;
}

}  // extern "C"
