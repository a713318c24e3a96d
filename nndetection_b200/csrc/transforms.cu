// Instance-segmentation -> detection targets, the `pre_trafo` chain the reference runs inside every training / validation step
// on GPU tensors (nndet/ptmodule/retinaunet/base.py:114-141): FindInstances (nndet/io/transforms/instances.py:25-39),
// Instances2Boxes / instances_to_boxes (:42-136) and Instances2Segmentation / instances_to_segmentation (:211-301).
// The reference loops over instances with `unique`, `(seg == id).nonzero()`, `.item()` and boolean-mask writes (one or more
// host synchronisations and full-volume passes PER INSTANCE).  Here: ONE streaming pass over the volume (4 B read + 4 B
// written per voxel) that min/max-reduces voxel coordinates per (sample, instance id) and writes the semantic map through a
// per-sample id -> class lookup table, then one tiny compaction kernel (ids ascending, like `unique(sorted=True)`).
// Bit-exact integer work: ids, box corners (min-1 / max+1 as floats), classes, semantic labels.
#include "common.cuh"

namespace {

constexpr int INST_THREADS = 256;

// bounds[b][id][0..2] = min index along the three spatial axes, [3..5] = max index; initialised to INT_MAX / -1
__global__ void __launch_bounds__(INST_THREADS)
inst_scan_kernel(const float* __restrict__ target, int B, int D, int H, int W, const int* __restrict__ lut, int max_id,
                 float* __restrict__ sem, int* __restrict__ bounds, int* __restrict__ err) {
  const long long V = (long long)D * H * W;
  const long long total = (long long)B * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float tv = target[i];
    const int id = (int)tv;                               // .to(torch.int): truncation (instances.py:34)
    float out = 0.f;
    if (id > 0) {
      if (id >= max_id || (float)id != tv) {
        atomicOr(err, 1);                                 // id outside the table / not an integer label
      } else {
        const int b = (int)(i / V);
        const long long r = i - (long long)b * V;
        const int w = (int)(r % W); const long long r2 = r / W;
        const int h = (int)(r2 % H); const int d = (int)(r2 / H);
        // warp-aggregate voxels of the same instance: one leader issues the six atomics
        const unsigned peers = __match_any_sync(__activemask(), b * max_id + id);
        int d0 = d, d1 = d, h0 = h, h1 = h, w0 = w, w1 = w;
        for (unsigned m = peers & (peers - 1) ? peers : 0u; m; m &= m - 1) {
          const int src = __ffs(m) - 1;
          const int od = __shfl_sync(peers, d, src), oh = __shfl_sync(peers, h, src), ow = __shfl_sync(peers, w, src);
          d0 = min(d0, od); d1 = max(d1, od); h0 = min(h0, oh); h1 = max(h1, oh); w0 = min(w0, ow); w1 = max(w1, ow);
        }
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) {
          int* bp = bounds + ((size_t)b * max_id + id) * 6;
          atomicMin(bp + 0, d0); atomicMin(bp + 1, h0); atomicMin(bp + 2, w0);
          atomicMax(bp + 3, d1); atomicMax(bp + 4, h1); atomicMax(bp + 5, w1);
        }
        const int cls = lut[(size_t)b * max_id + id];     // class (+1 with add_background), -1 = id missing in the mapping
        if (cls < 0) atomicOr(err, 2);
        out = (float)max(cls, 0);
      }
    }
    sem[i] = out;
  }
}

// one block per sample: ids ascending -> present ids, boxes (x1,y1,x2,y2,z1,z2) = (min0-1, min1-1, max0+1, max1+1, min2-1, max2+1)
__global__ void __launch_bounds__(256)
inst_compact_kernel(const int* __restrict__ bounds, const int* __restrict__ lut_class, int max_id, int cap,
                    int* __restrict__ out_ids, float* __restrict__ out_boxes, long long* __restrict__ out_classes,
                    int* __restrict__ counts) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int id0 = 0; id0 < max_id; id0 += 256) {
    const int id = id0 + tid;
    const int* bp = bounds + ((size_t)b * max_id + id) * 6;
    const bool present = id > 0 && id < max_id && bp[3] >= 0;
    const unsigned bal = __ballot_sync(0xffffffffu, present);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (present) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap) {
        out_ids[(size_t)b * cap + pos] = id;
        float* ob = out_boxes + ((size_t)b * cap + pos) * 6;
        ob[0] = (float)(bp[0] - 1); ob[1] = (float)(bp[1] - 1); ob[2] = (float)(bp[3] + 1);
        ob[3] = (float)(bp[4] + 1); ob[4] = (float)(bp[2] - 1); ob[5] = (float)(bp[5] + 1);
        out_classes[(size_t)b * cap + pos] = (long long)lut_class[(size_t)b * max_id + id];
      }
    }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
  if (tid == 0) counts[b] = s_base;
}

__global__ void inst_init_bounds_kernel(int* bounds, long long n6) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n6) bounds[i] = (i % 6) < 3 ? 0x7fffffff : -1;
}

}  // namespace

extern "C" {

size_t nnd_instances_workspace_bytes(int B, int max_id) { return nnd_align_up((size_t)B * max_id * 6 * sizeof(int)) + 256; }

// target / sem_out: fp32 [B, D, H, W] (the channel axis of size 1 squeezed); lut_sem[b][id] = label written into the semantic
// map (mapping[id] + add_background, -1 = id not in the mapping), lut_class[b][id] = mapping[id]; both int32 [B, max_id] on the
// device.  Outputs padded to `cap` instances per sample: out_ids int32 [B, cap], out_boxes fp32 [B, cap, 6], out_classes int64
// [B, cap], counts int32 [B]; err_out int32 (bit 0: id >= max_id or non-integer label, bit 1: id missing in the mapping).
int nnd_instances_to_targets(const float* target, int B, int D, int H, int W, const int* lut_sem, const int* lut_class,
                             int max_id, int cap, float* sem_out, int* out_ids, float* out_boxes, long long* out_classes,
                             int* counts, int* err_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!target || !lut_sem || !lut_class || !sem_out || !out_ids || !out_boxes || !out_classes || !counts || !err_out || !ws)
    return NND_ERR_ARG;
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || max_id < 2 || cap < 1) return NND_ERR_ARG;
  if (ws_bytes < nnd_instances_workspace_bytes(B, max_id)) return NND_ERR_WORKSPACE;
  int* bounds = reinterpret_cast<int*>(ws);
  const long long n6 = (long long)B * max_id * 6;
  inst_init_bounds_kernel<<<(unsigned)((n6 + 255) / 256), 256, 0, stream>>>(bounds, n6);
  NND_LAUNCH_CHECK("inst_init_bounds_kernel");
  NND_CUDA_TRY(cudaMemsetAsync(err_out, 0, sizeof(int), stream));
  const long long total = (long long)B * D * H * W;
  long long blocks = (total + INST_THREADS - 1) / INST_THREADS;
  if (blocks > NND_NUM_SMS * 16) blocks = NND_NUM_SMS * 16;
  inst_scan_kernel<<<(unsigned)blocks, INST_THREADS, 0, stream>>>(target, B, D, H, W, lut_sem, max_id, sem_out, bounds, err_out);
  NND_LAUNCH_CHECK("inst_scan_kernel");
  inst_compact_kernel<<<B, 256, 0, stream>>>(bounds, lut_class, max_id, cap, out_ids, out_boxes, out_classes, counts);
  NND_LAUNCH_CHECK("inst_compact_kernel");
  return NND_OK;
}

}  // extern "C"
