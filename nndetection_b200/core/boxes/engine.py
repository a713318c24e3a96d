"""Functional wrappers over the box-engine entry points of the C ABI (include/nndet_b200.h).

Every function takes/returns CUDA tensors, launches on the current stream, allocates through torch's caching
allocator and never synchronises with the host unless stated.  The classes in anchors.py / matcher.py /
sampler.py / coder.py (mirrors of nndet/core/boxes/*) and the fused train step are built on these.
"""
import math
from ctypes import c_double, c_float, c_int, c_longlong, c_size_t, c_uint, c_void_p, POINTER, byref
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from ... import _lib as L

BBOX_XFORM_CLIP = math.log(1000.0 / 16)      # torchvision BoxCoder default (nndet/core/boxes/coder.py:158-201)


def _ws(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _i3(v):
    return (c_int * 3)(*[int(x) for x in v])


def _f32c(t: Tensor) -> Tensor:
    return t.detach().contiguous().float()


# ------------------------------------------------------------------ anchors
def anchor_grid(base: Tensor, size: Sequence[int], stride: Sequence[int]) -> Tensor:
    """One pyramid level of AnchorGenerator3D.grid_anchors (nndet/core/boxes/anchors.py:337-377)."""
    L.require_cuda(base)
    base = _f32c(base)
    nb = base.shape[0]
    out = torch.empty((int(size[0]) * int(size[1]) * int(size[2]) * nb, 6), dtype=torch.float32, device=base.device)
    L.check(L.lib().nnd_anchor_grid_f32(L.ptr(out), L.ptr(base), c_int(nb), _i3(size), _i3(stride), L.stream_ptr()),
            "nnd_anchor_grid_f32")
    return out


# ------------------------------------------------------------------ pairwise metrics
def pairwise(b1: Tensor, b2: Tensor, mode: int, eps: float = 0.0) -> Tensor:
    """mode 0 IoU, 1 GIoU, 2 centre distance -> [N, M] fp32 (nndet/core/boxes/ops.py:131-185,262-287)."""
    L.require_cuda(b1, b2)
    b1, b2 = _f32c(b1), _f32c(b2)
    n, m = b1.shape[0], b2.shape[0]
    out = torch.empty((n, m), dtype=torch.float32, device=b1.device)
    for lo in range(0, n, 65535):
        hi = min(n, lo + 65535)
        L.check(L.lib().nnd_box_pairwise_f32(L.ptr(b1[lo:hi]), L.ptr(b2), c_int(hi - lo), c_int(m), c_float(eps),
                                             c_int(mode), L.ptr(out[lo:hi]), L.stream_ptr()), "nnd_box_pairwise_f32")
    return out


# ------------------------------------------------------------------ decode / probabilities
def decode_boxes(deltas: Tensor, anchors: Tensor, clip_shape: Optional[Sequence[int]] = None,
                 xform_clip: float = BBOX_XFORM_CLIP) -> Tensor:
    """decode_single (coder.py:90-155, weights 1) for deltas [n,6]; anchors [A,6] reused cyclically (n = B*A);
    optional clip_boxes_to_image_3d_ (clip.py:83-101) fused in."""
    L.require_cuda(deltas, anchors)
    deltas, anchors = _f32c(deltas), _f32c(anchors)
    out = torch.empty_like(deltas)
    shp = (c_float * 3)(*[float(s) for s in clip_shape]) if clip_shape is not None else None
    L.check(L.lib().nnd_decode_boxes_f32(L.ptr(deltas), L.ptr(anchors), c_longlong(deltas.shape[0]),
                                         c_longlong(anchors.shape[0]), c_float(xform_clip), shp, L.ptr(out),
                                         L.stream_ptr()), "nnd_decode_boxes_f32")
    return out


def sigmoid_fg(logits: Tensor, want_probs: bool = True, want_fg: bool = True) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """sigmoid(logits) [n,C] and max over classes [n] (nndet/arch/heads/comb.py:262-263)."""
    L.require_cuda(logits)
    logits = _f32c(logits)
    n, C = logits.shape
    probs = torch.empty_like(logits) if want_probs else None
    fg = torch.empty(n, dtype=torch.float32, device=logits.device) if want_fg else None
    L.check(L.lib().nnd_sigmoid_fg_f32(L.ptr(logits), c_longlong(n), c_int(C), L.ptr(probs), L.ptr(fg), L.stream_ptr()),
            "nnd_sigmoid_fg_f32")
    return probs, fg


# ------------------------------------------------------------------ ATSS
class GtBatch:
    """Ground truth of a batch packed for the device: boxes [G,6] fp32, classes [G] int64, img/local index per
    box (int32) and per-image offsets [B+1] (int32).  Built on the host from the reference's target lists
    (`target_boxes`, `target_classes`: nndet/core/retina.py:126-128) with one H2D copy per tensor."""

    def __init__(self, target_boxes: List[Tensor], target_classes: List[Tensor], device):
        counts = [int(b.shape[0]) for b in target_boxes]
        self.B = len(target_boxes)
        self.G = sum(counts)
        off = [0]
        for c in counts:
            off.append(off[-1] + c)
        img = [i for i, c in enumerate(counts) for _ in range(c)]
        loc = [j for c in counts for j in range(c)]
        self.counts = counts
        self.offsets = torch.tensor(off, dtype=torch.int32).to(device, non_blocking=True)
        self.img = torch.tensor(img, dtype=torch.int32).to(device, non_blocking=True)
        self.local = torch.tensor(loc, dtype=torch.int32).to(device, non_blocking=True)
        if self.G > 0:
            self.boxes = torch.cat([b.reshape(-1, 6) for b in target_boxes]).to(device, torch.float32).contiguous()
            self.classes = torch.cat([c.reshape(-1) for c in target_classes]).to(device, torch.int64).contiguous()
        else:
            self.boxes = torch.zeros((1, 6), dtype=torch.float32, device=device)
            self.classes = torch.zeros((1,), dtype=torch.int64, device=device)


_LEVEL_CACHE = {}


def _level_offsets(per_level: Sequence[int], device):
    key = (tuple(int(x) for x in per_level), str(device))
    if key not in _LEVEL_CACHE:
        off = [0]
        for a in per_level:
            off.append(off[-1] + int(a))
        _LEVEL_CACHE[key] = (torch.tensor(off, dtype=torch.int32, device=device), (c_int * len(off))(*off), off)
    return _LEVEL_CACHE[key]


def atss_match(gt: GtBatch, anchors: Tensor, per_level: Sequence[int], kc: int) -> Tensor:
    """ATSSMatcher.compute_matches for all images of a batch (matcher/atss.py:48-122) -> matches [B*A] int64."""
    L.require_cuda(anchors)
    lib = L.lib()
    A = anchors.shape[0]
    off_dev, off_host, off = _level_offsets(per_level, anchors.device)
    assert off[-1] == A, "num_anchors_per_level does not sum to the number of anchors"
    nl = len(per_level)
    matches = torch.empty(gt.B * A, dtype=torch.int64, device=anchors.device)
    lib.nnd_atss_workspace_bytes.restype = c_size_t
    nbytes = lib.nnd_atss_workspace_bytes(c_int(gt.G), c_int(A), off_host, c_int(nl), c_int(kc))
    ws = _ws(nbytes, anchors.device)
    L.check(lib.nnd_atss_match(L.ptr(gt.boxes), L.ptr(gt.img), L.ptr(gt.local), c_int(gt.G), L.ptr(anchors), c_int(A),
                               c_int(gt.B), L.ptr(off_dev), off_host, c_int(nl), c_int(kc), L.ptr(matches), L.ptr(ws),
                               c_size_t(ws.numel()), L.stream_ptr()), "nnd_atss_match")
    return matches


def assign_labels(matches: Tensor, gt: GtBatch, A: int) -> Tensor:
    """labels [B*A] fp32: 0 background, c+1 foreground, -1 ignore (nndet/core/retina.py:256-288)."""
    labels = torch.empty(matches.shape[0], dtype=torch.float32, device=matches.device)
    L.check(L.lib().nnd_assign_labels(L.ptr(matches), c_longlong(matches.shape[0]), c_longlong(A), L.ptr(gt.classes),
                                      L.ptr(gt.offsets), L.ptr(labels), L.stream_ptr()), "nnd_assign_labels")
    return labels


# ------------------------------------------------------------------ sampler
class SamplerPlan:
    """Host-side arithmetic of HardNegativeSamplerBatched (sampler.py:154-185,251-252), in Python floats."""

    def __init__(self, batch_size: int, batch_size_per_image: int = 32, positive_fraction: float = 0.33,
                 min_neg: int = 1, pool_size: float = 20):
        bspi = batch_size_per_image * batch_size
        self.max_pos = int(bspi * positive_fraction)
        self.neg_ratio = abs(1 - 1. / float(positive_fraction))
        self.min_neg = int(min_neg)
        self.pool_size = float(pool_size)
        self.max_neg = max(int(max(1, self.max_pos) * self.neg_ratio), self.min_neg)
        self.max_pool = int(self.max_neg * self.pool_size)


def hnm_sample(labels: Tensor, fg_probs: Tensor, plan: SamplerPlan, seed: int, want_pool: bool = False):
    """-> counts int32[8] (0 #pos, 1 #neg, 2 num_pos, 3 num_neg, 4 pool), pos_idx int64[max_pos], neg_idx
    int64[max_neg] (ascending, only the first num_pos / num_neg entries are valid) [, pool int32[max_pool]]."""
    L.require_cuda(labels, fg_probs)
    lib = L.lib()
    dev = labels.device
    n = labels.shape[0]
    if max(plan.max_pos, plan.max_neg) > lib.nnd_hnm_max_select():
        raise L.NndError("sampler batch too large for the single-CTA pick kernel")
    pos_cap, pool_cap = 1 << 20, max(plan.max_pool, 1)
    counts = torch.empty(8, dtype=torch.int32, device=dev)
    pos = torch.zeros(max(plan.max_pos, 1), dtype=torch.int64, device=dev)
    neg = torch.zeros(max(plan.max_neg, 1), dtype=torch.int64, device=dev)
    lib.nnd_hnm_workspace_bytes.restype = c_size_t
    nbytes = lib.nnd_hnm_workspace_bytes(c_longlong(n), c_int(pos_cap), c_int(pool_cap))
    ws = _ws(nbytes, dev)
    pool_ptr = c_void_p(0)
    L.check(lib.nnd_hnm_sample(L.ptr(labels), L.ptr(fg_probs), c_longlong(n), c_int(plan.max_pos),
                               c_double(plan.neg_ratio), c_int(plan.min_neg), c_double(plan.pool_size),
                               c_uint(seed & 0xFFFFFFFF), L.ptr(counts), L.ptr(pos), L.ptr(neg), c_int(pos_cap),
                               c_int(pool_cap), byref(pool_ptr), L.ptr(ws), c_size_t(ws.numel()), L.stream_ptr()),
            "nnd_hnm_sample")
    if want_pool:
        off = pool_ptr.value - ws.data_ptr()
        pool = ws[off: off + 4 * pool_cap].view(torch.int32)
        return counts, pos, neg, pool, ws
    return counts, pos, neg


# ------------------------------------------------------------------ head loss
def head_loss_fwd(logits: Tensor, deltas: Tensor, anchors: Tensor, matches: Tensor, gt: GtBatch, labels: Tensor,
                  pos: Tensor, neg: Tensor, counts: Tensor, giou_eps: float = 1e-7,
                  xform_clip: float = BBOX_XFORM_CLIP):
    """DetectionHeadHNMNative.compute_loss (comb.py:383-405) -> losses[2] = (reg, cls) + compact gradients."""
    dev = logits.device
    C = logits.shape[1]
    losses = torch.empty(2, dtype=torch.float32, device=dev)
    g_deltas = torch.empty((pos.shape[0], 6), dtype=torch.float32, device=dev)
    g_logits = torch.empty((pos.shape[0] + neg.shape[0], C), dtype=torch.float32, device=dev)
    L.check(L.lib().nnd_head_loss_fwd(L.ptr(logits), L.ptr(deltas), L.ptr(anchors), c_longlong(anchors.shape[0]),
                                      c_int(C), L.ptr(matches), L.ptr(gt.boxes), L.ptr(gt.offsets), L.ptr(labels),
                                      L.ptr(pos), L.ptr(neg), L.ptr(counts), c_float(xform_clip), c_float(giou_eps),
                                      L.ptr(losses), L.ptr(g_deltas), L.ptr(g_logits), L.stream_ptr()),
            "nnd_head_loss_fwd")
    return losses, g_deltas, g_logits


def head_loss_bwd(g_deltas: Tensor, g_logits: Tensor, pos: Tensor, neg: Tensor, counts: Tensor, n: int,
                  up_reg: Optional[Tensor], up_cls: Optional[Tensor]):
    """Dense d_deltas [n,6], d_logits [n,C] (zero except the sampled rows) scaled by the upstream grads."""
    dev = g_deltas.device
    C = g_logits.shape[1]
    d_deltas = torch.zeros((n, 6), dtype=torch.float32, device=dev)
    d_logits = torch.zeros((n, C), dtype=torch.float32, device=dev)
    L.check(L.lib().nnd_head_loss_bwd(L.ptr(g_deltas), L.ptr(g_logits), c_int(C), L.ptr(pos), L.ptr(neg), L.ptr(counts),
                                      L.ptr(up_reg), L.ptr(up_cls), L.ptr(d_deltas), L.ptr(d_logits), L.stream_ptr()),
            "nnd_head_loss_bwd")
    return d_deltas, d_logits


# ------------------------------------------------------------------ post-processing
def detect_postprocess(boxes: Tensor, probs: Tensor, B: int, A: int, C: int, topk: int = 10000,
                       score_thresh: Optional[float] = 0.0, min_size: Optional[float] = 0.01, nms_thresh: float = 0.6,
                       det_per_img: int = 100):
    """postprocess_detections_single_image for every image (retina.py:332-379), sync-free.
    -> (boxes [B,det,6], scores [B,det], labels [B,det] int64, counts [B] int32 on device)."""
    L.require_cuda(boxes, probs)
    lib = L.lib()
    dev = boxes.device
    ob = torch.empty((B, det_per_img, 6), dtype=torch.float32, device=dev)
    os_ = torch.empty((B, det_per_img), dtype=torch.float32, device=dev)
    ol = torch.empty((B, det_per_img), dtype=torch.int64, device=dev)
    oc = torch.empty(B, dtype=torch.int32, device=dev)
    lib.nnd_detect_postprocess_workspace_bytes.restype = c_size_t
    nbytes = lib.nnd_detect_postprocess_workspace_bytes(c_longlong(A), c_int(C), c_int(topk))
    ws = _ws(nbytes, dev)
    L.check(lib.nnd_detect_postprocess(L.ptr(boxes), L.ptr(probs), c_int(B), c_longlong(A), c_int(C), c_int(topk),
                                       c_float(score_thresh if score_thresh is not None else 0.0),
                                       c_int(score_thresh is not None),
                                       c_float(min_size if min_size is not None else 0.0), c_int(min_size is not None),
                                       c_float(nms_thresh), c_int(det_per_img), L.ptr(ob), L.ptr(os_), L.ptr(ol),
                                       L.ptr(oc), L.ptr(ws), c_size_t(ws.numel()), L.stream_ptr()),
            "nnd_detect_postprocess")
    return ob, os_, ol, oc
