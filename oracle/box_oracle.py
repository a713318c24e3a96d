"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the nnDetection box engine.

A from-scratch CPU restatement (torch-CPU tensor arithmetic, fp32, so that the
rounding of every intermediate equals the reference's own CPU path) of the box
engine the hot path uses.  Nothing in `nndetection_b200/` may import this
module; only tests/, bench.py's cpu_baseline / --impl reference leg and
__graft_entry__.smoke() do, and only as the checker.

Pinning: scripts/gen_golden.py runs the *unmodified* reference (imported from
/root/reference through oracle/ref_import.py) and this file on the same seeded
inputs, asserts equality, and stores the vectors in tests/golden/.  The
reference's own tests pin nothing on this path (tests/test_imports.py:1-18), so
the executed reference is the anchor.

All citations are file:line under /root/reference/.
"""
from __future__ import annotations

import math
from itertools import product
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

INF = 100.0                     # nndet/core/boxes/matcher/atss.py:19
BELOW_LOW_THRESHOLD = -1        # nndet/core/boxes/matcher/base.py:13-15
BETWEEN_THRESHOLDS = -2
BBOX_XFORM_CLIP = math.log(1000.0 / 16)   # torchvision BoxCoder default used by coder.py:158-201


# --------------------------------------------------------------------------- anchors
def base_anchors_3ds(width: Sequence[float], height: Sequence[float], depth: Sequence[float]) -> torch.Tensor:
    """AnchorGenerator3DS.generate_anchors, nndet/core/boxes/anchors.py:526-549.

    product(width, height, depth)/2 laid out as (-w, -h, w, h, -d, d); no rounding.
    """
    sizes = torch.tensor(list(product(width, height, depth)), dtype=torch.float32) / 2
    return torch.stack([-sizes[:, 0], -sizes[:, 1], sizes[:, 0], sizes[:, 1], -sizes[:, 2], sizes[:, 2]], dim=1)


def grid_anchors_3d(grid_sizes: Sequence[Sequence[int]], strides: Sequence[Sequence[int]],
                    cell_anchors: Sequence[torch.Tensor]) -> Tuple[List[torch.Tensor], List[int]]:
    """AnchorGenerator3D.grid_anchors, nndet/core/boxes/anchors.py:337-377.

    Position-major / anchor-minor, positions enumerated with axis 0 slowest and
    axis 2 fastest (meshgrid indexing="ij", :360-368); coordinate order of a box
    is (x1, y1, x2, y2, z1, z2) with x=axis0, y=axis1, z=axis2 (:369).
    """
    out, per_level = [], []
    for size, stride, base in zip(grid_sizes, strides, cell_anchors):
        s0 = torch.arange(0, size[0], dtype=torch.float32) * stride[0]
        s1 = torch.arange(0, size[1], dtype=torch.float32) * stride[1]
        s2 = torch.arange(0, size[2], dtype=torch.float32) * stride[2]
        g0, g1, g2 = torch.meshgrid(s0, s1, s2, indexing="ij")
        g0, g1, g2 = g0.reshape(-1), g1.reshape(-1), g2.reshape(-1)
        shifts = torch.stack((g0, g1, g0, g1, g2, g2), dim=1)
        a = (shifts[:, None, :] + base[None, :, :]).reshape(-1, 6)
        out.append(a)
        per_level.append(a.shape[0])
    return out, per_level


def anchors_for_image(image_size: Sequence[int], fmap_sizes: Sequence[Sequence[int]],
                      width, height, depth) -> Tuple[torch.Tensor, List[int]]:
    """AnchorGenerator2D.forward for one image, anchors.py:211-242 (strides = int(img/fm), :225)."""
    strides = [[int(i / s) for i, s in zip(image_size, fm)] for fm in fmap_sizes]
    cells = [base_anchors_3ds(w, h, d) for w, h, d in zip(width, height, depth)]
    per_fm, per_level = grid_anchors_3d(fmap_sizes, strides, cells)
    return torch.cat(per_fm, dim=0), per_level


# --------------------------------------------------------------------------- pairwise metrics
def box_volume(b: torch.Tensor) -> torch.Tensor:
    """box_area_3d, nndet/core/boxes/ops.py:27-38: (x2-x1)*(y2-y1)*(z2-z1)."""
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) * (b[:, 5] - b[:, 4])


def iou_union_3d(b1: torch.Tensor, b2: torch.Tensor, eps: float = 0.0):
    """box_iou_union_3d, ops.py:131-159: inter = dx+ * dy+ * dz+ + eps; union = v1 + v2 - inter."""
    b1, b2 = b1.float(), b2.float()
    v1, v2 = box_volume(b1), box_volume(b2)
    lo_x = torch.max(b1[:, None, 0], b2[:, 0]); lo_y = torch.max(b1[:, None, 1], b2[:, 1])
    hi_x = torch.min(b1[:, None, 2], b2[:, 2]); hi_y = torch.min(b1[:, None, 3], b2[:, 3])
    lo_z = torch.max(b1[:, None, 4], b2[:, 4]); hi_z = torch.min(b1[:, None, 5], b2[:, 5])
    inter = ((hi_x - lo_x).clamp(min=0) * (hi_y - lo_y).clamp(min=0) * (hi_z - lo_z).clamp(min=0)) + eps
    union = v1[:, None] + v2 - inter
    return inter / union, union


def box_iou(b1: torch.Tensor, b2: torch.Tensor, eps: float = 0.0) -> torch.Tensor:
    """box_iou, ops.py:74-102 (empty input -> empty 1-D tensor, :96-97)."""
    if b1.numel() == 0 or b2.numel() == 0:
        return torch.tensor([]).to(b1)
    return iou_union_3d(b1, b2, eps)[0]


def generalized_box_iou(b1: torch.Tensor, b2: torch.Tensor, eps: float = 0.0) -> torch.Tensor:
    """generalized_box_iou_3d, ops.py:162-185.  NB the inner IoU is called WITHOUT eps (:175)."""
    if b1.numel() == 0 or b2.numel() == 0:
        return torch.tensor([]).to(b1)
    b1, b2 = b1.float(), b2.float()
    iou, union = iou_union_3d(b1, b2)
    lo_x = torch.min(b1[:, None, 0], b2[:, 0]); lo_y = torch.min(b1[:, None, 1], b2[:, 1])
    hi_x = torch.max(b1[:, None, 2], b2[:, 2]); hi_y = torch.max(b1[:, None, 3], b2[:, 3])
    lo_z = torch.min(b1[:, None, 4], b2[:, 4]); hi_z = torch.max(b1[:, None, 5], b2[:, 5])
    hull = ((hi_x - lo_x).clamp(min=0) * (hi_y - lo_y).clamp(min=0) * (hi_z - lo_z).clamp(min=0)) + eps
    return iou - (hull - union) / hull


def box_center(b: torch.Tensor) -> torch.Tensor:
    """box_center, ops.py:314-327: per-axis (hi + lo) / 2 stacked (x, y, z)."""
    return torch.stack([(b[:, 2] + b[:, 0]) / 2.0, (b[:, 3] + b[:, 1]) / 2.0, (b[:, 5] + b[:, 4]) / 2.0], dim=1)


def box_center_dist(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """box_center_dist (euclidean), ops.py:262-287: sqrt(sum((c1[:,None]-c2[None])**2))."""
    c1, c2 = box_center(b1), box_center(b2)
    return (c1[:, None] - c2[None]).pow(2).sum(-1).sqrt()


# --------------------------------------------------------------------------- ATSS
def atss_match(gt: torch.Tensor, anchors: torch.Tensor, num_anchors_per_level: Sequence[int],
               num_anchors_per_loc: int, num_candidates: int = 4,
               canonical_ties: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """ATSSMatcher.compute_matches, nndet/core/boxes/matcher/atss.py:48-122 (center_in_gt=False),
    with Matcher.__call__'s no-GT shortcut (matcher/base.py:51-56).

    canonical_ties=True orders equal centre distances by ascending anchor index
    (the tie-break the CUDA path defines; torch.topk's own choice is
    unspecified).  On tie-free k-boundaries both agree.
    Returns (iou[G, A] fp32, matches[A] int64).
    """
    A = anchors.shape[0]
    if gt.numel() == 0:
        return torch.tensor([]).to(anchors), torch.full((A,), BELOW_LOW_THRESHOLD, dtype=torch.int64)
    G = gt.shape[0]
    dist = box_center_dist(gt, anchors)
    cand = []
    start = 0
    for apl in num_anchors_per_level:
        k = min(num_candidates * num_anchors_per_loc, apl)
        d = dist[:, start:start + apl]
        if canonical_ties:
            idx = torch.argsort(d, dim=1, stable=True)[:, :k]
        else:
            idx = d.topk(k, dim=1, largest=False)[1]
        cand.append(idx + start)
        start += apl
    cand = torch.cat(cand, dim=1)                                   # [G, K]
    iou = box_iou(gt, anchors)                                      # [G, A]
    c_iou = iou.gather(1, cand)
    thr = c_iou.mean(dim=1) + c_iou.std(dim=1)                      # unbiased std, atss.py:99
    is_pos = c_iou >= thr[:, None]                                  # atss.py:101
    flat = torch.full((G * A,), -INF)
    lin = (cand + torch.arange(G)[:, None] * A).reshape(-1)[is_pos.reshape(-1)]
    flat[lin] = iou.reshape(-1)[lin]
    vals, matches = flat.view(G, A).max(dim=0)                      # first GT wins on equal IoU
    matches[vals == -INF] = BELOW_LOW_THRESHOLD
    return iou, matches


def assign_targets(matches: torch.Tensor, gt_boxes: torch.Tensor, gt_classes: torch.Tensor,
                   num_anchors: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """BaseRetinaNet.assign_targets_to_anchors body for one image, nndet/core/retina.py:256-288."""
    if gt_boxes.numel() > 0:
        m = matches.clamp(min=0)
        matched = gt_boxes[m]
        labels = gt_classes[m].to(torch.float32) + 1
    else:
        matched = torch.zeros(num_anchors, 6)
        labels = torch.zeros(num_anchors)
    labels[matches == BELOW_LOW_THRESHOLD] = 0.0
    labels[matches == BETWEEN_THRESHOLDS] = -1.0
    return labels, matched


# --------------------------------------------------------------------------- sampler (deterministic parts)
def hnm_counts(num_positive: int, num_negative: int, batch_size: int, batch_size_per_image: int = 32,
               positive_fraction: float = 0.33, min_neg: int = 1, pool_size: float = 20) -> Tuple[int, int, int]:
    """HardNegativeSamplerBatched arithmetic, nndet/core/boxes/sampler.py:154-185,85-86,251-252.

    Returns (num_pos, num_neg, pool).
    """
    bspi = batch_size_per_image * batch_size
    num_pos = min(num_positive, int(bspi * positive_fraction))
    num_neg = int(max(1, num_pos) * abs(1 - 1.0 / float(positive_fraction)))
    num_neg = min(num_negative, max(num_neg, min_neg))
    pool = min(num_negative, int(num_neg * pool_size))
    return num_pos, num_neg, pool


def hnm_pool(labels_batch: torch.Tensor, fg_probs: torch.Tensor, pool: int) -> torch.Tensor:
    """Pool of hardest negatives, sampler.py:67-90: top-`pool` fg_prob among labels == 0.

    Ties broken by ascending anchor index (canonical).  Returns sorted anchor indices.
    """
    neg = torch.where(labels_batch == 0)[0]
    p = fg_probs[neg]
    order = torch.argsort(-p, stable=True)[:pool]
    return torch.sort(neg[order])[0]


def mix32(x: np.ndarray) -> np.ndarray:
    """The counter hash nndetection_b200 uses to draw sampling priorities (NOT torch.randperm;
    the reference's RNG stream cannot be reproduced on device, SURVEY 7 hard part 3)."""
    x = x.astype(np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x.astype(np.uint32)


def hash_priority(idx: np.ndarray, seed: int, stream: int) -> np.ndarray:
    s = (int(seed) * 0x9E3779B1 + int(stream) * 0x85EBCA77) & 0xFFFFFFFF
    return mix32(np.asarray(idx, dtype=np.uint64) ^ np.uint64(s))


def hnm_select(labels_batch: torch.Tensor, fg_probs: torch.Tensor, batch_size: int, seed: int, **kw):
    """Full sampler with the device hash as the random source: positives = num_pos smallest
    priorities among labels >= 1; negatives = num_neg smallest priorities in the pool.
    Returns ascending index tensors like torch.where (comb.py:270-271)."""
    pos = torch.where(labels_batch >= 1)[0]
    neg = torch.where(labels_batch == 0)[0]
    num_pos, num_neg, pool = hnm_counts(pos.numel(), neg.numel(), batch_size, **kw)
    pr = hash_priority(pos.numpy(), seed, 1).astype(np.uint64) * (1 << 32) + pos.numpy().astype(np.uint64)
    sel_pos = torch.sort(pos[torch.from_numpy(np.argsort(pr, kind="stable")[:num_pos])])[0]
    pool_idx = hnm_pool(labels_batch, fg_probs, pool)
    pr = hash_priority(pool_idx.numpy(), seed, 2).astype(np.uint64) * (1 << 32) + pool_idx.numpy().astype(np.uint64)
    sel_neg = torch.sort(pool_idx[torch.from_numpy(np.argsort(pr, kind="stable")[:num_neg])])[0]
    return sel_pos, sel_neg, pool_idx


# --------------------------------------------------------------------------- coder / clip / filter
def decode_single(rel: torch.Tensor, boxes: torch.Tensor, clip: float = BBOX_XFORM_CLIP) -> torch.Tensor:
    """decode_single with all weights 1.0, nndet/core/boxes/coder.py:90-155."""
    boxes = boxes.to(rel.dtype)
    w = boxes[:, 2] - boxes[:, 0]; h = boxes[:, 3] - boxes[:, 1]; d = boxes[:, 5] - boxes[:, 4]
    cx = boxes[:, 0] + 0.5 * w; cy = boxes[:, 1] + 0.5 * h; cz = boxes[:, 4] + 0.5 * d
    dx, dy, dw, dh, dz, dd = (rel[:, i] / 1.0 for i in range(6))
    dw = dw.clamp(max=clip); dh = dh.clamp(max=clip); dd = dd.clamp(max=clip)
    pcx = dx * w + cx; pcy = dy * h + cy; pcz = dz * d + cz
    pw = torch.exp(dw) * w; ph = torch.exp(dh) * h; pd = torch.exp(dd) * d
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph,
                        pcz - 0.5 * pd, pcz + 0.5 * pd], dim=1)


def clip_boxes_3d(b: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    """clip_boxes_to_image_3d_, nndet/core/boxes/clip.py:83-101: x in [0,s0], y in [0,s1], z in [0,s2]."""
    b = b.clone()
    b[:, 0] = b[:, 0].clamp(0, shape[0]); b[:, 2] = b[:, 2].clamp(0, shape[0])
    b[:, 1] = b[:, 1].clamp(0, shape[1]); b[:, 3] = b[:, 3].clamp(0, shape[1])
    b[:, 4] = b[:, 4].clamp(0, shape[2]); b[:, 5] = b[:, 5].clamp(0, shape[2])
    return b


def keep_not_small(b: torch.Tensor, min_size: float) -> torch.Tensor:
    """remove_small_boxes, ops.py:241-259: all three extents >= min_size; returns indices."""
    ok = ((b[:, 2] - b[:, 0]) >= min_size) & ((b[:, 3] - b[:, 1]) >= min_size) & ((b[:, 5] - b[:, 4]) >= min_size)
    return torch.where(ok)[0]


# --------------------------------------------------------------------------- NMS
def nms_greedy(boxes: torch.Tensor, scores: torch.Tensor, thr: float, cuda_semantics: bool = True) -> torch.Tensor:
    """Greedy 3-D NMS.  Restates nms_cuda + nms_kernel_3d + devIoU_3d
    (nndet/csrc/cuda/nms.cu:148-221,99-145,36-51) when cuda_semantics=True: IoU =
    inter/(Sa+Sb-inter) in fp32 with no eps, j suppressed by kept i iff IoU > thr (strict, :138), so a NaN
    IoU never suppresses.  cuda_semantics=False restates nms_cpu (nndet/core/boxes/nms.py:31-53): survivors
    are those with IoU <= thr, so NaN IoU suppresses.  Score ties -> ascending index (stable sort).
    Returns int64 indices into `boxes`, descending score.
    """
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64)
    b = boxes.float().numpy()
    order = torch.argsort(-scores.float(), stable=True).numpy()
    b = b[order]
    vol = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) * (b[:, 5] - b[:, 4])
    alive = np.ones(n, dtype=bool)
    keep = []
    thr32 = np.float32(thr)
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n):
            if not alive[i]:
                continue
            keep.append(order[i])
            if i + 1 == n:
                break
            r = b[i + 1:]
            w = np.maximum(np.minimum(b[i, 2], r[:, 2]) - np.maximum(b[i, 0], r[:, 0]), np.float32(0))
            h = np.maximum(np.minimum(b[i, 3], r[:, 3]) - np.maximum(b[i, 1], r[:, 1]), np.float32(0))
            d = np.maximum(np.minimum(b[i, 5], r[:, 5]) - np.maximum(b[i, 4], r[:, 4]), np.float32(0))
            inter = (w * h * d).astype(np.float32)
            iou = inter / ((vol[i] + vol[i + 1:]).astype(np.float32) - inter)
            if cuda_semantics:
                alive[i + 1:] &= ~(iou > thr32)
            else:
                alive[i + 1:] &= (iou <= thr32)
    return torch.from_numpy(np.asarray(keep, dtype=np.int64))


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, thr: float,
                cuda_semantics: bool = True) -> torch.Tensor:
    """batched_nms, nndet/core/boxes/nms.py:81-106: offset = idx * (max_coord + 1) in fp32."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    off = idxs.to(boxes) * (boxes.max() + 1)
    return nms_greedy(boxes + off[:, None], scores, thr, cuda_semantics)


# --------------------------------------------------------------------------- weighted box clustering
def wbc(boxes: torch.Tensor, scores: torch.Tensor, weights: torch.Tensor, n_exp_preds: torch.Tensor, iou_thresh: float,
        score_thresh: float, use_area: bool = True, missing_weight: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """wbc + compute_cluster_consolidation, nndet/inference/detection/wbc.py:94-198, restated without the N x N matrix:
    boxes in descending score order (stable: ties by ascending index); the highest remaining box heads a cluster of all
    remaining boxes with IoU > iou_thresh to it (:133-134, itself included since IoU(h, h) = 1; a zero-volume head has
    NaN IoU with itself and disappears with an empty cluster); score = sum(iou*w*s) / (sum(iou*w) + max(0, mean(n_exp) -
    n_found) * mean(iou*w) * missing_weight) (:185-191); box = sum(box * iou*w*s) / sum(iou*w*s) (:193-194); clusters whose
    score is not > score_thresh are dropped (:149)."""
    n = boxes.shape[0]
    if n == 0:
        return boxes.new_zeros((0, boxes.shape[1])), scores.new_zeros((0,))
    b = boxes.float()
    w = weights.float() * box_volume(b) if use_area else weights.float()
    order = torch.argsort(-scores.float(), stable=True)
    alive = torch.ones(n, dtype=torch.bool)
    out_b, out_s = [], []
    for oi in range(n):
        h = int(order[oi])
        if not alive[h]:
            continue
        pool = order[oi:][alive[order[oi:]]]
        iou = iou_union_3d(b[h:h + 1], b[pool])[0][0]
        m = iou > iou_thresh
        idx = pool[m]
        alive[h] = False                       # NaN IoU with itself: neither matched nor kept in the pool (:152-153)
        alive[idx] = False
        if idx.numel() == 0:
            continue
        iou_m = iou[m]
        msw = iou_m * w[idx]
        ms = msw * scores[idx].float()
        n_missing = torch.clamp(n_exp_preds[idx].float().mean() - idx.numel(), min=0.0)
        denom = msw.sum() + n_missing * msw.mean() * missing_weight
        sc = ms.sum() / denom
        if bool(sc > score_thresh):
            out_b.append((b[idx] * ms[:, None]).sum(0) / ms.sum())
            out_s.append(sc)
    if not out_b:
        return boxes.new_zeros((0, boxes.shape[1])), scores.new_zeros((0,))
    return torch.stack(out_b), torch.stack(out_s)


def batched_wbc(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh, use_area=False, missing_weight=1.0):
    """batched_wbc, wbc.py:30-91: per label (ascending), results concatenated; labels returned in the scores' dtype."""
    ob, os_, ol = [], [], []
    for lab in labels.unique():
        m = labels == lab
        b, s = wbc(boxes[m], scores[m], weights[m], n_exp_preds[m], iou_thresh, score_thresh, use_area, missing_weight)
        ob.append(b); os_.append(s); ol.append(torch.empty_like(s).fill_(float(lab)))
    if not ob:
        return boxes.new_zeros((0, boxes.shape[1])), scores.new_zeros((0,)), scores.new_zeros((0,))
    return torch.cat(ob), torch.cat(os_), torch.cat(ol)


# --------------------------------------------------------------------------- detection post-processing
def postprocess_single_image(boxes: torch.Tensor, probs: torch.Tensor, image_shape: Sequence[int],
                             num_classes: int, topk: int = 10000, score_thresh: Optional[float] = 0.0,
                             min_size: Optional[float] = 0.01, nms_thresh: float = 0.6,
                             detections_per_img: Optional[int] = 100, cuda_semantics: bool = True):
    """BaseRetinaNet.postprocess_detections_single_image, nndet/core/retina.py:332-379.

    boxes [A,6] decoded, probs [A,C] sigmoid.  Score ties -> ascending flat index.
    """
    boxes = clip_boxes_3d(boxes, image_shape)
    flat = probs.flatten()
    k = min(topk, boxes.size(0))
    idx = torch.argsort(-flat, stable=True)[:k]
    p = flat[idx]
    if score_thresh is not None:
        m = p > score_thresh
        p, idx = p[m], idx[m]
    a_idx = torch.div(idx, num_classes, rounding_mode="floor")
    labels = idx % num_classes
    b = boxes[a_idx]
    if min_size is not None:
        keep = keep_not_small(b, min_size)
        b, p, labels = b[keep], p[keep], labels[keep]
    keep = batched_nms(b, p, labels, nms_thresh, cuda_semantics)
    if detections_per_img is not None:
        keep = keep[:detections_per_img]
    return b[keep], p[keep], labels[keep]


# --------------------------------------------------------------------------- losses on sampled anchors
def giou_loss_sum(pred: torch.Tensor, target: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """GIoULoss(reduction='sum', loss_weight=1), nndet/losses/regression.py:118-162: -sum(diag(GIoU))."""
    if pred.numel() == 0:
        return torch.zeros(())
    return -1 * torch.diag(generalized_box_iou(pred, target, eps=eps)).sum()


def bce_onehot_mean(logits: torch.Tensor, labels: torch.Tensor, num_classes: int) -> torch.Tensor:
    """BCEWithLogitsLossOneHot(reduction='mean', smoothing=0), nndet/losses/classification.py:137-181:
    one-hot over C+1 classes, background column dropped, mean over N*C."""
    onehot = torch.zeros(labels.shape[0], num_classes + 1)
    onehot.scatter_(1, labels.long()[:, None], 1.0)
    return torch.nn.functional.binary_cross_entropy_with_logits(logits, onehot[:, 1:], reduction="mean")


def head_loss(box_logits: torch.Tensor, box_deltas: torch.Tensor, labels_batch: torch.Tensor,
              matched_batch: torch.Tensor, anchors_batch: torch.Tensor, pos_idx: torch.Tensor,
              neg_idx: torch.Tensor, num_classes: int):
    """DetectionHeadHNMNative.compute_loss given the sampled indices, nndet/arch/heads/comb.py:383-405."""
    losses = {}
    pred = decode_single(box_deltas[pos_idx], anchors_batch[pos_idx])
    if pos_idx.numel() > 0:
        losses["reg"] = giou_loss_sum(pred, matched_batch[pos_idx]) / max(1, pos_idx.numel())
    sel = torch.cat([pos_idx, neg_idx])
    losses["cls"] = bce_onehot_mean(box_logits[sel], labels_batch[sel], num_classes)
    return losses


def seg_loss(seg_logits: torch.Tensor, target: torch.Tensor, alpha: float = 0.5, smooth: float = 1e-5):
    """DiCESegmenterFgBg.compute_loss, nndet/arch/heads/segmenter.py:273-290,184-203 with
    SoftDiceLoss(batch_dice=True, do_bg=False, softmax), nndet/losses/segmentation.py:84-151.
    seg_logits [N,2,D,H,W]; target [N,D,H,W] (binarised: >0 -> 1)."""
    t = (target > 0).long()
    ce = torch.nn.functional.cross_entropy(seg_logits, t)
    p = torch.softmax(seg_logits, dim=1)
    onehot = torch.zeros_like(p).scatter_(1, t[:, None], 1.0)
    axes = [0, 2, 3, 4]
    tp = (p * onehot).sum(axes); fp = (p * (1 - onehot)).sum(axes); fn = ((1 - p) * onehot).sum(axes)
    dc = (2 * tp + smooth) / (2 * tp + fp + fn + smooth)
    return {"seg_ce": alpha * ce, "seg_dice": (1 - alpha) * (1 - dc[1:].mean())}
