"""Weighted box clustering on the B200: host mirror of nndet/inference/detection/wbc.py (`wbc` :94-160,
`batched_wbc` :30-91, `compute_cluster_consolidation` :163-198) -- same names, argument meaning and return values.

The reference builds the N x N IoU matrix and runs a python `while` loop with two `torch.where` host syncs per cluster;
here the clusters are the suppression sets of the on-device greedy scan (csrc/nms.cu, nnd_wbc3d_f32) and the weighted
consolidation is one atomics pass + one compaction block.  CUDA tensors only (no CPU fallback); 3-D boxes.
Deviation: equal scores are ordered by ascending index (torch.sort leaves ties unspecified, wbc.py:126).
"""
from typing import Tuple

import torch
from torch import Tensor

from .. import _lib as L


def wbc_device(boxes: Tensor, scores: Tensor, weights: Tensor, n_exp_preds: Tensor, iou_thresh: float, score_thresh: float,
               use_area: bool = True, missing_weight: float = 1.) -> Tuple[Tensor, Tensor, Tensor]:
    """Sync-free core: (boxes [N, 6] padded, scores [N] padded, count [1] int64 on the device)."""
    L.require_cuda(boxes, scores, weights, n_exp_preds)
    if boxes.dim() != 2 or boxes.shape[1] != 6:
        raise NotImplementedError("weighted box clustering is implemented for 3-D boxes [N, 6]")
    n = boxes.shape[0]
    dev = boxes.device
    ob = torch.empty((max(n, 1), 6), dtype=torch.float32, device=dev)
    os_ = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    if n == 0:
        return ob[:0], os_[:0], cnt
    lib = L.lib()
    lib.nnd_wbc_workspace_bytes.restype = L.c_size_t
    lib.nnd_wbc_workspace_bytes.argtypes = [L.c_ll]
    with torch.cuda.device(dev):
        b = boxes.detach().contiguous().float()
        s = scores.detach().contiguous().float()
        w = weights.detach().contiguous().float()
        ne = n_exp_preds.detach().contiguous().float()
        ws_bytes = lib.nnd_wbc_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        L.check(lib.nnd_wbc3d_f32(L.ptr(b), L.ptr(s), L.ptr(w), L.ptr(ne), L.c_ll(n), L.c_float(float(iou_thresh)),
                                  L.c_float(float(score_thresh)), L.c_int(1 if use_area else 0), L.c_float(float(missing_weight)),
                                  L.ptr(ob), L.ptr(os_), L.ptr(cnt), L.ptr(ws), L.c_size_t(ws_bytes), L.stream_ptr()),
                "nnd_wbc3d_f32")
    return ob, os_, cnt


def wbc(boxes: Tensor, scores: Tensor, weights: Tensor, n_exp_preds: Tensor, iou_thresh: float, score_thresh: float,
        use_area: bool = True, missing_weight: float = 1.) -> Tuple[Tensor, Tensor]:
    """nndet/inference/detection/wbc.py:94-160: consolidated boxes [K, 6] and scores [K], clusters in descending order of
    their highest-scoring member.  One host read (the cluster count)."""
    ob, os_, cnt = wbc_device(boxes, scores, weights, n_exp_preds, iou_thresh, score_thresh, use_area, missing_weight)
    k = int(cnt.item())
    return ob[:k].to(boxes.dtype), os_[:k].to(scores.dtype)


def batched_wbc(boxes: Tensor, scores: Tensor, labels: Tensor, weights: Tensor, iou_thresh: float, n_exp_preds: Tensor,
                score_thresh: float, use_area: bool = False, missing_weight: float = 1.) -> Tuple[Tensor, Tensor, Tensor]:
    """nndet/inference/detection/wbc.py:30-91: clustering per class (ascending label), results concatenated; the labels
    come back in the scores' dtype like the reference's `torch.empty_like(s).fill_(label)`.  All classes are enqueued
    before the single host read of their cluster counts."""
    L.require_cuda(boxes, scores, labels, weights, n_exp_preds)
    parts = []
    for label in labels.unique().tolist():
        m = labels == label
        parts.append((label, wbc_device(boxes[m], scores[m], weights[m], n_exp_preds[m], iou_thresh, score_thresh, use_area,
                                        missing_weight)))
    if not parts:
        return (torch.zeros((0, boxes.shape[1]), dtype=boxes.dtype, device=boxes.device),
                torch.zeros((0,), dtype=scores.dtype, device=scores.device),
                torch.zeros((0,), dtype=scores.dtype, device=scores.device))
    counts = torch.cat([p[1][2] for p in parts]).tolist()
    ob = torch.cat([p[1][0][:k] for p, k in zip(parts, counts)]).to(boxes.dtype)
    os_ = torch.cat([p[1][1][:k] for p, k in zip(parts, counts)]).to(scores.dtype)
    ol = torch.cat([torch.full((k,), float(p[0]), dtype=scores.dtype, device=scores.device) for p, k in zip(parts, counts)])
    return ob, os_, ol
