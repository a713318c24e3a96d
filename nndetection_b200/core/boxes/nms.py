"""Host mirror of nndet/core/boxes/nms.py (nms :56-78, batched_nms :81-106) on top of the sm_100a NMS."""
import torch
from torch import Tensor

from ... import _C


def nms(boxes: Tensor, scores: Tensor, iou_threshold: float) -> Tensor:
    """keep indices (int64, descending score) of greedy NMS; boxes [N, 4|6] (x1, y1, x2, y2, (z1, z2)).

    Reference: nndet/core/boxes/nms.py:56-78 (autocast off, inputs cast to fp32).  CUDA tensors only --
    the reference's `nms_cpu` fallback has no counterpart here by design.
    """
    return _C.nms(boxes.float(), scores.float(), iou_threshold)


def batched_nms(boxes: Tensor, scores: Tensor, idxs: Tensor, iou_threshold: float) -> Tensor:
    """Per-class NMS through the coordinate-offset trick, nndet/core/boxes/nms.py:81-106:
    offset = idx * (max_coordinate + 1) evaluated in the boxes' dtype."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + 1)
    return nms(boxes + offsets[:, None], scores, iou_threshold)
