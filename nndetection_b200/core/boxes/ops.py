"""Pairwise box metrics.  Mirrors nndet/core/boxes/ops.py: box_iou (:74-102), generalized_box_iou (:105-128),
box_center_dist (:262-287), box_area (:41-71)."""
import torch
from torch import Tensor

from . import engine as E


def box_iou(boxes1: Tensor, boxes2: Tensor, eps: float = 0) -> Tensor:
    if boxes1.numel() == 0 or boxes2.numel() == 0:
        return torch.tensor([]).to(boxes1)
    if boxes1.shape[-1] != 6:
        raise NotImplementedError("3-D boxes (x1, y1, x2, y2, z1, z2) only")
    return E.pairwise(boxes1, boxes2, 0, eps)


def generalized_box_iou(boxes1: Tensor, boxes2: Tensor, eps: float = 0) -> Tensor:
    if boxes1.nelement() == 0 or boxes2.nelement() == 0:
        return torch.tensor([]).to(boxes1)
    return E.pairwise(boxes1, boxes2, 1, eps)


def box_center_dist(boxes1: Tensor, boxes2: Tensor, euclidean: bool = True):
    if not euclidean:
        raise NotImplementedError
    c1 = torch.stack([(boxes1[:, 2] + boxes1[:, 0]) / 2., (boxes1[:, 3] + boxes1[:, 1]) / 2., (boxes1[:, 5] + boxes1[:, 4]) / 2.], 1)
    c2 = torch.stack([(boxes2[:, 2] + boxes2[:, 0]) / 2., (boxes2[:, 3] + boxes2[:, 1]) / 2., (boxes2[:, 5] + boxes2[:, 4]) / 2.], 1)
    return E.pairwise(boxes1, boxes2, 2), c1, c2
