"""ATSS matcher with the reference's callable protocol (nndet/core/boxes/matcher/base.py:13-66, atss.py:22-122)."""
from typing import Callable, Sequence, Tuple

import torch
from torch import Tensor

from . import engine as E
from .ops import box_iou


class ATSSMatcher:
    BELOW_LOW_THRESHOLD: int = -1
    BETWEEN_THRESHOLDS: int = -2

    def __init__(self, num_candidates: int, similarity_fn: Callable = box_iou, center_in_gt: bool = True):
        if center_in_gt:
            raise NotImplementedError("center_in_gt=True (v001 trains with False, nndet/conf/train/v001.yaml:105-107)")
        self.similarity_fn = similarity_fn
        self.num_candidates = num_candidates
        self.center_in_gt = center_in_gt

    def __call__(self, boxes: Tensor, anchors: Tensor, num_anchors_per_level: Sequence[int],
                 num_anchors_per_loc: int) -> Tuple[Tensor, Tensor]:
        """Single-image protocol of the reference: (match_quality_matrix [G, A], matches [A] int64)."""
        gt = E.GtBatch([boxes], [torch.zeros(boxes.shape[0], dtype=torch.int64)], anchors.device)
        matches = E.atss_match(gt, anchors, num_anchors_per_level, self.num_candidates * num_anchors_per_loc)
        if boxes.numel() == 0:
            return torch.tensor([]).to(anchors), matches
        return self.similarity_fn(boxes, anchors), matches

    def match_batch(self, gt: "E.GtBatch", anchors: Tensor, num_anchors_per_level, num_anchors_per_loc: int) -> Tensor:
        """Whole batch in one go (no [G, A] matrix): matches [B*A] int64."""
        return E.atss_match(gt, anchors, num_anchors_per_level, self.num_candidates * num_anchors_per_loc)
