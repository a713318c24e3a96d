"""Hard-negative sampler with the reference's protocol (nndet/core/boxes/sampler.py:212-270) plus a sync-free
index interface used by the fused train step.  Random draws come from a counter hash (see csrc/sampler.cu)."""
from typing import List

import torch
from torch import Tensor

from . import engine as E


class HardNegativeSamplerBatched:
    def __init__(self, batch_size_per_image: int, positive_fraction: float, min_neg: int = 0, pool_size: float = 10):
        self.min_neg = min_neg
        self._batch_size_per_image = batch_size_per_image
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.pool_size = pool_size
        self.seed = 0
        self._plans = {}

    def plan(self, batch_size: int) -> E.SamplerPlan:
        if batch_size not in self._plans:
            self._plans[batch_size] = E.SamplerPlan(batch_size, self._batch_size_per_image, self.positive_fraction,
                                                    self.min_neg, self.pool_size)
        return self._plans[batch_size]

    def sample_indices(self, labels_batch: Tensor, fg_probs: Tensor, batch_size: int):
        """-> (counts int32[8], pos_idx int64[max_pos], neg_idx int64[max_neg]); valid prefixes counts[2], counts[3]."""
        self.seed = (self.seed + 1) & 0x7FFFFFFF
        return E.hnm_sample(labels_batch, fg_probs, self.plan(batch_size), self.seed)

    def __call__(self, target_labels: List[Tensor], fg_probs: Tensor):
        """Reference protocol: uint8 masks over the concatenated batch, wrapped in one-element lists (:267-270)."""
        labels = torch.cat(target_labels, dim=0).float()
        counts, pos, neg = self.sample_indices(labels, fg_probs, len(target_labels))
        c = counts.tolist()
        pm = torch.zeros_like(labels, dtype=torch.uint8); pm[pos[:c[2]]] = 1
        nm = torch.zeros_like(labels, dtype=torch.uint8); nm[neg[:c[3]]] = 1
        return [pm], [nm]
