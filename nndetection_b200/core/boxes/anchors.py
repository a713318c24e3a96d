"""Anchor generators on device.  Mirrors nndet/core/boxes/anchors.py: `AnchorGenerator3DS` (:472-559),
`forward` / cache / `get_num_acnhors_per_level` [sic] / `num_anchors_per_location` (:211-263), `get_anchor_generator` (:20-37)."""
from itertools import product
from typing import List, Sequence

import torch
from torch import Tensor

from . import engine as E


class AnchorGenerator3DS(torch.nn.Module):
    def __init__(self, width, height, depth, **kwargs):
        super().__init__()
        if not isinstance(width[0], Sequence):
            width = [(w,) for w in width]
        if not isinstance(height[0], Sequence):
            height = [(h,) for h in height]
        if not isinstance(depth[0], Sequence):
            depth = [(d,) for d in depth]
        self.width, self.height, self.depth = width, height, depth
        assert len(self.width) == len(self.height) == len(self.depth)
        self.cell_anchors = None
        self._cache = {}
        self._by_image = {}
        self.num_anchors_per_level: List[int] = None

    @staticmethod
    def generate_anchors(width, height, depth, dtype=torch.float, device="cpu") -> Tensor:
        """anchors.py:526-549: product(w, h, d) / 2 as (-w, -h, w, h, -d, d)."""
        s = torch.tensor(list(product(width, height, depth)), dtype=dtype, device=device) / 2
        return torch.stack([-s[:, 0], -s[:, 1], s[:, 0], s[:, 1], -s[:, 2], s[:, 2]], dim=1)

    def set_cell_anchors(self, dtype, device):
        if self.cell_anchors is None or self.cell_anchors[0].device != torch.device(device):
            self.cell_anchors = [self.generate_anchors(w, h, d, torch.float32, "cpu").to(device)
                                 for w, h, d in zip(self.width, self.height, self.depth)]

    def grid_anchors(self, grid_sizes, strides):
        assert len(grid_sizes) == len(strides) == len(self.cell_anchors)
        out = [E.anchor_grid(b, g, s) for g, s, b in zip(grid_sizes, strides, self.cell_anchors)]
        return out, [a.shape[0] for a in out]

    def forward(self, image_list: Tensor, feature_maps: List[Tensor]) -> List[Tensor]:
        """One anchor tensor per image (all images share the same storage; the reference re-concatenates and
        re-uploads 24 MB per image per step, anchors.py:231-237)."""
        grid_sizes = [tuple(int(v) for v in fm.shape[2:]) for fm in feature_maps]
        image_size = image_list.shape[2:]
        strides = [[int(i / s) for i, s in zip(image_size, g)] for g in grid_sizes]
        self.set_cell_anchors(torch.float32, feature_maps[0].device)
        key = str(grid_sizes + strides) + str(feature_maps[0].device)
        if key not in self._cache:
            per_fm, per_level = self.grid_anchors(grid_sizes, strides)
            self._cache[key] = (torch.cat(per_fm), per_level)
        anchors, self.num_anchors_per_level = self._cache[key]
        self._by_image[(tuple(int(v) for v in image_size), str(feature_maps[0].device))] = self._cache[key]
        return [anchors] * image_list.shape[0]

    def lookup(self, image_list: Tensor):
        """(anchors, anchors per level) of an image shape seen before, else None -- lets the caller start target
        assignment (which needs only anchors + ground truth) before / beside the network forward."""
        return self._by_image.get((tuple(int(v) for v in image_list.shape[2:]), str(image_list.device)))

    def num_anchors_per_location(self) -> List[int]:
        return [len(w) * len(h) * len(d) for w, h, d in zip(self.width, self.height, self.depth)]

    def get_num_acnhors_per_level(self) -> List[int]:
        if self.num_anchors_per_level is None:
            raise RuntimeError("Need to forward features maps before get_num_acnhors_per_level can be called")
        return self.num_anchors_per_level


def get_anchor_generator(dim: int, s_param: bool = False):
    if dim == 3 and s_param:
        return AnchorGenerator3DS
    raise NotImplementedError("nndetection_b200 implements AnchorGenerator3DS (the generator the v001 planner emits)")
