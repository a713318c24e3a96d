"""Box coder.  Mirrors BoxCoderND.decode / decode_single (nndet/core/boxes/coder.py:203-245, :90-155)."""
import math
from typing import List, Sequence

import torch
from torch import Tensor

from . import engine as E


class BoxCoderND:
    def __init__(self, weights: Sequence[float], bbox_xform_clip: float = math.log(1000. / 16)):
        if any(float(w) != 1.0 for w in weights):
            raise NotImplementedError("coder weights other than 1.0 (v001 uses (1.0,) * 6, ptmodule/retinaunet/base.py:388)")
        self.weights = tuple(weights)
        self.bbox_xform_clip = bbox_xform_clip

    def decode_single(self, rel_codes: Tensor, boxes: Tensor) -> Tensor:
        return E.decode_boxes(rel_codes, boxes, xform_clip=self.bbox_xform_clip)

    def decode(self, rel_codes: Tensor, boxes: List[Tensor]) -> Tensor:
        return E.decode_boxes(rel_codes, torch.cat(boxes, dim=0), xform_clip=self.bbox_xform_clip)
