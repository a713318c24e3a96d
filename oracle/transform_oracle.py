"""TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy) of the reference's `pre_trafo` chain that runs inside every
training / validation step on the batch (nndet/ptmodule/retinaunet/base.py:114-141):

  FindInstances          nndet/io/transforms/instances.py:25-39    per sample: sorted unique instance ids > 0 (cast to int)
  Instances2Boxes        :42-89  -> instances_to_boxes :92-136     box = (min0-1, min1-1, max0+1, max1+1, min2-1, max2+1) over the
                                                                     voxel indices of the instance, float32; classes from the
                                                                     per-sample mapping dict (:176-190)
  Instances2Segmentation :211-262 -> instances_to_segmentation :265-301   out[instances == id] = mapping[id] (+1 with add_background)

Pinned against the executed reference by scripts/gen_golden.py (`transforms`), fixtures tests/golden/transforms.npz.
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np


def find_instances(target: np.ndarray) -> List[np.ndarray]:
    """FindInstances.forward, instances.py:31-39.  target [B, 1, ...]; ids are cast to int (truncation) before unique."""
    out = []
    for b in range(target.shape[0]):
        ids = np.unique(target[b].astype(np.int32))
        out.append(ids[ids > 0])
    return out


def instances_to_boxes(seg: np.ndarray, dim: int, instances: Sequence[int]) -> np.ndarray:
    """instances_to_boxes, instances.py:92-136: seg [1, D, H, W] (one sample incl. channel axis), last `dim` axes are spatial.
    NB the comparison `_seg == _idx` is done on the ORIGINAL (float) values against the int id."""
    boxes = []
    for idx in instances:
        pos = np.stack(np.nonzero(seg == idx), axis=1)[:, -dim:]
        mins, maxs = pos.min(0), pos.max(0)
        box = [mins[0] - 1, mins[1] - 1, maxs[0] + 1, maxs[1] + 1]
        if dim > 2:
            box += [mins[2] - 1, maxs[2] + 1]
        boxes.append(box)
    if not boxes:
        return np.zeros((1, 0), dtype=np.float32)          # torch.tensor([[]]) of the reference (:134)
    return np.asarray(boxes, dtype=np.float32)


def pre_trafo(target: np.ndarray, mappings: Sequence[Dict], add_background: bool = True
              ) -> Tuple[List[np.ndarray], List[np.ndarray], List[np.ndarray], np.ndarray]:
    """The whole chain on a batch: (present_instances, boxes, classes, semantic target)."""
    present = find_instances(target)
    boxes, classes = [], []
    sem = np.zeros_like(target)
    for b in range(target.shape[0]):
        mp = {int(k): int(v) for k, v in mappings[b].items()}
        boxes.append(instances_to_boxes(target[b], target[b].ndim - 1, present[b]))
        classes.append(np.asarray([mp[int(i)] for i in present[b]], dtype=np.int64))
        for i in present[b]:
            sem[b][target[b] == i] = mp[int(i)] + (1 if add_background else 0)
    return present, boxes, classes, sem
