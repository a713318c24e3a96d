"""Instance segmentation -> detection targets on the device: host mirror of nndet/io/transforms/instances.py
(`FindInstances` :25-39, `Instances2Boxes` :42-89, `Instances2Segmentation` :211-262) as used by the training module's
`pre_trafo` (nndet/ptmodule/retinaunet/base.py:114-141).  Same class names, constructor arguments and batch-dict keys, so
`Compose(FindInstances(...), Instances2Boxes(...), Instances2Segmentation(...))` keeps working; underneath the three
transforms share ONE fused kernel launch (nnd_instances_to_targets) and ONE host read (the per-sample instance counts,
which the variable-length target lists need anyway) instead of `unique` / `nonzero` / `.item()` per instance.
CUDA tensors only, 3 spatial dimensions, instance ids < `max_instance_id` (default 1024).
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from .. import _lib as L

_CACHE_KEY = "_nnd_b200_instances"


def instances_to_targets(target: Tensor, mappings: Sequence[Dict], add_background: bool = True, max_instance_id: int = 1024
                         ) -> Tuple[List[Tensor], List[Tensor], List[Tensor], Tensor]:
    """(present_instances, boxes, classes, semantic target) of a batch.  target [B, 1, D, H, W]; mappings[b]: instance id ->
    class (keys may be str, instances.py:186).  Boxes (x1, y1, x2, y2, z1, z2) = (min0-1, min1-1, max0+1, max1+1, min2-1,
    max2+1) fp32 (:124-127); an image without instances yields the reference's `tensor([[]])` box and `tensor([])` class."""
    L.require_cuda(target)
    if target.dim() != 5 or target.shape[1] != 1:
        raise NotImplementedError("instances_to_targets expects [B, 1, D, H, W]")
    B, _, D, H, W = target.shape
    dev = target.device
    lib = L.lib()
    lib.nnd_instances_workspace_bytes.restype = L.c_size_t
    lut_cls = torch.full((B, max_instance_id), -1, dtype=torch.int32)
    for b, mp in enumerate(mappings):
        for k, v in mp.items():
            if 0 < int(k) < max_instance_id:
                lut_cls[b, int(k)] = int(v)
    lut_sem = torch.where(lut_cls >= 0, lut_cls + (1 if add_background else 0), lut_cls)
    cap = max(1, max(len(mp) for mp in mappings)) if len(mappings) else 1
    with torch.cuda.device(dev):
        lut_cls_d, lut_sem_d = lut_cls.to(dev, non_blocking=True), lut_sem.to(dev, non_blocking=True)
        t32 = target.detach().contiguous().float()
        sem = torch.empty_like(t32)
        ids = torch.empty((B, cap), dtype=torch.int32, device=dev)
        boxes = torch.empty((B, cap, 6), dtype=torch.float32, device=dev)
        classes = torch.empty((B, cap), dtype=torch.int64, device=dev)
        meta = torch.empty(B + 1, dtype=torch.int32, device=dev)            # counts [B] | error flags
        ws_bytes = lib.nnd_instances_workspace_bytes(L.c_int(B), L.c_int(max_instance_id))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        L.check(lib.nnd_instances_to_targets(L.ptr(t32), L.c_int(B), L.c_int(D), L.c_int(H), L.c_int(W), L.ptr(lut_sem_d),
                                             L.ptr(lut_cls_d), L.c_int(max_instance_id), L.c_int(cap), L.ptr(sem), L.ptr(ids),
                                             L.ptr(boxes), L.ptr(classes), L.ptr(meta), L.c_void_p(meta.data_ptr() + 4 * B),
                                             L.ptr(ws), L.c_size_t(ws_bytes), L.stream_ptr()), "nnd_instances_to_targets")
        m = meta.tolist()                                                    # the chain's one host read
    counts, err = m[:B], m[B]
    if err & 1:
        raise ValueError(f"instance ids must be integers in [0, {max_instance_id})")
    if err & 2 or any(c > cap for c in counts):
        raise KeyError("an instance id present in the target is missing from instance_mapping")   # reference: KeyError at :189
    present = [ids[b, :c].to(torch.int32) for b, c in enumerate(counts)]
    out_boxes = [boxes[b, :c] if c > 0 else torch.zeros((1, 0), dtype=torch.float32, device=dev) for b, c in enumerate(counts)]
    out_classes = [classes[b, :c] if c > 0 else torch.zeros((0,), dtype=torch.float32, device=dev) for b, c in enumerate(counts)]
    return present, out_boxes, out_classes, sem.to(target.dtype)


class _Transform(torch.nn.Module):
    """AbstractTransform protocol (nndet/io/transforms/base.py): called with the batch dict as keyword arguments."""

    def __init__(self, grad: bool = False, **kwargs):
        super().__init__()
        self.grad = grad

    def __call__(self, **data) -> dict:
        with torch.set_grad_enabled(self.grad):
            return self.forward(**data)


def _fused(data: dict, instance_key: str, map_key: str, add_background: bool = True):
    c = data.get(_CACHE_KEY)
    if c is None or c["key"] != (instance_key, map_key, add_background) or c["src"] is not data[instance_key]:
        res = instances_to_targets(data[instance_key], data[map_key], add_background)
        c = {"key": (instance_key, map_key, add_background), "src": data[instance_key], "res": res}
        data[_CACHE_KEY] = c
    return c["res"]


class FindInstances(_Transform):
    """instances.py:25-39.  When the batch carries `instance_mapping` the fused kernel runs here and the two following
    transforms reuse its outputs."""

    def __init__(self, instance_key: str, save_key: str = "present_instances", map_key: str = "instance_mapping", **kwargs):
        super().__init__(grad=False)
        self.instance_key, self.save_key, self.map_key = instance_key, save_key, map_key

    def forward(self, **data) -> dict:
        data[self.save_key] = _fused(data, self.instance_key, self.map_key)[0]
        return data


class Instances2Boxes(_Transform):
    """instances.py:42-89."""

    def __init__(self, instance_key: str, map_key: str, box_key: str, class_key: str, grad: bool = False,
                 present_instances: Optional[str] = None, **kwargs):
        super().__init__(grad=grad)
        self.instance_key, self.map_key, self.box_key, self.class_key = instance_key, map_key, box_key, class_key
        self.present_instances = present_instances

    def forward(self, **data) -> dict:
        _, boxes, classes, _ = _fused(data, self.instance_key, self.map_key)
        data[self.box_key], data[self.class_key] = list(boxes), list(classes)
        return data


class Instances2Segmentation(_Transform):
    """instances.py:211-262 (`seg_key=None` overwrites the instance map, as the training module does)."""

    def __init__(self, instance_key: str, map_key: str, seg_key: str = None, add_background: bool = True, grad: bool = False,
                 present_instances: Optional[str] = None):
        super().__init__(grad=grad)
        self.instance_key, self.map_key, self.add_background = instance_key, map_key, add_background
        self.seg_key = seg_key if seg_key is not None else instance_key
        self.present_instances = present_instances

    def forward(self, **data) -> dict:
        sem = _fused(data, self.instance_key, self.map_key, self.add_background)[3]
        data.pop(_CACHE_KEY, None)                 # last user of the shared result: do not leak it into the step
        data[self.seg_key] = sem
        return data
