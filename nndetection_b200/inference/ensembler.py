"""Case-level box ensembling kept on the device: host mirror of `BoxEnsemblerSelective`
(nndet/inference/ensembler/detection.py:901-1130; base class methods `BoxEnsembler.postprocess_image` :166-217,
`_apply_offsets_to_boxes` :219-252, `get_case_result` :422-474, `BaseEnsembler.add_model`, ensembler/base.py:90-113) and of the
model- / ensemble-level suppression functions it is parameterised with (nndet/inference/detection/model.py:25-93,
nndet/inference/detection/ensemble.py:25-131).  Same class / method / parameter names and result dict.

What changed underneath: the reference moves every tile's predictions to the CPU (detection.py:1015-1017) and runs its CPU NMS /
python-loop WBC there; here the tiles' boxes stay where the model produced them, the bookkeeping is a handful of tensor ops, and
NMS / WBC are the sm_100a kernels (`nnd_nms3d_f32`, `nnd_wbc3d_f32`).  The suppression functions are looked up in `parameters`
exactly like in the reference (`model_nms_fn`, `ensemble_nms_fn`), so the sweep (`sweep_parameters`) can swap them.
`restore=True` (resampling back to the original image space via `nndet.inference.restore`, an ITK-side utility) is out of scope.
Deviation: equal scores are ordered by ascending index (torch.sort leaves ties unspecified, detection.py:194,1112).
"""
from collections import defaultdict
from typing import Any, Dict, Hashable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from ..core.boxes.nms import batched_nms, nms
from .wbc import batched_wbc


# ------------------------------------------------------------------ nndet/inference/detection/model.py
def batched_nms_model(boxes: Tensor, scores: Tensor, labels: Tensor, weights: Tensor, iou_thresh: float, *args, **kwargs
                      ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """model.py:25-55."""
    keep = batched_nms(boxes, scores, labels, iou_thresh)
    return boxes[keep], scores[keep], labels[keep], weights[keep]


def batched_weighted_nms_model(boxes: Tensor, scores: Tensor, labels: Tensor, weights: Tensor, iou_thresh: float, *args, **kwargs
                               ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """model.py:58-93: NMS ordered by score * weight; surviving boxes keep their score and get weight 1."""
    keep = batched_nms(boxes, scores * weights, labels, iou_thresh)
    return boxes[keep], scores[keep], labels[keep], torch.ones_like(weights)[keep]


# ------------------------------------------------------------------ nndet/inference/detection/ensemble.py
def batched_nms_ensemble(boxes: Tensor, scores: Tensor, labels: Tensor, weights: Tensor, iou_thresh: float, *args, **kwargs
                         ) -> Tuple[Tensor, Tensor, Tensor]:
    """ensemble.py:25-53."""
    keep = batched_nms(boxes, scores, labels, iou_thresh)
    return boxes[keep], scores[keep], labels[keep]


def batched_wbc_ensemble(boxes: Tensor, scores: Tensor, labels: Tensor, weights: Tensor, iou_thresh: float, n_exp_preds: Tensor,
                         score_thresh: float, *args, **kwargs) -> Tuple[Tensor, Tensor, Tensor]:
    """ensemble.py:56-91."""
    return batched_wbc(boxes, scores, labels, weights=weights, n_exp_preds=n_exp_preds, iou_thresh=iou_thresh,
                       score_thresh=score_thresh)


def wbc_nms_no_label_ensemble(boxes: Tensor, scores: Tensor, labels: Tensor, weights: Tensor, iou_thresh: float,
                              n_exp_preds: Tensor, score_thresh: float, *args, **kwargs) -> Tuple[Tensor, Tensor, Tensor]:
    """ensemble.py:94-131: WBC per class, then one class-agnostic NMS."""
    boxes, scores, labels = batched_wbc(boxes, scores, labels, weights=weights, n_exp_preds=n_exp_preds, iou_thresh=iou_thresh,
                                        score_thresh=score_thresh)
    keep = nms(boxes, scores, iou_thresh)
    return boxes[keep], scores[keep], labels[keep]


# ------------------------------------------------------------------ small box helpers (nndet/core/boxes/{ops,clip}.py)
def box_center(boxes: Tensor) -> Tensor:
    """ops.py:314-327."""
    return torch.stack([(boxes[:, 2] + boxes[:, 0]) / 2., (boxes[:, 3] + boxes[:, 1]) / 2., (boxes[:, 5] + boxes[:, 4]) / 2.], dim=1)


def clip_boxes_to_image(boxes: Tensor, img_shape: Sequence[int]) -> Tensor:
    """clip.py:83-101 (out of place): x in [0, s0], y in [0, s1], z in [0, s2]."""
    hi = torch.tensor([img_shape[0], img_shape[1], img_shape[0], img_shape[1], img_shape[2], img_shape[2]],
                      dtype=boxes.dtype, device=boxes.device)
    return torch.minimum(boxes.clamp(min=0), hi)


def remove_small_boxes(boxes: Tensor, min_size: float) -> Tensor:
    """ops.py:241-259: indices of the boxes whose three sides are all >= min_size."""
    keep = ((boxes[:, 2] - boxes[:, 0]) >= min_size) & ((boxes[:, 3] - boxes[:, 1]) >= min_size) & \
           ((boxes[:, 5] - boxes[:, 4]) >= min_size)
    return torch.where(keep)[0]


def _cat(ts: List[Tensor], like: Optional[Tensor] = None, width: Optional[int] = None) -> Tensor:
    """nndet/utils/tensor.py:146-155 (`cat`), with a defined result for "no predictions at all"."""
    if len(ts) == 0:
        shape = (0,) if width is None else (0, width)
        return torch.zeros(shape, dtype=torch.float32, device=like.device if like is not None else "cpu")
    return ts[0] if len(ts) == 1 else torch.cat(ts, dim=0)


class BoxEnsemblerSelective:
    """detection.py:901-1130.  Usage as in the reference's predictor (nndet/inference/predictor.py:237-306):
    `add_model(...)` per model / TTA pass, `process_batch(result, batch)` per tile batch, `get_case_result()` once."""
    ID = "boxes"

    def __init__(self, properties: Dict[str, Any], parameters: Dict[str, Any], box_key: str = 'pred_boxes',
                 score_key: str = 'pred_scores', label_key: str = 'pred_labels', data_key: str = 'data',
                 device: Optional[Union[torch.device, str]] = None, **kwargs):
        self.model_current = None
        self.model_results: Dict[Hashable, Dict[str, List[Tensor]]] = {}
        self.model_weights: Dict[Hashable, float] = {}
        self.properties = properties
        self.case_result: Optional[Dict] = None
        self.parameters = parameters
        self.parameters.update(kwargs)
        self.device = torch.device(device) if device is not None else None       # None: stay where the predictions are
        self.data_key, self.score_key, self.label_key, self.box_key = data_key, score_key, label_key, box_key

    @classmethod
    def get_default_parameters(cls) -> Dict[str, Any]:
        """detection.py:940-973."""
        return {
            "model_iou": 0.1, "model_nms_fn": batched_weighted_nms_model, "model_score_thresh": 0.0, "model_topk": 1000,
            "model_detections_per_image": 100,
            "ensemble_iou": 0.5, "ensemble_nms_fn": batched_wbc_ensemble, "ensemble_topk": 1000, "remove_small_boxes": 1e-2,
            "ensemble_score_thresh": 0.0,
        }

    @classmethod
    def sweep_parameters(cls) -> Tuple[Dict[str, Any], Dict[str, Sequence[Any]]]:
        """detection.py:975-995."""
        iou_threshs = np.linspace(0.0, 0.5, 6)
        iou_threshs[0] = 1e-5
        small_boxes_thresh = [1e-2] + np.linspace(2., 7., 6).tolist()
        return cls.get_default_parameters(), {
            "model_iou": iou_threshs, "model_nms_fn": [batched_weighted_nms_model, batched_nms_model],
            "ensemble_iou": iou_threshs, "model_score_thresh": [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6],
            "remove_small_boxes": small_boxes_thresh,
        }

    @classmethod
    def from_case(cls, case: Dict, properties: Optional[Dict] = None, parameters: Optional[Dict] = None,
                  box_key: str = 'pred_boxes', score_key: str = 'pred_scores', label_key: str = 'pred_labels',
                  data_key: str = 'data', device=None, **kwargs):
        """detection.py:75-129: default parameters, `shape` from the case (channel axis removed); other properties pass through."""
        _parameters = cls.get_default_parameters()
        _parameters.update(parameters or {})
        _properties = dict(properties or {})
        _properties["shape"] = tuple(case[data_key].shape[1:])
        return cls(properties=_properties, parameters=_parameters, box_key=box_key, score_key=score_key, label_key=label_key,
                   data_key=data_key, device=device, **kwargs)

    def update_parameters(self, **parameters) -> None:
        """ensembler/base.py:153-160 (the sweep overwrites one parameter at a time)."""
        self.parameters.update(parameters)

    def add_model(self, name: Optional[Hashable] = None, model_weight: Optional[float] = None) -> Hashable:
        """ensembler/base.py:90-113."""
        if name is None:
            name = len(self.model_weights) + 1
        if name in self.model_results:
            raise ValueError(f"Invalid model name, model {name} is already present")
        self.model_weights[name] = 1.0 if model_weight is None else model_weight
        self.model_results[name] = defaultdict(list)
        self.model_current = name
        return name

    @staticmethod
    def _get_box_in_tile_weight(box_centers: Tensor, tile_size: Sequence[int]) -> Tensor:
        """detection.py:1036-1060: weight 1 on a plateau around the tile centre, linearly down to 0.5 in the corners."""
        plateau_length = 0.5
        if box_centers.numel() == 0:
            return box_centers.new_zeros((0,))
        tile_center = torch.tensor(tile_size).to(box_centers) / 2.
        max_dist = tile_center.norm(p=2)
        boxes_dist = (box_centers - tile_center[None]).norm(p=2, dim=1)
        return -(boxes_dist / max_dist - plateau_length).clamp_(min=0) + 1

    @staticmethod
    def _apply_offsets_to_boxes(boxes: List[Tensor], tile_offset: Sequence[Sequence[int]]) -> List[Tensor]:
        """detection.py:219-252: (x1, y1, x2, y2, z1, z2) += (o0, o1, o0, o1, o2, o2)."""
        out = []
        for img_boxes, offset in zip(boxes, tile_offset):
            if img_boxes.nelement() == 0:
                out.append(img_boxes)
                continue
            o = torch.tensor([offset[0], offset[1], offset[0], offset[1], offset[2], offset[2]]).to(img_boxes)
            out.append(img_boxes + o[None])
        return out

    @torch.no_grad()
    def process_batch(self, result: Dict, batch: Dict):
        """detection.py:998-1034 without the `.cpu()` round trip."""
        boxes = [r.float() for r in result[self.box_key]]
        scores = [r.float() for r in result[self.score_key]]
        labels = [r.float() for r in result[self.label_key]]
        centers = [box_center(b) if b.numel() > 0 else b.new_zeros((0,)) for b in boxes]
        tile_origins = [to for to in zip(*batch["tile_origin"])]
        tile_size = batch[self.data_key].shape[2:]
        weights = [self._get_box_in_tile_weight(c, tile_size) * self.model_weights[self.model_current] for c in centers]
        boxes = self._apply_offsets_to_boxes(boxes, [[int(v) for v in o] for o in tile_origins])
        res = self.model_results[self.model_current]
        res["boxes"].extend(boxes); res["scores"].extend(scores); res["labels"].extend(labels); res["weights"].extend(weights)

    def postprocess_image(self, boxes: Tensor, probs: Tensor, labels: Tensor, weights: Tensor, shape: Optional[Tuple[int]] = None
                          ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """detection.py:166-217: top-k -> score threshold -> clip -> remove small boxes -> model NMS -> detections per image."""
        idx_sorted = torch.argsort(probs, descending=True, stable=True)[:self.parameters["model_topk"]]
        idx_sorted = idx_sorted[probs[idx_sorted] > self.parameters["model_score_thresh"]]
        b, p, l, w = boxes[idx_sorted], probs[idx_sorted], labels[idx_sorted], weights[idx_sorted]
        b = clip_boxes_to_image(b, shape)
        keep = remove_small_boxes(b, min_size=self.parameters["remove_small_boxes"])
        b, p, l, w = b[keep], p[keep], l[keep], w[keep]
        _b, _p, _l, _w = self.parameters["model_nms_fn"](boxes=b, scores=p, labels=l, weights=w,
                                                        iou_thresh=self.parameters["model_iou"])
        n = self.parameters.get("model_detections_per_image", 1000)
        return _b[:n], _p[:n], _l[:n], _w[:n]

    def process_model(self, name: Hashable) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """detection.py:1062-1089."""
        r = self.model_results[name]
        like = r["boxes"][0] if r["boxes"] else None
        boxes, probs = _cat(r["boxes"], like, 6), _cat(r["scores"], like)
        labels, weights = _cat(r["labels"], like), _cat(r["weights"], like)
        if self.device is not None:
            boxes, probs, labels, weights = (t.to(self.device) for t in (boxes, probs, labels, weights))
        return self.postprocess_image(boxes=boxes, probs=probs, labels=labels, weights=weights,
                                      shape=tuple(self.properties["shape"]))

    def process_ensemble(self, boxes: List[Tensor], probs: List[Tensor], labels: List[Tensor], weights: List[Tensor]
                         ) -> Tuple[Tensor, Tensor, Tensor]:
        """detection.py:1091-1130 (results stay on the device)."""
        num_models = len(boxes)
        boxes, probs, labels, weights = _cat(boxes), _cat(probs), _cat(labels), _cat(weights)
        idx = torch.argsort(probs, descending=True, stable=True)[:self.parameters["ensemble_topk"]]
        boxes, probs, labels, weights = boxes[idx], probs[idx], labels[idx], weights[idx]
        n_exp_preds = torch.tensor([num_models] * len(boxes)).to(boxes)
        return self.parameters["ensemble_nms_fn"](boxes, probs, labels, weights=weights, iou_thresh=self.parameters["ensemble_iou"],
                                                  n_exp_preds=n_exp_preds, score_thresh=self.parameters["ensemble_score_thresh"])

    # ---- on-disk state (ensembler/base.py:176-222, detection.py:276-318,1132-1163): `<case>_boxes.pt`, same keys
    def save_state(self, target_dir, name: str, **kwargs):
        """Per model only the `model_topk` best predictions are kept (detection.py:1150-1161), tensors are stored on the CPU (the
        reference's note at :288-290), keys as written by `BaseEnsembler.save_state` + `BoxEnsembler.save_state`."""
        from pathlib import Path
        for model, r in self.model_results.items():
            like = r["boxes"][0] if r["boxes"] else None
            boxes, probs = _cat(list(r["boxes"]), like, 6), _cat(list(r["scores"]), like)
            labels, weights = _cat(list(r["labels"]), like), _cat(list(r["weights"]), like)
            if len(probs) > self.parameters["model_topk"]:
                idx = torch.argsort(probs, descending=True, stable=True)[:self.parameters["model_topk"]]
                boxes, probs, labels, weights = boxes[idx], probs[idx], labels[idx], weights[idx]
            r["boxes"], r["scores"], r["labels"], r["weights"] = [boxes.cpu()], [probs.cpu()], [labels.cpu()], [weights.cpu()]
        state = dict(kwargs)
        state.update(properties=self.properties, parameters=self.parameters, model_current=self.model_current,
                     model_results={k: dict(v) for k, v in self.model_results.items()}, model_weights=self.model_weights,
                     case_result=self.case_result, score_key=self.score_key, label_key=self.label_key, box_key=self.box_key,
                     data_key=self.data_key, overlap_map=None)
        with open(Path(target_dir) / f"{name}_{self.ID}.pt", "wb") as f:
            torch.save(state, f)

    @classmethod
    def from_checkpoint(cls, base_dir, case_id: str, **kwargs):
        """detection.py:304-318."""
        from pathlib import Path
        ckp = torch.load(str(Path(base_dir) / f"{case_id}_{cls.ID}.pt"), weights_only=False)
        t = cls(properties=ckp["properties"], parameters=ckp["parameters"], box_key=ckp["box_key"], score_key=ckp["score_key"],
                label_key=ckp["label_key"], data_key=ckp["data_key"], **kwargs)
        for key, item in ckp.items():
            if key == "model_results":
                item = {k: defaultdict(list, {kk: (list(vv) if isinstance(vv, (list, tuple)) else [vv]) for kk, vv in v.items()})
                        for k, v in item.items()}
            if key != "overlap_map":
                setattr(t, key, item)
        return t

    @classmethod
    def get_case_ids(cls, base_dir) -> List[str]:
        """ensembler/base.py:224-227."""
        from pathlib import Path
        return [c.stem.rsplit(f"_{cls.ID}", 1)[0] for c in Path(base_dir).glob(f"*_{cls.ID}.pt")]

    def restore_prediction(self, boxes: Tensor) -> Tensor:
        """detection.py:254-274 + nndet/inference/restore.py:30-66 (`restore_detection`) + core/boxes/ops.py:330-374: boxes from the
        preprocessed space into the original image space -- axes permuted by `transpose_backward`, scaled by resampled / original
        spacing, shifted by the crop offset.  At most `ensemble_topk` boxes: the arithmetic is the reference's numpy float64 on the
        host (one small D2H copy), the result returns to the boxes' device and dtype."""
        import numpy as np
        pr = self.properties
        tb = [int(d) for d in pr["transpose_backward"]]
        b = boxes.detach().cpu().numpy()
        if 2 * len(tb) != b.shape[1]:
            raise TypeError(f"Need same number of dimensions, found dims {tb} but boxes with shape {b.shape}")
        pairs = [[0, 2], [1, 3], [4, 5]]
        axis = [pairs[tb[0]][0], pairs[tb[1]][0], pairs[tb[0]][1], pairs[tb[1]][1]]
        for d in tb[2:]:
            axis.extend(pairs[d])
        b = b[:, axis]
        expand = [0, 1, 0, 1] + ([2, 2] if len(tb) == 3 else [])
        scaling = np.asarray(pr["spacing_after_resampling"])[tb] / np.asarray(pr["original_spacing"])
        offset = np.asarray([i[0] for i in pr["crop_bbox"]])
        b = b * scaling[None][:, expand] + offset[None][:, expand]
        return torch.from_numpy(b).to(device=boxes.device, dtype=boxes.dtype)

    @torch.no_grad()
    def get_case_result(self, restore: bool = False, names: Optional[Sequence[Hashable]] = None) -> Dict[str, Any]:
        """detection.py:422-474."""
        names = list(self.model_results.keys()) if names is None else names
        per_model = [self.process_model(name) for name in names]
        boxes, probs, labels = self.process_ensemble(boxes=[m[0] for m in per_model], probs=[m[1] for m in per_model],
                                                     labels=[m[2] for m in per_model], weights=[m[3] for m in per_model])
        if restore:
            boxes = self.restore_prediction(boxes)
        out = {"pred_boxes": boxes, "pred_scores": probs, "pred_labels": labels, "restore": restore}
        for k in ("original_size_of_raw_data", "itk_origin", "itk_spacing", "itk_direction"):
            if k in self.properties:
                out[k] = self.properties[k]
        return out
