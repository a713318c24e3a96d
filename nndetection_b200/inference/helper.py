"""The caller and the on-disk formats on the inference side of the path: host mirror of `predict_dir`
(nndet/inference/helper.py:29-111), the checkpoint loaders (nndet/inference/loading.py:28-146), `get_predictor`
(nndet/ptmodule/retinaunet/base.py:697-745) and the small io helpers they use (`load_pickle` / `save_pickle`,
nndet/io/load.py:304-341; `get_case_id_from_path`, nndet/io/paths.py:147-181).  Same function names, argument meaning, file
names and file contents:

  <source_dir>/<case>.npz ["data"] (or <case>.npy), <case>.pkl          preprocessed case + properties            (read)
  <source_models>/*<identifier>*.ckpt                                   Lightning checkpoint: ["state_dict"]["model.<key>"] (read)
  <target_dir>/<case>_boxes.pkl                                         dict of numpy arrays pred_boxes / pred_scores / pred_labels
                                                                        + restore + the itk properties            (written)
  <target_dir>/<case>_boxes.pt, <case>_properties.pkl                   ensembler state when save_state=True      (written)

What changed underneath: the network, the predictor loop and the ensembler are the device-resident ones of this package; a case's
result leaves the GPU once, as the final <= ensemble_topk detections.  With `shard=(rank, world)` the tiles of every case are spread
over the ranks and rank 0 writes the files.
"""
import os
import pickle
from functools import partial
from pathlib import Path
from typing import Any, Callable, Dict, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .ensembler import BoxEnsemblerSelective
from .predictor import SlidingWindowPredictor

Pathlike = Union[str, Path]
CKPT_MODEL_PREFIX = "model."            # the LightningModule keeps the network in `self.model` (nndet/ptmodule/base_module.py:55-59)


# ------------------------------------------------------------------ nndet/io/load.py, nndet/io/paths.py, nndet/utils/tensor.py
def load_pickle(path: Pathlike, **kwargs) -> Any:
    """load.py:304-322: the suffix .pkl is appended unless the path already ends in .pkl / .pickle."""
    path = Path(path)
    if path.suffix not in (".pickle", ".pkl"):
        path = Path(str(path) + ".pkl")
    with open(path, "rb") as f:
        return pickle.load(f, **kwargs)


def save_pickle(data: Any, path: Pathlike, **kwargs) -> None:
    """load.py:325-341."""
    path = Path(path)
    if path.suffix not in (".pickle", ".pkl"):
        path = Path(str(path) + ".pkl")
    with open(str(path), "wb") as f:
        pickle.dump(data, f, **kwargs)


def get_case_id_from_path(file_path: Pathlike, remove_modality: bool = True) -> str:
    """paths.py:147-181: file name without its ending (.nii.gz counts as one ending); optionally without the `_0000` modality tag."""
    file_name = str(file_path).rsplit(os.path.sep, 1)[1]
    file_name = file_name.rsplit(".", 2)[0] if file_name.endswith(".nii.gz") else file_name.rsplit(".", 1)[0]
    return file_name[:-5] if remove_modality else file_name


def to_numpy(inp: Any) -> Any:
    """nndet/utils/tensor.py:90-111: tensors -> numpy arrays inside (nested) lists / tuples / dicts, everything else untouched."""
    if isinstance(inp, (tuple, list)):
        return type(inp)([to_numpy(i) for i in inp])
    if isinstance(inp, dict):
        return type(inp)({k: to_numpy(i) for k, i in inp.items()})
    if isinstance(inp, torch.Tensor):
        return inp.detach().cpu().numpy()
    return inp


# ------------------------------------------------------------------ checkpoints (nndet/inference/loading.py)
def network_state_from_checkpoint(path: Pathlike) -> Dict[str, torch.Tensor]:
    """The network's state_dict inside a reference Lightning checkpoint: `torch.load(path)["state_dict"]` (loading.py:96) holds the
    LightningModule's entries, the network's under `model.`; anything else in there (none in v001) is not the network's."""
    state = torch.load(str(path), map_location="cpu", weights_only=False)["state_dict"]
    return {k[len(CKPT_MODEL_PREFIX):]: v for k, v in state.items() if k.startswith(CKPT_MODEL_PREFIX)}


def save_checkpoint(model: torch.nn.Module, path: Pathlike, **extra) -> None:
    """Write the network so that the reference's `load_final_model` / `load_all_models` read it back (state_dict keys `model.<key>`,
    fp32 CPU tensors).  `extra` entries (epoch, global_step ...) are stored next to `state_dict` like Lightning does."""
    state = {CKPT_MODEL_PREFIX + k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
    torch.save(dict(extra, state_dict=state), str(path))


def _build_model(cfg: dict, plan: dict, path: Pathlike, device) -> torch.nn.Module:
    from ..ptmodule import RetinaUNetV001
    model = RetinaUNetV001.from_config_plan(cfg.get("model_cfg"), plan["architecture"], plan["anchors"])
    model.load_state_dict(network_state_from_checkpoint(path))         # strict: a checkpoint of another architecture raises
    model.float()
    model.eval()
    return model.to(device) if device is not None else model


def load_final_model(source_models: Pathlike, cfg: dict, plan: dict, num_models: int = 1, identifier: str = "last",
                     device: Optional[Union[str, torch.device]] = "cuda") -> Sequence[dict]:
    """loading.py:58-101: the one checkpoint whose name contains `identifier`."""
    assert num_models == 1, f"load_final_model only supports num_models=1, found {num_models}"
    names = [m for m in Path(source_models).glob("*.ckpt") if identifier in str(m.stem)]
    assert len(names) == 1, f"Found wrong number of models, {names} in {source_models} with {identifier}"
    return [{"model": _build_model(cfg, plan, names[0], device), "rank": 0}]


def load_all_models(source_models: Pathlike, cfg: dict, plan: dict, *args, device: Optional[Union[str, torch.device]] = "cuda",
                    **kwargs) -> Sequence[dict]:
    """loading.py:104-146: every checkpoint of the directory, rank = position."""
    names = list(Path(source_models).glob("*.ckpt"))
    if not names:
        raise RuntimeError(f"Did not find any models in {source_models}")
    return [{"model": _build_model(cfg, plan, p, device), "rank": r} for r, p in enumerate(names)]


def get_loader_fn(mode: str, **kwargs) -> Callable:
    """loading.py:28-33."""
    return load_all_models if mode.lower() == "all" else partial(load_final_model, identifier=mode, **kwargs)


# ------------------------------------------------------------------ predictor factory (nndet/ptmodule/retinaunet/base.py:697-745)
def get_predictor(plan: Dict, models: Sequence[Any], num_tta_transforms: Optional[int] = None, do_seg: bool = False,
                  ensembler_cls: Callable = BoxEnsemblerSelective, **kwargs) -> SlidingWindowPredictor:
    """Patch size and batch size from the plan, the box ensembler (`get_ensembler_cls(key="boxes", dim=3)`, base.py:677-695)
    parameterised by the plan's `inference_plan` (the sweep's result), 8 mirror passes for 3-D networks."""
    if plan.get("network_dim", 3) != 3:
        raise NotImplementedError("2-D networks are outside this path (the reference raises here too, base.py:742-743)")
    if do_seg:
        raise NotImplementedError("segmentation ensembling is not part of this row")
    inference_plan = plan.get("inference_plan", {})
    if num_tta_transforms is None:
        num_tta_transforms = 8
    return SlidingWindowPredictor(
        ensembler_fn=partial(ensembler_cls.from_case, parameters=inference_plan),
        models=models, crop_size=plan["patch_size"], num_tta_transforms=num_tta_transforms, batch_size=plan["batch_size"], **kwargs)


# ------------------------------------------------------------------ nndet/inference/helper.py:29-111
def predict_dir(source_dir: Pathlike, target_dir: Pathlike, cfg: dict, plan: dict, source_models: Pathlike,
                model_fn: Callable[[Path, dict, dict, int], Sequence[dict]] = load_final_model, num_models: Optional[int] = None,
                num_tta_transforms: Optional[int] = None, restore: bool = False, case_ids: Optional[Sequence[str]] = None,
                save_state: bool = False, **kwargs) -> SlidingWindowPredictor:
    """Predict all preprocessed cases of a directory; arguments as in the reference.  `kwargs` go to `get_predictor`
    (e.g. device=..., shard=(rank, world))."""
    source_dir, target_dir = Path(source_dir), Path(target_dir)
    models = model_fn(Path(source_models), cfg, plan, num_models) if num_models is not None else model_fn(Path(source_models), cfg, plan)
    predictor = get_predictor(plan=plan, models=[m["model"] for m in models], num_tta_transforms=num_tta_transforms, **kwargs)

    if case_ids is None:
        case_paths = [cp for cp in source_dir.glob("*.npz") if "_gt.npz" not in str(cp)]
    else:
        case_paths = [source_dir / f"{cid}.npz" for cid in case_ids]

    for path in case_paths:
        case_id = get_case_id_from_path(str(path), remove_modality=False)
        if path.is_file():
            case = np.load(str(path), allow_pickle=True)["data"]
        else:
            case = np.load(str(path)[:-4] + ".npy", allow_pickle=True)
        properties = load_pickle(path.parent / f"{case_id}.pkl")
        properties["transpose_backward"] = plan["transpose_backward"]
        if save_state:
            predictor.predict_case({"data": case}, properties, save_dir=target_dir, case_id=case_id, restore=restore)
        else:
            result = predictor.predict_case({"data": case}, properties, save_dir=None, case_id=None, restore=restore)
            if result is None:                   # tile sharding: only rank 0 holds the case result
                continue
            target_dir.mkdir(parents=True, exist_ok=True)
            for key, item in to_numpy(result).items():
                save_pickle(item, target_dir / f"{case_id}_{key}.pkl")
    return predictor


# ------------------------------------------------------------------ nndet/ptmodule/retinaunet/base.py:747-815
def sweep(cfg: dict, plan: dict, save_dir: Pathlike, train_data_dir: Pathlike, case_ids: Sequence[str], run_prediction: bool = True,
          sweep_ckpt: str = "last", eval_score_key: str = "mAP_IoU_0.10_0.50_0.05_MaxDet_100", ensembler_cls: Callable = BoxEnsemblerSelective,
          evaluator_cls: Optional[Callable] = None, sweep_device: str = "cuda", **kwargs) -> Dict[str, Any]:
    """`RetinaUNetModule.sweep`: predict the validation cases with the default post-processing and keep the ensembler states
    (`<save_dir>/sweep_predictions`), then search the post-processing parameters on them (`<save_dir>/sweep/sweep_<param>.json`).
    Returns the inference plan (to be stored as `plan["inference_plan"]`).  Ground truth: `<preprocessed>/labelsTr/<case>_boxes_gt.npz`
    next to `train_data_dir`, classes from `cfg["data"]["labels"]`, like the reference.  `kwargs` go to `predict_dir`."""
    from .sweeper import BoxSweeper
    save_dir, train_data_dir = Path(save_dir), Path(train_data_dir)
    processed_eval_labels = train_data_dir.parent / "labelsTr"
    (save_dir / "sweep").mkdir(parents=True, exist_ok=True)
    prediction_dir = save_dir / "sweep_predictions"
    prediction_dir.mkdir(parents=True, exist_ok=True)
    if run_prediction:
        predict_dir(source_dir=train_data_dir, target_dir=prediction_dir, cfg=cfg, plan=plan, source_models=save_dir, num_models=1,
                    num_tta_transforms=None, case_ids=case_ids, save_state=True, model_fn=get_loader_fn(mode=sweep_ckpt),
                    ensembler_cls=ensembler_cls, **kwargs)
    sweeper = BoxSweeper(classes=[item for _, item in cfg["data"]["labels"].items()], pred_dir=prediction_dir, gt_dir=processed_eval_labels,
                         target_metric=eval_score_key, ensembler_cls=ensembler_cls, save_dir=save_dir / "sweep",
                         evaluator_cls=evaluator_cls, device=sweep_device)
    return sweeper.run_postprocessing_sweep()
