"""Post-processing parameter sweep on saved ensembler states: host mirror of `BoxSweeper` (nndet/inference/sweeper.py:78-219) -- the
caller that re-runs the whole-case model NMS + ensemble WBC once per (parameter value, validation case): 6 + 2 + 6 + 7 + 7 = 28 settings
with `BoxEnsemblerSelective.sweep_parameters()`.  Same constructor, methods, search order (coordinate-wise: every
parameter in turn, the best value is kept for the following ones), tie rule (`np.argmax`: first best) and `sweep_<param>.json` files.

What changed underneath: each case's state is restored onto the device once per evaluation and the suppression kernels run there
(`device`, default "cuda"; the reference pins the sweep to the CPU, sweeper.py:53); only the final <= ensemble_topk detections of a
case come back to the host for the evaluator.  The evaluator itself (COCO-style matching / mAP, nndet/evaluator) is outside the hot
path: pass the reference's `BoxEvaluator` (or any class with the same `create` / `run_online_evaluation` / `finish_online_evaluation`
protocol) as `evaluator_cls`; when omitted it is imported from an installed reference.
"""
import json
import time
from pathlib import Path
from typing import Any, Callable, Dict, Optional, Sequence, Tuple, Union

import numpy as np

from .helper import to_numpy

Pathlike = Union[str, Path]


class BoxSweeper:
    def __init__(self, classes: Sequence[str], pred_dir: Pathlike, gt_dir: Pathlike, target_metric: str, ensembler_cls: Callable,
                 save_dir: Optional[Pathlike] = None, evaluator_cls: Optional[Callable] = None, device: str = "cuda"):
        """sweeper.py:27-58,79-110.  pred_dir: `<case>_boxes.pt` ensembler states (predict_dir(save_state=True));
        gt_dir: `<case>_boxes_gt.npz` with `boxes` and `classes`."""
        self.classes = classes
        self.save_dir = save_dir if save_dir is None else Path(save_dir)
        if self.save_dir is not None:
            self.save_dir.mkdir(parents=True, exist_ok=True)
        self.target_metric = target_metric
        self.device = device
        self.pred_dir, self.gt_dir = Path(pred_dir), Path(gt_dir)
        if evaluator_cls is None:
            try:
                from nndet.evaluator.registry import BoxEvaluator as evaluator_cls
            except ImportError as e:
                raise ImportError("BoxSweeper needs an evaluator: pass evaluator_cls (e.g. the reference's "
                                  "nndet.evaluator.registry.BoxEvaluator)") from e
        self.evaluator_cls = evaluator_cls
        self.ensembler_cls = ensembler_cls

    def run_postprocessing_sweep(self) -> Dict[str, Any]:
        """sweeper.py:112-141: the determined parameters (the plan's `inference_plan`).  Coordinate-wise: each parameter is searched
        with the winners of the earlier ones already fixed in `state`."""
        state, grid = self.ensembler_cls.sweep_parameters()
        for name in grid:
            state[name] = self.run_parameter(values=grid[name], param_name=name, state=state)[0]
        return state

    def run_parameter(self, values: Sequence[Any], param_name: str, state: Dict[str, Any]) -> Tuple[Any, float]:
        """sweeper.py:143-178: score every candidate value of one parameter, first best wins; the per-value report goes to
        `<save_dir>/sweep_<param_name>.json` (same keys as the reference's file + the seconds each evaluation took)."""
        report, scores = {}, []
        for candidate in values:
            started = time.perf_counter()
            metrics = self._evaluate_value(state=state, **{param_name: candidate})
            scores.append(metrics[self.target_metric])
            report[f"{param_name}_{candidate}".replace(".", "_")] = dict(
                state=str(state), overwrite={param_name: str(candidate)}, scores=str(metrics), seconds=time.perf_counter() - started)
        winner = int(np.argmax(scores))
        if self.save_dir is not None:
            report[f"best_{param_name}"] = dict(value=str(values[winner]), score=str(scores[winner]))
            (self.save_dir / f"sweep_{param_name}.json").write_text(json.dumps(report, indent=4))
        return values[winner], scores[winner]

    def _evaluate_value(self, state: Dict[str, Any], **overwrite) -> Dict[str, float]:
        """sweeper.py:180-216: every case's saved state is restored (onto `self.device`), post-processed with `state` + `overwrite`,
        and handed to a fresh evaluator together with its ground truth."""
        evaluator = self.evaluator_cls.create(classes=self.classes, fast=True, verbose=False, save_dir=None)
        settings = dict(state, **overwrite)
        for case_id in self.ensembler_cls.get_case_ids(self.pred_dir):
            ens = self.ensembler_cls.from_checkpoint(base_dir=self.pred_dir, case_id=case_id, device=self.device)
            ens.update_parameters(**settings)
            det = to_numpy(ens.get_case_result(restore=False))
            truth = np.load(str(self.gt_dir / f"{case_id}_boxes_gt.npz"), allow_pickle=True)
            evaluator.run_online_evaluation(pred_boxes=[det["pred_boxes"]], pred_classes=[det["pred_labels"]],
                                            pred_scores=[det["pred_scores"]], gt_boxes=[truth["boxes"]],
                                            gt_classes=[truth["classes"]], gt_ignore=None)
        return evaluator.finish_online_evaluation()[0]
