"""Post-processing sweep (nndetection_b200/inference/sweeper.py) against tests/golden/sweeper.npz, written by scripts/gen_golden.py
sweeper after EXECUTING the reference's `BoxSweeper` on the same saved ensembler states and finding identical determined parameters
and identical scores for every swept value -- with the reference's real COCO evaluator and with the stand-in evaluator used here
(the reference's evaluator is not part of this repository).  The oracle's NMS / WBC stand in for the device kernels."""
import json

import numpy as np
import pytest

import tutil as util


def test_sweep_determines_the_reference_parameters(tmp_path):
    from nndetection_b200.inference.sweeper import BoxSweeper
    g = util.golden("sweeper")
    util.write_sweep_cases(tmp_path / "pred", tmp_path / "gt")
    ens_cls = util.oracle_ensembler_cls()
    sw = BoxSweeper(["class0", "class1"], tmp_path / "pred", tmp_path / "gt", "stand_in", ens_cls, save_dir=tmp_path / "sweep",
                    evaluator_cls=util.StandInEvaluator, device="cpu")
    state = sw.run_postprocessing_sweep()
    _, sweep = ens_cls.sweep_parameters()
    assert list(sweep.keys()) == ["model_iou", "model_nms_fn", "ensemble_iou", "model_score_thresh", "remove_small_boxes"]   # search order
    assert sum(len(v) for v in sweep.values()) == 28
    for k in ("model_iou", "ensemble_iou", "model_score_thresh", "remove_small_boxes"):
        assert float(state[k]) == float(g["state_" + k]), k
    assert state["model_nms_fn"].__name__ == str(g["state_model_nms_fn"])
    assert state["model_topk"] == 1000 and state["ensemble_nms_fn"] is util.o_wbc_ensemble        # untouched entries of the default state
    for name, values in sweep.items():
        ov = json.load(open(tmp_path / "sweep" / f"sweep_{name}.json"))
        scores = [float(eval(v["scores"], {"np": np})["stand_in"]) for k, v in ov.items() if not k.startswith("best_")]
        assert len(scores) == len(values)
        assert np.allclose(scores, g["scores_" + name], rtol=1e-6, atol=1e-9), name
        assert ov[f"best_{name}"]["value"] == str(values[int(np.argmax(scores))])                 # first best wins (np.argmax)


def test_sweeper_needs_an_evaluator(tmp_path):
    from nndetection_b200.inference.sweeper import BoxSweeper
    with pytest.raises(ImportError):
        BoxSweeper(["c"], tmp_path, tmp_path, "m", util.oracle_ensembler_cls())
