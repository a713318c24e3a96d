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


def test_module_level_sweep_composes_prediction_and_search(tmp_path):
    """`helper.sweep` = `RetinaUNetModule.sweep` (ptmodule/retinaunet/base.py:747-815): predict_dir(save_state=True) on the validation
    cases, then the parameter search on the saved states; directory layout and returned inference plan as in the reference."""
    import torch
    from nndetection_b200.inference import helper as H
    pre = tmp_path / "preprocessed"
    data_dir, labels = pre / "imagesTr", pre / "labelsTr"
    data_dir.mkdir(parents=True); labels.mkdir()
    for ci in range(2):
        gen = torch.Generator().manual_seed(30 + ci)
        np.savez(data_dir / f"case_{ci}.npz", data=torch.rand(1, 40, 56, 48, generator=gen).numpy())
        H.save_pickle({"original_size_of_raw_data": (40, 56, 48), "itk_origin": (0., 0., 0.), "itk_spacing": (1., 1., 1.),
                       "itk_direction": (1., 0., 0., 0., 1., 0., 0., 0., 1.)}, data_dir / f"case_{ci}")
        np.savez(labels / f"case_{ci}_boxes_gt.npz", boxes=np.asarray([[4., 6., 14., 18., 8., 20.]], dtype=np.float32), classes=np.asarray([0]))
    plan = {"patch_size": (32, 32, 32), "batch_size": 4, "network_dim": 3, "transpose_backward": [0, 1, 2]}
    cfg = {"data": {"labels": {"0": "a", "1": "b"}}}
    save_dir = tmp_path / "model"
    save_dir.mkdir()
    # `sweep(run_prediction=True)` loads the real network from `<save_dir>/*last*.ckpt` (CUDA only); here its first half is done with
    # the stand-in detector exactly the way `sweep` calls `predict_dir`, then `sweep` runs the search on the saved states
    ens_cls = util.oracle_ensembler_cls()
    H.predict_dir(data_dir, save_dir / "sweep_predictions", cfg, plan, save_dir, model_fn=lambda *a: [{"model": util.FakeDetector(), "rank": 0}],
                  num_models=1, case_ids=["case_0", "case_1"], save_state=True, ensembler_cls=ens_cls, device="cpu")
    state = H.sweep(cfg, plan, save_dir, data_dir, case_ids=["case_0", "case_1"], run_prediction=False, eval_score_key="stand_in",
                    ensembler_cls=ens_cls, evaluator_cls=util.StandInEvaluator, sweep_device="cpu")
    assert sorted(p.name for p in (save_dir / "sweep_predictions").iterdir()) == [
        "case_0_boxes.pt", "case_0_properties.pkl", "case_1_boxes.pt", "case_1_properties.pkl"]
    assert sorted(p.name for p in (save_dir / "sweep").iterdir()) == [f"sweep_{k}.json" for k in sorted(
        ["model_iou", "model_nms_fn", "ensemble_iou", "model_score_thresh", "remove_small_boxes"])]
    assert set(state) == set(ens_cls.get_default_parameters()) and 1e-5 <= float(state["model_iou"]) <= 0.5
