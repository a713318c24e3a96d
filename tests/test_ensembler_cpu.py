"""Host logic of the device-resident case ensembler (nndetection_b200/inference/ensembler.py: tile offsets, in-tile weights, top-k /
threshold / clip / small-box filtering, per-model and cross-model bookkeeping) against the executed reference
`BoxEnsemblerSelective` (tests/golden/ensembler.npz, scripts/gen_golden.py ensembler).  The suppression steps are injected through
`parameters` exactly like the reference does (`model_nms_fn`, `ensemble_nms_fn`): on CPU the oracle's NMS / WBC stand in for the
sm_100a kernels (which have their own GPU parity tests: test_nms_gpu.py, test_wbc_gpu.py), so the comparison is bit-exact."""
import numpy as np
import torch

from oracle import box_oracle as bo
import tutil as util


def _o_weighted_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
    keep = bo.batched_nms(boxes, scores * weights, labels, iou_thresh, cuda_semantics=False)
    return boxes[keep], scores[keep], labels[keep], torch.ones_like(weights)[keep]


def _o_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
    keep = bo.batched_nms(boxes, scores, labels, iou_thresh, cuda_semantics=False)
    return boxes[keep], scores[keep], labels[keep], weights[keep]


def _o_wbc_ensemble(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh, *a, **k):
    return bo.batched_wbc(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh)


CASES = [(1, {"model_nms_fn": _o_weighted_nms_model, "ensemble_nms_fn": _o_wbc_ensemble}),
         (2, {"model_iou": 0.3, "ensemble_iou": 0.2, "model_score_thresh": 0.1, "remove_small_boxes": 2.0,
              "model_nms_fn": _o_nms_model, "ensemble_nms_fn": _o_wbc_ensemble})]


def test_case_ensembling_matches_reference_fixtures():
    from nndetection_b200.inference.ensembler import BoxEnsemblerSelective
    g = util.golden("ensembler")
    for ci, (seed, over) in enumerate(CASES):
        models, shape = util.synth_tile_predictions(seed)
        params = BoxEnsemblerSelective.get_default_parameters()
        params.update(over)
        ens = BoxEnsemblerSelective.from_case({"data": torch.zeros(1, *shape)}, properties={"itk_spacing": (1, 1, 1)}, parameters=params)
        for mi, batches in enumerate(models):
            ens.add_model(name=f"model0_t{mi}", model_weight=1.0 if mi == 0 else 0.7)
            for res, batch in batches:
                ens.process_batch(result=res, batch=batch)
        out = ens.get_case_result()
        assert np.array_equal(out["pred_boxes"].numpy(), g[f"c{ci}_boxes"])
        assert np.array_equal(out["pred_scores"].numpy(), g[f"c{ci}_scores"])
        assert np.array_equal(out["pred_labels"].numpy(), g[f"c{ci}_labels"])
        assert out["restore"] is False and out["itk_spacing"] == (1, 1, 1)


def test_defaults_and_protocol():
    from nndetection_b200.inference import ensembler as E
    d = E.BoxEnsemblerSelective.get_default_parameters()
    assert d["model_nms_fn"] is E.batched_weighted_nms_model and d["ensemble_nms_fn"] is E.batched_wbc_ensemble
    assert (d["model_iou"], d["model_topk"], d["model_detections_per_image"], d["ensemble_iou"], d["ensemble_topk"]) == (0.1, 1000, 100, 0.5, 1000)
    _, sweep = E.BoxEnsemblerSelective.sweep_parameters()
    assert abs(sweep["model_iou"][0] - 1e-5) < 1e-12 and len(sweep["model_score_thresh"]) == 7 and len(sweep["remove_small_boxes"]) == 7
    ens = E.BoxEnsemblerSelective(properties={"shape": (8, 8, 8)}, parameters=d)
    assert ens.add_model() == 1 and ens.add_model(model_weight=0.5) == 2
    try:
        ens.add_model(name=2)
        assert False
    except ValueError:
        pass
    # in-tile weight: 1 on the plateau, 0.5 in the corner (detection.py:1036-1060)
    w = E.BoxEnsemblerSelective._get_box_in_tile_weight(torch.tensor([[16., 24, 20], [0., 0, 0]]), (32, 48, 40))
    assert torch.allclose(w, torch.tensor([1.0, 0.5]))


def test_state_round_trip(tmp_path):
    """`<case>_boxes.pt` (ensembler/base.py:176-222, detection.py:276-318): saved state reloads into an ensembler that gives the
    same case result; only `model_topk` predictions per model are stored."""
    from nndetection_b200.inference.ensembler import BoxEnsemblerSelective
    seed, over = CASES[0]
    models, shape = util.synth_tile_predictions(seed)
    params = BoxEnsemblerSelective.get_default_parameters()
    params.update(over)
    params["model_topk"] = 150
    ens = BoxEnsemblerSelective.from_case({"data": torch.zeros(1, *shape)}, properties={}, parameters=params)
    for mi, batches in enumerate(models):
        ens.add_model(name=f"model0_t{mi}", model_weight=1.0 if mi == 0 else 0.7)
        for res, batch in batches:
            ens.process_batch(result=res, batch=batch)
    ref = ens.get_case_result()
    ens.save_state(tmp_path, "case_007")
    assert BoxEnsemblerSelective.get_case_ids(tmp_path) == ["case_007"]
    ens2 = BoxEnsemblerSelective.from_checkpoint(tmp_path, "case_007")
    assert all(v["boxes"][0].shape[0] <= 150 for v in ens2.model_results.values())
    out = ens2.get_case_result()
    for k in ("pred_boxes", "pred_scores", "pred_labels"):
        assert torch.equal(out[k], ref[k])


def test_restore_prediction_matches_the_executed_reference():
    """`get_case_result(restore=True)` (the reference's production path: scripts/predict.py:99 -> predict_dir(restore=True)): boxes from the
    preprocessed into the original image space, bit-exact against the executed `restore_detection` (nndet/inference/restore.py:30-66)
    on four axis orders / spacings / crops (tests/golden/restore.npz, scripts/gen_golden.py restore)."""
    from nndetection_b200.inference.ensembler import BoxEnsemblerSelective
    g = util.golden("restore")
    for i in range(int(g["n_cases"])):
        props = dict(transpose_backward=g[f"tb{i}"].tolist(), original_spacing=g[f"osp{i}"], spacing_after_resampling=g[f"rsp{i}"],
                     crop_bbox=[tuple(r) for r in g[f"crop{i}"].tolist()], itk_spacing=(1, 1, 1))
        ens = BoxEnsemblerSelective.from_case({"data": torch.zeros(1, 8, 8, 8)}, properties=props)
        out = ens.restore_prediction(torch.from_numpy(g[f"boxes{i}"]))
        assert out.dtype == torch.float32 and torch.equal(out, torch.from_numpy(g[f"restored{i}"])), i
    # through get_case_result: restored boxes + the flag, nothing raises
    models, shape = util.synth_tile_predictions(1)
    params = BoxEnsemblerSelective.get_default_parameters()
    params.update({"model_nms_fn": _o_weighted_nms_model, "ensemble_nms_fn": _o_wbc_ensemble})
    props = dict(transpose_backward=[2, 0, 1], original_spacing=(2.5, 0.7, 0.7), spacing_after_resampling=(0.8, 1.25, 0.9),
                 crop_bbox=[(3, 90), (11, 200), (7, 150)], itk_spacing=(1, 1, 1))
    ens = BoxEnsemblerSelective.from_case({"data": torch.zeros(1, *shape)}, properties=props, parameters=params)
    for mi, batches in enumerate(models):
        ens.add_model(name=f"m{mi}", model_weight=1.0)
        for res, batch in batches:
            ens.process_batch(result=res, batch=batch)
    plain, restored = ens.get_case_result(restore=False), ens.get_case_result(restore=True)
    assert restored["restore"] is True and plain["restore"] is False
    assert torch.equal(restored["pred_boxes"], ens.restore_prediction(plain["pred_boxes"]))
    assert torch.equal(restored["pred_scores"], plain["pred_scores"])
