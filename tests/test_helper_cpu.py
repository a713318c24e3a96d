"""The caller / on-disk formats on the inference side (nndetection_b200/inference/helper.py): `predict_dir`, checkpoint loading,
pickle helpers -- against tests/golden/helper.npz, written by scripts/gen_golden.py helper after it found this package's
`<case>_boxes.pkl` BYTE-IDENTICAL to the file the executed reference helpers (`to_numpy`, `save_pickle`) write for the reference
ensembler's case result, and a reference-layout Lightning checkpoint loading strictly in both directions.
A deterministic stand-in detector replaces the network, the oracle's NMS / WBC the device kernels (both have their own GPU tests)."""
import pickle

import numpy as np
import pytest
import torch

from oracle import box_oracle as bo, model_oracle as mo
import tutil as util

PROPS = {"original_size_of_raw_data": (40, 56, 48), "itk_origin": (0.0, 0.0, 0.0), "itk_spacing": (1.0, 1.0, 1.0),
         "itk_direction": (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)}


def _o_weighted_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
    keep = bo.batched_nms(boxes, scores * weights, labels, iou_thresh, cuda_semantics=False)
    return boxes[keep], scores[keep], labels[keep], torch.ones_like(weights)[keep]


def _o_wbc_ensemble(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh, *a, **k):
    return bo.batched_wbc(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh)


def _plan():
    return {"patch_size": (32, 32, 32), "batch_size": 4, "network_dim": 3, "transpose_backward": [0, 1, 2],
            "inference_plan": {"model_nms_fn": _o_weighted_nms_model, "ensemble_nms_fn": _o_wbc_ensemble}}


def _write_case(src, name="case_a"):
    from nndetection_b200.inference import helper as H
    gen = torch.Generator().manual_seed(17)
    np.savez(src / f"{name}.npz", data=torch.rand(1, 40, 56, 48, generator=gen).numpy())
    H.save_pickle(dict(PROPS), src / name)                  # suffix appended like the reference helper does
    assert (src / f"{name}.pkl").is_file() and H.load_pickle(src / name) == PROPS


def test_case_ids_and_pickle_suffix_rules():
    from nndetection_b200.inference import helper as H
    assert H.get_case_id_from_path("/data/Task000/imagesTr/case_001_0000.nii.gz") == "case_001"
    assert H.get_case_id_from_path("/data/Task000/imagesTr/case_001_0000.nii.gz", remove_modality=False) == "case_001_0000"
    assert H.get_case_id_from_path("/x/y/LUNA_17.npz", remove_modality=False) == "LUNA_17"
    assert H.get_case_id_from_path("/p/q.r/abc.def.npy", remove_modality=False) == "abc.def"


def test_predict_dir_writes_the_reference_files(tmp_path):
    from nndetection_b200.inference import helper as H
    g = util.golden("helper")
    src, dst = tmp_path / "src", tmp_path / "dst"
    src.mkdir()
    _write_case(src)
    np.savez(src / "case_a_gt.npz", data=np.zeros(1))                 # ground-truth files are skipped (helper.py:82)
    H.predict_dir(src, dst, cfg={}, plan=_plan(), source_models=tmp_path, model_fn=lambda *a: [{"model": util.FakeDetector(), "rank": 0}],
                  num_models=1, device="cpu")
    assert sorted(p.name for p in dst.iterdir()) == ["case_a_boxes.pkl"]
    with open(dst / "case_a_boxes.pkl", "rb") as f:
        res = pickle.load(f)
    assert list(res.keys()) == g["keys"].tolist()
    assert [str(getattr(v, "dtype", type(v).__name__)) for v in res.values()] == g["dtypes"].tolist()
    for k in ("pred_boxes", "pred_scores", "pred_labels"):
        assert isinstance(res[k], np.ndarray) and np.array_equal(res[k], g[k]), k
    assert res["restore"] is False and res["itk_spacing"] == PROPS["itk_spacing"]


def test_predict_dir_save_state_and_case_selection(tmp_path):
    """save_state=True keeps the ensembler state (`<case>_boxes.pt`, `<case>_properties.pkl`, predictor.py:180-185) instead of the final
    result; re-creating the ensembler from it gives the same case result; `case_ids` selects cases."""
    from nndetection_b200.inference import helper as H
    from nndetection_b200.inference.ensembler import BoxEnsemblerSelective
    g = util.golden("helper")
    src, dst = tmp_path / "src", tmp_path / "dst"
    src.mkdir()
    _write_case(src, "case_a")
    _write_case(src, "case_b")
    H.predict_dir(src, dst, cfg={}, plan=_plan(), source_models=tmp_path, model_fn=lambda *a: [{"model": util.FakeDetector(), "rank": 0}],
                  num_models=1, case_ids=["case_b"], save_state=True, device="cpu")
    assert sorted(p.name for p in dst.iterdir()) == ["case_b_boxes.pt", "case_b_properties.pkl"]
    assert BoxEnsemblerSelective.get_case_ids(dst) == ["case_b"]
    props = H.load_pickle(dst / "case_b_properties.pkl")
    assert props["transpose_backward"] == [0, 1, 2] and props["itk_origin"] == PROPS["itk_origin"]
    res = BoxEnsemblerSelective.from_checkpoint(dst, "case_b").get_case_result()
    for k in ("pred_boxes", "pred_scores", "pred_labels"):
        assert np.array_equal(res[k].numpy(), g[k]), k


def test_checkpoint_round_trip_in_the_reference_layout(tmp_path):
    """Lightning layout: {"state_dict": {"model.<network key>": tensor}} (nndet/inference/loading.py:96-97).  The oracle network has the
    reference's state_dict keys (asserted against the real model in scripts/gen_golden.py)."""
    from nndetection_b200.inference import helper as H

    class LM(torch.nn.Module):
        def __init__(self, net):
            super().__init__()
            self.model = net
    arch, anc, _, _ = mo.make_plan("tiny")
    plan = {"architecture": arch, "anchors": anc}
    torch.manual_seed(1)
    lm = LM(mo.RetinaUNetOracle(dict(arch), dict(anc)))
    torch.save({"state_dict": lm.state_dict(), "epoch": 3}, tmp_path / "model_last.ckpt")
    net = H.load_final_model(tmp_path, {"model_cfg": None}, plan, num_models=1, identifier="last", device=None)[0]["model"]
    assert not net.training
    for k, v in lm.model.state_dict().items():
        assert torch.equal(v, net.state_dict()[k]), k
    H.save_checkpoint(net, tmp_path / "model_best.ckpt", epoch=4)
    ck = torch.load(tmp_path / "model_best.ckpt", map_location="cpu", weights_only=False)
    assert ck["epoch"] == 4 and all(k.startswith("model.") for k in ck["state_dict"])
    t = lm.load_state_dict(ck["state_dict"])
    assert not t.missing_keys and not t.unexpected_keys
    assert len(H.get_loader_fn("all")(tmp_path, {"model_cfg": None}, plan, device=None)) == 2
    with pytest.raises(AssertionError):
        H.get_loader_fn("model")(tmp_path, {"model_cfg": None}, plan, num_models=1, device=None)       # two files match "model"
    arch2 = dict(arch, start_channels=16)
    with pytest.raises(RuntimeError):                                                                  # other architecture: strict load raises
        H.load_final_model(tmp_path, {"model_cfg": None}, {"architecture": arch2, "anchors": anc}, identifier="last", device=None)
