"""SURVEY 8f rows 1-2 composed ON THE DEVICE: `SlidingWindowPredictor` + `BoxEnsemblerSelective` with the sm_100a NMS / WBC kernels
(and, in the second test, the real Retina U-Net) against the same pipeline on the CPU with the oracle's NMS / WBC, whose host
logic is pinned on the executed reference (tests/test_predictor_cpu.py, tests/test_ensembler_cpu.py).

Both sides consume IDENTICAL per-tile detections (the device run's model outputs are recorded and replayed on the CPU side), so
what is gated is everything between `inference_step` and the case result: device-resident tiles, mirror TTA and box un-mirroring,
tile offsets and in-tile weights, per-model top-k / clip / small-box filter / weighted NMS kernel, ensemble top-k and the WBC kernel.
Tolerance: keep lists and labels exact, consolidated boxes / scores 1e-5 relative (fp32 atomics order inside the WBC kernel).

Strict since round 2 (round 1 carried a non-strict xfail: first device run).  Round-2 bisect (scripts/diag_inference.py on the B200):
all 787 case-level detections of the real-network case equal the CPU-pinned pipeline's row for row; see DESIGN.md section 2."""
import pytest
import torch

import tutil as util
from oracle import box_oracle as bo

pytestmark = pytest.mark.gpu


def _o_weighted_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
    keep = bo.batched_nms(boxes, scores * weights, labels, iou_thresh, cuda_semantics=False)
    return boxes[keep], scores[keep], labels[keep], torch.ones_like(weights)[keep]


def _o_wbc_ensemble(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh, *a, **k):
    return bo.batched_wbc(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh)


class _Recorder:
    """Wraps a detector: runs it, keeps a CPU copy of every call's detections (in call order)."""

    def __init__(self, model, run_on_host=False):
        self.model, self.run_on_host, self.calls = model, run_on_host, []

    def eval(self):
        return self

    def inference_step(self, images):
        assert images.is_cuda
        if self.run_on_host:               # a stand-in model whose arithmetic must not depend on the device
            out = self.model.inference_step(images.cpu())
            out = {k: [t.to(images.device) for t in out[k]] for k in ("pred_boxes", "pred_scores", "pred_labels")}
        else:
            out = self.model.inference_step(images)
        self.calls.append({k: [t.detach().cpu().clone() for t in out[k]] for k in ("pred_boxes", "pred_scores", "pred_labels")})
        return out


class _Replay:
    def __init__(self, calls):
        self.calls, self.i = calls, 0

    def eval(self):
        return self

    def inference_step(self, images):
        out = self.calls[self.i]
        self.i += 1
        assert len(out["pred_boxes"]) == images.shape[0]
        return {k: [t.clone() for t in v] for k, v in out.items()}


def _run_pair(model, case, crop, batch_size, run_on_host, params=None):
    from nndetection_b200.inference.ensembler import BoxEnsemblerSelective
    from nndetection_b200.inference.predictor import SlidingWindowPredictor
    params = dict(params or {})
    rec = _Recorder(model, run_on_host)
    dev = SlidingWindowPredictor(
        ensembler_fn=lambda c, properties=None: BoxEnsemblerSelective.from_case(c, properties, parameters=dict(params), device="cuda:0"),
        models=[rec], crop_size=crop, overlap=0.5, num_tta_transforms=8, batch_size=batch_size, device="cuda:0")
    out_d = dev.predict_case(case)["boxes"]
    cpu_params = dict(params, model_nms_fn=_o_weighted_nms_model, ensemble_nms_fn=_o_wbc_ensemble)
    cpu = SlidingWindowPredictor(
        ensembler_fn=lambda c, properties=None: BoxEnsemblerSelective.from_case(c, properties, parameters=dict(cpu_params)),
        models=[_Replay(rec.calls)], crop_size=crop, overlap=0.5, num_tta_transforms=8, batch_size=batch_size, device="cpu")
    out_c = cpu.predict_case({"data": case["data"].cpu()})["boxes"]
    return out_d, out_c, rec


def _compare(out_d, out_c):
    for k in ("pred_boxes", "pred_scores", "pred_labels"):
        assert out_d[k].is_cuda, k                                   # the case result never left the device
    assert out_d["pred_boxes"].shape == out_c["pred_boxes"].shape
    assert torch.equal(out_d["pred_labels"].cpu(), out_c["pred_labels"])
    assert torch.allclose(out_d["pred_scores"].cpu(), out_c["pred_scores"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(out_d["pred_boxes"].cpu(), out_c["pred_boxes"], rtol=1e-5, atol=1e-4)


def test_device_pipeline_with_stand_in_detector_matches_cpu_pipeline():
    g = torch.Generator().manual_seed(17)
    case = {"data": torch.rand(1, 40, 56, 48, generator=g)}
    out_d, out_c, rec = _run_pair(util.FakeDetector(), case, (32, 32, 32), 4, run_on_host=True)
    assert len(rec.calls) > 0 and len(rec.calls) % 8 == 0 and out_c["pred_boxes"].shape[0] > 0      # 8 mirror passes over all tiles
    _compare(out_d, out_c)


def test_device_pipeline_with_the_network_matches_cpu_pipeline():
    """Toy Retina U-Net (BASELINE config 1 architecture) over a 48 x 96 x 80 case: 3 x 3 x 3 symmetric-grid tiles of 32 x 64 x 64, 8 mirror passes.
    The network's detections are whatever the randomly initialised heads produce (<= 100 per tile, scores ~ 0.01): the low model
    score threshold keeps them all, so the NMS / WBC kernels see a few thousand boxes per pass."""
    from nndetection_b200.configs import make_plan
    from nndetection_b200.ptmodule import RetinaUNetV001
    arch, anc, patch, _ = make_plan("toy")
    torch.manual_seed(3)
    net = RetinaUNetV001.from_config_plan(None, arch, anc).cuda().eval()
    g = torch.Generator().manual_seed(23)
    case = {"data": torch.randn(1, 48, 96, 80, generator=g)}
    out_d, out_c, rec = _run_pair(net, case, patch, 2, run_on_host=False)
    assert len(rec.calls) == 8 * 14 and sum(len(b) for c in rec.calls for b in c["pred_boxes"]) > 0     # 3 x 3 x 3 symmetric tiles, batches of 2
    _compare(out_d, out_c)
    b = out_d["pred_boxes"]
    if b.shape[0]:
        lim = torch.tensor([48, 96, 48, 96, 80, 80], device=b.device, dtype=b.dtype)
        # consolidated boxes are score-weighted MEANS of clipped boxes (wbc.py:193-194): inside the case in exact arithmetic, but the
        # fp32 quotient sum(w * b) / sum(w) of boxes that all touch a border may land an ulp outside (this assert without the slack is
        # what failed in the round-1 driver run and again in the round-2 full-suite run: 96.00001 <= 96 -- the atomics' order decides)
        assert (b >= -1e-3).all() and (b <= lim + 1e-3).all(), (b.min(0).values, b.max(0).values)
