"""End-to-end learning check on the reference's toy task (scripts/generate_example.py: uniform noise + one cuboid of +0.4 intensity):
60 optimizer steps of the tiny Retina U-Net through `Trainer.train_step` (all kernels of the hot path, bf16 tensor-core convolutions,
fused SGD with the reference's warm-up / poly schedule), then inference on unseen patches.

Calibration: the CPU oracle (the reference's fp32 operators) trained on the SAME batches with the same hyper-parameters reaches
seg_dice 0.04, cls 0.25, reg -0.35 (means of the last 10 steps) and, on the 10 validation images below, a top-1 detection of class 0
with score 0.86-0.92 and IoU 0.37-0.62 (mean 0.51) with the cuboid.  The bounds here leave room for bf16 arithmetic and the different
sampling RNG.  Written after the round-1 GPU budget was spent -> non-strict xfail until its first B200 run (XPASS = it works)."""
import numpy as np
import pytest
import torch

import tutil as util
from oracle import box_oracle as bo

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first device run of the toy-task training check (round-1 GPU budget spent)")]


def test_tiny_network_learns_the_toy_task():
    from nndetection_b200.configs import make_plan
    from nndetection_b200.ptmodule import RetinaUNetV001
    from nndetection_b200.training import Trainer
    arch, anc, patch, bs = make_plan("tiny")
    torch.manual_seed(0)
    net = RetinaUNetV001.from_config_plan(None, arch, anc).cuda()
    trainer = Trainer(net, initial_lr=0.01, warm_iterations=10, warm_lr=1e-6, poly_gamma=0.9, num_iterations=200)
    hist = []
    for step in range(60):
        images, targets = util.toy_learning_batch(patch, bs, 1000 + step)
        tg = {"target_boxes": [b.cuda() for b in targets["target_boxes"]], "target_classes": [c.cuda() for c in targets["target_classes"]],
              "target_seg": targets["target_seg"].cuda()}
        losses, _ = trainer.train_step(images.cuda(), tg, evaluation=False)
        hist.append({k: float(v) for k, v in losses.items()})
        assert all(np.isfinite(v) for v in hist[-1].values()), (step, hist[-1])
    first = {k: np.mean([h[k] for h in hist[:5]]) for k in hist[0]}
    last = {k: np.mean([h[k] for h in hist[-10:]]) for k in hist[0]}
    assert last["seg_dice"] < 0.15 and last["seg_dice"] < 0.5 * first["seg_dice"], (first, last)
    assert last["cls"] < 0.45 and last["cls"] < first["cls"], (first, last)
    assert last["reg"] < -0.15, (first, last)

    net.eval()
    ious, scores, labels = [], [], []
    for v in range(5):
        images, targets = util.toy_learning_batch(patch, bs, 5000 + v)
        pred = net.inference_step(images.cuda())
        for i in range(bs):
            b, s, l = pred["pred_boxes"][i], pred["pred_scores"][i], pred["pred_labels"][i]
            if b.shape[0] == 0:
                ious.append(0.0); scores.append(0.0); labels.append(-1)
                continue
            ious.append(float(bo.box_iou(targets["target_boxes"][i], b[:1].float().cpu())[0, 0]))
            scores.append(float(s[0])); labels.append(int(l[0]))
    assert sum(i > 0.2 for i in ious) >= 8 and np.mean(ious) > 0.35, (ious, scores, labels)
    assert sum(l == 0 for l in labels) >= 8 and np.mean(scores) > 0.5, (ious, scores, labels)
