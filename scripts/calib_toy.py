"""Spread of the toy task's learning outcome (tests/test_zz_toy_training_gpu.py, gate 2) over equivalent arithmetic:
    python scripts/calib_toy.py       # TMA wgrad on / off x three initialisations, 60 and 100 optimizer steps, 20 unseen images"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import tutil as util
from oracle import box_oracle as bo
from nndetection_b200.arch import conv_ops as ops
from nndetection_b200.configs import make_plan
from nndetection_b200.ptmodule import RetinaUNetV001
from nndetection_b200.training import Trainer

arch, anc, patch, bs = make_plan("tiny")


def evaluate(net, n_batches):
    net.eval()
    ious, scores, labels, top5, anyk = [], [], [], [], []
    for v in range(n_batches):
        images, targets = util.toy_learning_batch(patch, bs, 5000 + v)
        pred = net.inference_step(images.cuda())
        for i in range(bs):
            b, s, l = pred["pred_boxes"][i], pred["pred_scores"][i], pred["pred_labels"][i]
            if b.shape[0] == 0:
                ious.append(0.0); scores.append(0.0); labels.append(-1); top5.append(0.0); anyk.append(0.0)
                continue
            allv = bo.box_iou(targets["target_boxes"][i], b.float().cpu())[0]
            top5.append(float(allv[:5].max())); anyk.append(float(allv.max()))
            ious.append(float(bo.box_iou(targets["target_boxes"][i], b[:1].float().cpu())[0, 0]))
            scores.append(float(s[0])); labels.append(int(l[0]))
    net.train()
    evaluate.top5, evaluate.anyk = top5, anyk
    return ious, scores, labels


for tma in (1,):
    for seed in (0, 1, 2):
        ops.set_wgrad_tma(tma)
        torch.manual_seed(seed)
        net = RetinaUNetV001.from_config_plan(None, arch, anc).cuda()
        trainer = Trainer(net, initial_lr=0.01, warm_iterations=10, warm_lr=1e-6, poly_gamma=0.9, num_iterations=200)
        hist = []
        for step in range(100):
            images, targets = util.toy_learning_batch(patch, bs, 1000 + step)
            tg = {"target_boxes": [b.cuda() for b in targets["target_boxes"]], "target_classes": [c.cuda() for c in targets["target_classes"]],
                  "target_seg": targets["target_seg"].cuda()}
            losses, _ = trainer.train_step(images.cuda(), tg, evaluation=False)
            hist.append({k: float(v.detach()) for k, v in losses.items()})
            if step + 1 in (60, 100):
                ious, scores, labels = evaluate(net, 10)
                last = {k: round(float(np.mean([h[k] for h in hist[-10:]])), 3) for k in hist[0]}
                print(f"tma={tma} seed={seed} steps={step + 1}: first10 n(iou>0.15)={sum(i > 0.15 for i in ious[:10])} mean={np.mean(ious[:10]):.3f} | "
                      f"20 images n(iou>0.15)={sum(i > 0.15 for i in ious)} mean iou={np.mean(ious):.3f} n(label 0)={sum(l == 0 for l in labels)} "
                      f"mean score={np.mean(scores):.3f} | top-5 best iou mean={np.mean(evaluate.top5):.3f} n>0.15={sum(v > 0.15 for v in evaluate.top5)} any-det best iou mean="
                      f"{np.mean(evaluate.anyk):.3f} n>0.15={sum(v > 0.15 for v in evaluate.anyk)} last10 {last}", flush=True)
