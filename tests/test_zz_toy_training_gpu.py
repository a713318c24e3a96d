"""End-to-end check of the TRAINING LOOP on the reference's toy task (scripts/generate_example.py: uniform noise + one cuboid of +0.4
intensity): 60 optimizer steps of the tiny Retina U-Net through `Trainer.train_step` (every kernel of the hot path: bf16 tensor-core
convolutions, side-stream target assignment from step 2 on, direct gradient accumulation into the flat buffer, fused SGD with the
reference's warm-up / poly schedule, lazy bf16 weight re-packing), then inference on unseen patches.

Two kinds of gates:
(1) LOCK-STEP, deterministic: at steps 0, 1, 2, 10, 30, 59 the fp32 CPU oracle (the reference's operators, oracle/model_oracle.py) is
    loaded with the device net's PRE-step weights and evaluates the same batch with the device sampler's indices injected: ATSS labels
    bit-exact, the four losses within 5e-3 + 1e-2 relative (bf16 activations; measured <= 1.3e-3), parameter gradients with a median
    relative error <= 0.35 and none above 0.6 (random-ish early weights make bf16 gradients of a normalised network noisy: stock
    PyTorch bf16 autocast shows the same 0.1-0.3, tests/test_net_gpu.py; by step 20 the median is < 0.01), and after step 0 the
    weights equal nesterov-SGD(weight decay on non-norm parameters) applied to those gradients (1e-5).
(2) LEARNING, statistical: the losses fall and the detections on 10 unseen images find the cuboid.  Bounds from five runs of the
    same arithmetic (default / no side stream / torch SGD / autograd accumulation / mma.sync kernels on the B200, plus the oracle:
    gpurun_out r2 call 1) and 18 more (scripts/calib_toy.py: three initialisations x TMA kernels on / off x 60 / 100 steps, twice):
    last-10-step means seg_dice 0.03-0.06, cls 0.23-0.31, reg -0.33..-0.45, top-1 score 0.65-0.9, label always 0.
    The trajectories are chaotic (fp32 atomics order, bf16 rounding) and the TOP-1 box is bimodal: every image holds exactly one cuboid,
    so the patch-sized coarse anchor learns "there is an object" (IoU with the cuboid = its volume fraction, 0.02-0.03) and outranks
    the well-localised boxes in about half of all runs -- with the same seed, with either kernel set (top-1 IoU > 0.15 on 0 / 20, 5 / 20
    or 20 / 20 images).  What every run shows is the localisation itself: the best of the five highest-scoring boxes has IoU 0.58-0.62
    on average and > 0.15 on 20 / 20 images.  That is the gate; the top-1 gates of rounds 1 and 2 sat inside the spread (DESIGN.md
    section 2)."""
import numpy as np
import torch
import pytest

import tutil as util
from oracle import box_oracle as bo, model_oracle as mo

pytestmark = pytest.mark.gpu
CHECK_STEPS = (0, 1, 2, 10, 30, 59)


def _rel(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-20))


def test_tiny_network_learns_the_toy_task_in_lock_step_with_the_oracle():
    from nndetection_b200.configs import make_plan
    from nndetection_b200.ptmodule import RetinaUNetV001
    from nndetection_b200.training import Trainer, poly_lr
    arch, anc, patch, bs = make_plan("tiny")
    torch.manual_seed(0)
    net = RetinaUNetV001.from_config_plan(None, arch, anc).cuda()
    sched = dict(initial_lr=0.01, warm_iterations=10, warm_lr=1e-6, poly_gamma=0.9, num_iterations=200)
    trainer = Trainer(net, **sched)
    chk = mo.RetinaUNetOracle(dict(arch), dict(anc))
    hist, side_stream_steps = [], 0
    for step in range(60):
        images, targets = util.toy_learning_batch(patch, bs, 1000 + step)
        tg = {"target_boxes": [b.cuda() for b in targets["target_boxes"]], "target_classes": [c.cuda() for c in targets["target_classes"]],
              "target_seg": targets["target_seg"].cuda()}
        pre = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()} if step in CHECK_STEPS else None
        side_stream_steps += net.anchor_generator.lookup(images.cuda()) is not None
        losses, _ = trainer.train_step(images.cuda(), tg, evaluation=False)
        hist.append({k: float(v.detach()) for k, v in losses.items()})
        assert all(np.isfinite(v) for v in hist[-1].values()), (step, hist[-1])
        if pre is None:
            continue
        pos_idx, neg_idx, counts, labels, matches = net.last_sample
        cnt = counts.cpu().tolist()
        pos, neg = pos_idx[:cnt[2]].cpu(), neg_idx[:cnt[3]].cpu()
        chk.load_state_dict(pre)
        chk.train(); chk.zero_grad()
        lc, lb, _ = util.oracle_losses_with_indices(chk, images, targets, pos, neg)
        sum(lc.values()).backward()
        assert torch.equal(labels.cpu(), lb.float()), step
        for k in lc:
            o = float(lc[k].detach())
            assert abs(hist[-1][k] - o) <= 5e-3 + 1e-2 * abs(o), (step, k, hist[-1][k], o)
        named = dict(net.named_parameters())
        errs = {k: _rel(named[k].grad, p.grad) for k, p in chk.named_parameters() if p.grad is not None and float(p.grad.norm()) > 0}
        assert float(np.median(list(errs.values()))) <= 0.35, (step, sorted(errs.items(), key=lambda kv: -kv[1])[:5])
        assert max(errs.values()) <= 0.6, (step, sorted(errs.items(), key=lambda kv: -kv[1])[:5])
        if step == 0:            # first optimizer step: momentum buffer = gradient, nesterov update = lr * (g + 0.9 g)
            lr = poly_lr(0, **sched)
            for k, p in named.items():
                w0, g = pre[k].double(), p.grad.detach().cpu().double()
                g = g + (0.0 if k.endswith(("norm.weight", "norm.bias")) else 3e-5) * w0
                assert _rel(p.detach().cpu().double(), w0 - lr * 1.9 * g) <= 1e-5, k
    assert side_stream_steps >= 58            # target assignment ran beside the forward pass from the second step on
    first = {k: np.mean([h[k] for h in hist[:5]]) for k in hist[0]}
    last = {k: np.mean([h[k] for h in hist[-10:]]) for k in hist[0]}
    assert last["seg_dice"] < 0.15 and last["seg_dice"] < 0.5 * first["seg_dice"], (first, last)
    assert last["cls"] < 0.45 and last["cls"] < first["cls"], (first, last)
    assert last["reg"] < -0.15, (first, last)

    net.eval()
    ious, scores, labels_ = [], [], []            # ious: best IoU among the five highest-scoring detections of an image
    for v in range(5):
        images, targets = util.toy_learning_batch(patch, bs, 5000 + v)
        pred = net.inference_step(images.cuda())
        for i in range(bs):
            b, s, l = pred["pred_boxes"][i], pred["pred_scores"][i], pred["pred_labels"][i]
            if b.shape[0] == 0:
                ious.append(0.0); scores.append(0.0); labels_.append(-1)
                continue
            ious.append(float(bo.box_iou(targets["target_boxes"][i], b[:5].float().cpu())[0].max()))
            scores.append(float(s[0])); labels_.append(int(l[0]))
    assert sum(i > 0.15 for i in ious) >= 9 and np.mean(ious) > 0.4, (ious, scores, labels_)
    assert sum(l == 0 for l in labels_) >= 8 and np.mean(scores) > 0.5, (ious, scores, labels_)
