"""Bisect aid for the toy-task learning check (tests/test_zz_toy_training_gpu.py), run on the GPU box:

  python scripts/diag_toy.py [--steps 60] [--variant default|noside|torchsgd|nodirect|mma] > gpurun_out/diag_toy_<variant>.txt

Per step: the four losses of `Trainer.train_step` on the device next to the fp32 CPU oracle TRAINED ON THE SAME BATCHES (own
weights, torch.optim.SGD, the reference's schedule), and -- at the listed check steps -- a lock-step comparison: the oracle is
loaded with the device net's PRE-step weights, evaluates the same batch with the device sampler's indices injected, and its
losses / ATSS labels / per-parameter gradients / post-step weights are compared with the device's.
TEST INFRASTRUCTURE (imports oracle/)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tutil as util                                      # noqa: E402
from oracle import box_oracle as bo, model_oracle as mo   # noqa: E402


def rel_err(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-20))


def oracle_losses_with_indices(orc, images, targets, pos, neg):
    pred, anchors, pseg = orc(images)
    labels, matched = [], []
    for a, gb, gc in zip(anchors, targets["target_boxes"], targets["target_classes"]):
        _, m = bo.atss_match(gb, a, orc.per_level, orc.apos, orc.num_candidates)
        l, mb = bo.assign_targets(m, gb, gc, a.shape[0])
        labels.append(l); matched.append(mb)
    lb, mb, ab = torch.cat(labels), torch.cat(matched), torch.cat(anchors)
    losses = bo.head_loss(pred["box_logits"], pred["box_deltas"], lb, mb, ab, pos, neg, orc.num_classes)
    losses.update(bo.seg_loss(pseg["seg_logits"], targets["target_seg"]))
    return losses, lb, pred


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--variant", default="default")
    ap.add_argument("--plan", default="tiny")
    ap.add_argument("--checks", default="0,1,2,3,5,10,20,40,59")
    ap.add_argument("--no-oracle", action="store_true", help="skip the independent oracle trajectory (fast variants)")
    a = ap.parse_args()
    from nndetection_b200.arch import conv_ops as ops
    from nndetection_b200.configs import make_plan
    from nndetection_b200.ptmodule import RetinaUNetV001
    from nndetection_b200.training import Trainer, poly_lr
    from nndetection_b200.arch.conv import NormParams
    arch, anc, patch, bs = make_plan(a.plan)
    torch.manual_seed(0)
    net = RetinaUNetV001.from_config_plan(None, arch, anc).cuda()
    sched = dict(initial_lr=0.01, warm_iterations=10, warm_lr=1e-6, poly_gamma=0.9, num_iterations=200)
    trainer = Trainer(net, **sched)
    if a.variant == "noside":
        net.anchor_generator.lookup = lambda images: None
    if a.variant == "nodirect":
        for p in net.parameters():
            p._nnd_direct_grad = False
    if a.variant == "mma":
        from ctypes import c_int
        from nndetection_b200 import _lib as L
        ops.set_tensor_path(False)
        L.lib().nnd_conv_set_wgrad_tc(c_int(0))
    tsgd = None
    if a.variant == "torchsgd":
        norm_ids = {id(p) for m in net.modules() if isinstance(m, NormParams) for p in m.parameters(recurse=False)}
        dec = [p for p in net.parameters() if id(p) not in norm_ids]
        nod = [p for p in net.parameters() if id(p) in norm_ids]
        tsgd = torch.optim.SGD([{"params": dec, "weight_decay": 3e-5}, {"params": nod, "weight_decay": 0.0}], lr=0.01, momentum=0.9,
                               nesterov=True)

    orc = mo.RetinaUNetOracle(dict(arch), dict(anc))
    orc.load_state_dict({k: v.detach().cpu().clone() for k, v in net.state_dict().items()})
    chk = mo.RetinaUNetOracle(dict(arch), dict(anc))
    o_norm = {id(p) for m in orc.modules() if isinstance(m, (torch.nn.InstanceNorm3d, torch.nn.GroupNorm)) for p in m.parameters(recurse=False)}
    o_opt = torch.optim.SGD([{"params": [p for p in orc.parameters() if id(p) not in o_norm], "weight_decay": 3e-5},
                             {"params": [p for p in orc.parameters() if id(p) in o_norm], "weight_decay": 0.0}], lr=0.01, momentum=0.9,
                            nesterov=True)
    checks = {int(c) for c in a.checks.split(",") if c}
    hist = []
    for step in range(a.steps):
        images, targets = util.toy_learning_batch(patch, bs, 1000 + step)
        tg = {"target_boxes": [b.cuda() for b in targets["target_boxes"]], "target_classes": [c.cuda() for c in targets["target_classes"]],
              "target_seg": targets["target_seg"].cuda()}
        pre = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()} if step in checks else None
        if tsgd is None:
            losses, _ = trainer.train_step(images.cuda(), tg, evaluation=False)
        else:                       # same forward / backward, torch's optimizer on the same parameters
            from nndetection_b200.arch.conv import bump_weights_epoch
            net.train(); trainer.fp.zero_grad()
            losses, _ = net.train_step(images.cuda(), tg, evaluation=False, batch_num=step)
            sum(losses.values()).backward()
            for g_ in tsgd.param_groups:
                g_["lr"] = poly_lr(step, **sched)
            tsgd.step(); bump_weights_epoch()
        dl = {k: float(v.detach()) for k, v in losses.items()}
        # independent oracle trajectory
        ol = {}
        if not a.no_oracle:
            orc.train()
            o_opt.zero_grad()
            lo, aux = orc.train_step(images, targets, seed=step)
            sum(lo.values()).backward()
            for g_ in o_opt.param_groups:
                g_["lr"] = poly_lr(step, **sched)
            o_opt.step()
            ol = {k: float(v.detach()) for k, v in lo.items()}
        rec = {"step": step, "dev": dl, "orc": ol}
        if pre is not None:
            pos_idx, neg_idx, counts, labels, matches = net.last_sample
            cnt = counts.cpu().tolist()
            pos, neg = pos_idx[:cnt[2]].cpu(), neg_idx[:cnt[3]].cpu()
            chk.load_state_dict(pre)
            chk.train()
            chk.zero_grad()
            lc, lb, pred_c = oracle_losses_with_indices(chk, images, targets, pos, neg)
            sum(lc.values()).backward()
            rec["lockstep_losses"] = {k: (dl[k], float(lc[k].detach())) for k in lc}
            rec["labels_equal"] = bool(torch.equal(labels.cpu(), lb.float()))
            rec["counts"] = cnt[:5]
            grads = {}
            named = dict(net.named_parameters())
            for k, p2 in chk.named_parameters():
                g_dev = named[k].grad
                grads[k] = (rel_err(g_dev, p2.grad) if p2.grad is not None else None, float(p2.grad.norm()) if p2.grad is not None else None,
                            float(g_dev.norm()))
            rec["grad_rel_err"] = grads
            worst = sorted(((v[0], k) for k, v in grads.items() if v[0] is not None), reverse=True)[:8]
            rec["worst"] = worst
            # post-step weights: torch SGD on the device gradients from the pre-step weights
            lr = poly_lr(step, **sched)
            wd_err = {}
            for k, p in named.items():
                w0 = pre[k].double()
                g = p.grad.detach().cpu().double()
                is_norm = k.endswith("norm.weight") or k.endswith("norm.bias")
                g = g + (0.0 if is_norm else 3e-5) * w0
                # first step: buf = g; later steps need the momentum history -> only check step 0 exactly
                if step == 0:
                    upd = g + 0.9 * g
                    wd_err[k] = rel_err(p.detach().cpu().double(), w0 - lr * upd)
            if wd_err:
                rec["post_step_weight_err_max"] = max(wd_err.values())
        hist.append(rec)
        line = f"step {step:3d} dev " + " ".join(f"{k}={v:+.4f}" for k, v in dl.items()) + " | orc " + " ".join(f"{k}={v:+.4f}" for k, v in ol.items())
        print(line, flush=True)
        if pre is not None:
            print("   lockstep losses (dev, oracle@same weights+indices):", {k: (round(x, 5), round(y, 5)) for k, (x, y) in rec["lockstep_losses"].items()},
                  "labels_equal", rec["labels_equal"], "counts", rec["counts"], flush=True)
            print("   worst grad rel err:", [(round(e, 3), k) for e, k in rec["worst"]], flush=True)
            med = float(np.median([v[0] for v in grads.values() if v[0] is not None]))
            print("   median grad rel err:", round(med, 4), "post-step weight err:", rec.get("post_step_weight_err_max"), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(hist, open(os.path.join(ROOT, "gpurun_out", f"diag_toy_{a.variant}.json"), "w"), indent=1, default=str)
    # validation like the test
    net.eval(); orc.eval()
    ious, scores, labels_, o_ious = [], [], [], []
    for v in range(5):
        images, targets = util.toy_learning_batch(patch, bs, 5000 + v)
        pred = net.inference_step(images.cuda())
        with torch.no_grad():
            po, anchors, _ = orc(images)
            dets = orc.postprocess(images, po, anchors)
        for i in range(bs):
            b, s, l = pred["pred_boxes"][i], pred["pred_scores"][i], pred["pred_labels"][i]
            ious.append(float(bo.box_iou(targets["target_boxes"][i], b[:1].float().cpu())[0, 0]) if b.shape[0] else 0.0)
            scores.append(float(s[0]) if b.shape[0] else 0.0); labels_.append(int(l[0]) if b.shape[0] else -1)
            ob = dets[i][0]
            o_ious.append(float(bo.box_iou(targets["target_boxes"][i], ob[:1])[0, 0]) if ob.shape[0] else 0.0)
    print("validation dev ious", [round(x, 2) for x in ious], "mean", round(float(np.mean(ious)), 3))
    print("validation dev scores", [round(x, 2) for x in scores], "labels", labels_)
    print("validation orc ious", [round(x, 2) for x in o_ious], "mean", round(float(np.mean(o_ious)), 3))


if __name__ == "__main__":
    main()
