"""Seeded synthetic inputs shared by CPU and GPU tests (identical to scripts/gen_golden.py generators)."""
import os
import zlib

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rand_boxes(n, g, extent=160.0, lo=2.0, hi=22.0):
    c = torch.rand(n, 3, generator=g) * extent
    h = torch.rand(n, 3, generator=g) * (hi - lo) + lo
    return torch.stack([c[:, 0] - h[:, 0], c[:, 1] - h[:, 1], c[:, 0] + h[:, 0], c[:, 1] + h[:, 1],
                        c[:, 2] - h[:, 2], c[:, 2] + h[:, 2]], dim=1)


def unique_scores(n, g):
    return (torch.randperm(n, generator=g).float() + 0.5) / max(n, 1)


def nms_case(n, extent=None):
    g = torch.Generator().manual_seed(1000 + n)
    boxes = rand_boxes(n, g, extent=(60.0 if n <= 129 else 160.0) if extent is None else extent)
    return boxes, unique_scores(n, g)


def det_fill(sd, seed=0):
    """Deterministic, platform-independent weights keyed by parameter name (same as scripts/gen_golden.py)."""
    out = {}
    for k, v in sd.items():
        rs = np.random.RandomState((zlib.crc32(k.encode()) + seed) & 0x7FFFFFFF)
        if v.ndim == 0:
            out[k] = torch.tensor(1.0 + 0.1 * rs.standard_normal(), dtype=v.dtype)
        elif k.endswith("norm.weight"):
            out[k] = torch.from_numpy(1.0 + 0.1 * rs.standard_normal(v.shape)).to(v.dtype)
        elif k.endswith("bias"):
            out[k] = torch.from_numpy(0.05 * rs.standard_normal(v.shape)).to(v.dtype)
        else:
            fan_in = int(np.prod(v.shape[1:])) if v.ndim > 1 else 1
            out[k] = torch.from_numpy(rs.standard_normal(v.shape) * (1.5 / np.sqrt(fan_in))).to(v.dtype)
    return out


def wbc_case(n, seed, extent=60.0):
    """Same generator as scripts/gen_golden.py:wbc_case."""
    g = torch.Generator().manual_seed(seed)
    boxes = rand_boxes(n, g, extent=extent, lo=2.0, hi=10.0)
    scores = unique_scores(n, g)
    weights = torch.rand(n, generator=g) * 0.9 + 0.1
    n_exp = torch.randint(1, 9, (n,), generator=g).float()
    return boxes, scores, weights, n_exp


TRANSFORM_CASES = [(2, (12, 16, 20), 1), (4, (32, 32, 32), 2), (1, (8, 8, 8), 3), (3, (24, 40, 36), 4)]


def synth_instances(B, shape, seed, nmax=6):
    """Same generator as scripts/gen_golden.py:synth_instances."""
    rs = np.random.RandomState(seed)
    t = np.zeros((B, 1) + tuple(shape), dtype=np.float32)
    maps = []
    for b in range(B):
        k = rs.randint(0, nmax + 1)
        ids = sorted(rs.choice(np.arange(1, 40), size=k, replace=False).tolist())
        mp = {}
        for i in ids:
            lo = [rs.randint(0, s - 3) for s in shape]
            sz = [rs.randint(1, max(2, s // 3)) for s in shape]
            sl = tuple(slice(l, min(l + z, s)) for l, z, s in zip(lo, sz, shape))
            t[b, 0][sl] = i
            mp[str(i)] = int(rs.randint(0, 3))
        maps.append(mp)
    return t, maps


def synth_tile_predictions(seed, case_shape=(64, 96, 80), tile=(32, 48, 40), n_models=2):
    """Same generator as scripts/gen_golden.py:synth_tile_predictions."""
    g = torch.Generator().manual_seed(seed)
    origins = [(a, b, c) for a in range(0, case_shape[0] - tile[0] + 1, 16) for b in range(0, case_shape[1] - tile[1] + 1, 24)
               for c in range(0, case_shape[2] - tile[2] + 1, 20)]
    total = 0
    models = []
    for m in range(n_models):
        batches = []
        for i in range(0, len(origins), 2):
            bo_, res = origins[i:i + 2], {"pred_boxes": [], "pred_scores": [], "pred_labels": []}
            for _ in bo_:
                n = int(torch.randint(0, 25, (1,), generator=g))
                lo = torch.rand(n, 3, generator=g) * torch.tensor(tile) * 0.8 - 2.0
                sz = torch.rand(n, 3, generator=g) * 10 + 1.0
                res["pred_boxes"].append(torch.stack([lo[:, 0], lo[:, 1], lo[:, 0] + sz[:, 0], lo[:, 1] + sz[:, 1], lo[:, 2],
                                                      lo[:, 2] + sz[:, 2]], dim=1))
                res["pred_scores"].append(torch.empty(n))
                res["pred_labels"].append(torch.randint(0, 2, (n,), generator=g))
                total += n
            batch = {"tile_origin": [torch.tensor([o[ax] for o in bo_]) for ax in range(3)], "data": torch.zeros(len(bo_), 1, *tile)}
            batches.append((res, batch))
        models.append(batches)
    sc = (torch.randperm(total, generator=g).float() + 0.5) / total
    k = 0
    for batches in models:
        for res, _ in batches:
            for j, t in enumerate(res["pred_scores"]):
                res["pred_scores"][j] = sc[k:k + t.numel()]; k += t.numel()
    return models, case_shape


GRID_CASES = [(32, 40, 16), (32, 56, 16), (32, 32, 16), (48, 40, 24), (16, 100, 4), (20, 61, 10), (8, 8, 0)]


class FakeDetector:
    """Same deterministic stand-in model as scripts/gen_golden.py:FakeDetector (detections = function of the tile content)."""

    def eval(self):
        return self

    def inference_step(self, images):
        out = {"pred_boxes": [], "pred_scores": [], "pred_labels": [], "pred_seg": torch.zeros(images.shape[0], 2, *images.shape[2:])}
        D, H, W = images.shape[2:]
        for img in images[:, 0]:
            flat = img.reshape(-1)
            idx = torch.argsort(flat, descending=True, stable=True)[:6]
            z = torch.div(idx, H * W, rounding_mode="floor"); y = torch.div(idx % (H * W), W, rounding_mode="floor"); x = idx % W
            c = torch.stack([z, y, x], 1).float()
            half = 2.0 + 6.0 * flat[idx][:, None] * torch.tensor([1.0, 0.7, 0.5])
            out["pred_boxes"].append(torch.stack([c[:, 0] - half[:, 0], c[:, 1] - half[:, 1], c[:, 0] + half[:, 0], c[:, 1] + half[:, 1],
                                                  c[:, 2] - half[:, 2], c[:, 2] + half[:, 2]], 1))
            ramp = torch.arange(flat.numel(), dtype=torch.float32)
            salt = float((flat * ((ramp * 0.6180339887) % 1.0)).sum() % 1.0)
            out["pred_scores"].append((flat[idx] * 0.9 + 0.1 * salt).clone())
            out["pred_labels"].append(((z + y + x) % 2).long())
        return out


# ------------------------------------------------------------------ sweep fixtures (tests/test_sweeper_cpu.py, scripts/gen_golden.py sweeper)
def o_weighted_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
    from oracle import box_oracle as bo
    keep = bo.batched_nms(boxes, scores * weights, labels, iou_thresh, cuda_semantics=False)
    return boxes[keep], scores[keep], labels[keep], torch.ones_like(weights)[keep]


def o_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
    from oracle import box_oracle as bo
    keep = bo.batched_nms(boxes, scores, labels, iou_thresh, cuda_semantics=False)
    return boxes[keep], scores[keep], labels[keep], weights[keep]


def o_wbc_ensemble(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh, *a, **k):
    from oracle import box_oracle as bo
    return bo.batched_wbc(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh)


def oracle_ensembler_cls():
    """`BoxEnsemblerSelective` whose default / sweep parameters name the oracle's CPU NMS / WBC instead of the device kernels."""
    from nndetection_b200.inference.ensembler import BoxEnsemblerSelective, batched_nms_model, batched_weighted_nms_model

    class OracleEnsembler(BoxEnsemblerSelective):
        @classmethod
        def get_default_parameters(cls):
            d = super().get_default_parameters()
            d.update(model_nms_fn=o_weighted_nms_model, ensemble_nms_fn=o_wbc_ensemble)
            return d

        @classmethod
        def sweep_parameters(cls):
            _, sweep = BoxEnsemblerSelective.sweep_parameters()
            swap = {batched_weighted_nms_model: o_weighted_nms_model, batched_nms_model: o_nms_model}
            sweep["model_nms_fn"] = [swap[f] for f in sweep["model_nms_fn"]]
            return cls.get_default_parameters(), sweep
    return OracleEnsembler


class StandInEvaluator:
    """Evaluator protocol of nndet/evaluator/det.py:BoxEvaluator (create / run_online_evaluation / finish_online_evaluation) with a
    small deterministic metric: per ground-truth box the best score * IoU among same-class predictions, averaged over all boxes, minus
    a 0.0003 penalty per prediction per case -- sensitive to every swept parameter, no ties on the fixtures."""

    @classmethod
    def create(cls, classes, fast=True, verbose=False, save_dir=None):
        return cls()

    def __init__(self):
        self.hits, self.n_gt, self.n_pred, self.n_cases = 0.0, 0, 0, 0

    def run_online_evaluation(self, pred_boxes, pred_classes, pred_scores, gt_boxes, gt_classes, gt_ignore=None):
        from oracle import box_oracle as bo
        for pb, pc, ps, gb, gc in zip(pred_boxes, pred_classes, pred_scores, gt_boxes, gt_classes):
            self.n_cases += 1
            self.n_pred += len(pb)
            self.n_gt += len(gb)
            if len(pb) == 0 or len(gb) == 0:
                continue
            iou = bo.box_iou(torch.as_tensor(np.asarray(gb, dtype=np.float32)), torch.as_tensor(np.asarray(pb, dtype=np.float32))).numpy()
            same = np.asarray(gc).reshape(-1, 1) == np.asarray(pc).reshape(1, -1)
            self.hits += float((iou * same * np.asarray(ps, dtype=np.float64).reshape(1, -1)).max(axis=1).sum())

    def finish_online_evaluation(self):
        return {"stand_in": self.hits / max(self.n_gt, 1) - 0.0003 * self.n_pred / max(self.n_cases, 1)}, None


def write_sweep_cases(pred_dir, gt_dir, n_cases=3):
    """Ensembler states (`<case>_boxes.pt`) of synthetic tile predictions + ground truth derived from each case's own strongest,
    mutually distant predictions (so that good post-processing parameters exist)."""
    import os
    ens_cls = oracle_ensembler_cls()
    os.makedirs(pred_dir, exist_ok=True); os.makedirs(gt_dir, exist_ok=True)
    for ci in range(n_cases):
        models, shape = synth_tile_predictions(100 + ci)
        # second "model" (TTA pass) = jittered copy of the first one's detections + its own: overlapping duplicates inside a pass
        # (tiles overlap) and across passes, so the NMS / WBC thresholds matter
        gj = torch.Generator().manual_seed(900 + ci)
        for (res1, _), (res2, _) in zip(models[0], models[1]):
            for j in range(len(res1["pred_boxes"])):
                b1, s1, l1 = res1["pred_boxes"][j], res1["pred_scores"][j], res1["pred_labels"][j]
                jit = b1 + (torch.rand(b1.shape, generator=gj) - 0.5) * 1.5
                res2["pred_boxes"][j] = torch.cat([res2["pred_boxes"][j], jit])
                res2["pred_scores"][j] = torch.cat([res2["pred_scores"][j], (s1 * (0.6 + 0.4 * torch.rand(s1.shape, generator=gj))).clamp(0.001, 0.999)])
                res2["pred_labels"][j] = torch.cat([res2["pred_labels"][j], l1])
        ens = ens_cls.from_case({"data": torch.zeros(1, *shape)}, properties={}, parameters=None)
        for mi, batches in enumerate(models):
            ens.add_model(name=f"model0_t{mi}", model_weight=1.0 if mi == 0 else 0.7)
            for res, batch in batches:
                ens.process_batch(result=res, batch=batch)
        out = ens.get_case_result()
        b, l = out["pred_boxes"].numpy(), out["pred_labels"].numpy()
        pick = list(range(0, min(len(b), 24), 2))                  # every second of the 24 best consolidated detections
        rs = np.random.RandomState(ci)
        gtb = b[pick] + rs.uniform(-1.0, 1.0, size=(len(pick), 6)).astype(np.float32)
        np.savez(os.path.join(gt_dir, f"case_{ci}_boxes_gt.npz"), boxes=gtb, classes=l[pick].astype(np.int64))
        ens.save_state(pred_dir, f"case_{ci}")


def toy_learning_batch(patch, bs, seed):
    """The reference's toy data (scripts/generate_example.py:49-98) at patch size: uniform noise, one cuboid of side U[6, 12) per image
    with +0.4 intensity = the object (class 0); seg = the cuboid."""
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(bs, 1, *patch, generator=g)
    tb, tc = [], []
    seg = torch.zeros(bs, *patch)
    for i in range(bs):
        lo = [int(torch.randint(1, patch[ax] - 13, (1,), generator=g)) for ax in range(3)]
        sd = [int(torch.randint(6, 12, (1,), generator=g)) for _ in range(3)]
        sl = tuple(slice(l, l + s) for l, s in zip(lo, sd))
        images[i, 0][sl] += 0.4
        seg[i][sl] = 1
        tb.append(torch.tensor([[lo[0] + 0.318, lo[1] + 0.328, lo[0] + sd[0] - 0.359, lo[1] + sd[1] - 0.349, lo[2] + 0.338, lo[2] + sd[2] - 0.339]]))
        tc.append(torch.zeros(1, dtype=torch.int64))
    return images, dict(target_boxes=tb, target_classes=tc, target_seg=seg)


def oracle_losses_with_indices(orc, images, targets, pos, neg):
    """The oracle network's four losses on a batch with the sampled anchor indices INJECTED (instead of the oracle's own sampling),
    plus its ATSS labels and raw outputs: lock-step checks of a device train step (same weights, same batch, same indices)."""
    from oracle import box_oracle as bo
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
