"""GPU parity of the box engine (C ABI via nndetection_b200.core.boxes.engine) against the oracle + goldens.
Gates (north star): bit-exact anchors / match indices / sampler index sets / NMS keep; fp32 metrics, decoded
boxes and losses within 1e-4 relative (written per assert)."""
import zlib

import numpy as np
import pytest
import torch

import tutil as util
from oracle import box_oracle as bo, model_oracle as mo

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def E():
    from nndetection_b200.core.boxes import engine
    return engine


@pytest.mark.parametrize("name", ["tiny", "toy", "luna"])
def test_anchor_grid_bit_exact(E, name):
    g = util.golden(f"anchors_{name}")
    arch, anc, patch, _ = mo.make_plan(name)
    fm = g["fmap_sizes"].tolist()
    strides = [[int(i / s) for i, s in zip(patch, f)] for f in fm]
    levels = []
    for f, s, w, h, d in zip(fm, strides, anc["width"], anc["height"], anc["depth"]):
        levels.append(E.anchor_grid(bo.base_anchors_3ds(w, h, d).cuda(), f, s))
    a = torch.cat(levels).cpu()
    assert [l.shape[0] for l in levels] == g["per_level"].tolist()
    assert zlib.crc32(a.numpy().tobytes()) == int(g["crc"])          # bit-exact vs the executed reference
    assert torch.equal(a[T(g["sample_idx"])], T(g["sample"]))


def test_pairwise_metrics(E):
    g = util.golden("pairwise")
    a, b = T(g["a"]).cuda(), T(g["b"]).cuda()
    assert torch.equal(E.pairwise(a, b, 0).cpu(), T(g["iou"]))       # same fp32 op sequence -> bit-exact
    torch.testing.assert_close(E.pairwise(a, b, 2).cpu(), T(g["dist"]), rtol=1e-6, atol=1e-6)   # sum order of 3 squares
    torch.testing.assert_close(E.pairwise(a, b, 1, 1e-7).cpu(), T(g["giou"]), rtol=1e-6, atol=1e-7)


def test_decode_and_clip(E):
    g = util.golden("coder")
    dec = E.decode_boxes(T(g["rel"]).cuda(), T(g["anchors"]).cuda()).cpu()
    torch.testing.assert_close(dec, T(g["decoded"]), rtol=1e-5, atol=1e-4)         # expf ulp differences only
    cl = E.decode_boxes(T(g["rel"]).cuda(), T(g["anchors"]).cuda(), clip_shape=(128, 128, 128)).cpu()
    torch.testing.assert_close(cl, T(g["clipped"]), rtol=1e-5, atol=1e-4)
    assert float(cl.min()) >= 0 and float(cl.max()) <= 128


def _toy_anchors():
    arch, anc, patch, _ = mo.make_plan("toy")
    fm = [[16, 32, 32], [8, 16, 16]]
    anchors, per_level = bo.anchors_for_image(patch, fm, anc["width"], anc["height"], anc["depth"])
    return anchors, per_level, patch


def test_atss_golden_bit_exact(E):
    g = util.golden("atss_toy")
    anchors, per_level, _ = _toy_anchors()
    gts = [T(g[f"gt{i}"]) for i in range(int(g["n_cases"]))] + [torch.zeros(0, 6)]
    gtb = E.GtBatch(gts, [torch.zeros(x.shape[0], dtype=torch.int64) for x in gts], "cuda")
    m = E.atss_match(gtb, anchors.cuda(), per_level, 4 * 27).cpu().view(len(gts), -1)
    for i in range(int(g["n_cases"])):
        assert torch.equal(torch.where(m[i] >= 0)[0], T(g[f"pos_idx{i}"]))
        assert torch.equal(m[i][m[i] >= 0], T(g[f"pos_gt{i}"]))
    assert (m[-1] == -1).all()


def test_atss_vs_oracle_many_gt_and_ties(E):
    anchors, per_level, patch = _toy_anchors()
    # 20 GT boxes incl. on-grid centres (27-fold + positional distance ties): canonical tie-break must agree
    _, tg = mo.synth_batch(patch, 2, 1, 2, 7, max_gt=20)
    gt = tg["target_boxes"][0]
    gt = torch.cat([gt, torch.tensor([[4., 4, 12, 12, 4, 12], [0., 0, 30, 60, 0, 60], [10., 10, 10.5, 10.5, 10, 10.5]])])
    gtb = E.GtBatch([gt], [torch.zeros(gt.shape[0], dtype=torch.int64)], "cuda")
    m = E.atss_match(gtb, anchors.cuda(), per_level, 4 * 27).cpu()
    _, mo_ = bo.atss_match(gt, anchors, per_level, 27, 4, canonical_ties=True)
    assert torch.equal(m, mo_)


def test_labels_sampler_and_loss(E):
    arch, anc, patch, bs = mo.make_plan("toy")
    anchors, per_level, _ = _toy_anchors()
    A = anchors.shape[0]
    _, tg = mo.synth_batch(patch, bs, 1, 2, 31)
    gtb = E.GtBatch(tg["target_boxes"], tg["target_classes"], "cuda")
    m = E.atss_match(gtb, anchors.cuda(), per_level, 108)
    labels = E.assign_labels(m, gtb, A)
    # oracle
    lab_o, mb_o = [], []
    for gb, gc in zip(tg["target_boxes"], tg["target_classes"]):
        _, mm = bo.atss_match(gb, anchors, per_level, 27, 4)
        l, mb = bo.assign_targets(mm, gb, gc, A)
        lab_o.append(l); mb_o.append(mb)
    lab_o, mb_o = torch.cat(lab_o), torch.cat(mb_o)
    assert torch.equal(labels.cpu(), lab_o)
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(bs * A, 2, generator=g) - 3.0
    deltas = torch.randn(bs * A, 6, generator=g) * 0.3
    probs, fg = E.sigmoid_fg(logits.cuda())
    torch.testing.assert_close(probs.cpu(), torch.sigmoid(logits), rtol=1e-6, atol=1e-7)
    plan = E.SamplerPlan(bs)
    counts, pos, neg, pool, _ws = E.hnm_sample(labels, fg, plan, seed=99, want_pool=True)
    c = counts.cpu().tolist()
    pos_o, neg_o, pool_o = bo.hnm_select(lab_o, fg.cpu(), bs, 99)          # oracle on the SAME probabilities
    assert c[0] == int((lab_o >= 1).sum()) and c[1] == int((lab_o == 0).sum())
    assert (c[2], c[3], c[4]) == bo.hnm_counts(c[0], c[1], bs)
    assert torch.equal(torch.sort(pool[:c[4]].cpu().long())[0], pool_o)
    assert torch.equal(pos[:c[2]].cpu(), pos_o) and torch.equal(neg[:c[3]].cpu(), neg_o)
    # losses + gradients
    losses, gd, gl = E.head_loss_fwd(logits.cuda(), deltas.cuda(), anchors.cuda(), m, gtb, labels, pos, neg, counts)
    lg = logits.clone().requires_grad_(True); dl = deltas.clone().requires_grad_(True)
    lo = bo.head_loss(lg, dl, lab_o, mb_o, anchors.repeat(bs, 1), pos_o, neg_o, 2)
    (lo["reg"] + lo["cls"]).backward()
    torch.testing.assert_close(losses.cpu(), torch.stack([lo["reg"], lo["cls"]]).detach(), rtol=1e-4, atol=1e-6)
    dd, dlg = E.head_loss_bwd(gd, gl, pos, neg, counts, bs * A, None, None)
    torch.testing.assert_close(dd.cpu(), dl.grad, rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(dlg.cpu(), lg.grad, rtol=1e-3, atol=1e-7)


def test_sampler_golden_and_edge_cases(E):
    g = util.golden("sampler")
    labels, probs = T(g["labels"]).cuda(), T(g["probs"]).cuda()
    plan = E.SamplerPlan(4)
    counts, pos, neg, pool, _ws = E.hnm_sample(labels, probs, plan, seed=int(g["hash_seed"]), want_pool=True)
    c = counts.cpu().tolist()
    assert c[4] == int(g["pool_size"])
    assert torch.equal(torch.sort(pool[:c[4]].cpu().long())[0], T(g["pool"]))      # reference topk pool set
    assert torch.equal(pos[:c[2]].cpu(), T(g["hash_pos"])) and torch.equal(neg[:c[3]].cpu(), T(g["hash_neg"]))
    # no positives; all-equal probabilities (every negative ties): canonical = lowest indices
    lab = torch.zeros(20000); pr = torch.full((20000,), 0.25)
    counts, pos, neg, pool, _ws = E.hnm_sample(lab.cuda(), pr.cuda(), plan, seed=1, want_pool=True)
    c = counts.cpu().tolist()
    assert (c[2], c[3], c[4]) == bo.hnm_counts(0, 20000, 4)
    assert torch.equal(torch.sort(pool[:c[4]].cpu().long())[0], torch.arange(c[4]))
    # fewer negatives than the pool
    lab = torch.full((500,), -1.0); lab[:7] = 0; lab[100:103] = 2
    pr = torch.rand(500)
    counts, pos, neg, pool, _ws = E.hnm_sample(lab.cuda(), pr.cuda(), plan, seed=1, want_pool=True)
    c = counts.cpu().tolist()
    assert (c[0], c[1], c[2], c[3], c[4]) == (3, 7) + bo.hnm_counts(3, 7, 4)
    pos_o, neg_o, pool_o = bo.hnm_select(lab, pr, 4, 1)
    assert pos[:3].cpu().tolist() == [100, 101, 102] and torch.equal(neg[:c[3]].cpu(), neg_o)
    assert torch.sort(pool[:c[4]].cpu().long())[0].tolist() == list(range(7))


def test_postprocess_vs_oracle(E):
    anchors, per_level, patch = _toy_anchors()
    A, C, B = anchors.shape[0], 2, 2
    g = torch.Generator().manual_seed(17)
    deltas = torch.randn(B * A, 6, generator=g) * 0.2
    logits = torch.randn(B * A, C, generator=g) * 2 - 4
    boxes = E.decode_boxes(deltas.cuda(), anchors.cuda(), clip_shape=patch)
    probs, _ = E.sigmoid_fg(logits.cuda(), want_fg=False)
    ob, os_, ol, oc = E.detect_postprocess(boxes, probs, B, A, C, topk=10000, score_thresh=0.0, min_size=0.01,
                                           nms_thresh=0.6, det_per_img=100)
    oc = oc.cpu().tolist()
    bc, pc = boxes.cpu(), probs.cpu()
    for i in range(B):
        rb, rs, rl = bo.postprocess_single_image(bc[i * A:(i + 1) * A], pc[i * A:(i + 1) * A], patch, C)
        assert oc[i] == rb.shape[0]
        assert torch.equal(ol[i, :oc[i]].cpu(), rl)                  # bit-exact keep set / order / labels
        assert torch.equal(os_[i, :oc[i]].cpu(), rs)
        assert torch.equal(ob[i, :oc[i]].cpu(), rb)
