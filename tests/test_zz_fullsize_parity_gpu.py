"""Oracle parity AT BASELINE SIZES (BASELINE.json configs 2, 5, 3): one full-size patch of the LUNA16-shaped
(128^3, 1 ch, C = 1), ADAM-shaped (128^3, 2 ch, C = 3, 20 ground-truth boxes = ATSS stress) and LIDC-shaped (96 x 192 x 192, first
stride (1, 2, 2)) plans through the CUDA path and through the CPU oracle (fp32 torch operators = the reference's arithmetic,
oracle/model_oracle.py) with IDENTICAL weights and inputs.

Gates: ATSS matches / labels bit-exact at A = 1 010 880 resp. 3 411 720 anchors (reference core/boxes/matcher/atss.py:48-122,
core/retina.py:228-290); network outputs <= 5e-2 relative in norm (bf16 activations through ~40 layers against fp32); the four
losses <= 1e-4 relative given identical logits and identical sampled indices (north-star tolerance for fp32 losses), <= 3e-2 + 2e-3
end to end (the oracle evaluates ITS logits at the indices the device sampler picked).  The oracle needs ~10-40 s per patch on the
GPU box's host cores."""
import numpy as np
import pytest
import torch

from oracle import box_oracle as bo, model_oracle as mo

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-12))


@pytest.mark.parametrize("name,seed,max_gt", [("luna", 3, 4), ("adam", 11, 20), ("lidc", 5, 4)])
def test_one_full_size_patch_against_the_oracle(name, seed, max_gt):
    from nndetection_b200.configs import make_plan, synth_batch
    from nndetection_b200.ptmodule import RetinaUNetV001
    arch, anc, patch, _ = make_plan(name)
    C = arch["classifier_classes"]
    torch.manual_seed(100 + seed)
    net = RetinaUNetV001.from_config_plan(None, arch, anc)
    orc = mo.RetinaUNetOracle(dict(arch), dict(anc))
    orc.load_state_dict({k: v.detach().clone() for k, v in net.state_dict().items()})
    net = net.cuda().train()
    images, targets = synth_batch(patch, 1, arch["in_channels"], C, seed, max_gt=max_gt)
    G = targets["target_boxes"][0].shape[0]
    assert G >= 1 and (name != "adam" or G == 20)
    tg = {"target_boxes": [b.cuda() for b in targets["target_boxes"]], "target_classes": [c.cuda() for c in targets["target_classes"]],
          "target_seg": targets["target_seg"].cuda()}
    # --- device
    # outputs of the SAME forward pass the losses were computed from (a second pass differs in the last bits: the norm statistics are
    # accumulated with fp32 atomics)
    seen = {}
    h1 = net.head.register_forward_hook(lambda m, i, o: seen.__setitem__("det", {k: v.detach().clone() for k, v in o.items()}))
    h2 = net.segmenter.register_forward_hook(lambda m, i, o: seen.__setitem__("seg", {k: v.detach().clone() for k, v in o.items()}))
    with torch.no_grad():
        losses, _ = net.train_step(images.cuda(), tg, evaluation=False, batch_num=0)
        pos_idx, neg_idx, counts, labels, matches = net.last_sample
    h1.remove(); h2.remove()
    pdet, pseg = seen["det"], seen["seg"]
    anchors_d = net.anchor_generator.lookup(images.cuda())
    assert anchors_d is not None
    anchors_d = [anchors_d[0]]
    cnt = counts.cpu().tolist()
    pos, neg = pos_idx[:cnt[2]].cpu(), neg_idx[:cnt[3]].cpu()
    # --- oracle
    with torch.no_grad():
        po, anchors_o, so = orc(images)
        a = anchors_o[0]
        A = a.shape[0]
        assert A == {"luna": 1010880, "adam": 1010880, "lidc": 3411720}[name] and anchors_d[0].shape[0] == A
        assert torch.equal(anchors_d[0].cpu(), a)
        _, m_o = bo.atss_match(targets["target_boxes"][0], a, orc.per_level, orc.apos, orc.num_candidates)
        lab_o, mb_o = bo.assign_targets(m_o, targets["target_boxes"][0], targets["target_classes"][0], A)
    # ATSS: bit-exact
    assert torch.equal(matches.cpu(), m_o), f"ATSS matches differ at {int((matches.cpu() != m_o).sum())} anchors"
    assert torch.equal(labels.cpu(), lab_o.float())
    n_pos = int((lab_o >= 1).sum())
    assert n_pos > 0 and cnt[0] == n_pos and cnt[1] == int((lab_o == 0).sum())
    assert (cnt[2], cnt[3], cnt[4]) == bo.hnm_counts(cnt[0], cnt[1], 1)
    assert (lab_o[pos] >= 1).all() and (lab_o[neg] == 0).all()
    # forward
    assert rel_err(pdet["box_logits"], po["box_logits"]) < 5e-2
    assert rel_err(pdet["box_deltas"], po["box_deltas"]) < 5e-2
    assert rel_err(pseg["seg_logits"], so["seg_logits"]) < 5e-2
    # losses given identical logits + indices: the oracle's loss functions on the DEVICE's outputs
    with torch.no_grad():
        l_same = bo.head_loss(pdet["box_logits"].cpu(), pdet["box_deltas"].cpu(), lab_o, mb_o, a, pos, neg, C)
        l_same.update(bo.seg_loss(pseg["seg_logits"].float().cpu(), targets["target_seg"]))
        l_e2e = bo.head_loss(po["box_logits"], po["box_deltas"], lab_o, mb_o, a, pos, neg, C)
        l_e2e.update(bo.seg_loss(so["seg_logits"], targets["target_seg"]))
    for k in ("reg", "cls", "seg_ce", "seg_dice"):
        d, s, e = float(losses[k]), float(l_same[k]), float(l_e2e[k])
        assert abs(d - s) <= 1e-4 * abs(s) + 1e-6, (k, d, s)
        assert abs(d - e) <= 3e-2 * abs(e) + 2e-3, (k, d, e)


@pytest.mark.parametrize("shape,frac", [((2, 12, 16, 20), 0.2), ((1, 8, 8, 8), 0.0), ((3, 5, 7, 9), 0.9)])
def test_seg_loss_forward_backward_vs_oracle_autograd(shape, frac):
    """csrc/seg.cu seg_loss_fwd / seg_loss_bwd against autograd of the oracle's restatement of DiCESegmenterFgBg.compute_loss
    (nndet/arch/heads/segmenter.py:184-203,273-290; SoftDiceLoss batch dice, nndet/losses/segmentation.py:84-151): losses and
    d(loss)/d(logits) at 1e-4 relative, for both losses separately and summed, incl. an all-background target."""
    from nndetection_b200.arch.net import _SegLossFn
    g = torch.Generator().manual_seed(int(1000 * frac) + shape[1])
    logits = torch.randn(shape[0], 2, *shape[1:], generator=g) * 2.0
    target = (torch.rand(shape[0], *shape[1:], generator=g) < frac).float() * 3.0          # instance-style values > 1: binarised
    for w_ce, w_dice in ((1.0, 0.0), (0.0, 1.0), (1.0, 1.0), (0.3, 2.0)):
        lo = logits.clone().requires_grad_(True)
        o = bo.seg_loss(lo, target)
        (w_ce * o["seg_ce"] + w_dice * o["seg_dice"]).backward()
        ld = logits.cuda().requires_grad_(True)
        ce, dice = _SegLossFn.apply(ld, target.cuda(), 0.5, 1e-5)
        (w_ce * ce + w_dice * dice).backward()
        assert abs(float(ce) - float(o["seg_ce"])) <= 1e-4 * abs(float(o["seg_ce"])) + 1e-7
        assert abs(float(dice) - float(o["seg_dice"])) <= 1e-4 * abs(float(o["seg_dice"])) + 1e-7
        assert rel_err(ld.grad, lo.grad) < 1e-4, (w_ce, w_dice)
        torch.testing.assert_close(ld.grad.cpu(), lo.grad, rtol=1e-3, atol=1e-7 * float(lo.grad.abs().max()) + 1e-9)


def test_inference_patch_160_against_the_oracle():
    """BASELINE config 4's patch (160^3, A = 1 974 375 anchors): `inference_step` on the device against the oracle -- network outputs
    <= 5e-2 in norm against the fp32 CPU forward, and the whole post-processing (decode + clip + sigmoid + top-10 000 + score / size
    filters + 3-D NMS + top-100: nndet/core/retina.py:292-379) BIT-EXACT against the oracle's restatement applied to the device's own
    logits / deltas (boxes within one ulp of expf)."""
    from nndetection_b200.configs import make_plan
    from nndetection_b200.ptmodule import RetinaUNetV001
    arch, anc, patch, _ = make_plan("infer160")
    torch.manual_seed(160)
    net = RetinaUNetV001.from_config_plan(None, arch, anc)
    orc = mo.RetinaUNetOracle(dict(arch), dict(anc))
    orc.load_state_dict({k: v.detach().clone() for k, v in net.state_dict().items()})
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(4)
    images = torch.rand(1, 1, *patch, generator=g)
    with torch.no_grad():
        pdet, anchors, pseg = net(images.cuda())
        pred = net.postprocess_for_inference(images=images.cuda(), pred_detection=pdet, pred_seg=pseg, anchors=anchors)
        po, anchors_o, so = orc(images)
    A = anchors_o[0].shape[0]
    assert A == 1974375 and torch.equal(anchors[0].cpu(), anchors_o[0])
    assert rel_err(pdet["box_logits"], po["box_logits"]) < 5e-2
    assert rel_err(pdet["box_deltas"], po["box_deltas"]) < 5e-2
    assert rel_err(pseg["seg_logits"], so["seg_logits"]) < 5e-2
    # decode + sigmoid on the device against torch on the same raw outputs (expf ulp), then the oracle's selection / filters / NMS on
    # the DEVICE's decoded boxes and probabilities -- the same ordering problem, so bit-exact
    from nndetection_b200.core.boxes import engine as E
    boxes_d = E.decode_boxes(pdet["box_deltas"], anchors[0], clip_shape=patch)
    probs_d = E.sigmoid_fg(pdet["box_logits"], want_fg=False)[0]
    torch.testing.assert_close(probs_d.cpu(), torch.sigmoid(pdet["box_logits"].cpu()), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(boxes_d.cpu(), bo.clip_boxes_3d(bo.decode_single(pdet["box_deltas"].cpu(), anchors_o[0]), patch), rtol=1e-5, atol=1e-3)
    rb, rs, rl = bo.postprocess_single_image(boxes_d.cpu(), probs_d.cpu(), patch, arch["classifier_classes"], topk=10000, score_thresh=0,
                                             min_size=0.01, nms_thresh=net.nms_thresh, detections_per_img=100)
    b, s, l = pred["pred_boxes"][0].cpu(), pred["pred_scores"][0].cpu(), pred["pred_labels"][0].cpu()
    assert b.shape == rb.shape and b.shape[0] > 0
    assert torch.equal(l, rl) and torch.equal(s, rs) and torch.equal(b, rb)
