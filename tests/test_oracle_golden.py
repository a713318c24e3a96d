"""CPU: the oracle restatement against the golden vectors produced by the executed reference."""
import zlib

import numpy as np
import pytest
import torch

from oracle import box_oracle as bo, model_oracle as mo
import tutil as util


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_pairwise_metrics_bit_exact():
    g = util.golden("pairwise")
    a, b = T(g["a"]), T(g["b"])
    assert torch.equal(bo.box_iou(a, b), T(g["iou"]))
    assert torch.equal(bo.generalized_box_iou(a, b, eps=1e-7), T(g["giou"]))
    assert torch.equal(bo.box_center_dist(a, b), T(g["dist"]))


@pytest.mark.parametrize("name", ["tiny", "toy", "luna"])
def test_anchor_grid_matches_reference(name):
    g = util.golden(f"anchors_{name}")
    arch, anc, patch, _ = mo.make_plan(name)
    a, per_level = bo.anchors_for_image(tuple(g["patch"]), g["fmap_sizes"].tolist(), anc["width"], anc["height"], anc["depth"])
    assert per_level == g["per_level"].tolist()
    assert zlib.crc32(a.numpy().tobytes()) == int(g["crc"])
    assert torch.equal(a[T(g["sample_idx"])], T(g["sample"]))


def test_atss_matches_bit_exact():
    g = util.golden("atss_toy")
    arch, anc, patch, _ = mo.make_plan("toy")
    anchors, per_level = bo.anchors_for_image(patch, g["fmap_sizes"].tolist(), anc["width"], anc["height"], anc["depth"])
    for i in range(int(g["n_cases"])):
        _, m = bo.atss_match(T(g[f"gt{i}"]), anchors, per_level, 27, 4)
        assert torch.equal(torch.where(m >= 0)[0], T(g[f"pos_idx{i}"]))
        assert torch.equal(m[m >= 0], T(g[f"pos_gt{i}"]))
    _, m = bo.atss_match(torch.zeros(0, 6), anchors, per_level, 27, 4)
    assert (m == -1).all() and m.dtype == torch.int64


def test_sampler_counts_and_pool():
    g = util.golden("sampler")
    for npos, nneg, bs, p, n, pool in g["counts"].tolist():
        assert bo.hnm_counts(npos, nneg, bs) == (p, n, pool)
    labels, probs = T(g["labels"]), T(g["probs"])
    assert torch.equal(bo.hnm_pool(labels, probs, int(g["pool_size"])), T(g["pool"]))
    pos, neg, _ = bo.hnm_select(labels, probs, 4, seed=int(g["hash_seed"]))
    assert torch.equal(pos, T(g["hash_pos"])) and torch.equal(neg, T(g["hash_neg"]))


def test_coder_clip_small():
    g = util.golden("coder")
    dec = bo.decode_single(T(g["rel"]), T(g["anchors"]))
    assert torch.equal(dec, T(g["decoded"]))
    cl = bo.clip_boxes_3d(dec, (128, 128, 128))
    assert torch.equal(cl, T(g["clipped"]))
    assert torch.equal(bo.keep_not_small(cl, 0.01), T(g["keep"]))


def test_nms_keep_lists_bit_exact():
    g = util.golden("nms")
    for n, thr in g["cases"].tolist():
        n = int(n)
        boxes, scores = util.nms_case(n)
        keep = bo.nms_greedy(boxes, scores, thr)
        assert torch.equal(keep, T(g[f"n{n}_t{thr}_keep"])), (n, thr)
    gen = torch.Generator().manual_seed(4242)
    boxes = util.rand_boxes(1500, gen); scores = util.unique_scores(1500, gen)
    idxs = torch.randint(0, 3, (1500,), generator=gen)
    assert torch.equal(bo.batched_nms(boxes, scores, idxs, 0.5), T(g["batched_keep"]))


def test_nms_nan_semantics_differ_as_documented():
    # two zero-volume boxes: CUDA semantics keep both (NaN > thr false), nms_cpu drops the second
    b = torch.tensor([[1., 1, 1, 1, 1, 1], [1., 1, 1, 1, 1, 1]])
    s = torch.tensor([0.9, 0.8])
    assert bo.nms_greedy(b, s, 0.5, cuda_semantics=True).tolist() == [0, 1]
    assert bo.nms_greedy(b, s, 0.5, cuda_semantics=False).tolist() == [0]


def test_model_oracle_matches_reference_golden():
    g = util.golden("model_tiny")
    arch, anc, patch, bs = mo.make_plan("tiny")
    torch.manual_seed(0)
    net = mo.RetinaUNetOracle(dict(arch), dict(anc))
    net.load_state_dict(util.det_fill(net.state_dict(), int(g["seed"])))
    images, targets = mo.synth_batch(patch, bs, arch["in_channels"], arch["classifier_classes"], 2024 + int(g["seed"]))
    assert zlib.crc32(images.numpy().tobytes()) == int(g["images_crc"])
    losses, aux = net.train_step(images, targets, seed=int(g["sampler_seed"]))
    sum(losses.values()).backward()
    for k, v in losses.items():
        assert torch.allclose(v.detach(), T(g["loss_" + k]), rtol=1e-5, atol=1e-6), k
    assert torch.equal(aux["pos"], T(g["pos_idx"])) and torch.equal(aux["neg"], T(g["neg_idx"]))
    assert torch.allclose(aux["pred"]["box_logits"].detach(), T(g["box_logits"]), rtol=1e-4, atol=1e-5)
    for k, p in net.named_parameters():
        ref = float(g["gnorm/" + k])
        assert abs(float(p.grad.double().norm()) - ref) <= 1e-3 * max(ref, 1e-6) + 1e-7, k
    post = net.postprocess(images, {k: v.detach() for k, v in aux["pred"].items()}, aux["anchors"])
    for i in range(bs):
        assert torch.equal(post[i][2], T(g[f"det_labels{i}"]))
        assert torch.allclose(post[i][0], T(g[f"det_boxes{i}"]), atol=1e-4)


def test_wbc_oracle_matches_reference_fixtures():
    """oracle.box_oracle.wbc / batched_wbc vs the executed reference (nndet/inference/detection/wbc.py), bit-exact."""
    g = util.golden("wbc")
    for i, (n, seed, thr, st, ua, mw) in enumerate(g["cases"].tolist()):
        n, seed = int(n), int(seed)
        b, s, w, ne = util.wbc_case(n, seed, extent=60.0 if n <= 1000 else 100.0)
        ob, os_ = bo.wbc(b, s, w, ne, thr, st, use_area=bool(ua), missing_weight=mw)
        assert torch.equal(ob, torch.from_numpy(g[f"c{i}_boxes"])) and torch.equal(os_, torch.from_numpy(g[f"c{i}_scores"]))
    b, s, w, ne = util.wbc_case(600, 9)
    lab = torch.from_numpy(g["batched_labels_in"])
    o = bo.batched_wbc(b, s, lab, w, 0.2, ne, 0.02, use_area=True, missing_weight=1.0)
    assert torch.equal(o[0], torch.from_numpy(g["batched_boxes"])) and torch.equal(o[1], torch.from_numpy(g["batched_scores"]))
    assert torch.equal(o[2], torch.from_numpy(g["batched_labels"]))


def test_instance_transform_oracle_matches_reference_fixtures():
    """oracle.transform_oracle.pre_trafo vs the executed FindInstances -> Instances2Boxes -> Instances2Segmentation chain."""
    import zlib
    from oracle import transform_oracle as to
    g = util.golden("transforms")
    for ci, (B, shape, seed) in enumerate(util.TRANSFORM_CASES):
        t, maps = util.synth_instances(B, shape, seed)
        p, bx, cl, sem = to.pre_trafo(t, maps)
        for b in range(B):
            assert np.array_equal(p[b], g[f"c{ci}_ids{b}"]) and np.array_equal(cl[b], g[f"c{ci}_classes{b}"])
            assert bx[b].shape == g[f"c{ci}_boxes{b}"].shape and np.array_equal(bx[b], g[f"c{ci}_boxes{b}"])
        assert zlib.crc32(sem.tobytes()) == int(g[f"c{ci}_sem_crc"][0])
