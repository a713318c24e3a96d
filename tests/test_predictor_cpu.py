"""Host logic of the sliding-window predictor (nndetection_b200/inference/predictor.py): tile grid, shifted crops, mirror TTA with
box un-mirroring, the per-model / per-TTA / per-batch loop and the tile sharding over ranks -- against fixtures produced by
EXECUTING the reference's `create_grid` / `save_get_crop` / `Mirror` / `BoxEnsemblerSelective` (scripts/gen_golden.py predictor).
A deterministic fake detector replaces the network, the oracle's NMS / WBC the device kernels (both have their own GPU tests)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import box_oracle as bo
import tutil as util


def _o_weighted_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
    keep = bo.batched_nms(boxes, scores * weights, labels, iou_thresh, cuda_semantics=False)
    return boxes[keep], scores[keep], labels[keep], torch.ones_like(weights)[keep]


def _o_wbc_ensemble(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh, *a, **k):
    return bo.batched_wbc(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh)


def _predictor(shard=(0, 1)):
    from nndetection_b200.inference.ensembler import BoxEnsemblerSelective
    from nndetection_b200.inference.predictor import SlidingWindowPredictor
    return SlidingWindowPredictor(
        ensembler_fn=lambda c, properties=None: BoxEnsemblerSelective.from_case(
            c, properties, parameters={"model_nms_fn": _o_weighted_nms_model, "ensemble_nms_fn": _o_wbc_ensemble}),
        models=[util.FakeDetector()], crop_size=(32, 32, 32), overlap=0.5, num_tta_transforms=8, batch_size=4, device="cpu", shard=shard)


def _case():
    g = torch.Generator().manual_seed(17)
    return {"data": torch.rand(1, 40, 56, 48, generator=g)}


def test_tile_grid_matches_reference():
    from nndetection_b200.inference.predictor import create_grid, get_tta_dims
    g = util.golden("predictor")
    for gi, (ps, dl, ov) in enumerate(util.GRID_CASES):
        for mode in ("fixed", "symmetric"):
            for cb in (False, True):
                m = create_grid((ps, ps), (dl, dl + 7), (ov, ov), mode=mode, center_boarder=cb)
                assert np.array_equal(np.asarray([[(s.start, s.stop) for s in c] for c in m], dtype=np.int64), g[f"grid{gi}_{mode}_{int(cb)}"])
    assert get_tta_dims(0) == [()] and len(get_tta_dims(4)) == 4 and get_tta_dims(8)[-1] == (0, 1, 2)


def test_case_prediction_matches_reference_pipeline():
    g = util.golden("predictor")
    pred = _predictor()
    case = _case()
    tiles = pred.tile_case(case)
    assert np.array_equal(np.asarray([t["tile_origin"] for t in tiles]), g["tile_origins"])
    assert all(t["data"].shape == (1, 32, 32, 32) and t["data"].data_ptr() >= case["data"].data_ptr() for t in tiles)     # views, no copies
    out = pred.predict_case(case)["boxes"]
    assert np.array_equal(out["pred_boxes"].numpy(), g["case_boxes"])
    assert np.array_equal(out["pred_scores"].numpy(), g["case_scores"])
    assert np.array_equal(out["pred_labels"].numpy(), g["case_labels"])


def _shard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    out = _predictor(shard=(rank, world)).predict_case(_case())
    if rank == 0:
        q.put({k: out["boxes"][k].numpy() for k in ("pred_boxes", "pred_scores", "pred_labels")})
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_tile_sharding_over_two_ranks_gives_the_single_process_result():
    """SURVEY 8e (inference): tiles rank::world per rank, one gather of the per-tile detections, whole-case NMS / WBC on rank 0."""
    g = util.golden("predictor")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res["pred_boxes"], g["case_boxes"]) and np.array_equal(res["pred_scores"], g["case_scores"])
    assert np.array_equal(res["pred_labels"], g["case_labels"])


def test_patch_larger_than_the_case_falls_back_to_symmetric_padding():
    """predictor.py:223-228 / patching.py:396-452: `np.pad(mode="symmetric")` of the clipped crop, origin = crop start (negative);
    also with pads wider than the data (several reflections)."""
    from nndetection_b200.inference.predictor import SlidingWindowPredictor, create_grid, padded_crop_symmetric
    g = torch.Generator().manual_seed(4)
    data = torch.rand(2, 5, 40, 7, generator=g)
    for crop in [(slice(-3, 9), slice(4, 36), slice(-2, 10)), (slice(-14, 18), slice(-1, 31), slice(-13, 19)), (slice(1, 4), slice(0, 40), slice(5, 9))]:
        tile, origin, c = padded_crop_symmetric(data, crop)
        clipped = tuple(slice(max(s.start, 0), min(s.stop, d)) for s, d in zip(crop, data.shape[1:]))
        pads = [(0, 0)] + [(max(-s.start, 0), max(s.stop - d, 0)) for s, d in zip(crop, data.shape[1:])]
        ref = np.pad(data.numpy()[(slice(None), *clipped)], pads, mode="symmetric")
        assert np.array_equal(tile.numpy(), ref) and origin == [s.start for s in crop] and list(c) == list(crop)
    pred = SlidingWindowPredictor(lambda c, properties=None: None, [], (16, 32, 16), 0.5, 8, 4, device="cpu")
    tiles = pred.tile_case({"data": data})
    crops = create_grid((16, 32, 16), (5, 40, 7), [8, 16, 8], mode="symmetric")
    assert len(tiles) == len(crops) and all(t["data"].shape == (2, 16, 32, 16) for t in tiles)
    assert [t["tile_origin"] for t in tiles] == [[s.start for s in c] for c in crops]
