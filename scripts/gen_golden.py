"""Generate tests/golden/* by EXECUTING THE UNMODIFIED REFERENCE (build container only).

Run:  python scripts/gen_golden.py
Needs /root/reference (imported via oracle/ref_import.py).  For every case it
  1. runs the reference function on seeded inputs,
  2. runs the oracle restatement (oracle/box_oracle.py, oracle/model_oracle.py) on the same inputs
     and ASSERTS agreement (bit-exact for indices, tight fp32 tolerance for floats) -- this is what
     pins the oracle, since the reference ships no golden vectors of its own,
  3. stores inputs + reference outputs as compressed .npz fixtures under tests/golden/.
The fixtures are what travels to the GPU box; /root/reference does not.
"""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, box_oracle as bo, model_oracle as mo  # noqa: E402

ref_import.load()
from nndet.core.boxes import ops as rops                      # noqa: E402
import importlib                                                # noqa: E402
rnms = importlib.import_module('nndet.core.boxes.nms')          # noqa: E402
from nndet.core.boxes.anchors import AnchorGenerator3DS       # noqa: E402
from nndet.core.boxes.matcher import ATSSMatcher              # noqa: E402
from nndet.core.boxes.sampler import HardNegativeSamplerBatched  # noqa: E402
from nndet.core.boxes.coder import BoxCoderND                 # noqa: E402
from nndet.core.boxes.clip import clip_boxes_to_image_        # noqa: E402
from nndet.core.retina import BaseRetinaNet                   # noqa: E402
from nndet.arch.conv import ConvInstanceRelu, ConvGroupRelu, Generator  # noqa: E402
from nndet.arch.blocks.basic import StackedConvBlock2         # noqa: E402
from nndet.arch.encoder.modular import Encoder                # noqa: E402
from nndet.arch.decoder.base import UFPNModular               # noqa: E402
from nndet.arch.heads.classifier import BCECLassifier         # noqa: E402
from nndet.arch.heads.regressor import GIoURegressor          # noqa: E402
from nndet.arch.heads.comb import DetectionHeadHNMNative      # noqa: E402
from nndet.arch.heads.segmenter import DiCESegmenterFgBg      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def save(name, **arrs):
    arrs = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(f"  wrote {name}.npz ({os.path.getsize(os.path.join(OUT, name + '.npz')) / 1024:.1f} KiB)")


def rand_boxes(n, g, extent=160.0, lo=2.0, hi=22.0):
    """NMS stress boxes: centres U[0,extent)^3, half sizes U[lo,hi)."""
    c = torch.rand(n, 3, generator=g) * extent
    h = torch.rand(n, 3, generator=g) * (hi - lo) + lo
    return torch.stack([c[:, 0] - h[:, 0], c[:, 1] - h[:, 1], c[:, 0] + h[:, 0], c[:, 1] + h[:, 1],
                        c[:, 2] - h[:, 2], c[:, 2] + h[:, 2]], dim=1)


def unique_scores(n, g):
    return (torch.randperm(n, generator=g).float() + 0.5) / max(n, 1)


# ------------------------------------------------------------------------------------------ weighted box clustering
def wbc_case(n, seed, extent=60.0):
    g = torch.Generator().manual_seed(seed)
    boxes = rand_boxes(n, g, extent=extent, lo=2.0, hi=10.0)
    scores = unique_scores(n, g)
    weights = torch.rand(n, generator=g) * 0.9 + 0.1
    n_exp = torch.randint(1, 9, (n,), generator=g).float()
    return boxes, scores, weights, n_exp


def gen_wbc():
    """nndet/inference/detection/wbc.py loaded from its file (the package __init__ pulls SimpleITK-dependent io modules)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_wbc", os.path.join(ref_import.REF_ROOT, "nndet/inference/detection/wbc.py"))
    rw = importlib.util.module_from_spec(spec); spec.loader.exec_module(rw)
    out, cases = {}, []
    for n, seed, thr, st, ua, mw in [(1, 1, 0.1, 0.0, True, 1.0), (50, 2, 0.1, 0.0, True, 0.7), (300, 3, 0.3, 0.05, False, 1.0),
                                     (1000, 4, 1e-5, 0.1, True, 0.5), (1000, 5, 0.5, 0.0, True, 1.0), (2500, 6, 0.2, 0.02, True, 1.0)]:
        b, s, w, ne = wbc_case(n, seed, extent=60.0 if n <= 1000 else 100.0)
        rb, rs = rw.wbc(b, s, w, ne, thr, st, use_area=ua, missing_weight=mw)
        ob, os_ = bo.wbc(b, s, w, ne, thr, st, use_area=ua, missing_weight=mw)
        rs = rs.flatten()
        assert rb.shape == ob.shape and torch.equal(rb, ob) and torch.equal(rs, os_), (n, thr)
        cases.append((n, seed, thr, st, int(ua), mw))
        out[f"c{len(cases) - 1}_boxes"] = rb; out[f"c{len(cases) - 1}_scores"] = rs
    b, s, w, ne = wbc_case(600, 9)
    lab = torch.randint(0, 3, (600,), generator=torch.Generator().manual_seed(5))
    r = rw.batched_wbc(b, s, lab, w, 0.2, ne, 0.02, use_area=True, missing_weight=1.0)
    o = bo.batched_wbc(b, s, lab, w, 0.2, ne, 0.02, use_area=True, missing_weight=1.0)
    assert torch.equal(r[0], o[0]) and torch.equal(r[1].flatten(), o[1]) and torch.equal(r[2].flatten(), o[2])
    save("wbc", cases=np.asarray(cases, dtype=np.float64), batched_labels_in=lab, batched_boxes=r[0], batched_scores=r[1].flatten(),
         batched_labels=r[2].flatten(), **out)


# ------------------------------------------------------------------------------------------ instance transforms (pre_trafo)
def synth_instances(B, shape, seed, nmax=6):
    """Random cuboid instances (later ones overwrite earlier ones; ids fully overwritten stay in the mapping)."""
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


TRANSFORM_CASES = [(2, (12, 16, 20), 1), (4, (32, 32, 32), 2), (1, (8, 8, 8), 3), (3, (24, 40, 36), 4)]


def gen_transforms():
    """nndet/io/transforms/instances.py executed from its file with a stand-in for AbstractTransform (the nndet.io package
    __init__ pulls SimpleITK-dependent modules); the three transforms are chained exactly like the module's pre_trafo."""
    import importlib.util, types
    from oracle import transform_oracle as to

    class _AT(torch.nn.Module):
        def __init__(self, grad=False, **kw):
            super().__init__(); self.grad = grad

        def __call__(self, **data):
            return self.forward(**data)
    base = types.ModuleType("nndet.io.transforms.base"); base.AbstractTransform = _AT
    for name in ("nndet.io", "nndet.io.transforms"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["nndet.io.transforms.base"] = base
    spec = importlib.util.spec_from_file_location("ref_inst", os.path.join(ref_import.REF_ROOT, "nndet/io/transforms/instances.py"))
    ri = importlib.util.module_from_spec(spec); spec.loader.exec_module(ri)
    out = {}
    for ci, (B, shape, seed) in enumerate(TRANSFORM_CASES):
        t, maps = synth_instances(B, shape, seed)
        data = {"target": torch.from_numpy(t.copy()), "instance_mapping": maps}
        data = ri.FindInstances(instance_key="target", save_key="present_instances")(**data)
        data = ri.Instances2Boxes(instance_key="target", map_key="instance_mapping", box_key="boxes", class_key="classes",
                                  present_instances="present_instances")(**data)
        data = ri.Instances2Segmentation(instance_key="target", map_key="instance_mapping", present_instances="present_instances")(**data)
        p, bx, cl, sem = to.pre_trafo(t, maps)
        for b in range(B):
            assert np.array_equal(data["present_instances"][b].numpy(), p[b])
            rb = data["boxes"][b].numpy()
            assert rb.shape == bx[b].shape and np.array_equal(rb, bx[b])
            assert np.array_equal(data["classes"][b].numpy(), cl[b])
            out[f"c{ci}_ids{b}"] = p[b]; out[f"c{ci}_boxes{b}"] = rb; out[f"c{ci}_classes{b}"] = cl[b]
        assert np.array_equal(data["target"].numpy(), sem)
        out[f"c{ci}_sem_crc"] = np.asarray([zlib.crc32(sem.tobytes())], dtype=np.int64)
        out[f"c{ci}_sem_sum"] = np.asarray([sem.sum()], dtype=np.float64)
    save("transforms", **out)


# ------------------------------------------------------------------------------------------ case-level ensembling
def synth_tile_predictions(seed, case_shape=(64, 96, 80), tile=(32, 48, 40), n_models=2):
    """Per model / TTA pass: a list of tile batches (2 tiles per batch) with random boxes in tile coordinates."""
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
                lo = torch.rand(n, 3, generator=g) * torch.tensor(tile) * 0.8 - 2.0          # some boxes stick out of the tile
                sz = torch.rand(n, 3, generator=g) * 10 + 1.0
                res["pred_boxes"].append(torch.stack([lo[:, 0], lo[:, 1], lo[:, 0] + sz[:, 0], lo[:, 1] + sz[:, 1], lo[:, 2],
                                                      lo[:, 2] + sz[:, 2]], dim=1))
                res["pred_scores"].append(torch.empty(n))                                    # filled below (unique scores)
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


def gen_ensembler():
    """BoxEnsemblerSelective (nndet/inference/ensembler/detection.py:901-1130) executed from its files; the package __init__
    modules that pull ITK / predictor code are replaced by path-only stand-ins, `nndet.io.load` / `nndet.inference.restore` by stubs."""
    import importlib, types
    root = ref_import.REF_ROOT
    for name, sub in (("nndet.inference", "nndet/inference"), ("nndet.inference.ensembler", "nndet/inference/ensembler")):
        if name not in sys.modules or not hasattr(sys.modules[name], "__path__"):
            pkg = types.ModuleType(name); pkg.__path__ = [os.path.join(root, sub)]; sys.modules[name] = pkg
    io = sys.modules.setdefault("nndet.io", types.ModuleType("nndet.io"))
    load = types.ModuleType("nndet.io.load"); load.save_pickle = lambda *a, **k: None
    sys.modules["nndet.io.load"] = load; io.load = load
    rest = types.ModuleType("nndet.inference.restore"); rest.restore_detection = lambda boxes, **k: boxes
    sys.modules["nndet.inference.restore"] = rest
    det = importlib.import_module("nndet.inference.ensembler.detection")
    from nndetection_b200.inference import ensembler as mine

    def o_weighted_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
        keep = bo.batched_nms(boxes, scores * weights, labels, iou_thresh, cuda_semantics=False)
        return boxes[keep], scores[keep], labels[keep], torch.ones_like(weights)[keep]

    def o_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
        keep = bo.batched_nms(boxes, scores, labels, iou_thresh, cuda_semantics=False)
        return boxes[keep], scores[keep], labels[keep], weights[keep]

    def o_wbc_ensemble(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh, *a, **k):
        return bo.batched_wbc(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh)

    out = {}
    for ci, (seed, overrides, mine_over) in enumerate([
            (1, {}, {"model_nms_fn": o_weighted_nms_model, "ensemble_nms_fn": o_wbc_ensemble}),
            (2, {"model_iou": 0.3, "ensemble_iou": 0.2, "model_score_thresh": 0.1, "remove_small_boxes": 2.0,
                 "model_nms_fn": det.batched_nms_model},
             {"model_iou": 0.3, "ensemble_iou": 0.2, "model_score_thresh": 0.1, "remove_small_boxes": 2.0,
              "model_nms_fn": o_nms_model, "ensemble_nms_fn": o_wbc_ensemble})]):
        models, shape = synth_tile_predictions(seed)
        props = {"shape": shape, "original_size_of_raw_data": shape, "itk_origin": (0, 0, 0), "itk_spacing": (1, 1, 1),
                 "itk_direction": (1, 0, 0, 0, 1, 0, 0, 0, 1)}
        rp = det.BoxEnsemblerSelective.get_default_parameters(); rp.update(overrides)
        ref = det.BoxEnsemblerSelective(properties=dict(props), parameters=rp)
        mp = mine.BoxEnsemblerSelective.get_default_parameters(); mp.update(mine_over)
        my = mine.BoxEnsemblerSelective(properties=dict(props), parameters=mp)
        for mi, batches in enumerate(models):
            ref.add_model(name=f"model0_t{mi}", model_weight=1.0 if mi == 0 else 0.7)
            my.add_model(name=f"model0_t{mi}", model_weight=1.0 if mi == 0 else 0.7)
            for res, batch in batches:
                ref.process_batch(result={k: [t.clone() for t in v] for k, v in res.items()}, batch=batch)
                my.process_batch(result={k: [t.clone() for t in v] for k, v in res.items()}, batch=batch)
        r, m = ref.get_case_result(restore=False), my.get_case_result(restore=False)
        for k in ("pred_boxes", "pred_scores", "pred_labels"):
            assert r[k].shape == m[k].shape and torch.equal(r[k].float(), m[k].float()), (ci, k)
        assert r["pred_boxes"].shape[0] > 10
        out[f"c{ci}_boxes"] = r["pred_boxes"]; out[f"c{ci}_scores"] = r["pred_scores"]; out[f"c{ci}_labels"] = r["pred_labels"]
    save("ensembler", **out)


# ------------------------------------------------------------------------------------------ sliding-window predictor
class FakeDetector:
    """Deterministic stand-in for a model: detections are a function of the tile content only (so mirrored tiles give mirrored
    detections only if the pipeline un-mirrors them correctly)."""

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
            # unique scores: the same voxel is seen by overlapping tiles and by every mirrored copy -- a content- and
            # orientation-dependent offset keeps all scores distinct (torch.sort leaves the order of ties unspecified)
            ramp = torch.arange(flat.numel(), dtype=torch.float32)
            salt = float((flat * ((ramp * 0.6180339887) % 1.0)).sum() % 1.0)
            out["pred_scores"].append((flat[idx] * 0.9 + 0.1 * salt).clone())
            out["pred_labels"].append(((z + y + x) % 2).long())
        return out


GRID_CASES = [(32, 40, 16), (32, 56, 16), (32, 32, 16), (48, 40, 24), (16, 100, 4), (20, 61, 10), (8, 8, 0)]


def gen_predictor():
    """Tile grid (nndet/io/patching.py), Mirror (nndet/io/transforms/spatial.py) and BoxEnsemblerSelective executed from their files;
    the predictor loop of nndet/inference/predictor.py:192-306 is restated around them (its module needs a DataLoader / ITK stack)."""
    import importlib, importlib.util, types
    root = ref_import.REF_ROOT
    sk = types.ModuleType("skimage"); skm = types.ModuleType("skimage.measure"); skm.regionprops = None
    sys.modules.setdefault("skimage", sk); sys.modules.setdefault("skimage.measure", skm)

    class _AT(torch.nn.Module):
        def __init__(self, grad=False, **kw):
            super().__init__(); self.grad = grad

        def __call__(self, **data):
            return self.forward(**data)
    base = types.ModuleType("nndet.io.transforms.base"); base.AbstractTransform = _AT
    for name in ("nndet.io", "nndet.io.transforms"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["nndet.io.transforms.base"] = base

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(root, rel))
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m
    rpatch = load("ref_patching", "nndet/io/patching.py")
    rspat = load("ref_spatial", "nndet/io/transforms/spatial.py")
    for name, sub in (("nndet.inference", "nndet/inference"), ("nndet.inference.ensembler", "nndet/inference/ensembler")):
        if name not in sys.modules or not hasattr(sys.modules[name], "__path__"):
            pkg = types.ModuleType(name); pkg.__path__ = [os.path.join(root, sub)]; sys.modules[name] = pkg
    ld = types.ModuleType("nndet.io.load"); ld.save_pickle = lambda *a, **k: None; sys.modules["nndet.io.load"] = ld
    rest = types.ModuleType("nndet.inference.restore"); rest.restore_detection = lambda boxes, **k: boxes
    sys.modules["nndet.inference.restore"] = rest
    det = importlib.import_module("nndet.inference.ensembler.detection")
    from nndetection_b200.inference import predictor as mp, ensembler as me

    out = {}
    # ---- grids: both modes, with / without centre-border crops
    for gi, (ps, dl, ov) in enumerate(GRID_CASES):
        for mode in ("fixed", "symmetric"):
            for cb in (False, True):
                r = rpatch.create_grid((ps, ps), (dl, dl + 7), (ov, ov), mode=mode, center_boarder=cb)
                m = mp.create_grid((ps, ps), (dl, dl + 7), (ov, ov), mode=mode, center_boarder=cb)
                assert [[(s.start, s.stop) for s in c] for c in r] == [[(s.start, s.stop) for s in c] for c in m]
                out[f"grid{gi}_{mode}_{int(cb)}"] = np.asarray([[(s.start, s.stop) for s in c] for c in r], dtype=np.int64)
    # ---- box mirroring
    g = torch.Generator().manual_seed(3)
    bx = rand_boxes(50, g, extent=30.0, lo=1.0, hi=6.0)
    for dims in [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2), (0, 1, 2)]:
        r = rspat.Mirror(keys=["pred_seg"], box_keys=["pred_boxes"], dims=dims)(pred_seg=torch.zeros(1, 2, 32, 40, 24), pred_boxes=[bx])
        assert torch.equal(r["pred_boxes"][0], mp.mirror_boxes(bx, dims, (32, 40, 24)))
    # ---- patch larger than the case: the predictor's fallback save_get_crop(mode="symmetric") (predictor.py:223-228)
    small = torch.rand(2, 5, 40, 7, generator=g)
    for crop in rpatch.create_grid(cshape=(16, 32, 16), dshape=(5, 40, 7), overlap=[8, 16, 8], mode="symmetric"):
        r_tile, r_origin, r_crop = rpatch.save_get_crop(small.numpy(), crop, mode="symmetric")
        m_tile, m_origin, m_crop = mp.padded_crop_symmetric(small, crop)
        assert np.array_equal(r_tile, m_tile.numpy()) and list(r_origin) == m_origin and list(r_crop) == m_crop
    # ---- whole loop: case -> tiles -> 8 mirror TTAs -> ensembler
    g = torch.Generator().manual_seed(17)
    case = {"data": torch.rand(1, 40, 56, 48, generator=g).numpy()}
    crop_size, bs = (32, 32, 32), 4
    props = {"transpose_backward": None, "original_spacing": None, "spacing_after_resampling": None, "crop_bbox": None,
             "original_size_of_raw_data": (40, 56, 48), "itk_origin": (0, 0, 0), "itk_spacing": (1, 1, 1), "itk_direction": (1, 0, 0, 0, 1, 0, 0, 0, 1)}
    model = FakeDetector()
    ens = det.BoxEnsemblerSelective.from_case(case, props, parameters={})
    crops = rpatch.create_grid(cshape=crop_size, dshape=case["data"].shape[1:], overlap=[int(c * 0.5) for c in crop_size], mode="symmetric")
    tiles = []
    for crop in crops:                                                    # predictor.py:214-235
        tile = {"data": rpatch.save_get_crop(case["data"], crop, mode="shift")[0]}
        _, tile["tile_origin"], tile["crop"] = rpatch.save_get_crop(case["data"], crop, mode="shift")
        tiles.append(tile)
    for t, dims in enumerate(mp.get_tta_dims(8)):                         # predictor.py:258-273, inference/transforms.py:25-72
        ens.add_model(name=f"model0_t{t}", model_weight=1.0)
        for i in range(0, len(tiles), bs):
            chunk = tiles[i:i + bs]
            batch = {"data": torch.stack([torch.from_numpy(np.ascontiguousarray(c["data"])) for c in chunk]),
                     "tile_origin": [torch.tensor([c["tile_origin"][ax] for c in chunk]) for ax in range(3)]}
            tr = rspat.Mirror(keys=["data"], dims=dims)(**batch) if dims else batch
            res = model.inference_step(tr["data"])
            if dims:
                res = rspat.Mirror(keys=["pred_seg"], box_keys=["pred_boxes"], dims=dims)(**res)
            ens.process_batch(result=res, batch=batch)
    ref = ens.get_case_result(restore=False)

    def o_weighted_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
        keep = bo.batched_nms(boxes, scores * weights, labels, iou_thresh, cuda_semantics=False)
        return boxes[keep], scores[keep], labels[keep], torch.ones_like(weights)[keep]

    def o_wbc_ensemble(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh, *a, **k):
        return bo.batched_wbc(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh)
    pred = mp.SlidingWindowPredictor(
        ensembler_fn=lambda c, properties=None: me.BoxEnsemblerSelective.from_case(
            c, properties, parameters={"model_nms_fn": o_weighted_nms_model, "ensemble_nms_fn": o_wbc_ensemble}),
        models=[model], crop_size=crop_size, overlap=0.5, num_tta_transforms=8, batch_size=bs, device="cpu")
    mine_res = pred.predict_case({"data": torch.from_numpy(case["data"])}, properties=props)["boxes"]
    for k in ("pred_boxes", "pred_scores", "pred_labels"):
        assert ref[k].shape == mine_res[k].shape and torch.equal(ref[k].float(), mine_res[k].float()), k
    print(f"  predictor case: {len(tiles)} tiles, {ref['pred_boxes'].shape[0]} final boxes")
    assert ref["pred_boxes"].shape[0] >= 5 and len(tiles) >= 8
    out["case_boxes"] = ref["pred_boxes"]; out["case_scores"] = ref["pred_scores"]; out["case_labels"] = ref["pred_labels"]
    out["tile_origins"] = np.asarray([t["tile_origin"] for t in tiles], dtype=np.int64)
    save("predictor", **out)


# ------------------------------------------------------------------------------------------ predict_dir + on-disk formats
HELPER_PROPS = {"original_size_of_raw_data": (40, 56, 48), "itk_origin": (0.0, 0.0, 0.0), "itk_spacing": (1.0, 1.0, 1.0),
                "itk_direction": (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)}


def gen_helper():
    """`get_case_id_from_path` (nndet/io/paths.py), `save_pickle` / `load_pickle` (nndet/io/load.py) and `to_numpy` (nndet/utils/tensor.py)
    executed from their files: the `<case>_boxes.pkl` the reference's `predict_dir` (nndet/inference/helper.py:109-110) writes for the
    case result of tests/golden/predictor.npz (itself produced by the executed ensembler), against the file this package's `predict_dir`
    writes for the same case -- byte for byte here, content (key order, dtypes, values) in the fixture."""
    import importlib.util, tempfile, types, pickle
    root = ref_import.REF_ROOT

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(root, rel))
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m
    rpaths = load("ref_paths", "nndet/io/paths.py")
    for name in ("nndet.io",):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    saved = sys.modules.get("nndet.io.paths")
    sys.modules["nndet.io.paths"] = rpaths
    rload = load("ref_load", "nndet/io/load.py")
    if saved is not None:
        sys.modules["nndet.io.paths"] = saved
    rtensor = load("ref_tensor", "nndet/utils/tensor.py")
    from nndetection_b200.inference import helper as mh

    names = ["/data/Task000/imagesTr/case_001_0000.nii.gz", "/x/y/LUNA_17.npz", "/p/q.r/abc.def.npy", "/a/b_0000.nii.gz"]
    for n in names:
        for rm in (True, False):
            assert rpaths.get_case_id_from_path(n, remove_modality=rm) == mh.get_case_id_from_path(n, remove_modality=rm), (n, rm)

    g = np.load(os.path.join(OUT, "predictor.npz"))
    ref_result = {"boxes": {"pred_boxes": torch.from_numpy(g["case_boxes"]), "pred_scores": torch.from_numpy(g["case_scores"]),
                            "pred_labels": torch.from_numpy(g["case_labels"]), "restore": False, **HELPER_PROPS}}

    def o_weighted_nms_model(boxes, scores, labels, weights, iou_thresh, *a, **k):
        keep = bo.batched_nms(boxes, scores * weights, labels, iou_thresh, cuda_semantics=False)
        return boxes[keep], scores[keep], labels[keep], torch.ones_like(weights)[keep]

    def o_wbc_ensemble(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh, *a, **k):
        return bo.batched_wbc(boxes, scores, labels, weights, iou_thresh, n_exp_preds, score_thresh)

    with tempfile.TemporaryDirectory() as td:
        src, dst_ref, dst_mine = os.path.join(td, "src"), os.path.join(td, "ref"), os.path.join(td, "mine")
        for d in (src, dst_ref, dst_mine):
            os.makedirs(d)
        gen = torch.Generator().manual_seed(17)
        np.savez(os.path.join(src, "case_a.npz"), data=torch.rand(1, 40, 56, 48, generator=gen).numpy())
        rload.save_pickle(dict(HELPER_PROPS), os.path.join(src, "case_a"))                 # suffix added by the reference helper
        assert os.path.isfile(os.path.join(src, "case_a.pkl"))
        assert mh.load_pickle(os.path.join(src, "case_a")) == rload.load_pickle(os.path.join(src, "case_a")) == HELPER_PROPS
        for key, item in rtensor.to_numpy(ref_result).items():                             # nndet/inference/helper.py:109-110
            rload.save_pickle(item, os.path.join(dst_ref, f"case_a_{key}.pkl"))
        plan = {"patch_size": (32, 32, 32), "batch_size": 4, "network_dim": 3, "transpose_backward": [0, 1, 2],
                "inference_plan": {"model_nms_fn": o_weighted_nms_model, "ensemble_nms_fn": o_wbc_ensemble}}
        mh.predict_dir(src, dst_mine, cfg={}, plan=plan, source_models=td, model_fn=lambda *a: [{"model": FakeDetector(), "rank": 0}],
                       num_models=1, device="cpu")
        a = open(os.path.join(dst_ref, "case_a_boxes.pkl"), "rb").read()
        b = open(os.path.join(dst_mine, "case_a_boxes.pkl"), "rb").read()
        assert a == b, "case_a_boxes.pkl differs from the file the reference helpers write"
        res = pickle.loads(a)
    # ---- checkpoints: the REAL reference network inside a stand-in LightningModule (attribute `model`, nndet/ptmodule/base_module.py:55-59)
    #      written the way Lightning does ({"state_dict": module.state_dict()}) and read the way loading.py:96-97 does
    from nndetection_b200.ptmodule import RetinaUNetV001
    arch, anc, patch, bs = mo.make_plan("tiny")

    class LM(torch.nn.Module):
        def __init__(self, net):
            super().__init__(); self.model = net
    torch.manual_seed(1)
    lm = LM(build_reference_model(dict(arch), dict(anc)))
    with tempfile.TemporaryDirectory() as td:
        torch.save({"state_dict": lm.state_dict(), "epoch": 3, "global_step": 7500}, os.path.join(td, "model_last.ckpt"))
        mine = mh.load_final_model(td, {"model_cfg": None}, {"architecture": arch, "anchors": anc}, num_models=1, identifier="last", device=None)[0]["model"]
        for k, v in lm.model.state_dict().items():
            assert torch.equal(v, mine.state_dict()[k]), k
        torch.manual_seed(2)
        other = RetinaUNetV001.from_config_plan(None, arch, anc)
        mh.save_checkpoint(other, os.path.join(td, "model_best.ckpt"), epoch=1)
        state_dict = torch.load(os.path.join(td, "model_best.ckpt"), map_location="cpu")["state_dict"]      # loading.py:96
        t = lm.load_state_dict(state_dict)                                                                    # loading.py:97 (strict)
        assert not t.missing_keys and not t.unexpected_keys
        for k, v in other.state_dict().items():
            assert torch.equal(v.float(), lm.model.state_dict()[k]), k
        assert len(mh.load_all_models(td, {"model_cfg": None}, {"architecture": arch, "anchors": anc}, device=None)) == 2
    print("  checkpoints: reference-layout .ckpt loads into this package's network and back (strict, all tensors equal)")
    out = {"keys": np.asarray(list(res.keys())), "dtypes": np.asarray([str(getattr(v, "dtype", type(v).__name__)) for v in res.values()])}
    for k in ("pred_boxes", "pred_scores", "pred_labels"):
        out[k] = res[k]
    print(f"  case_a_boxes.pkl: {len(a)} bytes identical, keys {list(res.keys())}")
    save("helper", **out)


# ------------------------------------------------------------------------------------------ post-processing sweep
def gen_sweeper():
    """`BoxSweeper` (nndet/inference/sweeper.py:78-219) EXECUTED from its file on saved ensembler states, against this package's
    mirror: identical determined parameters and identical score per swept value, (a) with the reference's real `BoxEvaluator` (COCO mAP;
    matplotlib stubbed) and (b) with the small stand-in evaluator the committed test can run; (b) goes into tests/golden/sweeper.npz."""
    import importlib, importlib.util, json, tempfile, types
    root = ref_import.REF_ROOT
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tutil
    from nndetection_b200.inference.sweeper import BoxSweeper

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(root, rel))
        m = importlib.util.module_from_spec(spec); sys.modules[modname] = m; spec.loader.exec_module(m); return m
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.ticker"].FuncFormatter = object
    sys.modules["matplotlib.pyplot"].Figure = sys.modules["matplotlib.pyplot"].Axes = object      # only used in annotations
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "nndet.io" not in sys.modules or not hasattr(sys.modules["nndet.io"], "__path__"):
        pkg = types.ModuleType("nndet.io"); pkg.__path__ = []; sys.modules["nndet.io"] = pkg
    load("nndet.io.paths", "nndet/io/paths.py")
    load("nndet.io.load", "nndet/io/load.py")
    from nndet.evaluator.registry import BoxEvaluator
    rsw = load("ref_sweeper", "nndet/inference/sweeper.py")
    os.environ["det_verbose"] = "0"

    ens_cls = tutil.oracle_ensembler_cls()
    classes = ["class0", "class1"]
    out = {}
    with tempfile.TemporaryDirectory() as td:
        pred, gt = os.path.join(td, "pred"), os.path.join(td, "gt")
        tutil.write_sweep_cases(pred, gt)
        for tag, ev, metric in (("coco", BoxEvaluator, "mAP_IoU_0.10_0.50_0.05_MaxDet_100"), ("standin", tutil.StandInEvaluator, "stand_in")):
            ref = rsw.BoxSweeper(classes=classes, pred_dir=pred, gt_dir=gt, target_metric=metric, ensembler_cls=ens_cls,
                                 save_dir=os.path.join(td, f"ref_{tag}"))
            ref.evaluator_cls = ev
            state_ref = ref.run_postprocessing_sweep()
            mine = BoxSweeper(classes, pred, gt, metric, ens_cls, save_dir=os.path.join(td, f"mine_{tag}"), evaluator_cls=ev, device="cpu")
            state_mine = mine.run_postprocessing_sweep()
            assert list(state_ref.keys()) == list(state_mine.keys())
            for k in state_ref:
                assert state_ref[k] == state_mine[k] or state_ref[k] is state_mine[k], (tag, k, state_ref[k], state_mine[k])
            for f in sorted(os.listdir(os.path.join(td, f"ref_{tag}"))):
                a = json.load(open(os.path.join(td, f"ref_{tag}", f))); b = json.load(open(os.path.join(td, f"mine_{tag}", f)))
                assert list(a.keys()) == list(b.keys()), f
                for k in a:
                    if k.startswith("best_"):
                        assert a[k] == b[k], (tag, f, k)
                    else:
                        assert a[k]["scores"] == b[k]["scores"] and a[k]["overwrite"] == b[k]["overwrite"], (tag, f, k)
                if tag == "standin":
                    vals = [float(eval(v["scores"], {"np": np, "nan": float("nan")})["stand_in"]) for k, v in a.items() if not k.startswith("best_")]
                    assert len(set(vals)) == len(vals) or np.argmax(vals) == 0 or sorted(vals)[-1] > sorted(vals)[-2], f"tie at the top in {f}"
                    out["scores_" + f[len("sweep_"):-len(".json")]] = np.asarray(vals)
            print(f"  {tag}: determined", {k: (v.__name__ if callable(v) else float(v)) for k, v in state_ref.items()
                                           if k in ("model_iou", "model_nms_fn", "ensemble_iou", "model_score_thresh", "remove_small_boxes")})
            if tag == "standin":
                for k in ("model_iou", "ensemble_iou", "model_score_thresh", "remove_small_boxes"):
                    out["state_" + k] = np.float64(state_ref[k])
                out["state_model_nms_fn"] = np.asarray(state_ref["model_nms_fn"].__name__)
    save("sweeper", **out)


# ------------------------------------------------------------------------------------------ learning-rate schedule
def gen_restore():
    """`restore_detection` (nndet/inference/restore.py:30-66, with permute_boxes / expand_to_boxes of core/boxes/ops.py:330-374) executed
    from its file (its module-level imports of loguru / the ITK-side resampling helpers are stubbed: the function uses neither), as
    `BoxEnsembler.restore_prediction` calls it (ensembler/detection.py:254-274): float32 boxes in, float64 numpy arithmetic, float32 out."""
    import importlib, types
    root = ref_import.REF_ROOT
    if "nndet.inference" not in sys.modules or not hasattr(sys.modules["nndet.inference"], "__path__"):
        pkg = types.ModuleType("nndet.inference"); pkg.__path__ = [os.path.join(root, "nndet/inference")]; sys.modules["nndet.inference"] = pkg
    if "loguru" not in sys.modules:
        lg = types.ModuleType("loguru"); lg.logger = types.SimpleNamespace(info=print, warning=print, error=print); sys.modules["loguru"] = lg
    if "nndet.preprocessing.resampling" not in sys.modules:
        pp = sys.modules.setdefault("nndet.preprocessing", types.ModuleType("nndet.preprocessing"))
        rs = types.ModuleType("nndet.preprocessing.resampling")
        rs.resample_data_or_seg = rs.get_do_separate_z = rs.get_lowres_axis = None
        sys.modules["nndet.preprocessing.resampling"] = rs; pp.resampling = rs
    sys.modules.pop("nndet.inference.restore", None)
    restore = importlib.import_module("nndet.inference.restore")
    out = {}
    cases = [((0, 1, 2), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0), [(0, 10), (0, 20), (0, 30)]),
             ((2, 0, 1), (2.5, 0.7, 0.7), (0.8, 1.25, 0.9), [(3, 90), (11, 200), (7, 150)]),
             ((1, 2, 0), (0.5, 0.5, 3.0), (1.0, 1.0, 1.0), [(0, 64), (5, 69), (17, 81)]),
             ((2, 1, 0), (1.37, 0.91, 0.66), (1.5, 0.75, 0.75), [(21, 50), (0, 40), (9, 99)])]
    g = torch.Generator().manual_seed(77)
    for i, (tb, osp, rsp, crop) in enumerate(cases):
        boxes = rand_boxes(40 + i, g).float()
        props = dict(transpose_backward=list(tb), original_spacing=np.asarray(osp), spacing_after_resampling=np.asarray(rsp), crop_bbox=crop)
        ref = restore.restore_detection(boxes.numpy(), **props)
        res = torch.from_numpy(ref).to(dtype=boxes.dtype)
        out[f"boxes{i}"] = boxes.numpy(); out[f"restored{i}"] = res.numpy()
        out[f"tb{i}"] = np.asarray(tb); out[f"osp{i}"] = np.asarray(osp); out[f"rsp{i}"] = np.asarray(rsp); out[f"crop{i}"] = np.asarray(crop)
        from nndetection_b200.inference.ensembler import BoxEnsemblerSelective
        mine = BoxEnsemblerSelective.from_case({"data": torch.zeros(1, 8, 8, 8)}, properties=props)
        assert torch.equal(mine.restore_prediction(boxes), res), i
    out["n_cases"] = np.asarray(len(cases))
    save("restore", **out)


def gen_lr():
    """LinearWarmupPolyLR (nndet/training/learning_rate.py:126-183) executed: the lr the optimizer holds at every step."""
    import importlib.util
    from nndetection_b200.training import poly_lr
    spec = importlib.util.spec_from_file_location("ref_lr", os.path.join(ref_import.REF_ROOT, "nndet/training/learning_rate.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    cfgs = [(0.01, 40, 1e-6, 0.9, 200), (0.01, 4000, 1e-6, 0.9, 6000)]
    out = {"cfgs": np.asarray(cfgs, dtype=np.float64)}
    for i, (lr0, warm, wlr, gamma, n) in enumerate(cfgs):
        p_ = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p_], lr=lr0)
        sch = m.LinearWarmupPolyLR(opt, warm_iterations=warm, warm_lr=wlr, poly_gamma=gamma, num_iterations=n)
        lrs = []
        for step in range(n - 1):
            lrs.append(opt.param_groups[0]["lr"])
            assert lrs[-1] == poly_lr(step, lr0, warm, wlr, gamma, n), (i, step, lrs[-1], poly_lr(step, lr0, warm, wlr, gamma, n))
            opt.step(); sch.step()
        out[f"lrs{i}"] = np.asarray(lrs, dtype=np.float64)
    save("lr", **out)


# ------------------------------------------------------------------------------------------ box metrics
def gen_pairwise():
    g = torch.Generator().manual_seed(11)
    a, b = rand_boxes(37, g), rand_boxes(301, g)
    b[5] = a[3]                                   # identical pair
    b[6] = a[3] + 1000.0                          # disjoint
    iou = rops.box_iou(a, b); giou = rops.generalized_box_iou(a, b, eps=1e-7)
    dist = rops.box_center_dist(a, b)[0]
    assert torch.equal(iou, bo.box_iou(a, b)), "oracle box_iou != reference"
    assert torch.equal(giou, bo.generalized_box_iou(a, b, eps=1e-7)), "oracle giou != reference"
    assert torch.equal(dist, bo.box_center_dist(a, b)), "oracle center_dist != reference"
    save("pairwise", a=a, b=b, iou=iou, giou=giou, dist=dist)


# ------------------------------------------------------------------------------------------ anchors
def gen_anchors():
    for name in ("tiny", "toy", "luna"):
        arch, anc, patch, _ = mo.make_plan(name)
        fm = []
        for lvl in arch["decoder_levels"]:
            s = [1, 1, 1]
            for st in arch["strides"][:lvl]:
                s = [a * b for a, b in zip(s, st)]
            fm.append([p // q for p, q in zip(patch, s)])
        gen = AnchorGenerator3DS(**anc)
        img = torch.zeros(1, 1, *patch)
        fmaps = [torch.zeros(1, 1, *f) for f in fm]
        ref = gen(img, fmaps)[0]
        per_level = gen.get_num_acnhors_per_level()
        mine, pl = bo.anchors_for_image(patch, fm, anc["width"], anc["height"], anc["depth"])
        assert torch.equal(ref, mine) and list(per_level) == list(pl), f"anchors differ for {name}"
        # store a strided sample + checksums (full luna tensor is 24 MB)
        step = max(1, ref.shape[0] // 4096)
        save(f"anchors_{name}", patch=patch, fmap_sizes=np.asarray(fm), width=np.asarray(anc["width"]),
             per_level=np.asarray(per_level), sample_idx=np.arange(0, ref.shape[0], step),
             sample=ref[::step], colsum=ref.double().sum(0), crc=zlib.crc32(ref.numpy().tobytes()))


# ------------------------------------------------------------------------------------------ ATSS
def gen_atss():
    arch, anc, patch, _ = mo.make_plan("toy")
    fm = [[16, 32, 32], [8, 16, 16]]
    anchors, per_level = bo.anchors_for_image(patch, fm, anc["width"], anc["height"], anc["depth"])
    matcher = ATSSMatcher(num_candidates=4, similarity_fn=rops.box_iou, center_in_gt=False)
    cases = {}
    for seed in range(4):
        _, tg = mo.synth_batch(patch, 2, 1, 2, 100 + seed, max_gt=6)
        gt = tg["target_boxes"][0]
        iou_r, m_r = matcher(gt, anchors, num_anchors_per_level=per_level, num_anchors_per_loc=27)
        iou_o, m_o = bo.atss_match(gt, anchors, per_level, 27, 4, canonical_ties=True)
        assert torch.equal(m_r, m_o), f"ATSS matches differ (seed {seed}): {(m_r != m_o).sum()}"
        assert torch.equal(iou_r, iou_o)
        cases[f"gt{seed}"] = gt
        cases[f"pos_idx{seed}"] = torch.where(m_r >= 0)[0]
        cases[f"pos_gt{seed}"] = m_r[m_r >= 0]
    # no-GT image
    _, m_r = matcher(torch.zeros(0, 6), anchors, num_anchors_per_level=per_level, num_anchors_per_loc=27)
    assert (m_r == -1).all()
    save("atss_toy", patch=patch, fmap_sizes=np.asarray(fm), per_level=np.asarray(per_level), n_cases=4, **cases)


# ------------------------------------------------------------------------------------------ sampler
def gen_sampler():
    smp = HardNegativeSamplerBatched(batch_size_per_image=32, positive_fraction=0.33, min_neg=1, pool_size=20)
    rows = []
    for npos, nneg, bs in [(0, 1000, 4), (1, 1000, 4), (5, 100000, 4), (500, 4000000, 4), (10, 3, 2), (0, 0, 1),
                           (41, 50, 4), (43, 5000, 4), (7, 900, 1)]:
        smp.batch_size_per_image = smp._batch_size_per_image * bs
        p = smp.get_num_pos(torch.zeros(npos))
        n = smp.get_num_neg(torch.zeros(nneg), p)
        pool = min(nneg, int(n * smp.pool_size))
        assert (p, n, pool) == bo.hnm_counts(npos, nneg, bs), (npos, nneg, bs)
        rows.append([npos, nneg, bs, p, n, pool])
    # pool set: reference topk over fg_probs[negative] (tie-free probs)
    g = torch.Generator().manual_seed(5)
    A = 50000
    labels = torch.zeros(A); labels[torch.randperm(A, generator=g)[:60]] = 1.0
    labels[torch.randperm(A, generator=g)[:500]] = -1.0
    probs = torch.rand(A, generator=g)
    negative = torch.where(labels == 0)[0]
    _, _, pool = bo.hnm_counts(int((labels >= 1).sum()), negative.numel(), 4)
    ref_pool = torch.sort(negative[probs[negative].topk(pool, sorted=True)[1]])[0]
    assert torch.equal(ref_pool, bo.hnm_pool(labels, probs, pool))
    pos, neg, _ = bo.hnm_select(labels, probs, 4, seed=77)
    save("sampler", counts=np.asarray(rows), labels=labels, probs=probs, pool=ref_pool, pool_size=pool,
         hash_pos=pos, hash_neg=neg, hash_seed=77)


# ------------------------------------------------------------------------------------------ coder / clip
def gen_coder():
    g = torch.Generator().manual_seed(3)
    anchors = rand_boxes(2000, g, extent=128.0)
    rel = torch.randn(2000, 6, generator=g) * 0.5
    rel[0, 2] = 10.0                                # exercises the log(1000/16) clamp
    coder = BoxCoderND(weights=(1.0,) * 6)
    dec = coder.decode_single(rel, anchors)
    assert torch.equal(dec, bo.decode_single(rel, anchors)), "decode differs"
    clipped = clip_boxes_to_image_(dec.clone(), (128, 128, 128))
    assert torch.equal(clipped, bo.clip_boxes_3d(dec, (128, 128, 128)))
    keep = rops.remove_small_boxes(clipped, 0.01)
    assert torch.equal(keep, bo.keep_not_small(clipped, 0.01))
    save("coder", anchors=anchors, rel=rel, decoded=dec, clipped=clipped, keep=keep)


# ------------------------------------------------------------------------------------------ NMS
def gen_nms():
    out = {}
    cases = []
    for n in (0, 1, 2, 63, 64, 65, 129, 1000, 3000, 10000):          # 10 000 = BASELINE's headline NMS size
        for thr in (1e-5, 0.1, 0.5, 0.6, 0.9):
            if (n == 3000 and thr not in (0.1, 0.6)) or (n == 10000 and thr != 0.1):
                continue
            g = torch.Generator().manual_seed(1000 + n)
            boxes = rand_boxes(n, g, extent=60.0 if n <= 129 else 160.0)
            scores = unique_scores(n, g)
            keep = rnms.nms(boxes, scores, thr) if n > 0 else torch.empty(0, dtype=torch.int64)
            mine = bo.nms_greedy(boxes, scores, thr, cuda_semantics=True)
            mine_cpu = bo.nms_greedy(boxes, scores, thr, cuda_semantics=False)
            assert torch.equal(keep, mine_cpu), f"nms_cpu restatement differs n={n} thr={thr}"
            assert torch.equal(keep, mine), f"cuda-semantics restatement differs on NaN-free input n={n} thr={thr}"
            key = f"n{n}_t{thr}"
            out[key + "_keep"] = keep
            cases.append((n, thr))
    # batched (3 classes)
    g = torch.Generator().manual_seed(4242)
    boxes = rand_boxes(1500, g); scores = unique_scores(1500, g)
    idxs = torch.randint(0, 3, (1500,), generator=g)
    keep = rnms.batched_nms(boxes, scores, idxs, 0.5)
    assert torch.equal(keep, bo.batched_nms(boxes, scores, idxs, 0.5))
    out["batched_keep"] = keep
    save("nms", cases=np.asarray(cases), **out)
    # NOTE: inputs are regenerated from the seeds by tests (rand_boxes/unique_scores live in tests/util.py too)


# ------------------------------------------------------------------------------------------ model
def det_fill(sd, seed=0):
    """Deterministic, platform-independent weights keyed by parameter name."""
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


def build_reference_model(arch, anc):
    """Mirror of RetinaUNetModule.from_config_plan (nndet/ptmodule/retinaunet/base.py:387-466)
    using the reference's own classes (pytorch_lightning is absent so the LightningModule cannot import)."""
    conv_b = Generator(ConvInstanceRelu, 3)
    conv_h = Generator(ConvGroupRelu, 3)
    enc = Encoder(conv=conv_b, conv_kernels=arch["conv_kernels"], strides=arch["strides"],
                  block_cls=StackedConvBlock2, in_channels=arch["in_channels"],
                  start_channels=arch["start_channels"], stage_kwargs=None, max_channels=arch["max_channels"])
    dec = UFPNModular(conv=conv_b, conv_kernels=arch["conv_kernels"], strides=enc.get_strides(),
                      in_channels=enc.get_channels(), decoder_levels=arch["decoder_levels"],
                      fixed_out_channels=arch["fpn_channels"], min_out_channels=8, upsampling_mode="transpose",
                      num_lateral=1, norm_lateral=False, activation_lateral=False, num_out=1, norm_out=False,
                      activation_out=False)
    ag = AnchorGenerator3DS(**anc)
    apos = ag.num_anchors_per_location()[0]
    cls = BCECLassifier(conv=conv_h, in_channels=arch["fpn_channels"], internal_channels=arch["head_channels"],
                        num_classes=arch["classifier_classes"], anchors_per_pos=apos,
                        num_levels=len(arch["decoder_levels"]), num_convs=1, norm_channels_per_group=16,
                        norm_affine=True, reduction="mean", loss_weight=1., prior_prob=0.01)
    reg = GIoURegressor(conv=conv_h, in_channels=arch["fpn_channels"], internal_channels=arch["head_channels"],
                        anchors_per_pos=apos, num_levels=len(arch["decoder_levels"]), num_convs=1,
                        norm_channels_per_group=16, norm_affine=True, reduction="sum", loss_weight=1.,
                        learn_scale=True)
    sampler = HardNegativeSamplerBatched(batch_size_per_image=32, positive_fraction=0.33, pool_size=20, min_neg=1)
    head = DetectionHeadHNMNative(classifier=cls, regressor=reg, coder=BoxCoderND(weights=(1.0,) * 6),
                                  sampler=sampler, log_num_anchors=None)
    seg = DiCESegmenterFgBg(conv_b, seg_classes=arch["seg_classes"], in_channels=dec.get_channels(),
                            decoder_levels=arch["decoder_levels"], dice_kwargs={"batch_dice": True})
    matcher = ATSSMatcher(similarity_fn=rops.box_iou, num_candidates=4, center_in_gt=False)
    return BaseRetinaNet(dim=3, encoder=enc, decoder=dec, head=head, anchor_generator=ag, matcher=matcher,
                         num_classes=arch["classifier_classes"], decoder_levels=arch["decoder_levels"],
                         segmenter=seg, detections_per_img=100, score_thresh=0, topk_candidates=10000,
                         remove_small_boxes=0.01, nms_thresh=0.6)


class _HashSamplerMixin:
    """Test double for the RNG only: replaces torch.randperm draws (sampler.py:93,205) with the
    counter-hash priorities of oracle.box_oracle so reference and CUDA path pick the same anchors."""
    seed = 0

    def select_positives(self, positive, num_pos, img_labels, img_fg_probs):
        pr = bo.hash_priority(positive.numpy(), self.seed, 1).astype(np.uint64) * (1 << 32) + positive.numpy().astype(np.uint64)
        sel = positive[torch.from_numpy(np.argsort(pr, kind="stable")[:num_pos])]
        m = torch.zeros_like(img_labels, dtype=torch.uint8); m[sel] = 1
        return m

    def select_negatives(self, negative, num_neg, img_labels, img_fg_probs):
        pool = min(negative.numel(), int(num_neg * self.pool_size))
        _, ip = img_fg_probs[negative].topk(pool, sorted=True)
        negative = torch.sort(negative[ip])[0]
        pr = bo.hash_priority(negative.numpy(), self.seed, 2).astype(np.uint64) * (1 << 32) + negative.numpy().astype(np.uint64)
        sel = negative[torch.from_numpy(np.argsort(pr, kind="stable")[:num_neg])]
        m = torch.zeros_like(img_labels, dtype=torch.uint8); m[sel] = 1
        return m


class HashSampler(_HashSamplerMixin, HardNegativeSamplerBatched):
    pass


def gen_model(name="tiny", seed=0):
    arch, anc, patch, bs = mo.make_plan(name)
    torch.manual_seed(0)
    ref = build_reference_model(dict(arch), dict(anc))
    orc = mo.RetinaUNetOracle(dict(arch), dict(anc))
    rk, ok = list(ref.state_dict().keys()), list(orc.state_dict().keys())
    assert sorted(rk) == sorted(ok), f"state_dict keys differ:\n{set(rk) ^ set(ok)}"
    for k in rk:
        assert ref.state_dict()[k].shape == orc.state_dict()[k].shape, k
    sd = det_fill(ref.state_dict(), seed)
    ref.load_state_dict(sd); orc.load_state_dict(sd)
    hs = HashSampler(batch_size_per_image=32, positive_fraction=0.33, pool_size=20, min_neg=1)
    hs.seed = 123
    ref.head.fg_bg_sampler = hs

    images, targets = mo.synth_batch(patch, bs, arch["in_channels"], arch["classifier_classes"], 2024 + seed)
    t_ref = {k: ([x.clone() for x in v] if isinstance(v, list) else v.clone()) for k, v in targets.items()}
    ref.train(); orc.train()
    losses_r, pred_r = ref.train_step(images, t_ref, evaluation=True, batch_num=0)
    sum(losses_r.values()).backward()
    losses_o, aux = orc.train_step(images, targets, seed=123)
    sum(losses_o.values()).backward()
    for k in losses_r:
        assert torch.allclose(losses_r[k], losses_o[k], rtol=1e-5, atol=1e-6), (k, losses_r[k], losses_o[k])
    gr = {k: p.grad for k, p in ref.named_parameters()}
    go = {k: p.grad for k, p in orc.named_parameters()}
    for k in gr:
        assert torch.allclose(gr[k], go[k], rtol=1e-3, atol=1e-5), f"grad {k}: {(gr[k] - go[k]).abs().max()}"
    post_o = orc.postprocess(images, {k: v.detach() for k, v in aux["pred"].items()}, aux["anchors"])
    for i in range(bs):
        assert torch.equal(pred_r["pred_labels"][i], post_o[i][2]), "postprocess labels differ"
        assert torch.allclose(pred_r["pred_boxes"][i], post_o[i][0], atol=1e-4), "postprocess boxes differ"
        assert torch.allclose(pred_r["pred_scores"][i], post_o[i][1], atol=1e-6)

    with torch.no_grad():
        pd, _, ps = ref(images)
    arrs = dict(seed=seed, sampler_seed=123, images_crc=zlib.crc32(images.numpy().tobytes()),
                box_logits=pd["box_logits"], box_deltas=pd["box_deltas"][::7],
                seg_logits=ps["seg_logits"][:, :, ::2, ::2, ::2],
                pos_idx=aux["pos"], neg_idx=aux["neg"], labels_nonzero_idx=torch.where(aux["labels"] != 0)[0],
                labels_nonzero=aux["labels"][aux["labels"] != 0])
    for k, v in losses_r.items():
        arrs["loss_" + k] = v.detach()
    for k, gv in gr.items():
        arrs["gnorm/" + k] = gv.double().norm()
        arrs["ghead/" + k] = gv.reshape(-1)[:16]
    for i in range(bs):
        arrs[f"det_boxes{i}"] = pred_r["pred_boxes"][i]
        arrs[f"det_scores{i}"] = pred_r["pred_scores"][i]
        arrs[f"det_labels{i}"] = pred_r["pred_labels"][i]
    save(f"model_{name}", **arrs)


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["pairwise", "anchors", "atss", "sampler", "coder", "nms", "wbc", "transforms", "ensembler", "predictor", "helper", "sweeper", "restore", "lr", "model"]
    for w in which:
        print("==", w)
        globals()["gen_" + w]()
    print("all reference/oracle agreement assertions passed")
