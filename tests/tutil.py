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
