"""Sliding-window case prediction with mirror TTA, kept on the device: host mirror of the reference's predictor loop
(nndet/inference/predictor.py:155-306: `predict_case` -> `tile_case` -> `predict_tiles` -> `predict_with_transformation`), its tile
grid (nndet/io/patching.py:157-343: `create_grid`, `_fixed_slices`, `_symmetric_slices`, `save_get_crop(mode="shift")` :345-392) and
its TTA transforms (nndet/inference/transforms.py:25-72 `get_tta_transforms`; nndet/io/transforms/spatial.py:24-240 `Mirror`).

What changed underneath: the case volume is uploaded once and tiles are views of it (the reference collates numpy crops through a
DataLoader and uploads every batch per TTA pass, predictor.py:253-299); predictions never leave the device on their way into the
ensembler (inference/ensembler.py); the model is not shuttled CPU <-> GPU per case (predictor.py:255,275-276).  Tiles of one case
are independent, so with `shard=(rank, world)` every rank predicts tiles rank::world and `gather_case_result` merges the per-tile
detections on rank 0 before the whole-case NMS / WBC (the only exchange step of the inference path).
Segmentation ensembling (`pred_seg`) is not part of this row.
"""
import itertools
from typing import Any, Callable, Dict, Hashable, List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor


# ------------------------------------------------------------------ tile grid (nndet/io/patching.py)
def _fixed_slices(psize: int, dlim: int, overlap: int, start: int = 0) -> Tuple[slice, ...]:
    """patching.py:258-284: patches of size psize, consecutive ones overlap by `overlap`; only the last one may exceed dlim."""
    upper_limit, lower_limit, idx, crops = 0, start, 0, []
    while upper_limit < dlim:
        if idx != 0:
            lower_limit = lower_limit - overlap
        upper_limit = lower_limit + psize
        crops.append(slice(lower_limit, upper_limit))
        lower_limit = upper_limit
        idx += 1
    return tuple(crops)


def _symmetric_slices(psize: int, dlim: int, overlap: int) -> Tuple[slice, ...]:
    """patching.py:287-306: first and last patch stick out of the data by the same amount."""
    if psize >= dlim:
        return _fixed_slices(psize, dlim, overlap, start=-(psize - dlim) // 2)
    pmod = dlim % (psize - overlap)
    return _fixed_slices(psize, dlim, overlap, start=(pmod - psize) // 2)


def create_grid(cshape: Union[Sequence[int], int], dshape: Sequence[int], overlap: Union[Sequence[int], int] = 0, mode: str = "fixed",
                center_boarder: bool = False, **kwargs) -> List[Tuple[slice, ...]]:
    """patching.py:157-255 for matching dimensionality of patch and data (the 3-D path of the predictor)."""
    fns = {"fixed": _fixed_slices, "symmetric": _symmetric_slices}
    if isinstance(cshape, int):
        cshape = tuple([cshape] * len(dshape))
    if isinstance(overlap, int):
        overlap = tuple([overlap] * len(dshape))
    if len(cshape) != len(dshape):
        raise TypeError("cshape and dshape must be defined for same dimensionality.")
    if len(overlap) != len(dshape):
        raise TypeError("overlap and dshape must be defined for same dimensionality.")
    if any(c - o < 0 for c, o in zip(cshape, overlap)):
        raise TypeError("Overlap must be smaller than size of patches.")
    grid_slices = [fns[mode](p, d, o, **kwargs) for p, d, o in zip(cshape, dshape, overlap)]
    if center_boarder:
        for idx, (p, d, _) in enumerate(zip(cshape, dshape, overlap)):
            lo, hi = int(-0.5 * p), d - int(0.5 * p)
            grid_slices[idx] = (slice(lo, lo + p), *grid_slices[idx], slice(hi, hi + p))
    return list(itertools.product(*grid_slices))


def shifted_crop(dshape: Sequence[int], crop: Sequence[slice]) -> Tuple[List[int], List[slice]]:
    """`_shifted_crop`, patching.py:345-392 (`save_get_crop(mode="shift")`): crops sticking out of the data are moved inside.
    dshape: the spatial shape the crop refers to.  Returns (origin, shifted crop)."""
    out = []
    for dim, c in zip(dshape, crop):
        if c.start < 0:
            new = slice(0, c.stop - c.start, c.step)
            if new.stop > dim:
                raise RuntimeError("Patch is bigger than entire data. shift is not supported in this case.")
            out.append(new)
        elif c.stop > dim:
            new = slice(c.start - (c.stop - dim), dim, c.step)
            if new.start < 0:
                raise RuntimeError("Patch is bigger than entire data. shift is not supported in this case.")
            out.append(new)
        else:
            out.append(c)
    return [int(s.start) for s in out], out


def padded_crop_symmetric(data: Tensor, crop: Sequence[slice]) -> Tuple[Tensor, List[int], List[slice]]:
    """`_padded_crop(mode="symmetric")`, patching.py:396-452 = `np.pad(data[clipped crop], mode="symmetric")`, as index gathers on
    the device: position p of an axis clipped to [lb, ub) (n = ub - lb) reads lb + m if m < n else lb + 2n - 1 - m with
    m = (p - lb) mod 2n (edge value repeated, reflections continue periodically).  Returns (tile, origin = crop starts, crop)."""
    out = data
    nd = data.dim() - len(crop)
    for ax, (c, dim) in enumerate(zip(crop, data.shape[nd:])):
        lb, ub = max(c.start, 0), min(c.stop, dim)
        n = ub - lb
        if n <= 0:
            raise RuntimeError("crop lies completely outside of the data")
        m = (torch.arange(c.start, c.stop, device=data.device) - lb) % (2 * n)
        idx = lb + torch.where(m < n, m, 2 * n - 1 - m)
        out = out.index_select(nd + ax, idx)
    return out, [int(c.start) for c in crop], list(crop)


# ------------------------------------------------------------------ mirror TTA (nndet/io/transforms/spatial.py, inference/transforms.py)
def mirror(data: Tensor, dims: Sequence[int]) -> Tensor:
    """spatial.py:87-99: flip the given spatial dims of [N, C, spatial...]."""
    return data.flip([d + 2 for d in dims])


def mirror_boxes(boxes: Tensor, dims: Sequence[int], shape: Sequence[int]) -> Tensor:
    """`Mirror` on a `box_keys` entry (spatial.py:64-67 via boxes2points -> mirror_points -> points2boxes, :102-232): a point p of a
    mirrored axis becomes shape - p, so (lo, hi) -> (shape - hi, shape - lo).  Box layout (x1, y1, x2, y2, z1, z2): axis 0 -> columns
    (0, 2), axis 1 -> (1, 3), axis 2 -> (4, 5)."""
    if boxes.numel() == 0:
        return boxes.new_zeros((0, 6))
    out = boxes.clone()
    for d in dims:
        lo_c, hi_c = ((0, 2), (1, 3), (4, 5))[d]
        s = float(shape[d])
        out[:, lo_c] = torch.minimum(s - boxes[:, lo_c], s - boxes[:, hi_c])
        out[:, hi_c] = torch.maximum(s - boxes[:, lo_c], s - boxes[:, hi_c])
    return out


def get_tta_dims(num_tta_transforms: int) -> List[Tuple[int, ...]]:
    """Mirror dims of `get_tta_transforms` (inference/transforms.py:25-72) in its order; () = NoOp.  0: no TTA, 4: the 2-D
    mirrors, 8: all 3-D mirrors."""
    dims: List[Tuple[int, ...]] = [()]
    if num_tta_transforms >= 4:
        dims += [(0,), (1,), (0, 1)]
    if num_tta_transforms >= 8:
        dims += [(2,), (0, 2), (1, 2), (0, 1, 2)]
    return dims


# ------------------------------------------------------------------ predictor loop (nndet/inference/predictor.py)
class SlidingWindowPredictor:
    """`Predictor` for the detection branch: `predict_case(case, properties)` -> {"boxes": ensembler result dict}.

    ensembler_fn: callable(case, properties=...) -> ensembler with add_model / process_batch / get_case_result
        (e.g. `lambda case, properties: BoxEnsemblerSelective.from_case(case, properties, parameters)`)
    models: modules with `inference_step(images) -> {"pred_boxes": [...], "pred_scores": [...], "pred_labels": [...]}`
    """

    def __init__(self, ensembler_fn: Callable, models: Sequence[Any], crop_size: Sequence[int], overlap: float = 0.5,
                 num_tta_transforms: int = 8, batch_size: int = 4, model_weights: Optional[Sequence[float]] = None,
                 device: Union[str, torch.device] = "cuda:0", shard: Tuple[int, int] = (0, 1), data_key: str = "data"):
        self.ensembler_fn, self.models = ensembler_fn, list(models)
        self.model_weights = [1.] * len(self.models) if model_weights is None else list(model_weights)
        self.crop_size, self.overlap, self.batch_size = tuple(int(c) for c in crop_size), overlap, batch_size
        self.tta_dims = get_tta_dims(num_tta_transforms)
        self.device, self.shard, self.data_key = torch.device(device), shard, data_key
        self.grid_mode, self.save_get_mode = "symmetric", "shift"          # predictor.py:120-121
        self.ensembler = None

    def tile_case(self, case: Dict) -> List[Dict]:
        """predictor.py:192-235: tiles as VIEWS of the (device-resident) case + their origin / crop.  When the patch is larger than
        the case in some axis the crop is taken with symmetric padding instead (the reference's fallback, :223-228): a copy, origin =
        the (possibly negative) crop start."""
        data = case[self.data_key]
        dshape = tuple(data.shape[1:])
        overlap = [int(c * self.overlap) for c in self.crop_size]
        tiles = []
        for crop in create_grid(cshape=self.crop_size, dshape=dshape, overlap=overlap, mode=self.grid_mode):
            try:
                origin, sc = shifted_crop(dshape, crop)
                tile = data[(slice(None), *sc)]
            except RuntimeError:
                tile, origin, sc = padded_crop_symmetric(data, crop)
            tiles.append({self.data_key: tile, "tile_origin": origin, "crop": sc})
        return tiles

    @torch.no_grad()
    def predict_tiles(self, tiles: Sequence[Dict]) -> None:
        """predictor.py:237-306: per model, per TTA transform (each one registered as its own ensembler "model"), per tile batch."""
        rank, world = self.shard
        mine = list(tiles)[rank::world]
        for model_idx, (model, model_weight) in enumerate(zip(self.models, self.model_weights)):
            if hasattr(model, "eval"):
                model.eval()
            for t, dims in enumerate(self.tta_dims):
                self.ensembler.add_model(name=f"model{model_idx}_t{t}", model_weight=model_weight)
                for i in range(0, len(mine), self.batch_size):
                    chunk = mine[i:i + self.batch_size]
                    data = torch.stack([c[self.data_key] for c in chunk]).to(self.device)
                    batch = {self.data_key: data, "tile_origin": [torch.tensor([c["tile_origin"][ax] for c in chunk]) for ax in range(3)]}
                    result = model.inference_step(mirror(data, dims) if dims else data)
                    if dims:
                        shape = tuple(data.shape[2:])
                        result = dict(result)
                        result["pred_boxes"] = [mirror_boxes(b, dims, shape) for b in result["pred_boxes"]]
                    self.ensembler.process_batch(result=result, batch=batch)

    def gather_case_result(self) -> None:
        """Tile sharding: move every rank's per-tile detections to rank 0's ensembler (variable-length lists, a few KB)."""
        import torch.distributed as dist
        rank, world = self.shard
        if world == 1:
            return
        payload = {name: {k: [t.cpu() for t in v] for k, v in res.items()} for name, res in self.ensembler.model_results.items()}
        gathered = [None] * world
        dist.all_gather_object(gathered, payload)
        if rank == 0:
            dev = self.device
            for r in range(1, world):
                for name, res in gathered[r].items():
                    for k, v in res.items():
                        self.ensembler.model_results[name][k].extend(t.to(dev) for t in v)

    def predict_case(self, case: Dict, properties: Optional[Dict] = None, save_dir=None, case_id: Optional[str] = None,
                     restore: bool = False) -> Optional[Dict[Hashable, Dict]]:
        """predictor.py:155-190 (detection branch), same arguments: with `save_dir` the ensembler state (`<case_id>_boxes.pt`) and the
        properties (`<case_id>_properties.pkl`) are saved next to returning the result.  With tile sharding only rank 0 saves and
        returns (others None)."""
        case = dict(case)
        if isinstance(case[self.data_key], Tensor):
            case[self.data_key] = case[self.data_key].to(self.device)
        else:
            case[self.data_key] = torch.as_tensor(case[self.data_key]).to(self.device)
        self.ensembler = self.ensembler_fn(case, properties=properties)
        self.predict_tiles(self.tile_case(case))
        self.gather_case_result()
        if self.shard[0] != 0:
            return None
        result = {"boxes": self.ensembler.get_case_result(restore=restore)}
        if save_dir is not None:                                   # predictor.py:180-185
            import pickle
            from pathlib import Path
            save_dir = Path(save_dir)
            save_dir.mkdir(parents=True, exist_ok=True)
            self.ensembler.save_state(save_dir, name=case_id)
            with open(save_dir / f"{case_id}_properties.pkl", "wb") as f:
                pickle.dump(properties, f)
        return result
