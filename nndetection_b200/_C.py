"""Drop-in for the reference's only native symbol, `nndet._C.nms` (nndet/csrc/ops.cpp:13-15).

    nms(dets: Tensor[N, 6 or 4], scores: Tensor[N], iou_threshold: float) -> Tensor[K] (int64)

Same contract as nms_cuda (nndet/csrc/cuda/nms.cu:148-221): returns indices into `dets` of the kept boxes,
ordered by descending score, on the input device; empty input -> empty int64 tensor
(nndet/csrc/cpu/nms.cpp:24-27); CPU tensors raise RuntimeError (cpu/nms.cpp:31, nms.cu:150-151).
To use it from an unmodified reference checkout: `sys.modules["nndet._C"] = nndetection_b200._C`
before `import nndet`.
"""
import torch

from . import _lib as L


def nms(dets: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    if not dets.is_cuda:
        raise RuntimeError("Not compiled with CPU support")           # cpu/nms.cpp:31
    if not scores.is_cuda:
        raise RuntimeError("scores must be a CUDA tensor")            # nms.cu:151
    n = dets.shape[0]
    if dets.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=dets.device)
    dim = dets.shape[1] // 2
    if dets.shape[1] not in (4, 6):
        raise RuntimeError(f"dets must be [N, 4] or [N, 6], got {tuple(dets.shape)}")
    lib = L.lib()
    with torch.cuda.device(dets.device):
        d = dets.detach().contiguous().float()
        s = scores.detach().contiguous().float()
        ws_bytes = lib.nnd_nms_workspace_bytes(n, dim)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dets.device)
        keep = torch.empty(n, dtype=torch.int64, device=dets.device)
        n_keep = torch.empty(1, dtype=torch.int64, device=dets.device)
        fn = lib.nnd_nms3d_f32 if dim == 3 else lib.nnd_nms2d_f32
        st = fn(L.ptr(d), L.ptr(s), L.c_ll(n), L.c_float(float(iou_threshold)), L.ptr(keep), L.ptr(n_keep),
                L.ptr(ws), L.c_size_t(ws_bytes), L.stream_ptr())
        L.check(st, "nnd_nms")
        k = int(n_keep.item())      # the reference call is host-synchronous too (nms.cu:193)
    return keep[:k]


def nms_device(dets: torch.Tensor, scores: torch.Tensor, iou_threshold: float):
    """Sync-free variant: returns (keep[N] int64 padded, n_keep[1] int64) without reading the count back."""
    L.require_cuda(dets, scores)
    n = dets.shape[0]
    dim = dets.shape[1] // 2
    lib = L.lib()
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=dets.device)
    n_keep = torch.zeros(1, dtype=torch.int64, device=dets.device)
    if n == 0:
        return keep[:0], n_keep
    ws_bytes = lib.nnd_nms_workspace_bytes(n, dim)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dets.device)
    fn = lib.nnd_nms3d_f32 if dim == 3 else lib.nnd_nms2d_f32
    st = fn(L.ptr(dets), L.ptr(scores), L.c_ll(n), L.c_float(float(iou_threshold)), L.ptr(keep), L.ptr(n_keep),
            L.ptr(ws), L.c_size_t(ws_bytes), L.stream_ptr())
    L.check(st, "nnd_nms")
    return keep, n_keep
