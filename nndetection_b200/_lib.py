"""ctypes binding of libnndet_b200.so -- the C-ABI boundary of the package.

The product path has NO CPU fallback: if the shared library is missing or was not built for sm_100a the import
of any op raises.  PyTorch is used only for device memory (torch's caching allocator, so the planner's VRAM
estimator sees every byte: nndet/planning/estimator.py:228,240), streams and torch.distributed.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnndet_b200.so")

_lib = None

c_void_p, c_int, c_ll, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_size_t


class NndError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NndError(f"{LIB_PATH} not found: build it with `python -m nndetection_b200.build` "
                           f"(nndetection_b200 has no CPU/PyTorch fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.nnd_last_error.restype = ctypes.c_char_p
        _lib.nnd_build_arch.restype = ctypes.c_char_p
        _lib.nnd_nms_workspace_bytes.restype = c_size_t
        _lib.nnd_nms_workspace_bytes.argtypes = [c_ll, c_int]
    return _lib


def check(status: int, what: str):
    if status != 0:
        msg = lib().nnd_last_error().decode() if status == 3 else {1: "bad argument", 2: "workspace too small"}.get(status, "?")
        raise NndError(f"{what} failed with status {status}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NndError("nndetection_b200 ops need CUDA tensors (no CPU fallback)")
