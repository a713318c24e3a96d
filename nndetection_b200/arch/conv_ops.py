"""Host side of the convolution family: geometry tables for the gather-convolution kernels and thin launch
wrappers over the C ABI (nnd_conv_gather_bf16 / nnd_conv_wgrad_bf16 / nnd_conv_first_* / nnd_norm_* ...).

Tensors are torch tensors of logical shape [N, C, D, H, W] in channels_last_3d memory format (= NDHWC in HBM),
bf16 for activations.  See nndetection_b200/csrc/conv_common.cuh for the kernel-side description.
"""
import itertools
from ctypes import c_float, c_int, c_longlong, byref
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from .. import _lib as L

CL3D = torch.channels_last_3d


def pad32(c: int) -> int:
    return (c + 31) // 32 * 32


def t3(v) -> Tuple[int, int, int]:
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return tuple(int(x) for x in v)
    return (int(v),) * 3


def empty_cl(n: int, c: int, sp: Sequence[int], dtype=torch.bfloat16, device="cuda") -> Tensor:
    """[N, C, D, H, W] tensor with NDHWC memory."""
    return torch.empty((n, sp[0], sp[1], sp[2], c), dtype=dtype, device=device).permute(0, 4, 1, 2, 3)


def zeros_cl(n: int, c: int, sp: Sequence[int], dtype=torch.bfloat16, device="cuda") -> Tensor:
    return torch.zeros((n, sp[0], sp[1], sp[2], c), dtype=dtype, device=device).permute(0, 4, 1, 2, 3)


def as_cl(x: Tensor, dtype=torch.bfloat16) -> Tensor:
    """Make x [N,C,D,H,W] dense NDHWC of `dtype` (no copy when it already is)."""
    if x.dtype != dtype:
        x = x.to(dtype)
    if x.shape[1] == 1 and x.is_contiguous():
        return x                                   # one channel: NCDHW memory == NDHWC memory
    return x.contiguous(memory_format=CL3D)


def _geom(N, in_sp, cin, L_, s, out_sp, om, oo, taps):
    """taps: list of (od, oh, ow, weight_tap)."""
    a = [N, in_sp[0], in_sp[1], in_sp[2], cin, L_[0], L_[1], L_[2], s[0], s[1], s[2], out_sp[0], out_sp[1], out_sp[2],
         om[0], om[1], om[2], oo[0], oo[1], oo[2], len(taps)]
    for t in taps:
        a.extend(t)
    return (c_int * len(a))(*a)


class ConvPlan:
    """All launch geometries of one conv layer for one input shape (cached by the layer)."""

    def __init__(self, N: int, cin: int, cdy: int, in_sp: Sequence[int], k, s, p, transposed: bool):
        """cin: channels of the layer input tensor; cdy: channels of the (padded) output-gradient tensor."""
        k, s, p = t3(k), t3(s), t3(p)
        self.k, self.s, self.p, self.transposed = k, s, p, transposed
        self.T = k[0] * k[1] * k[2]
        self.in_sp = tuple(int(v) for v in in_sp)
        flat = lambda a, b, c: (a * k[1] + b) * k[2] + c
        if not transposed:
            self.out_sp = tuple((i + 2 * pp - kk) // ss + 1 for i, pp, kk, ss in zip(in_sp, p, k, s))
            taps = [(a - p[0], b - p[1], c - p[2], flat(a, b, c)) for a, b, c in itertools.product(*[range(x) for x in k])]
            self.fprop = [_geom(N, in_sp, cin, self.out_sp, s, self.out_sp, (1, 1, 1), (0, 0, 0), taps)]
            # dgrad: dy -> dx, one launch per output parity class
            self.dgrad, self.dgrad_covers_all = [], True
            per_axis = []
            for ax in range(3):
                cls = []
                for par in range(s[ax]):
                    tl = [((par + p[ax] - kk) // s[ax], kk) for kk in range(k[ax]) if (par + p[ax] - kk) % s[ax] == 0]
                    cls.append((par, tl))
                per_axis.append(cls)
            for (pd, td), (ph, th), (pw, tw) in itertools.product(*per_axis):
                Lc = tuple(-(-(in_sp[ax] - par) // s[ax]) for ax, par in enumerate((pd, ph, pw)))
                if min(Lc) <= 0:
                    continue
                if not (td and th and tw):
                    self.dgrad_covers_all = False
                    continue
                tp = [(a[0], b[0], c[0], flat(a[1], b[1], c[1])) for a, b, c in itertools.product(td, th, tw)]
                self.dgrad.append(_geom(N, self.out_sp, cdy, Lc, (1, 1, 1), in_sp, s, (pd, ph, pw), tp))
            self.wgrad = self.fprop
        else:
            assert k == s and p == (0, 0, 0), "transposed convolutions are supported with kernel == stride, padding 0"
            self.out_sp = tuple(i * ss for i, ss in zip(in_sp, s))
            self.fprop = []
            for a, b, c in itertools.product(*[range(x) for x in s]):
                self.fprop.append(_geom(N, in_sp, cin, in_sp, (1, 1, 1), self.out_sp, s, (a, b, c), [(0, 0, 0, flat(a, b, c))]))
            taps = [(a, b, c, flat(a, b, c)) for a, b, c in itertools.product(*[range(x) for x in s])]
            self.dgrad = [_geom(N, self.out_sp, cdy, in_sp, s, in_sp, (1, 1, 1), (0, 0, 0), taps)]
            self.dgrad_covers_all = True
            self.wgrad = self.fprop
            # the same weight gradient as ONE strided gather with the operands' roles swapped: dense operand = the layer input,
            # strided operand = dy read at s * i + tap (used by the opt-in tcgen05 path, arch/conv.py)
            self.wgrad_swapped = self.dgrad[0]



def conv_gather(x: Tensor, w_packed: Tensor, geom, out: Tensor, cout: int, cout_pad: int, *, out_fp32=False,
                out_n_stride=None, out_v_stride=None, bias=None, scale=None, residual=None, stat_sum=None, stat_sq=None,
                w_items=None):
    """One launch of the gather convolution.  `out` may be any tensor whose data_ptr is the destination base.
    w_items: optional `ItemPack` of the same weights (opt-in bulk-copy variant of the tile kernel)."""
    used = c_int(0)
    if out_n_stride is None:
        out_n_stride = geom[11] * geom[12] * geom[13] * cout
    if out_v_stride is None:
        out_v_stride = cout
    if w_items is None:
        L.check(L.lib().nnd_conv_gather_bf16(L.ptr(x), L.ptr(w_packed), geom, L.ptr(out), c_longlong(out_n_stride),
                                             c_longlong(out_v_stride), c_int(1 if out_fp32 else 0), c_int(cout),
                                             c_int(cout_pad), L.ptr(bias), L.ptr(scale), L.ptr(residual), L.ptr(stat_sum),
                                             L.ptr(stat_sq), byref(used), L.stream_ptr()), "nnd_conv_gather_bf16")
    else:
        L.check(L.lib().nnd_conv_gather_bf16_items(L.ptr(x), L.ptr(w_packed), geom, L.ptr(out), c_longlong(out_n_stride),
                                                   c_longlong(out_v_stride), c_int(1 if out_fp32 else 0), c_int(cout),
                                                   c_int(cout_pad), L.ptr(bias), L.ptr(scale), L.ptr(residual), L.ptr(stat_sum),
                                                   L.ptr(stat_sq), byref(used), L.stream_ptr(), L.ptr(w_items.data),
                                                   c_int(w_items.n_tile), c_int(w_items.T)), "nnd_conv_gather_bf16_items")
    return used.value


class ItemPack:
    """A K-major weight pack [T][rows_pad][K] re-ordered for the bulk-copy variant of the tile kernel: [row tile][k chunk][T][k group]
    [n][8] -- the slice one pipeline item needs (one tap, one 32-channel chunk, one n_tile-row tile) is contiguous."""

    def __init__(self, data: Tensor, n_tile: int, T: int):
        self.data, self.n_tile, self.T = data, n_tile, T


def item_n_tile(rows_pad: int) -> int:
    """The row tile the tile kernel uses for `rows_pad` output channels (csrc/conv_tc.cu: nnd_conv_tc)."""
    return 128 if rows_pad % 128 == 0 else (64 if rows_pad % 64 == 0 else 32)


def repack_items(w_packed: Tensor) -> Optional["ItemPack"]:
    """w_packed: bf16 [T][rows_pad][K] (fprop or dgrad pack of `pack_weights`).  None when the shape cannot take the bulk path."""
    T, rows_pad, K = (int(v) for v in w_packed.shape)
    if K % 32 or rows_pad % 32:
        return None
    n_tile = item_n_tile(rows_pad)
    out = torch.empty_like(w_packed)
    L.check(L.lib().nnd_repack_items_bf16(L.ptr(w_packed), c_int(T), c_int(rows_pad), c_int(K), c_int(n_tile), c_int(4), L.ptr(out),
                                          L.stream_ptr()), "nnd_repack_items_bf16")
    return ItemPack(out, n_tile, T)


_TC_BULK = False


def set_tc_bulk(enable: bool):
    """Do not enable: this variant of the tile kernel (weights via cp.async.bulk on item-order packs) deadlocks on the device."""
    global _TC_BULK
    L.lib().nnd_conv_set_tc_bulk(c_int(1 if enable else 0))
    _TC_BULK = bool(enable)


def tc_bulk_enabled() -> bool:
    return _TC_BULK


_WGRAD_WS = {}          # (device index, stream handle) -> uint8 workspace tensor, grown on demand


def _wgrad_workspace(nbytes: int, device) -> Tensor:
    """Split-K partials of the TMA-fed wgrad kernels: ONE buffer per (device, launch stream), grown to the largest request and kept --
    consecutive launches of a stream reuse it in stream order (the finishing pass of a launch runs before the next launch's kernel), and
    the ~40 allocations / frees per step it replaces no longer churn the caching allocator (a first-time cudaMalloc inside a step costs
    ~10 ms).  Still allocated through torch: the planner's VRAM estimator sees the bytes."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    ws = _WGRAD_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WGRAD_WS[key] = ws
    return ws


def conv_wgrad(dy: Tensor, cdy: int, x: Tensor, cx: int, geom, dw: Tensor, s_co: int, s_ci: int, s_tap: int, cout: int,
               cin: int):
    lib = L.lib()
    lib.nnd_conv_wgrad_workspace_bytes.restype = c_longlong
    nbytes = int(lib.nnd_conv_wgrad_workspace_bytes(geom, c_int(cdy), c_int(cx), c_int(cout), c_int(cin)))
    ws = _wgrad_workspace(nbytes, dy.device) if nbytes > 0 else None
    L.check(lib.nnd_conv_wgrad_bf16_ws(L.ptr(dy), c_int(cdy), L.ptr(x), c_int(cx), geom, L.ptr(dw), c_longlong(s_co),
                                       c_longlong(s_ci), c_longlong(s_tap), c_int(cout), c_int(cin), L.ptr(ws), c_longlong(nbytes),
                                       L.stream_ptr()), "nnd_conv_wgrad_bf16_ws")


def conv_first_fprop(x: Tensor, w: Tensor, geom, cout: int, out: Tensor, stat_sum, stat_sq):
    L.check(L.lib().nnd_conv_first_fprop_f32(L.ptr(x), L.ptr(w), geom, c_int(cout), L.ptr(out), L.ptr(stat_sum),
                                             L.ptr(stat_sq), L.stream_ptr()), "nnd_conv_first_fprop_f32")


def conv_first_wgrad(x: Tensor, dy: Tensor, geom, cout: int, dw: Tensor):
    L.check(L.lib().nnd_conv_first_wgrad_f32(L.ptr(x), L.ptr(dy), geom, c_int(cout), L.ptr(dw), L.stream_ptr()),
            "nnd_conv_first_wgrad_f32")


def pack_weights(w: Tensor, cout: int, cin: int, T: int, transposed: bool, want_bwd: bool = True):
    """fp32 master weight -> (fwd bf16 [T][pad32(cout)][pad32(cin)], bwd bf16 [T][pad32(cin)][pad32(cout)])."""
    cop, cip = pad32(cout), pad32(cin)
    fwd = torch.empty((T, cop, cip), dtype=torch.bfloat16, device=w.device)
    bwd = torch.empty((T, cip, cop), dtype=torch.bfloat16, device=w.device) if want_bwd else None
    L.check(L.lib().nnd_pack_weights(L.ptr(w), c_int(cout), c_int(cin), c_int(T), c_int(1 if transposed else 0), L.ptr(fwd),
                                     c_int(cop), c_int(cip), L.ptr(bwd), c_int(cip), c_int(cop), L.stream_ptr()),
            "nnd_pack_weights")
    return fwd, bwd


def norm_finalize(ssum, ssq, gamma, beta, N, C, cpg, count, eps):
    dev = ssum.device
    buf = torch.empty((4, N, C), dtype=torch.float32, device=dev)
    a, b, mean, rstd = buf[0], buf[1], buf[2], buf[3]
    L.check(L.lib().nnd_norm_finalize(L.ptr(ssum), L.ptr(ssq), L.ptr(gamma), L.ptr(beta), c_int(N), c_int(C), c_int(cpg),
                                      c_longlong(count), c_float(eps), L.ptr(a), L.ptr(b), L.ptr(mean), L.ptr(rstd),
                                      L.stream_ptr()), "nnd_norm_finalize")
    return a, b, mean, rstd


def norm_apply(y: Tensor, a, b, N, V, C, relu: bool, z: Tensor):
    L.check(L.lib().nnd_norm_apply(L.ptr(y), L.ptr(a), L.ptr(b), c_int(N), c_longlong(V), c_int(C), c_int(1 if relu else 0),
                                   L.ptr(z), L.stream_ptr()), "nnd_norm_apply")


def norm_backward(dz, y, a, b, mean, rstd, gamma, N, V, C, cpg, relu, dy, dgamma, dbeta):
    ws = torch.empty(5 * N * C, dtype=torch.float32, device=dz.device)
    L.check(L.lib().nnd_norm_backward(L.ptr(dz), L.ptr(y), L.ptr(a), L.ptr(b), L.ptr(mean), L.ptr(rstd), L.ptr(gamma), c_int(N),
                                      c_longlong(V), c_int(C), c_int(cpg), c_int(1 if relu else 0), L.ptr(dy), L.ptr(dgamma),
                                      L.ptr(dbeta), L.ptr(ws), L.stream_ptr()), "nnd_norm_backward")


def channel_sum(src: Tensor, rows: int, C: int, stride: int, out: Tensor, scale: float = 1.0):
    L.check(L.lib().nnd_channel_sum(L.ptr(src), c_int(1 if src.dtype == torch.bfloat16 else 0), c_longlong(rows), c_int(C),
                                    c_longlong(stride), c_float(scale), L.ptr(out), L.stream_ptr()), "nnd_channel_sum")


def set_tensor_path(enable_tcgen05: bool):
    L.lib().nnd_conv_set_tensor_path(c_int(1 if enable_tcgen05 else 0))


def set_stream_path(mode: int = 1, issuers: int = 2):
    """0: off, 1: streaming z-window tcgen05 kernel where profitable (default), 2: wherever the shape is supported."""
    L.lib().nnd_conv_set_stream_path(c_int(mode), c_int(issuers))



_WGRAD_STRIDED_TC = True        # default since round 2 (validated on a B200); the C side has the same default


def set_wgrad_strided_tc(enable: bool):
    """De-interleaved tcgen05 weight gradient for stride-2 convolutions and kernel == stride transposed convolutions (default on;
    False = the mma.sync kernels, for A/B)."""
    global _WGRAD_STRIDED_TC
    L.lib().nnd_conv_set_wgrad_strided_tc(c_int(1 if enable else 0))
    _WGRAD_STRIDED_TC = bool(enable)


def set_gather_strided_tc(enable: bool):
    """De-interleaved-halo tcgen05 kernel for stride-2 convolutions and the dgrad of up-convolutions (default on; False = mma.sync)."""
    L.lib().nnd_conv_set_gather_strided_tc(c_int(1 if enable else 0))


GATHER_TMA_DEFAULT = 3      # on since its validation on a B200 (round 2); the C side has the same default


def set_gather_tma(mode: int):
    """TMA-fed variant of the tcgen05 tile kernel (csrc/conv_tct.cu).  bit 0: stride-1 forms (128-channel layers, small volumes, the
    >= 4-tap stride-2 dgrad classes); bit 1: stride-2 convolutions and the dgrad of up-convolutions."""
    L.lib().nnd_conv_set_gather_tma(c_int(int(mode)))


WGRAD_TMA_DEFAULT = 1      # on since its validation on a B200 (round 2); the C side has the same default


def set_wgrad_tma(mode: int):
    """TMA-fed tcgen05 weight gradient (csrc/conv_wgrad_tma.cu) for stride-1 3x3x3 / 1x3x3 layers with channel counts in multiples of
    64.  bit 0: on; bit 1: shared-memory descriptors carry base_offset for row-shifted starts; bit 2: one MMA per dx tap (A/B)."""
    L.lib().nnd_conv_set_wgrad_tma(c_int(int(mode)))


_PW_TMA = True           # default since its validation on a B200 (round 2); the C side has the same default


def set_pointwise_tma(enable: bool):
    """TMA-fed tcgen05 GEMM (csrc/conv_pw.cu) for single-tap gathers -- 1x1x1 convolutions and their input gradients, parity classes of
    up-convolutions -- and, through `conv_upconv`, whole kernel == stride transposed convolutions in one launch."""
    global _PW_TMA
    L.lib().nnd_conv_set_pointwise_tma(c_int(1 if enable else 0))
    _PW_TMA = bool(enable)


def pointwise_tma_enabled() -> bool:
    return _PW_TMA


def conv_upconv(x: Tensor, w_packed: Tensor, N: int, in_sp, cin: int, cout: int, s, out: Tensor, bias=None, residual=None):
    """nn.ConvTranspose3d with kernel == stride in one launch (taps stacked along N, each input voxel read once, lateral added)."""
    L.check(L.lib().nnd_conv_upconv_bf16(L.ptr(x), L.ptr(w_packed), c_int(N), c_int(in_sp[0]), c_int(in_sp[1]), c_int(in_sp[2]), c_int(cin),
                                         c_int(cout), c_int(s[0]), c_int(s[1]), c_int(s[2]), L.ptr(out), L.ptr(bias), L.ptr(residual),
                                         L.stream_ptr()), "nnd_conv_upconv_bf16")


def wgrad_strided_tc_enabled() -> bool:
    return _WGRAD_STRIDED_TC


def set_norm_bwd_narrow(enable: bool):
    """Opt-in: norm backward passes with four channels per thread (more resident warps; not yet measured on a device)."""
    L.lib().nnd_norm_set_bwd_narrow(c_int(1 if enable else 0))


def trace_start():
    """Profiling aid: record one row per convolution-family launch (kernel chosen, layer geometry, duration) until `trace_dump`."""
    L.lib().nnd_conv_trace(c_int(1))


def trace_dump(path: str) -> int:
    """Stop recording, synchronise, write the rows as CSV to `path`; returns the number of rows."""
    lib = L.lib()
    lib.nnd_conv_trace(c_int(0))
    lib.nnd_conv_trace_count.restype = c_longlong
    L.check(lib.nnd_conv_trace_dump(str(path).encode()), "nnd_conv_trace_dump")
    return int(lib.nnd_conv_trace_count())
