"""TMA-fed tcgen05 weight gradient (csrc/conv_wgrad_tma.cu: 5-D tensor-map boxes out of the NDHWC tensors, MN-major SWIZZLE_128B operands,
the three dx taps stacked along N through row-shifted descriptor starts) against the fp32 CPU oracle of the same bf16-exact operands
(torch.nn.grad.conv3d_weight == the autograd of nndet/arch/conv.py:344-348) and against the cp.async kernel it replaces
(conv_wgrad_tc.cu; only the accumulation order differs).  Shapes cover both unit forms (16 w x 4 h / 8 w x 8 h), ragged widths and
heights, one- and two-block co tiles, partial co tiles, several ci tiles, the 4^3 level and a 1x3x3 filter."""
import csv
import os
import tempfile

import pytest
import torch

import tutil as util  # noqa: F401
from test_net_gpu import q, rel_err

pytestmark = pytest.mark.gpu

CASES = [  # cin, cout, (D, H, W), batch, kernel
    (64, 64, (4, 8, 16), 1, 3),
    (128, 128, (5, 6, 32), 2, 3),
    (64, 128, (3, 9, 20), 1, 3),          # wide units, second segment 4 wide, H = 9
    (128, 64, (4, 8, 8), 2, 3),           # narrow units
    (128, 128, (6, 7, 24), 1, 3),         # narrow, three segments per row, odd H
    (256, 320, (4, 4, 8), 1, 3),          # three co tiles (the last one half empty), two ci tiles
    (320, 128, (4, 4, 4), 2, 3),          # five one-block ci tiles, 4^3 level (boxes larger than the tensor)
    (64, 64, (5, 12, 12), 1, (1, 3, 3)),  # 1x3x3 filter: three filter rows
    (128, 192, (2, 5, 40), 1, 3),         # narrow (40 < 48), co tile 2 with one real block
    (192, 64, (3, 7, 20), 2, 3),          # 64 output channels: paired filter rows, odd H (the h = -1 row), three one-block ci tiles
    (64, 64, (2, 8, 8), 2, (1, 3, 3)),    # paired rows, narrow units, 1x3x3
]


def _wgrad(ops, layer, x, dy, cin, cout, sp, bs):
    plan = layer.plan(bs, tuple(sp))
    T = plan.T
    dw = torch.zeros(tuple(layer.conv.weight.shape), dtype=torch.float32, device=x.device)
    ops.trace_start()
    for g in plan.wgrad:
        ops.conv_wgrad(dy, cout, x, cin, g, dw, cin * T, T, 1, cout, cin)
    with tempfile.TemporaryDirectory() as td:
        ops.trace_dump(os.path.join(td, "t.csv"))
        kernels = [r["kernel"] for r in csv.DictReader(open(os.path.join(td, "t.csv"))) if r["kind"] == "wgrad"]
    return dw.cpu(), kernels


@pytest.mark.parametrize("cin,cout,sp,bs,k", CASES)
def test_wgrad_tma_vs_oracle_and_cp_async_kernel(cin, cout, sp, bs, k):
    from nndetection_b200.arch import conv_ops as ops
    from nndetection_b200.arch.conv import ConvInstanceRelu
    pad = 1 if isinstance(k, int) else tuple(v // 2 for v in k)
    layer = ConvInstanceRelu(3, cin, cout, kernel_size=k, stride=1, padding=pad).cuda()
    g = torch.Generator().manual_seed(71 + cin + cout + sp[2])
    x = q(torch.randn(bs, cin, *sp, generator=g))
    dy = q(torch.randn(bs, cout, *sp, generator=g))
    ref = torch.nn.grad.conv3d_weight(x, tuple(layer.conv.weight.shape), dy, stride=1, padding=pad)     # fp32 on the CPU
    xm = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    dym = dy.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    res = {}
    try:
        for mode in (1, 65, 0):     # split-K partials in a workspace + finishing pass (default) | atomics straight into dW | cp.async kernel
            ops.set_wgrad_tma(mode)
            res[mode] = _wgrad(ops, layer, xm, dym, cin, cout, sp, bs)
    finally:
        ops.set_wgrad_tma(ops.WGRAD_TMA_DEFAULT)
    assert res[1][1] == ["wgrad_tma"] and res[65][1] == ["wgrad_tma"] and "wgrad_tma" not in res[0][1]
    assert rel_err(res[65][0], ref) < 1e-4
    assert rel_err(res[1][0], ref) < 1e-4            # fp32 accumulation of exact bf16 products: only the summation order differs
    assert rel_err(res[1][0], res[0][0]) < 1e-4


def test_descriptor_model_base_offset_must_stay_zero():
    """The finding the kernel rests on (first B200 run of scripts/probe_wgrad_tma.py): a SWIZZLE_128B operand may start at ANY 128-byte row
    of a TMA-written box with base_offset = 0 -- the swizzle is a function of the absolute shared-memory address -- while filling
    base_offset with (start >> 7) & 7 shifts the pattern a second time and yields garbage."""
    from nndetection_b200.arch import conv_ops as ops
    from nndetection_b200.arch.conv import ConvInstanceRelu
    layer = ConvInstanceRelu(3, 64, 64, kernel_size=3, stride=1, padding=1).cuda()
    g = torch.Generator().manual_seed(5)
    sp, bs = (4, 8, 16), 1
    x = q(torch.randn(bs, 64, *sp, generator=g))
    dy = q(torch.randn(bs, 64, *sp, generator=g))
    ref = torch.nn.grad.conv3d_weight(x, tuple(layer.conv.weight.shape), dy, stride=1, padding=1)
    xm = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    dym = dy.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    err = {}
    try:
        for mode in (1, 5, 3):          # stacked N = 192 | one N = 64 MMA per tap | stacked with base_offset
            ops.set_wgrad_tma(mode)
            err[mode] = rel_err(_wgrad(ops, layer, xm, dym, 64, 64, sp, bs)[0], ref)
    finally:
        ops.set_wgrad_tma(ops.WGRAD_TMA_DEFAULT)
    assert err[1] < 1e-4 and err[5] < 1e-4 and err[3] > 0.5


def _module_wgrad(ops, mine, x, gy, mode):
    ops.set_wgrad_tma(mode)
    mine.zero_grad(set_to_none=True)
    xm = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    ops.trace_start()
    mine(xm).backward(gy.cuda().to(torch.bfloat16))
    with tempfile.TemporaryDirectory() as td:
        ops.trace_dump(os.path.join(td, "t.csv"))
        kernels = [r["kernel"] for r in csv.DictReader(open(os.path.join(td, "t.csv"))) if r["kind"] == "wgrad"]
    return mine.conv.weight.grad.cpu().clone(), kernels


@pytest.mark.parametrize("cin,cout,s,shape", [
    (32, 64, 2, (2, 12, 16, 40)),           # the first encoder stride: 32-channel x (SWIZZLE_64B planes), paired filter rows +1 / -1
    (64, 128, (1, 2, 2), (2, 6, 12, 12)),   # LIDC-style first stride (unstrided depth), 6 outputs per row: narrow units
    (128, 256, 2, (1, 8, 10, 34)),          # two co tiles, two 64-channel blocks per CTA, odd output width (17)
    (256, 320, 2, (2, 8, 8, 8)),            # three co tiles (the last one half empty), two ci tiles
    (32, 64, 2, (1, 9, 11, 13)),            # odd input sizes in every axis
    (64, 64, 2, (2, 8, 16, 32)),            # 64 -> 64: pair mode with one 64-channel block
])
def test_strided_wgrad_tma_vs_oracle_and_cp_async_kernel(cin, cout, s, shape):
    """conv_wgrad_tma_s2.cu: x through element-stride-2 tensor maps (odd / even planes), dx = -1 / +1 as one MMA.  fp32 dW against the CPU
    oracle on bf16-exact operands and against the cp.async kernel (conv_wgrad_tc.cu, SW = 2: accumulation order only)."""
    from nndetection_b200.arch import conv_ops as ops
    from test_net_gpu import make_pair
    mine, ref = make_pair("instance", cin, cout, 3, s, norm=False)
    g = torch.Generator().manual_seed(91)
    x = q(torch.randn(shape[0], cin, *shape[1:], generator=g))
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = q(torch.randn(yr.shape, generator=g))
    yr.backward(gy)
    try:
        new, k_new = _module_wgrad(ops, mine, x, gy, 1)
        old, k_old = _module_wgrad(ops, mine, x, gy, 1 | 128)
    finally:
        ops.set_wgrad_tma(ops.WGRAD_TMA_DEFAULT)
    assert k_new == ["wgrad_tma_s2"] and k_old == ["wgrad_tc_s2"]
    assert rel_err(new, ref.conv.weight.grad) < 5e-3            # dy passes through one bf16 rounding on the device
    assert rel_err(new, old) < 1e-4


@pytest.mark.parametrize("cin,cout,s,shape", [(64, 32, 2, (2, 4, 6, 40)), (128, 64, 2, (1, 5, 3, 9)), (128, 128, (1, 2, 2), (1, 4, 4, 4)),
                                             (128, 128, 2, (2, 2, 2, 2)), (64, 32, 2, (1, 8, 16, 16))])
def test_transposed_conv_wgrad_tma_vs_oracle_and_cp_async_kernel(cin, cout, s, shape):
    """Up-convolutions (kernel == stride): ONE strided launch with the operands' roles swapped (dense = the layer input, strided = dy at
    2 i + {0, 1}: taps dx in {0, +1})."""
    from nndetection_b200.arch import conv_ops as ops
    from test_net_gpu import make_pair
    mine, ref = make_pair("instance", cin, cout, None, s, transposed=True)
    g = torch.Generator().manual_seed(92)
    x = q(torch.randn(shape[0], cin, *shape[1:], generator=g))
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = q(torch.randn(yr.shape, generator=g))
    yr.backward(gy)
    try:
        new, k_new = _module_wgrad(ops, mine, x, gy, 1)
        old, k_old = _module_wgrad(ops, mine, x, gy, 1 | 128)
    finally:
        ops.set_wgrad_tma(ops.WGRAD_TMA_DEFAULT)
    assert k_new == ["wgrad_tma_s2"] and k_old == ["wgrad_tc_s2"]
    assert rel_err(new, ref.conv.weight.grad) < 5e-3
    assert rel_err(new, old) < 1e-4


@pytest.mark.parametrize("sp,bs,k", [((12, 16, 40), 2, 3), ((5, 9, 20), 1, 3), ((16, 32, 48), 4, 3), ((4, 24, 16), 1, (1, 3, 3))])
def test_32_channel_stacked_tap_wgrad_tma_vs_oracle_and_cp_async_kernel(sp, bs, k):
    """conv_wgrad_tma32.cu (all 27 taps per CTA: dy slices stacked along M, x rows along N through descriptor strides of two SWIZZLE_64B
    tensor-map boxes per tile) against the fp32 CPU oracle and against conv_wgrad_tc32.cu; ragged tiles in every axis, a 1x3x3 filter,
    workspace (split-K partials per CTA) and atomics epilogue."""
    from ctypes import c_int
    from nndetection_b200 import _lib as L
    from nndetection_b200.arch import conv_ops as ops
    from nndetection_b200.arch.conv import ConvInstanceRelu
    pad = 1 if isinstance(k, int) else tuple(v // 2 for v in k)
    layer = ConvInstanceRelu(3, 32, 32, kernel_size=k, stride=1, padding=pad).cuda()
    g = torch.Generator().manual_seed(95 + sp[2])
    x = q(torch.randn(bs, 32, *sp, generator=g))
    dy = q(torch.randn(bs, 32, *sp, generator=g))
    ref = torch.nn.grad.conv3d_weight(x, tuple(layer.conv.weight.shape), dy, stride=1, padding=pad)
    xm = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    dym = dy.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    res = {}
    try:
        L.lib().nnd_conv_set_wgrad_tc(c_int(2))             # the stacked-tap kernels also on volumes too small to fill the grid
        for mode in (1, 1 | 64, 1 | 256):                   # TMA + workspace | TMA + atomics | cp.async kernel
            ops.set_wgrad_tma(mode)
            res[mode] = _wgrad(ops, layer, xm, dym, 32, 32, sp, bs)
    finally:
        ops.set_wgrad_tma(ops.WGRAD_TMA_DEFAULT)
        L.lib().nnd_conv_set_wgrad_tc(c_int(1))
    assert res[1][1] == ["wgrad_tma32"] and res[1 | 64][1] == ["wgrad_tma32"] and res[1 | 256][1] == ["wgrad_tc32"]
    assert rel_err(res[1][0], ref) < 1e-4 and rel_err(res[1 | 64][0], ref) < 1e-4
    assert rel_err(res[1][0], res[1 | 256][0]) < 1e-4


def test_32_channel_wgrad_at_full_resolution_matches_the_cp_async_kernel():
    """The layer the kernel exists for (32 -> 32 @128^3, batch 4: 32 768 tiles over 148 CTAs) -- too large for the CPU oracle, so the A/B
    partner is conv_wgrad_tc32.cu (itself pinned on the oracle at small sizes): same products, different summation order."""
    from nndetection_b200.arch import conv_ops as ops
    from nndetection_b200.arch.conv import ConvInstanceRelu
    layer = ConvInstanceRelu(3, 32, 32, kernel_size=3, stride=1, padding=1).cuda()
    sp, bs = (128, 128, 128), 4
    g = torch.Generator(device="cuda").manual_seed(7)
    xm = torch.randn(bs, 32, *sp, generator=g, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    dym = torch.randn(bs, 32, *sp, generator=g, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    res = {}
    try:
        for mode in (1, 1 | 256):
            ops.set_wgrad_tma(mode)
            res[mode] = _wgrad(ops, layer, xm, dym, 32, 32, sp, bs)
    finally:
        ops.set_wgrad_tma(ops.WGRAD_TMA_DEFAULT)
    assert res[1][1] == ["wgrad_tma32"] and res[1 | 256][1] == ["wgrad_tc32"]
    assert rel_err(res[1][0], res[1 | 256][0]) < 1e-4
