"""TMA-fed tcgen05 tile kernel (csrc/conv_tct.cu: halo planes as 5-D tensor-map boxes -- element stride 2 de-interleaves the stride-2 forms --
weights as 2-D boxes, K-major SWIZZLE_64B / SWIZZLE_32B operands with row-shifted descriptor starts) against the CPU oracle on bf16-exact
operands and against the cp.async tile kernel it replaces (conv_tc.cu; same MMAs, same accumulation order per tile)."""
import csv
import os
import tempfile

import pytest
import torch

import tutil as util  # noqa: F401
from test_net_gpu import make_pair, q, rel_err

pytestmark = pytest.mark.gpu


def _run(ops, mine, x, gy, mode, residual=None):
    ops.set_gather_tma(mode)
    mine.zero_grad(set_to_none=True)
    xm = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    ops.trace_start()
    y = mine(xm) if residual is None else mine(xm, residual=residual)
    y.backward(gy.cuda().to(torch.bfloat16))
    with tempfile.TemporaryDirectory() as td:
        ops.trace_dump(os.path.join(td, "t.csv"))
        kernels = [r["kernel"] for r in csv.DictReader(open(os.path.join(td, "t.csv"))) if r["kind"] == "fprop"]
    return y.float().cpu(), xm.grad.float().cpu(), kernels


STRIDE1 = [  # kind, cin, cout, kernel, shape: forward + dgrad both ride the tile kernel
    ("instance", 128, 128, 3, (2, 8, 16, 8)),            # N_TILE 128, two depth slices per tile
    ("instance", 128, 128, 3, (4, 32, 32, 32)),          # the 128 -> 128 @32^3 layer itself: four slices per tile, persistent grid
    ("instance", 256, 320, 3, (1, 5, 17, 9)),            # ragged in every axis, N_TILE 64 forward (320) / 128 dgrad (256)
    ("instance", 128, 128, (1, 3, 3), (1, 4, 8, 8)),     # nine taps
    ("group", 64, 64, 3, (2, 5, 8, 9)),                  # small volume: 64-channel layer on the tile kernel
    ("instance", 32, 32, 3, (1, 6, 12, 16)),             # N_TILE 32
]


@pytest.mark.parametrize("kind,cin,cout,k,shape", STRIDE1)
def test_stride1_layers_on_the_tma_tile_kernel(kind, cin, cout, k, shape):
    from nndetection_b200.arch import conv_ops as ops
    mine, ref = make_pair(kind, cin, cout, k, 1)
    g = torch.Generator().manual_seed(81)
    x = q(torch.randn(shape[0], cin, *shape[1:], generator=g))
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = q(torch.randn(yr.shape, generator=g))
    yr.backward(gy)
    try:
        ops.set_stream_path(0)                               # keep the 32 / 64-channel cases off the streaming kernel
        y1, dx1, k1 = _run(ops, mine, x, gy, 1)
        y0, dx0, k0 = _run(ops, mine, x, gy, 0)
    finally:
        ops.set_gather_tma(ops.GATHER_TMA_DEFAULT)
        ops.set_stream_path(1)
    assert k1 == ["conv_tct", "conv_tct"] and k0 == ["conv_tc", "conv_tc"]
    assert rel_err(y1, yr.detach()) < 1e-2 and rel_err(dx1, xr.grad) < 3e-2
    assert rel_err(y1, y0) < 2e-3 and rel_err(dx1, dx0) < 2e-3


STRIDE2 = [(32, 64, 2, (2, 12, 34, 36)), (64, 128, 2, (1, 16, 16, 16)), (128, 256, 2, (2, 9, 20, 17)), (256, 320, 2, (1, 16, 16, 16)),
           (32, 64, (1, 2, 2), (1, 6, 24, 40))]


@pytest.mark.parametrize("cin,cout,s,shape", STRIDE2)
def test_stride2_convolutions_on_the_tma_tile_kernel(cin, cout, s, shape):
    """forward through the de-interleaving tensor maps (mode bit 1); the dgrad's >= 4-tap parity classes through the stride-1 form (bit 0)"""
    from nndetection_b200.arch import conv_ops as ops
    mine, ref = make_pair("instance", cin, cout, 3, s)
    g = torch.Generator().manual_seed(82)
    x = q(torch.randn(shape[0], cin, *shape[1:], generator=g))
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = q(torch.randn(yr.shape, generator=g))
    yr.backward(gy)
    try:
        y1, dx1, k1 = _run(ops, mine, x, gy, 3)
        y0, dx0, k0 = _run(ops, mine, x, gy, 0)
    finally:
        ops.set_gather_tma(ops.GATHER_TMA_DEFAULT)
    assert k1[0] == "conv_tct_s2" and k0[0] == "conv_tc_s2"
    assert "conv_tct" in k1 and "conv_tc" not in k1
    assert rel_err(y1, yr.detach()) < 1e-2 and rel_err(dx1, xr.grad) < 3e-2
    assert rel_err(y1, y0) < 2e-3 and rel_err(dx1, dx0) < 2e-3


@pytest.mark.parametrize("cin,cout,shape", [(64, 32, (2, 4, 9, 12)), (128, 64, (1, 5, 8, 8)), (128, 128, (2, 4, 8, 16))])
def test_upconv_input_gradient_on_the_tma_tile_kernel(cin, cout, shape):
    """dgrad of kernel == stride == 2 up-convolutions: a 2x2x2 stride-2 gather of dy"""
    from nndetection_b200.arch import conv_ops as ops
    mine, ref = make_pair("instance", cin, cout, None, 2, transposed=True)
    g = torch.Generator().manual_seed(83)
    x = q(torch.randn(shape[0], cin, *shape[1:], generator=g))
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = q(torch.randn(yr.shape, generator=g))
    yr.backward(gy)
    try:
        y1, dx1, k1 = _run(ops, mine, x, gy, 3)
        y0, dx0, k0 = _run(ops, mine, x, gy, 0)
    finally:
        ops.set_gather_tma(ops.GATHER_TMA_DEFAULT)
    assert "conv_tct_s2" in k1 and "conv_tc_s2" in k0
    assert rel_err(dx1, xr.grad) < 5e-3 and rel_err(dx1, dx0) < 2e-3
