"""The TMA-fed pointwise tcgen05 GEMM (csrc/conv_pw.cu: cp.async.bulk.tensor operand loads with 128B / 64B swizzle, K-major swizzled
UMMA descriptors): 1x1x1 convolutions (U-FPN laterals, nndet/arch/decoder/base.py:216-241) forward + input gradient, and whole
kernel == stride transposed convolutions (up-convolutions, :272-304, with the lateral added in the epilogue, :405) in ONE launch --
against the CPU oracle's torch operators on bf16-exact operands (5e-3 in norm: fp32 accumulation order + one bf16 rounding of the
output) and against the mma.sync gather kernel on identical operands (2e-3)."""
import csv
import os
import tempfile

import pytest
import torch

import tutil as util  # noqa: F401
from test_net_gpu import make_pair, q, rel_err

pytestmark = pytest.mark.gpu


def _kernels(fn):
    from nndetection_b200.arch import conv_ops as ops
    ops.trace_start()
    out = fn()
    with tempfile.TemporaryDirectory() as td:
        ops.trace_dump(os.path.join(td, "t.csv"))
        ks = [(r["kind"], r["kernel"]) for r in csv.DictReader(open(os.path.join(td, "t.csv")))]
    return out, ks


@pytest.mark.parametrize("cin,cout,shape", [
    (32, 32, (2, 8, 12, 16)),        # K = 32: 64-byte swizzle, one k block
    (64, 32, (2, 8, 8, 8)),          # K = 64: 128-byte swizzle
    (320, 128, (2, 4, 4, 4)),        # five k blocks, exactly one 128-row tile
    (128, 128, (1, 5, 7, 9)),        # 315 rows: partial last tile (TMA zero-fills rows beyond M, stores are predicated)
    (256, 128, (3, 4, 8, 8)),        # several tiles per CTA on a small grid? no: 768 rows = 6 tiles
    (64, 64, (4, 16, 16, 20)),       # 20480 rows: persistent CTAs walk more tiles than pipeline stages
])
def test_pointwise_conv_forward_and_input_gradient(cin, cout, shape):
    from nndetection_b200.arch import conv_ops as ops
    mine, ref = make_pair("instance", cin, cout, 1, 1, norm=False)
    g = torch.Generator().manual_seed(5)
    x = q(torch.randn(shape[0], cin, *shape[1:], generator=g))
    res = q(torch.randn(shape[0], cout, *shape[1:], generator=g))
    xr, rr = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    yr = ref(xr) + rr
    gy = q(torch.randn(yr.shape, generator=g))
    yr.backward(gy)
    out = {}
    try:
        for mode in (True, False):
            ops.set_pointwise_tma(mode)
            mine.zero_grad(set_to_none=True)
            xm = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
            rm = res.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)

            def run():
                ym = mine(xm, residual=rm)
                ym.backward(gy.cuda().to(torch.bfloat16))
                return ym
            ym, ks = _kernels(run)
            out[mode] = (ym.detach().float().cpu(), xm.grad.float().cpu(), mine.conv.weight.grad.cpu().clone(), ks)
    finally:
        ops.set_pointwise_tma(True)
    ym, gx, gw, ks = out[True]
    assert [k for kind, k in ks if kind == "fprop"] == ["conv_pw", "conv_pw"]            # forward + input gradient
    assert "conv_pw" not in [k for _, k in out[False][3]]
    assert rel_err(ym, yr.detach()) < 5e-3
    assert rel_err(gx, xr.grad) < 5e-3
    assert rel_err(gw, ref.conv.weight.grad) < 5e-3
    assert rel_err(ym, out[False][0]) < 2e-3 and rel_err(gx, out[False][1]) < 2e-3


@pytest.mark.parametrize("cin,cout,s,shape", [
    (64, 32, 2, (2, 4, 6, 8)),               # N = 8 x 32 = 256 columns: one n tile
    (128, 128, 2, (1, 4, 8, 8)),             # N = 1024: four n tiles of 256 re-reading the A box from L2
    (128, 64, 2, (2, 3, 5, 7)),              # 210 rows: partial tile, N = 512
    (64, 32, (1, 2, 2), (1, 4, 8, 8)),       # anisotropic stride (LIDC-style): four taps
    (32, 32, 2, (1, 8, 8, 8)),               # K = 32: 64-byte swizzle
])
def test_whole_upconvolution_in_one_launch(cin, cout, s, shape):
    from nndetection_b200.arch import conv_ops as ops
    mine, ref = make_pair("instance", cin, cout, None, s, transposed=True)
    st = s if isinstance(s, tuple) else (s,) * 3
    g = torch.Generator().manual_seed(3)
    x = q(torch.randn(shape[0], cin, *shape[1:], generator=g))
    lat = q(torch.randn(shape[0], cout, *[a * b for a, b in zip(shape[1:], st)], generator=g))
    xr, lr = x.clone().requires_grad_(True), lat.clone().requires_grad_(True)
    yr = lr + ref(xr)                                   # decoder/base.py:405
    out = {}
    try:
        for mode in (True, False):
            ops.set_pointwise_tma(mode)
            xm = x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
            lm = lat.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
            with torch.no_grad():
                ym, ks = _kernels(lambda: mine(xm, residual=lm))
            out[mode] = (ym.float().cpu(), ks)
    finally:
        ops.set_pointwise_tma(True)
    assert out[True][1] == [("fprop", "conv_pw_up")] and len(out[False][1]) == st[0] * st[1] * st[2]     # one launch instead of one per tap
    assert rel_err(out[True][0], yr.detach()) < 5e-3
    assert rel_err(out[True][0], out[False][0]) < 2e-3
