"""Host logic of the convolution family, checked on CPU: the geometry tables `ConvPlan` hands to the gather-convolution
kernels (csrc/conv_common.cuh: out[n, lo*om + oo, co] = sum_taps sum_ci in[n, lo*s + off_tap, ci] * W[tap_w][co][ci]) are
emulated in numpy and compared with torch's Conv3d / ConvTranspose3d forward and input-gradient on small shapes -- every
(kernel, stride, transposed) form the Retina U-Net uses, incl. the parity-class decomposition of the strided dgrad."""
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from nndetection_b200.arch.conv_ops import ConvPlan


def _emulate(geom, x, w_taps, cout):
    """x [N, Di, Hi, Wi, Cin] (NDHWC), w_taps[tap] = [cout, cin] matrix; returns (out [N, Do, Ho, Wo, cout] with untouched voxels NaN)."""
    a = list(geom)
    N, Di, Hi, Wi, Cin = a[0:5]
    L = a[5:8]; s = a[8:11]; Do, Ho, Wo = a[11:14]; om = a[14:17]; oo = a[17:20]; T = a[20]
    taps = [a[21 + 4 * t: 25 + 4 * t] for t in range(T)]
    out = np.full((N, Do, Ho, Wo, cout), np.nan, dtype=np.float64)
    for n in range(N):
        for ld, lh, lw in itertools.product(range(L[0]), range(L[1]), range(L[2])):
            acc = np.zeros(cout)
            for od, oh, ow, tw in taps:
                i = (ld * s[0] + od, lh * s[1] + oh, lw * s[2] + ow)
                if 0 <= i[0] < Di and 0 <= i[1] < Hi and 0 <= i[2] < Wi:
                    acc += w_taps[tw] @ x[n, i[0], i[1], i[2]]
            out[n, ld * om[0] + oo[0], lh * om[1] + oo[1], lw * om[2] + oo[2]] = acc
    return out


CASES = [((3, 3, 3), (1, 1, 1), False), ((3, 3, 3), (2, 2, 2), False), ((1, 3, 3), (1, 2, 2), False), ((1, 1, 1), (1, 1, 1), False),
         ((2, 2, 2), (2, 2, 2), True), ((1, 2, 2), (1, 2, 2), True)]


@pytest.mark.parametrize("k,s,transposed", CASES)
def test_plan_geometry_reproduces_torch_conv_forward_and_dgrad(k, s, transposed):
    rs = np.random.RandomState(0)
    N, cin, cout, sp = 1, 3, 2, (5, 6, 4)
    p = (0, 0, 0) if transposed else tuple((kk - 1) // 2 for kk in k)
    x = rs.standard_normal((N, cin) + sp)
    T = k[0] * k[1] * k[2]
    plan = ConvPlan(N, cin, cout, sp, k, s, p, transposed)
    xt = torch.from_numpy(x).requires_grad_(True)
    if transposed:
        w = rs.standard_normal((cin, cout) + k)
        y = F.conv_transpose3d(xt, torch.from_numpy(w), stride=s)
        w_f = [w[:, :, a, b, c].T for a, b, c in itertools.product(*[range(v) for v in k])]       # [cout, cin] per tap
        w_b = [w[:, :, a, b, c] for a, b, c in itertools.product(*[range(v) for v in k])]         # dgrad: [cin, cout]
    else:
        w = rs.standard_normal((cout, cin) + k)
        y = F.conv3d(xt, torch.from_numpy(w), stride=s, padding=p)
        w_f = [w[:, :, a, b, c] for a, b, c in itertools.product(*[range(v) for v in k])]
        w_b = [w[:, :, a, b, c].T for a, b, c in itertools.product(*[range(v) for v in k])]
    assert tuple(y.shape[2:]) == plan.out_sp and len(w_f) == T
    # forward: the union of the launches writes every output voxel exactly once
    xn = np.transpose(x, (0, 2, 3, 4, 1))
    out = np.full((N,) + plan.out_sp + (cout,), np.nan)
    for g in plan.fprop:
        o = _emulate(g, xn, w_f, cout)
        m = ~np.isnan(o)
        assert np.isnan(out[m]).all()
        out[m] = o[m]
    assert not np.isnan(out).any()
    np.testing.assert_allclose(out, np.transpose(y.detach().numpy(), (0, 2, 3, 4, 1)), rtol=1e-10, atol=1e-10)
    # input gradient: parity classes of the strided conv / the k = s conv of the transposed conv
    gy = rs.standard_normal(tuple(y.shape))
    y.backward(torch.from_numpy(gy))
    gyn = np.transpose(gy, (0, 2, 3, 4, 1))
    dx = np.full((N,) + sp + (cin,), np.nan)
    for g in plan.dgrad:
        o = _emulate(g, gyn, w_b, cin)
        m = ~np.isnan(o)
        assert np.isnan(dx[m]).all()
        dx[m] = o[m]
    if plan.dgrad_covers_all:
        assert not np.isnan(dx).any()
    dx = np.nan_to_num(dx, nan=0.0)              # voxels no launch writes receive no gradient (zero-filled by the caller)
    np.testing.assert_allclose(dx, np.transpose(xt.grad.numpy(), (0, 2, 3, 4, 1)), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("k,s,transposed", CASES)
def test_plan_geometry_reproduces_torch_weight_gradient(k, s, transposed):
    """dW[tap][co][ci] = sum_lo dy[lo*om + oo][co] * x[lo*s + off_tap][ci] over the plan's wgrad launches (csrc/conv_wgrad*.cu)."""
    rs = np.random.RandomState(1)
    N, cin, cout, sp = 2, 3, 2, (4, 5, 6)
    p = (0, 0, 0) if transposed else tuple((kk - 1) // 2 for kk in k)
    x = rs.standard_normal((N, cin) + sp)
    plan = ConvPlan(N, cin, cout, sp, k, s, p, transposed)
    xt = torch.from_numpy(x)
    if transposed:
        w = torch.from_numpy(rs.standard_normal((cin, cout) + k)).requires_grad_(True)
        y = F.conv_transpose3d(xt, w, stride=s)
    else:
        w = torch.from_numpy(rs.standard_normal((cout, cin) + k)).requires_grad_(True)
        y = F.conv3d(xt, w, stride=s, padding=p)
    gy = rs.standard_normal(tuple(y.shape))
    y.backward(torch.from_numpy(gy))
    xn, gyn = np.transpose(x, (0, 2, 3, 4, 1)), np.transpose(gy, (0, 2, 3, 4, 1))
    T = k[0] * k[1] * k[2]
    dw = np.zeros((T, cout, cin))
    for g in plan.wgrad:
        a = list(g)
        Di, Hi, Wi = a[1:4]; L = a[5:8]; st = a[8:11]; om = a[14:17]; oo = a[17:20]
        taps = [a[21 + 4 * t: 25 + 4 * t] for t in range(a[20])]
        for n in range(N):
            for ld, lh, lw in itertools.product(range(L[0]), range(L[1]), range(L[2])):
                d = gyn[n, ld * om[0] + oo[0], lh * om[1] + oo[1], lw * om[2] + oo[2]]
                for od, oh, ow, tw in taps:
                    i = (ld * st[0] + od, lh * st[1] + oh, lw * st[2] + ow)
                    if 0 <= i[0] < Di and 0 <= i[1] < Hi and 0 <= i[2] < Wi:
                        dw[tw] += np.outer(d, xn[n, i[0], i[1], i[2]])
    ref = w.grad.numpy().reshape(w.shape[0], w.shape[1], T)
    ref = np.transpose(ref, (2, 1, 0)) if transposed else np.transpose(ref, (2, 0, 1))      # -> [tap][co][ci]
    np.testing.assert_allclose(dw, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("s,shape", [((2, 2, 2), (2, 7, 9, 40)), ((1, 2, 2), (1, 4, 6, 13)), ((2, 2, 2), (1, 5, 4, 34))])
def test_deinterleaved_row_layout_of_the_strided_wgrad_kernel(s, shape):
    """Index arithmetic of conv_wgrad_tc.cu, SW = 2, emulated in numpy (the MMA itself is the validated stride-1 machinery): a row of
    16 outputs loads the 33 input voxels 2*w0 - 1 .. 2*w0 + 31 into slots [odd plane 0..16 | even plane 17..32] (even v -> slot v/2,
    odd v -> 17 + v/2), tap dx reads 16 consecutive slots from 0 / 17 / 1 -- summed over rows this must be torch's weight gradient."""
    RW, XW = 16, 33
    N, D, H, W = shape
    cin, cout = 3, 4
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, cin, D, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv3d(x, w, stride=s, padding=1)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    Ld, Lh, Lw = y.shape[2:]
    xn, dyn = x.numpy(), gy.numpy()
    dw = np.zeros((cout, cin, 3, 3, 3))
    wsegs = -(-Lw // RW)
    tap_start = {0: 0, 1: RW + 1, 2: 1}                                   # slot where tap dx = k - 1 starts
    for n, d, h, ws in itertools.product(range(N), range(Ld), range(Lh), range(wsegs)):
        A = np.zeros((RW, cout))
        for v in range(RW):
            if ws * RW + v < Lw:
                A[v] = dyn[n, :, d, h, ws * RW + v]
        for dz, dyo in itertools.product((-1, 0, 1), repeat=2):
            dd, hh = d * s[0] + dz, h * s[1] + dyo
            B = np.zeros((XW, cin))
            if 0 <= dd < D and 0 <= hh < H:
                w_in0 = ws * RW * 2 - 1
                for v in range(XW):
                    if 0 <= w_in0 + v < W:
                        slot = (RW + 1) + (v >> 1) if (v & 1) else (v >> 1)
                        B[slot] = xn[n, :, dd, hh, w_in0 + v]
            for k in range(3):
                dw[:, :, dz + 1, dyo + 1, k] += A.T @ B[tap_start[k]:tap_start[k] + RW]
    assert np.allclose(dw, w.grad.numpy(), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("s,shape", [((2, 2, 2), (2, 3, 5, 20)), ((1, 2, 2), (1, 4, 3, 7))])
def test_deinterleaved_row_layout_serves_transposed_conv_weight_gradient(s, shape):
    """Same kernel arithmetic with the operands' roles swapped (arch/conv.py, opt-in path): rows run over the layer INPUT grid (dense
    operand x), the strided operand is dy read at s * i + tap with taps in {0, 1}: group offsets (dz, dy) in {0, 1}^2, tap dx = 0 ->
    even plane (slot 17), dx = 1 -> odd plane from slot 1.  Must equal torch's ConvTranspose3d weight gradient [Cin, Cout, k, k, k]."""
    RW, XW = 16, 33
    N, D, H, W = shape
    cin, cout = 4, 3
    g = torch.Generator().manual_seed(6)
    x = torch.randn(N, cin, D, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(cin, cout, *s, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose3d(x, w, stride=s)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    Do, Ho, Wo = y.shape[2:]
    xn, dyn = x.numpy(), gy.numpy()
    dw = np.zeros((cin, cout, *s))
    wsegs = -(-W // RW)
    tap_start = {0: RW + 1, 1: 1}                                         # off_w = 0 -> even plane, +1 -> odd plane shifted
    for n, d, h, ws in itertools.product(range(N), range(D), range(H), range(wsegs)):
        A = np.zeros((RW, cin))
        for v in range(RW):
            if ws * RW + v < W:
                A[v] = xn[n, :, d, h, ws * RW + v]
        for dz, dyo in itertools.product(range(s[0]), range(s[1])):
            dd, hh = d * s[0] + dz, h * s[1] + dyo
            B = np.zeros((XW, cout))
            if 0 <= dd < Do and 0 <= hh < Ho:
                w_in0 = ws * RW * 2 - 1
                for v in range(XW):
                    if 0 <= w_in0 + v < Wo:
                        slot = (RW + 1) + (v >> 1) if (v & 1) else (v >> 1)
                        B[slot] = dyn[n, :, dd, hh, w_in0 + v]
            for dx in range(2):
                dw[:, :, dz, dyo, dx] += A.T @ B[tap_start[dx]:tap_start[dx] + RW]
    assert np.allclose(dw, w.grad.numpy(), rtol=1e-10, atol=1e-10)


def test_transposed_plan_exposes_the_swapped_weight_gradient_geometry():
    """ConvPlan.wgrad_swapped: a single gather geometry over the INPUT grid reading the output grid at s * i + tap, all k^3 taps."""
    plan = ConvPlan(2, 64, 32, (4, 6, 8), 2, 2, 0, True)
    g = list(plan.wgrad_swapped)
    assert g[:21] == [2, 8, 12, 16, 32, 4, 6, 8, 2, 2, 2, 4, 6, 8, 1, 1, 1, 0, 0, 0, 8]
    taps = [tuple(g[21 + 4 * t: 25 + 4 * t]) for t in range(8)]
    assert taps == [(a, b, c, (a * 2 + b) * 2 + c) for a, b, c in itertools.product(range(2), repeat=3)]


def _emulate_s2_tile_kernel(x, w_taps, offs, out_sp, MT=2, BH=16, BW=8, sd=2):
    """Index arithmetic of conv_tc.cu, S2 = 1 in numpy: per tile (MT x 16 x 8 outputs) the halo is staged de-interleaved per axis
    (odd plane: slots 0..n <- positions 2 (o0 + p) - 1; even plane: slots n+1..2n <- 2 (o0 + p)), tap offset -1 / 0 / +1 reads from
    slot 0 / n + 1 / 1 of its axis, MMA row r = (hy, wx) = (r // 8, r % 8) adds hy row pitches and wx slots, depth slice mt adds mt
    slice pitches.  x: [Di, Hi, Wi, Cin]; w_taps[t]: [Cout, Cin]; offs[t] = (od, oh, ow)."""
    Di, Hi, Wi, Cin = x.shape
    Ld, Lh, Lw = out_sp
    HY, HX, ZS = 2 * BH + 1, 2 * BW + 1, 2 * MT + 1
    sl = lambda off, n: n + 1 if off == 0 else (1 if off > 0 else 0)
    out = np.zeros((Ld, Lh, Lw, w_taps[0].shape[0]))
    for d0, h0, w0 in itertools.product(range(0, Ld, MT), range(0, Lh, BH), range(0, Lw, BW)):
        halo = np.zeros((ZS, HY, HX, Cin))
        for z, y in itertools.product(range(ZS), range(HY)):
            d = (2 * (d0 + z) - 1 if z <= MT else 2 * (d0 + z - (MT + 1))) if sd == 2 else d0 - 1 + z
            h = 2 * (h0 + y) - 1 if y <= BH else 2 * (h0 + y - (BH + 1))
            if not ((sd == 2 or z < MT + 2) and 0 <= d < Di and 0 <= h < Hi):
                continue
            w_in0 = 2 * w0 - 1
            for v in range(HX):
                if 0 <= w_in0 + v < Wi:
                    slot = (BW + 1) + (v >> 1) if (v & 1) else (v >> 1)
                    halo[z, y, slot] = x[d, h, w_in0 + v]
        flat = halo.reshape(ZS * HY * HX, Cin)
        for t, (od, oh, ow) in enumerate(offs):
            tapoff = ((sl(od, MT) if sd == 2 else od + 1) * HY + sl(oh, BH)) * HX + sl(ow, BW)
            for mt, r in itertools.product(range(MT), range(BH * BW)):
                hy, wx = r // 8, r % 8
                d, h, ww = d0 + mt, h0 + hy, w0 + wx
                if d < Ld and h < Lh and ww < Lw:
                    out[d, h, ww] += w_taps[t] @ flat[tapoff + mt * HY * HX + hy * HX + wx]
    return out


@pytest.mark.parametrize("in_sp", [(7, 18, 20), (8, 34, 17), (5, 9, 16)])
def test_deinterleaved_halo_of_the_strided_tile_kernel_conv(in_sp):
    """3x3x3 stride-2 padding-1 convolution through the emulated S2 addressing == torch."""
    g = torch.Generator().manual_seed(8)
    cin, cout = 3, 5
    x = torch.randn(1, cin, *in_sp, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv3d(x, w, stride=2, padding=1)[0].permute(1, 2, 3, 0).numpy()
    offs = [(a - 1, b - 1, c - 1) for a, b, c in itertools.product(range(3), repeat=3)]
    taps = [w[:, :, a, b, c].numpy() for a, b, c in itertools.product(range(3), repeat=3)]
    out = _emulate_s2_tile_kernel(x[0].permute(1, 2, 3, 0).numpy(), taps, offs, ref.shape[:3])
    assert np.allclose(out, ref, rtol=1e-10, atol=1e-10)


def test_deinterleaved_halo_of_the_strided_tile_kernel_upconv_dgrad():
    """dgrad of a kernel == stride == 2 transposed convolution = 2x2x2 stride-2 convolution of dy with taps at offsets {0, 1}
    (ConvPlan.dgrad of a transposed layer) through the emulated S2 addressing == torch autograd."""
    g = torch.Generator().manual_seed(9)
    cin, cout, in_sp = 4, 3, (3, 9, 10)
    x = torch.randn(1, cin, *in_sp, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cin, cout, 2, 2, 2, generator=g, dtype=torch.float64)
    y = F.conv_transpose3d(x, w, stride=2)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    offs = list(itertools.product(range(2), repeat=3))
    taps = [w[:, :, a, b, c].numpy() for a, b, c in offs]                  # dx[ci] += W[ci, co, tap] dy[co]
    out = _emulate_s2_tile_kernel(gy[0].permute(1, 2, 3, 0).numpy(), taps, offs, in_sp)
    assert np.allclose(out, x.grad[0].permute(1, 2, 3, 0).numpy(), rtol=1e-10, atol=1e-10)


def test_deinterleaved_halo_with_an_unstrided_depth_axis():
    """First stride (1, 2, 2) of anisotropic plans (LIDC-shaped config): ordinary halo slots along d, de-interleaved h / w."""
    g = torch.Generator().manual_seed(10)
    cin, cout, in_sp = 3, 4, (5, 18, 20)
    x = torch.randn(1, cin, *in_sp, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv3d(x, w, stride=(1, 2, 2), padding=1)[0].permute(1, 2, 3, 0).numpy()
    offs = [(a - 1, b - 1, c - 1) for a, b, c in itertools.product(range(3), repeat=3)]
    taps = [w[:, :, a, b, c].numpy() for a, b, c in itertools.product(range(3), repeat=3)]
    out = _emulate_s2_tile_kernel(x[0].permute(1, 2, 3, 0).numpy(), taps, offs, ref.shape[:3], sd=1)
    assert np.allclose(out, ref, rtol=1e-10, atol=1e-10)


def test_dispatch_table_of_the_luna_train_step():
    """scripts/dispatch_report.py: every convolution launch of the LUNA-shaped train step through the library's dry-run dispatch
    queries (host-only predicates, the ones the real dispatch uses).  Pins which layers ride on tcgen05 -- 90 % of the step's
    8.24 TFLOP with the strided / transposed forms on mma.sync (round-1 default, experimental=False), 97 % with their tcgen05 kernels and the TMA pointwise GEMM (defaults since round 2) -- picked for exactly the strided forms."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("dispatch_report", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                                 "scripts", "dispatch_report.py"))
    dr = importlib.util.module_from_spec(spec); spec.loader.exec_module(dr)
    rows, totals, frac = dr.report("luna", experimental=False, quiet=True)
    by = {r[0]: r for r in rows}
    assert abs(sum(totals.values()) - 8242) < 15 and 0.89 < frac < 0.94
    assert by["encoder.stage0.conv2"][5:8] == ("conv_tcs", "conv_tcs", "conv_wgrad_tma32 (TMA)")
    # 128 -> 128: TMA-fed tile kernel (fprop + dgrad) and TMA-fed wgrad (defaults since round 2)
    assert by["encoder.stage2.conv2"][5:8] == ("conv_tct (TMA)", "conv_tct (TMA)", "conv_wgrad_tma (TMA)")
    assert by["encoder.stage1.conv2"][7] == "conv_wgrad_tma (TMA)" and by["head.regressor.c_in@P3"][7] == "conv_wgrad_tma (TMA)"
    assert by["encoder.stage1.conv1"][5] == "conv_igemm (mma.sync)" and by["encoder.stage1.conv1"][7] == "wgrad halo (mma.sync)"
    assert by["head.regressor.conv_out@P2"][5:8] == ("conv_tct (TMA)", "conv_tct (TMA)", "conv_wgrad_tma (TMA)")   # dy padded to 192 channels: 3 x 64
    rows, totals, frac_x = dr.report("luna", experimental=True, quiet=True)
    by = {r[0]: r for r in rows}
    assert 0.97 < frac_x < 0.995                                                                # 99.0 % of the FLOPs on tcgen05 kernels
    for l in (0, 1, 2, 3):
        assert by[f"decoder.lateral.P{l}"][5] == "conv_pw (TMA)" and by[f"decoder.lateral.P{l}"][6] == "conv_pw (TMA)"        # 1x1x1: fprop + dgrad
    for l in (1, 2, 3, 4):
        assert by[f"decoder.up.P{l}"][5] == "conv_pw (TMA)"                                     # whole up-convolution in one launch
    for l in (1, 2, 3, 4):
        assert by[f"encoder.stage{l}.conv1"][5] == "conv_tct S2 (TMA)" and by[f"encoder.stage{l}.conv1"][7] == "conv_wgrad_tma S2 (TMA)"
        assert by[f"encoder.stage{l}.conv1"][6] == "conv_pw (TMA)+conv_tct (TMA)"     # stride-2 dgrad: 1-tap class pointwise, 2- / 4- / 8-tap classes tile kernel
    for l in (1, 2, 3, 4):
        assert by[f"decoder.up.P{l}"][6] == "conv_tct S2 (TMA)" and by[f"decoder.up.P{l}"][7] == "conv_wgrad_tma S2 (TMA)"
    assert by["encoder.stage0.conv2"][5:8] == ("conv_tcs", "conv_tcs", "conv_wgrad_tma32 (TMA)")  # stride-1 layers untouched
    _, _, frac_after = dr.report("luna", experimental=False, quiet=True)
    assert frac_after == frac                                                                      # switches restored


def test_item_order_weight_pack_layout():
    """Layout contract between `repack_items_kernel` (csrc/misc.cu) and the BULK producer of csrc/conv_tc.cu, in numpy: destination
    16-byte group i = (((nt * KC + kc) * T + t) * kg + g) * n_tile + n  <-  source group ((t * rows_pad + nt * n_tile + n) * K / 8 +
    kc * kg + g); the producer fetches slice (nt, kc, t) = n_tile * kg groups starting at ((nt * KC + kc) * T + t) * n_tile * kg and the
    MMA's B operand layout inside a slice is [k group][n][8]."""
    T, rows_pad, K, n_tile, kg = 27, 256, 64, 128, 4
    src = np.arange(T * rows_pad * K, dtype=np.int64).reshape(T, rows_pad, K // 8, 8)            # element ids, 8 per 16-byte group
    KC, NT = K // (8 * kg), rows_pad // n_tile
    dst = np.empty((T * rows_pad * K // 8, 8), dtype=np.int64)
    i = np.arange(dst.shape[0])
    n = i % n_tile; r = i // n_tile
    g = r % kg; r //= kg
    t = r % T; r //= T
    kc = r % KC; nt = r // KC
    dst[i] = src[t, nt * n_tile + n, kc * kg + g]
    for (nt_, kc_, t_) in [(0, 0, 0), (1, 1, 26), (1, 0, 13)]:
        start = ((nt_ * KC + kc_) * T + t_) * n_tile * kg
        sl = dst[start:start + n_tile * kg].reshape(kg, n_tile, 8)                                # what one bulk copy lands in smem
        want = src[t_, nt_ * n_tile:(nt_ + 1) * n_tile, kc_ * kg:(kc_ + 1) * kg].transpose(1, 0, 2)   # [k group][n][8] of that tap / chunk / tile
        assert np.array_equal(sl, want)


def test_split_k_plan_of_the_tma_weight_gradients():
    """Host-only: the workspace the TMA-fed wgrad kernels ask for = splits x taps x Cout x Cin fp32 partials, i.e. the split-K plan
    (csrc/conv_wgrad_tma.cu:wm_plan, conv_wgrad_tma_s2.cu:ws_plan): one CTA group per filter row -- per PAIR of rows when dy has 64
    channels (rows (dy, dy - 1) at stride 1, (+1, -1) at stride 2) --, one wave of 148 SMs, >= 2 units per split."""
    from ctypes import c_int, c_longlong
    from nndetection_b200 import _lib as L
    lib = L.lib()
    lib.nnd_conv_wgrad_workspace_bytes.restype = c_longlong

    def splits(cin, cout, in_sp, s, n=4):
        plan = ConvPlan(n, cin, cout, in_sp, 3, s, 1, False)
        code = lib.nnd_conv_wgrad_dispatch(plan.wgrad[0], c_int(cout), c_int(cin))
        b = int(lib.nnd_conv_wgrad_workspace_bytes(plan.wgrad[0], c_int(cout), c_int(cin), c_int(cout), c_int(cin)))
        assert b % (27 * cout * cin * 4) == 0
        return code, b // (27 * cout * cin * 4)

    assert splits(128, 128, (32, 32, 32), 1) == (6, 16)          # 9 filter rows x 1 co tile x 1 ci tile -> 148 // 9 = 16 splits
    assert splits(64, 64, (64, 64, 64), 1) == (6, 24)            # 64-channel dy: 6 paired groups -> 24 splits
    assert splits(256, 256, (16, 16, 16), 1) == (6, 4)           # 9 rows x 2 co tiles x 2 ci tiles = 36 CTAs per split
    assert splits(320, 320, (8, 8, 8), 1) == (6, 1)              # 9 x 3 x 5 = 135 tiles: no split
    assert splits(32, 64, (128, 128, 128), 2) == (8, 24)         # stride 2, 64-channel dy: rows (+1, -1) paired, row 0 alone: 6 groups
    assert splits(64, 128, (64, 64, 64), 2) == (8, 16)           # 9 groups
    assert splits(32, 32, (128, 128, 128), 1) == (9, 148)        # 32-channel layers: TMA-fed stacked-tap kernel, one partial block per CTA
    assert splits(128, 128, (4, 4, 4), 1)[1] >= 1                # 4^3 level: boxes larger than the tensor, still planned
