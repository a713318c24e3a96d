"""Host plans of the TMA-fed kernels, replayed in numpy (no GPU): `nnd_conv_tct_plan_debug` returns the tables csrc/conv_tct.cu hands to
its kernel -- halo planes (entries, first coordinate, byte offset inside a shared-memory stage, tensor-map box extents) and per tap the
byte offset of its first operand row, the pitch of its 8-row groups (SBO) and of its depth slices.  The replay fills a stage the way
the TMA unit does (box rows densely in (d, h, w) order, every s-th tensor element with element stride s, zeros out of bounds) and reads
every tap's 128 operand rows the way the UMMA descriptor does (16 groups of 8 consecutive rows, SBO apart, from the tap's start row):
each row must be the voxel the convolution multiplies with that tap.  (The swizzle is transparent to this model: both engines apply it
to the same absolute shared-memory address -- the device-verified property, tests/test_wgrad_tma_gpu.py.)"""
import itertools
from ctypes import c_int

import numpy as np
import pytest

from nndetection_b200.arch.conv_ops import ConvPlan

TBH, TBW = 16, 8


def _plan(geom, mt, s2):
    from nndetection_b200 import _lib as L
    lib = L.lib()
    out = (c_int * 512)()
    n = lib.nnd_conv_tct_plan_debug(geom, c_int(mt), c_int(1 if s2 else 0), out, c_int(512))
    assert n > 0
    v = list(out[:n])
    n_planes, a_stage, a_bytes, rowb = v[:4]
    planes = [v[4 + 10 * p: 14 + 10 * p] for p in range(n_planes)]
    T = geom[20]
    taps = [v[4 + 10 * n_planes + 3 * t: 7 + 10 * n_planes + 3 * t] for t in range(T)]
    return n_planes, a_stage, a_bytes, rowb, planes, taps


def _replay(geom, mt, s2, tiles):
    g = list(geom)
    N, Di, Hi, Wi = g[0:4]
    Ld, Lh, Lw = g[5:8]
    sd, sh, sw = g[8:11]
    T = g[20]
    offs = [(g[21 + 4 * t], g[22 + 4 * t], g[23 + 4 * t]) for t in range(T)]
    n_planes, a_stage, a_bytes, rowb, planes, taps = _plan(geom, mt, s2)
    x = np.arange(1, N * Di * Hi * Wi + 1, dtype=np.int64).reshape(N, Di, Hi, Wi)           # voxel ids; 0 = zero fill

    def voxel(n, d, h, w):
        return int(x[n, d, h, w]) if 0 <= d < Di and 0 <= h < Hi and 0 <= w < Wi else 0

    assert sum(p[0] * p[1] * p[2] for p in planes) * rowb == a_bytes                         # expect_tx of a stage
    for (n, d0, h0, w0) in tiles:
        stage = {}
        for pd, ph, pw, cd, ch, cw, base, bw, bh, bd in planes:
            assert base % 1024 == 0 and base + pd * ph * pw * rowb <= a_stage
            assert (bw, bh, bd) == ((pw - 1) * sw + 1, (ph - 1) * sh + 1, (pd - 1) * sd + 1)  # box spans the strided entries
            for iz, iy, ix in itertools.product(range(pd), range(ph), range(pw)):
                addr = base + ((iz * ph + iy) * pw + ix) * rowb
                assert addr not in stage
                stage[addr] = voxel(n, d0 * sd + cd + iz * sd, h0 * sh + ch + iy * sh, w0 * sw + cw + ix * sw)
        for t, (od, oh, ow) in enumerate(offs):
            tap_off, sbo, slice_step = taps[t]
            for m in range(mt):
                for r in range(128):
                    hy, wx = r // 8, r % 8
                    addr = tap_off + m * slice_step + hy * sbo + wx * rowb
                    assert addr in stage, (t, m, r)                                           # even masked rows read inside the stage
                    if d0 + m < Ld and h0 + hy < Lh and w0 + wx < Lw:
                        want = voxel(n, (d0 + m) * sd + od, (h0 + hy) * sh + oh, (w0 + wx) * sw + ow)
                        assert stage[addr] == want, (t, m, r, stage[addr], want)


@pytest.mark.parametrize("k,s,in_sp,mt", [(3, 1, (6, 20, 11), 4), (3, 1, (3, 16, 8), 2), ((1, 3, 3), 1, (2, 17, 9), 2)])
def test_stride1_halo_plane_and_tap_table(k, s, in_sp, mt):
    plan = ConvPlan(2, 32, 32, in_sp, k, s, tuple(v // 2 for v in ((k,) * 3 if isinstance(k, int) else k)), False)
    tiles = [(0, 0, 0, 0), (1, mt, 16, 8), (1, 0, 0, 8)]
    _replay(plan.fprop[0], mt, False, tiles)                 # forward
    for g in plan.dgrad:                                     # input gradient: the same gather with mirrored taps
        _replay(g, mt, False, tiles[:2])


@pytest.mark.parametrize("s,in_sp", [(2, (9, 34, 19)), ((1, 2, 2), (5, 33, 18))])
def test_stride2_planes_through_element_stride_boxes(s, in_sp):
    plan = ConvPlan(1, 32, 64, in_sp, 3, s, 1, False)
    _replay(plan.fprop[0], 2, True, [(0, 0, 0, 0), (0, 2, 16, 8), (0, 0, 0, 8)])
    # the dgrad of a stride-2 convolution runs as parity classes: stride-1 gathers of dy with 1, 2, 4 or 8 of the taps each
    for g in plan.dgrad:
        if g[20] >= 2:
            _replay(g, 2, False, [(0, 0, 0, 0)])


def test_upconv_input_gradient_is_a_stride2_gather_with_taps_0_and_1():
    plan = ConvPlan(1, 64, 32, (4, 9, 6), 2, 2, 0, True)
    for g in plan.dgrad:
        _replay(g, 2, True, [(0, 0, 0, 0), (0, 2, 0, 0)])


# ---------------------------------------------------------------------------------------------------------------- weight gradients
def _wgrad_plan(geom, cdy, cx, strided):
    from nndetection_b200 import _lib as L
    lib = L.lib()
    out = (c_int * 256)()
    fn = lib.nnd_conv_wgrad_tma_s2_plan_debug if strided else lib.nnd_conv_wgrad_tma_plan_debug
    n = fn(geom, c_int(cdy), c_int(cx), out, c_int(256))
    assert n > 0
    v = list(out[:n])
    if strided:
        keys = ("narrow", "bw", "bh", "pair", "nb", "cb", "HB", "WS", "n_groups", "ci_tiles", "splits", "units_per_split", "total_units", "dxmask")
    else:
        keys = ("narrow", "bw", "bh", "pair", "nb", "HB", "WS", "n_groups", "ci_tiles", "splits", "units_per_split", "total_units")
    p = dict(zip(keys, v))
    base = len(keys)
    p["groups"] = [dict(dz=v[base + 8 * i], ty=v[base + 8 * i + 1], tw=v[base + 8 * i + 2: base + 8 * i + 5], tw2=v[base + 8 * i + 5: base + 8 * i + 8])
                   for i in range(p["n_groups"])]
    return p


def _replay_wgrad(geom, cdy, cx, strided):
    """Which (weight tap, dy voxel) products does a launch accumulate?  Units, K-steps, MMA row halves and tap columns as the kernels of
    csrc/conv_wgrad_tma.cu / conv_wgrad_tma_s2.cu index them (restated here), driven by the plan the library returns: every tap must meet
    every dy voxel whose x partner exists EXACTLY once -- across CTA groups, paired filter rows (units start at h = -1), split ranges
    and skipped padding slices."""
    g = list(geom)
    N, Di, Hi, Wi = g[0:4]
    D, H, W = g[5:8]
    sd, sh, sw = g[8:11]
    T = g[20]
    taps = {(g[21 + 4 * t], g[22 + 4 * t], g[23 + 4 * t]): g[24 + 4 * t] for t in range(T)}
    p = _wgrad_plan(geom, cdy, cx, strided)
    bw, bh, pair = p["bw"], p["bh"], p["pair"]
    assert p["HB"] == -(-(H + pair) // bh) and p["WS"] == -(-W // bw)
    assert p["total_units"] == N * D * p["HB"] * p["WS"]
    assert p["splits"] * p["units_per_split"] >= p["total_units"] > (p["splits"] - 1) * p["units_per_split"]
    seen = {}
    for grp in p["groups"]:
        dz, ty = grp["dz"], grp["ty"]
        for u in range(p["total_units"]):                       # the split ranges tile [0, total_units): every unit once per group
            ws_, r = u % p["WS"], u // p["WS"]
            hb, r = r % p["HB"], r // p["HB"]
            d, n = r % D, r // D
            xd = d * sd + dz
            if not 0 <= xd < Di:
                continue                                        # the kernels skip the unit: its x slice is padding
            w0, h0 = ws_ * bw, hb * bh - pair
            for kk in range(bw * bh):                           # the 64 voxels of the unit (4 K-steps x 16)
                if p["narrow"]:
                    h, w = h0 + kk // 8, w0 + kk % 8
                else:
                    h, w = h0 + kk // 16, w0 + kk % 16
                for half in range(2 if pair else 1):            # MMA rows 64..127: dy one h row further
                    hd = h + half
                    if not (0 <= hd < H and 0 <= w < W):
                        continue                                # zero-filled dy row
                    for t in range(3):                          # dx = t - 1
                        tw = (grp["tw2"] if half else grp["tw"])[t]
                        if tw == 255:
                            continue
                        xh, xw = h * sh + ty, w * sw + t - 1
                        if not (0 <= xh < Hi and 0 <= xw < Wi):
                            continue                            # zero-filled x voxel
                        key = (tw, n, d, hd, w)
                        seen[key] = seen.get(key, 0) + 1
    want = {}
    for (od, oh, ow), tw in taps.items():
        for n, d, h, w in itertools.product(range(N), range(D), range(H), range(W)):
            if 0 <= d * sd + od < Di and 0 <= h * sh + oh < Hi and 0 <= w * sw + ow < Wi:
                want[(tw, n, d, h, w)] = 1
    assert seen == want


@pytest.mark.parametrize("cin,cout,k,in_sp", [(128, 128, 3, (3, 6, 20)), (64, 64, 3, (3, 7, 9)), (192, 64, (1, 3, 3), (2, 9, 24)), (64, 128, 3, (2, 4, 4))])
def test_stride1_weight_gradient_plan_covers_every_product_once(cin, cout, k, in_sp):
    pad = tuple(v // 2 for v in ((k,) * 3 if isinstance(k, int) else k))
    plan = ConvPlan(2, cin, cout, in_sp, k, 1, pad, False)
    _replay_wgrad(plan.wgrad[0], cout, cin, False)


@pytest.mark.parametrize("cin,cout,s,in_sp", [(32, 64, 2, (5, 9, 13)), (64, 128, (1, 2, 2), (3, 12, 12)), (64, 64, 2, (4, 8, 34))])
def test_strided_weight_gradient_plan_covers_every_product_once(cin, cout, s, in_sp):
    plan = ConvPlan(1, cin, cout, in_sp, 3, s, 1, False)
    _replay_wgrad(plan.wgrad[0], cout, cin, True)


def test_upconv_weight_gradient_plan_with_swapped_operands():
    """kernel == stride transposed convolution: the dense operand is the layer input, the strided one dy read at 2 i + {0, 1}"""
    plan = ConvPlan(1, 64, 32, (3, 5, 9), 2, 2, 0, True)
    _replay_wgrad(plan.wgrad_swapped, 64, 32, True)
