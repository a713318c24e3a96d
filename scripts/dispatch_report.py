"""Which kernel serves which convolution launch of a train step, and how many FLOPs ride on each kernel family -- WITHOUT a GPU:
the layer list is derived from the plan the way `RetinaUNetV001.from_config_plan` builds the network (encoder stages, U-FPN laterals /
up-convolutions / output convs, shared head convs per decoder level), every launch geometry comes from `ConvPlan`, and the kernel from
the library's dry-run dispatch queries (`nnd_conv_gather_dispatch`, `nnd_conv_wgrad_dispatch`: the same `*_supported` predicates the
real dispatch uses).  The image layer (Cin <= 4) and the 1x1x1 segmentation conv have kernels of their own and are listed as such.

    python scripts/dispatch_report.py [config] [--mma-s2]      # config: luna (default) | adam | lidc | infer160 | toy | tiny
"""
import os
import sys
from ctypes import c_int, c_longlong

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

GATHER = {0: "conv_igemm (mma.sync)", 1: "conv_tc", 2: "conv_tcs", 3: "conv_tc S2", 4: "conv_pw (TMA)", 6: "conv_tct (TMA)", 7: "conv_tct S2 (TMA)"}
WGRAD = {0: "wgrad generic (mma.sync)", 1: "wgrad halo (mma.sync)", 2: "conv_wgrad_tc", 3: "conv_wgrad_tc32", 4: "conv_wgrad_tcn",
         5: "conv_wgrad_tc SW=2", 6: "conv_wgrad_tma (TMA)", 8: "conv_wgrad_tma S2 (TMA)", 9: "conv_wgrad_tma32 (TMA)"}


def pad32(c):
    return (c + 31) // 32 * 32


def layers_of(arch, patch, bs):
    """(name, cin, cout, k, stride, transposed, in_sp, norm, residual, head_out) for every convolution of the network."""
    t3 = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v,) * 3
    out = []
    sp, c = tuple(patch), arch["in_channels"]
    chans, sps = [], []
    for i, k in enumerate(arch["conv_kernels"]):
        co = arch["start_channels"] if i == 0 else min(c * 2, arch.get("max_channels", 320))
        s = (1, 1, 1) if i == 0 else t3(arch["strides"][i - 1])
        out.append((f"encoder.stage{i}.conv1", c, co, t3(k), s, False, sp, True, False, None))
        sp = tuple((a + 2 * ((kk - 1) // 2) - kk) // ss + 1 for a, kk, ss in zip(sp, t3(k), s))
        out.append((f"encoder.stage{i}.conv2", co, co, t3(k), (1, 1, 1), False, sp, True, False, None))
        c = co
        chans.append(co); sps.append(sp)
    n = len(chans)
    oc = [arch["fpn_channels"]] * n
    for ol in [l for l in range(n) if l < min(arch["decoder_levels"])][::-1]:
        oc[ol] = max(8, oc[ol + 1] // 2)
    for l in range(n):
        out.append((f"decoder.lateral.P{l}", chans[l], oc[l], (1, 1, 1), (1, 1, 1), False, sps[l], False, False, None))
    for l in range(n - 1, 0, -1):
        s = t3(arch["strides"][l - 1])
        out.append((f"decoder.up.P{l}", oc[l], oc[l - 1], s, s, True, sps[l], False, True, None))
    for l in range(n):
        out.append((f"decoder.out.P{l}", oc[l], oc[l], t3(arch["conv_kernels"][l]), (1, 1, 1), False, sps[l], False, False, None))
    hc, C = arch["head_channels"], arch["classifier_classes"]
    for l in arch["decoder_levels"]:
        for br, co_out in (("classifier", 27 * C), ("regressor", 162)):
            out.append((f"head.{br}.c_in@P{l}", oc[l], hc, (3, 3, 3), (1, 1, 1), False, sps[l], True, False, None))
            out.append((f"head.{br}.c_internal0@P{l}", hc, hc, (3, 3, 3), (1, 1, 1), False, sps[l], True, False, None))
            out.append((f"head.{br}.conv_out@P{l}", hc, co_out, (3, 3, 3), (1, 1, 1), False, sps[l], False, False, co_out))
    return out


def report(config="luna", experimental=True, quiet=False):      # experimental=False: strided forms on mma.sync (round-1 default, A/B)
    from nndetection_b200 import _lib as L
    from nndetection_b200.arch.conv_ops import ConvPlan
    from nndetection_b200.configs import make_plan
    lib = L.lib()
    lib.nnd_conv_set_gather_strided_tc(c_int(1 if experimental else 0))
    lib.nnd_conv_set_wgrad_strided_tc(c_int(1 if experimental else 0))
    arch, anc, patch, bs = make_plan(config)
    rows, totals = [], {}

    def add(kind, kernel, gflop):
        totals[kernel] = totals.get(kernel, 0.0) + gflop
        return kernel

    for name, cin, cout, k, s, tr, in_sp, norm, residual, head_out in layers_of(arch, patch, bs):
        T = int(np.prod(k))
        if cin <= 4:                                     # image layer: conv_first_* kernels
            vox = bs * int(np.prod(in_sp))
            gf = 2.0 * cin * cout * T * vox * 1e-9
            rows.append((name, cin, cout, in_sp, s, add("fprop", "conv_first_mma (mma.sync)", gf), "-", add("wgrad", "conv_first_mma (mma.sync)", gf), gf))
            continue
        cdy = pad32(cout)
        plan = ConvPlan(bs, cin, cdy, in_sp, k, s, tuple((kk - 1) // 2 for kk in k) if not tr else 0, tr)
        vox_out = bs * int(np.prod(plan.out_sp))
        gf = 2.0 * cin * cout * T * (bs * int(np.prod(in_sp)) if tr else vox_out) * 1e-9
        fk = set()
        from nndetection_b200.arch import conv_ops
        stacked = tr and not norm and conv_ops.pointwise_tma_enabled() and bs * int(np.prod(in_sp)) >= 128 and all(v in (1, 2) for v in plan.s)
        if stacked:                                      # arch/conv.py: the whole up-convolution in one launch (nnd_conv_upconv_bf16)
            fk.add("conv_pw (TMA)")
        for g in ([] if stacked else plan.fprop):
            if head_out is not None:                    # fp32 outputs written straight into the [N, anchors, C] tensors
                code = lib.nnd_conv_gather_dispatch(g, c_longlong(10 ** 9), c_longlong(head_out), c_int(1), c_int(head_out), c_int(cdy), c_int(1), c_int(0), c_int(0))
            else:
                code = lib.nnd_conv_gather_dispatch(g, c_longlong(int(np.prod(plan.out_sp)) * cout), c_longlong(cout), c_int(0), c_int(cout), c_int(cdy),
                                                    c_int(0 if norm else 1), c_int(1 if residual else 0), c_int(1 if norm else 0))
            fk.add(GATHER[code])
        dk = set()
        for g in plan.dgrad:
            code = lib.nnd_conv_gather_dispatch(g, c_longlong(int(np.prod(in_sp)) * cin), c_longlong(cin), c_int(0), c_int(cin), c_int(pad32(cin)), c_int(0), c_int(0), c_int(0))
            dk.add(GATHER[code])
        if tr and experimental and s[2] == 2 and cin >= 64 and cin % 32 == 0 and cout % 32 == 0:
            wk = {WGRAD[lib.nnd_conv_wgrad_dispatch(plan.wgrad_swapped, c_int(cin), c_int(cout))]}
        else:
            wk = {WGRAD[lib.nnd_conv_wgrad_dispatch(g, c_int(cdy if head_out is not None else cout), c_int(cin))] for g in plan.wgrad}
        f, d, w = "+".join(sorted(fk)), "+".join(sorted(dk)) or "-", "+".join(sorted(wk))
        add("fprop", f, gf); add("wgrad", w, gf)
        if not name.startswith("encoder.stage0.conv1"):
            add("dgrad", d, gf)
        rows.append((name, cin, cout, in_sp, s, f, d, w, gf))
    lib.nnd_conv_set_gather_strided_tc(c_int(1))      # back to the defaults
    lib.nnd_conv_set_wgrad_strided_tc(c_int(1))
    tot = sum(totals.values())
    tensor = sum(v for k, v in totals.items() if "mma.sync" not in k)
    if not quiet:
        print(f"# {config}: batch {bs} x {patch}, strided / transposed forms on {'tcgen05 (default)' if experimental else 'mma.sync (A/B)'}")
        print(f"{'layer':34s} {'Cin':>4s} {'Cout':>4s} {'input':>14s} {'stride':>8s}  {'GFLOP':>7s}  fprop | dgrad | wgrad")
        for name, cin, cout, in_sp, s, f, d, w, gf in rows:
            print(f"{name:34s} {cin:4d} {cout:4d} {'x'.join(map(str, in_sp)):>14s} {'x'.join(map(str, s)):>8s}  {gf:7.1f}  {f} | {d} | {w}")
        print(f"\nGFLOP per kernel family (fprop + dgrad + wgrad of one step, {tot:.0f} GFLOP):")
        for k, v in sorted(totals.items(), key=lambda kv: -kv[1]):
            print(f"  {v:8.1f}  {100 * v / tot:5.1f} %  {k}")
        print(f"  on tcgen05: {100 * tensor / tot:.1f} %")
    return rows, totals, tensor / tot


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    report(args[0] if args else "luna", "--mma-s2" not in sys.argv)
