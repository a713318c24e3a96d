"""Device probe of the TMA-fed wgrad kernel (csrc/conv_wgrad_tma.cu): one descriptor variant per process
    python scripts/probe_wgrad_tma.py <mode> [quick]
mode = value for nnd_conv_set_wgrad_tma (bit 0 on, bit 1 base_offset descriptors, bit 2 unstacked taps).  For every shape: relative error of
dW against the fp32 cuDNN weight gradient of the same bf16 operands and against the cp.async kernel, then (if correct) timings of both."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nndetection_b200.arch import conv_ops as ops
from nndetection_b200.arch.conv import ConvInstanceRelu

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
quick = len(sys.argv) > 2
dev = torch.device("cuda")

SMALL = [  # cin, cout, (D, H, W), batch, kernel
    (64, 64, (4, 8, 16), 1, 3),
    (128, 128, (5, 6, 32), 2, 3),
    (64, 128, (3, 9, 20), 1, 3),          # ragged: W = 20 (wide, second segment 4 wide), H = 9
    (128, 64, (4, 8, 8), 2, 3),           # narrow units
    (128, 128, (6, 7, 24), 1, 3),         # narrow, three segments, odd H
    (256, 320, (4, 4, 8), 1, 3),          # three co tiles (last half), NB = 2 over 256 -> ci tiles 2
    (320, 128, (4, 4, 4), 2, 3),          # NB = 1, five ci tiles, 4^3 level
    (64, 64, (5, 12, 12), 1, (1, 3, 3)),  # 1x3x3 filter
]
BIG = [(128, 128, (32, 32, 32), 4, 3), (64, 64, (64, 64, 64), 4, 3), (256, 256, (16, 16, 16), 4, 3), (128, 128, (16, 16, 16), 4, 3),
       (128, 128, (8, 8, 8), 4, 3), (320, 320, (8, 8, 8), 4, 3)]


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def setup(cin, cout, sp, bs, k):
    pad = 1 if isinstance(k, int) else tuple(v // 2 for v in k)
    layer = ConvInstanceRelu(3, cin, cout, kernel_size=k, stride=1, padding=pad).to(dev)
    g = torch.Generator(device="cpu").manual_seed(cin * 7 + cout + sp[2])
    x = torch.randn(bs, cin, *sp, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    dy = torch.randn(bs, cout, *sp, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    plan = layer.plan(bs, tuple(sp))
    T = plan.T
    return layer, x, dy, plan, T, pad


def run(x, dy, plan, T, cin, cout, shape):
    dw = torch.zeros(shape, dtype=torch.float32, device=dev)
    for g in plan.wgrad:
        ops.conv_wgrad(dy, cout, x, cin, g, dw, cin * T, T, 1, cout, cin)
    return dw


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ok_all = True
for cin, cout, sp, bs, k in SMALL:
    layer, x, dy, plan, T, pad = setup(cin, cout, sp, bs, k)
    wshape = tuple(layer.conv.weight.shape)
    ref = torch.nn.grad.conv3d_weight(x.float().contiguous(), wshape, dy.float().contiguous(), stride=1, padding=pad)
    ops.set_wgrad_tma(0)
    old = run(x, dy, plan, T, cin, cout, wshape)
    ops.set_wgrad_tma(mode)
    ops.trace_start()
    new = run(x, dy, plan, T, cin, cout, wshape)
    import csv, tempfile
    with tempfile.TemporaryDirectory() as td:
        ops.trace_dump(os.path.join(td, "t.csv"))
        kern = [r["kernel"] for r in csv.DictReader(open(os.path.join(td, "t.csv")))]
    torch.cuda.synchronize()
    e_new, e_old = rel(new, ref), rel(old, ref)
    good = e_new < 5e-3
    ok_all = ok_all and good
    print(f"mode {mode} {cin}->{cout} {sp} x{bs} k={k}: kernels {kern}  rel(new, cudnn fp32) {e_new:.2e}  rel(old, cudnn) {e_old:.2e}  "
          f"rel(new, old) {rel(new, old):.2e}  {'OK' if good else 'WRONG'}", flush=True)
print(f"mode {mode}: {'ALL CORRECT' if ok_all else 'MISMATCH'}", flush=True)
if ok_all and not quick:
    for cin, cout, sp, bs, k in BIG:
        layer, x, dy, plan, T, pad = setup(cin, cout, sp, bs, k)
        wshape = tuple(layer.conv.weight.shape)
        dw = torch.zeros(wshape, dtype=torch.float32, device=dev)
        fn = lambda: [ops.conv_wgrad(dy, cout, x, cin, g, dw, cin * T, T, 1, cout, cin) for g in plan.wgrad]
        ops.set_wgrad_tma(0)
        t_old = timeit(fn)
        ops.set_wgrad_tma(mode)
        t_new = timeit(fn)
        ops.set_wgrad_tma(0)
        old = run(x, dy, plan, T, cin, cout, wshape)
        ops.set_wgrad_tma(mode)
        new = run(x, dy, plan, T, cin, cout, wshape)
        fl = 2.0 * T * cin * cout * bs * sp[0] * sp[1] * sp[2]
        print(f"mode {mode} {cin}->{cout} {sp} x{bs}: cp.async {t_old * 1e3:.1f} us ({fl / t_old / 1e9:.0f} TFLOP/s)  TMA {t_new * 1e3:.1f} us "
              f"({fl / t_new / 1e9:.0f} TFLOP/s)  rel(new, old) {rel(new, old):.2e}", flush=True)
