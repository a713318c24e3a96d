"""Timing experiments on the TMA-fed wgrad kernel: python scripts/time_wgrad_tma.py cin cout size batch mode [mode ...]
(mode bits: 1 on, 8 no epilogue atomics, 16 no MMAs, 32 no TMA loads; 0 = the cp.async kernel)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nndetection_b200.arch import conv_ops as ops
from nndetection_b200.arch.conv import ConvInstanceRelu

cin, cout, size, bs = (int(v) for v in sys.argv[1:5])
modes = [int(v) for v in sys.argv[5:]] or [1]
reps = int(os.environ.get("REPS", "10"))
dev = torch.device("cuda")
layer = ConvInstanceRelu(3, cin, cout, kernel_size=3, stride=1, padding=1).to(dev)
x = torch.randn(bs, cin, size, size, size, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
dy = torch.randn(bs, cout, size, size, size, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
plan = layer.plan(bs, (size,) * 3)
dw = torch.zeros_like(layer.conv.weight)
fl = 2.0 * 27 * cin * cout * bs * size ** 3
for m in modes:
    ops.set_wgrad_tma(m)
    fn = lambda: ops.conv_wgrad(dy, cout, x, cin, plan.wgrad[0], dw, cin * 27, 27, 1, cout, cin)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"wgrad {cin}->{cout} @{size}^3 x{bs} mode {m}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
