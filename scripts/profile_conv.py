"""Run single conv layers in isolation (for `ncu --set full -k regex:conv_`):
    python scripts/profile_conv.py [cin cout size batch [fprop|wgrad [wgrad_tc_mode]]]
env: NND_STREAM=mode,issuers  NND_TC_RING=n  NND_STRIDE=2 (3x3x3 stride-2 layer, `size` = input size)  NND_S2=1 (opt-in tcgen05 kernels
for the strided forms: conv_tc S2 / conv_wgrad_tc SW=2)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nndetection_b200.arch import conv_ops as ops
from nndetection_b200.arch.conv import ConvInstanceRelu

cin, cout, size, bs = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 32, 128, 4)))
mode = sys.argv[5] if len(sys.argv) > 5 else "fprop"
if len(sys.argv) > 6:
    from nndetection_b200 import _lib as L
    from ctypes import c_int
    L.lib().nnd_conv_set_wgrad_tc(c_int(int(sys.argv[6])))
if os.environ.get("NND_STREAM"):                      # "mode,issuers" e.g. NND_STREAM=0,2 disables the streaming kernel
    m_, i_ = (int(v) for v in os.environ["NND_STREAM"].split(","))
    ops.set_stream_path(m_, i_)
if os.environ.get("NND_TC_RING"):
    from nndetection_b200 import _lib as L2
    from ctypes import c_int as _ci
    L2.lib().nnd_conv_set_tc_ring(_ci(int(os.environ["NND_TC_RING"])))
if os.environ.get("NND_TCS_MAP"):
    from nndetection_b200 import _lib as L3
    from ctypes import c_int as _ci3
    L3.lib().nnd_conv_set_tcs_map(_ci3(int(os.environ["NND_TCS_MAP"])))
stride = int(os.environ.get("NND_STRIDE", "1"))
if os.environ.get("NND_S2") == "0":          # A/B: strided forms on the mma.sync kernels
    ops.set_gather_strided_tc(False)
    ops.set_wgrad_strided_tc(False)
dev = torch.device("cuda")
layer = ConvInstanceRelu(3, cin, cout, kernel_size=3, stride=stride, padding=1).to(dev)
x = torch.randn(bs, cin, size, size, size, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
plan = layer.plan(bs, (size,) * 3)
wp, wb = layer.packed()
y = ops.empty_cl(bs, cout, plan.out_sp, device=dev)
st = torch.zeros((2, bs, cout), dtype=torch.float32, device=dev)
dw = torch.zeros_like(layer.conv.weight)
if mode != "fprop":
    y.normal_()                      # stands in for dy
def run():
    if mode == "fprop":
        return ops.conv_gather(x, wp, plan.fprop[0], y, cout, cout, stat_sum=st[0], stat_sq=st[1])
    ops.conv_wgrad(y, cout, x, cin, plan.wgrad[0], dw, cin * 27, 27, 1, cout, cin)
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
fl = 2.0 * 27 * cin * cout * bs * plan.out_sp[0] * plan.out_sp[1] * plan.out_sp[2]
print(f"{mode} {cin}->{cout} @{size}^3 x{bs} stride={stride} s2={os.environ.get('NND_S2', '0')} stream={os.environ.get('NND_STREAM', 'default')}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  (kernel code {run()})")
