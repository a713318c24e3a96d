"""Small launch loops for `ncu --set full -k regex:<kernel> -s <skip> -c 1` captures of single kernels at BASELINE-config shapes:
    python scripts/ncu_targets.py pw    cin cout size batch      # 1x1x1 lateral (conv_pw / conv_igemm), forward
    python scripts/ncu_targets.py up    cin cout size batch      # kernel == stride 2 up-convolution + lateral add, forward (size = input)
    python scripts/ncu_targets.py block cin cout size batch      # 3x3x3 conv + instance norm + ReLU, forward + backward (conv, wgrad, norm kernels)
    python scripts/ncu_targets.py block2 cin cout size batch     # the same with stride 2 (size = input size)
    python scripts/ncu_targets.py nms   n                        # one nndet._C.nms call on the SURVEY 8d stress boxes (mask + scan kernels)
env NND_PW=0: pointwise forms on the mma.sync gather kernel (A/B)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nndetection_b200.arch import conv_ops as ops  # noqa: E402
from nndetection_b200.arch.conv import ConvInstanceRelu  # noqa: E402

mode = sys.argv[1]
ops.set_pointwise_tma(os.environ.get("NND_PW", "1") != "0")
dev = torch.device("cuda")
REPS = 3
if mode == "nms":
    from nndetection_b200 import _C
    import bench
    n = int(sys.argv[2])
    boxes, scores = bench._nms_stress(n, dev)
    for _ in range(REPS):
        keep = _C.nms(boxes, scores, 0.1)
    torch.cuda.synchronize()
    print("nms", n, "kept", keep.numel())
    sys.exit(0)
cin, cout, size, bs = (int(a) for a in sys.argv[2:6])
rnd = lambda c, s: torch.randn(bs, c, s, s, s, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
if mode == "pw":
    layer = ConvInstanceRelu(3, cin, cout, kernel_size=1, stride=1, padding=0, add_norm=False, add_act=False).to(dev)
    x = rnd(cin, size)
    with torch.no_grad():
        for _ in range(REPS):
            y = layer(x)
elif mode == "up":
    layer = ConvInstanceRelu(3, cin, cout, kernel_size=2, stride=2, transposed=True, add_norm=False, add_act=False).to(dev)
    x, lat = rnd(cin, size), rnd(cout, 2 * size)
    with torch.no_grad():
        for _ in range(REPS):
            y = layer(x, residual=lat)
elif mode in ("block", "block2"):              # block2: the stride-2 form (size = input size)
    stride = 2 if mode == "block2" else 1
    layer = ConvInstanceRelu(3, cin, cout, kernel_size=3, stride=stride, padding=1).to(dev)
    x = rnd(cin, size).requires_grad_(True)
    gy = rnd(cout, size // stride)
    for _ in range(REPS):
        layer.zero_grad(set_to_none=True)
        layer(x).backward(gy)
torch.cuda.synchronize()
print(mode, cin, cout, size, bs, "done")
