"""Named workload configurations (BASELINE.json `configs`) and the synthetic batch generator."""
import math
from typing import Sequence

import torch


def make_plan(name: str):
    """Architecture/anchor plans shaped like BASELINE.json's configs."""
    if name == "toy":        # config 1: 32x64x64, 2 pools, levels (1, 2), C=2
        arch = dict(dim=3, in_channels=1, classifier_classes=2, seg_classes=2, start_channels=32, fpn_channels=128,
                    head_channels=128, max_channels=320, decoder_levels=(1, 2),
                    conv_kernels=[(3, 3, 3)] * 3, strides=[(2, 2, 2)] * 2)
        sizes = [(4, 8, 16), (8, 16, 32)]
        patch, bs = (32, 64, 64), 2
    elif name == "tiny":     # unit-test size: 16x32x32, levels (1, 2)
        arch = dict(dim=3, in_channels=1, classifier_classes=2, seg_classes=2, start_channels=32, fpn_channels=64,
                    head_channels=64, max_channels=320, decoder_levels=(1, 2),
                    conv_kernels=[(3, 3, 3)] * 3, strides=[(2, 2, 2)] * 2)
        sizes = [(4, 8, 16), (8, 16, 32)]
        patch, bs = (16, 32, 32), 2
    elif name in ("luna", "adam", "infer160"):
        cin, ncls = (2, 3) if name == "adam" else (1, 1)
        arch = dict(dim=3, in_channels=cin, classifier_classes=ncls, seg_classes=ncls, start_channels=32,
                    fpn_channels=128, head_channels=128, max_channels=320, decoder_levels=(2, 3, 4, 5),
                    conv_kernels=[(3, 3, 3)] * 6, strides=[(2, 2, 2)] * 5)
        sizes = [(4, 8, 16), (8, 16, 32), (16, 32, 64), (32, 64, 128)]
        patch, bs = ((160, 160, 160) if name == "infer160" else (128, 128, 128)), 4
    elif name == "lidc":     # config 3: 96x192x192, first stride (1, 2, 2)
        arch = dict(dim=3, in_channels=1, classifier_classes=1, seg_classes=1, start_channels=32, fpn_channels=128,
                    head_channels=128, max_channels=320, decoder_levels=(2, 3, 4, 5),
                    conv_kernels=[(3, 3, 3)] * 6, strides=[(1, 2, 2)] + [(2, 2, 2)] * 4)
        sizes = [(4, 8, 16), (8, 16, 32), (16, 32, 64), (32, 64, 128)]
        patch, bs = (96, 192, 192), 4
    else:
        raise KeyError(name)
    anchors = dict(width=sizes, height=sizes, depth=sizes)
    return arch, anchors, patch, bs


def synth_batch(patch: Sequence[int], bs: int, cin: int, ncls: int, seed: int, max_gt: int = 4):
    """Synthetic patches + cuboid targets: uniform noise images, 0..max_gt boxes of side U[6,32)
    (scaled to the patch), at least one empty-GT image when bs >= 2; seg = union of boxes.
    GT corners get an irrational-ish fractional offset so centre-distance ties at the ATSS k-boundary
    do not occur (SURVEY 7 hard part 2)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(bs, cin, *patch, generator=g)
    tb, tc = [], []
    seg = torch.zeros(bs, *patch)
    for i in range(bs):
        n = 0 if (bs >= 2 and i == bs - 1) else int(torch.randint(1, max_gt + 1, (1,), generator=g))
        boxes = []
        for _ in range(n):
            lo, hi = [], []
            for ax in range(3):
                side = min(float(torch.randint(6, 32, (1,), generator=g)), patch[ax] - 2.0)
                start = float(torch.rand(1, generator=g)) * (patch[ax] - side - 1)
                start = math.floor(start) + 0.3183098861 + 0.01 * ax
                lo.append(start); hi.append(start + side + 0.1415926)
            boxes.append([lo[0], lo[1], hi[0], hi[1], lo[2], hi[2]])
            seg[i, int(lo[0]):int(hi[0]) + 1, int(lo[1]):int(hi[1]) + 1, int(lo[2]):int(hi[2]) + 1] = 1
        tb.append(torch.tensor(boxes, dtype=torch.float32).reshape(-1, 6))
        tc.append(torch.randint(0, ncls, (n,), generator=g))
    return images, dict(target_boxes=tb, target_classes=tc, target_seg=seg)
