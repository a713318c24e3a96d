"""The per-launch convolution trace (nnd_conv_trace / nnd_conv_trace_dump, a profiling aid behind `bench.py --trace-layers`): one row
per convolution-family launch of a train step, with the kernel the dispatch chose, the layer geometry and a positive duration.
First B200 run: round-1 driver run (XPASS); strict since round 2."""
import csv

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trace_lists_every_convolution_launch_of_a_train_step(tmp_path):
    from nndetection_b200.arch import conv_ops as ops
    from nndetection_b200.configs import make_plan, synth_batch
    from nndetection_b200.ptmodule import RetinaUNetV001
    from nndetection_b200.training import Trainer
    arch, anc, patch, bs = make_plan("toy")
    torch.manual_seed(0)
    net = RetinaUNetV001.from_config_plan(None, arch, anc).cuda()
    images, targets = synth_batch(patch, bs, arch["in_channels"], arch["classifier_classes"], 3)
    tg = {"target_boxes": [b.cuda() for b in targets["target_boxes"]], "target_classes": [c.cuda() for c in targets["target_classes"]],
          "target_seg": targets["target_seg"].cuda()}
    tr = Trainer(net)
    tr.train_step(images.cuda(), tg)                       # untraced warm-up
    ops.trace_start()
    tr.train_step(images.cuda(), tg)
    n = ops.trace_dump(str(tmp_path / "layers.csv"))
    rows = list(csv.DictReader(open(tmp_path / "layers.csv")))
    assert n == len(rows) and n > 60                        # ~30 conv layers x (fprop + dgrad + wgrad)
    kinds = {r["kind"] for r in rows}
    assert {"fprop", "wgrad", "first_fprop", "first_wgrad"} <= kinds
    assert all(float(r["ms"]) > 0 for r in rows) and all(float(r["gflop"]) > 0 for r in rows)
    assert any(r["kernel"].startswith("conv_tc") for r in rows) and any(r["kernel"].startswith("wgrad_") for r in rows)
    tr.train_step(images.cuda(), tg)                       # tracing is off again: the table must not grow
    assert ops.trace_dump(str(tmp_path / "again.csv")) == n
