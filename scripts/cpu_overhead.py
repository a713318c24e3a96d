"""How long does the host need to ENQUEUE one train step (no sync) vs how long the GPU needs to execute it?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nndetection_b200.configs import make_plan, synth_batch
from nndetection_b200.ptmodule import RetinaUNetV001
from nndetection_b200.training import Trainer
ev = "--no-eval" not in sys.argv
arch, anc, patch, bs = make_plan("luna")
dev = torch.device("cuda")
net = RetinaUNetV001.from_config_plan(None, arch, anc).to(dev)
tr = Trainer(net)
im, tg = synth_batch(patch, bs, 1, 1, 1)
im = im.to(dev); tg = {"target_boxes": [b.to(dev) for b in tg["target_boxes"]], "target_classes": [c.to(dev) for c in tg["target_classes"]], "target_seg": tg["target_seg"].to(dev)}
for _ in range(3):
    tr.train_step(im, tg, evaluation=ev)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    tr.model.train(); tr.fp.zero_grad(); tr.model.defer_prediction_sync = True
    losses, pred = tr.model.train_step(im, tg, evaluation=ev, batch_num=0)
    t1 = time.perf_counter()
    sum(losses.values()).backward()
    t2 = time.perf_counter()
    tr.optimizer_step()
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"enqueue fwd+loss+post {1e3*(t1-t0):.1f} ms | bwd {1e3*(t2-t1):.1f} ms | opt {1e3*(t3-t2):.1f} ms | wait for GPU {1e3*(t4-t3):.1f} ms | total {1e3*(t4-t0):.1f} ms")
