"""The fused optimizer step (nnd_sgd_step, csrc/misc.cu) against torch.optim.SGD as the reference configures it
(nndet/ptmodule/retinaunet/base.py:300-336: momentum 0.9, nesterov, weight decay 3e-5 on everything except norm parameters,
nndet/training/optimizer/utils.py:30-50) over several steps with changing learning rates -- including the first step's momentum-buffer
initialisation and the 1 / world gradient scale of the data-parallel path.  fp32; differences = fused multiply-adds only (2e-6 relative + 5e-7)."""
from ctypes import c_float, c_int, c_longlong

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nesterov,world", [(True, 1), (True, 4), (False, 1)])
def test_sgd_step_matches_torch_optim_sgd(nesterov, world):
    from nndetection_b200 import _lib as L
    from nndetection_b200.training import poly_lr
    n, n_decay = 100_003, 90_001
    g = torch.Generator().manual_seed(11)
    p0 = torch.randn(n, generator=g)
    flat = p0.clone().cuda()
    mom = torch.zeros(n, device="cuda")
    pd = torch.nn.Parameter(p0[:n_decay].clone().cuda())
    pn = torch.nn.Parameter(p0[n_decay:].clone().cuda())
    opt = torch.optim.SGD([{"params": [pd], "weight_decay": 3e-5}, {"params": [pn], "weight_decay": 0.0}], lr=0.01, momentum=0.9,
                          nesterov=nesterov)
    lib = L.lib()
    for step in range(4):
        lr = poly_lr(step, 0.01, 2, 1e-6, 0.9, 10)
        grad_sum = torch.randn(n, generator=g).cuda() * world          # what the all-reduce (SUM) leaves in the flat buffer
        for grp in opt.param_groups:
            grp["lr"] = lr
        pd.grad, pn.grad = (grad_sum[:n_decay] / world).clone(), (grad_sum[n_decay:] / world).clone()
        opt.step()
        L.check(lib.nnd_sgd_step(L.ptr(flat), L.ptr(grad_sum), L.ptr(mom), c_longlong(n), c_longlong(n_decay), c_float(lr), c_float(0.9),
                                 c_float(3e-5), c_int(1 if nesterov else 0), c_int(1 if step == 0 else 0), c_float(1.0 / world),
                                 L.stream_ptr()), "nnd_sgd_step")
        ref = torch.cat([pd.detach(), pn.detach()])
        assert torch.allclose(flat, ref, rtol=2e-6, atol=5e-7), (step, float((flat - ref).abs().max()))


def test_parameters_without_a_gradient_are_left_alone_like_torch_sgd_does():
    """torch.optim.SGD skips parameters whose `.grad is None` (no weight decay, no momentum).  In the reference that is e.g. decoder
    `out.P1` of the LUNA plan (its output feeds neither head nor segmenter).  The flat-buffer optimizer finds those layers after the first
    backward pass (layers whose backward never ran) and skips their element ranges (nnd_sgd_step_skip)."""
    from nndetection_b200.configs import make_plan, synth_batch
    from nndetection_b200.ptmodule import RetinaUNetV001
    from nndetection_b200.training import Trainer
    arch, anc, patch, bs = make_plan("tiny")
    arch = dict(arch, decoder_levels=(2,), fpn_channels=128, head_channels=128)   # P0 (32 ch) -> segmenter, P1 (64) -> nothing, P2 (128) -> head
    anc = {k: v[:1] for k, v in anc.items()}
    torch.manual_seed(1)
    net = RetinaUNetV001.from_config_plan(None, arch, anc).cuda()
    before = {k: v.detach().clone() for k, v in net.named_parameters()}
    trainer = Trainer(net, initial_lr=0.05, warm_iterations=1, warm_lr=0.05, num_iterations=10, weight_decay=1e-2)
    for step in range(2):
        images, targets = synth_batch(patch, bs, 1, arch["classifier_classes"], 50 + step)
        tg = {"target_boxes": [b.cuda() for b in targets["target_boxes"]], "target_classes": [c.cuda() for c in targets["target_classes"]],
              "target_seg": targets["target_seg"].cuda()}
        trainer.train_step(images.cuda(), tg, evaluation=False)
    assert sorted(trainer.fp.unused) == ["decoder.out.P1.0.conv.bias", "decoder.out.P1.0.conv.weight"]
    after = dict(net.named_parameters())
    for k in trainer.fp.unused:
        assert torch.equal(after[k].detach(), before[k]), k      # bit-identical: neither decayed nor moved
    changed = [k for k in before if k not in trainer.fp.unused and not torch.equal(after[k].detach(), before[k])]
    assert len(changed) == len(before) - 2                        # every other parameter was updated
