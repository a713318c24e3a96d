"""CPU (gloo, world size 2): the data-parallel plumbing of the train step -- flat-buffer layout, broadcast of rank-0
weights, gradient all-reduce + mean -- without a GPU.  The CUDA kernels are exercised by the `-m gpu` tests; here a
stand-in SGD (same formula as csrc/misc.cu:sgd_kernel) checks that two ranks end up with identical parameters equal to a
single-process run on the concatenated batch."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _sgd_reference(p, g, mom, lr, momentum, wd, nesterov, first):
    g = g + wd * p
    b = g.clone() if first else momentum * mom + g
    step = g + momentum * b if nesterov else b
    return p - lr * step, b


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nndetection_b200.arch.conv import NormParams
    from nndetection_b200.training import FlatParameters, poly_lr
    torch.manual_seed(100 + rank)                      # different init per rank: broadcast must fix it
    model = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 4))
    model.add_module("norm", NormParams(4, 1))
    fp = FlatParameters(model)
    assert fp.n == sum(p.numel() for p in model.parameters()) and fp.n_decay == fp.n - 8
    for p in model.parameters():                       # parameters and grads are views of the flat buffers
        assert p.data_ptr() >= fp.flat.data_ptr() and p.grad.data_ptr() >= fp.grad.data_ptr()
    dist.broadcast(fp.flat, src=0)
    torch.manual_seed(7)
    x = torch.randn(8, 8)
    y = torch.randn(8, 4)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    fp.zero_grad()
    out = model[2](model[1](model[0](xs))) * model.norm.weight + model.norm.bias
    ((out - ys) ** 2).mean().backward()
    dist.all_reduce(fp.grad)
    lr = poly_lr(0, 0.01, 0, 1e-6, 0.9, 100)
    new, _ = _sgd_reference(fp.flat, fp.grad / world, fp.mom, lr, 0.9, 3e-5, True, True)
    new[fp.n_decay:] = _sgd_reference(fp.flat[fp.n_decay:], fp.grad[fp.n_decay:] / world, fp.mom[fp.n_decay:], lr, 0.9, 0.0, True, True)[0]
    ret[rank] = new.clone()
    dist.destroy_process_group()


def test_two_rank_gradient_mean_matches_single_process():
    mp.set_start_method("spawn", force=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert torch.allclose(ret[0], ret[1], atol=0, rtol=0)          # ranks stay bit-identical
    # single process on the full batch (mean over 8 samples == mean of the two 4-sample means)
    from nndetection_b200.arch.conv import NormParams
    torch.manual_seed(100)
    model = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 4))
    model.add_module("norm", NormParams(4, 1))
    params = [p for p in model.parameters()]
    torch.manual_seed(7)
    x, y = torch.randn(8, 8), torch.randn(8, 4)
    out = model[2](model[1](model[0](x))) * model.norm.weight + model.norm.bias
    ((out - y) ** 2).mean().backward()
    flat_p = torch.cat([p.data.reshape(-1) for p in params])
    flat_g = torch.cat([p.grad.reshape(-1) for p in params])
    from nndetection_b200.training import poly_lr
    lr = poly_lr(0, 0.01, 0, 1e-6, 0.9, 100)
    ref, _ = _sgd_reference(flat_p, flat_g, torch.zeros_like(flat_p), lr, 0.9, 3e-5, True, True)
    nd = flat_p.numel() - 8
    ref[nd:] = _sgd_reference(flat_p[nd:], flat_g[nd:], torch.zeros(8), lr, 0.9, 0.0, True, True)[0]
    assert torch.allclose(ret[0], ref, rtol=1e-5, atol=1e-7)


def test_lr_schedule_matches_reference_scheduler():
    """nndetection_b200.training.poly_lr vs the EXECUTED reference `LinearWarmupPolyLR` (one scheduler step per optimizer step),
    tests/golden/lr.npz from scripts/gen_golden.py lr: identical doubles for every step of two schedules."""
    import numpy as np
    import tutil as util
    from nndetection_b200.training import poly_lr
    g = util.golden("lr")
    for i, (lr0, warm, wlr, gamma, n) in enumerate(g["cfgs"].tolist()):
        ref = g[f"lrs{i}"]
        mine = np.asarray([poly_lr(s, lr0, int(warm), wlr, gamma, int(n)) for s in range(len(ref))])
        assert np.array_equal(mine, ref), i


# ------------------------------------------------------------------ opt-in gradient buckets (training.GradientBuckets)
class _DirectFn(torch.autograd.Function):
    """CPU stand-in for arch/conv.py:_ConvBlockFn: y = x * w (per channel); the weight gradient is ADDED straight into `w.grad` (a view
    of the flat buffer) and autograd gets None for it -- exactly the contract of the direct-accumulation kernels."""

    @staticmethod
    def forward(ctx, x, weight, layer):
        ctx.layer = layer
        w = weight.detach().view(-1)[:x.shape[1]]
        ctx.save_for_backward(x, w)
        return x * w

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        ctx.layer.conv.weight.grad.view(-1)[:x.shape[1]].add_((dy * x).sum(0))
        return dy * w, None, None


def _bucket_model():
    from nndetection_b200.arch.conv import BaseConvNormAct, notify_grad_observer

    class Direct(BaseConvNormAct):
        def forward(self, x):
            z = _DirectFn.apply(x, self.conv.weight, self)
            notify_grad_observer(self, z)
            return z

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = Direct(3, 32, 32, None, None, 1)          # 1x1x1 conv holder: weight [32, 32, 1, 1, 1] + bias
            self.shared = Direct(3, 32, 32, None, None, 1)
            self.lin = nn.Linear(32, 32)
            self.b = Direct(3, 32, 32, None, None, 1)
            self.unused = nn.Parameter(torch.zeros(700))

        def forward(self, x):
            x = self.a(x)
            x = self.shared(x)
            x = torch.tanh(self.lin(x))
            x = self.shared(x)                                  # second use of the same parameters
            return self.b(x)
    return Net()


def _bucket_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nndetection_b200.arch.conv import set_grad_observer
    from nndetection_b200.training import FlatParameters, GradientBuckets
    torch.manual_seed(5)
    model = _bucket_model()
    fp = FlatParameters(model)
    torch.manual_seed(50 + rank)
    x = torch.randn(6, 32)

    def backward_once():
        fp.zero_grad()
        (model(x) ** 2).sum().backward()

    backward_once()                                             # reference: one all-reduce after backward
    ref = fp.grad.clone()
    dist.all_reduce(ref)
    gb = GradientBuckets(model, fp, bucket_mb=1000 * 4 / 2 ** 20)          # 1000 floats per bucket -> 5 buckets
    set_grad_observer(gb)
    fired = []
    orig = gb._launch
    gb._launch = lambda b: (fired.append((b, gb.in_backward)), orig(b))[1]
    for _ in range(2):                                          # two steps: per-step state resets
        gb.begin()
        fired.clear()
        gb.in_backward = True
        backward_once()
        gb.in_backward = False
        gb.finish()
        assert torch.equal(fp.grad, ref)
    set_grad_observer(None)
    ret[rank] = (list(gb.order), list(fired), len(gb.bounds))
    dist.destroy_process_group()


def test_gradient_buckets_overlap_gives_the_single_all_reduce_result():
    """Opt-in bucketed exchange: same gradients as one all-reduce, identical issue order on both ranks, buckets of layers whose
    backward is over are exchanged DURING backward (the shared layer only after its second use), the rest in finish()."""
    mp.set_start_method("spawn", force=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.Process(target=_bucket_worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    order0, fired0, nb = ret[0]
    assert ret[1][0] == order0 and sorted(order0) == list(range(nb)) and nb >= 4
    during = [b for b, inside in fired0 if inside]
    assert len(during) >= 2 and any(not inside for _, inside in fired0)       # the bucket with the unused parameter waits for finish()
