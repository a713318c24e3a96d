"""CPU (gloo, world size 2): the data-parallel plumbing of the train step -- flat-buffer layout, broadcast of rank-0
weights, gradient all-reduce + mean -- without a GPU.  The CUDA kernels are exercised by the `-m gpu` tests; here a
stand-in SGD (same formula as csrc/misc.cu:sgd_kernel) checks that two ranks end up with identical parameters equal to a
single-process run on the concatenated batch."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


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
    port = 29500 + os.getpid() % 1000
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
