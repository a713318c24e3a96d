"""Training-step plumbing for the hot path: flat parameter / gradient buffers, one fused SGD kernel, the reference's
LR schedule, and patch-level data parallelism (one process per GPU, NCCL all-reduce on the gradient buffer only).

Reference semantics kept: torch.optim.SGD(nesterov, momentum 0.9, weight decay 3e-5 on everything except norm
parameters) -- nndet/ptmodule/retinaunet/base.py:300-336, nndet/training/optimizer/utils.py:30-50; LinearWarmupPolyLR --
nndet/training/learning_rate.py; DDP is what PL would do for `gpus > 1` (scripts/train.py:265-272): per-GPU batch,
gradient mean over ranks, no other collective (InstanceNorm / GroupNorm are per sample).
"""
from ctypes import c_float, c_int, c_longlong
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib as L
from .arch.conv import NormParams, bump_weights_epoch


class FlatParameters:
    """Re-homes every parameter of `model` into ONE fp32 buffer ([decayed | norm (no decay)]) and every gradient
    into a matching buffer, so the optimizer is one launch and the DDP exchange one NCCL call."""

    def __init__(self, model: nn.Module):
        norm_ids = set()
        for m in model.modules():
            if isinstance(m, NormParams):
                for p in m.parameters(recurse=False):
                    norm_ids.add(id(p))
        params = [p for p in model.parameters() if p.requires_grad]
        decay = [p for p in params if id(p) not in norm_ids]
        no_decay = [p for p in params if id(p) in norm_ids]
        self.params = decay + no_decay
        self.n_decay = sum(p.numel() for p in decay)
        self.n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.mom = torch.zeros(self.n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)         # autograd accumulates in place into the flat buffer
            p._nnd_direct_grad = True                             # ... and the conv / norm kernels add straight into it
            off += k
        self.first_step = True
        self.skip, self.n_skip, self.unused = None, 0, None
        ids = {id(p): n for n, p in model.named_parameters()}
        self.names = [(ids.get(id(p), "?"), p) for p in self.params]

    def zero_grad(self):
        self.grad.zero_()

    def find_unused(self, model: nn.Module):
        """Parameters that receive no gradient at all (e.g. decoder `out.P1` under the LUNA plan: its output feeds neither head nor
        segmenter) keep `grad is None` in the reference, and torch.optim.SGD skips them -- no weight decay, no momentum.  Here every
        parameter owns a slice of the flat gradient buffer, so "no gradient" is read off the layers instead: each layer's backward
        function marks the layer (`_nnd_grad_seen`, arch/conv.py / arch/net.py) when autograd runs it.  After the first backward pass the
        parameters of layers of THIS package that were never marked are handed to the optimizer kernel as element ranges to leave
        alone (the graph is static).  Parameters of foreign modules are always updated."""
        from .arch.conv import BaseConvNormAct
        from .arch.net import DetectionHeadHNMNative, DiCESegmenterFgBg
        unused = set()
        for m in model.modules():
            if getattr(m, "_nnd_grad_seen", False):
                continue
            if isinstance(m, BaseConvNormAct):
                owned = list(m.conv.parameters()) + (list(m.norm.parameters()) if m.norm is not None else [])
            elif isinstance(m, DetectionHeadHNMNative):
                owned = list(m.classifier.conv_out.parameters()) + list(m.regressor.conv_out.parameters()) + \
                    ([s.scale for s in m.regressor.scales] if m.n_scales else [])
            elif isinstance(m, DiCESegmenterFgBg):
                owned = list(m.conv_out.parameters())
            else:
                continue
            unused.update(id(p) for p in owned)
        ranges, off = [], 0
        for p in self.params:
            k = p.numel()
            if id(p) in unused:
                if ranges and ranges[-1][1] == off:
                    ranges[-1][1] = off + k
                else:
                    ranges.append([off, off + k])
            off += k
        self.unused = [n for n, p in self.names if id(p) in unused]
        self.n_skip = len(ranges)
        self.skip = torch.tensor(ranges, dtype=torch.int64, device=self.flat.device).reshape(-1) if ranges else None

    def check_homed(self):
        """`model.to()` / `.float()` / `.cuda()` after construction re-home `p.data`: the optimizer kernel would then update a buffer the
        model no longer reads.  One pointer comparison per parameter (host only), once per step."""
        base, end = self.flat.data_ptr(), self.flat.data_ptr() + 4 * self.n
        for p in self.params:
            if not (base <= p.data_ptr() < end) or p.grad is None or not (self.grad.data_ptr() <= p.grad.data_ptr() < self.grad.data_ptr() + 4 * self.n):
                raise RuntimeError("a parameter (or its .grad) no longer lives in the Trainer's flat buffer: build the Trainer after "
                                   "the last model.to()/.float()/.cuda(), and do not replace .grad tensors (zero them in place)")


class GradientBuckets:
    """Overlap of the data-parallel gradient exchange with the backward pass (default since round 2): the flat gradient buffer is cut into
    contiguous buckets of ~`bucket_mb` (parameter order = forward order, so they complete roughly in reverse); a bucket is all-reduced
    asynchronously as soon as every parameter in it has its final gradient of this step, `finish()` launches what is left (in index
    order) and waits for everything.  Readiness: layers that add their gradients straight into the flat buffer (arch/conv.py) report
    through autograd post-hooks on their nodes (one per use: shared head layers run once per pyramid level); every other parameter
    through `register_post_accumulate_grad_hook`.  All ranks run the same graph, so buckets complete -- and their collectives are
    issued -- in the same order everywhere.  Validated on gloo (tests/test_ddp_cpu.py: same gradients as the single all-reduce) and over
    NCCL on B200s (scripts/ddp_bucket_check.py at 2 GPUs; bench at 2 / 8 GPUs: 26.99 vs 27.08 ms and 27.20 vs 27.39 ms per step against
    `Trainer(bucket_mb=None)`, the single all-reduce after backward)."""

    def __init__(self, model: nn.Module, fp: FlatParameters, bucket_mb: float = 25.0):
        from .arch.conv import BaseConvNormAct
        self.fp = fp
        off, self.offsets = 0, {}
        for p in fp.params:
            self.offsets[id(p)] = (off, off + p.numel())
            off += p.numel()
        per = max(1, int(bucket_mb * 2 ** 20 / 4))
        self.bounds = [(lo, min(lo + per, fp.n)) for lo in range(0, fp.n, per)]
        bucket_of = lambda lo, hi: range(lo // per, (hi - 1) // per + 1)
        # parameters owned by direct-accumulation layers -> layer; the rest -> hook on the parameter
        self.layer_buckets: Dict[int, List[int]] = {}
        direct_ids = set()
        for m in model.modules():
            if isinstance(m, BaseConvNormAct):
                ps = [p for p in (m.conv.weight, m.conv.bias, getattr(m.norm, "weight", None), getattr(m.norm, "bias", None))
                      if p is not None and id(p) in self.offsets and getattr(p, "_nnd_direct_grad", False)]
                bs = sorted({b for p in ps for b in bucket_of(*self.offsets[id(p)])})
                self.layer_buckets[id(m)] = bs
                direct_ids.update(id(p) for p in ps)
        self.param_buckets = {id(p): list(bucket_of(*self.offsets[id(p)])) for p in fp.params if id(p) not in direct_ids}
        self._handles = [p.register_post_accumulate_grad_hook(lambda p_, _s=self: _s._param_done(p_))
                         for p in fp.params if id(p) not in direct_ids]
        self.static_pending = [0] * len(self.bounds)            # parameters reporting through their own hook
        for bs in self.param_buckets.values():
            for b in bs:
                self.static_pending[b] += 1
        self.begin()

    def begin(self):
        """Call before the forward pass of a step."""
        self.pending = list(self.static_pending)
        self.launched = [False] * len(self.bounds)
        self.works, self.order, self.in_backward = [], [], False
        self.main_stream = torch.cuda.current_stream(self.fp.grad.device) if self.fp.grad.is_cuda else None

    # ---- observer protocol of arch/conv.py
    def layer_used(self, layer):
        for b in self.layer_buckets.get(id(layer), ()):
            self.pending[b] += 1

    def layer_done(self, layer):
        for b in self.layer_buckets.get(id(layer), ()):
            self._dec(b)

    def _param_done(self, p):
        for b in self.param_buckets.get(id(p), ()):
            self._dec(b)

    def _dec(self, b):
        self.pending[b] -= 1
        if self.pending[b] == 0 and not self.launched[b]:
            self._launch(b)

    def _launch(self, b):
        lo, hi = self.bounds[b]
        self.launched[b] = True
        self.order.append(b)
        # The bucket's gradients were written by kernels on the stream each layer's forward ran on (the coarse pyramid levels and the
        # segmentation branch have side streams, arch/net.py / core/retina.py); the collective orders itself after the CURRENT stream
        # only, so join the side streams first.
        if self.fp.grad.is_cuda:
            from .arch.net import _LEVEL_STREAMS
            from .core.retina import _SIDE_STREAMS
            cur = torch.cuda.current_stream(self.fp.grad.device)
            for st in list(_LEVEL_STREAMS.values()) + list(_SIDE_STREAMS.values()) + [self.main_stream]:
                if st is not None and st.device == self.fp.grad.device and st != cur:
                    cur.wait_stream(st)
        self.works.append(dist.all_reduce(self.fp.grad[lo:hi], async_op=True))

    def finish(self):
        """After backward: exchange the buckets that never completed through hooks (unused parameters), wait for all of them."""
        for b in range(len(self.bounds)):
            if not self.launched[b]:
                self._launch(b)
        for w in self.works:
            w.wait()


def poly_lr(step: int, initial_lr: float, warm_iterations: int, warm_lr: float, poly_gamma: float, num_iterations: int) -> float:
    """Learning rate of optimizer step `step` (0-based) under the reference's `LinearWarmupPolyLR` stepped once per batch
    (nndet/training/learning_rate.py:126-183, configured at nndet/ptmodule/retinaunet/base.py:329-336).  The reference evaluates its
    formulas with torch's `_step_count`, which is already 1 when the first batch runs, so step s uses iteration s + 1:
        s <  warm_iterations : warm_lr + (initial_lr - warm_lr) * (s + 1) / warm_iterations            (linear_warm_up, :27-49)
        s >= warm_iterations : initial_lr * (1 - it / poly_iterations) ** gamma, it = s + 1 - warm_iterations, clamped to
                               poly_iterations - 1 once the schedule is over                           (poly_lr, :52-78)
    Pinned against the executed reference scheduler: tests/golden/lr.npz."""
    if step < warm_iterations:
        return warm_lr + (initial_lr - warm_lr) * (float(step + 1) / float(warm_iterations))
    poly_iterations = num_iterations - warm_iterations
    if poly_iterations <= 0:            # degenerate schedule (the reference would divide by zero): stay at the plateau
        return initial_lr
    it = step + 1 - warm_iterations
    if it >= poly_iterations:
        it = poly_iterations - 1
    return initial_lr * (1 - it / float(poly_iterations)) ** poly_gamma


class Trainer:
    """Full train step = forward + losses (+ optional detection post-processing) + backward + gradient all-reduce +
    SGD.  Nothing in here synchronises with the host; `losses` are device tensors (the reference's caller reads
    them with .item(), nndet/ptmodule/retinaunet/base.py:154)."""

    def __init__(self, model: nn.Module, initial_lr=0.01, momentum=0.9, nesterov=True, weight_decay=3e-5,
                 warm_iterations=4000, warm_lr=1e-6, poly_gamma=0.9, num_iterations=50 * 2500, distributed: bool = False,
                 bucket_mb: Optional[float] = 25.0):
        self.model = model
        self.fp = FlatParameters(model)
        self.cfg = dict(initial_lr=initial_lr, warm_iterations=warm_iterations, warm_lr=warm_lr, poly_gamma=poly_gamma,
                        num_iterations=num_iterations)
        self.momentum, self.nesterov, self.weight_decay = momentum, nesterov, weight_decay
        self.step_idx = 0
        self.distributed = distributed and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.distributed else 1
        self.buckets = None
        if self.distributed:
            dist.broadcast(self.fp.flat, src=0)
            bump_weights_epoch()
            if bucket_mb is not None:                 # default: exchange ~25 MB gradient buckets while the backward pass still runs (None: one all-reduce after it)
                from .arch.conv import set_grad_observer
                self.buckets = GradientBuckets(model, self.fp, bucket_mb)
                set_grad_observer(self.buckets)

    def optimizer_step(self):
        fp = self.fp
        fp.check_homed()
        lr = poly_lr(self.step_idx, **self.cfg)
        if fp.first_step:
            fp.find_unused(self.model)
        L.check(L.lib().nnd_sgd_step_skip(L.ptr(fp.flat), L.ptr(fp.grad), L.ptr(fp.mom), c_longlong(fp.n), c_longlong(fp.n_decay),
                                          c_float(lr), c_float(self.momentum), c_float(self.weight_decay),
                                          c_int(1 if self.nesterov else 0), c_int(1 if fp.first_step else 0),
                                          c_float(1.0 / self.world), L.ptr(fp.skip), c_int(fp.n_skip), L.stream_ptr()),
                "nnd_sgd_step_skip")
        fp.first_step = False
        self.step_idx += 1
        bump_weights_epoch()              # packed bf16 weight copies are refreshed lazily by the layers

    def train_step(self, images: torch.Tensor, targets: dict, evaluation: bool = False):
        self.model.train()
        self.fp.zero_grad()
        self.model.defer_prediction_sync = True
        if self.buckets is not None:
            self.buckets.begin()
        losses, prediction = self.model.train_step(images, targets, evaluation=evaluation, batch_num=self.step_idx)
        loss = sum(losses.values())
        loss.backward()
        if self.buckets is not None:
            self.buckets.finish()
        elif self.distributed:
            dist.all_reduce(self.fp.grad)      # 76 MB fp32, NCCL over NVLink; mean folded into the SGD kernel
        self.optimizer_step()
        if prediction is not None:
            prediction = prediction.resolve()      # the step's only host read, after everything is enqueued
        return losses, prediction
