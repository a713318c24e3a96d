"""Convolution blocks with the reference's factory protocol, running on the sm_100a kernels.

Mirrors nndet/arch/conv.py: `Generator` (:28-51), `BaseConvNormAct` (:54-143, Sequential(conv[, norm][, act]),
bias iff no norm :113), `ConvInstanceRelu` (:146-217), `ConvGroupRelu` (:220-294).  Sub-module names `conv`,
`norm`, `act` and parameter shapes equal the reference's, so its checkpoints load (`state_dict` contract).
Differences by design: activations are bf16 NDHWC; conv + norm-statistics are one kernel, norm-apply + ReLU one
streaming pass; backward = norm-bwd (2 passes) + wgrad + dgrad kernels; no cuDNN, no CPU path.
"""
from typing import Callable, Optional, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import conv_ops as ops
from .conv_ops import pad32, t3

_WEIGHTS_EPOCH = [0]      # bumped by nndetection_b200 optimizers that update parameters through raw pointers


def bump_weights_epoch():
    _WEIGHTS_EPOCH[0] += 1


class Generator:
    """nndet/arch/conv.py:28-51: binds a conv class to a dimension."""

    def __init__(self, conv_cls, dim: int):
        self.dim = dim
        self.conv_cls = conv_cls

    def __call__(self, *args, **kwargs) -> nn.Module:
        return self.conv_cls(self.dim, *args, **kwargs)


class ConvParams(nn.Module):
    """Parameter holder named `conv` (weight/bias in torch.nn.Conv3d / ConvTranspose3d layout and init)."""

    def __init__(self, cin, cout, k, stride, padding, bias, transposed):
        super().__init__()
        k = t3(k)
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride, self.padding, self.transposed = k, t3(stride), t3(padding), transposed
        shape = (cin, cout, *k) if transposed else (cout, cin, *k)
        self.weight = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        # same as torch.nn.modules.conv._ConvNd.reset_parameters
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / fan_in ** 0.5 if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, k={self.kernel_size}, s={self.stride}, p={self.padding}, T={self.transposed}"


class NormParams(nn.Module):
    """Parameter holder named `norm`: instance norm (channels_per_group = 1) or group norm, affine."""

    def __init__(self, channels, channels_per_group, eps=1e-5, affine=True):
        super().__init__()
        self.channels, self.cpg, self.eps = channels, channels_per_group, eps
        if affine:
            self.weight = nn.Parameter(torch.ones(channels))
            self.bias = nn.Parameter(torch.zeros(channels))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)


_GRAD_OBSERVER = [None]      # object with layer_used(layer) / layer_done(layer), installed by training.GradientBuckets


def set_grad_observer(obs):
    _GRAD_OBSERVER[0] = obs


def notify_grad_observer(layer, z):
    """Gradient-bucket overlap (training.GradientBuckets, opt-in): a layer's parameter gradients are complete when the backward of
    every one of its uses in this step has run; the autograd node's post-hook fires right after that use's weight-gradient kernels
    were enqueued on the backward stream."""
    obs = _GRAD_OBSERVER[0]
    if obs is not None and z.grad_fn is not None:
        obs.layer_used(layer)
        z.grad_fn.register_hook(lambda *_a, _l=layer: obs.layer_done(_l))


def _grad_target(param, shape, dev):
    """Where a parameter-gradient kernel accumulates (all of them ADD into their destination).  When the parameter already
    owns a dense fp32 `.grad` (the Trainer's flat gradient buffer, zeroed once per step) the kernels add straight into it
    and autograd gets `None` for that input -- no zero-fill, no temporary and no accumulate kernel per parameter
    (~220 tiny launches per step).  Only for parameters whose owner opted in (`param._nnd_direct_grad`, set by
    training.FlatParameters): a third-party trainer that relies on autograd's AccumulateGrad hooks (torch DDP, as PL
    would wrap the module) keeps the ordinary path -- a zeroed temporary returned through autograd."""
    g = getattr(param, "grad", None) if param is not None and getattr(param, "_nnd_direct_grad", False) else None
    if g is not None and g.dtype == torch.float32 and g.is_contiguous() and tuple(g.shape) == tuple(shape) and g.device == dev:
        return g, None
    t = torch.zeros(shape, dtype=torch.float32, device=dev)
    return t, t


class _ConvBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, residual, layer):
        conv: ConvParams = layer.conv
        N = x.shape[0]
        cin, cout = conv.in_channels, conv.out_channels
        first = layer.is_first
        if first:
            x = x.contiguous().float()                    # fp32 NCDHW image batch
        else:
            x = ops.as_cl(x)
        plan = layer.plan(N, tuple(x.shape[2:]))
        wp_f, _ = layer.packed()
        dev = x.device
        y = ops.empty_cl(N, cout, plan.out_sp, device=dev)
        has_norm = layer.norm is not None
        stats = torch.zeros((2, N, cout), dtype=torch.float32, device=dev) if has_norm else None
        res = ops.as_cl(residual) if residual is not None else None
        if first:
            ops.conv_first_fprop(x, weight.detach(), plan.fprop[0], cout, y, stats[0] if has_norm else None,
                                 stats[1] if has_norm else None)
        elif (conv.transposed and not has_norm and ops.pointwise_tma_enabled() and N * x.shape[2] * x.shape[3] * x.shape[4] >= 128
              and all(v in (1, 2) for v in plan.s)):
            # the whole up-convolution in one launch: taps stacked along N, the lateral added in the epilogue
            ops.conv_upconv(x, wp_f, N, tuple(x.shape[2:]), cin, cout, plan.s, y, bias.detach() if bias is not None else None, res)
        else:
            items_f = layer.packed_items()[0]
            for g in plan.fprop:
                ops.conv_gather(x, wp_f, g, y, cout, pad32(cout), bias=bias.detach() if bias is not None else None,
                                residual=res, stat_sum=stats[0] if has_norm else None,
                                stat_sq=stats[1] if has_norm else None, w_items=items_f)
        V = plan.out_sp[0] * plan.out_sp[1] * plan.out_sp[2]
        if has_norm:
            a, b, mean, rstd = ops.norm_finalize(stats[0], stats[1], gamma.detach() if gamma is not None else None,
                                                 beta.detach() if beta is not None else None, N, cout, layer.norm.cpg, V,
                                                 layer.norm.eps)
            z = ops.empty_cl(N, cout, plan.out_sp, device=dev)
            ops.norm_apply(y, a, b, N, V, cout, layer.has_act, z)
            ctx.save_for_backward(x, weight, gamma, y, a, b, mean, rstd)
        else:
            z = y
            ctx.save_for_backward(x, weight, gamma)
        ctx.layer, ctx.plan, ctx.N, ctx.V = layer, plan, N, V
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        return z

    @staticmethod
    def backward(ctx, dz):
        layer, plan, N, V = ctx.layer, ctx.plan, ctx.N, ctx.V
        layer._nnd_grad_seen = True           # training.FlatParameters.find_unused: this layer's parameters receive gradients
        conv: ConvParams = layer.conv
        cin, cout = conv.in_channels, conv.out_channels
        has_norm = layer.norm is not None
        dz = ops.as_cl(dz)
        dev = dz.device
        dgamma = dbeta = None
        if has_norm:
            x, weight, gamma, y, a, b, mean, rstd = ctx.saved_tensors
            dy = ops.empty_cl(N, cout, plan.out_sp, device=dev)
            affine = gamma is not None
            beta_p = layer.norm.bias if affine else None
            if affine:
                dgamma, dgamma_ret = _grad_target(layer.norm.weight, (cout,), dev)
                dbeta, dbeta_ret = _grad_target(beta_p, (cout,), dev)
            ops.norm_backward(dz, y, a, b, mean, rstd, gamma.detach() if affine else None, N, V, cout, layer.norm.cpg,
                              layer.has_act, dy, dgamma, dbeta)
        else:
            x, weight, gamma = ctx.saved_tensors
            dy = dz
        dbias_ret = None
        if ctx.has_bias:
            dbias, dbias_ret = _grad_target(conv.bias, (cout,), dev)
            ops.channel_sum(dy, N * V, cout, cout, dbias)
        dw, dw_ret = _grad_target(conv.weight, tuple(weight.shape), dev)
        T = plan.T
        if layer.is_first:
            ops.conv_first_wgrad(x, dy, plan.fprop[0], cout, dw)
        else:
            s_co, s_ci = (T, cout * T) if conv.transposed else (cin * T, T)
            if (conv.transposed and ops.wgrad_strided_tc_enabled() and plan.s[2] == 2 and cin >= 64 and cin % 32 == 0
                    and cout % 32 == 0):
                # opt-in: one strided gather with the roles swapped (dense = x, strided = dy) instead of one launch per tap
                ops.conv_wgrad(x, cin, dy, cout, plan.wgrad_swapped, dw, s_ci, s_co, 1, cin, cout)
            else:
                for g in plan.wgrad:
                    ops.conv_wgrad(dy, cout, x, cin, g, dw, s_co, s_ci, 1, cout, cin)
        dx = None
        if ctx.needs_input_grad[0] and not layer.is_first:
            _, wp_b = layer.packed()
            dx = (ops.empty_cl if plan.dgrad_covers_all else ops.zeros_cl)(N, cin, plan.in_sp, device=dev)
            items_b = layer.packed_items()[1]
            for g in plan.dgrad:
                ops.conv_gather(dy, wp_b, g, dx, cin, pad32(cin), w_items=items_b)
        dres = dy if ctx.has_res else None
        if has_norm and gamma is not None:
            dgamma, dbeta = dgamma_ret, dbeta_ret
        return dx, dw_ret, dbias_ret, dgamma, dbeta, dres, None


class BaseConvNormAct(nn.Module):
    """conv -> norm -> act block (nndet/arch/conv.py:54-143) on the B200 kernels."""

    def __init__(self, dim: int, in_channels: int, out_channels: int, norm: Optional[str], act: Optional[str],
                 kernel_size, stride=1, padding=0, dilation=1, groups: int = 1, bias: Optional[bool] = None,
                 transposed: bool = False, norm_kwargs: Optional[dict] = None, act_inplace: Optional[bool] = None,
                 act_kwargs: Optional[dict] = None, initializer: Callable[[nn.Module], None] = None):
        super().__init__()
        if dim != 3:
            raise NotImplementedError("nndetection_b200 implements the volumetric (3-D) hot path only")
        if t3(dilation) != (1, 1, 1) or groups != 1:
            raise NotImplementedError("dilation / grouped convolutions are not part of the v001 hot path")
        norm_kwargs = dict(norm_kwargs or {})
        bias = bool(norm is None) if bias is None else bias          # conv.py:113
        if bias and norm is not None:
            raise NotImplementedError("bias together with a normalisation layer")
        self.conv = ConvParams(in_channels, out_channels, kernel_size, stride, padding, bias, transposed)
        self.norm = None
        if norm is not None:
            kind = norm.lower() if isinstance(norm, str) else None
            if kind == "instance":
                cpg = 1
            elif kind == "group":
                cpg = norm_kwargs.get("channels_per_group", None) or out_channels // norm_kwargs["num_groups"]
            else:
                raise NotImplementedError(f"normalisation {norm!r}")
            self.norm = NormParams(out_channels, cpg, eps=norm_kwargs.get("eps", 1e-5), affine=norm_kwargs.get("affine", True))
        self.has_act = act is not None
        if self.has_act:
            if not (isinstance(act, str) and act.lower() == "relu"):
                raise NotImplementedError(f"activation {act!r}")
            if self.norm is None:
                raise NotImplementedError("activation without normalisation")
            self.act = nn.ReLU(inplace=True)          # parameter-free; kept for module-name parity (fused in-kernel)
        self.is_first = in_channels < 8
        if not self.is_first and in_channels % 32:
            raise NotImplementedError(f"in_channels={in_channels}: the kernels need multiples of 32 (or < 8 for the image layer)")
        if out_channels % 32:
            raise NotImplementedError(f"out_channels={out_channels}: use the head / segmenter modules for narrow outputs")
        if self.is_first and (transposed or t3(stride) != (1, 1, 1) or bias):
            raise NotImplementedError("image-input layer must be a stride-1 conv followed by a norm")
        self._plans, self._packed, self._packed_key = {}, None, None
        if initializer is not None:
            self.apply(initializer)

    # ---- caches
    def plan(self, N: int, in_sp) -> ops.ConvPlan:
        key = (N, tuple(in_sp))
        if key not in self._plans:
            c = self.conv
            self._plans[key] = ops.ConvPlan(N, c.in_channels, pad32(c.out_channels), in_sp, c.kernel_size, c.stride,
                                            c.padding, c.transposed)
        return self._plans[key]

    def packed(self):
        w = self.conv.weight
        key = (w._version, _WEIGHTS_EPOCH[0], w.data_ptr())
        if self._packed is None or self._packed_key != key:
            c = self.conv
            T = c.kernel_size[0] * c.kernel_size[1] * c.kernel_size[2]
            if self.is_first:
                self._packed = (None, None)
            else:
                self._packed = ops.pack_weights(w.detach(), c.out_channels, c.in_channels, T, c.transposed)
            self._packed_key = key
            self._items = None
        return self._packed

    def packed_items(self):
        """(fprop, dgrad) `ItemPack`s of the current weights for the opt-in bulk-copy variant of the tile kernel, or (None, None):
        refreshed together with `packed()`; only built while that variant is switched on."""
        if not ops.tc_bulk_enabled() or self.is_first:
            return None, None
        wp_f, wp_b = self.packed()
        if getattr(self, "_items", None) is None:
            self._items = (ops.repack_items(wp_f) if wp_f is not None else None, ops.repack_items(wp_b) if wp_b is not None else None)
        return self._items

    def forward(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        n = self.norm
        z = _ConvBlockFn.apply(x, self.conv.weight, self.conv.bias, n.weight if n is not None else None,
                               n.bias if n is not None else None, residual, self)
        notify_grad_observer(self, z)
        return z

    def __deepcopy__(self, memo):                  # planner deep-copies the model (nndet/planning/estimator.py:130)
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_plans", "_packed", "_packed_key", "_items", "_nnd_grad_seen"):
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new._plans, new._packed, new._packed_key = {}, None, None
        return new


class ConvInstanceRelu(BaseConvNormAct):
    """nndet/arch/conv.py:146-217."""

    def __init__(self, dim, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=None,
                 transposed=False, add_norm=True, add_act=True, act_inplace=None, norm_eps=1e-5, norm_affine=True,
                 initializer=None):
        super().__init__(dim, in_channels, out_channels, "Instance" if add_norm else None, "ReLU" if add_act else None,
                         kernel_size, stride, padding, dilation, groups, bias, transposed,
                         norm_kwargs={"eps": norm_eps, "affine": norm_affine}, act_inplace=act_inplace,
                         initializer=initializer)


class ConvGroupRelu(BaseConvNormAct):
    """nndet/arch/conv.py:220-294 (GroupNorm with `norm_channels_per_group` channels per group)."""

    def __init__(self, dim, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=None,
                 transposed=False, add_norm=True, add_act=True, act_inplace=None, norm_eps=1e-5, norm_affine=True,
                 norm_channels_per_group=16, initializer=None):
        super().__init__(dim, in_channels, out_channels, "Group" if add_norm else None, "ReLU" if add_act else None,
                         kernel_size, stride, padding, dilation, groups, bias, transposed,
                         norm_kwargs={"eps": norm_eps, "affine": norm_affine, "channels_per_group": norm_channels_per_group},
                         act_inplace=act_inplace, initializer=initializer)
