"""Encoder, U-FPN decoder, detection heads and segmenter of Retina U-Net v001 on the B200 kernels.

Each class mirrors the constructor / forward / helper interface and the state_dict keys of its reference:
  StackedConvBlock2   nndet/arch/blocks/basic.py:45-151
  Encoder             nndet/arch/encoder/modular.py:28-157
  UFPNModular         nndet/arch/decoder/base.py:28-417
  BCECLassifier       nndet/arch/heads/classifier.py:64-292
  GIoURegressor       nndet/arch/heads/regressor.py:51-202,260-310   (Scale: arch/layers/scale.py:21-43)
  DetectionHeadHNMNative  nndet/arch/heads/comb.py:60-158,206-276,351-405
  DiCESegmenterFgBg   nndet/arch/heads/segmenter.py:51-290
"""
import math
from ctypes import c_float, c_int, c_longlong
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib as L
from ..core.boxes import engine as E
from . import conv_ops as ops
from .conv_ops import pad32, t3


_LEVEL_STREAMS = {}          # per (device, pyramid level); module-level so that the model stays deep-copyable / picklable


def level_stream(device, level: int) -> "torch.cuda.Stream":
    """The side stream of a pyramid level (decoder `out` convolution and detection-head convolutions of the coarse levels)."""
    key = (str(device), int(level))
    if key not in _LEVEL_STREAMS:
        _LEVEL_STREAMS[key] = torch.cuda.Stream(device=device)
    return _LEVEL_STREAMS[key]


# ------------------------------------------------------------------------------------------------ encoder
class StackedConvBlock2(nn.Module):
    expansion = 2

    def __init__(self, conv, in_channels, conv_kernel, stride=None, out_channels=None, max_out_channels=None,
                 num_blocks: int = 1, **kwargs):
        super().__init__()
        if out_channels is not None and max_out_channels is not None and out_channels > max_out_channels:
            raise ValueError("Output channels can not be larger than max output channels")
        if out_channels is None:
            out_channels = in_channels * self.expansion
        if max_out_channels is not None and out_channels > max_out_channels:
            out_channels = max_out_channels
        if stride is None:
            stride = 1
        k = t3(conv_kernel)
        padding = tuple((i - 1) // 2 for i in k)
        blocks = [self.build_block(conv, in_channels, out_channels, k, stride, padding, **kwargs)]
        for _ in range(num_blocks - 1):
            blocks.append(self.build_block(conv, out_channels, out_channels, k, 1, padding, **kwargs))
        self.convs = nn.Sequential(*blocks)
        self.out_channels = out_channels

    @staticmethod
    def build_block(conv, in_channels, out_channels, kernel_size, stride, padding, **kwargs):
        return nn.Sequential(
            conv(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=stride,
                 padding=padding, **kwargs),
            conv(in_channels=out_channels, out_channels=out_channels, kernel_size=kernel_size, stride=1,
                 padding=padding, **kwargs))

    def forward(self, x):
        return self.convs(x)

    def get_output_channels(self):
        return self.out_channels


class Encoder(nn.Module):
    def __init__(self, conv, conv_kernels, strides, block_cls, in_channels: int, start_channels: int,
                 stage_kwargs=None, out_stages=None, max_channels: int = None, first_block_cls=None):
        super().__init__()
        self.num_stages = len(conv_kernels)
        self.dim = conv.dim
        if stage_kwargs is None:
            stage_kwargs = [{}] * self.num_stages
        elif isinstance(stage_kwargs, dict):
            stage_kwargs = [stage_kwargs] * self.num_stages
        self.out_stages = list(range(self.num_stages)) if out_stages is None else out_stages
        first_block_cls = first_block_cls or block_cls
        if isinstance(strides[0], int):
            strides = [tuple([s] * self.dim) for s in strides]
        self.strides = strides
        stages, self.out_channels = [], []
        for i in range(self.num_stages):
            if i == 0:
                blk = first_block_cls(conv=conv, in_channels=in_channels, out_channels=start_channels,
                                      conv_kernel=conv_kernels[i], stride=None, max_out_channels=max_channels,
                                      **stage_kwargs[i])
            else:
                blk = block_cls(conv=conv, in_channels=in_channels, out_channels=None, conv_kernel=conv_kernels[i],
                                stride=strides[i - 1], max_out_channels=max_channels, **stage_kwargs[i])
            in_channels = blk.get_output_channels()
            self.out_channels.append(in_channels)
            stages.append(blk)
        self.stages = nn.ModuleList(stages)

    def forward(self, x: Tensor) -> List[Tensor]:
        outs = []
        for i, m in enumerate(self.stages):
            x = m(x)
            if i in self.out_stages:
                outs.append(x)
        return outs

    def get_channels(self) -> List[int]:
        return [self.out_channels[i] for i in range(self.num_stages) if i in self.out_stages]

    def get_strides(self) -> List[List[int]]:
        out = []
        for i in range(self.num_stages):
            out.append([1] * self.dim if i == 0 else [a * b for a, b in zip(out[i - 1], self.strides[i - 1])])
        return out


# ------------------------------------------------------------------------------------------------ decoder
class UFPNModular(nn.Module):
    """U-shaped FPN with v001 settings (lateral 1x1x1 / transposed-conv up / out 3x3x3, all bias, no norm/act).
    The top-down sum `lateral + up` (decoder/base.py:405) is fused into the transposed conv's epilogue."""

    def __init__(self, conv, strides, in_channels, conv_kernels, decoder_levels, fixed_out_channels: int,
                 min_out_channels: int = 8, upsampling_mode: str = "nearest", num_lateral: int = 1,
                 norm_lateral: bool = False, activation_lateral: bool = False, num_out: int = 1, norm_out: bool = False,
                 activation_out: bool = False, num_fusion: int = 0, norm_fusion: bool = False,
                 activation_fusion: bool = False):
        super().__init__()
        if len(strides) != len(in_channels):
            raise ValueError("Strides must contain same number of elements as channels.")
        if upsampling_mode.lower() != "transpose" or num_lateral != 1 or num_out != 1 or num_fusion != 0 or \
                norm_lateral or activation_lateral or norm_out or activation_out:
            raise NotImplementedError("only the v001 decoder settings (transpose upsampling, 1 lateral, 1 out, no "
                                      "norm/act, no fusion convs) are implemented on this path")
        self.dim = conv.dim
        self.num_level = len(in_channels)
        self.in_channels = in_channels
        self.decoder_levels = decoder_levels
        st = [s if isinstance(s, Sequence) else (s,) * self.dim for s in strides]
        self.strides = [tuple(int(b / a) for a, b in zip(st[i - 1], st[i])) for i in range(1, len(st))]
        if isinstance(conv_kernels, int):
            conv_kernels = [conv_kernels] * self.num_level
        self.conv_kernels = [t3(k) for k in conv_kernels]
        self.conv_paddings = [tuple((i - 1) // 2 for i in k) for k in self.conv_kernels]
        self.min_out_channels, self.fixed_out_channels = min_out_channels, fixed_out_channels
        self.out_channels = self.compute_output_channels()
        oc = self.out_channels
        kw = dict(add_norm=False, add_act=False)
        self.lateral = nn.ModuleDict({f"P{l}": nn.Sequential(conv(in_channels[l], oc[l], kernel_size=1, padding=0, stride=1, **kw))
                                      for l in range(self.num_level)})
        self.out = nn.ModuleDict({f"P{l}": nn.Sequential(conv(oc[l], oc[l], kernel_size=self.conv_kernels[l],
                                                              padding=self.conv_paddings[l], stride=1, **kw))
                                  for l in range(self.num_level)})
        self.up = nn.ModuleDict({f"P{l}": conv(oc[l], oc[l - 1], kernel_size=self.strides[l - 1], stride=self.strides[l - 1],
                                               transposed=True, **kw) for l in range(1, self.num_level)})

    def compute_output_channels(self) -> List[int]:
        oc = [self.fixed_out_channels] * self.num_level
        if self.decoder_levels is not None:
            for ol in [l for l in range(self.num_level) if l < min(self.decoder_levels)][::-1]:
                oc[ol] = max(self.min_out_channels, oc[ol + 1] // 2)
        return oc

    def get_channels(self) -> List[int]:
        return self.out_channels

    def forward(self, inp_seq: Sequence[Tensor], levels: Optional[Sequence[int]] = None,
                side_levels: Sequence[int] = ()) -> List[Tensor]:
        """levels: pyramid levels whose `out` convolution is needed (default: all, the reference's protocol).  The detector passes
        its head levels + level 0 for the segmenter; an `out` map nobody reads (P1 under the LUNA plan) is then not computed -- dead
        code in the reference too: it has no consumer and receives no gradient.  Skipped levels are `None` in the returned list.
        side_levels: levels whose `out` convolution runs on that level's side stream (the coarse levels' 30-70 us launches overlap
        the rest of the top-down path; the consumer -- the detection head -- continues on the same stream)."""
        lat = [self.lateral[f"P{l}"](f) for l, f in enumerate(inp_seq)]
        xs = [None] * self.num_level
        x = lat[-1]
        xs[-1] = x
        outs = [None] * self.num_level
        need = set(range(self.num_level)) if levels is None else set(levels)
        cur = torch.cuda.current_stream(x.device) if x.is_cuda else None

        def out_on_side(l):
            if l in need and l in side_levels and cur is not None:
                st = level_stream(x.device, l)
                st.wait_stream(cur)
                self.out[f"P{l}"][0].packed()                     # pack on the current stream (no-op when cached)
                with torch.cuda.stream(st):
                    outs[l] = self.out[f"P{l}"](xs[l])
                xs[l].record_stream(st)
        out_on_side(self.num_level - 1)
        for l in range(self.num_level - 1, 0, -1):
            x = self.up[f"P{l}"](x, residual=lat[l - 1])          # lateral + up, one kernel
            xs[l - 1] = x
            out_on_side(l - 1)
        for l, v in enumerate(xs):
            if l in need and outs[l] is None:
                outs[l] = self.out[f"P{l}"](v)
        return outs


# ------------------------------------------------------------------------------------------------ heads
class Scale(nn.Module):
    def __init__(self, scale: float = 1.):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, inp):
        return inp * self.scale


class _OutConv(nn.Module):
    """Sequential-compatible holder for the final head conv (`conv_out.conv.{weight,bias}`)."""

    def __init__(self, cin, cout):
        super().__init__()
        from .conv import ConvParams
        self.conv = ConvParams(cin, cout, 3, 1, 1, True, False)


class _HeadBranch(nn.Module):
    def __init__(self, conv, in_channels, internal_channels, out_channels, num_convs, add_norm=True, **kwargs):
        super().__init__()
        ci = nn.Sequential()
        ci.add_module("c_in", conv(in_channels, internal_channels, kernel_size=3, stride=1, padding=1, add_norm=add_norm, **kwargs))
        for i in range(num_convs):
            ci.add_module(f"c_internal{i}", conv(internal_channels, internal_channels, kernel_size=3, stride=1, padding=1,
                                                 add_norm=add_norm, **kwargs))
        self.conv_internal = ci
        self.conv_out = _OutConv(internal_channels, out_channels)
        self.out_channels = out_channels

    def _normal_init(self):
        from .conv import ConvParams
        for m in self.modules():
            if isinstance(m, ConvParams):
                nn.init.normal_(m.weight, mean=0, std=0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)


class BCECLassifier(_HeadBranch):
    def __init__(self, conv, in_channels, internal_channels, num_classes, anchors_per_pos, num_levels, num_convs=3,
                 add_norm=True, prior_prob=None, weight=None, reduction="mean", smoothing=0.0, loss_weight=1., **kwargs):
        if weight is not None or reduction != "mean" or smoothing != 0.0 or loss_weight != 1.:
            raise NotImplementedError("only the v001 BCE settings (mean reduction, no smoothing/weights) are implemented")
        super().__init__(conv, in_channels, internal_channels, num_classes * anchors_per_pos, num_convs, add_norm, **kwargs)
        self.dim, self.num_levels, self.num_convs = conv.dim, num_levels, num_convs
        self.num_classes, self.anchors_per_pos = num_classes, anchors_per_pos
        self.prior_prob = prior_prob
        if prior_prob is not None:                       # classifier.py:210-228
            self._normal_init()
            nn.init.constant_(self.conv_out.conv.bias, -math.log((1 - prior_prob) / prior_prob))

    def box_logits_to_probs(self, box_logits: Tensor) -> Tensor:
        return E.sigmoid_fg(box_logits, want_fg=False)[0]


class GIoURegressor(_HeadBranch):
    def __init__(self, conv, in_channels, internal_channels, anchors_per_pos, num_levels, num_convs=3, add_norm=True,
                 reduction="sum", loss_weight=1., learn_scale=False, **kwargs):
        if reduction != "sum" or loss_weight != 1.:
            raise NotImplementedError("only the v001 GIoU settings (sum reduction, weight 1) are implemented")
        super().__init__(conv, in_channels, internal_channels, anchors_per_pos * 6, num_convs, add_norm, **kwargs)
        self.dim, self.num_levels, self.num_convs = conv.dim, num_levels, num_convs
        self.anchors_per_pos, self.learn_scale = anchors_per_pos, learn_scale
        if learn_scale:
            self.scales = nn.ModuleList([Scale() for _ in range(num_levels)])
        self._normal_init()                              # regressor.py:189-201


class _HeadOutFn(torch.autograd.Function):
    """Final 3x3x3 convs of classifier and regressor on all pyramid levels, writing straight into
    box_logits [N*A, C] / box_deltas [N*A, 6] (image-major, level, position, anchor -- comb.py:101-108)."""

    @staticmethod
    def forward(ctx, head, wc, bc, wr, br, *rest):
        nl = head.num_levels
        scales, feats = rest[:head.n_scales], rest[head.n_scales:]
        fc, fr = [ops.as_cl(f) for f in feats[:nl]], [ops.as_cl(f) for f in feats[nl:]]
        N = fc[0].shape[0]
        C, apos = head.classifier.num_classes, head.classifier.anchors_per_pos
        sps = [tuple(f.shape[2:]) for f in fc]
        vox = [s[0] * s[1] * s[2] for s in sps]
        offs = [0]
        for v in vox:
            offs.append(offs[-1] + v * apos)
        A = offs[-1]
        dev = fc[0].device
        logits = torch.empty((N * A, C), dtype=torch.float32, device=dev)
        deltas = torch.empty((N * A, 6), dtype=torch.float32, device=dev)
        wpc, wpr = head.packed_out()
        cin = head.classifier.conv_out.conv.in_channels
        plans = [head.out_plan(N, sp) for sp in sps]
        lf, df = logits.view(-1), deltas.view(-1)
        for l in range(nl):
            ops.conv_gather(fc[l], wpc[0], plans[l].fprop[0], lf[offs[l] * C:], apos * C, pad32(apos * C), out_fp32=True,
                            out_n_stride=A * C, out_v_stride=apos * C, bias=bc.detach())
            ops.conv_gather(fr[l], wpr[0], plans[l].fprop[0], df[offs[l] * 6:], apos * 6, pad32(apos * 6), out_fp32=True,
                            out_n_stride=A * 6, out_v_stride=apos * 6, bias=br.detach(),
                            scale=scales[l].detach() if scales else None)
        ctx.save_for_backward(wc, wr, deltas, *scales, *fc, *fr)
        ctx.head, ctx.meta = head, (N, A, C, apos, sps, vox, offs, cin, plans)
        return logits, deltas

    @staticmethod
    def backward(ctx, d_logits, d_deltas):
        head = ctx.head
        head._nnd_grad_seen = True
        N, A, C, apos, sps, vox, offs, cin, plans = ctx.meta
        nl, ns = head.num_levels, head.n_scales
        saved = ctx.saved_tensors
        wc, wr, deltas = saved[0], saved[1], saved[2]
        scales = saved[3:3 + ns]
        fc, fr = saved[3 + ns:3 + ns + nl], saved[3 + ns + nl:]
        dev = wc.device
        d_logits = d_logits.contiguous().float().view(-1)
        d_deltas = d_deltas.contiguous().float().view(-1)
        lib = L.lib()
        wpc, wpr = head.packed_out()
        coc, cor = apos * C, apos * 6
        pc, pr = pad32(coc), pad32(cor)
        dwc, dwr = torch.zeros_like(wc), torch.zeros_like(wr)
        dbc = torch.zeros(coc, dtype=torch.float32, device=dev)
        dbr = torch.zeros(cor, dtype=torch.float32, device=dev)
        dscales = [torch.zeros((), dtype=torch.float32, device=dev) for _ in range(ns)]
        dfc, dfr = [], []
        T = 27
        for l in range(nl):
            V = vox[l]
            gl = torch.empty((N * V, pc), dtype=torch.bfloat16, device=dev)
            gr = torch.empty((N * V, pr), dtype=torch.bfloat16, device=dev)
            L.check(lib.nnd_pad_cast_f32_bf16(L.ptr(d_logits[offs[l] * C:]), c_int(N), c_longlong(V), c_int(coc),
                                              c_longlong(A * C), None, L.ptr(gl), c_int(pc), L.stream_ptr()), "pad_cast")
            if ns:
                L.check(lib.nnd_scale_grad(L.ptr(d_deltas[offs[l] * 6:]), L.ptr(deltas.view(-1)[offs[l] * 6:]), c_int(N),
                                           c_longlong(V * cor), c_longlong(A * 6), L.ptr(scales[l]), L.ptr(dscales[l]),
                                           L.stream_ptr()), "nnd_scale_grad")
            L.check(lib.nnd_pad_cast_f32_bf16(L.ptr(d_deltas[offs[l] * 6:]), c_int(N), c_longlong(V), c_int(cor),
                                              c_longlong(A * 6), L.ptr(scales[l]) if ns else None, L.ptr(gr), c_int(pr),
                                              L.stream_ptr()), "pad_cast")
            ops.channel_sum(gl, N * V, coc, pc, dbc)
            ops.channel_sum(gr, N * V, cor, pr, dbr)
            g = plans[l].fprop[0]
            ops.conv_wgrad(gl, pc, fc[l], cin, g, dwc, cin * T, T, 1, coc, cin)
            ops.conv_wgrad(gr, pr, fr[l], cin, g, dwr, cin * T, T, 1, cor, cin)
            dplan_c, dplan_r = head.out_dplan(N, sps[l])
            dxc = ops.empty_cl(N, cin, sps[l], device=dev)
            dxr = ops.empty_cl(N, cin, sps[l], device=dev)
            for gg in dplan_c.dgrad:
                ops.conv_gather(gl, wpc[1], gg, dxc, cin, pad32(cin))
            for gg in dplan_r.dgrad:
                ops.conv_gather(gr, wpr[1], gg, dxr, cin, pad32(cin))
            dfc.append(dxc); dfr.append(dxr)
        return (None, dwc, dbc, dwr, dbr, *dscales, *dfc, *dfr)




class DetectionHeadHNMNative(nn.Module):
    """Detection head with hard-negative mining and GIoU on decoded boxes (comb.py:351-405), sync-free."""

    def __init__(self, classifier: BCECLassifier, regressor: GIoURegressor, coder, sampler, log_num_anchors=None, **kw):
        super().__init__()
        self.classifier, self.regressor, self.coder, self.fg_bg_sampler = classifier, regressor, coder, sampler
        self.num_levels = classifier.num_levels
        self.n_scales = regressor.num_levels if regressor.learn_scale else 0
        self._packed, self._packed_key, self._plans = None, None, {}
        self.sample_seed = 0
        self.parallel_levels = True      # coarse pyramid levels on side streams (False: everything on the current stream, A/B)
        self.pyramid_levels = None       # decoder levels of the feature maps (set by BaseRetinaNet): selects the streams the decoder used

    def packed_out(self):
        from .conv import _WEIGHTS_EPOCH
        wc, wr = self.classifier.conv_out.conv.weight, self.regressor.conv_out.conv.weight
        key = (wc._version, wr._version, _WEIGHTS_EPOCH[0], wc.data_ptr(), wr.data_ptr())
        if self._packed_key != key:
            cin = wc.shape[1]
            self._packed = (ops.pack_weights(wc.detach(), wc.shape[0], cin, 27, False),
                            ops.pack_weights(wr.detach(), wr.shape[0], cin, 27, False))
            self._packed_key = key
        return self._packed

    def out_plan(self, N, sp):
        key = ("f", N, sp)
        if key not in self._plans:
            cin = self.classifier.conv_out.conv.in_channels
            self._plans[key] = ops.ConvPlan(N, cin, 32, sp, 3, 1, 1, False)
        return self._plans[key]

    def out_dplan(self, N, sp):
        key = ("d", N, sp)
        if key not in self._plans:
            cin = self.classifier.conv_out.conv.in_channels
            self._plans[key] = (ops.ConvPlan(N, cin, pad32(self.classifier.out_channels), sp, 3, 1, 1, False),
                                ops.ConvPlan(N, cin, pad32(self.regressor.out_channels), sp, 3, 1, 1, False))
        return self._plans[key]

    def forward(self, fmaps: List[Tensor]) -> Dict[str, Tensor]:
        if self.parallel_levels and len(fmaps) > 1 and fmaps[0].is_cuda:
            fc, fr = self._internal_convs_on_level_streams(fmaps)
        else:
            fc = [self.classifier.conv_internal(p) for p in fmaps]
            fr = [self.regressor.conv_internal(p) for p in fmaps]
        scales = [s.scale for s in self.regressor.scales] if self.n_scales else []
        logits, deltas = _HeadOutFn.apply(self, self.classifier.conv_out.conv.weight, self.classifier.conv_out.conv.bias,
                                          self.regressor.conv_out.conv.weight, self.regressor.conv_out.conv.bias,
                                          *scales, *fc, *fr)
        return {"box_deltas": deltas, "box_logits": logits}

    def _internal_convs_on_level_streams(self, fmaps: List[Tensor]):
        """The internal convolutions of the coarser pyramid levels (16^3 ... 4^3 at the LUNA plan) are launches of 30-70 us that fill a
        fraction of the 148 SMs each (64, 16, 4 tiles): level l >= 1 runs on its own side stream, concurrently with the other levels
        and with level 0 on the current stream.  The levels are independent (weights shared, read-only in forward; their gradient
        kernels all ADD atomically into the same buffers); autograd runs every node's backward on the stream its forward ran on and
        orders the streams itself, so the backward pass overlaps the same way."""
        cur = torch.cuda.current_stream(fmaps[0].device)
        streams = [level_stream(fmaps[0].device, lv) for lv in self.pyramid_levels[1:len(fmaps)]] if self.pyramid_levels is not None \
            else [level_stream(fmaps[0].device, 100 + i) for i in range(len(fmaps) - 1)]
        for m in list(self.classifier.conv_internal) + list(self.regressor.conv_internal):
            m.packed()                                 # (re-)pack the shared weights ONCE, on the current stream, before any level reads them
            m.packed_items()
        fc, fr = [None] * len(fmaps), [None] * len(fmaps)
        for l in range(1, len(fmaps)):
            st = streams[l - 1]
            st.wait_stream(cur)                        # (a feature map produced on this very stream is already in order)
            with torch.cuda.stream(st):
                fc[l] = self.classifier.conv_internal(fmaps[l])
                fr[l] = self.regressor.conv_internal(fmaps[l])
        fc[0] = self.classifier.conv_internal(fmaps[0])
        fr[0] = self.regressor.conv_internal(fmaps[0])
        for l in range(1, len(fmaps)):
            cur.wait_stream(streams[l - 1])
            fc[l].record_stream(cur); fr[l].record_stream(cur)      # allocated under a side stream, consumed on the current one
        return fc, fr

    def postprocess_for_inference(self, prediction: Dict[str, Tensor], anchors: List[Tensor]) -> Dict[str, Tensor]:
        """comb.py:140-158: decode all anchors + sigmoid."""
        boxes = E.decode_boxes(prediction["box_deltas"], anchors[0])
        probs = E.sigmoid_fg(prediction["box_logits"], want_fg=False)[0]
        return {"pred_boxes": boxes, "pred_probs": probs}

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_packed", "_packed_key", "_plans", "_nnd_grad_seen"):
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new._packed, new._packed_key, new._plans = None, None, {}
        return new


# ------------------------------------------------------------------------------------------------ segmenter
class _SegConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, owner=None):
        ctx.owner = owner
        x = ops.as_cl(x)
        N, C = x.shape[0], x.shape[1]
        sp = tuple(x.shape[2:])
        total = N * sp[0] * sp[1] * sp[2]
        logits = torch.empty((N, *sp, 2), dtype=torch.float32, device=x.device)
        L.check(L.lib().nnd_seg_conv_fwd(L.ptr(x), c_int(C), L.ptr(w.detach()), L.ptr(b.detach()), c_longlong(total),
                                         L.ptr(logits), L.stream_ptr()), "nnd_seg_conv_fwd")
        ctx.save_for_backward(x, w)
        return logits.permute(0, 4, 1, 2, 3)               # [N, 2, D, H, W] view, channels-last memory

    @staticmethod
    def backward(ctx, dlogits):
        x, w = ctx.saved_tensors
        if ctx.owner is not None:
            ctx.owner._nnd_grad_seen = True
        N, C = x.shape[0], x.shape[1]
        sp = tuple(x.shape[2:])
        total = N * sp[0] * sp[1] * sp[2]
        dl = dlogits.permute(0, 2, 3, 4, 1).contiguous().float()
        dx = ops.empty_cl(N, C, sp, device=x.device)
        dw = torch.zeros_like(w)
        db = torch.zeros(2, dtype=torch.float32, device=x.device)
        L.check(L.lib().nnd_seg_conv_bwd(L.ptr(x), c_int(C), L.ptr(w.detach()), L.ptr(dl), c_longlong(total), L.ptr(dx),
                                         L.ptr(dw), L.ptr(db), L.stream_ptr()), "nnd_seg_conv_bwd")
        return dx, dw, db, None


class _SegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, alpha, smooth):
        lg = logits.permute(0, 2, 3, 4, 1).contiguous().float()      # no copy when produced by _SegConvFn
        tg = target.contiguous().float()
        total = tg.numel()
        sums = torch.empty(4, dtype=torch.float64, device=lg.device)
        losses = torch.empty(2, dtype=torch.float32, device=lg.device)
        L.check(L.lib().nnd_seg_loss_fwd(L.ptr(lg), L.ptr(tg), c_longlong(total), c_float(alpha), c_float(smooth), L.ptr(sums),
                                         L.ptr(losses), L.stream_ptr()), "nnd_seg_loss_fwd")
        ctx.save_for_backward(lg, tg, sums)
        ctx.cfg = (alpha, smooth, total)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_ce, g_dice):
        lg, tg, sums = ctx.saved_tensors
        alpha, smooth, total = ctx.cfg
        dl = torch.empty_like(lg)
        L.check(L.lib().nnd_seg_loss_bwd(L.ptr(lg), L.ptr(tg), L.ptr(sums), c_longlong(total), c_float(alpha), c_float(smooth),
                                         L.ptr(g_ce.contiguous().float()), L.ptr(g_dice.contiguous().float()), L.ptr(dl),
                                         L.stream_ptr()), "nnd_seg_loss_bwd")
        return dl.permute(0, 4, 1, 2, 3), None, None, None


class DiCESegmenterFgBg(nn.Module):
    def __init__(self, conv, seg_classes: int, in_channels: Sequence[int], decoder_levels: Sequence[int],
                 internal_channels=None, num_internal: int = 0, add_norm=True, add_act=True, kernel_size=3,
                 alpha: float = 0.5, ce_kwargs=None, dice_kwargs=None, **kwargs):
        super().__init__()
        if num_internal != 0 or ce_kwargs:
            raise NotImplementedError("segmenter with internal convolutions / CE kwargs")
        dk = dict(dice_kwargs or {})
        if not dk.get("batch_dice", False) or dk.get("do_bg", False):
            raise NotImplementedError("only batch dice without background (v001) is implemented")
        self.smooth = dk.get("smooth_nom", 1e-5)
        from .conv import ConvParams
        self.seg_classes = 2                                   # FgBg: one foreground class + background
        self.in_channels, self.decoder_levels, self.alpha = in_channels, decoder_levels, alpha
        holder = nn.Module()
        holder.conv = ConvParams(in_channels[0], 2, 1, 1, 0, True, False)
        self.conv_out = holder
        self.conv_intermediate = None

    def forward(self, x: List[Tensor]) -> Dict[str, Tensor]:
        c = self.conv_out.conv
        return {"seg_logits": _SegConvFn.apply(x[0], c.weight.view(2, -1), c.bias, self)}

    def compute_loss(self, pred_seg: Dict[str, Tensor], target: Tensor) -> Dict[str, Tensor]:
        ce, dice = _SegLossFn.apply(pred_seg["seg_logits"], target, self.alpha, self.smooth)
        return {"seg_ce": ce, "seg_dice": dice}

    def postprocess_for_inference(self, prediction: Dict[str, Tensor], *args, **kwargs) -> Dict[str, Tensor]:
        return {"pred_seg": torch.softmax(prediction["seg_logits"].float(), dim=1)}
