"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Retina U-Net (v001) network + train step.

A compact torch-CPU restatement of what `RetinaUNetV001.from_config_plan` assembles
(nndet/ptmodule/retinaunet/base.py:387-466 with the class set of v001.py:31-38 and
the kwargs of nndet/conf/train/v001.yaml:61-110).  The arithmetic of the reference on
this path *is* torch's Conv3d / ConvTranspose3d / InstanceNorm3d / GroupNorm / ReLU
(nndet/arch/conv.py:297-348,388-446, arch/layers/norm.py:26-50) -- a third-party
dependency the reference pins no tighter than its NGC image -- so the oracle calls the
same torch CPU operators in fp32.  Module/parameter names reproduce the reference's
state_dict keys, which scripts/gen_golden.py asserts against the real model.

Never imported by nndetection_b200/.  Pinned by scripts/gen_golden.py (outputs, losses
and gradients equal to the executed reference on seeded inputs -> tests/golden/).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import box_oracle as bo


def _t3(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v, v)


class ConvNormAct(nn.Sequential):
    """BaseConvNormAct, nndet/arch/conv.py:54-143: conv[, norm][, act]; bias iff no norm (:113)."""

    def __init__(self, cin, cout, k, stride=1, pad=0, norm=None, act=True, bias=None, transposed=False,
                 ch_per_group=16):
        super().__init__()
        bias = (norm is None) if bias is None else bias
        cls = nn.ConvTranspose3d if transposed else nn.Conv3d
        self.add_module("conv", cls(cin, cout, _t3(k), _t3(stride), _t3(pad), bias=bias))
        if norm == "instance":          # ConvInstanceRelu, conv.py:146-217 (eps 1e-5, affine)
            self.add_module("norm", nn.InstanceNorm3d(cout, eps=1e-5, affine=True))
        elif norm == "group":           # ConvGroupRelu, conv.py:220-294 (16 channels per group)
            self.add_module("norm", nn.GroupNorm(cout // ch_per_group, cout, eps=1e-5, affine=True))
        if act:
            self.add_module("act", nn.ReLU(inplace=norm is not None))


class _Stage(nn.Module):
    """StackedConvBlock2 with num_blocks=1, nndet/arch/blocks/basic.py:45-151."""

    def __init__(self, cin, cout, k, stride):
        super().__init__()
        pad = tuple((i - 1) // 2 for i in _t3(k))
        self.convs = nn.Sequential(nn.Sequential(
            ConvNormAct(cin, cout, k, stride, pad, norm="instance"),
            ConvNormAct(cout, cout, k, 1, pad, norm="instance")))

    def forward(self, x):
        return self.convs(x)


class Encoder(nn.Module):
    """Encoder, nndet/arch/encoder/modular.py:28-157: channels double per stage, capped at max_channels."""

    def __init__(self, conv_kernels, strides, in_channels, start_channels, max_channels):
        super().__init__()
        stages, self.out_channels = [], []
        c = in_channels
        for i, k in enumerate(conv_kernels):
            co = start_channels if i == 0 else min(c * 2, max_channels)
            stages.append(_Stage(c, co, k, 1 if i == 0 else strides[i - 1]))
            c = co
            self.out_channels.append(co)
        self.stages = nn.ModuleList(stages)
        self.strides = [_t3(s) for s in strides]

    def forward(self, x):
        outs = []
        for s in self.stages:
            x = s(x)
            outs.append(x)
        return outs


class UFPN(nn.Module):
    """UFPNModular, nndet/arch/decoder/base.py:316-417 with v001 kwargs: lateral 1x1x1 (bias, no norm/act),
    transposed-conv upsampling k=s=stride (:272-304), `out` 3x3x3 (bias, no norm/act; built before
    conv_settings['out'] is overwritten, :99-101 vs :382-384), no fusion convs.
    Output channels: compute_output_channels :182-199."""

    def __init__(self, enc: Encoder, conv_kernels, decoder_levels, fpn_channels, min_out=8):
        super().__init__()
        n = len(enc.out_channels)
        oc = [fpn_channels] * n
        for ol in [l for l in range(n) if l < min(decoder_levels)][::-1]:
            oc[ol] = max(min_out, oc[ol + 1] // 2)
        self.out_channels = oc
        self.lateral = nn.ModuleDict({f"P{l}": nn.Sequential(ConvNormAct(enc.out_channels[l], oc[l], 1, act=False))
                                      for l in range(n)})
        self.out = nn.ModuleDict({f"P{l}": nn.Sequential(ConvNormAct(
            oc[l], oc[l], conv_kernels[l], 1, tuple((i - 1) // 2 for i in _t3(conv_kernels[l])), act=False))
            for l in range(n)})
        self.up = nn.ModuleDict({f"P{l}": ConvNormAct(oc[l], oc[l - 1], enc.strides[l - 1], enc.strides[l - 1],
                                                      act=False, transposed=True) for l in range(1, n)})
        self.n = n

    def forward(self, feats):
        lat = [self.lateral[f"P{l}"](f) for l, f in enumerate(feats)]
        outs = [None] * self.n
        up = None
        for l in range(self.n - 1, -1, -1):
            x = lat[l] if up is None else lat[l] + up
            if l > 0:
                up = self.up[f"P{l}"](x)
            outs[l] = x
        return [self.out[f"P{l}"](o) for l, o in enumerate(outs)]


class _HeadBranch(nn.Module):
    def __init__(self, cin, cint, cout, num_convs):
        super().__init__()
        ci = nn.Sequential()
        ci.add_module("c_in", ConvNormAct(cin, cint, 3, 1, 1, norm="group"))
        for i in range(num_convs):
            ci.add_module(f"c_internal{i}", ConvNormAct(cint, cint, 3, 1, 1, norm="group"))
        self.conv_internal = ci
        self.conv_out = ConvNormAct(cint, cout, 3, 1, 1, act=False, bias=True)


class Classifier(_HeadBranch):
    """BCECLassifier, nndet/arch/heads/classifier.py:64-292 (prior-prob bias init :210-228)."""

    def __init__(self, cin, cint, num_classes, apos, num_convs=1, prior_prob=0.01):
        super().__init__(cin, cint, num_classes * apos, num_convs)
        self.num_classes = num_classes
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.normal_(m.weight, mean=0, std=0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        nn.init.constant_(self.conv_out.conv.bias, -math.log((1 - prior_prob) / prior_prob))

    def forward(self, x, level):
        y = self.conv_out(self.conv_internal(x))
        return y.permute(0, 2, 3, 4, 1).contiguous().view(x.size(0), -1, self.num_classes)


class Scale(nn.Module):
    """Scale, nndet/arch/layers/scale.py:21-43."""

    def __init__(self):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(1.0))

    def forward(self, x):
        return x * self.scale


class Regressor(_HeadBranch):
    """GIoURegressor, nndet/arch/heads/regressor.py:51-202,260-310 (normal(0, .01) init :189-201)."""

    def __init__(self, cin, cint, apos, num_levels, num_convs=1):
        super().__init__(cin, cint, apos * 6, num_convs)
        self.scales = nn.ModuleList([Scale() for _ in range(num_levels)])
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.normal_(m.weight, mean=0, std=0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x, level):
        y = self.scales[level](self.conv_out(self.conv_internal(x)))
        return y.permute(0, 2, 3, 4, 1).contiguous().view(x.size(0), -1, 6)


class Head(nn.Module):
    """DetectionHead.forward, nndet/arch/heads/comb.py:85-109."""

    def __init__(self, classifier, regressor):
        super().__init__()
        self.classifier, self.regressor = classifier, regressor

    def forward(self, fmaps):
        lg = [self.classifier(p, l) for l, p in enumerate(fmaps)]
        of = [self.regressor(p, l) for l, p in enumerate(fmaps)]
        return {"box_deltas": torch.cat(of, dim=1).reshape(-1, 6),
                "box_logits": torch.cat(lg, dim=1).flatten(0, -2)}


class Segmenter(nn.Module):
    """DiCESegmenterFgBg, nndet/arch/heads/segmenter.py:223-290: 1x1x1 conv P0 -> 2 channels."""

    def __init__(self, cin):
        super().__init__()
        self.conv_out = ConvNormAct(cin, 2, 1, act=False, bias=True)

    def forward(self, fmaps):
        return {"seg_logits": self.conv_out(fmaps[0])}


class RetinaUNetOracle(nn.Module):
    """BaseRetinaNet, nndet/core/retina.py:26-414, with matcher/sampler/losses from box_oracle."""

    def __init__(self, plan_arch: dict, plan_anchors: dict, head_sampler_kwargs=None, num_candidates=4):
        super().__init__()
        pa = plan_arch
        self.encoder = Encoder(pa["conv_kernels"], pa["strides"], pa["in_channels"], pa["start_channels"],
                               pa.get("max_channels", 320))
        self.decoder = UFPN(self.encoder, pa["conv_kernels"], pa["decoder_levels"], pa["fpn_channels"])
        apos = len(plan_anchors["width"][0]) * len(plan_anchors["height"][0]) * len(plan_anchors["depth"][0])
        self.head = Head(Classifier(pa["fpn_channels"], pa["head_channels"], pa["classifier_classes"], apos),
                         Regressor(pa["fpn_channels"], pa["head_channels"], apos, len(pa["decoder_levels"])))
        self.segmenter = Segmenter(self.decoder.out_channels[0])
        self.decoder_levels = tuple(pa["decoder_levels"])
        self.num_classes = pa["classifier_classes"]
        self.anchor_cfg = plan_anchors
        self.apos = apos
        self.num_candidates = num_candidates
        self.sampler_kw = dict(batch_size_per_image=32, positive_fraction=0.33, min_neg=1, pool_size=20)
        if head_sampler_kwargs:
            self.sampler_kw.update(head_sampler_kwargs)
        self.post = dict(topk=pa.get("topk_candidates", 10000), score_thresh=pa.get("score_thresh", 0),
                         min_size=pa.get("remove_small_boxes", 0.01), nms_thresh=pa.get("nms_thresh", 0.6),
                         detections_per_img=pa.get("detections_per_img", 100))

    def forward(self, x):
        """BaseRetinaNet.forward, retina.py:198-226."""
        fm_all = self.decoder(self.encoder(x))
        fm_head = [fm_all[i] for i in self.decoder_levels]
        pred = self.head(fm_head)
        anc, per_level = bo.anchors_for_image(x.shape[2:], [f.shape[2:] for f in fm_head],
                                              self.anchor_cfg["width"], self.anchor_cfg["height"],
                                              self.anchor_cfg["depth"])
        self.per_level = per_level
        return pred, [anc] * x.shape[0], self.segmenter(fm_all)

    def train_step(self, images, targets, seed: int = 0):
        """BaseRetinaNet.train_step, retina.py:86-159, with the hash sampler (box_oracle.hnm_select)
        standing in for torch.randperm.  Returns (losses, aux)."""
        pred, anchors, pseg = self(images)
        labels, matched = [], []
        for a, gb, gc in zip(anchors, targets["target_boxes"], targets["target_classes"]):
            _, m = bo.atss_match(gb, a, self.per_level, self.apos, self.num_candidates)
            l, mb = bo.assign_targets(m, gb, gc, a.shape[0])
            labels.append(l); matched.append(mb)
        lb, mb, ab = torch.cat(labels), torch.cat(matched), torch.cat(anchors)
        with torch.no_grad():
            fg = torch.sigmoid(pred["box_logits"]).max(dim=1)[0]
            pos, neg, pool = bo.hnm_select(lb, fg, images.shape[0], seed, **self.sampler_kw)
        losses = bo.head_loss(pred["box_logits"], pred["box_deltas"], lb, mb, ab, pos, neg, self.num_classes)
        losses.update(bo.seg_loss(pseg["seg_logits"], targets["target_seg"]))
        return losses, dict(pred=pred, anchors=anchors, pred_seg=pseg, labels=lb, pos=pos, neg=neg, pool=pool)

    @torch.no_grad()
    def postprocess(self, images, pred, anchors):
        """postprocess_detections, retina.py:292-330 + comb.py:140-158 + coder.py:219-241."""
        A = anchors[0].shape[0]
        boxes = bo.decode_single(pred["box_deltas"], torch.cat(anchors))
        probs = torch.sigmoid(pred["box_logits"])
        out = []
        for i in range(images.shape[0]):
            out.append(bo.postprocess_single_image(boxes[i * A:(i + 1) * A], probs[i * A:(i + 1) * A],
                                                   images.shape[2:], self.num_classes, **self.post))
        return out


# named configurations + synthetic batches live in the package (shared by bench.py, tests and the oracle)
from nndetection_b200.configs import make_plan, synth_batch  # noqa: E402,F401
