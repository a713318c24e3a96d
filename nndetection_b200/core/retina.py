"""Retina U-Net detector core on the B200 kernels.

Mirrors `BaseRetinaNet` (nndet/core/retina.py:26-414) -- same constructor, `forward`, `train_step`,
`inference_step`, `postprocess_for_inference`, `postprocess_detections[_single_image]` and
`assign_targets_to_anchors` signatures and return structures (AbstractModel contract, nndet/arch/abstract.py:25-77) --
so the reference's Lightning module, planner/VRAM estimator and predictor can drive it unchanged.
What changed underneath: no per-image Python loops, no [G, A] matrices, no torch.where / .item() host syncs in the
training path; matching, sampling, losses and post-processing are device-resident kernels.
"""
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from .boxes import engine as E


_SIDE_STREAMS = {}          # per device; kept out of the module so that the model stays deep-copyable / picklable


def _side_stream(device, name: str = "targets") -> torch.cuda.Stream:
    key = (str(device), name)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class _HeadLossFn(torch.autograd.Function):
    """DetectionHeadHNMNative.compute_loss (nndet/arch/heads/comb.py:352-405) incl. hard-negative sampling."""

    @staticmethod
    def forward(ctx, logits, deltas, anchors, matches, labels, gt, sampler, batch_size):
        logits_c, deltas_c = logits.detach().contiguous().float(), deltas.detach().contiguous().float()
        _, fg = E.sigmoid_fg(logits_c, want_probs=False)                   # comb.py:262-263
        counts, pos, neg = sampler.sample_indices(labels, fg, batch_size)  # sampler.py:212-270
        losses, g_d, g_l = E.head_loss_fwd(logits_c, deltas_c, anchors, matches, gt, labels, pos, neg, counts)
        ctx.save_for_backward(g_d, g_l, pos, neg, counts)
        ctx.n = logits.shape[0]
        ctx.mark_non_differentiable(pos, neg, counts)
        return losses[0], losses[1], pos, neg, counts

    @staticmethod
    def backward(ctx, g_reg, g_cls, *_):
        g_d, g_l, pos, neg, counts = ctx.saved_tensors
        d_deltas, d_logits = E.head_loss_bwd(g_d, g_l, pos, neg, counts, ctx.n, g_reg.contiguous().float(),
                                             g_cls.contiguous().float())
        return d_logits, d_deltas, None, None, None, None, None, None


class DeferredPrediction:
    """Detections of a train step whose per-image counts have not been read back yet.  `resolve()` performs the one
    host read and returns the reference's prediction dict (lists of [R, ...] tensors)."""

    def __init__(self, ob, os_, ol, oc, pred_seg):
        self.ob, self.os_, self.ol, self.oc, self.pred_seg = ob, os_, ol, oc, pred_seg

    def resolve(self) -> Dict[str, Any]:
        cnt = self.oc.tolist()
        out = {"pred_boxes": [self.ob[i, :c] for i, c in enumerate(cnt)],
               "pred_scores": [self.os_[i, :c] for i, c in enumerate(cnt)],
               "pred_labels": [self.ol[i, :c] for i, c in enumerate(cnt)]}
        if self.pred_seg is not None:
            out["pred_seg"] = self.pred_seg
        return out


class BaseRetinaNet(nn.Module):
    def __init__(self, dim: int, encoder, decoder, head, num_classes: int, anchor_generator, matcher,
                 decoder_levels: tuple = (2, 3, 4, 5), score_thresh: float = None, detections_per_img: int = 100,
                 topk_candidates: int = 10000, remove_small_boxes: float = 1e-2, nms_thresh: float = 0.9,
                 segmenter=None):
        super().__init__()
        assert dim == 3, "volumetric hot path only"
        self.dim = dim
        self.decoder_levels = decoder_levels
        self.encoder, self.decoder, self.head = encoder, decoder, head
        self.num_foreground_classes = num_classes
        self.anchor_generator = anchor_generator
        self.proposal_matcher = matcher
        self.score_thresh, self.topk_candidates = score_thresh, topk_candidates
        self.detections_per_img, self.remove_small_boxes, self.nms_thresh = detections_per_img, remove_small_boxes, nms_thresh
        self.segmenter = segmenter
        # True: train_step(evaluation=True) returns a DeferredPrediction so that the caller can enqueue backward and
        # the optimizer before the (only) host synchronisation of the step (nndetection_b200.training.Trainer does).
        self.defer_prediction_sync = False

    # ---------------------------------------------------------------- forward (retina.py:198-226)
    def forward(self, inp: Tensor):
        from ..arch.net import UFPNModular
        lazy = isinstance(self.decoder, UFPNModular)
        if lazy:
            # only the decoder outputs somebody reads: the head's levels + level 0 for the segmenter; the coarse head levels' `out`
            # convolutions run on their level's side stream, where the head continues with them
            need = set(self.decoder_levels) | ({0} if self.segmenter is not None else set())
            side = tuple(self.decoder_levels[1:]) if getattr(self.head, "parallel_levels", False) and inp.is_cuda else ()
            features_maps_all = self.decoder(self.encoder(inp), levels=need, side_levels=side)
            if hasattr(self.head, "pyramid_levels"):
                self.head.pyramid_levels = tuple(self.decoder_levels)
        else:
            features_maps_all = self.decoder(self.encoder(inp))
        feature_maps_head = [features_maps_all[i] for i in self.decoder_levels]
        pred_seg = None
        seg_stream = None
        if self.segmenter is not None and inp.is_cuda and getattr(self.head, "parallel_levels", False):
            # the segmentation branch (1x1x1 convolution over the finest map: a streaming, HBM-bound launch) beside the tensor-bound head
            cur = torch.cuda.current_stream(inp.device)
            seg_stream = _side_stream(inp.device, "seg")
            seg_stream.wait_stream(cur)
            with torch.cuda.stream(seg_stream):
                pred_seg = self.segmenter(features_maps_all)
        pred_detection = self.head(feature_maps_head)
        anchors = self.anchor_generator(inp, feature_maps_head)
        if seg_stream is not None:
            torch.cuda.current_stream(inp.device).wait_stream(seg_stream)
            for v in pred_seg.values():
                v.record_stream(torch.cuda.current_stream(inp.device))
        elif self.segmenter is not None:
            pred_seg = self.segmenter(features_maps_all)
        return pred_detection, anchors, pred_seg

    # ---------------------------------------------------------------- training (retina.py:86-159)
    def train_step(self, images: Tensor, targets: dict, evaluation: bool, batch_num: int):
        target_boxes: List[Tensor] = targets["target_boxes"]
        target_classes: List[Tensor] = targets["target_classes"]
        target_seg: Tensor = targets["target_seg"]

        # Target assignment needs anchors + ground truth only, not the network outputs: once the anchors of this image
        # shape are cached it runs on a side stream BESIDE the forward pass (its single-CTA selection kernels would
        # otherwise sit on the critical path between forward and loss) and is joined before the loss.
        early = self.anchor_generator.lookup(images) if hasattr(self.anchor_generator, "lookup") else None
        if early is not None:
            cur = torch.cuda.current_stream(images.device)
            side = _side_stream(images.device)
            side.wait_stream(cur)                     # targets may have been produced / copied on the current stream
            with torch.cuda.stream(side):
                gt = E.GtBatch(target_boxes, target_classes, images.device)
                matches = self.proposal_matcher.match_batch(gt, early[0], early[1],
                                                            self.anchor_generator.num_anchors_per_location()[0])
                labels = E.assign_labels(matches, gt, early[0].shape[0])
            pred_detection, anchors, pred_seg = self(images)
            cur.wait_stream(side)
            # allocated under the side stream, consumed on the current one: tell the caching allocator, so that a block freed after
            # this step cannot be handed out again on the side stream while the current stream still reads it
            for t in (matches, labels, gt.boxes, gt.classes, gt.offsets, gt.img, gt.local):
                t.record_stream(cur)
            a0 = anchors[0]
            A = a0.shape[0]
        else:
            pred_detection, anchors, pred_seg = self(images)
            gt = E.GtBatch(target_boxes, target_classes, images.device)
            a0 = anchors[0]
            A = a0.shape[0]
            matches = self.proposal_matcher.match_batch(
                gt, a0, self.anchor_generator.get_num_acnhors_per_level(),
                self.anchor_generator.num_anchors_per_location()[0])
            labels = E.assign_labels(matches, gt, A)

        losses = {}
        reg, cls, pos_idx, neg_idx, counts = _HeadLossFn.apply(
            pred_detection["box_logits"], pred_detection["box_deltas"], a0, matches, labels, gt,
            self.head.fg_bg_sampler, images.shape[0])
        # NB: the reference omits "reg" when no positive anchor was sampled (comb.py:396); reading that count
        # would cost a host sync, so the key is always present (value 0 then) -- the summed loss is identical.
        losses["reg"], losses["cls"] = reg, cls
        if self.segmenter is not None:
            losses.update(self.segmenter.compute_loss(pred_seg, target_seg))
        self.last_sample = (pos_idx, neg_idx, counts, labels, matches)

        prediction = None
        if evaluation:
            if self.defer_prediction_sync:
                with torch.no_grad():
                    det = {k: v.detach() for k, v in pred_detection.items()}
                    ob, os_, ol, oc = self.postprocess_detections_device(det, anchors, images.shape[2:])
                    seg = self.segmenter.postprocess_for_inference(pred_seg)["pred_seg"] if self.segmenter is not None else None
                prediction = DeferredPrediction(ob, os_, ol, oc, seg)
            else:
                prediction = self.postprocess_for_inference(images=images, pred_detection=pred_detection,
                                                            pred_seg=pred_seg, anchors=anchors)
        return losses, prediction

    @torch.no_grad()
    def assign_targets_to_anchors(self, anchors: List[Tensor], target_boxes: List[Tensor],
                                  target_classes: List[Tensor]) -> Tuple[List[Tensor], List[Tensor]]:
        """Reference protocol (retina.py:228-290): per-image label vectors and matched GT boxes."""
        gt = E.GtBatch(target_boxes, target_classes, anchors[0].device)
        A = anchors[0].shape[0]
        matches = self.proposal_matcher.match_batch(
            gt, anchors[0], self.anchor_generator.get_num_acnhors_per_level(),
            self.anchor_generator.num_anchors_per_location()[0])
        labels = E.assign_labels(matches, gt, A)
        out_l, out_b = [], []
        for i, gb in enumerate(target_boxes):
            m = matches[i * A:(i + 1) * A]
            out_l.append(labels[i * A:(i + 1) * A])
            if gb.numel() > 0:
                out_b.append(gb.to(anchors[0])[m.clamp(min=0)])
            else:
                out_b.append(torch.zeros_like(anchors[0]))
        return out_l, out_b

    # ---------------------------------------------------------------- inference (retina.py:161-196,292-414)
    @torch.no_grad()
    def postprocess_for_inference(self, images: Tensor, pred_detection: Dict[str, Tensor], pred_seg, anchors):
        image_shapes = [images.shape[2:]] * images.shape[0]
        boxes, probs, labels = self.postprocess_detections(pred_detection=pred_detection, anchors=anchors,
                                                           image_shapes=image_shapes)
        prediction = {"pred_boxes": boxes, "pred_scores": probs, "pred_labels": labels}
        if self.segmenter is not None:
            prediction["pred_seg"] = self.segmenter.postprocess_for_inference(pred_seg)["pred_seg"]
        return prediction

    @torch.no_grad()
    def postprocess_detections_device(self, pred_detection: Dict[str, Tensor], anchors: List[Tensor], image_shape):
        """Sync-free core: fixed-size outputs + per-image counts on the device."""
        B, A, C = len(anchors), anchors[0].shape[0], self.num_foreground_classes
        boxes = E.decode_boxes(pred_detection["box_deltas"], anchors[0], clip_shape=tuple(int(s) for s in image_shape))
        probs = E.sigmoid_fg(pred_detection["box_logits"], want_fg=False)[0]
        if self.topk_candidates is None or self.detections_per_img is None:
            raise NotImplementedError("topk_candidates / detections_per_img = None")
        return E.detect_postprocess(boxes, probs, B, A, C, topk=self.topk_candidates, score_thresh=self.score_thresh,
                                    min_size=self.remove_small_boxes, nms_thresh=self.nms_thresh,
                                    det_per_img=self.detections_per_img)

    @torch.no_grad()
    def postprocess_detections(self, pred_detection: Dict[str, Tensor], anchors: List[Tensor], image_shapes):
        ob, os_, ol, oc = self.postprocess_detections_device(pred_detection, anchors, image_shapes[0])
        cnt = oc.tolist()                     # the one host read of the inference path
        return ([ob[i, :c] for i, c in enumerate(cnt)], [os_[i, :c] for i, c in enumerate(cnt)],
                [ol[i, :c] for i, c in enumerate(cnt)])

    @torch.no_grad()
    def inference_step(self, images: Tensor, **kwargs) -> Dict[str, Any]:
        pred_detection, anchors, pred_seg = self(images)
        return self.postprocess_for_inference(images=images, pred_detection=pred_detection, pred_seg=pred_seg,
                                              anchors=anchors)
