"""Module-level plug-in: the v001 class set wired to the B200 implementations.

Mirrors the assembly of `RetinaUNetModule.from_config_plan` (nndet/ptmodule/retinaunet/base.py:338-466) with the
class attributes of `RetinaUNetV001` (nndet/ptmodule/retinaunet/v001.py:29-38) and the `model_cfg` defaults of
nndet/conf/train/v001.yaml:61-110.  When the reference package (with pytorch_lightning) is importable,
`register_with_reference()` registers a `RetinaUNetV001B200` LightningModule subclass in `MODULE_REGISTRY` so
`module=RetinaUNetV001B200` selects it from the unmodified trainer.
"""
import copy
from typing import Optional

from .arch.conv import ConvGroupRelu, ConvInstanceRelu, Generator
from .arch.net import (BCECLassifier, DetectionHeadHNMNative, DiCESegmenterFgBg, Encoder, GIoURegressor,
                       StackedConvBlock2, UFPNModular)
from .core.boxes import ATSSMatcher, BoxCoderND, HardNegativeSamplerBatched, box_iou, get_anchor_generator
from .core.retina import BaseRetinaNet

V001_MODEL_CFG = {
    "encoder_kwargs": {},
    "decoder_kwargs": {"min_out_channels": 8, "upsampling_mode": "transpose", "num_lateral": 1, "norm_lateral": False,
                       "activation_lateral": False, "num_out": 1, "norm_out": False, "activation_out": False},
    "head_kwargs": {},
    "head_classifier_kwargs": {"num_convs": 1, "norm_channels_per_group": 16, "norm_affine": True, "reduction": "mean",
                               "loss_weight": 1., "prior_prob": 0.01},
    "head_regressor_kwargs": {"num_convs": 1, "norm_channels_per_group": 16, "norm_affine": True, "reduction": "sum",
                              "loss_weight": 1., "learn_scale": True},
    "head_sampler_kwargs": {"batch_size_per_image": 32, "positive_fraction": 0.33, "pool_size": 20, "min_neg": 1},
    "segmenter_kwargs": {"dice_kwargs": {"batch_dice": True}},
    "matcher_kwargs": {"num_candidates": 4, "center_in_gt": False},
    "plan_arch_overwrites": {},
    "plan_anchors_overwrites": {},
}


class RetinaUNetV001:
    """Class-attribute injection points identical to nndet/ptmodule/retinaunet/base.py:75-85 / v001.py:31-38."""
    base_conv_cls = ConvInstanceRelu
    head_conv_cls = ConvGroupRelu
    block = StackedConvBlock2
    encoder_cls = Encoder
    decoder_cls = UFPNModular
    matcher_cls = ATSSMatcher
    head_cls = DetectionHeadHNMNative
    head_classifier_cls = BCECLassifier
    head_regressor_cls = GIoURegressor
    head_sampler_cls = HardNegativeSamplerBatched
    segmenter_cls = DiCESegmenterFgBg

    @classmethod
    def from_config_plan(cls, model_cfg: Optional[dict], plan_arch: dict, plan_anchors: dict, log_num_anchors=None,
                         **kwargs) -> BaseRetinaNet:
        model_cfg = copy.deepcopy(V001_MODEL_CFG if model_cfg is None else model_cfg)
        plan_arch, plan_anchors = dict(plan_arch), copy.deepcopy(dict(plan_anchors))
        plan_arch.update(model_cfg["plan_arch_overwrites"])
        plan_anchors.update(model_cfg["plan_anchors_overwrites"])
        dim = plan_arch["dim"]
        coder = BoxCoderND(weights=(1.0,) * (dim * 2))
        s_param = not (("aspect_ratios" in plan_anchors) and (plan_anchors["aspect_ratios"] is not None))
        anchor_generator = get_anchor_generator(dim, s_param=s_param)(**plan_anchors)

        conv = Generator(cls.base_conv_cls, dim)
        encoder = cls.encoder_cls(conv=conv, conv_kernels=plan_arch["conv_kernels"], strides=plan_arch["strides"],
                                  block_cls=cls.block, in_channels=plan_arch["in_channels"],
                                  start_channels=plan_arch["start_channels"], stage_kwargs=None,
                                  max_channels=plan_arch.get("max_channels", 320), **model_cfg["encoder_kwargs"])
        decoder = cls.decoder_cls(conv=Generator(cls.base_conv_cls, dim), conv_kernels=plan_arch["conv_kernels"],
                                  strides=encoder.get_strides(), in_channels=encoder.get_channels(),
                                  decoder_levels=plan_arch["decoder_levels"], fixed_out_channels=plan_arch["fpn_channels"],
                                  **model_cfg["decoder_kwargs"])
        matcher = cls.matcher_cls(similarity_fn=box_iou, **model_cfg["matcher_kwargs"])
        hconv = Generator(cls.head_conv_cls, dim)
        apos = anchor_generator.num_anchors_per_location()[0]
        classifier = cls.head_classifier_cls(conv=hconv, in_channels=plan_arch["fpn_channels"],
                                             internal_channels=plan_arch["head_channels"],
                                             num_classes=plan_arch["classifier_classes"], anchors_per_pos=apos,
                                             num_levels=len(plan_arch["decoder_levels"]),
                                             **model_cfg["head_classifier_kwargs"])
        regressor = cls.head_regressor_cls(conv=hconv, in_channels=plan_arch["fpn_channels"],
                                           internal_channels=plan_arch["head_channels"], anchors_per_pos=apos,
                                           num_levels=len(plan_arch["decoder_levels"]),
                                           **model_cfg["head_regressor_kwargs"])
        sampler = cls.head_sampler_cls(**model_cfg["head_sampler_kwargs"])
        head = cls.head_cls(classifier=classifier, regressor=regressor, coder=coder, sampler=sampler,
                            log_num_anchors=None, **model_cfg["head_kwargs"])
        segmenter = None
        if cls.segmenter_cls is not None:
            segmenter = cls.segmenter_cls(Generator(cls.base_conv_cls, dim), seg_classes=plan_arch["seg_classes"],
                                          in_channels=decoder.get_channels(), decoder_levels=plan_arch["decoder_levels"],
                                          **model_cfg["segmenter_kwargs"])
        return BaseRetinaNet(dim=dim, encoder=encoder, decoder=decoder, head=head, anchor_generator=anchor_generator,
                             matcher=matcher, num_classes=plan_arch["classifier_classes"],
                             decoder_levels=plan_arch["decoder_levels"], segmenter=segmenter,
                             detections_per_img=plan_arch.get("detections_per_img", 100),
                             score_thresh=plan_arch.get("score_thresh", 0),
                             topk_candidates=plan_arch.get("topk_candidates", 10000),
                             remove_small_boxes=plan_arch.get("remove_small_boxes", 0.01),
                             nms_thresh=plan_arch.get("nms_thresh", 0.6))


def register_with_reference():
    """Register a LightningModule subclass backed by these classes in the reference's MODULE_REGISTRY
    (nndet/ptmodule/__init__.py:4, nndet/utils/registry.py:17-46).  Needs `nndet` + pytorch_lightning importable."""
    from nndet.ptmodule import MODULE_REGISTRY
    from nndet.ptmodule.retinaunet.base import RetinaUNetModule

    class RetinaUNetV001B200(RetinaUNetModule):
        base_conv_cls = RetinaUNetV001.base_conv_cls
        head_conv_cls = RetinaUNetV001.head_conv_cls
        block = RetinaUNetV001.block
        encoder_cls = RetinaUNetV001.encoder_cls
        decoder_cls = RetinaUNetV001.decoder_cls
        matcher_cls = RetinaUNetV001.matcher_cls
        head_cls = RetinaUNetV001.head_cls
        head_classifier_cls = RetinaUNetV001.head_classifier_cls
        head_regressor_cls = RetinaUNetV001.head_regressor_cls
        head_sampler_cls = RetinaUNetV001.head_sampler_cls
        segmenter_cls = RetinaUNetV001.segmenter_cls
        from_config_plan = classmethod(lambda c, *a, **k: RetinaUNetV001.from_config_plan.__func__(c, *a, **k))

        # inference side: device-resident ensembler and predictor behind the reference's hooks
        # (`get_ensembler_cls` base.py:677-695, `get_predictor` :697-745); the reference's `predict_dir` / `sweep` call them unchanged
        @staticmethod
        def get_ensembler_cls(key, dim):
            if dim == 3 and key == "boxes":
                from .inference.ensembler import BoxEnsemblerSelective
                return BoxEnsemblerSelective
            return RetinaUNetModule.get_ensembler_cls(key, dim)

        @classmethod
        def get_predictor(cls, plan, models, num_tta_transforms=None, do_seg=False, **kwargs):
            from .inference.helper import get_predictor
            return get_predictor(plan, [getattr(m, "model", m) for m in models], num_tta_transforms, do_seg,
                                 ensembler_cls=cls.get_ensembler_cls("boxes", plan["network_dim"]), **kwargs)

    MODULE_REGISTRY.register(RetinaUNetV001B200)
    return RetinaUNetV001B200
