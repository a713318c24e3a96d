"""The module-level plug-in EXECUTED against the unmodified reference:
`nndetection_b200.ptmodule.register_with_reference()` subclasses the reference's `RetinaUNetModule`
(nndet/ptmodule/retinaunet/base.py:75-85: class-attribute injection) and registers it in `MODULE_REGISTRY`
(nndet/ptmodule/__init__.py:4, nndet/utils/registry.py:17-46), the table hydra's `module=<ClassName>` (conf/train/v001.yaml:5,
scripts/train.py:237) and the inference loader (nndet/inference/loading.py:82-92) look models up in.  pytorch_lightning & co. are absent
in this image: permissive stand-ins (tests/ref_stubs.py) let the reference's own modules import; the registry, the module skeleton, its
`from_config_plan` hook and the ensembler / predictor hooks are the reference's real code.  Runs in a subprocess (the stubs stay out of
the other tests' interpreter); skipped where /root/reference is absent (GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/nndet"), reason="needs the reference checkout (build container only)")
def test_register_with_reference_executes_against_the_unmodified_reference():
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import ref_stubs
        nndet = ref_stubs.import_reference()
        from nndet.ptmodule import MODULE_REGISTRY
        import nndet.ptmodule.retinaunet.v001 as ref_v001            # the reference's own registration runs too
        assert MODULE_REGISTRY.get("RetinaUNetV001") is ref_v001.RetinaUNetV001
        from nndet.ptmodule.retinaunet.base import RetinaUNetModule
        import nndetection_b200.ptmodule as P
        cls = P.register_with_reference()
        # registry: lookup by name like scripts/train.py / inference/loading.py do; duplicates raise like the reference's own
        assert MODULE_REGISTRY["RetinaUNetV001B200"] is cls and MODULE_REGISTRY.get("RetinaUNetV001B200") is cls
        assert issubclass(cls, RetinaUNetModule)
        try:
            P.register_with_reference()
            raise SystemExit("second registration must raise TypeError (utils/registry.py:33-35)")
        except TypeError:
            pass
        # class-attribute injection points of base.py:75-85 carry THIS package's classes
        from nndetection_b200.arch import conv as C, net as N
        from nndetection_b200.core import boxes as B
        assert cls.base_conv_cls is C.ConvInstanceRelu and cls.head_conv_cls is C.ConvGroupRelu
        assert cls.block is N.StackedConvBlock2 and cls.encoder_cls is N.Encoder and cls.decoder_cls is N.UFPNModular
        assert cls.head_cls is N.DetectionHeadHNMNative and cls.head_classifier_cls is N.BCECLassifier
        assert cls.head_regressor_cls is N.GIoURegressor and cls.segmenter_cls is N.DiCESegmenterFgBg
        assert cls.matcher_cls is B.ATSSMatcher and cls.head_sampler_cls is B.HardNegativeSamplerBatched
        for k in ("base_conv_cls", "head_conv_cls", "block", "encoder_cls", "decoder_cls", "matcher_cls", "head_cls", "head_classifier_cls",
                  "head_regressor_cls", "head_sampler_cls", "segmenter_cls"):
            assert hasattr(RetinaUNetModule, k), k                   # every hook exists on the reference skeleton
        # from_config_plan hook (base.py:338-466) builds this package's network with the reference's state_dict keys
        from nndetection_b200.configs import make_plan
        from nndetection_b200.core.retina import BaseRetinaNet
        arch, anc, patch, bs = make_plan("tiny")
        net = cls.from_config_plan(P.V001_MODEL_CFG, arch, anc)
        ref_net = ref_v001.RetinaUNetV001.from_config_plan(
            dict(P.V001_MODEL_CFG, head_kwargs={}, matcher_kwargs={"num_candidates": 4, "center_in_gt": False}), dict(arch), dict(anc))
        assert isinstance(net, BaseRetinaNet)
        mine, theirs = net.state_dict(), ref_net.state_dict()
        assert list(mine.keys()) == list(theirs.keys())
        assert all(tuple(mine[k].shape) == tuple(theirs[k].shape) for k in mine)
        # inference hooks (base.py:677-745) route to the device-resident ensembler
        from nndetection_b200.inference.ensembler import BoxEnsemblerSelective
        assert cls.get_ensembler_cls("boxes", 3) is BoxEnsemblerSelective
        print("REGISTRY_OK", len(mine))
    """) % (ROOT, os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert r.returncode == 0 and "REGISTRY_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
