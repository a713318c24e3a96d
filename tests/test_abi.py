"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/nndet_b200.h declares;
the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from nndetection_b200 import build
    lib = build.build()
    h = ctypes.CDLL(lib)
    header = open(os.path.join(ROOT, "include", "nndet_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(nnd_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 35
    missing = [n for n in declared if not hasattr(h, n)]
    assert not missing, missing
    h.nnd_build_arch.restype = ctypes.c_char_p
    assert h.nnd_build_arch() == b"sm_100a"
    assert h.nnd_abi_version() == 1


def test_no_cpu_fallback():
    from nndetection_b200 import _C
    from nndetection_b200.core.boxes import engine
    with pytest.raises(RuntimeError):
        _C.nms(torch.zeros(4, 6), torch.zeros(4), 0.5)
    with pytest.raises(RuntimeError):
        engine.pairwise(torch.zeros(4, 6), torch.zeros(4, 6), 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nndetection_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_sampler_plan_and_state_dict_contract():
    from nndetection_b200.configs import make_plan
    from nndetection_b200.core.boxes.engine import SamplerPlan
    from nndetection_b200.ptmodule import RetinaUNetV001
    from oracle import box_oracle as bo, model_oracle as mo
    p = SamplerPlan(4)
    assert p.max_pos == 42 and (p.max_pos, p.max_neg, p.max_pool) == bo.hnm_counts(10 ** 6, 10 ** 9, 4)
    arch, anc, patch, bs = make_plan("luna")
    net = RetinaUNetV001.from_config_plan(None, arch, anc)
    orc = mo.RetinaUNetOracle(dict(arch), dict(anc))
    a, b = net.state_dict(), orc.state_dict()       # oracle keys == reference keys (asserted in scripts/gen_golden.py)
    assert sorted(a) == sorted(b) and len(a) == 92
    assert all(a[k].shape == b[k].shape for k in a)
    net.load_state_dict(b)                          # reference-layout checkpoints load
    import copy
    copy.deepcopy(net)                              # planner deep-copies the model (nndet/planning/estimator.py:130)


def test_conv_plan_geometry():
    from nndetection_b200.arch.conv_ops import ConvPlan
    p = ConvPlan(2, 32, 64, (12, 16, 20), 3, 2, 1, False)
    assert p.out_sp == (6, 8, 10) and len(p.fprop) == 1 and p.fprop[0][20] == 27
    assert len(p.dgrad) == 8 and sorted(g[20] for g in p.dgrad) == [1, 2, 2, 2, 4, 4, 4, 8] and p.dgrad_covers_all
    t = ConvPlan(2, 64, 32, (4, 6, 8), 2, 2, 0, True)
    assert t.out_sp == (8, 12, 16) and len(t.fprop) == 8 and t.dgrad[0][20] == 8
    a = ConvPlan(1, 64, 128, (6, 12, 12), 3, (1, 2, 2), 1, False)
    assert a.out_sp == (6, 6, 6) and len(a.dgrad) == 4
