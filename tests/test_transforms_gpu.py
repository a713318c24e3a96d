"""GPU parity of the fused instance -> targets kernel (nnd_instances_to_targets) against the executed reference chain
(tests/golden/transforms.npz) and the oracle: ids, box corners, classes and the semantic map are bit-exact."""
import zlib

import numpy as np
import pytest
import torch

from oracle import transform_oracle as to
import tutil as util

pytestmark = pytest.mark.gpu


def test_golden_chain_bit_exact():
    from nndetection_b200.io import FindInstances, Instances2Boxes, Instances2Segmentation
    g = util.golden("transforms")
    for ci, (B, shape, seed) in enumerate(util.TRANSFORM_CASES):
        t, maps = util.synth_instances(B, shape, seed)
        data = {"target": torch.from_numpy(t).cuda(), "instance_mapping": maps, "data": None}
        # the composition of nndet/ptmodule/retinaunet/base.py:114-134
        data = FindInstances(instance_key="target", save_key="present_instances")(**data)
        data = Instances2Boxes(instance_key="target", map_key="instance_mapping", box_key="boxes", class_key="classes",
                               present_instances="present_instances")(**data)
        data = Instances2Segmentation(instance_key="target", map_key="instance_mapping", present_instances="present_instances")(**data)
        for b in range(B):
            assert np.array_equal(data["present_instances"][b].cpu().numpy(), g[f"c{ci}_ids{b}"])
            bx = data["boxes"][b].cpu().numpy()
            assert bx.shape == g[f"c{ci}_boxes{b}"].shape and np.array_equal(bx, g[f"c{ci}_boxes{b}"])
            assert np.array_equal(data["classes"][b].cpu().numpy(), g[f"c{ci}_classes{b}"])
        sem = data["target"].cpu().numpy()
        assert sem.dtype == np.float32 and zlib.crc32(sem.tobytes()) == int(g[f"c{ci}_sem_crc"][0])
        assert "_nnd_b200_instances" not in data


def test_full_size_patch_vs_oracle_and_errors():
    from nndetection_b200.io import instances_to_targets
    t, maps = util.synth_instances(4, (128, 128, 128), 11, nmax=8)            # BASELINE configs[1] target shape
    p, bx, cl, sem = instances_to_targets(torch.from_numpy(t).cuda(), maps)
    po, bo_, co, so = to.pre_trafo(t, maps)
    for b in range(4):
        assert np.array_equal(p[b].cpu().numpy(), po[b]) and np.array_equal(cl[b].cpu().numpy(), co[b])
        assert bx[b].shape == bo_[b].shape and np.array_equal(bx[b].cpu().numpy(), bo_[b])
    assert np.array_equal(sem.cpu().numpy(), so)
    with pytest.raises(KeyError):                                               # id in the volume but not in the mapping
        bad = torch.zeros(1, 1, 8, 8, 8).cuda(); bad[0, 0, 1:3, 1:3, 1:3] = 7
        instances_to_targets(bad, [{"3": 0}])
    with pytest.raises(RuntimeError):
        instances_to_targets(torch.zeros(1, 1, 8, 8, 8), [{}])
