"""GPU parity of the on-device weighted box clustering (nnd_wbc3d_f32) against the reference's golden outputs and the oracle.
Cluster membership is exact (same boxes, same order); consolidated scores / coordinates are fp32 sums accumulated with
atomics in a different order than torch.sum -> 1e-5 relative (BASELINE north_star: fp32 box coords within 1e-4)."""
import pytest
import torch

from oracle import box_oracle as bo
import tutil as util

pytestmark = pytest.mark.gpu


def _close(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)


def test_golden_cases():
    from nndetection_b200.inference import wbc
    g = util.golden("wbc")
    for i, (n, seed, thr, st, ua, mw) in enumerate(g["cases"].tolist()):
        n, seed = int(n), int(seed)
        b, s, w, ne = util.wbc_case(n, seed, extent=60.0 if n <= 1000 else 100.0)
        ob, os_ = wbc(b.cuda(), s.cuda(), w.cuda(), ne.cuda(), thr, st, use_area=bool(ua), missing_weight=mw)
        _close(ob.cpu(), torch.from_numpy(g[f"c{i}_boxes"]))
        _close(os_.cpu(), torch.from_numpy(g[f"c{i}_scores"]))


def test_batched_golden_and_edge_cases():
    from nndetection_b200.inference import batched_wbc, wbc
    g = util.golden("wbc")
    b, s, w, ne = util.wbc_case(600, 9)
    lab = torch.from_numpy(g["batched_labels_in"])
    ob, os_, ol = batched_wbc(b.cuda(), s.cuda(), lab.cuda(), w.cuda(), 0.2, ne.cuda(), 0.02, use_area=True, missing_weight=1.0)
    _close(ob.cpu(), torch.from_numpy(g["batched_boxes"]))
    _close(os_.cpu(), torch.from_numpy(g["batched_scores"]))
    assert torch.equal(ol.cpu(), torch.from_numpy(g["batched_labels"]))
    # empty input, CPU tensors, a zero-volume box (NaN IoU with itself: vanishes, wbc.py:133-153)
    e = wbc(torch.zeros(0, 6).cuda(), torch.zeros(0).cuda(), torch.zeros(0).cuda(), torch.zeros(0).cuda(), 0.1, 0.0)
    assert e[0].shape == (0, 6) and e[1].shape == (0,)
    with pytest.raises(RuntimeError):
        wbc(torch.zeros(3, 6), torch.zeros(3), torch.zeros(3), torch.zeros(3), 0.1, 0.0)
    bz = torch.tensor([[1., 1, 1, 1, 1, 1], [0., 0, 2, 2, 0, 2], [0.1, 0, 2, 2, 0, 2]])
    sz = torch.tensor([0.9, 0.8, 0.7]); wz = torch.ones(3); nz = torch.full((3,), 2.0)
    ob, os_ = wbc(bz.cuda(), sz.cuda(), wz.cuda(), nz.cuda(), 0.3, 0.0)
    rb, rs = bo.wbc(bz, sz, wz, nz, 0.3, 0.0)
    _close(ob.cpu(), rb); _close(os_.cpu(), rs)


@pytest.mark.parametrize("n,thr", [(10000, 0.1), (20000, 0.3)])
def test_vs_oracle_large(n, thr):
    """Multi-chunk scan (> 8192 boxes): first-remover-wins across column chunks."""
    from nndetection_b200.inference import wbc
    b, s, w, ne = util.wbc_case(n, 77, extent=160.0)
    ob, os_ = wbc(b.cuda(), s.cuda(), w.cuda(), ne.cuda(), thr, 0.01, use_area=True, missing_weight=0.8)
    rb, rs = bo.wbc(b, s, w, ne, thr, 0.01, use_area=True, missing_weight=0.8)
    _close(ob.cpu(), rb); _close(os_.cpu(), rs)
