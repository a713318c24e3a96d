"""GPU parity: nndetection_b200._C.nms (C ABI: nnd_nms3d_f32) vs the oracle and the reference's golden keep lists."""
import numpy as np
import pytest
import torch

from oracle import box_oracle as bo
import tutil as util

pytestmark = pytest.mark.gpu


def _nms(boxes, scores, thr):
    from nndetection_b200 import _C
    return _C.nms(boxes.cuda(), scores.cuda(), thr).cpu()


def test_golden_keep_lists_bit_exact():
    g = util.golden("nms")
    for n, thr in g["cases"].tolist():
        n = int(n)
        boxes, scores = util.nms_case(n)
        keep = _nms(boxes, scores, thr)
        assert keep.dtype == torch.int64
        assert torch.equal(keep, torch.from_numpy(g[f"n{n}_t{thr}_keep"])), (n, thr)


@pytest.mark.parametrize("n,thr", [(10000, 0.1), (10000, 0.6), (20000, 0.5)])
def test_vs_oracle_large(n, thr):
    boxes, scores = util.nms_case(n)
    assert torch.equal(_nms(boxes, scores, thr), bo.nms_greedy(boxes, scores, thr))


def test_batched_golden():
    from nndetection_b200.core.boxes.nms import batched_nms
    g = util.golden("nms")
    gen = torch.Generator().manual_seed(4242)
    boxes = util.rand_boxes(1500, gen); scores = util.unique_scores(1500, gen)
    idxs = torch.randint(0, 3, (1500,), generator=gen)
    keep = batched_nms(boxes.cuda(), scores.cuda(), idxs.cuda(), 0.5).cpu()
    assert torch.equal(keep, torch.from_numpy(g["batched_keep"]))


def test_edge_cases():
    from nndetection_b200 import _C
    e = _C.nms(torch.zeros(0, 6).cuda(), torch.zeros(0).cuda(), 0.5)
    assert e.shape == (0,) and e.dtype == torch.int64 and e.is_cuda
    with pytest.raises(RuntimeError):
        _C.nms(torch.zeros(3, 6), torch.zeros(3), 0.5)
    # zero-volume duplicates: NaN IoU never suppresses (CUDA semantics of the reference kernel)
    b = torch.tensor([[1., 1, 1, 1, 1, 1], [1., 1, 1, 1, 1, 1], [0., 0, 2, 2, 0, 2]])
    s = torch.tensor([0.9, 0.8, 0.7])
    assert _nms(b, s, 0.5).tolist() == [0, 1, 2]
    # equal scores: stable order (ascending index)
    b = util.rand_boxes(500, torch.Generator().manual_seed(1))
    s = torch.full((500,), 0.5)
    assert torch.equal(_nms(b, s, 0.3), bo.nms_greedy(b, s, 0.3))
    # negative threshold: every later box with non-NaN IoU is suppressed
    assert _nms(b, util.unique_scores(500, torch.Generator().manual_seed(2)), -1.0).numel() == 1
    # 2-D boxes
    b2 = b[:, :4].contiguous()
    s2 = util.unique_scores(500, torch.Generator().manual_seed(3))
    import torchvision
    assert torch.equal(_nms(b2, s2, 0.4), torchvision.ops.nms(b2, s2, 0.4))


def test_idempotent_and_sorted_property_100k():
    # size-independent properties at BASELINE config 4's N: keep is score-sorted, NMS(keep) == keep
    n = 100_000
    g = torch.Generator().manual_seed(9)
    boxes, scores = util.rand_boxes(n, g), util.unique_scores(n, g)
    keep = _nms(boxes, scores, 0.1)
    ks = scores[keep]
    assert (ks[:-1] > ks[1:]).all()
    keep2 = _nms(boxes[keep], ks, 0.1)
    assert torch.equal(keep2, torch.arange(keep.numel()))
    # pairwise IoU among the kept set never exceeds the threshold (checked on a slice)
    sub = boxes[keep[:3000]]
    iou = bo.box_iou(sub, sub)
    iou.fill_diagonal_(0)
    assert float(iou.max()) <= 0.1


@pytest.mark.parametrize("n,thr,mk", [(10000, 0.6, 100), (10000, 0.1, 100), (3000, 0.5, 1), (20000, 0.3, 2500)])
def test_prefix_variant_equals_prefix_of_full_result(n, thr, mk):
    """nnd_nms3d_topk_f32 (early-exit scan used by the detection post-processing for keep[:detections_per_img])."""
    from ctypes import c_float, c_longlong, c_size_t
    from nndetection_b200 import _lib as L
    boxes, scores = util.nms_case(n)
    full = _nms(boxes, scores, thr)
    lib = L.lib()
    b, s = boxes.cuda().contiguous(), scores.cuda().contiguous()
    ws_bytes = lib.nnd_nms_workspace_bytes(n, 3)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    keep = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    L.check(lib.nnd_nms3d_topk_f32(L.ptr(b), L.ptr(s), c_longlong(n), c_float(thr), c_longlong(mk), L.ptr(keep), L.ptr(cnt),
                                   L.ptr(ws), c_size_t(ws_bytes), L.stream_ptr()), "nnd_nms3d_topk_f32")
    k = min(int(cnt.item()), mk)
    assert k == min(mk, full.numel())
    assert torch.equal(keep[:k].cpu(), full[:k])


@pytest.mark.parametrize("n,thr", [(1000, 0.5), (10000, 0.1), (10000, 0.6), (100000, 0.1), (100000, 0.5)])
def test_ab_against_the_reference_cuda_nms(n, thr):
    """Head-to-head with the REFERENCE's own kernel (nndet/csrc/cuda/nms.cu:99-221 + ops.cpp, built unmodified except the two-token
    dispatch fix of nms.cu:172,182 by oracle/build_ref_nms.py into oracle/_ref/): identical keep lists on the SURVEY 8d stress boxes
    (unique scores, so the sort order is defined), 3-D and 2-D, fp32."""
    from oracle.build_ref_nms import load
    ref = load()
    if ref is None:
        pytest.skip("oracle/_ref/ref_nms*.so was not built (python oracle/build_ref_nms.py needs /root/reference)")
    g = torch.Generator().manual_seed(7000 + n)
    boxes, scores = util.rand_boxes(n, g), util.unique_scores(n, g)
    b, s = boxes.cuda(), scores.cuda()
    from nndetection_b200 import _C
    keep_ref = ref.nms(b, s, thr)
    keep = _C.nms(b, s, thr)
    assert keep_ref.dtype == keep.dtype == torch.int64 and keep.is_cuda
    assert torch.equal(keep, keep_ref), (n, thr, keep.numel(), keep_ref.numel())
    if n <= 10000:
        b2 = b[:, :4].contiguous()
        assert torch.equal(_C.nms(b2, s, thr), ref.nms(b2, s, thr))
