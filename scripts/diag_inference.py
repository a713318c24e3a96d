"""Bisect aid for tests/test_zz_inference_gpu.py::test_device_pipeline_with_the_network_matches_cpu_pipeline (GPU box):
runs the same pair of pipelines and prints where the case results part ways, plus tie statistics of the per-tile detections.
TEST INFRASTRUCTURE (imports oracle/ through the test module)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_zz_inference_gpu as T        # noqa: E402


def main():
    from nndetection_b200.configs import make_plan
    from nndetection_b200.ptmodule import RetinaUNetV001
    arch, anc, patch, _ = make_plan("toy")
    torch.manual_seed(3)
    net = RetinaUNetV001.from_config_plan(None, arch, anc).cuda().eval()
    g = torch.Generator().manual_seed(23)
    case = {"data": torch.randn(1, 48, 96, 80, generator=g)}
    out_d, out_c, rec = T._run_pair(net, case, patch, 2, run_on_host=False)
    print("calls", len(rec.calls), "boxes/call", [sum(len(b) for b in c["pred_boxes"]) for c in rec.calls][:16])
    alls = torch.cat([s for c in rec.calls for s in c["pred_scores"]])
    print("per-tile detections:", alls.numel(), "distinct scores:", alls.unique().numel(), "min/max", float(alls.min()), float(alls.max()))
    for k in ("pred_boxes", "pred_scores", "pred_labels"):
        d, c = out_d[k].cpu(), out_c[k]
        print(k, "dev", tuple(d.shape), "cpu", tuple(c.shape))
    n = min(out_d["pred_scores"].shape[0], out_c["pred_scores"].shape[0])
    sd, sc = out_d["pred_scores"].cpu()[:n], out_c["pred_scores"][:n]
    bd, bc = out_d["pred_boxes"].cpu()[:n], out_c["pred_boxes"][:n]
    ld, lc = out_d["pred_labels"].cpu()[:n], out_c["pred_labels"][:n]
    bad = torch.where(~(torch.isclose(sd, sc, rtol=1e-5, atol=1e-6) & torch.isclose(bd, bc, rtol=1e-5, atol=1e-4).all(1) & (ld == lc)))[0]
    print("rows compared", n, "mismatching rows", bad.numel(), "first", bad[:10].tolist())
    for i in bad[:6].tolist():
        print(i, "dev", sd[i].item(), ld[i].item(), bd[i].tolist())
        print(i, "cpu", sc[i].item(), lc[i].item(), bc[i].tolist())
    # are the two results the same SET (order / tie effects only)?
    def key(b, s, l):
        return sorted((round(float(x), 4), int(y), tuple(round(float(v), 2) for v in z)) for x, y, z in zip(s, l, b))
    kd, kc = key(out_d["pred_boxes"].cpu(), out_d["pred_scores"].cpu(), out_d["pred_labels"].cpu()), key(out_c["pred_boxes"], out_c["pred_scores"], out_c["pred_labels"])
    print("same multiset (rounded):", kd == kc, "only dev", len(set(kd) - set(kc)), "only cpu", len(set(kc) - set(kd)))
    print("sorted desc dev", bool((sd[:-1] >= sd[1:]).all()), "cpu", bool((sc[:-1] >= sc[1:]).all()))
    print("score ties within result (dev):", n - sd.unique().numel())


if __name__ == "__main__":
    main()
