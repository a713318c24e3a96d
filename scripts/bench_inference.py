"""BASELINE config 4 through the inference-side rows: sliding-window prediction of one synthetic case with the
LUNA-shaped network at 160^3 patches, 8 mirror passes, device-resident ensembling (NMS + WBC kernels).  Not the headline bench
(`bench.py` is); prints one JSON line: patches/s through `SlidingWindowPredictor.predict_case` (H2D of the case included), the
post-processing share, and the detections found.

    python scripts/bench_inference.py [--case 288 288 288] [--tta 8] [--batch 2] [--reps 3]
    python -m torch.distributed.run --nproc-per-node N ... scripts/bench_inference.py    # tiles sharded over ranks
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, nargs=3, default=[288, 288, 288])
    ap.add_argument("--tta", type=int, default=8)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from nndetection_b200.configs import make_plan
    from nndetection_b200.ptmodule import RetinaUNetV001
    from nndetection_b200.inference import BoxEnsemblerSelective, SlidingWindowPredictor
    arch, anc, patch, _ = make_plan("infer160")
    torch.manual_seed(1234)
    net = RetinaUNetV001.from_config_plan(None, arch, anc).to(dev).eval()
    g = torch.Generator().manual_seed(5)
    case = {"data": torch.randn(1, *args.case, generator=g).pin_memory()}
    pred = SlidingWindowPredictor(
        ensembler_fn=lambda c, properties=None: BoxEnsemblerSelective.from_case(c, properties, device=dev),
        models=[net], crop_size=patch, overlap=0.5, num_tta_transforms=args.tta, batch_size=args.batch, device=dev, shard=(rank, world))
    n_tiles = len(pred.tile_case({"data": torch.empty(1, *args.case, device="meta")}))
    out = pred.predict_case(case)                      # warm-up (plans, anchors, allocator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        out = pred.predict_case(case)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / args.reps
    if rank == 0:
        t1 = time.perf_counter()
        res = pred.ensembler.get_case_result()
        torch.cuda.synchronize()
        post = time.perf_counter() - t1
        # the post-processing SWEEP of the reference (ensembler/detection.py:975-995 through sweeper.py:108-213): the whole-case NMS + WBC
        # re-run for every candidate value of every swept parameter (6 + 2 + 6 + 7 + 7 = 28 settings) on the case's saved tile detections
        _, sweep = BoxEnsemblerSelective.sweep_parameters()
        base = dict(pred.ensembler.parameters)
        t2 = time.perf_counter()
        n_settings, n_boxes = 0, 0
        for key, values in sweep.items():
            for v in values:
                pred.ensembler.update_parameters(**{key: v})
                r = pred.ensembler.get_case_result()
                n_settings += 1; n_boxes += int(r["pred_boxes"].shape[0])
            pred.ensembler.update_parameters(**{key: base[key]})
        torch.cuda.synchronize()
        sweep_s = time.perf_counter() - t2
        cand = sum(int(b.shape[0]) for m in pred.ensembler.model_results.values() for b in m["boxes"])
        print(json.dumps({"metric": "sliding-window inference, 160^3 patches", "value": n_tiles * len(pred.tta_dims) / dt, "unit": "patches/s",
                          "sweep": {"settings": n_settings, "s_total": sweep_s, "s_per_setting": sweep_s / max(n_settings, 1),
                                    "tile_detections_in": cand, "candidate_boxes_per_s": cand * n_settings / max(sweep_s, 1e-9)},
                          "n_gpus": world, "case": args.case, "tiles": n_tiles, "tta": len(pred.tta_dims), "batch": args.batch,
                          "s_per_case": dt, "s_whole_case_nms_wbc": post, "detections": int(res["pred_boxes"].shape[0]),
                          "data": "synthetic", "weights": "random init"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
