"""NCCL check of the gradient buckets (torchrun, 2+ ranks): the same step from the same weights on the same per-rank batches, once with
ONE all-reduce of the flat gradient buffer after backward and once with 25 MB buckets all-reduced asynchronously during backward (side
streams of the coarse pyramid levels joined before each collective): the all-reduced gradient buffers must agree up to the order of
the fp32 atomics inside the weight-gradient kernels (run-to-run noise of the SAME mode is printed beside it).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ddp_bucket_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from nndetection_b200.arch.conv import set_grad_observer
    from nndetection_b200.configs import make_plan, synth_batch
    from nndetection_b200.ptmodule import RetinaUNetV001
    from nndetection_b200.training import Trainer
    arch, anc, patch, bs = make_plan("toy")
    images, tg = synth_batch(patch, bs, arch["in_channels"], arch["classifier_classes"], 100 + rank)
    tgd = {"target_boxes": [b.to(dev) for b in tg["target_boxes"]], "target_classes": [c.to(dev) for c in tg["target_classes"]],
           "target_seg": tg["target_seg"].to(dev)}
    grads = {}
    for mode in ("flat", "flat_again", "buckets"):
        torch.manual_seed(7)
        net = RetinaUNetV001.from_config_plan(None, arch, anc).to(dev)
        set_grad_observer(None)
        tr = Trainer(net, distributed=True, bucket_mb=(2.0 if mode == "buckets" else None))
        tr.fp.zero_grad()
        net.train()
        if tr.buckets is not None:
            tr.buckets.begin()
        losses, _ = net.train_step(images.to(dev), tgd, evaluation=False, batch_num=0)
        sum(losses.values()).backward()
        if tr.buckets is not None:
            tr.buckets.finish()
            n_b = len(tr.buckets.bounds)
        else:
            dist.all_reduce(tr.fp.grad)
        torch.cuda.synchronize()
        grads[mode] = tr.fp.grad.clone()
    rel = lambda a, b: float((a - b).double().norm() / b.double().norm())
    noise, diff = rel(grads["flat_again"], grads["flat"]), rel(grads["buckets"], grads["flat"])
    ok = diff <= max(10 * noise, 1e-4)
    t = torch.tensor([noise, diff, float(ok)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN if False else dist.ReduceOp.MAX)
    if rank == 0:
        print(f"BUCKET_CHECK world={world} buckets={n_b} run_to_run_noise={noise:.3e} buckets_vs_flat={diff:.3e} ok={bool(ok)}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
