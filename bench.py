#!/usr/bin/env python
"""Benchmark of the hot path: full Retina U-Net train step (forward + losses + detection post-processing with 3-D NMS
+ backward + gradient all-reduce + SGD) on synthetic LUNA16-shaped patches (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "patches/sec (128^3 1ch) full train step + 3D NMS"
UNIT = "patches/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken inside [t0, t1] (wall clock of the timed region; all samples if none fall inside).
        The sampler is started BEFORE the warm-up so that nvidia-smi's start-up (driver queries that can stall launches for
        tens of ms) never lands inside the timed region."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r for t, r in self.rows if t0 is not None and t0 <= t <= t1]
        self.rows = inside if inside else [r for _, r in self.rows]
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(sm)}


def conv_flops_per_patch(arch, patch):
    """Closed-form forward conv FLOPs (2 * Cin * Cout * k^3 * output voxels), SURVEY 8d."""
    from nndetection_b200.ptmodule import RetinaUNetV001
    import numpy as np
    fl = 0
    sp = list(patch)
    chans, sps = [], []
    c = arch["in_channels"]
    for i, k in enumerate(arch["conv_kernels"]):
        co = arch["start_channels"] if i == 0 else min(c * 2, arch.get("max_channels", 320))
        if i > 0:
            sp = [a // b for a, b in zip(sp, arch["strides"][i - 1])]
        v = sp[0] * sp[1] * sp[2]
        kk = int(np.prod(k))
        fl += 2 * c * co * kk * v + 2 * co * co * kk * v
        c = co
        chans.append(co); sps.append(list(sp))
    n = len(chans)
    oc = [arch["fpn_channels"]] * n
    for ol in [l for l in range(n) if l < min(arch["decoder_levels"])][::-1]:
        oc[ol] = max(8, oc[ol + 1] // 2)
    for l in range(n):
        v = sps[l][0] * sps[l][1] * sps[l][2]
        fl += 2 * chans[l] * oc[l] * v + 2 * oc[l] * oc[l] * int(np.prod(arch["conv_kernels"][l])) * v
        if l > 0:
            fl += 2 * oc[l] * oc[l - 1] * int(np.prod(arch["strides"][l - 1])) * v
    hc, C = arch["head_channels"], arch["classifier_classes"]
    for l in arch["decoder_levels"]:
        v = sps[l][0] * sps[l][1] * sps[l][2]
        fl += 2 * (2 * 27 * (oc[l] * hc + hc * hc) * v)            # c_in + c_internal0, classifier and regressor
        fl += 2 * 27 * hc * (27 * C + 162) * v
    fl += 2 * oc[0] * 2 * sps[0][0] * sps[0][1] * sps[0][2]
    return fl


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port: identical torch-CPU operators), all host
    threads; each step = ONE 128^3 patch (bounded sample of the batch-4 workload) train step + post-processing."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import model_oracle as mo
    torch.set_num_threads(min(os.cpu_count(), 64))
    arch, anc, patch, bs = mo.make_plan("luna")
    net = mo.RetinaUNetOracle(dict(arch), dict(anc))
    images, targets = mo.synth_batch(patch, 1, arch["in_channels"], arch["classifier_classes"], 1234)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=3e-5)

    def step():
        opt.zero_grad()
        losses, aux = net.train_step(images, targets, seed=1)
        net.postprocess(images, {k: v.detach() for k, v in aux["pred"].items()}, aux["anchors"])
        sum(losses.values()).backward()
        opt.step()

    # bounded sample: at most 1 warm-up + args.steps steps, and stop after ~150 s of timed work (>= 1 step)
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    done = 0
    while done < args.steps and (done == 0 or time.perf_counter() - t0 < 150.0):
        step()
        done += 1
    dt = time.perf_counter() - t0
    v = done * 1 / dt
    out = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": min(args.warmup, 1),
           "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": {"workload": "luna16 128^3 1ch (reference CPU path, 1 patch per step)", "batch_per_step": 1},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                            "sample": f"{done} train steps of 1 patch (fwd+loss+postprocess/nms_cpu+bwd+SGD), torch CPU fp32, {torch.get_num_threads()} threads"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="luna")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="timed region only (for ncu): no e2e / roofline / cpu baseline")
    ap.add_argument("--igemm-only", action="store_true", help="disable the tcgen05 conv kernel (A/B)")
    ap.add_argument("--experimental", default="", help="comma list of A/B switches: mma_s2 (strided / transposed forms on the mma.sync kernels instead of tcgen05), tc_bulk (tile kernel weights via cp.async.bulk: deadlocks, do not use), norm_narrow (4-channel norm backward passes), no_buckets (N > 1: ONE all-reduce of the flat gradient after backward instead of 25 MB buckets all-reduced during it), no_streams (coarse pyramid levels on the main stream), tcs_map, no_pw_tma")
    ap.add_argument("--stock-tuned", action="store_true", help="internal: only the tuned stock-PyTorch comparator (cudnn.benchmark + channels_last_3d), one JSON line")
    ap.add_argument("--trace-layers", default=None, metavar="CSV",
                    help="after the timed regions run ONE extra step with the per-launch convolution trace on and write it here "
                         "(kernel chosen, layer geometry, ms, GFLOP per launch) -- maps the step time onto the network")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        args.steps = min(args.steps, 5)
        return run_reference(args)
    if args.stock_tuned:
        print(json.dumps(stock_gpu_baseline("cuda:0", args.config, tuned=True)), flush=True)
        return

    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # The contract is ONE JSON line on stdout: anything a library prints there while the job runs (NCCL's version banner at
    # communicator creation, for one) is sent to stderr instead; fd 1 is restored just before the result line.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from nndetection_b200 import _lib as L
    from nndetection_b200.configs import make_plan, synth_batch
    from nndetection_b200.ptmodule import RetinaUNetV001
    from nndetection_b200.training import Trainer
    from nndetection_b200.arch import conv_ops

    lib = L.lib()                       # raises if the CUDA extension is missing: no fallback
    lib.nnd_launch_count.restype = __import__("ctypes").c_ulonglong
    if args.igemm_only:
        conv_ops.set_tensor_path(False)
    if "mma_s2" in args.experimental.split(","):             # A/B: strided / transposed forms back on the mma.sync kernels (round-1 default)
        conv_ops.set_wgrad_strided_tc(False)
        conv_ops.set_gather_strided_tc(False)
    if "pw_tma" in args.experimental.split(","):             # TMA-fed pointwise GEMM for laterals / up-convolutions (conv_pw.cu)
        conv_ops.set_pointwise_tma(True)
    if "no_pw_tma" in args.experimental.split(","):
        conv_ops.set_pointwise_tma(False)
    for tok in args.experimental.split(","):                  # gather_tma=<mode>: TMA-fed tile kernel (bit 0 stride-1 forms, bit 1 stride-2 forms)
        if tok.startswith("gather_tma="):
            conv_ops.set_gather_tma(int(tok.split("=")[1]))
    if "no_wgrad_tma" in args.experimental.split(","):        # A/B: >= 64-channel stride-1 weight gradients back on the cp.async kernel
        conv_ops.set_wgrad_tma(0)
    if "tcs_map" in args.experimental.split(","):            # A/B: coalesced halo copy mapping in the streaming kernel
        L.lib().nnd_conv_set_tcs_map(1)
    if "tc_bulk" in args.experimental.split(","):
        conv_ops.set_tc_bulk(True)
    if "norm_narrow" in args.experimental.split(","):
        conv_ops.set_norm_bwd_narrow(True)
    arch, anc, patch, bs = make_plan(args.config)
    torch.manual_seed(1234 + rank)
    net = RetinaUNetV001.from_config_plan(None, arch, anc).to(dev)
    if "no_streams" in args.experimental.split(","):          # A/B: coarse pyramid levels / segmentation branch on the main stream
        net.head.parallel_levels = False
    trainer = Trainer(net, distributed=world > 1, bucket_mb=None if "no_buckets" in args.experimental.split(",") else 25.0)

    # ---- synthetic data: 4 distinct batches (> L2: one batch of activations alone is GBs), pinned on the host
    # as many distinct batches as warm-up steps (<= 4): every batch's allocation pattern (it depends on the number of ground-truth
    # boxes) is seen once before the timed region -- a first-time cudaMalloc inside it costs ~10 ms
    n_batches = max(1, min(4, args.warmup))
    host = []
    for i in range(n_batches):
        im, tg = synth_batch(patch, bs, arch["in_channels"], arch["classifier_classes"], 1234 + 97 * rank + i)
        host.append((im.pin_memory(), [b.pin_memory() for b in tg["target_boxes"]], [c.pin_memory() for c in tg["target_classes"]],
                     tg["target_seg"].pin_memory()))
    resident = [(im.to(dev), {"target_boxes": [b.to(dev) for b in tb], "target_classes": [c.to(dev) for c in tc],
                              "target_seg": sg.to(dev)}) for im, tb, tc, sg in host]
    h2d_bytes = host[0][0].numel() * 4 + host[0][3].numel() * 4 + sum(b.numel() * 4 for b in host[0][1]) + sum(c.numel() * 8 for c in host[0][2])

    def step_resident(i):
        im, tg = resident[i % n_batches]
        return trainer.train_step(im, tg, evaluation=True)

    def step_e2e(i):
        im, tb, tc, sg = host[i % n_batches]
        imd = im.to(dev, non_blocking=True)
        tg = {"target_boxes": [b.to(dev, non_blocking=True) for b in tb], "target_classes": [c.to(dev, non_blocking=True) for c in tc],
              "target_seg": sg.to(dev, non_blocking=True)}
        losses, pred = trainer.train_step(imd, tg, evaluation=True)
        vals = torch.stack([losses[k].detach().float() for k in ("reg", "cls", "seg_ce", "seg_dice")]).cpu()   # D2H result read
        nb = sum(int(b.shape[0]) for b in pred["pred_boxes"])
        return vals, nb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for i in range(args.warmup):
        step_resident(i)
    barrier()

    # ---- timed region 1: device-resident inputs
    l0 = lib.nnd_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_w0 = time.time()
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    barrier()
    t_w1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = lib.nnd_launch_count() - l0
    clk = clocks.stop(t_w0, t_w1) if rank == 0 else None

    if args.profile:
        sys.stdout.flush(); os.dup2(saved_stdout, 1)
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms / args.steps, "gpu_launches": int(launches)}), flush=True)
        return
    # ---- timed region 2: end to end through the public API with host buffers
    for i in range(2):
        step_e2e(i)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    d2h = 0
    for i in range(args.steps):
        vals, nb = step_e2e(i)
        d2h = 16 + 4 * bs + nb * (24 + 4 + 8)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- whole-step roofline: ONE extra step (outside the timed regions) with the per-launch convolution trace on
    roof_step = None
    if rank == 0 and world == 1:
        import tempfile
        trace_path = args.trace_layers or os.path.join(tempfile.mkdtemp(prefix="nnd_trace_"), "layers.csv")
        conv_ops.trace_start()
        step_resident(0)
        rows = conv_ops.trace_dump(trace_path)
        roof_step = step_roofline(trace_path, ms / args.steps)
        if args.trace_layers:
            print(f"[bench] wrote {rows} convolution launches of one train step to {args.trace_layers}", file=sys.stderr)

    # ---- roofline of the dominant kernel family (gather convolution), measured live with CUDA events
    roof = None
    if rank == 0:
        roof = conv_roofline(net, dev, arch, patch, bs)

    if rank == 0:
        value = world * bs * args.steps / (ms / 1e3)
        e2e_v = world * bs * args.steps / (ms_e2e / 1e3)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "bf16", "data": "synthetic",
               "config": {"workload": f"{args.config}: {patch[0]}x{patch[1]}x{patch[2]} {arch['in_channels']}ch, batch {bs}/GPU, "
                                      f"full train step (fwd+ATSS+HNM+losses+postprocess/3D-NMS+bwd+SGD)",
                          "global_batch": bs * world, "parallelism": f"dp{world}", "experimental": args.experimental,
                          "precision": "bf16 operands (8-bit significand) with fp32 accumulation, fp32 master weights / norm statistics / "
                                       "box engine / losses; the reference trains under fp16 AMP (11-bit significand, "
                                       "nndet/conf/train/v001.yaml:32-33): same operand width, narrower mantissa -- parity gates for the "
                                       "bf16 layers are 1e-2 (tests/test_net_gpu.py), 1e-4 only for the fp32 box engine and losses",
                          "l2": f"{n_batches} distinct input batches; per-step activations (> 4 GB) exceed the 126 MB L2"},
               "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
                       "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "roofline_step": roof_step,
               "roofline_kernels": kernel_rooflines(dev, bs) if world == 1 else None,
               "train_tflops": 3 * conv_flops_per_patch(arch, patch) * bs * world * args.steps / (ms / 1e3) / 1e12}
        try:
            out["nms"] = nms_rates(dev)
        except Exception as e:
            out["nms"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                # free this arm's buffers first: the stock path keeps fp32 + fp16 activations of every layer (tens of GB)
                del trainer, net, resident
                torch.cuda.empty_cache()
                out["gpu_baseline"] = stock_gpu_baseline(dev, args.config)
            except Exception as e:                       # a comparator, never allowed to cost the result line
                out["gpu_baseline"] = {"value": None, "unit": UNIT, "kind": "stock PyTorch/cuDNN modules", "sample": f"failed: {e!r}"}
            try:         # the stock path's best (cudnn.benchmark + channels_last_3d) in a subprocess under a time guard: autotuning ~180 3-D
                # convolution problems can take minutes, and must never cost the result line
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stock-tuned", "--config", args.config], capture_output=True,
                                   text=True, timeout=200)
                out["gpu_baseline"]["tuned"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-300:]}
            except Exception as e:
                out["gpu_baseline"]["tuned"] = {"error": repr(e)[:300]}
            try:
                out["nms_vs_reference_cuda"] = ref_nms_rates(dev)
            except Exception as e:
                out["nms_vs_reference_cuda"] = {"error": repr(e)}
            try:
                out["cpu_baseline"] = cpu_baseline()
            except Exception as e:                       # never lose the GPU line to a host-side problem
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}
        sys.stdout.flush(); os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def conv_roofline(net, dev, arch, patch, bs):
    """Dominant kernel = the gather convolution of the largest layer (encoder stage 0, conv 2: 32->32 at full
    resolution, 464 GFLOP per batch-4 launch = 16.8 % of the forward FLOPs).  achieved = algorithmic FLOPs / mean
    launch time over 5 launches (CUDA events on the launch stream, after 2 warm-ups)."""
    from nndetection_b200.arch import conv_ops as ops
    hbm, tf_burst, tf_sus, kind = peaks()
    layer = net.encoder.stages[0].convs[0][1]
    cin, cout = layer.conv.in_channels, layer.conv.out_channels
    x = torch.randn(bs, cin, *patch, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    plan = layer.plan(bs, tuple(patch))
    wp, _ = layer.packed()
    y = ops.empty_cl(bs, cout, plan.out_sp, device=dev)
    st = torch.zeros((2, bs, cout), dtype=torch.float32, device=dev)
    used = 0
    for _ in range(2):
        used = ops.conv_gather(x, wp, plan.fprop[0], y, cout, cout, stat_sum=st[0], stat_sq=st[1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        ops.conv_gather(x, wp, plan.fprop[0], y, cout, cout, stat_sum=st[0], stat_sq=st[1])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    vox = bs * patch[0] * patch[1] * patch[2]
    flops = 2.0 * 27 * cin * cout * vox
    achieved = flops / (ms / 1e3) / 1e12
    kname = {2: "conv_tcs_kernel (tcgen05, streaming z-window N=96 MMAs)", 1: "conv_tc_kernel (tcgen05 tile kernel)"}.get(used, "conv_igemm_kernel<32> (mma.sync)")
    # DRAM traffic of this launch from the committed `ncu --set full` capture (profiles/r01_ncu_conv_tcs32_summary.txt):
    # dram__bytes_read.sum 545.0 MB + dram__bytes_write.sum 483.8 MB -- equals the algorithmic bytes (no re-reads)
    luna_shape = used == 2 and tuple(patch) == (128, 128, 128) and bs == 4
    traffic = 544.989952e6 + 483.824384e6 if luna_shape else None       # the capture is of THIS shape only; other configs: null
    return {"bound": "tensor", "kernel": kname,
            "layer": f"encoder.stage0.conv2 {cin}->{cout} 3x3x3 @ {patch[0]}x{patch[1]}x{patch[2]} x batch {bs}",
            "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst, "peak_kind": kind + " burst bf16 cuBLAS",
            "ms_per_launch": ms, "algorithmic_flops_per_launch": flops,
            "algorithmic_bytes_per_launch": 2.0 * vox * (cin + cout) + 2.0 * 27 * cin * cout, "traffic": traffic,
            "traffic_source": "profiles/r01_ncu_conv_tcs32_summary.txt (ncu --set full, same layer and shape)" if luna_shape else None}


def kernel_rooflines(dev, bs):
    """The other tensor-core kernels of the step, each on its largest layer of the LUNA plan, launched in isolation (CUDA events on the
    launch stream, 2 warm-ups + 5 timed launches): TFLOP/s against the measured burst bf16 peak.  `roofline` above stays the single
    largest launch; `roofline_step` is the whole step."""
    from nndetection_b200.arch import conv_ops as ops
    from nndetection_b200.arch.conv import ConvInstanceRelu
    hbm, tf_burst, tf_sus, kind = peaks()
    cases = [  # kind, cin, cout, input size, stride
        ("fprop", 128, 128, 32, 1), ("wgrad", 128, 128, 32, 1), ("wgrad", 64, 64, 64, 1), ("wgrad", 32, 32, 128, 1),
        ("fprop", 32, 64, 128, 2), ("wgrad", 32, 64, 128, 2), ("fprop", 64, 64, 64, 1),
    ]
    out = []
    for what, cin, cout, size, stride in cases:
        try:
            layer = ConvInstanceRelu(3, cin, cout, kernel_size=3, stride=stride, padding=1).to(dev)
            x = torch.randn(bs, cin, size, size, size, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
            plan = layer.plan(bs, (size,) * 3)
            wp, _ = layer.packed()
            y = ops.empty_cl(bs, cout, plan.out_sp, device=dev)
            y.normal_()
            st = torch.zeros((2, bs, cout), dtype=torch.float32, device=dev)
            dw = torch.zeros_like(layer.conv.weight)
            if what == "fprop":
                fn = lambda: ops.conv_gather(x, wp, plan.fprop[0], y, cout, cout, stat_sum=st[0], stat_sq=st[1])
            else:
                fn = lambda: ops.conv_wgrad(y, cout, x, cin, plan.wgrad[0], dw, cin * 27, 27, 1, cout, cin)
            ops.trace_start()
            fn()
            import tempfile, csv
            with tempfile.TemporaryDirectory() as td:
                ops.trace_dump(os.path.join(td, "t.csv"))
                kern = [r["kernel"] for r in csv.DictReader(open(os.path.join(td, "t.csv")))][-1]
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            vo = bs * plan.out_sp[0] * plan.out_sp[1] * plan.out_sp[2]
            fl = 2.0 * 27 * cin * cout * vo
            out.append({"kernel": kern, "pass": what, "layer": f"{cin}->{cout} 3x3x3 stride {stride} @ {size}^3 x batch {bs}", "ms_per_launch": ms,
                        "achieved": fl / ms / 1e9, "unit": "TFLOP/s", "peak": tf_burst, "frac": fl / ms / 1e9 / tf_burst,
                        "algorithmic_flops_per_launch": fl})
            del layer, x, y, dw
        except Exception as e:                      # a comparator table: never fail the bench line over it
            out.append({"pass": what, "layer": f"{cin}->{cout} stride {stride} @ {size}^3", "error": repr(e)})
    return out


def _nms_stress(n, dev):
    g = torch.Generator().manual_seed(7)
    c = torch.rand(n, 3, generator=g) * 160
    h = torch.rand(n, 3, generator=g) * 20 + 2
    boxes = torch.stack([c[:, 0] - h[:, 0], c[:, 1] - h[:, 1], c[:, 0] + h[:, 0], c[:, 1] + h[:, 1], c[:, 2] - h[:, 2], c[:, 2] + h[:, 2]], 1).to(dev)
    scores = ((torch.randperm(n, generator=g).float() + 0.5) / n).to(dev)
    return boxes, scores


def _time_nms(fn, boxes, scores, thr, it):
    for _ in range(2):
        keep = fn(boxes, scores, thr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        keep = fn(boxes, scores, thr)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it, keep


def ref_nms_rates(dev):
    """Head-to-head with the REFERENCE's own CUDA NMS (nndet/csrc/cuda/nms.cu:99-221, built by oracle/build_ref_nms.py into
    oracle/_ref/ -- a comparator, like `gpu_baseline`): one call each on the same stress boxes, keep lists compared."""
    from oracle.build_ref_nms import load
    from nndetection_b200 import _C
    ref = load()
    if ref is None:
        return {"unavailable": "oracle/_ref/ref_nms*.so not built"}
    res = {"kind": "reference nndet._C.nms (nms.cu + ops.cpp, two-token dispatch fix), same GPU, same boxes, thr 0.1"}
    for n in (1_000, 10_000, 100_000):
        boxes, scores = _nms_stress(n, dev)
        it = 10 if n <= 10_000 else 3
        ms_ref, keep_ref = _time_nms(ref.nms, boxes, scores, 0.1, it)
        ms, keep = _time_nms(_C.nms, boxes, scores, 0.1, it)
        res[f"n{n}"] = {"ref_ms": ms_ref, "ms": ms, "speedup": ms_ref / ms, "ref_boxes_per_s": n / (ms_ref / 1e3),
                        "boxes_per_s": n / (ms / 1e3), "keep_equal": bool(torch.equal(keep, keep_ref)), "kept": int(keep.numel())}
    return res


FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12          # 148 SMs x 128 FMA lanes x 2 flop x 1.965 GHz (no measured fp32 peak in MEASURED_PEAKS.json)


def step_roofline(trace_csv, ms_step):
    """The WHOLE step against the tensor roof (the single-kernel `roofline` block describes the best kernel,
    not the step).  From the per-launch convolution trace of one extra step (nnd_conv_trace: two CUDA events around every
    convolution-family launch, on its stream): per kernel family the summed launch time, the algorithmic FLOPs (2 * taps * Cin * Cout
    per logical output voxel -- the transposed / parity forms counted once) and TFLOP/s as a fraction of the SUSTAINED measured bf16
    peak (the launches run inside a long step); `other_ms` = step time - convolution launches = norm passes, box engine, losses, NMS,
    optimizer, eager residual adds; `step` = all conv FLOPs / step time.  `tcgen05_flop_share` / `mma_sync_time_share` answer "how much
    of the step still rides on mma.sync"."""
    import csv, collections
    _, _, tf_sus, kind = peaks()
    fam = collections.defaultdict(lambda: [0.0, 0.0, 0])
    for r in csv.DictReader(open(trace_csv)):
        ms = float(r["ms"])
        if ms < 0:
            continue
        f = fam[(r["kind"] if r["kind"] in ("wgrad", "first_fprop", "first_wgrad") else "fprop/dgrad", r["kernel"])]
        f[0] += ms; f[1] += float(r["gflop"]); f[2] += 1
    conv_ms = sum(v[0] for v in fam.values()); gf = sum(v[1] for v in fam.values())
    mma_sync = {"conv_igemm", "wgrad_generic", "wgrad_halo", "conv_first"}
    rows = [{"kind": k[0], "kernel": k[1], "launches": v[2], "ms": round(v[0], 4), "gflop": round(v[1], 1),
             "tflops": round(v[1] / max(v[0], 1e-9), 1), "frac_of_sustained_peak": round(v[1] / max(v[0], 1e-9) / tf_sus, 4),
             "share_of_step": round(v[0] / ms_step, 4), "path": "mma.sync" if k[1] in mma_sync else "tcgen05"}
            for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])]
    tc_gf = sum(v[1] for k, v in fam.items() if k[1] not in mma_sync)
    return {"peak": tf_sus, "peak_kind": kind + " sustained bf16 cuBLAS", "unit": "TFLOP/s", "ms_step": ms_step, "conv_ms": round(conv_ms, 3),
            "other_ms": round(ms_step - conv_ms, 3), "conv_gflop": round(gf, 1),
            "step": {"tflops": round(gf / ms_step, 1), "frac": round(gf / ms_step / tf_sus, 4)},
            "conv_only": {"tflops": round(gf / max(conv_ms, 1e-9), 1), "frac": round(gf / max(conv_ms, 1e-9) / tf_sus, 4)},
            "tcgen05_flop_share": round(tc_gf / max(gf, 1e-9), 4),
            "mma_sync_time_share": round(sum(v[0] for k, v in fam.items() if k[1] in mma_sync) / ms_step, 4),
            "families": rows,
            "note": "launch times from CUDA events inside one traced step (launches on forked streams may overlap: shares can sum past the "
                    "step on those)"}


def nms_rates(dev):
    """BASELINE.json's second metric: 3-D NMS boxes/s = N / time of ONE nndet._C.nms call (sort included), SURVEY 8d stress
    boxes (centres U[0,160)^3, half sizes U[2,22), unique scores), thr 0.1; bytes = 28N + 8N + 8 N ceil(N/64) (upper
    triangle written once, read once) + 8K; pair tests = N (N-1) / 2.  `roofline`: the all-pairs phase is fp32-ALU-bound
    (~22 flop per pair test, 176 flop per mask byte), so the fraction is pair tests x 22 / nominal fp32 peak; the
    HBM fraction of the algorithmic bytes is reported beside it (the north star's 70 % HBM target is the wrong roof for N >= 2 k)."""
    from nndetection_b200 import _C
    res = {}
    for n in (1_000, 10_000, 100_000):
        boxes, scores = _nms_stress(n, dev)
        ms, keep = _time_nms(_C.nms, boxes, scores, 0.1, 10 if n <= 10_000 else 3)
        k = int(keep.numel())
        byt = 36.0 * n + 8.0 * n * ((n + 63) // 64) + 8.0 * k
        pairs = 0.5 * n * (n - 1) / (ms / 1e3)
        res[f"n{n}"] = {"ms": ms, "boxes_per_s": n / (ms / 1e3), "kept": k, "algorithmic_GB_per_s": byt / (ms / 1e3) / 1e9,
                        "pair_tests_per_s": pairs,
                        "roofline": {"bound": "fp32 alu", "achieved": pairs * 22 / 1e12, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                                     "frac": pairs * 22 / 1e12 / FP32_PEAK_TFLOPS, "peak_kind": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz",
                                     "hbm_frac_of_measured": byt / (ms / 1e3) / 1e9 / peaks()[0]}}
    return res


def stock_gpu_baseline(dev, config="luna", steps=3, warmup=2, tuned=False):
    """SURVEY 8d's second comparator, labelled separately from the CPU baseline: the reference's network as STOCK PyTorch modules
    (nn.Conv3d / ConvTranspose3d / InstanceNorm3d / GroupNorm through cuDNN: the oracle's modules are the reference's operators) on
    this GPU, the way the reference trains it (fp16 autocast + GradScaler = Lightning `precision: 16`, torch.optim.SGD nesterov,
    cudnn.benchmark off = nndet/conf/train/v001.yaml:39, which also keeps this leg free of minutes of cuDNN autotuning).  Network forward + backward + optimizer step ONLY, same batch shape -- no anchors, ATSS, sampling, box losses,
    post-processing or NMS (a surrogate loss on the three outputs drives the backward), so the number FAVOURS the stock path."""
    from oracle import model_oracle as mo
    arch, anc, patch, bs = mo.make_plan(config)
    cuda = torch.device(dev).type == "cuda"
    old_bench = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = bool(tuned)       # tuned: the stock path's best -- cuDNN autotuning + channels_last_3d
    try:
        torch.manual_seed(4321)
        net = mo.RetinaUNetOracle(dict(arch), dict(anc)).to(dev)
        if tuned:
            net = net.to(memory_format=torch.channels_last_3d)
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=3e-5)
        scaler = torch.amp.GradScaler("cuda", enabled=cuda)
        images = torch.rand(bs, arch["in_channels"], *patch, device=dev)
        if tuned:
            images = images.contiguous(memory_format=torch.channels_last_3d)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast(device_type="cuda" if cuda else "cpu", dtype=torch.float16 if cuda else torch.bfloat16, enabled=cuda):
                fm_all = net.decoder(net.encoder(images))
                pred = net.head([fm_all[i] for i in net.decoder_levels])
                seg = net.segmenter(fm_all)
                loss = (pred["box_logits"].float().square().mean() + pred["box_deltas"].float().square().mean()
                        + seg["seg_logits"].float().square().mean())
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()

        for _ in range(warmup):
            step()
        if cuda:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        if cuda:
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
        else:
            ms = 1e3 * (time.perf_counter() - t0) / steps
        peak_gb = torch.cuda.max_memory_allocated(dev) / 2**30 if cuda else None
        return {"value": bs / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "kind": ("stock PyTorch/cuDNN modules, fp16 autocast + GradScaler, channels_last_3d, cudnn.benchmark ON (the stock path's best)" if tuned
                         else "stock PyTorch/cuDNN modules, fp16 autocast + GradScaler, NCDHW, cudnn.benchmark off (reference default)"),
                "sample": f"{steps} steps of batch {bs} {patch[0]}x{patch[1]}x{patch[2]}: network fwd + bwd + SGD only (no box engine / NMS; surrogate loss)",
                "torch": torch.__version__, "cudnn": torch.backends.cudnn.version() if cuda else None, "max_memory_GiB": peak_gb}
    finally:
        torch.backends.cudnn.benchmark = old_bench
        if cuda:
            torch.cuda.empty_cache()


def cpu_baseline():
    """Oracle port (the reference's torch-CPU operators) on the host cores: ONE train step of ONE 128^3 patch."""
    from oracle import model_oracle as mo
    torch.set_num_threads(min(os.cpu_count(), 64))
    arch, anc, patch, bs = mo.make_plan("luna")
    net = mo.RetinaUNetOracle(dict(arch), dict(anc))
    images, targets = mo.synth_batch(patch, 1, arch["in_channels"], arch["classifier_classes"], 1234)
    with torch.no_grad():
        net(images[:, :, :64, :64, :64])       # touch the operators once (oneDNN primitive creation); 64^3 -> 2^3 at the bottleneck
    t0 = time.perf_counter()
    losses, aux = net.train_step(images, targets, seed=1)
    net.postprocess(images, {k: v.detach() for k, v in aux["pred"].items()}, aux["anchors"])
    sum(losses.values()).backward()
    dt = time.perf_counter() - t0
    out = {"value": 1.0 / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
           "sample": "1 train step of 1 patch (fwd+loss+postprocess+bwd), torch CPU fp32"}
    try:
        out["nms_cpu"] = cpu_nms_rates()
    except Exception as e:                              # the patches/s baseline above must survive
        out["nms_cpu"] = {"error": repr(e)}
    return out


def cpu_nms_rates():
    """The reference's CPU NMS path (nms_cpu, nndet/core/boxes/nms.py:31-53, restated in oracle.box_oracle.nms_greedy) on the
    same stress boxes as the GPU `nms` block, N = 1 k and 10 k (N = 100 k would need a 40 GB IoU matrix in the reference)."""
    from oracle import box_oracle as bo
    res = {}
    for n in (1_000, 10_000):
        g = torch.Generator().manual_seed(7)
        c = torch.rand(n, 3, generator=g) * 160
        h = torch.rand(n, 3, generator=g) * 20 + 2
        boxes = torch.stack([c[:, 0] - h[:, 0], c[:, 1] - h[:, 1], c[:, 0] + h[:, 0], c[:, 1] + h[:, 1], c[:, 2] - h[:, 2], c[:, 2] + h[:, 2]], 1)
        scores = (torch.randperm(n, generator=g).float() + 0.5) / n
        t0 = time.perf_counter()
        keep = bo.nms_greedy(boxes, scores, 0.1)
        dt = time.perf_counter() - t0
        res[f"n{n}"] = {"ms": 1e3 * dt, "boxes_per_s": n / dt, "kept": int(keep.numel())}
    res["note"] = ("oracle port: row-wise numpy restatement, 1 thread; the reference's nms_cpu materialises the N x N IoU matrix "
                   "(SURVEY 8d: 2.09 s at N = 10 k on the survey host)")
    return res


if __name__ == "__main__":
    main()
