#!/usr/bin/env python
"""Benchmark of the S-sample multi-exit inference path (BASELINE.json metric: images/sec at S MC samples,
multi-exit ResNet-18, 32x32, and % of tensor-core peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--dtype fp16|bf16|fp32]
    python bench.py --impl reference ...        # the reference's algorithm on the host cores (oracle port)

A "step" is one pass of the hot path over one batch: B images x S stochastic samples -> per-exit predictive
mean / entropy statistics.  N = 1 runs BASELINE.json configs[1] (ResNet-18 ME, MC dropout after every stage +
at every exit, S = 32, B = 256, 3x32x32).  N > 1 (torchrun, one rank per GPU): every rank owns its own batch of
B images (weak scaling, global batch N*B) and the per-exit statistics are all-gathered over NCCL inside the
timed region; `--workload c5` instead shards the S = 128 samples of ONE batch over the ranks (strong scaling)
with a single all-reduce of the per-exit sums.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec at S MC samples (multi-exit ResNet-18, 32x32) & % tensor-core peak"

WORKLOADS = {
    # name: (description, model factory name, B, S, classes)
    "c1": ("multi-exit LeNet-5 MC-dropout (last-layer, p=0.2), S=8, 1x28x28, batch 64 (the reference's CPU-runnable "
           "case; CUDA-core kernels only: no tensor-core claim)", "lenet", 64, 8, 10),
    "c2": ("multi-exit ResNet-18 MC-dropout (block+exit, p=0.5), S=32, 3x32x32, batch 256", "resnet_mcd", 256, 32, 10),
    "c3": ("multi-exit ResNet-18 Masksembles (4 masks, scale 2.0), S=4, 3x32x32, batch 256, 100 classes",
           "resnet_mask", 256, 4, 100),
    "c4": ("multi-exit VGG-19 MC-dropout on the last 3 blocks + exits, S=64, 3x32x32, batch 512, 100 classes",
           "vgg_last3", 512, 64, 100),
    "c5": ("multi-exit ResNet-18 MC-dropout S=128 sample-sharded over the ranks, batch 256", "resnet_mcd", 256, 128, 10),
}


def input_shape(kind):
    return (1, 28, 28) if kind == "lenet" else (3, 32, 32)


def build_model(kind, classes):
    import numpy as np
    import torch
    from bayesnn_fpga_b200 import lenet, resnet18, vgg19
    torch.manual_seed(0)
    np.random.seed(0)
    if kind == "lenet":
        m = lenet.LeNetMCEarlyExit(dropout_p=0.2, out_dim=classes)
    elif kind == "resnet_mcd":
        m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", dropout_p=0.5, out_dim=classes)
    elif kind == "resnet_mask":
        m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", out_dim=classes, mask_type="mask",
                                         num_masks=4, mask_scale=2.0)
    elif kind == "vgg_last3":
        m = vgg19.VGG19MCEarlyExit(dropout_exit=True, dropout=None, dropout_p=0.5, out_dim=classes, image_size=32,
                                   n_exits=5).append_block_dropout((2, 3, 4))
    else:
        raise ValueError(kind)
    # random-init weights of the reference architecture + non-trivial BatchNorm statistics
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1, generator=g)
            mod.running_var.uniform_(0.5, 1.5, generator=g)
            mod.weight.data.uniform_(0.5, 1.5, generator=g)
            mod.bias.data.normal_(0, 0.1, generator=g)
    return m.eval()


def oracle_forward_fn(kind):
    from oracle import nets
    if kind == "lenet":
        return lambda sd, x, site: nets.lenet_forward(sd, x, site)
    if kind in ("resnet_mcd", "resnet_mask"):
        return lambda sd, x, site: nets.resnet18_forward(sd, x, site, "block", True)
    return lambda sd, x, site: nets.vgg19_forward(sd, x, site, True, (2, 3, 4))


def time_cpu_reference(kind, classes, S_nominal, budget_images=256, passes=4):
    """The reference's algorithm (results_analyzer.py:236-270: S sequential full forward passes, softmax per exit
    per pass, fp64 host means) through the oracle port, torch's own dropout RNG, all host threads.
    Bounded sample: `budget_images` images x `passes` passes, scaled linearly to S_nominal passes."""
    import torch
    from oracle import nets, stats
    torch.set_num_threads(os.cpu_count() or 1)
    model = build_model(kind, classes)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tables = {k[:-len(".masks")]: v for k, v in sd.items() if k.endswith(".masks")}
    spec = nets.SiteSpec("mask" if kind == "resnet_mask" else "mc", 0.2 if kind == "lenet" else 0.5, tables)
    x = torch.randn(budget_images, *input_shape(kind))
    fwd = oracle_forward_fn(kind)

    def run():
        sites = nets.TorchRngSites(spec)

        def one(i):
            out = fwd(sd, x, sites)
            sites.next_pass()
            return out
        with torch.no_grad():
            stats.mc_get_output(one, passes)
    run()                                   # warm-up
    reps = []
    for _ in range(3):
        t = time.perf_counter()
        run()
        reps.append(time.perf_counter() - t)
    t_med = sorted(reps)[1]
    img_per_s = budget_images / (t_med * S_nominal / passes)
    return {"value": img_per_s, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "%d images x %d of %d passes (full recompute per pass, like the reference), median of 3, "
                      "scaled linearly to S=%d; %.2f s per repetition" % (budget_images, passes, S_nominal, S_nominal,
                                                                         t_med)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx, self.proc, self.lines = device_index, None, []
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        window = "timed region"
        if not inside:                       # region shorter than one sampling period: nearest samples
            inside, window = [ln for (_, ln) in self.lines[-3:]], "nearest samples (timed region < sampling period)"
        for ln in inside:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "sm_mhz_min": sm[0], "power_w_max": max(pw),
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


NCU_SUMMARY = "profiles/r01_ncu_conv_tc_full_v3.json"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per conv_tc launch (mean over the 20 launches of one C2 step) from
    the committed `ncu --set full` summary (tools/ncu_summary.py), or None."""
    p = os.path.join(ROOT, NCU_SUMMARY)
    if not os.path.exists(p):
        return None
    rows = json.load(open(p))
    return sum((r["dram_read_MB"] + r["dram_write_MB"]) * 1e6 for r in rows) / max(len(rows), 1)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"burst": d["bf16_tflops"], "sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "src": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


def reference_arm(args, wl):
    desc, kind, B, S, classes = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm, reps = max(args.warmup, 1), max(args.steps, 1)
    # bounded so that the whole run ends within a few minutes: each step = 64 images x 2 passes, scaled to S
    import torch
    from oracle import nets, stats
    torch.set_num_threads(os.cpu_count() or 1)
    model = build_model(kind, classes)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tables = {k[:-len(".masks")]: v for k, v in sd.items() if k.endswith(".masks")}
    spec = nets.SiteSpec("mask" if kind == "resnet_mask" else "mc", 0.2 if kind == "lenet" else 0.5, tables)
    nb, passes = 64, 2
    x = torch.randn(nb, *input_shape(kind))
    fwd = oracle_forward_fn(kind)

    def step():
        sites = nets.TorchRngSites(spec)
        with torch.no_grad():
            stats.mc_get_output(lambda i: fwd(sd, x, sites), passes)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    val = nb / (dt * S / passes)
    sample = "each step = %d images x %d of %d passes on %d host threads, scaled linearly to S=%d" % (
        nb, passes, S, os.cpu_count() or 1, S)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "batch": B, "S": S, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch (debugging)")
    ap.add_argument("--samples", type=int, default=0, help="override the workload's S (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-ops", action="store_true", help="print the per-op device times of one step")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return reference_arm(args, wl)

    import torch
    import torch.distributed as dist
    from bayesnn_fpga_b200 import mc_predict, predict

    desc, kind, B, S, classes = wl
    B = args.batch or B
    S = args.samples or S
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shard_samples = args.workload == "c5" and world > 1
    W = max(args.warmup, 3)

    model = build_model(kind, classes).to(dev)
    eng = model.bnn_engine(args.dtype)
    g = torch.Generator().manual_seed(100 + (0 if shard_samples else rank))
    x_host = torch.randn(B, *input_shape(kind), generator=g).pin_memory()
    x_dev = x_host.to(dev)
    E = eng.graph.n_exits
    gather_buf = [torch.empty(4 * E * B * classes + 3 * E * B, device=dev) for _ in range(world)] if world > 1 else None

    def step_device():
        """inputs already resident in HBM"""
        if shard_samples:
            s0, sl = predict.shard_samples(S, world, rank)
            r = eng.run(x_dev, sl, sample0=s0, S_total=S, reduce_fn=predict.allreduce_sums)
        else:
            r = eng.run(x_dev, S)
            if world > 1:      # weak scaling: every rank's statistics are gathered (one small NCCL collective)
                dist.all_gather(gather_buf, eng._bufs[(B, S, False)]["out"])
        return r

    out_host = torch.empty((4, E, B, classes), dtype=torch.float32).pin_memory()

    def step_e2e():
        """the public API with HOST buffers: pinned H2D of the batch, D2H of the statistics, every step"""
        r = mc_predict(model, x_host, S, dtype=args.dtype, distributed=shard_samples)
        out_host[0].copy_(r.mean_probs, non_blocking=True)
        out_host[1].copy_(r.mean_logits, non_blocking=True)
        out_host[2].copy_(r.ens_probs, non_blocking=True)
        out_host[3].copy_(r.ens_logits, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # nvidia-smi needs ~1 s to start: launch it before the warm-up
    for _ in range(W):
        step_device()
    launches0 = eng.launches
    sampler.mark_start()
    ms = timed(step_device, args.steps)
    sampler.mark_end()
    launches = eng.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    images_per_step = B * (1 if shard_samples else world)
    value = images_per_step * args.steps / (ms * 1e-3)
    e2e_value = images_per_step * args.steps / (ms_e2e * 1e-3)

    # ---- per-op device times of one step (CUDA events on the launching stream) -> roofline of the dominant kernel
    prof = eng.profile_step(x_dev, S if not shard_samples else predict.shard_samples(S, world, rank)[1])
    peaks = load_peaks()
    tc = [o for o in prof if o["kernel"] == "conv_tc"]
    pre_macs, suf_macs = eng.graph.macs()
    if tc:
        flops = sum(o["flops"] for o in tc)
        t_tc = sum(o["ms"] for o in tc) * 1e-3
        ach = flops / t_tc / 1e12
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit GEMM)", "achieved": ach,
                "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": ach / peaks["sustained"],
                "peak_source": "%s bf16 cuBLAS sustained (kernel timed inside a long step); burst %.1f" % (
                    peaks["src"], peaks["burst"]),
                "launches_per_step": len(tc), "avg_launch_ms": t_tc * 1e3 / len(tc),
                "flops_per_launch": flops / len(tc), "share_of_step": t_tc * 1e3 / sum(o["ms"] for o in prof),
                "algorithmic_bytes_per_launch": sum(o["bytes"] for o in tc) / len(tc),
                "traffic": ncu_traffic() if args.workload == "c2" else None,
                "traffic_source": NCU_SUMMARY + " (ncu --set full --clock-control none over the 20 conv launches of one "
                                  "C2 step: mean dram__bytes_read.sum + dram__bytes_write.sum per launch)"}
        if roof["frac"] > 1.0:
            roof["note"] = ("above the SUSTAINED cuBLAS figure (dense random operands at the 1 kW cap): this workload's "
                            "small maps make most im2col taps zero padding, which draws less power, so the SM clock "
                            "stays higher; against the burst figure the fraction is %.3f" % (ach / peaks["burst"]))
    else:
        flops = sum(o["flops"] for o in prof)
        t_all = sum(o["ms"] for o in prof) * 1e-3
        roof = {"bound": "tensor", "kernel": "conv2d_simt (CUDA cores, fp32 parity path)", "achieved": flops / t_all / 1e12,
                "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": flops / t_all / 1e12 / peaks["sustained"],
                "peak_source": peaks["src"], "traffic": None}
    step_flops = 2.0 * B * (pre_macs + (S if not shard_samples else S / world) * suf_macs)
    kernels = {}
    for o in prof:
        k = kernels.setdefault(o["kernel"], {"ms": 0.0, "launches": 0, "bytes": 0, "flops": 0})
        k["ms"] += o["ms"]; k["launches"] += 1; k["bytes"] += o["bytes"]; k["flops"] += o["flops"]
    for k in kernels.values():
        k["GBps"] = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
        k["TFLOPs"] = k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else None
    if args.profile_ops and rank == 0:
        for o in prof:
            print("  %-10s %-28s %8.3f ms  %8.1f TFLOP/s %8.1f GB/s" % (
                o["kernel"], o["name"], o["ms"], o["flops"] / (o["ms"] * 1e-3) / 1e12 if o["ms"] else 0,
                o["bytes"] / (o["ms"] * 1e-3) / 1e9 if o["ms"] else 0), file=sys.stderr)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = time_cpu_reference(kind, classes, S)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if shard_samples else "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.dtype] + " operands, f32 accumulate",
            "data": "synthetic",
            "config": {"workload": desc if not (args.batch or args.samples) else desc + " [overridden B=%d S=%d]" % (B, S),
                       "batch_per_gpu": B, "global_batch": images_per_step, "S": S, "exits": E, "classes": classes,
                       "partition": "samples" if shard_samples else ("batch" if world > 1 else "none"),
                       "l2": "per-step activation working set (GiB) >> 126 MB L2; no explicit flush",
                       "cuda_graph": os.environ.get("BNN_CUDA_GRAPH", "1") != "0",
                       "masksembles_mode": eng.gather_mode if kind == "resnet_mask" else None,
                       "algorithmic_gflop_per_image": 2e-9 * (pre_macs + S * suf_macs)},
            "tflops_whole_step": step_flops * (world if not shard_samples else world) / (ms / args.steps * 1e-3) / 1e12,
            "frac_of_tensor_peak_whole_step": step_flops / (ms / args.steps * 1e-3) / 1e12 / peaks["sustained"],
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": launches,
            "roofline": roof,
            "kernels": kernels,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
