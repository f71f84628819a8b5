#!/usr/bin/env python
"""Benchmark of the S-sample multi-exit inference path (BASELINE.json metric: images/sec at S MC samples,
multi-exit ResNet-18, 32x32, and % of tensor-core peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--dtype fp16|bf16|fp32]
    python bench.py --impl reference ...        # the reference's own CPU implementation (oracle/_ref) on the host cores

A "step" is one pass of the hot path over one batch: B images x S stochastic samples -> per-exit predictive
mean / entropy statistics.  The HEADLINE (`value`, `e2e`, `roofline`) is BASELINE.json configs[1] (C2: ResNet-18 ME,
MC dropout after every stage + at every exit, S = 32, B = 256, 3x32x32).  The same JSON line carries, each measured in
its own short timed region with its own clock sample:
  * N = 1:  `configs` = {c1, c3, c4, c5} (LeNet S=8; Masksembles S=4 C=100; VGG-19 S=64 B=512; ResNet S=128 on one GPU)
  * N > 1:  the headline is C2 weak-scaled (every rank its own B images: independent batches, NO data-path collective;
            BNN_BENCH_WEAK_GATHER=1 adds an `all_gather_into_tensor` of the statistics to every step) and `c5_strong` = BASELINE configs[4]: the S = 128 samples of ONE batch
            sharded over the ranks with a single all-reduce of the per-exit sums (strong scaling), with the unsharded
            single-GPU time measured in the same process for `efficiency_vs_n1`.
`--workload cX` makes cX the headline instead (debugging / profiling).
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


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref, staged by oracle/make_ref.py) - or, where no PyTorch
# reference exists (LeNet) or the staged tree is missing, the oracle port
# ---------------------------------------------------------------------------------------------------------------
def oracle_forward_fn(kind):
    from oracle import nets
    if kind == "lenet":
        return lambda sd, x, site: nets.lenet_forward(sd, x, site)
    if kind in ("resnet_mcd", "resnet_mask"):
        return lambda sd, x, site: nets.resnet18_forward(sd, x, site, "block", True)
    return lambda sd, x, site: nets.vgg19_forward(sd, x, site, True, (2, 3, 4))


def cpu_step_fn(kind, classes, S):
    """-> (step(x) running the FULL S passes of the reference's `_get_output` on the host for the images in x, kind)
    kind "reference": the unmodified reference modules from oracle/_ref; "port": the oracle restatement."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import ref_arm
    if kind != "lenet" and ref_arm.available():
        runner = ref_arm.ReferenceRunner(kind, classes, S)
        return runner.step, "reference"
    from oracle import nets, stats
    model = build_model(kind, classes)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tables = {k[:-len(".masks")]: v for k, v in sd.items() if k.endswith(".masks")}
    spec = nets.SiteSpec("mask" if kind == "resnet_mask" else "mc", 0.2 if kind == "lenet" else 0.5, tables)
    fwd = oracle_forward_fn(kind)

    def step(x):
        sites = nets.TorchRngSites(spec)

        def one(i):
            out = fwd(sd, x, sites)
            sites.next_pass()
            return out
        with torch.no_grad():
            return stats.mc_get_output(one, S)
    return step, "port"


def calibrate_images(step, kind, B, budget_s):
    """How many images of the batch one CPU step may take to stay inside `budget_s` (full S passes each; images are
    independent, so images/s does not depend on the slice).  -> (n_images, seconds of the calibration step)"""
    import torch
    n0 = min(B, 8)
    x = torch.randn(n0, *input_shape(kind))
    step(x)                                         # first call: allocator / thread-pool warm-up
    t = time.perf_counter()
    step(x)
    dt = time.perf_counter() - t
    n = int(budget_s / max(dt / n0, 1e-9))
    n = max(8, min(B, n - n % 8 if n >= 8 else 8))
    return n, dt


def time_cpu_reference(kind, classes, S, B, budget_s=6.0):
    """cpu_baseline: median of 3 steps of the reference's S-pass loop over a bounded slice of the batch."""
    import torch
    step, impl = cpu_step_fn(kind, classes, S)
    n, _ = calibrate_images(step, kind, B, budget_s)
    x = torch.randn(n, *input_shape(kind))
    reps = []
    for _ in range(3):
        t = time.perf_counter()
        step(x)
        reps.append(time.perf_counter() - t)
    t_med = sorted(reps)[1]
    return {"value": n / t_med, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": impl,
            "sample": "%d of the batch's %d images x ALL %d passes (FullAnalysis._get_output of %s, full recompute per "
                      "pass, torch's own dropout RNG, fp64 host means), median of 3 steps of %.2f s" % (
                          n, B, S, "the unmodified reference modules in oracle/_ref" if impl == "reference"
                          else "the oracle port", t_med)}


def reference_arm(args, wl_name):
    desc, kind, B, S, classes = WORKLOADS[wl_name]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    warm, reps = max(args.warmup, 1), max(args.steps, 1)
    step, impl = cpu_step_fn(kind, classes, S)
    # bounded so that the whole run ends within a few minutes: every step = n images of the batch x ALL S passes
    n, _ = calibrate_images(step, kind, B, 150.0 / (warm + reps))
    x = torch.randn(n, *input_shape(kind))
    for _ in range(warm):
        step(x)
    t0 = time.perf_counter()
    for _ in range(reps):
        step(x)
    dt = (time.perf_counter() - t0) / reps
    val = n / dt
    cores = os.cpu_count() or 1
    sample = "each step = %d of the batch's %d images x ALL %d passes through %s on %d host threads" % (
        n, B, S, "the unmodified reference (oracle/_ref: models.model_loader.get_network + FullAnalysis._get_output "
        "source)" if impl == "reference" else "the oracle port (no PyTorch reference exists for this network)", cores)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "batch": B, "S": S, "images_per_step": n, "same_S": True, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": impl, "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe): one sampler process,
    any number of [t0, t1] windows read from it afterwards."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx, self.proc, self.lines = device_index, None, []

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
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1]
        where = "timed region"
        if not inside:                       # region shorter than one sampling period: nearest samples
            near = sorted(self.lines, key=lambda e: abs(e[0] - 0.5 * (t0 + t1)))[:3]
            inside, where = [ln for (_, ln) in near], "nearest samples (timed region < sampling period)"
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
                "samples": len(sm), "window": where, "seconds": t1 - t0, "reasons": sorted(reasons)}


NCU_SUMMARY = "profiles/r02_ncu_conv_tc_full.json"
NCU_SUMMARY_FALLBACK = "profiles/r01_ncu_conv_tc_full_v3.json"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per conv_tc launch (mean over the conv launches of one C2 step) from
    the committed `ncu --set full` summary (tools/ncu_summary.py), or (None, None)."""
    for rel in (NCU_SUMMARY, NCU_SUMMARY_FALLBACK):
        p = os.path.join(ROOT, rel)
        if os.path.exists(p):
            rows = json.load(open(p))
            return sum((r["dram_read_MB"] + r["dram_write_MB"]) * 1e6 for r in rows) / max(len(rows), 1), rel
    return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"burst": d["bf16_tflops"], "sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "src": "MEASURED_PEAKS.json"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "src": "B200_PROFILING.md fallback"}


def summarize_profile(prof, peaks, clocks, region_s, with_traffic):
    """Per-launch CUDA-event times of one step -> (roofline of the dominant tensor kernel, per-kernel table with HBM
    sub-rooflines).  The denominator is chosen from how the TIMED REGION ran: a region of a few tenths of a second at
    (near) maximum SM clock is the burst regime -> burst peak; seconds-long, power-capped regions -> sustained peak.
    Both fractions are always printed."""
    kernels = {}
    for o in prof:
        k = kernels.setdefault(o["kernel"], {"ms": 0.0, "launches": 0, "bytes": 0, "flops": 0, "flops_exec": 0})
        k["ms"] += o["ms"]; k["launches"] += 1; k["bytes"] += o["bytes"]; k["flops"] += o["flops"]
        k["flops_exec"] += o["flops_exec"]
    total_ms = sum(o["ms"] for o in prof)
    for name, k in kernels.items():
        k["share_of_step"] = k["ms"] / total_ms if total_ms else None
        k["GBps"] = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
        k["TFLOPs_exec"] = k["flops_exec"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else None
        if name in ("dropout", "exit_head", "maxpool", "layout") and k["GBps"] is not None:
            # HBM-bound kernels: algorithmic bytes (read once + write once) / CUDA-event time vs the measured copy peak
            k["roofline"] = {"bound": "hbm", "achieved": k["GBps"], "peak": peaks["hbm"], "unit": "GB/s",
                             "frac": k["GBps"] / peaks["hbm"]}
    sm = (clocks or {}).get("sm_mhz")
    mx = (clocks or {}).get("sm_max_mhz")
    burst_regime = region_s < 2.0 and (sm is None or mx is None or sm >= 0.9 * mx)
    kind = "burst" if burst_regime else "sustained"
    tc = [o for o in prof if o["kernel"] == "conv_tc"]
    if tc:
        fx = sum(o["flops_exec"] for o in tc)
        fa = sum(o["flops"] for o in tc)
        t_tc = sum(o["ms"] for o in tc) * 1e-3
        ach = fx / t_tc / 1e12
        traffic, src = ncu_traffic() if with_traffic else (None, None)
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit GEMM)", "achieved": ach,
                "peak": peaks[kind], "unit": "TFLOP/s", "frac": ach / peaks[kind], "peak_kind": kind,
                "frac_of_burst_peak": ach / peaks["burst"], "frac_of_sustained_peak": ach / peaks["sustained"],
                "peak_source": "%s: bf16 cuBLAS %s %.1f TFLOP/s chosen because the timed region ran %.2f s at a median "
                               "%s MHz of %s MHz (burst %.1f / sustained %.1f)" % (
                                   peaks["src"], kind, peaks[kind], region_s, sm, mx, peaks["burst"],
                                   peaks["sustained"]),
                "flops_counted": "EXECUTED multiply-adds (3x3 windows on 1x1 maps run their centre tap only, on 2x2 maps the "
                                 "16 of 36 tap blocks that do not multiply zero padding, position-major tiles on 4x4 / 8x8 "
                                 "maps skip the padding-only taps: 25 of 36 / 121 of 144; zero-padded stem channels and "
                                 "Masksembles-dropped channels are not counted); achieved_algorithmic counts the "
                                 "reference layer's full multiply-adds over the same time",
                "achieved_algorithmic": fa / t_tc / 1e12,
                "launches_per_step": len(tc), "avg_launch_ms": t_tc * 1e3 / len(tc),
                "flops_per_launch": fx / len(tc), "share_of_step": t_tc * 1e3 / total_ms,
                "algorithmic_bytes_per_launch": sum(o["bytes"] for o in tc) / len(tc),
                "traffic": traffic,
                "traffic_source": (src + " (ncu --set full --clock-control none over the conv launches of one C2 step: "
                                   "mean dram__bytes_read.sum + dram__bytes_write.sum per launch)") if src else None}
    else:
        flops = sum(o["flops_exec"] for o in prof)
        t_all = total_ms * 1e-3
        roof = {"bound": "tensor", "kernel": "conv2d_simt (CUDA cores: fp32 parity path / LeNet)",
                "achieved": flops / t_all / 1e12, "peak": peaks[kind], "unit": "TFLOP/s",
                "frac": flops / t_all / 1e12 / peaks[kind], "peak_kind": kind, "peak_source": peaks["src"], "traffic": None,
                "note": "CUDA-core kernels: no tensor-core claim for this workload"}
    return roof, kernels


WEAK_GATHER = os.environ.get("BNN_BENCH_WEAK_GATHER") == "1"
_RESULT_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on fd 1 when the
    box sets NCCL_DEBUG), so the real stdout is kept aside for the result line and fd 1 is pointed at stderr for
    everything else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _RESULT_OUT


def emit(line):
    out = claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
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
    ap.add_argument("--no-extra-configs", action="store_true", help="headline only (profiling runs)")
    ap.add_argument("--profile-ops", action="store_true", help="print the per-op device times of one step")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args, args.workload)

    import torch
    import torch.distributed as dist
    from bayesnn_fpga_b200 import mc_predict, predict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    peaks = load_peaks()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # nvidia-smi needs ~1 s to start: launch it before the first warm-up

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """-> (ms max over ranks, per-rank ms list, wall-clock window).  CUDA events on the launching stream,
        barrier + synchronize on both sides."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        per_rank = [ms]
        if world > 1:
            t = torch.tensor([ms], device=dev)
            allt = torch.empty(world, device=dev)
            dist.all_gather_into_tensor(allt, t)
            per_rank = [float(v) for v in allt.tolist()]
            ms = max(per_rank)
        return ms, per_rank, (t0, t1)

    def measure(wl_name, steps, headline, shard, B_over=0, S_over=0):
        """One workload on this rank set.  shard: None (N=1) | "batch" (weak: every rank its own batch, statistics
        all-gathered) | "samples" (strong: the S samples of one batch over the ranks, sums all-reduced)."""
        desc, kind, B, S, classes = WORKLOADS[wl_name]
        B, S = B_over or B, S_over or S
        model = build_model(kind, classes).to(dev)
        eng = model.bnn_engine(args.dtype)
        g = torch.Generator().manual_seed(100 + (rank if shard == "batch" else 0))
        x_host = torch.randn(B, *input_shape(kind), generator=g).pin_memory()
        x_dev = x_host.to(dev)
        E = eng.graph.n_exits
        n_stats = 4 * E * B * classes + 3 * E * B
        gathered = torch.empty(world * n_stats, device=dev) if shard == "batch" else None
        s0, sl = predict.shard_samples(S, world, rank) if shard == "samples" else (0, S)

        def gather(out):
            dist.all_gather_into_tensor(gathered, out)

        def step_device():
            """inputs already resident in HBM"""
            if shard == "samples":
                return eng.run(x_dev, sl, sample0=s0, S_total=S, reduce_fn=predict.allreduce_sums)
            # batch sharding: the ranks' batches are independent - nothing is exchanged (optionally ONE NCCL all-gather
            # right behind the finaliser, captured inside the step's CUDA graph)
            return eng.run(x_dev, S, gather_fn=gather if (shard == "batch" and WEAK_GATHER) else None)

        out_host = torch.empty((4 * E * B * classes + 3 * E * B,), dtype=torch.float32).pin_memory()

        def step_e2e():
            """the public API with HOST buffers: pinned H2D of the batch, D2H of the statistics, every step"""
            r = mc_predict(model, x_host, S, dtype=args.dtype, distributed=(shard == "samples"))
            out_host.copy_(r.flat, non_blocking=True)          # all statistics (means, ensembles, entropies): one D2H copy
            torch.cuda.current_stream().synchronize()
            return r

        for _ in range(W):
            step_device()
        launches0 = eng.launches
        ms, per_rank, win = timed(step_device, steps)
        launches = eng.launches - launches0
        clocks = sampler.window(*win) if rank == 0 else None
        for _ in range(2):
            step_e2e()
        ms_e2e, _, _ = timed(step_e2e, steps)
        images = B * (world if shard == "batch" else 1)
        rec = {"workload": desc, "value": images * steps / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms / steps,
               "steps": steps, "batch_per_gpu": B, "global_batch": images, "S": S, "exits": E, "classes": classes,
               "partition": shard or "none", "per_rank_ms_per_step": [v / steps for v in per_rank],
               "e2e": {"value": images * steps / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_step": ms_e2e / steps,
                       "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4},
               "gpu_launches": launches, "clocks": clocks,
               "collective_in_graph": (bool(eng._graph_collectives) if (shard == "samples" or (shard and WEAK_GATHER))
                                       else None)}
        prof = eng.profile_step(x_dev, sl)
        roof, kernels = summarize_profile(prof, peaks, clocks, ms * 1e-3, with_traffic=(wl_name == "c2"))
        pre_macs, suf_macs = eng.graph.macs()
        rec["algorithmic_gflop_per_image"] = 2e-9 * (pre_macs + S * suf_macs)
        step_flops = 2.0 * B * (pre_macs + sl * suf_macs) * world        # all ranks (each replicates the prefix)
        rec["tflops_whole_step"] = step_flops / (ms / steps * 1e-3) / 1e12
        rec["frac_of_tensor_peak_whole_step"] = {
            "burst": rec["tflops_whole_step"] / world / peaks["burst"],
            "sustained": rec["tflops_whole_step"] / world / peaks["sustained"]}
        rec["roofline"] = roof
        rec["kernels"] = kernels
        if kind == "resnet_mask":
            rec["masksembles_mode"] = eng.gather_mode
        if args.profile_ops and rank == 0 and headline:
            for o in prof:
                print("  %-10s %-44s %8.3f ms  %8.1f TFLOP/s %8.1f GB/s" % (
                    o["kernel"], o["name"], o["ms"], o["flops_exec"] / (o["ms"] * 1e-3) / 1e12 if o["ms"] else 0,
                    o["bytes"] / (o["ms"] * 1e-3) / 1e9 if o["ms"] else 0), file=sys.stderr)
        extra = None
        if shard == "samples":
            # the unsharded run on ONE GPU in the same process (every rank does it, no collective) and the collective
            # alone: strong-scaling efficiency without a second launch of the benchmark
            n1 = max(3, min(steps, 5))
            for _ in range(2):
                eng.run(x_dev, S)
            ms1, _, _ = timed(lambda: eng.run(x_dev, S), n1)
            flat = eng._bufs[(B, sl, False)]["sums"]
            for _ in range(5):
                predict.allreduce_sums(flat.clone())
            buf = flat.clone()
            ms_c, _, _ = timed(lambda: predict.allreduce_sums(buf), 50)
            for _ in range(5):
                predict.allreduce_sums_nccl(buf)
            ms_nccl, _, _ = timed(lambda: predict.allreduce_sums_nccl(buf), 50)
            peer = predict._PEER.get(dist.group.WORLD)
            via = ("bnn_peer_allreduce: one kernel over NVLink peer memory, rank-ordered sums"
                   if peer is not None and peer.disabled is None else "NCCL")
            if peer is not None:
                peer.check()
            t_n1 = ms1 / n1
            extra = {"n1_ms_per_step": t_n1, "speedup_vs_n1": t_n1 / (ms / steps),
                     "efficiency_vs_n1": t_n1 / (ms / steps) / world, "collective_us": ms_c / 50 * 1e3,
                     "collective_us_nccl": ms_nccl / 50 * 1e3,
                     "collective": "all_reduce(sum) of %d fp32 statistics via %s" % (flat.numel(), via),
                     "ideal_efficiency_with_replicated_prefix":
                         (pre_macs + S * suf_macs) / (world * (pre_macs + (S / world) * suf_macs))}
            rec.update(extra)
        eng.release_buffers()
        del eng, model
        torch.cuda.empty_cache()
        return rec

    head_wl = args.workload
    if world > 1 and head_wl == "c5":
        head = measure("c5", args.steps, True, "samples", args.batch, args.samples)
    else:
        head = measure(head_wl, args.steps, True, "batch" if world > 1 else None, args.batch, args.samples)

    extras, c5_strong = {}, None
    if not args.no_extra_configs and not (args.batch or args.samples) and head_wl == "c2":
        sub_steps = max(5, min(args.steps, 20))
        if world == 1:
            for name in ("c1", "c3", "c4", "c5"):
                extras[name] = measure(name, sub_steps if name != "c5" else max(3, min(args.steps, 8)), False, None)
        else:
            c5_strong = measure("c5", sub_steps, False, "samples")

    if extras and world == 1 and args.dtype == "fp16":
        # BASELINE config 4 with its stochastic suffix in 8 bits (q8.Q8Plan: tcgen05 kind::i8)
        from bayesnn_fpga_b200 import q8
        desc, kind, B, S, classes = WORKLOADS["c4"]
        model = build_model(kind, classes).to(dev)
        xq = torch.randn(B, *input_shape(kind), generator=torch.Generator().manual_seed(100)).to(dev)
        plan = q8.Q8Plan(model, xq[:64], S_calib=4)
        for _ in range(W):
            plan.run(xq, S)
        k8 = max(5, min(args.steps, 20))
        ms8, _, win8 = timed(lambda: plan.run(xq, S), k8)
        pre_macs, suf_macs = plan.eng.graph.macs()
        extras["c4_int8_suffix"] = {
            "workload": desc + " - stochastic suffix on unsigned 8-bit activations x signed 8-bit weights (tcgen05 "
                               "kind::i8, int32 accumulate), prefix and classifiers unchanged",
            "value": B * k8 / (ms8 * 1e-3), "unit": "images/s", "ms_per_step": ms8 / k8, "steps": k8,
            "suffix_tera_ops_per_s": 2.0 * B * S * suf_macs / (ms8 / k8 * 1e-3) / 1e12,
            "speedup_vs_fp16_c4": extras["c4"]["ms_per_step"] / (ms8 / k8), "cuda_graph": False,
            "clocks": sampler.window(*win8) if rank == 0 else None}
        del plan, model
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        desc, kind, B, S, classes = WORKLOADS[head_wl]
        cpu = time_cpu_reference(kind, classes, args.samples or S, args.batch or B)
    if rank == 0:
        sampler.stop()
        desc, kind, B, S, classes = WORKLOADS[head_wl]
        shard_samples = head["partition"] == "samples"
        line = {
            "metric": METRIC, "value": head["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if shard_samples else "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.dtype],
            "data": "synthetic",
            "config": {"workload": head["workload"] if not (args.batch or args.samples) else
                       head["workload"] + " [overridden B=%d S=%d]" % (head["batch_per_gpu"], head["S"]),
                       "batch_per_gpu": head["batch_per_gpu"], "global_batch": head["global_batch"], "S": head["S"],
                       "exits": head["exits"], "classes": head["classes"], "partition": head["partition"],
                       "arithmetic": ("%s operands, f32 accumulate (tcgen05.mma kind::f16), f32 statistics" %
                                      {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.dtype]
                                      if args.dtype != "fp32" else "f32 CUDA-core kernels, f32 statistics"),
                       "l2": "per-step activation working set (GiB) >> 126 MB L2; no explicit flush",
                       "cuda_graph": os.environ.get("BNN_CUDA_GRAPH", "1") != "0",
                       "collective_inside_cuda_graph": head.get("collective_in_graph"),
                       "collective": (None if world == 1 else
                                      "all_reduce(sum) of the per-exit sums before the finaliser" if shard_samples else
                                      ("all_gather_into_tensor of the finished statistics, on the compute stream"
                                       if WEAK_GATHER else "none: every rank owns its batch (independent replicas)")),
                       "algorithmic_gflop_per_image": head["algorithmic_gflop_per_image"]},
            "tflops_whole_step": head["tflops_whole_step"],
            "frac_of_tensor_peak_whole_step": head["frac_of_tensor_peak_whole_step"],
            "per_rank_ms_per_step": head["per_rank_ms_per_step"],
            "clocks": head["clocks"],
            "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"],
            "kernels": head["kernels"],
            "cpu_baseline": cpu,
        }
        for k in ("n1_ms_per_step", "speedup_vs_n1", "efficiency_vs_n1", "collective_us"):
            if k in head:
                line[k] = head[k]
        if extras:
            line["configs"] = extras
        if c5_strong is not None:
            line["c5_strong"] = c5_strong
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
