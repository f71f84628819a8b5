"""GPU (needs >= 2 devices, skipped otherwise): sample sharding over NCCL reproduces the single-GPU statistics.
The host logic of the same path is covered on CPU by tests/test_dist_gloo.py.

Two cases, each launched as `torch.distributed.run` with one rank per GPU:
* the fixture network (B=4) with a RAGGED split (S=7) at world size 2;
* BASELINE configs[4] (C5: ResNet-18 ME MCD, S=128, B=256, fp16 operands) over ALL visible GPUs (2, 4 or 8):
  every rank's per-sample logits must be BIT-IDENTICAL to the single-GPU run's logits of the same GLOBAL sample
  indices (Philox counters carry the global index, so the masks do not depend on the sharding), and the all-reduced
  statistics must equal the single-GPU ones up to fp32 summation order.
Each worker appends its measured errors to gpurun_out/multi_gpu_report.jsonl."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from bayesnn_fpga_b200 import mc_predict, predict
from tests.cases import build_seeded
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
case = %(case)r
rep = {"case": case, "world": world, "rank": rank}
if case == "peer":
    # the all-reduce kernel over NVLink peer memory (csrc/kernels_comm.cu) against NCCL on random payloads: several sizes
    # (one not a multiple of 4, one tiny), several epochs each, eager and replayed from a CUDA graph
    g = torch.Generator(device="cuda").manual_seed(1000 + rank)
    worst = 0.0
    for n in (21504, 1031, 7, 262144):
        for it in range(4):
            flat = torch.randn(n, device="cuda", generator=g)
            want = predict.allreduce_sums_nccl(flat.clone())
            got = predict.allreduce_sums(flat.clone())
            worst = max(worst, ((got - want).abs().max() / want.abs().max()).item())
            everyone = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(everyone, got)
            assert all(torch.equal(everyone[0], t) for t in everyone), "ranks disagree on the bits of the total"
    red = predict._PEER[dist.group.WORLD]
    assert red.disabled is None, "peer-memory path not taken: %%s" %% red.disabled
    buf = torch.zeros(21504, device="cuda")
    predict.allreduce_sums(buf)                               # workspace exists before the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        predict.allreduce_sums(buf)
    for it in range(3):
        src = torch.randn(21504, device="cuda", generator=g)
        want = predict.allreduce_sums_nccl(src.clone())
        buf.copy_(src)
        graph.replay()
        worst = max(worst, ((buf - want).abs().max() / want.abs().max()).item())
    red.check()                                               # no peer ever timed out
    assert worst <= 1e-6, worst
    rep["peer_allreduce_max_rel_err_vs_nccl"] = worst
    dtypes = ()
elif case == "ragged":
    model, sd, gold = build_seeded("resnet18_mcd_block")
    model.cuda()
    x = torch.from_numpy(gold["x"])
    S, dtypes = 7, (("fp32", 2e-6), ("fp16", 2e-6))          # ragged over 2 ranks: 4 + 3
else:
    import bench                                              # the benchmark's own C5 model (configs[4])
    model = bench.build_model("resnet_mcd", 10).cuda()
    x = torch.randn(256, 3, 32, 32, generator=torch.Generator().manual_seed(100))
    S, dtypes = 128, (("fp16", 2e-6),)
for dt, tol in dtypes:
    r = mc_predict(model, x, S, seed=11, dtype=dt, distributed=True, want_logits=True)
    s0, sl = predict.shard_samples(S, world, rank)
    shard = [t.clone() for t in (r.mean_probs, r.mean_logits, r.entropy, r.expected_entropy, r.ens_probs)]
    shard_logits = r.all_logits.clone()                       # [S_local, E, B, C] of this rank's GLOBAL samples
    assert shard_logits.shape[0] == sl
    del r
    model.bnn_engine(dt).release_buffers()
    full = mc_predict(model, x, S, seed=11, dtype=dt, distributed=False, want_logits=True)
    errs = []
    for a, b in zip(shard, (full.mean_probs, full.mean_logits, full.entropy, full.expected_entropy, full.ens_probs)):
        err = (a - b).abs().max().item()
        errs.append(err)
        assert err <= tol * max(1.0, b.abs().max().item()), (dt, err)
    same = torch.equal(shard_logits, full.all_logits[s0:s0 + sl])
    rep[dt] = {"max_err_stats": max(errs), "per_sample_logits_bit_identical": bool(same), "samples": [s0, sl]}
    assert same, "per-sample logits of global samples [%%d, %%d) differ from the single-GPU run" %% (s0, s0 + sl)
    model.bnn_engine(dt).release_buffers()
os.makedirs(os.path.join(%(root)r, "gpurun_out"), exist_ok=True)
with open(os.path.join(%(root)r, "gpurun_out", "multi_gpu_report.jsonl"), "a") as f:
    f.write(json.dumps(rep) + "\n")
dist.barrier()
if rank == 0:
    print("MULTI_GPU_OK")
dist.destroy_process_group()
'''


def _run(tmp_path, case, world, port):
    script = tmp_path / ("worker_%s.py" % case)
    script.write_text(WORKER % {"root": ROOT, "case": case})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_sample_sharding_over_nccl_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(tmp_path, "ragged", 2, 29533)


def test_peer_memory_allreduce_kernel_matches_nccl(tmp_path):
    """bnn_peer_allreduce (one kernel over NVLink peer loads, rank-ordered sums) == NCCL's all-reduce up to fp32
    summation order, bit-identical across ranks, eager and inside a CUDA graph, over all visible GPUs."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    _run(tmp_path, "peer", 8 if n >= 8 else (4 if n >= 4 else 2), 29535)


def test_c5_s128_sample_sharded_over_all_gpus_is_bit_identical_per_sample(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    _run(tmp_path, "c5", world, 29534)
