"""GPU (needs >= 2 devices, skipped otherwise): sample sharding over NCCL reproduces the single-GPU statistics.
The host logic of the same path is covered on CPU by tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from bayesnn_fpga_b200 import mc_predict
from tests.cases import build_seeded
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
model, sd, gold = build_seeded("resnet18_mcd_block")
model.cuda()
x = torch.from_numpy(gold["x"])
S = 7                                             # ragged over 2 ranks: 4 + 3
for dt, tol in (("fp32", 2e-6), ("fp16", 2e-6)):
    r = mc_predict(model, x, S, seed=11, dtype=dt, distributed=True)
    shard = [t.clone() for t in (r.mean_probs, r.mean_logits, r.entropy, r.expected_entropy)]
    full = mc_predict(model, x, S, seed=11, dtype=dt, distributed=False)
    for a, b in zip(shard, (full.mean_probs, full.mean_logits, full.entropy, full.expected_entropy)):
        err = (a - b).abs().max().item()
        assert err <= tol * max(1.0, b.abs().max().item()), (dt, err)
dist.barrier()
if rank == 0:
    print("MULTI_GPU_OK")
dist.destroy_process_group()
'''


def test_sample_sharding_over_nccl_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
