"""CPU, world_size 2 over gloo: the multi-GPU host logic of the path - sample sharding with global
Philox sample indices + ONE all-reduce of the per-exit sums reproduces the single-process statistics."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesnn_fpga_b200 import predict
from oracle import nets, seeded, stats


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial_sums(sd, x, spec, seed, start, count):
    """What one rank's exit heads accumulate: sums over its samples of probs / logits / p log p."""
    with torch.no_grad():
        r = stats.mc_get_output(lambda i: nets.lenet_forward(sd, x, nets.InjectedSites(spec, seed, start + i)), count)
    sp, sl = r["all_probs"].sum(0), r["all_logits"].sum(0)
    spl = (r["all_probs"] * np.log(r["all_probs"])).sum(-1).sum(0)
    return np.concatenate([sp.ravel(), sl.ravel(), spl.ravel()])


def _worker(rank, world, port, S, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sd = seeded.seeded_state_dict(nets.LENET_SHAPES, seed=3)
    x = seeded.seeded_input((5, 1, 28, 28), seed=4)
    spec = nets.SiteSpec("mc", 0.2)
    start, count = predict.shard_samples(S, world, rank)
    flat = torch.from_numpy(_partial_sums(sd, x, spec, 0x5EED, start, count))
    predict.allreduce_sums(flat)                      # the path's single collective
    if rank == 0:
        np.save(out, flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sample_sharding_allreduce_equals_single_process(tmp_path):
    S, world = 7, 2                                   # ragged split: 4 + 3
    out = str(tmp_path / "sums.npy")
    mp.spawn(_worker, args=(world, _free_port(), S, out), nprocs=world, join=True)
    got = np.load(out)
    sd = seeded.seeded_state_dict(nets.LENET_SHAPES, seed=3)
    x = seeded.seeded_input((5, 1, 28, 28), seed=4)
    want = _partial_sums(sd, x, nets.SiteSpec("mc", 0.2), 0x5EED, 0, S)
    assert np.abs(got - want).max() < 1e-9
