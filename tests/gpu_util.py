"""Helpers of the -m gpu tests: raw C-ABI calls on torch tensors, and an error report that is
written to gpurun_out/ so the numbers can be read back after a gpurun call."""
import ctypes
import json
import os

import torch

from bayesnn_fpga_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def drop_desc(kind=0, p=0.0, seed=0, stream_id=0, sample0=0, batch=1, masks=None, cnt0=0):
    d = _lib.DropDesc()
    d.kind, d.p, d.seed, d.stream_id, d.sample0, d.batch, d.cnt0 = kind, p, seed, stream_id, sample0, batch, cnt0
    if masks is not None:
        d.masks, d.n_masks = masks.data_ptr(), masks.shape[0]
    return d


def report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")


TORCH_DT = {"fp32": (torch.float32, _lib.F32), "fp16": (torch.float16, _lib.F16), "bf16": (torch.bfloat16, _lib.BF16)}
