"""Deterministic synthetic parameters (test infrastructure, see oracle/__init__.py).

Neither datasets nor checkpoints are available offline, so every parity case runs on seeded
synthetic weights.  They are generated from the Philox restatement (not from torch's RNG) so
that the build container (where the live reference is imported to freeze the golden vectors)
and the GPU box (where the reference does not exist) see bit-identical parameters.

Distributions follow the reference initialisers (resnet18.py:120-126, vgg19.py:98-108) with
NON-trivial BatchNorm statistics so that BN folding is really exercised.
"""
import zlib

import numpy as np
import torch

from . import philox


def _stream(name):
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


def seeded_state_dict(template, seed=1234, keep=("masks",)):
    """template: mapping name -> tensor (or shape tuple). Returns name -> float32 tensor.
    Entries whose last path component is in ``keep`` are passed through untouched (Masksembles
    masks come from the reference generator, not from here)."""
    out = {}
    seen = {}          # storage -> first name (the reference VGG registers every block twice,
    #                    vgg19.py:92, so its state_dict lists the same tensor under two names)
    for name, ref in template.items():
        shape = tuple(ref.shape) if hasattr(ref, "shape") else tuple(ref)
        leaf = name.split(".")[-1]
        if hasattr(ref, "data_ptr") and ref.numel() > 0:
            key = (ref.data_ptr(), tuple(ref.shape))
            if key in seen:
                out[name] = out[seen[key]]
                continue
            seen[key] = name
        if leaf in keep:
            out[name] = ref.clone() if hasattr(ref, "clone") else ref
            continue
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros((), dtype=torch.long)
            continue
        n = int(np.prod(shape)) if shape else 1
        st = _stream(name)
        if leaf == "running_var":
            v = 0.5 + philox.uniform01(seed, st, 0, n)
        elif leaf == "running_mean":
            v = 0.1 * philox.normal(seed, st, 0, n)
        elif leaf == "weight" and len(shape) == 1:            # BN gamma
            v = 0.5 + philox.uniform01(seed, st, 0, n)
        elif leaf == "bias":
            v = 0.1 * philox.normal(seed, st, 0, n)
        elif leaf == "weight" and len(shape) == 4:            # conv: N(0, sqrt(2/(k*k*Cout)))
            std = np.sqrt(2.0 / (shape[2] * shape[3] * shape[0]))
            v = std * philox.normal(seed, st, 0, n)
        elif leaf == "weight" and len(shape) == 2:            # linear: U(+-1/sqrt(fan_in))
            bound = 1.0 / np.sqrt(shape[1])
            v = bound * (2.0 * philox.uniform01(seed, st, 0, n) - 1.0)
        else:
            v = philox.normal(seed, st, 0, n)
        out[name] = torch.from_numpy(v.astype(np.float32).reshape(shape))
    return out


def seeded_input(shape, seed=99):
    """x ~ N(0,1) float32 (Gaussian-noise inputs, precedent data_utils.py:73-90)."""
    n = int(np.prod(shape))
    return torch.from_numpy(philox.normal(seed, 0x1A2B, 0, n).astype(np.float32).reshape(shape))


def seeded_labels(n, classes, seed=7):
    w = philox.random_words(seed, 0x3C4D, 0, n)
    return (w % np.uint32(classes)).astype(np.int64)
