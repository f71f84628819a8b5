"""The reference's OWN CPU implementation of the hot path, loaded from ``oracle/_ref`` (staged by oracle/make_ref.py)
and driven exactly as the reference drives it (TEST INFRASTRUCTURE: ``bench.py --impl reference`` / ``cpu_baseline``
and the tests that validate the oracle against it; nothing in ``bayesnn_fpga_b200`` imports this).

* models: ``models.model_loader.get_network(hyperparams)`` of the staged tree (models/model_loader.py:8-23) with the
  hyper-parameter dict the reference's drivers build (train/hyperparameters.py:86-109) - the unmodified
  ``ResNet18MCEarlyExit`` / ``VGG19MCEarlyExit`` classes with their own ``MCDropout`` / ``Masksembles`` layers and
  torch's own RNG;
* the S-pass loop: the SOURCE of ``FullAnalysis._get_output`` (train/results_analyzer.py:236-270) exec'd unchanged
  inside a stub class - the module itself cannot be imported (KDEpy / matplotlib / sacred are not installed).

BASELINE config 4 (VGG-19 with dropout on the last three blocks) is built the only way the reference can build it
(SURVEY.md 8c): ``dropout=None`` + ``blocks[i].append(MCDropout(p))`` with the reference's own MCDropout class - its
``dropout="block"`` constructor path raises AttributeError (vgg19.py:365).  Config 1 (LeNet) has no PyTorch reference.
"""
import importlib
import json
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SW = os.path.join(REF, "Software_Artifact", "software")


def available():
    return os.path.exists(os.path.join(REF, "MANIFEST.json")) and os.path.isdir(SW)


def _import_models():
    if SW not in sys.path:
        sys.path.insert(0, SW)
    for name in ("utils", "models"):           # the staged tree must win over anything of the same name
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(SW):
            del sys.modules[name]
    return importlib.import_module("models.model_loader"), importlib.import_module("models.vgg19.vgg19")


def get_output_stub():
    """A class holding the reference's `_get_output` source (results_analyzer.py:236-270), unchanged."""
    src = open(os.path.join(SW, "train", "results_analyzer.py")).read().split("\n")
    body = "\n".join(src[235:270])
    assert body.lstrip().startswith("def _get_output(self, b_x):"), "reference line numbers moved"
    ns = {"np": np, "torch": torch, "nn": torch.nn}
    exec("class Stub:\n" + textwrap.indent(textwrap.dedent(body), "    "), ns)
    return ns["Stub"]


def build(kind, classes, seed_bn=True):
    """The reference's own network for a bench.py workload kind ('resnet_mcd' | 'resnet_mask' | 'vgg_last3')."""
    loader, vgg = _import_models()
    torch.manual_seed(0)
    np.random.seed(0)
    base = dict(load_model=None, out_dim=classes, dropout_exit=True, dropout_p=0.5, mask_type="mc", num_masks=4,
                mask_scale=4.0)
    if kind == "resnet_mcd":
        m = loader.get_network(dict(base, call="ResNet18", resnet_type="mc_early_exit", dropout="block", n_exits=4))
    elif kind == "resnet_mask":
        m = loader.get_network(dict(base, call="ResNet18", resnet_type="mc_early_exit", dropout="block", n_exits=4,
                                    mask_type="mask", mask_scale=2.0))
    elif kind == "vgg_last3":
        m = loader.get_network(dict(base, call="VGG19", resnet_type="mc_early_exit", dropout=None, n_exits=5,
                                    image_size=32))
        for i in (2, 3, 4):
            m.blocks[i].append(vgg.MCDropout(0.5))
    else:
        raise ValueError("no PyTorch reference for workload kind %r" % (kind,))
    if seed_bn:                                  # same non-trivial BatchNorm statistics as bench.build_model
        g = torch.Generator().manual_seed(1)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1, generator=g)
                mod.running_var.uniform_(0.5, 1.5, generator=g)
                mod.weight.data.uniform_(0.5, 1.5, generator=g)
                mod.bias.data.normal_(0, 0.1, generator=g)
    return m.eval()


class ReferenceRunner:
    """`FullAnalysis._get_output` of the reference over its own model, on the host cores."""

    def __init__(self, kind, classes, S):
        self.model = build(kind, classes)
        self.stub = get_output_stub()()
        self.stub.model = self.model
        self.stub.mc_dropout = True
        self.stub.mc_passes = int(S)
        n_exits = 5 if kind == "vgg_last3" else 4
        self.stub.outputs = list(range(n_exits))

    def step(self, x):
        with torch.no_grad():
            return self.stub._get_output(x)


def manifest():
    with open(os.path.join(REF, "MANIFEST.json")) as f:
        return json.load(f)
