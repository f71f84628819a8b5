"""Freeze golden vectors of the PyTorch NN -> BNN converter from the reference's OWN modules (build container only).

    python tests/golden/make_golden_converter.py     # needs /root/reference, writes tests/golden/converter.npz

What runs here is the reference's code: `Hardware_Artifact/converter/pytorch/Dropouts.py` is imported as is
(`BayesianDropout`, `BayesianDropout2D`, `BayesianDropout3D`), and `nn2bnn.py` is imported as is after satisfying its
stray `from test.ThreeLayerNet import ThreeLayerNet` (a git-ignored file) with an empty stand-in module.
`_convert_model` (nn2bnn.py:32-45) converts each network; the eval forward of `MCDropout` (nn2bnn.py:19-27) raises
`UnboundLocalError` as shipped (line 25 prints `pred` before assigning it), so its two working lines are executed
literally: `pred = [self.model(x) for _ in range(self.nSamples)]; sum(pred) / len(pred)`.

torch's dropout RNG is not reproducible across devices, so the masks are INJECTED: the name `F` inside the reference's
Dropouts module is bound to a shim whose dropout / dropout2d / dropout3d multiply by the Philox keep masks of the mask
contract (oracle/philox.py) - stream id = order of first use of a wrapper inside one forward, sample = pass index -
and by 1/(1-p).  Everything else (which leaves get wrapped, with which dropout flavour, in which order they execute,
the mean over nSamples) is the reference's behaviour and is what the fixture pins.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/Hardware_Artifact/converter/pytorch"

from oracle import philox, seeded      # noqa: E402
from tests.nets_converter import NETS   # noqa: E402


class InjectedF:
    """Stands in for `torch.nn.functional` inside the reference's Dropouts module."""

    def __init__(self, seed):
        self.seed, self.sample, self.calls = seed, 0, 0

    def next_pass(self):
        self.sample += 1
        self.calls = 0

    def _apply(self, x, p, mode):
        stream = self.calls            # wrappers execute in the same order in every pass
        self.calls += 1
        if p >= 1.0:
            return torch.zeros_like(x)
        keep = philox.keep_mask(self.seed, stream, self.sample, x.shape, p, mode)
        scale = np.float32(1.0) / np.float32(1.0 - p)
        return x * (torch.from_numpy(keep).to(x.dtype) * float(scale))

    def dropout(self, x, p, training, inplace):
        assert training is True
        return self._apply(x, p, "element")

    def dropout2d(self, x, p, training, inplace):
        assert training is True
        return self._apply(x, p, "channel")

    dropout3d = dropout2d


def load_reference():
    sys.path.insert(0, REF)
    stub = types.ModuleType("test.ThreeLayerNet")
    stub.ThreeLayerNet = object
    sys.modules["test.ThreeLayerNet"] = stub
    import Dropouts as ref_dropouts
    import nn2bnn as ref_nn2bnn
    return ref_dropouts, ref_nn2bnn


def structure(model):
    """'path:ClassName(p)' of every module, in registration order - what _convert_model did to the network."""
    return [("%s:%s" % (n, type(m).__name__)) + ("(%g)" % m.p if hasattr(m, "p") and hasattr(m, "layer") else "")
            for n, m in model.named_modules()]


def main():
    ref_dropouts, ref_nn2bnn = load_reference()
    out = {}
    for tag, (make, in_shape, p, n_samples, seed) in NETS.items():
        torch.manual_seed(0)
        net = make()
        sd = seeded.seeded_state_dict(net.state_dict(), seed=77)
        net.load_state_dict(sd)
        wrapper = ref_nn2bnn.MCDropout(net, nSamples=n_samples, p=p).eval()
        x = seeded.seeded_input((5,) + in_shape, seed=31)
        shim = InjectedF(seed)
        ref_dropouts.F = shim
        try:
            with torch.no_grad():
                pred = []
                for _ in range(wrapper.nSamples):          # nn2bnn.py:26
                    pred.append(wrapper.model(x))
                    shim.next_pass()
                mean = sum(pred) / len(pred)                # nn2bnn.py:27
        finally:
            ref_dropouts.F = torch.nn.functional
        out[tag + "/x"] = x.numpy()
        out[tag + "/passes"] = np.stack([t.numpy() for t in pred])
        out[tag + "/mean"] = mean.numpy()
        out[tag + "/structure"] = np.array(structure(wrapper.model))
        out[tag + "/repr"] = np.array([wrapper.extra_repr()])
        print("%-12s %d wrapped leaves, output %s, |mean| max %.4f" % (
            tag, sum(":Bayesian" in s for s in structure(wrapper.model)), tuple(mean.shape), float(mean.abs().max())))
    # constructor contract of the wrappers (Dropouts.py:13-23)
    for bad in (-0.1, 1.5):
        try:
            ref_dropouts.BayesianDropout(nn.Linear(2, 2), bad)
            raise AssertionError("p=%g accepted" % bad)
        except ValueError as e:
            out["err/%g" % bad] = np.array([str(e)])
    out["extra_repr"] = np.array([ref_dropouts.BayesianDropout2D(nn.Conv2d(1, 1, 1), 0.25).extra_repr()])
    np.savez_compressed(os.path.join(HERE, "converter.npz"), **out)


if __name__ == "__main__":
    main()
