"""PyTorch NN -> BNN converter - drop-in for the reference's
``Hardware_Artifact/converter/pytorch/nn2bnn.py`` (``MCDropout`` :7-30, ``_convert_model`` :32-45).

Same type map and the same recursive ``named_children`` walk.  The reference's eval branch raises
``UnboundLocalError`` as shipped (``pred`` is printed before assignment, nn2bnn.py:25);
the intended semantics - mean over ``nSamples`` stochastic passes of the RAW model output
(nn2bnn.py:26-27) - is what this class implements.
"""
import torch.nn as nn

from .Dropouts import BayesianDropout, BayesianDropout2D, BayesianDropout3D


def _convert_model(model, p):
    base_layers = {nn.Linear: BayesianDropout,
                   nn.MaxPool1d: BayesianDropout,
                   nn.MaxPool2d: BayesianDropout,
                   nn.MaxPool3d: BayesianDropout,
                   nn.Conv1d: BayesianDropout,
                   nn.Conv2d: BayesianDropout2D,
                   nn.Conv3d: BayesianDropout3D}
    if type(model) in base_layers:             # exact type match, like the reference
        return base_layers[type(model)](model, p)
    for name, layer in model.named_children():
        setattr(model, name, _convert_model(layer, p))
    return model


class MCDropout(nn.Module):
    """Monte-Carlo-dropout wrapper: training -> one stochastic pass; eval -> mean of ``nSamples`` passes."""

    def __init__(self, model, nSamples=10, p=0.5):
        super().__init__()
        self.model = _convert_model(model, p)
        self.nSamples = nSamples
        self.p = p

    def reseed(self, seed):
        """Make the run reproducible: site k of the converted model gets Philox stream k."""
        k = 0
        for m in self.model.modules():
            if hasattr(m, "reseed"):
                m.reseed(seed, k)
                k += 1
        return self

    def forward(self, x):
        if self.training:
            return self.model(x)
        pred = [self.model(x) for _ in range(self.nSamples)]
        return sum(pred) / len(pred)

    def extra_repr(self) -> str:
        return "nSamples: {}\nprobability: {}".format(self.nSamples, self.p)
