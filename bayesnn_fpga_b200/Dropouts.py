"""Bayesian dropout wrappers - drop-in for the reference's
``Hardware_Artifact/converter/pytorch/Dropouts.py`` (``_DropoutBase`` :5-23, ``BayesianDropout`` :25-34,
``BayesianDropout2D`` :36-45, ``BayesianDropout3D`` :47-56) plus the models' ``MCDropout``
(``Software_Artifact/software/models/resnet18/resnet18.py:207-210``, ``vgg19.py:384-387``).

Every class applies dropout with ``training=True`` regardless of ``.eval()``, like the reference;
the mask comes from the counter-based Philox stream of the C-ABI library (``bnn_dropout``) instead of
torch's global RNG.  Each module owns a (seed, stream_id) pair and a call counter, so successive
forward calls draw successive Monte-Carlo samples and a run is reproducible from its seed.
"""
import itertools

import torch.nn as nn

from . import _lib
from .utils import _site_forward

_stream_ids = itertools.count(1 << 16)      # stand-alone modules get ids outside the planners' range
DEFAULT_SEED = 0x5EED


class _PhiloxSite:
    """Mixin: per-module Philox coordinates."""

    def _init_site(self):
        self.bnn_seed = DEFAULT_SEED
        self.bnn_stream = next(_stream_ids)
        self.bnn_calls = 0

    def reseed(self, seed, stream_id=None):
        self.bnn_seed = int(seed)
        if stream_id is not None:
            self.bnn_stream = int(stream_id)
        self.bnn_calls = 0
        return self

    def _draw(self, x, kind, p):
        sample = self.bnn_calls
        self.bnn_calls += 1
        return _site_forward(x, kind, p, None, None, self.bnn_seed, self.bnn_stream, sample)


class MCDropout(nn.Dropout, _PhiloxSite):
    """``F.dropout(x, p, training=True)`` in every mode (resnet18.py:207-210)."""

    def __init__(self, p: float = 0.5, inplace: bool = False):
        super().__init__(p, inplace)
        self._init_site()

    def forward(self, x):
        out = self._draw(x, _lib.DROP_ELEMENT, self.p)
        if self.inplace:
            x.copy_(out)
            return x
        return out


class _DropoutBase(nn.Module, _PhiloxSite):
    """Dropouts.py:5-23: wraps ``layer``; ValueError unless 0 <= p <= 1."""
    __constants__ = ['p', 'inplace']
    p: float
    inplace: bool

    def __init__(self, layer: nn.Module, p: float = 0.5, inplace: bool = False) -> None:
        super().__init__()
        if p < 0 or p > 1:
            raise ValueError("dropout probability has to be between 0 and 1, "
                             "but got {}".format(p))
        self.layer = layer
        self.p = p
        self.inplace = inplace
        self._init_site()

    def extra_repr(self) -> str:
        return 'p={}, inplace={}'.format(self.p, self.inplace)


class BayesianDropout(_DropoutBase):
    """layer(x) then element-wise dropout, always on (Dropouts.py:25-34)."""

    def forward(self, input):
        return self._draw(self.layer(input), _lib.DROP_ELEMENT, self.p)


class BayesianDropout2D(_DropoutBase):
    """layer(x) then channel-wise dropout (``F.dropout2d``), always on (Dropouts.py:36-45)."""

    def forward(self, input):
        return self._draw(self.layer(input), _lib.DROP_CHANNEL, self.p)


class BayesianDropout3D(_DropoutBase):
    """layer(x) then channel-wise dropout over (D, H, W) volumes (``F.dropout3d``; Dropouts.py:47-56)."""

    def forward(self, input):
        return self._draw(self.layer(input), _lib.DROP_CHANNEL, self.p)
