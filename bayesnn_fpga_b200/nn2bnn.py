"""PyTorch NN -> BNN converter - drop-in for the reference's
``Hardware_Artifact/converter/pytorch/nn2bnn.py`` (``MCDropout`` :7-30, ``_convert_model`` :32-45).

Same type map and the same recursive ``named_children`` walk.  The reference's eval branch raises
``UnboundLocalError`` as shipped (``pred`` is printed before assignment, nn2bnn.py:25);
the intended semantics - mean over ``nSamples`` stochastic passes of the RAW model output
(nn2bnn.py:26-27) - is what this class implements.

Eval mode runs the converted network through the fused plan (``lowering.lower_module`` -> ``engine.Engine``): the
layers in front of the first wrapped leaf once per image, the rest once per sample, the mean accumulated on the
device.  The masks are the ones ``nSamples`` successive stand-alone forward calls would draw (every wrapper keeps its
own Philox stream and call counter), so both forms agree to rounding.  A network the lowering cannot express RAISES
(``NotImplementedError`` naming the construct): there is no silent library path.  ``eager_fallback=True`` opts in to the
reference's literal loop instead - ``nSamples`` stand-alone calls, each leaf through torch (cuBLAS / cuDNN) and each
dropout through ``bnn_dropout`` - with a warning.
"""
import warnings

import torch
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


# ---- insertion strategies (the Keras converter's, Hardware_Artifact/converter/keras/nn2bnn.py:9-72, for PyTorch) -------
def _leaves(model):
    """(parent, attribute name, module) of every leaf module in registration order = `model.layers` of a Keras
    Sequential / functional model."""
    out = []
    for parent in model.modules():
        for name, child in parent.named_children():
            if next(child.children(), None) is None:
                out.append((parent, name, child))
    return out


def _is_weight_layer(m):
    return type(m) is nn.Conv2d or isinstance(m, nn.Linear)        # `type(layer) in conv_list or isinstance(layer, Dense)`


def _default_strategy(layers, num=1):
    """A Bayesian layer in front of each of the LAST `num` dense / conv layers (:9-29)."""
    pos = [i - 1 for i, m in enumerate(layers) if _is_weight_layer(m)]
    return pos[max(0, len(pos) - num):]


def _last_strategy(layers, num=1):
    """`num` Bayesian layers going backward from the input of the first dense layer behind the last conv layer (:31-59)."""
    after_conv, last_conv, count = False, -1, -1
    for i, m in enumerate(layers):
        if type(m) is nn.Conv2d:
            last_conv, after_conv = i, True
        if isinstance(m, nn.Linear) and after_conv:
            count, after_conv = i - 1, False
    if last_conv < 0:
        return []
    if count < 0:
        count = last_conv
    pos = []
    while num > 0 and count >= 0:
        pos.append(count)
        count -= 1
        num -= 1
    return pos


def _full_strategy(layers, num=None):
    """A Bayesian layer in front of every dense / conv layer (:61-72)."""
    return [i - 1 for i, m in enumerate(layers) if _is_weight_layer(m)]


strategy_fn = {"default": _default_strategy, "last": _last_strategy, "full": _full_strategy}


def convert_model(model, strategy="default", num=1, type="BayesianDropout", p=0.5, n=4, scale=2.0):
    """Insert Bayesian layers into `model` (in place) with one of the Keras converter's strategies.  Position k means
    "behind layer k" exactly as there (`supported_layers[k]`, :120-128), i.e. in front of layer k + 1, which is replaced
    by ``nn.Sequential(site, layer)``; a position in front of the first layer is ignored like in the reference.
    type: "BayesianDropout" (element-wise MC dropout, converter/keras/MCDropout.py:37-38) or "Masksembles" (n masks,
    `scale`; the mask width is the consuming layer's input width)."""
    from .Dropouts import MCDropout as _Site
    from .utils import Masksembles1D, Masksembles2D
    if strategy not in strategy_fn:
        raise ValueError("unknown strategy %r (expected one of %s)" % (strategy, sorted(strategy_fn)))
    if type not in ("BayesianDropout", "Masksembles"):
        raise Exception("The type %s is not supported yet!" % type)
    leaves = _leaves(model)
    layers = [m for _, _, m in leaves]
    for k in sorted(set(strategy_fn[strategy](layers, num))):
        if k < 0 or k + 1 >= len(leaves):
            continue
        parent, name, layer = leaves[k + 1]
        if type == "BayesianDropout":
            site = _Site(p)
        elif isinstance(layer, nn.Conv2d):
            site = Masksembles2D(layer.in_channels, n, scale)
        elif isinstance(layer, nn.Linear):
            site = Masksembles1D(layer.in_features, n, scale)
        else:
            raise ValueError("Masksembles in front of %s: the channel count is not known" % layer.__class__.__name__)
        setattr(parent, name, nn.Sequential(site, layer))
    return model


class MCDropout(nn.Module):
    """Monte-Carlo-dropout wrapper: training -> one stochastic pass; eval -> mean of ``nSamples`` passes.
    ``strategy`` (None | "default" | "last" | "full", with ``num``) selects the Keras converter's insertion strategies
    instead of the PyTorch converter's wrap-every-leaf rule."""

    def __init__(self, model, nSamples=10, p=0.5, dtype="fp32", fused=True, strategy=None, num=1,
                 eager_fallback=False):
        super().__init__()
        self.model = _convert_model(model, p) if strategy is None else convert_model(model, strategy, num, p=p)
        self.nSamples = nSamples
        self.p = p
        self.bnn_dtype = dtype          # "fp32": exact CUDA-core path; "fp16" / "bf16": tensor cores
        self.bnn_fused = fused          # False: the stand-alone loop, explicitly requested
        self.bnn_eager_fallback = eager_fallback

    def _site_modules(self):
        return [m for m in self.model.modules() if hasattr(m, "bnn_calls")]

    def _plan(self, shape, device):
        """(engine, site modules) for inputs [B, *shape], or None when the network cannot be lowered."""
        from . import _plans
        key = ("nn2bnn", tuple(shape), self.bnn_dtype, str(device))
        plans = _plans.plans_for(self)       # outside the module (picklable), dropped when the parameters change
        if key not in plans:
            from . import engine, lowering
            try:
                graph, sites = lowering.lower_module(self.model, shape)
                plans[key] = (engine.Engine(graph, dtype=self.bnn_dtype, device=device), [m for _, m in sites])
            except NotImplementedError as e:
                if not self.bnn_eager_fallback:
                    raise NotImplementedError(
                        "nn2bnn.MCDropout: this network cannot be lowered to the fused B200 plan (%s). Pass "
                        "eager_fallback=True to run %d stand-alone passes through torch's own layers instead." % (
                            e, self.nSamples)) from e
                warnings.warn("nn2bnn.MCDropout: running %d stand-alone passes instead of the fused plan: %s" % (
                    self.nSamples, e))
                plans[key] = None
        return plans[key]

    def _forward_fused(self, x):
        flat = x.dim() == 2
        shape = (x.shape[1], 1, 1) if flat else tuple(x.shape[1:])
        if len(shape) != 3 or not x.is_cuda:
            return None
        plan = self._plan(shape, x.device)
        if plan is None:
            return None
        eng, mods = plan
        sites = self._site_modules()
        seeds = {int(m.bnn_seed) for m in sites}
        calls = {int(m.bnn_calls) for m in sites}
        if len(seeds) > 1 or len(calls) > 1:
            return None                      # wrappers were driven individually: keep their own streams exact
        r = eng.run(x.reshape(x.shape[0], *shape).float(), self.nSamples, seed=seeds.pop() if seeds else 0,
                    sample0=calls.pop() if calls else 0, S_total=self.nSamples)
        for m in sites:
            m.bnn_calls += self.nSamples
        outs = [r.mean_logits[e].clone() for e in eng.graph.out_order]
        return outs[0] if len(outs) == 1 else outs

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
        if self.bnn_fused:
            out = self._forward_fused(x)
            if out is not None:
                return out
        pred = [self.model(x) for _ in range(self.nSamples)]
        return sum(pred) / len(pred)

    def extra_repr(self) -> str:
        return "nSamples: {}\nprobability: {}".format(self.nSamples, self.p)
