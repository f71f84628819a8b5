"""``get_network`` - the factory the reference's drivers call (Software_Artifact/software/models/model_loader.py:8-23).

Contract kept: a hyper-parameter dict with the keys ``call`` ("ResNet18" | "VGG19"), ``resnet_type`` (None |
"early_exit" | "mc" | "mc_early_exit"), ``load_model`` (None or the path of a pickled model, optionally with
``gpu_device`` as the map location) and the constructor kwargs of the selected class; an unknown ``call`` raises
``AttributeError``.  A pickle of the reference's OWN model class keeps working downstream: ``mc_predict`` and
``FullAnalysis`` lower such objects through ``lowering.lower_module``.
"""
import torch

from . import resnet18 as _resnet18
from . import vgg19 as _vgg19
from .utils import dict_drop

_FACTORIES = {"ResNet18": _resnet18.get_res_net_18, "VGG19": _vgg19.get_vgg_19}
_NOT_CTOR_KWARGS = ("call", "load_model", "resnet_type")


def _load_pickled(hp):
    path = hp["load_model"]
    if not torch.cuda.is_available():
        where = torch.device("cpu")
    else:
        where = hp.get("gpu_device")          # absent -> torch.load's default placement, like the reference's KeyError branch
    return torch.load(path, map_location=where, weights_only=False)


def get_network(network_hyperparams):
    hp = network_hyperparams
    if hp["load_model"] is not None:
        return _load_pickled(hp)
    make = _FACTORIES.get(hp["call"])
    if make is None:
        raise AttributeError("unknown network %r" % (hp["call"],))
    return make(hp["resnet_type"], dict_drop(hp, *_NOT_CTOR_KWARGS))
