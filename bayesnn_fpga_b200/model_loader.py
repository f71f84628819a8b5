"""``get_network`` - drop-in for ``Software_Artifact/software/models/model_loader.py:8-23``.

``network_hyperparams["load_model"]`` (a pickled model file) is loaded with ``torch.load`` exactly like there - a
pickle of the reference's OWN model class keeps working, because ``mc_predict`` / ``FullAnalysis`` lower such objects
through ``lowering.lower_module``; otherwise ``"call"`` selects the ResNet-18 / VGG-19 factory and ``"resnet_type"``
the kind (None | "early_exit" | "mc" | "mc_early_exit").  Unknown ``call`` -> ``AttributeError`` like the reference.
"""
import torch

from .resnet18 import get_res_net_18
from .utils import dict_drop
from .vgg19 import get_vgg_19


def get_network(network_hyperparams):
    if network_hyperparams["load_model"] is not None:
        if torch.cuda.is_available():
            try:
                model = torch.load(network_hyperparams["load_model"], map_location=network_hyperparams["gpu_device"],
                                   weights_only=False)
            except KeyError:
                model = torch.load(network_hyperparams["load_model"], weights_only=False)
        else:
            model = torch.load(network_hyperparams["load_model"], map_location=torch.device("cpu"), weights_only=False)
    elif network_hyperparams["call"] == "ResNet18":
        model = get_res_net_18(network_hyperparams["resnet_type"],
                               dict_drop(network_hyperparams, "call", "load_model", "resnet_type"))
    elif network_hyperparams["call"] == "VGG19":
        model = get_vgg_19(network_hyperparams["resnet_type"],
                           dict_drop(network_hyperparams, "call", "load_model", "resnet_type"))
    else:
        raise AttributeError
    return model
