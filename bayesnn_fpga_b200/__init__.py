"""bayesnn_fpga_b200 - B200-native multi-exit MC-dropout / Masksembles inference path.

Drop-in for the PyTorch half of os-hxfan/BayesNN_FPGA on this path:

    reference module                                   here
    Software_Artifact/software/utils.py                bayesnn_fpga_b200.utils
    .../models/resnet18/resnet18.py                    bayesnn_fpga_b200.resnet18
    .../models/vgg19/vgg19.py                          bayesnn_fpga_b200.vgg19
    .../train/results_analyzer.py (FullAnalysis)       bayesnn_fpga_b200.results_analyzer
    Hardware_Artifact/converter/pytorch/Dropouts.py    bayesnn_fpga_b200.Dropouts
    .../models/model_loader.py (get_network)           bayesnn_fpga_b200.model_loader
    Hardware_Artifact/converter/pytorch/nn2bnn.py      bayesnn_fpga_b200.nn2bnn  (+ the Keras converter's strategies)
    any other nn.Module / the reference's own models   bayesnn_fpga_b200.lowering (traced with torch.fx)
    bayes_hw/models/t_qmodels_bayes_me.py (LeNet spec) bayesnn_fpga_b200.lenet

All arithmetic runs in the sm_100a kernels behind include/bnn_b200.h; there is no CPU fallback.
"""
from . import _lib                                   # noqa: F401
from .engine import Engine, Graph, MCResult          # noqa: F401
from .predict import masksembles_batched_predict, mc_predict   # noqa: F401

__all__ = ["mc_predict", "masksembles_batched_predict", "Engine", "Graph", "MCResult"]
