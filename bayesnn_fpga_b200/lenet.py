"""Multi-exit LeNet-5 with last-layer MC dropout - the PyTorch-facing form of the reference's Keras
spec ``Hardware_Artifact/bayes_hw/models/t_qmodels_bayes_me.py:41-147`` (float layers as in
``models.py:34-73``; dropout kind ``model_utils.py:43-47``).  The reference has no PyTorch LeNet; layer
names follow the Keras layer names so the mapping is one to one.

"Temporal mapping" (t_qmodels_bayes_me.py:84-94,130-139: the graph holds the prefix once and S copies
of dropout -> FC -> softmax on a cached tensor) is exactly what the engine's prefix/suffix split does.
Returns [main exit, second exit] like the Keras model (:141).
"""
import torch.nn as nn

from . import engine as _engine
from .Dropouts import MCDropout
from .resnet18 import _BnnModel, _head_site
from .utils import Masksembles1D


class LeNetMCEarlyExit(_BnnModel):
    def __init__(self, dropout_p=0.2, out_dim=10, mask_type="mc", num_masks=4, mask_scale=2.0, n_exits=2):
        super().__init__()
        self.n_exits, self.out_dim, self.dropout_p, self.mask_type = n_exits, out_dim, dropout_p, mask_type
        self.dropout, self.dropout_exit = None, True
        self.conv2d_1 = nn.Conv2d(1, 20, 5, padding=2)                    # :49-52, "same"
        self.conv2d_2 = nn.Conv2d(20, 20, 5, padding=2)                   # :99-102
        self.conv2d_2_2nd_exit = nn.Conv2d(20, 20, 5, stride=7, padding=0)  # :58-61 ("same" at stride 7 -> no pad)
        self.fc_1 = nn.Linear(80, 100)                                    # :110
        self.fc_1_2nd_exit = nn.Linear(80, 100)                           # :66
        self.fc_exit_1st = nn.Linear(100, out_dim)                        # :118-119
        self.fc_2nd_exit = nn.Linear(100, out_dim)                        # :70-71
        mk = (lambda: MCDropout(dropout_p)) if mask_type == "mc" else (lambda: Masksembles1D(100, num_masks, mask_scale))
        self.bayes_2nd_exit = mk()                                        # :69
        self.bayes_1st_exit = mk()                                        # :117

    def _bnn_graph(self):
        g = _engine.Graph(1, 28, 28)
        t = g.conv(g.input, self.conv2d_1, None, relu=True, name="conv2d_1")
        t = g.maxpool(t, 2, name="pool1")
        e = g.conv(t, self.conv2d_2_2nd_exit, None, relu=True, name="conv2d_2_2nd_exit")
        e = g.linear(e, self.fc_1_2nd_exit, relu=True, name="fc_1_2nd_exit")
        site_e = _head_site(g, self.bayes_2nd_exit, "bayes_2nd_exit")     # visited first, like the spec
        m = g.conv(t, self.conv2d_2, None, relu=True, name="conv2d_2")
        m = g.maxpool(m, 7, name="pool2")
        m = g.linear(m, self.fc_1, relu=True, name="fc_1")
        site_m = _head_site(g, self.bayes_1st_exit, "bayes_1st_exit")
        g.head(m, self.fc_exit_1st, site_m, name="fc_exit_1st")           # output 0 = main exit
        g.head(e, self.fc_2nd_exit, site_e, name="fc_2nd_exit")           # output 1 = second exit
        return g
