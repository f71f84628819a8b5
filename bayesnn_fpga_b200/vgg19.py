"""Multi-exit VGG-19-BN - drop-in for the reference's
``Software_Artifact/software/models/vgg19/vgg19.py`` (``VGG`` :88-119, ``make_layers`` :121-143,
``make_classifier`` :146-183, ``VGG19`` :186-192, ``VGG19MC`` :194-254, ``VGG19EarlyExit`` :256-324,
``VGG19MCEarlyExit`` :327-382, ``MCDropout`` :384-387, ``get_vgg_19`` :16-42) for 32x32 inputs.

Same names / kwargs / parameter names / return structure; ``forward`` lowers to the B200 op graph.
The 224x224 ChestX path with ImageNet weight remapping (vgg19.py:44-84) is dataset-specific and not
part of this path.  Reference defects kept visible instead of replicated: the
reference's ``dropout="block"`` / ``"layer"`` VGG constructors raise AttributeError
(vgg19.py:224,235,365,376); here ``dropout="block"`` works and follows the reference's intended
placement (end of every block except the last two for the multi-exit net, except the last for the
single-exit net), using 2-D Masksembles where the reference (wrongly) names ``Masksembles1D``.
BASELINE config 4 ("dropout on the last 3 blocks") is built with :meth:`append_block_dropout`.
"""
import math

import torch.nn as nn

from . import engine as _engine
from .Dropouts import MCDropout
from .resnet18 import _BnnModel, _head_site, _lower_site
from .utils import Masksembles1D, Masksembles2D, dict_drop

CFG19 = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M']


def make_layers(cfg, batch_norm=False):
    """vgg19.py:121-143: returns (ModuleList of per-block Sequentials, ModuleList of per-block ModuleLists);
    both views share the same layer objects."""
    blocks, layers, cin = nn.ModuleList(), nn.ModuleList(), 3
    for l in cfg:
        if l == 'M':
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            blocks.append(layers)
            layers = nn.ModuleList()
            continue
        layers.append(nn.Conv2d(cin, l, kernel_size=3, padding=1))
        if batch_norm:
            layers.append(nn.BatchNorm2d(l))
        layers.append(nn.ReLU(inplace=True))
        cin = l
    return nn.ModuleList(nn.Sequential(*b) for b in blocks), blocks


def make_classifier(size, num_classes, mc_dropout_p=0, mask_type='mask', num_masks=4, mask_scale=4.0):
    """vgg19.py:146-183, 32x32 branch: [MCDropout | Masksembles1D,] Linear(512, C)."""
    if size == 224:
        raise NotImplementedError("the 224x224 ChestX classifier is outside the 32x32 inference path")
    if mc_dropout_p == 0:
        return nn.Sequential(nn.Linear(512, num_classes))
    if mask_type == 'mc':
        return nn.Sequential(MCDropout(p=mc_dropout_p), nn.Linear(512, num_classes))
    return nn.Sequential(Masksembles1D(512, num_masks, mask_scale), nn.Linear(512, num_classes))


class VGG(_BnnModel):
    def __init__(self, blocks, num_class=100, image_size=32):
        super().__init__()
        self.blocks, self.non_sequentialized_blocks = blocks
        self.image_size = image_size
        self.avg_pool = nn.AdaptiveAvgPool2d((7, 7))
        self.classifier = make_classifier(self.image_size, num_class, mask_type=None)
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.fill_(0.01)

    def append_block_dropout(self, block_indices, module_factory=None):
        """Append a stochastic layer at the end (after the max-pool) of the given blocks, e.g.
        ``m.append_block_dropout((2, 3, 4))`` == the reference idiom
        ``for i in (2,3,4): m.blocks[i].append(MCDropout(p))`` used for BASELINE config 4."""
        for i in block_indices:
            self.blocks[i].append(module_factory() if module_factory else MCDropout(self.dropout_p))
        return self                              # (plans re-validate themselves: _plans.fingerprint sees the new module)

    # ---- lowering ---------------------------------------------------------------------------
    def _lower_block(self, g, t, b):
        seq = list(self.blocks[b].named_children())
        i = 0
        while i < len(seq):
            k, m = seq[i]
            name = "blocks.%d.%s" % (b, k)
            if isinstance(m, nn.Conv2d):
                bn = seq[i + 1][1] if i + 1 < len(seq) and isinstance(seq[i + 1][1], nn.BatchNorm2d) else None
                j = i + (2 if bn is not None else 1)
                relu = j < len(seq) and isinstance(seq[j][1], nn.ReLU)
                t = g.conv(t, m, bn, relu=relu, name=name)
                i = j + (1 if relu else 0)
            elif isinstance(m, nn.MaxPool2d):
                t = g.maxpool(t, m.kernel_size, name=name)
                i += 1
            elif isinstance(m, (MCDropout, Masksembles1D, Masksembles2D)):
                t = _lower_site(g, t, m, name)
                i += 1
            else:
                raise NotImplementedError("cannot lower %s (%s)" % (name, type(m).__name__))
        return t

    @staticmethod
    def _lower_classifier(g, t, seq, name):
        mods = list(seq.children())
        site = _head_site(g, mods[0], name + ".0") if len(mods) == 2 else None
        g.head(t, mods[-1], site, name=name)

    _early_exit = False

    def _bnn_graph(self):
        if self.image_size != 32:
            raise NotImplementedError("only the 32x32 path is lowered")
        g = _engine.Graph(3, 32, 32)
        t = g.input
        n_branch = {0: 3, 1: 2, 2: 1}
        for b in range(5):
            t = self._lower_block(g, t, b)
            if not self._early_exit or b == 4:
                continue
            if b in n_branch:                                   # vgg19.py:295-312
                fe = getattr(self, "ex%dfeatureextractor" % (b + 1))
                o = t
                for j in range(n_branch[b]):
                    o = g.conv(o, fe[3 * j], fe[3 * j + 1], relu=True, name="ex%dfeatureextractor.%d" % (b + 1, 3 * j))
                self._lower_classifier(g, o, getattr(self, "ex%dlinear" % (b + 1)), "ex%dlinear" % (b + 1))
            else:                                               # block 3: avg_pool2d(out, 2) -> ex4linear
                self._lower_classifier(g, t, self.ex4linear, "ex4linear")
        self._lower_classifier(g, t, self.classifier, "classifier")
        return g


class VGG19(VGG):
    def __init__(self, n_exits=1, out_dim=100, *args, **kwargs):
        super().__init__(make_layers(CFG19, batch_norm=True), num_class=out_dim, *args, **kwargs)
        self.n_exits = n_exits
        self.out_dim = out_dim
        self.init_weights()


def _install_block_sites(model, skip_last):
    """dropout="block": a site at the end of every block except the last `skip_last`."""
    nb = len(model.blocks)
    for b in range(nb - skip_last):
        ch = [m for m in model.blocks[b] if isinstance(m, nn.Conv2d)][-1].out_channels
        mod = MCDropout(model.dropout_p) if model.mask_type == "mc" else Masksembles2D(ch, model.num_masks,
                                                                                      model.mask_scale)
        model.blocks[b].append(mod)


class VGG19MC(VGG19):
    def __init__(self, dropout_exit=False, dropout=None, dropout_p=0.5, n_exits=1, out_dim=100, mask_type="mc",
                 num_masks=4, mask_scale=4.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.n_exits, self.out_dim = n_exits, out_dim
        self.dropout, self.dropout_p, self.dropout_exit = dropout, dropout_p, dropout_exit
        self.mask_type, self.num_masks, self.mask_scale = mask_type, num_masks, mask_scale
        if self.dropout_exit:
            self.classifier = make_classifier(self.image_size, self.out_dim, self.dropout_p, self.mask_type,
                                              self.num_masks, self.mask_scale)
        self.init_weights()
        if self.dropout == "block":
            _install_block_sites(self, skip_last=1)
        elif self.dropout is not None:
            raise NotImplementedError('VGG dropout=%r is not constructible in the reference either' % (self.dropout,))


class VGG19EarlyExit(VGG19):
    _early_exit = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        widths = {1: [64, 128, 256, 512], 2: [128, 256, 512], 3: [256, 512]}
        for e, ch in widths.items():
            mods = []
            for j in range(len(ch) - 1):
                mods += [nn.Conv2d(ch[j], ch[j + 1], kernel_size=3, stride=2, padding=1, bias=False),
                         nn.BatchNorm2d(ch[j + 1]), nn.ReLU(inplace=True)]
            setattr(self, "ex%dfeatureextractor" % e, nn.Sequential(*mods))
            setattr(self, "ex%dlinear" % e, make_classifier(self.image_size, self.out_dim, mask_type=None))
        self.ex4linear = make_classifier(self.image_size, self.out_dim, mask_type=None)
        self.init_weights()


class VGG19MCEarlyExit(VGG19EarlyExit):
    def __init__(self, dropout_exit=False, dropout=None, dropout_p=0.5, n_exits=4, out_dim=100, mask_type="mc",
                 num_masks=4, mask_scale=4.0, *args, **kwargs):
        super().__init__(n_exits=n_exits, out_dim=out_dim, *args, **kwargs)
        self.n_exits, self.out_dim = n_exits, out_dim
        self.dropout, self.dropout_p, self.dropout_exit = dropout, dropout_p, dropout_exit
        self.mask_type, self.num_masks, self.mask_scale = mask_type, num_masks, mask_scale
        if self.dropout_exit:
            mk = lambda: make_classifier(self.image_size, self.out_dim, self.dropout_p, self.mask_type,
                                         self.num_masks, self.mask_scale)
            self.ex1linear, self.ex2linear, self.ex3linear, self.ex4linear = mk(), mk(), mk(), mk()
            self.classifier = mk()
        self.init_weights()
        if self.dropout == "block":
            _install_block_sites(self, skip_last=2)
        elif self.dropout is not None:
            raise NotImplementedError('VGG dropout=%r is not constructible in the reference either' % (self.dropout,))


_MC_ONLY_KWARGS = ("dropout", "dropout_exit", "dropout_p", "mask_type", "num_masks", "mask_scale")


def get_vgg_19(network_type, hyperparams):
    """vgg19.py:16-42.  The plain and "early_exit" classes are built without the six MC-only keys of a full
    hyper-parameter dict, like the reference."""
    kw = dict_drop(hyperparams, "call", "load_model", "resnet_type")
    table = {None: VGG19, "early_exit": VGG19EarlyExit, "mc": VGG19MC, "mc_early_exit": VGG19MCEarlyExit}
    if network_type not in table:
        raise ValueError("unknown vgg type %r" % (network_type,))
    if network_type in (None, "early_exit"):
        kw = dict_drop(kw, *_MC_ONLY_KWARGS)
    return table[network_type](**kw)
