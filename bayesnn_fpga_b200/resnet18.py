"""Multi-exit ResNet-18 for 32x32 inputs - drop-in for the reference's
``Software_Artifact/software/models/resnet18/resnet18.py`` (``BasicBlock`` :17-48, ``ResNet`` :88-180,
``ResNet18EarlyExit`` :182-186, ``ResNet18Base`` :189-204, ``MCDropout`` :207-210, ``ResNet18MC``
:212-258, ``ResNet18MCEarlyExit`` :260-346).

Same class names, constructor kwargs, parameter names (so reference ``state_dict``s load) and
return structure (list of logits tensors, shallow exit first).  The modules only HOLD parameters:
``forward`` lowers the network to an op graph (:meth:`_bnn_graph`) and runs it on the B200 through
``bayesnn_fpga_b200.engine`` - one stochastic pass per call, like ``model(b_x)`` in the reference.
For S passes use :func:`bayesnn_fpga_b200.mc_predict`, which computes the deterministic prefix once.
"""
import math

import torch
import torch.nn as nn

from . import _plans
from . import engine as _engine
from .Dropouts import MCDropout
from .utils import Masksembles1D, Masksembles2D


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    """conv3x3-BN-ReLU-conv3x3-BN (+ 1x1/BN shortcut) - add - ReLU. Parameter holder; lowered by
    :meth:`lower`."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.planes = planes
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=False)
        self.downsample = downsample
        self.stride = stride

    def lower(self, g, x, name):
        out = g.conv(x, self.conv1, self.bn1, relu=True, name=name + ".conv1")
        res = x
        if self.downsample is not None:
            res = g.conv(x, self.downsample[0], self.downsample[1], relu=False, name=name + ".downsample")
        return g.conv(out, self.conv2, self.bn2, relu=True, residual=res, name=name + ".conv2")


class _BnnModel(nn.Module):
    """Shared forward machinery of the multi-exit models."""
    bnn_dtype = "fp16"
    bnn_seed = 0x5EED

    def _bnn_graph(self):
        raise NotImplementedError

    def bnn_engine(self, dtype=None, rebuild=False, **kw):
        """Plan (fold BN, pack weights, fuse sites) for the current parameters; cached per dtype in ``_plans`` (outside
        the module, so ``torch.save(model)`` / ``copy.deepcopy(model)`` keep working after a run) and rebuilt whenever
        the parameters, buffers or sub-modules changed since it was made."""
        dtype = dtype or self.bnn_dtype
        if rebuild:
            _plans.drop(self)
        cache = _plans.plans_for(self)
        key = (dtype, tuple(sorted(kw.items())))
        if key not in cache:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("bayesnn_fpga_b200 models run on a CUDA (sm_100) device only; there is no "
                                   "CPU fallback - call model.cuda() first")
            cache[key] = _engine.Engine(self._bnn_graph(), dtype=dtype, device=dev, **kw)
        return cache[key]

    def _masksembles_modules(self):
        from .utils import _MasksemblesBase
        return [m for m in self.modules() if isinstance(m, _MasksemblesBase)]

    def bnn_group_engine(self, dtype=None):
        """Plan of the Masksembles BATCHED formulation (training branch of utils.py:158-164, :220-226): the batch is n
        contiguous groups, group g runs through mask g of every Masksembles layer; nothing is shared between groups."""
        dtype = dtype or self.bnn_dtype
        cache = _plans.plans_for(self)
        key = ("groups", dtype)
        if key not in cache:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("bayesnn_fpga_b200 models run on a CUDA (sm_100) device only; there is no "
                                   "CPU fallback - call model.cuda() first")
            cache[key] = _engine.Engine(self._bnn_graph().mark_input_stochastic(), dtype=dtype, device=dev, mask_gather=0)
        return cache[key]

    def _forward_groups(self, x, mods):
        """One forward in which every Masksembles module is in TRAINING mode: image group g uses mask row g
        (no rotation of `cnt`, no rescale) - the reference modules' `if self.training` branch - through the fused plan."""
        n = int(mods[0].n)
        if any(int(m.n) != n for m in mods):
            raise ValueError("Masksembles layers with different n in one network")
        batch = x.shape[0]
        if batch % n != 0:
            raise ValueError('Batch size must be divisible by n, got batch {} and n {}'.format(batch, n))
        eng = self.bnn_group_engine()
        saved = [int(m.cnt) for m in mods]
        for m in mods:
            m.cnt = 0                                   # rows 0 .. n-1 in order, whatever the eval-mode counter says
        try:
            r = eng.run(x, n, seed=self.bnn_seed, want_logits=True, mask_offset=0)
            # all_logits [n groups, E, batch / n, C] -> per exit [batch, C] in the original image order
            outs = [r.all_logits[:, e].reshape(batch, -1).clone() for e in range(r.all_logits.shape[1])]
        finally:
            for m, c in zip(mods, saved):
                m.cnt = c
        return outs

    def forward(self, x):
        """One stochastic forward pass -> list of E logits tensors [B, out_dim]."""
        mods = self._masksembles_modules()
        if mods and all(m.training for m in mods):
            outs = self._forward_groups(x, mods)
            self.intermediary_output_list = (outs[-1], outs[:-1], None, [])
            return outs
        eng = self.bnn_engine()
        sample = self.__dict__.get("_bnn_pass", 0)
        self.__dict__["_bnn_pass"] = sample + 1
        r = eng.run(x, 1, seed=self.bnn_seed, sample0=sample, want_logits=True, mask_offset=0)
        outs = [r.all_logits[0, e].clone() for e in range(r.all_logits.shape[1])]
        self.intermediary_output_list = (outs[-1], outs[:-1], None, [])
        return outs


def _make_site_module(mask_type, p, channels, num_masks, mask_scale, two_d):
    if mask_type == "mc":
        return MCDropout(p)
    return (Masksembles2D if two_d else Masksembles1D)(channels, num_masks, mask_scale)


def _lower_site(g, t, mod, name):
    if isinstance(mod, MCDropout):
        return g.site(t, "mc", mod.p, name=name)
    return g.site(t, "mask", 0.0, module=mod, name=name)


def _head_site(g, mod, name):
    if mod is None:
        return None
    if isinstance(mod, MCDropout):
        return g.new_site("mc", mod.p, name=name)
    return g.new_site("mask", 0.0, module=mod, name=name)


class ResNet(_BnnModel):
    def __init__(self, block=BasicBlock, num_blocks=[2, 2, 2, 2], num_classes=100):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=False)
        self.layer1 = self._make_layer(block, 64, num_blocks[0], stride=1)
        self.layer2 = self._make_layer(block, 128, num_blocks[1], stride=2)
        self.layer3 = self._make_layer(block, 256, num_blocks[2], stride=2)
        self.layer4 = self._make_layer(block, 512, num_blocks[3], stride=2)
        self.linear = nn.Linear(512 * block.expansion, num_classes)

        widths = {1: [64, 128, 256, 512], 2: [128, 256, 512], 3: [256, 512]}
        for e, ch in widths.items():                     # exit branches: stride-2 conv3x3 + BN chains
            for j in range(len(ch) - 1):
                setattr(self, "ex%dconv%d" % (e, j + 1),
                        nn.Conv2d(ch[j], ch[j + 1], kernel_size=3, stride=2, padding=1, bias=False))
            for j in range(len(ch) - 1):
                setattr(self, "ex%dbn%d" % (e, j + 1), nn.BatchNorm2d(ch[j + 1]))
            setattr(self, "ex%dlinear" % e, nn.Linear(512, num_classes))

        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    # ---- lowering ---------------------------------------------------------------------------
    _early_exit = True

    def _lower_stage(self, g, t, stage):
        """layerN may be: Sequential(blocks), Sequential(Sequential(blocks), site) [block mode] or a
        Sequential whose entries are Sequential(block, site) [layer mode] (resnet18.py:273-290)."""
        name = "layer%d" % stage

        def walk(mod, t, path):
            if isinstance(mod, BasicBlock):
                return mod.lower(g, t, path)
            if isinstance(mod, nn.Sequential):
                for k, child in mod.named_children():
                    t = walk(child, t, path + "." + k)
                return t
            return _lower_site(g, t, mod, path)
        return walk(getattr(self, name), t, name)

    def _bnn_graph(self):
        size = getattr(self, "image_size", 32)
        g = _engine.Graph(3, size, size)
        t = g.conv(g.input, self.conv1, self.bn1, relu=False, name="conv1")        # :303, no ReLU
        n_branch = {1: 3, 2: 2, 3: 1}
        for stage in (1, 2, 3, 4):
            t = self._lower_stage(g, t, stage)
            if stage == 4 or not self._early_exit:
                continue
            o = t                                                                     # exit branch :306-314
            for j in range(1, n_branch[stage] + 1):
                o = g.conv(o, getattr(self, "ex%dconv%d" % (stage, j)), getattr(self, "ex%dbn%d" % (stage, j)),
                           relu=True, name="ex%dconv%d" % (stage, j))
            site = _head_site(g, getattr(self, "exit%d_dropout" % stage, None), "exit%d_dropout" % stage)
            g.head(o, getattr(self, "ex%dlinear" % stage), site, name="ex%dlinear" % stage)
        site = _head_site(g, getattr(self, "exit_dropout", None), "exit_dropout")
        g.head(t, self.linear, site, name="linear")                                  # :339-344
        return g


class ResNet18EarlyExit(ResNet):
    def __init__(self, n_exits=4, out_dim=100, image_size=32, *args, **kwargs):
        super().__init__(block=BasicBlock, num_blocks=[2, 2, 2, 2], num_classes=out_dim, *args, **kwargs)
        self.n_exits = n_exits
        self.out_dim = out_dim
        self.image_size = image_size


class ResNet18Base(ResNet):
    _early_exit = False

    def __init__(self, n_exits=1, out_dim=100, *args, **kwargs):
        super().__init__(block=BasicBlock, num_blocks=[2, 2, 2, 2], num_classes=out_dim, *args, **kwargs)
        self.n_exits = n_exits
        self.out_dim = out_dim


class _ResNet18MCMixin:
    def _install_sites(self, multi_exit):
        stages = [self.layer1, self.layer2, self.layer3, self.layer4]
        mk = lambda ch, two_d: _make_site_module(self.mask_type, self.dropout_p, ch, self.num_masks,
                                                 self.mask_scale, two_d)
        if self.dropout == "block":                      # after layer1..3, not layer4 (:273-280)
            for i in range(3):
                stages[i] = nn.Sequential(stages[i], mk(stages[i][-1].planes, True))
            self.layer1, self.layer2, self.layer3, self.layer4 = stages
        elif self.dropout == "layer":                    # after every BasicBlock but the last (:281-288)
            if self.mask_type != "mc":
                # the reference raises UnboundLocalError here (resnet18.py:288 indexes an unbound `i`)
                raise NotImplementedError('dropout="layer" with Masksembles is not constructible in the reference')
            for b, stage in enumerate(stages):
                for l in range(len(stage)):
                    if not (b == 3 and l == len(stage) - 1):
                        stage[l] = nn.Sequential(stage[l], MCDropout(self.dropout_p))
        if self.dropout_exit:
            names = ["exit1_dropout", "exit2_dropout", "exit3_dropout", "exit_dropout"] if multi_exit else ["exit_dropout"]
            for nme in names:
                setattr(self, nme, mk(512 * BasicBlock.expansion, False))


class ResNet18MC(ResNet, _ResNet18MCMixin):
    _early_exit = False

    def __init__(self, dropout_exit=False, dropout=None, dropout_p=0.5, n_exits=1, out_dim=100, image_size=32,
                 mask_type="mc", num_masks=4, mask_scale=4.0, *args, **kwargs):
        super().__init__(block=BasicBlock, num_blocks=[2, 2, 2, 2], num_classes=out_dim, *args, **kwargs)
        self.n_exits, self.out_dim, self.image_size = n_exits, out_dim, image_size
        self.dropout_exit, self.dropout, self.dropout_p = dropout_exit, dropout, dropout_p
        self.mask_type, self.num_masks, self.mask_scale = mask_type, num_masks, mask_scale
        self._install_sites(multi_exit=False)


class ResNet18MCEarlyExit(ResNet, _ResNet18MCMixin):
    def __init__(self, dropout_exit=False, dropout=None, dropout_p=0.5, n_exits=4, out_dim=100, image_size=32,
                 mask_type="mc", num_masks=4, mask_scale=4.0, *args, **kwargs):
        super().__init__(block=BasicBlock, num_blocks=[2, 2, 2, 2], num_classes=out_dim, *args, **kwargs)
        self.n_exits, self.out_dim, self.image_size = n_exits, out_dim, image_size
        self.dropout_exit, self.dropout, self.dropout_p = dropout_exit, dropout, dropout_p
        self.mask_type, self.num_masks, self.mask_scale = mask_type, num_masks, mask_scale
        self._install_sites(multi_exit=True)


_MC_ONLY_KWARGS = ("dropout", "dropout_exit", "dropout_p", "mask_type", "num_masks", "mask_scale")


def get_res_net_18(network_type, hyperparams):
    """resnet18_loader.py:4-15.  Like the reference: ``None`` and "early_exit" both build ``ResNet18EarlyExit`` and
    drop the six MC-only keys of a full hyper-parameter dict; "mc" / "mc_early_exit" keep them."""
    from .utils import dict_drop
    kw = dict_drop(hyperparams, "call", "load_model", "resnet_type")
    if network_type is None or network_type == "early_exit":
        return ResNet18EarlyExit(**dict_drop(kw, *_MC_ONLY_KWARGS))
    table = {"mc": ResNet18MC, "mc_early_exit": ResNet18MCEarlyExit}
    if network_type not in table:
        raise ValueError("unknown resnet type %r" % (network_type,))
    return table[network_type](**kw)
