"""Functional fp32 torch-CPU restatement of the reference networks' forward passes
(test infrastructure, see oracle/__init__.py).

Every function takes a plain ``state_dict`` (name -> tensor, the reference's own parameter
names) and a ``site`` callback that implements the stochastic layers, so the same code runs
with injected Philox masks (parity) or with torch's own RNG (CPU baseline timing).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import philox


# --------------------------------------------------------------------------------------
# stochastic sites
# --------------------------------------------------------------------------------------
class SiteSpec:
    """Describes one network's stochastic layers.

    kind "mc"   : element-wise dropout, ``F.dropout(x, p, training=True)`` (resnet18.py:207-210)
    kind "mc2d" : channel-wise dropout, ``F.dropout2d`` (converter Dropouts.py:43-45)
    kind "mask" : Masksembles eval branch, ``x * masks[cnt]`` with rotation (utils.py:165-169)
    """

    def __init__(self, kind="mc", p=0.5, masks=None, cnt0=0):
        self.kind = kind
        self.p = float(p)
        self.masks = masks or {}       # site name -> float tensor [n, C]
        self.cnt0 = int(cnt0)


class InjectedSites:
    """site(name, x) with masks from the Philox contract; stream id = order of first use
    within one forward pass (the product's plan builder numbers its sites the same way)."""

    def __init__(self, spec, seed, sample, batch_offset=0):
        self.spec, self.seed, self.sample = spec, int(seed), int(sample)
        self.batch_offset = int(batch_offset)      # x holds images batch_offset.. of the batch the masks are drawn for
        self.streams = {}

    def __call__(self, name, x):
        sp = self.spec
        if name not in self.streams:
            self.streams[name] = len(self.streams)
        stream = self.streams[name]
        if sp.kind == "mask":
            m = sp.masks[name]
            row = m[(sp.cnt0 + self.sample) % m.shape[0]].to(x.dtype)
            return x * row.reshape(1, -1, *([1] * (x.dim() - 2)))
        mode = "channel" if sp.kind == "mc2d" else "element"
        keep = philox.keep_mask(self.seed, stream, self.sample, x.shape, sp.p, mode, self.batch_offset)
        if sp.p >= 1.0:
            return torch.zeros_like(x)
        scale = np.float32(1.0) / np.float32(1.0 - sp.p)
        return x * (torch.from_numpy(keep).to(x.dtype) * float(scale))


class GroupSites:
    """Masksembles TRAINING-branch / batched formulation: the batch is split into n contiguous groups and group g is
    multiplied by mask row g (Software_Artifact/software/utils.py:158-164, :220-226; the Keras layer does the same
    to an input tiled n times, Hardware_Artifact/converter/keras/Masksembles.py:177-181).  No rotation, no rescale."""

    def __init__(self, spec):
        self.spec = spec

    def __call__(self, name, x):
        m = self.spec.masks[name]
        n, batch = m.shape[0], x.shape[0]
        if batch % n != 0:
            raise ValueError('Batch size must be divisible by n, got batch {} and n {}'.format(batch, n))
        rows = m[torch.arange(batch) // (batch // n)].to(x.dtype)                  # [batch, C]
        return x * rows.reshape(batch, -1, *([1] * (x.dim() - 2)))


class TorchRngSites:
    """site(name, x) exactly as the reference executes it (torch global RNG; stateful
    Masksembles counter).  Used for the CPU baseline timing leg."""

    def __init__(self, spec):
        self.spec = spec
        self.cnt = spec.cnt0

    def next_pass(self):
        self.cnt += 1

    def __call__(self, name, x):
        sp = self.spec
        if sp.kind == "mask":
            m = sp.masks[name]
            row = m[self.cnt % m.shape[0]]
            return (x * row.reshape(1, -1, *([1] * (x.dim() - 2)))).float()
        if sp.kind == "mc2d":
            return F.dropout2d(x, sp.p, True)
        return F.dropout(x, sp.p, True)


def _bn(sd, prefix, x):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, 1e-5)


# --------------------------------------------------------------------------------------
# ResNet-18, multi-exit  (resnet18.py)
# --------------------------------------------------------------------------------------
def _basic_block(sd, pfx, x, stride):
    # resnet18.py:32-48
    out = F.conv2d(x, sd[pfx + ".conv1.weight"], None, stride, 1)
    out = F.relu(_bn(sd, pfx + ".bn1", out))
    out = F.conv2d(out, sd[pfx + ".conv2.weight"], None, 1, 1)
    out = _bn(sd, pfx + ".bn2", out)
    res = x
    if (pfx + ".downsample.0.weight") in sd:
        res = F.conv2d(x, sd[pfx + ".downsample.0.weight"], None, stride, 0)
        res = _bn(sd, pfx + ".downsample.1", res)
    return F.relu(out + res)


def _resnet_stage(sd, stage, x, dropout, site):
    """layerN of resnet18.py:278-290. Module paths change with the dropout mode:
    block -> Sequential(layerN, drop): blocks at 'layerN.0.i', site 'layerN.1';
    layer -> layerN[i] = Sequential(block, drop): blocks at 'layerN.i.0', site 'layerN.i.1'."""
    name = "layer%d" % stage
    stride = 1 if stage == 1 else 2
    for i in range(2):
        if dropout == "block" and stage != 4:
            pfx = "%s.0.%d" % (name, i)
        elif dropout == "layer" and not (stage == 4 and i == 1):
            pfx = "%s.%d.0" % (name, i)
        else:
            pfx = "%s.%d" % (name, i)
        x = _basic_block(sd, pfx, x, stride if i == 0 else 1)
        if dropout == "layer" and not (stage == 4 and i == 1):
            x = site("%s.%d.1" % (name, i), x)
    if dropout == "block" and stage != 4:
        x = site(name + ".1", x)
    return x


def resnet18_forward(sd, x, site, dropout=None, dropout_exit=False, early_exit=True):
    """ResNet18MCEarlyExit.forward (resnet18.py:302-346) / ResNet18MC.forward (:245-258)."""
    out = _bn(sd, "bn1", F.conv2d(x, sd["conv1.weight"], None, 1, 1))   # :303 (no ReLU)
    out = _resnet_stage(sd, 1, out, dropout, site)
    outs = []
    exit_convs = {1: 3, 2: 2, 3: 1}
    for stage in (1, 2, 3):
        if early_exit:
            o = out
            for j in range(1, exit_convs[stage] + 1):                   # :306-308
                o = F.conv2d(F.relu(o), sd["ex%dconv%d.weight" % (stage, j)], None, 2, 1)
                o = _bn(sd, "ex%dbn%d" % (stage, j), o)
            o = F.avg_pool2d(F.relu(o), 4).flatten(1)                   # :309-311
            if dropout_exit:
                o = site("exit%d_dropout" % stage, o)                   # :312-313
            outs.append(F.linear(o, sd["ex%dlinear.weight" % stage], sd["ex%dlinear.bias" % stage]))
        out = _resnet_stage(sd, stage + 1, out, dropout, site)          # :316,:327,:337
    o = F.avg_pool2d(F.relu(out), 4).flatten(1)                         # :339-341
    if dropout_exit:
        o = site("exit_dropout", o)
    outs.append(F.linear(o, sd["linear.weight"], sd["linear.bias"]))
    return outs


# --------------------------------------------------------------------------------------
# VGG-19-BN, multi-exit  (vgg19.py)
# --------------------------------------------------------------------------------------
VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M",
             512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]          # vgg19.py:188


def _vgg_block(sd, b, x, nconv):
    # make_layers vgg19.py:121-143: [conv(bias), bn, relu] * nconv + maxpool
    for j in range(nconv):
        k = "blocks.%d.%d" % (b, 3 * j)
        x = F.conv2d(x, sd[k + ".weight"], sd[k + ".bias"], 1, 1)
        x = F.relu(_bn(sd, "blocks.%d.%d" % (b, 3 * j + 1), x))
    return F.max_pool2d(x, 2, 2)


def _vgg_head(sd, name, x, site, dropout_exit):
    # make_classifier vgg19.py:173-183: [dropout,] Linear(512, C)
    if dropout_exit:
        x = site(name + ".0", x)
        return F.linear(x, sd[name + ".1.weight"], sd[name + ".1.bias"])
    return F.linear(x, sd[name + ".0.weight"], sd[name + ".0.bias"])


def vgg19_forward(sd, x, site, dropout_exit=False, dropout_blocks=()):
    """VGG19EarlyExit.forward (vgg19.py:294-324) with MCDropout / Masksembles appended at the
    END of the blocks in ``dropout_blocks`` (after the max-pool) - the construction SURVEY.md
    section 8c uses for BASELINE config 4 because the reference's own ``dropout="block"`` VGG
    constructor raises (vgg19.py:365)."""
    nconv = [2, 2, 4, 4, 4]
    exit_convs = {0: 3, 1: 2, 2: 1}
    outs = []
    out = x
    for b in range(5):
        out = _vgg_block(sd, b, out, nconv[b])
        if b in dropout_blocks:
            out = site("blocks.%d.%d" % (b, 3 * nconv[b] + 1), out)
        if b in exit_convs:                                             # :295-312
            o = F.relu(out)
            for j in range(exit_convs[b]):
                fe = "ex%dfeatureextractor" % (b + 1)
                o = F.conv2d(o, sd["%s.%d.weight" % (fe, 3 * j)], None, 2, 1)
                o = F.relu(_bn(sd, "%s.%d" % (fe, 3 * j + 1), o))
            o = F.avg_pool2d(o, 2).flatten(1)
            outs.append(_vgg_head(sd, "ex%dlinear" % (b + 1), o, site, dropout_exit))
        elif b == 3:                                                    # :314-317
            o = F.avg_pool2d(out, 2).flatten(1)
            outs.append(_vgg_head(sd, "ex4linear", o, site, dropout_exit))
    outs.append(_vgg_head(sd, "classifier", out.flatten(1), site, dropout_exit))   # :319-322
    return outs


# --------------------------------------------------------------------------------------
# multi-exit LeNet (float restatement of t_qmodels_bayes_me.py:41-147)
# --------------------------------------------------------------------------------------
def lenet_forward(sd, x, site):
    """Returns [main logits, exit-2 logits] (Keras output order, t_qmodels_bayes_me.py:141)."""
    t = F.relu(F.conv2d(x, sd["conv2d_1.weight"], sd["conv2d_1.bias"], 1, 2))        # :49-52
    t = F.max_pool2d(t, 2, 2)                                                         # :54
    # second exit :58-71 ("same" padding with stride 7 on 14x14, k=5 -> zero padding)
    e = F.relu(F.conv2d(t, sd["conv2d_2_2nd_exit.weight"], sd["conv2d_2_2nd_exit.bias"], 7, 0))
    e = F.relu(F.linear(e.flatten(1), sd["fc_1_2nd_exit.weight"], sd["fc_1_2nd_exit.bias"]))
    e = site("bayes_2nd_exit", e)                                                     # :84
    e = F.linear(e, sd["fc_2nd_exit.weight"], sd["fc_2nd_exit.bias"])                 # :85
    # main exit :99-117
    m = F.relu(F.conv2d(t, sd["conv2d_2.weight"], sd["conv2d_2.bias"], 1, 2))
    m = F.max_pool2d(m, 7, 7)
    m = F.relu(F.linear(m.flatten(1), sd["fc_1.weight"], sd["fc_1.bias"]))
    m = site("bayes_1st_exit", m)                                                     # :130
    m = F.linear(m, sd["fc_exit_1st.weight"], sd["fc_exit_1st.bias"])                 # :131
    return [m, e]


LENET_SHAPES = {
    "conv2d_1.weight": (20, 1, 5, 5), "conv2d_1.bias": (20,),
    "conv2d_2.weight": (20, 20, 5, 5), "conv2d_2.bias": (20,),
    "conv2d_2_2nd_exit.weight": (20, 20, 5, 5), "conv2d_2_2nd_exit.bias": (20,),
    "fc_1.weight": (100, 80), "fc_1.bias": (100,),
    "fc_1_2nd_exit.weight": (100, 80), "fc_1_2nd_exit.bias": (100,),
    "fc_exit_1st.weight": (10, 100), "fc_exit_1st.bias": (10,),
    "fc_2nd_exit.weight": (10, 100), "fc_2nd_exit.bias": (10,),
}
# In the oracle the two stochastic sites are visited exit-2 first (graph order of the spec,
# t_qmodels_bayes_me.py:84 precedes :130).
