"""Generic lowering of an ``nn.Module`` to the engine's op graph.

The reference's converter (``Hardware_Artifact/converter/pytorch/nn2bnn.py:32-45``) turns ANY network into a
Bayesian one by wrapping its Linear / Conv / MaxPool leaves in ``BayesianDropout*``; its eval forward then calls the
whole network ``nSamples`` times (:26-27).  Here the converted network is traced once with ``torch.fx`` and lowered
to :class:`bayesnn_fpga_b200.engine.Graph`, so it runs through the same plan as the named multi-exit models: the
layers in front of the first stochastic site are computed once per image, everything behind it once per sample,
BatchNorm is folded, ReLU / residual adds / dropout are fused into the convolution epilogues and the mean over the
samples is accumulated on the device.

Supported vocabulary (anything else raises ``NotImplementedError`` naming the node - there is no silent fallback):
  Conv2d (square kernel, dilation 1, groups 1) [-> BatchNorm2d] [-> + residual] [-> ReLU], Linear [-> ReLU],
  BatchNorm2d [-> ReLU] behind a stochastic layer (eval mode; a stand-alone per-channel affine),
  MaxPool2d (kernel == stride, no padding) in either order with ReLU, Flatten / flatten / view(B, -1),
  AdaptiveAvgPool2d(1) / full-map AvgPool2d, Dropout / Identity (eval: no-ops), MCDropout, Masksembles1D/2D,
  BayesianDropout(Linear | MaxPool2d), BayesianDropout2D(Conv2d), one or several (list / tuple) outputs.
"""
import operator

import torch
import torch.fx
import torch.nn as nn
import torch.nn.functional as F

from . import engine as _engine
from .Dropouts import BayesianDropout, BayesianDropout2D, BayesianDropout3D, MCDropout, _DropoutBase
from .utils import _MasksemblesBase

_RELU_FNS = (F.relu, torch.relu, torch.relu_)
_ADD_FNS = (operator.add, operator.iadd, torch.add)
_FLATTEN_FNS = (torch.flatten,)


def _is_mc_module(m):
    """This package's MCDropout, or the reference's own class of that name (an nn.Dropout subclass whose forward is
    F.dropout(x, p, training=True), resnet18.py:207-210 / vgg19.py:384-387)."""
    return isinstance(m, MCDropout) or (isinstance(m, nn.Dropout) and type(m).__name__ == "MCDropout")


def _is_masksembles(m):
    """This package's Masksembles modules, or the reference's own (utils.py:115-236): same attributes."""
    return isinstance(m, _MasksemblesBase) or (
        type(m).__name__ in ("Masksembles1D", "Masksembles2D") and all(hasattr(m, a) for a in ("masks", "n", "cnt")))


class _Tracer(torch.fx.Tracer):
    def is_leaf_module(self, m, qualname):
        return isinstance(m, _DropoutBase) or _is_mc_module(m) or _is_masksembles(m) or \
            super().is_leaf_module(m, qualname)


class _Lowering:
    def __init__(self, model, input_shape):
        if len(input_shape) != 3:
            raise ValueError("input_shape must be (C, H, W), got %r" % (tuple(input_shape),))
        self.model = model
        self.gm = torch.fx.GraphModule(model, _Tracer().trace(model))
        self.mods = dict(model.named_modules())
        self.g = _engine.Graph(*[int(v) for v in input_shape])
        self.val = {}              # fx node -> TensorRef
        self.consumed = set()      # nodes folded into a group emitted elsewhere
        self.deferred = {}         # add node -> group waiting for its residual operand
        self.sites = []            # (Site, module) in creation order
        self.exit_of = {}          # fx node -> exit index of the head that produces it
        self.nonneg = set()        # ids of lowered tensors known to be >= 0 (a ReLU on them is the identity)
        self.pool_relu_done = set()   # ReLU nodes already applied in front of the max-pool / site chain they follow
        self.flat_nodes = set()       # fx nodes holding a FLATTENED view of a spatial (H*W > 1) map

    # ---- node classification -----------------------------------------------------------------
    def _mod(self, node):
        return self.mods[node.target] if node.op == "call_module" else None

    def _is_relu(self, node):
        m = self._mod(node)
        return isinstance(m, nn.ReLU) or (node.op == "call_function" and node.target in _RELU_FNS) or \
            (node.op == "call_method" and node.target in ("relu", "relu_"))

    def _is_add(self, node):
        return (node.op == "call_function" and node.target in _ADD_FNS and len(node.args) == 2 and
                all(isinstance(a, torch.fx.Node) for a in node.args)) or \
            (node.op == "call_method" and node.target in ("add", "add_") and len(node.args) == 2)

    def _is_maxpool(self, node):
        return isinstance(self._mod(node), nn.MaxPool2d) or (node.op == "call_function" and node.target is F.max_pool2d)

    def _passes_relu(self, node):
        m = self._mod(node)
        inner = m.layer if isinstance(m, _DropoutBase) else m
        return self._is_maxpool(node) or isinstance(inner, nn.MaxPool2d) or _is_mc_module(m) or _is_masksembles(m)

    @staticmethod
    def _sole_user(node):
        users = list(node.users)
        return users[0] if len(users) == 1 else None

    def _fail(self, node, why):
        raise NotImplementedError("cannot lower node %r (%s %s): %s" % (node.name, node.op, node.target, why))

    # ---- sites ---------------------------------------------------------------------------------
    def _site(self, t, module, kind, p, name, node=None):
        flat = node is not None and bool(node.all_input_nodes) and node.all_input_nodes[0] in self.flat_nodes
        if flat and kind != "mc":
            self._fail(node, "a %s site behind a Flatten of a %dx%d map (only element-wise dropout is supported there)" % (
                kind, t.H, t.W))
        dst = self.g.site(t, kind, p, module=module if kind == "mask" else None, name=name)
        self.g.sites[-1].nchw_flat = flat     # masks indexed like a stand-alone call on the flattened tensor would
        if flat:
            self.flat_nodes.add(node)
        self._bind(self.g.sites[-1], module)
        if t.id in self.nonneg:
            self.nonneg.add(dst.id)            # mask multipliers are >= 0
        return dst

    def _bind(self, site, module):
        """The site draws from the module's own Philox stream: the fused run reproduces what successive
        stand-alone forward calls of the module would draw."""
        if hasattr(module, "bnn_stream"):
            site.stream = int(module.bnn_stream)
        self.sites.append((site, module))

    # ---- groups --------------------------------------------------------------------------------
    def _conv_group(self, node):
        """conv [-> BN] [-> + residual] [-> ReLU] with look-ahead; wrapped conv: conv -> dropout2d [-> ReLU]
        (ReLU commutes with the non-negative mask multiplier)."""
        m = self._mod(node)
        wrapper = m if isinstance(m, _DropoutBase) else None
        conv = wrapper.layer if wrapper is not None else m
        if conv.kernel_size[0] != conv.kernel_size[1] or conv.stride[0] != conv.stride[1] or \
                conv.padding[0] != conv.padding[1] or conv.dilation != (1, 1) or conv.groups != 1 or \
                isinstance(conv.padding, str) or conv.padding_mode != "zeros":
            self._fail(node, "only square, undilated, ungrouped, zero-padded Conv2d")
        grp = dict(node=node, conv=conv, bn=None, add=None, res=None, relu=False, wrapper=wrapper, last=node)
        cur = node
        nxt = self._sole_user(cur)
        if wrapper is None:
            if nxt is not None and isinstance(self._mod(nxt), nn.BatchNorm2d):
                grp["bn"] = self._mod(nxt)
                cur, nxt = nxt, self._sole_user(nxt)
            if nxt is not None and self._is_add(nxt) and nxt not in self.deferred:   # else: we ARE the shortcut
                other = nxt.args[1] if nxt.args[0] is cur else nxt.args[0]
                grp["add"], grp["res"] = nxt, other
                cur, nxt = nxt, self._sole_user(nxt)
        if nxt is not None and self._is_relu(nxt):
            grp["relu"] = True
            cur = nxt
        elif nxt is not None:
            # a ReLU further down a chain of max-pools and dropout sites commutes with them (max is monotonic, the
            # mask multipliers are >= 0): relu(drop(maxpool(x))) == drop(maxpool(relu(x)))
            probe = nxt
            while probe is not None and self._passes_relu(probe):
                probe = self._sole_user(probe)
            if probe is not None and probe is not nxt and self._is_relu(probe):
                grp["relu"] = True
                grp["relu_after_pool"] = probe
        grp["last"] = cur
        return grp

    def _emit_conv_group(self, grp):
        node, conv = grp["node"], grp["conv"]
        src = self.val[node.args[0]]
        res = self.val[grp["res"]] if grp["res"] is not None else None
        t = self.g.conv(src, conv, grp["bn"], relu=grp["relu"], residual=res, name=node.name)
        if grp["relu"]:
            self.nonneg.add(t.id)
        w = grp["wrapper"]
        if w is not None:
            if isinstance(w, BayesianDropout3D) or not isinstance(w, (BayesianDropout, BayesianDropout2D)):
                self._fail(node, "unsupported dropout wrapper %s" % type(w).__name__)
            t = self._site(t, w, "mc2d" if isinstance(w, BayesianDropout2D) else "mc", w.p, node.name + ".dropout")
        cur = node
        while True:
            self.val[cur] = t
            if cur is grp["last"]:
                break
            cur = self._sole_user(cur)
        if "relu_after_pool" in grp:
            self.pool_relu_done.add(grp["relu_after_pool"])

    def _linear(self, node):
        m = self._mod(node)
        wrapper = m if isinstance(m, _DropoutBase) else None
        lin = wrapper.layer if wrapper is not None else m
        src = self.val[node.args[0]]
        if src.C * src.H * src.W != lin.in_features:
            self._fail(node, "Linear expects %d features, the lowered tensor has %d" % (lin.in_features,
                                                                                      src.C * src.H * src.W))
        relu = False
        nxt = self._sole_user(node)
        if nxt is not None and self._is_relu(nxt):
            relu = True
        t = self.g.linear(src, lin, relu=relu, name=node.name)
        if relu:
            self.nonneg.add(t.id)
        if wrapper is None and not relu and all(u.op == "output" for u in node.users):
            # `x -> [global pool] -> [site] -> Linear -> output`: a real exit head (pool + site + Linear + softmax +
            # accumulation over the samples in ONE kernel) instead of the generic ops just emitted
            tail = self._pop_tail_linear(t)
            if tail is not None:
                hsrc, lin_w, lin_b, site, name = tail
                hl = nn.Linear(lin_w.shape[1], lin_w.shape[0])
                hl.weight.data.copy_(lin_w)
                hl.bias.data.copy_(lin_b)
                self.exit_of[node] = self.g.n_exits
                self.g.head(hsrc, hl, site, name=name)
                self.val[node] = ("exit", self.exit_of[node])
                return
        if wrapper is not None:
            if not isinstance(wrapper, BayesianDropout):
                self._fail(node, "Linear wrapped in %s" % type(wrapper).__name__)
            t = self._site(t, wrapper, "mc", wrapper.p, node.name + ".dropout")
        self.val[node] = t
        if relu:
            self.val[nxt] = t
            self.consumed.add(nxt)

    def _maxpool(self, node):
        m = self._mod(node)
        wrapper = m if isinstance(m, _DropoutBase) else None
        if node.op == "call_function":
            k = node.args[1] if len(node.args) > 1 else node.kwargs.get("kernel_size")
            s = node.kwargs.get("stride", node.args[2] if len(node.args) > 2 else None) or k
            pad, dil, ceil = node.kwargs.get("padding", 0), node.kwargs.get("dilation", 1), node.kwargs.get("ceil_mode", False)
        else:
            pool = wrapper.layer if wrapper is not None else m
            k, s, pad, dil, ceil = pool.kernel_size, pool.stride or pool.kernel_size, pool.padding, pool.dilation, pool.ceil_mode
        one = lambda v: v[0] if isinstance(v, (tuple, list)) else v
        if isinstance(k, (tuple, list)) and k[0] != k[1]:
            self._fail(node, "square pooling windows only")
        if one(k) != one(s) or one(pad) != 0 or one(dil) != 1 or ceil:
            self._fail(node, "MaxPool2d needs kernel == stride, no padding, no dilation, floor mode")
        src_t = self.val[node.args[0]]
        t = self.g.maxpool(src_t, int(one(k)), name=node.name)
        if src_t.id in self.nonneg:
            self.nonneg.add(t.id)
        if wrapper is not None:
            t = self._site(t, wrapper, "mc", wrapper.p, node.name + ".dropout")
        self.val[node] = t

    def _global_avgpool(self, node, src):
        """Global average pool = a convolution whose kernel covers the map with weight 1/(H*W) on the diagonal.
        When it only feeds the final classifier it is removed again in favour of the exit-head kernel's own fused
        pool (see _pop_tail_linear)."""
        w = torch.zeros(src.C, src.C, src.H, src.W)
        idx = torch.arange(src.C)
        w[idx, idx] = 1.0 / (src.H * src.W)
        t = self.g.conv_raw(src, w, torch.zeros(src.C), 1, 0, False, None, node.name)
        self.g.ops[-1].is_gap = True
        if src.id in self.nonneg:
            self.nonneg.add(t.id)
        return t

    # ---- main walk -----------------------------------------------------------------------------
    def run(self):
        g = self.g
        outputs = None
        for node in self.gm.graph.nodes:
            if node in self.consumed:
                continue
            if node.op == "get_attr" and not node.users:
                continue                      # e.g. a leaf module's parameter that fx materialises but nobody reads
            m = self._mod(node)
            inner = m.layer if isinstance(m, _DropoutBase) else m
            if node.op == "placeholder":
                if self.val:
                    self._fail(node, "models with a single tensor input only")
                self.val[node] = g.input
            elif node.op == "output":
                outputs = node.args[0]
            elif isinstance(inner, nn.Conv2d):
                grp = self._conv_group(node)
                cur = node
                while cur is not grp["last"]:
                    cur = self._sole_user(cur)
                    if cur is not grp["add"]:
                        self.consumed.add(cur)
                if grp["res"] is not None and grp["res"] not in self.val:
                    self.deferred[grp["add"]] = grp         # the shortcut branch is traced after the main branch
                else:
                    if grp["add"] is not None:
                        self.consumed.add(grp["add"])
                    self._emit_conv_group(grp)
            elif isinstance(m, nn.BatchNorm2d):
                # a BatchNorm the convolution in front could not absorb (a stochastic site sits in between, e.g. the
                # converter's BayesianDropout2D(conv) -> BN -> ReLU): per-channel affine [+ the ReLU behind it]
                if m.training or not m.track_running_stats:
                    self._fail(node, "BatchNorm2d must be in eval mode with running statistics")
                nxt = self._sole_user(node)
                relu = nxt is not None and self._is_relu(nxt)
                t = g.affine(self.val[node.args[0]], m, relu=relu, name=node.name)
                self.val[node] = t
                if relu:
                    self.nonneg.add(t.id)
                    self.val[nxt] = t
                    self.consumed.add(nxt)
            elif self._is_add(node):
                if node not in self.deferred:
                    self._fail(node, "an add is only supported as the residual of a convolution (conv [-> BN] -> +)")
                self._emit_conv_group(self.deferred.pop(node))
            elif isinstance(inner, nn.Linear):
                self._linear(node)
            elif self._is_maxpool(node) or isinstance(inner, nn.MaxPool2d):
                self._maxpool(node)
            elif self._is_relu(node):
                if node in self.pool_relu_done:
                    self.val[node] = self.val[node.args[0]]     # already applied in front of the max-pool
                elif getattr(self.val.get(node.args[0]), "id", None) in self.nonneg:
                    self.val[node] = self.val[node.args[0]]     # ReLU of a tensor that is already >= 0
                else:
                    self._fail(node, "a ReLU must follow a convolution, a Linear or a max-pool of a convolution")
            elif _is_mc_module(m):                               # (a subclass of nn.Dropout: test it first)
                self.val[node] = self._site(self.val[node.args[0]], m, "mc", m.p, node.name, node)
            elif node.op == "call_function" and node.target is F.dropout:
                a = list(node.args) + [None] * 4
                p_ = node.kwargs.get("p", a[1] if a[1] is not None else 0.5)
                training = node.kwargs.get("training", a[2] if a[2] is not None else True)
                src = self.val[node.args[0]]
                self.val[node] = self._site(src, None, "mc", float(p_), node.name, node) if training else src
            elif isinstance(m, (nn.Dropout, nn.Dropout2d, nn.Identity)):
                self.val[node] = self.val[node.args[0]]          # eval-mode no-ops
                if node.args[0] in self.flat_nodes:
                    self.flat_nodes.add(node)
            elif _is_masksembles(m):
                self.val[node] = self._site(self.val[node.args[0]], m, "mask", 0.0, node.name, node)
            elif isinstance(m, nn.Flatten) or (node.op == "call_function" and node.target in _FLATTEN_FNS) or \
                    (node.op == "call_method" and node.target in ("flatten", "view", "reshape")):
                self.val[node] = self.val[node.args[0]]          # g.linear flattens in NCHW order itself
                v_ = self.val[node]
                if v_ is not None and (v_.H * v_.W > 1 or node.args[0] in self.flat_nodes):
                    self.flat_nodes.add(node)
            elif isinstance(m, (nn.AdaptiveAvgPool2d, nn.AvgPool2d)) or \
                    (node.op == "call_function" and node.target in (F.avg_pool2d, F.adaptive_avg_pool2d)):
                src = self.val[node.args[0]]
                if node.op == "call_function":
                    arg = node.args[1] if len(node.args) > 1 else node.kwargs.get("kernel_size", node.kwargs.get("output_size"))
                    k = arg if isinstance(arg, (tuple, list)) else (arg,) * 2
                    ok = tuple(k) == ((1, 1) if node.target is F.adaptive_avg_pool2d else (src.H, src.W)) and \
                        not node.kwargs.get("padding", 0) and len(node.args) <= 2
                elif isinstance(m, nn.AdaptiveAvgPool2d):
                    ok = m.output_size in (1, (1, 1))
                else:
                    k = m.kernel_size if isinstance(m.kernel_size, (tuple, list)) else (m.kernel_size,) * 2
                    ok = tuple(k) == (src.H, src.W) and m.padding in (0, (0, 0))
                if not ok:
                    self._fail(node, "only global average pooling")
                self.val[node] = self._global_avgpool(node, src)
            elif node.op == "call_method" and node.target == "size":
                self.val[node] = None
            elif node.op == "call_function" and node.target is operator.getitem and self.val.get(node.args[0], 0) is None:
                self.val[node] = None
            else:
                self._fail(node, "unsupported operation")
        if self.deferred:
            self._fail(next(iter(self.deferred)), "residual operand never produced")
        self._outputs(outputs)
        return g

    def _outputs(self, outputs):
        g = self.g
        outs = list(outputs) if isinstance(outputs, (tuple, list)) else [outputs]
        order = []
        for o in outs:
            v = self.val.get(o)
            if v is None:
                raise NotImplementedError("output %r is not a lowered tensor" % (o,))
            if isinstance(v, tuple):                         # already an exit head
                order.append(v[1])
                continue
            if v.H * v.W != 1:
                raise NotImplementedError("output %r is a %dx%d map: only [B, F] outputs are supported" % (
                    o.name, v.H, v.W))
            order.append(g.n_exits)
            g.head(v, _identity_linear(v.C), None, name=o.name + ".mean")      # mean over the samples of a raw tensor
        g.out_order = order                                  # exit index of every model output, in return order

    def _pop_tail_linear(self, t):
        """If `t` is produced by the LAST op, a bias/ReLU-free-form Linear over a 1x1 map (optionally behind a
        stand-alone site that only it reads), remove those ops and return what a head needs."""
        g = self.g
        if not g.ops or g.ops[-1].dst is not t:
            return None
        op = g.ops[-1]
        if op.kind != "conv" or op.relu or op.res is not None or op.site is not None or op.src.H * op.src.W != 1 or \
                op.ksize != (1, 1):
            return None
        readers = lambda t: sum(1 for o in g.ops if o.src is t or o.res is t)
        src, site, n_pop = op.src, None, 1
        if len(g.ops) >= 2 and g.ops[-2].kind == "site" and g.ops[-2].dst is src and readers(src) == 1:
            site, src, n_pop = g.ops[-2].site, g.ops[-2].src, 2
        if len(g.ops) > n_pop and getattr(g.ops[-n_pop - 1], "is_gap", False) and g.ops[-n_pop - 1].dst is src and \
                readers(src) == 1:
            src, n_pop = g.ops[-n_pop - 1].src, n_pop + 1     # the head kernel pools the map itself
        del g.ops[-n_pop:]
        w = op.weight.reshape(op.weight.shape[0], -1)
        return src, w, op.bias, site, op.name


def _identity_linear(c):
    lin = nn.Linear(c, c)
    lin.weight.data.copy_(torch.eye(c))
    lin.bias.data.zero_()
    return lin


def lower_module(model, input_shape):
    """-> (Graph, [(Site, module)]) for ``model`` applied to one [B, *input_shape] batch."""
    lo = _Lowering(model, input_shape)
    g = lo.run()
    return g, lo.sites
