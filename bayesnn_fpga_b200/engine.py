"""Plan builder and executor of the S-sample multi-exit inference path.

Replaces the reference's sample loop (``FullAnalysis._get_output``,
Software_Artifact/software/train/results_analyzer.py:236-270: ``for i in range(mc_passes): model(b_x)``
+ per-exit softmax + host-side fp64 means).  What changes:

* the network is described once as a small op graph (:class:`Graph`), built by each model class in
  the same order its reference ``forward`` executes (so stochastic sites get the same stream ids as
  the oracle's first-use numbering);
* ops whose inputs do not depend on a stochastic site form the deterministic PREFIX and run on B
  images; everything downstream of the first site runs on S_local*B images (sample-major).  This is
  the reference's own cost model (results_analyzer.py:632-637) turned into the execution plan;
* eval-mode BatchNorm (and conv bias) is folded into the conv weights at plan time; dropout /
  Masksembles sites that directly follow a stochastic conv are fused into its epilogue; each exit is
  one ``bnn_exit_head`` call that pools, masks, applies the classifier, soft-maxes and accumulates
  the per-exit sums over the samples on the device;
* samples are addressed by their GLOBAL index in the Philox counters, so a rank that owns samples
  [s0, s0 + S_local) produces exactly the partial sums of those samples (see distributed.py).

Only the C-ABI library does arithmetic; torch is used for allocation, streams and NCCL.
"""
import ctypes
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _lib

DTYPES = {"fp32": (_lib.F32, torch.float32), "fp16": (_lib.F16, torch.float16),
          "bf16": (_lib.BF16, torch.bfloat16)}
KINDS = {"mc": _lib.DROP_ELEMENT, "mc2d": _lib.DROP_CHANNEL, "mask": _lib.DROP_MASKSEMBLES}


# --------------------------------------------------------------------------------------------
# graph description (host logic, no device needed)
# --------------------------------------------------------------------------------------------
@dataclass
class TensorRef:
    id: int
    C: int
    H: int
    W: int
    stoch: bool = False          # carries the sample dimension


@dataclass
class Site:
    """One stochastic layer. kind: 'mc' (F.dropout), 'mc2d' (F.dropout2d), 'mask' (Masksembles)."""
    kind: str
    p: float
    stream: int
    name: str
    module: Optional[object] = None      # Masksembles module (owns `masks` and the rotating `cnt`)
    nchw_flat: bool = False              # element indices in NCHW-flattened order (site behind a Flatten of a map)


@dataclass
class Op:
    kind: str                            # conv | site | maxpool | affine | head
    src: Optional[TensorRef] = None
    dst: Optional[TensorRef] = None
    res: Optional[TensorRef] = None
    weight: Optional[torch.Tensor] = None    # folded, OIHW fp32 (conv) / [C, F] (head)
    bias: Optional[torch.Tensor] = None
    ksize: tuple = (1, 1)
    stride: int = 1
    pad: int = 0
    relu: bool = False
    site: Optional[Site] = None              # conv: fused epilogue site; site/head: the site itself
    pool_k: int = 1
    exit_index: int = -1
    name: str = ""


def fold_conv_bn(conv, bn=None):
    """conv weight/bias with eval-mode BatchNorm folded in:
    w' = w * g / sqrt(var + eps),  b' = beta + (b - mean) * g / sqrt(var + eps)."""
    f64 = lambda t: t.detach().cpu().double()        # plan-time host arithmetic, done once per model
    w = f64(conv.weight)
    b = f64(conv.bias) if conv.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    if bn is not None:
        g = f64(bn.weight) / torch.sqrt(f64(bn.running_var) + bn.eps)
        w = w * g.reshape(-1, 1, 1, 1)
        b = f64(bn.bias) + (b - f64(bn.running_mean)) * g
    return w.float().contiguous(), b.float().contiguous()


class Graph:
    """Op graph of one network, in execution order."""

    def __init__(self, in_channels, in_h, in_w):
        self.tensors: List[TensorRef] = []
        self.ops: List[Op] = []
        self.sites: List[Site] = []
        self.n_exits = 0
        self.n_classes = None
        self.input = self._new(in_channels, in_h, in_w, False)

    def _new(self, C, H, W, stoch):
        t = TensorRef(len(self.tensors), C, H, W, stoch)
        self.tensors.append(t)
        return t

    def conv(self, src, conv, bn=None, relu=False, residual=None, name=""):
        w, b = fold_conv_bn(conv, bn)
        return self.conv_raw(src, w, b, conv.stride[0], conv.padding[0], relu, residual, name)

    def conv_raw(self, src, w, b, stride, pad, relu=False, residual=None, name=""):
        cout, cin, kh, kw = w.shape
        assert cin == src.C, (name, cin, src.C)
        oh = (src.H + 2 * pad - kh) // stride + 1
        ow = (src.W + 2 * pad - kw) // stride + 1
        stoch = src.stoch or (residual is not None and residual.stoch)
        dst = self._new(cout, oh, ow, stoch)
        if residual is not None:
            assert (residual.C, residual.H, residual.W) == (cout, oh, ow)
        self.ops.append(Op("conv", src, dst, residual, w, b, (kh, kw), stride, pad, relu, name=name))
        return dst

    def linear(self, src, lin, relu=False, name=""):
        """nn.Linear over the NCHW-flattened src == a convolution whose kernel covers the whole map."""
        w = lin.weight.detach().cpu().float().reshape(lin.out_features, src.C, src.H, src.W).contiguous()
        b = lin.bias.detach().cpu().float().contiguous() if lin.bias is not None else torch.zeros(lin.out_features)
        return self.conv_raw(src, w, b, 1, 0, relu, None, name)

    def site(self, src, kind, p=0.5, module=None, name=""):
        s = Site(kind, float(p), len(self.sites), name, module)
        self.sites.append(s)
        dst = self._new(src.C, src.H, src.W, True)
        self.ops.append(Op("site", src, dst, site=s, name=name))
        return dst

    def new_site(self, kind, p=0.5, module=None, name=""):
        """Allocate a stream id for a site that a head will apply itself."""
        s = Site(kind, float(p), len(self.sites), name, module)
        self.sites.append(s)
        return s

    def maxpool(self, src, k, name=""):
        dst = self._new(src.C, src.H // k, src.W // k, src.stoch)
        self.ops.append(Op("maxpool", src, dst, pool_k=k, name=name))
        return dst

    def affine(self, src, bn, relu=False, name=""):
        """Eval-mode BatchNorm2d [+ ReLU] as a stand-alone per-channel affine (a BN that cannot be folded into the
        convolution in front of it: a stochastic site sits in between)."""
        f64 = lambda t: t.detach().cpu().double()
        g = f64(bn.weight) / torch.sqrt(f64(bn.running_var) + bn.eps)
        b = f64(bn.bias) - f64(bn.running_mean) * g
        dst = self._new(src.C, src.H, src.W, src.stoch)
        self.ops.append(Op("affine", src, dst, weight=g.float().contiguous(), bias=b.float().contiguous(), relu=relu,
                           name=name))
        return dst

    def head(self, src, lin, site=None, name=""):
        """global average pool -> [site] -> Linear -> softmax -> accumulate (one exit)."""
        w = lin.weight.detach().cpu().float().contiguous()
        b = lin.bias.detach().cpu().float().contiguous()
        assert w.shape[1] == src.C, (name, w.shape, src.C)
        if self.n_classes is None:
            self.n_classes = w.shape[0]
        assert self.n_classes == w.shape[0]
        self.ops.append(Op("head", src, None, weight=w, bias=b, site=site, exit_index=self.n_exits, name=name))
        self.n_exits += 1

    def mark_input_stochastic(self):
        """The INPUT already carries the sample / group dimension: x is [G * B', ...] and "sample" g is the g-th
        contiguous chunk of the batch.  This is the Masksembles training-branch / batched formulation
        (Software_Artifact/software/utils.py:158-164: batch split into n groups, group g uses mask g): nothing is
        shared between groups, so every tensor is per-group and every site picks its mask row by group index."""
        for t in self.tensors:
            t.stoch = True
        return self

    # ---- analysis used by the planner and by the tests -------------------------------------
    def fuse_sites(self):
        """Fuse a site into the epilogue of the conv that produces its input when that conv already
        runs per-sample and nobody else reads the un-masked tensor."""
        uses = {}
        for op in self.ops:
            for t in (op.src, op.res):
                if t is not None:
                    uses[t.id] = uses.get(t.id, 0) + 1
        producer = {op.dst.id: op for op in self.ops if op.dst is not None}
        out = []
        for op in self.ops:
            if op.kind == "site":
                prod = producer.get(op.src.id)
                if (prod is not None and prod.kind == "conv" and prod.site is None and op.src.stoch
                        and uses.get(op.src.id, 0) == 1 and not op.site.nchw_flat):
                    prod.site = op.site
                    prod.dst = op.dst           # conv now writes the masked tensor
                    producer[op.dst.id] = prod
                    continue
            out.append(op)
        self.ops = out

    def fuse_shortcuts(self, eligible, allow_masked=lambda producer: False):
        """A residual block's 1x1 stride-2 projection shortcut (BasicBlock.forward resnet18.py:41-46) becomes extra
        K of the convolution that adds it: conv2(h) + ds(x) is ONE GEMM over [patches(h) | x at the stride-2 centre
        tap] with the weights concatenated along K.  The shortcut tensor is then never written nor re-read as a
        residual, and the sibling-group launch in front loses its one-k-block (epilogue-bound) tiles."""
        producer = {op.dst.id: op for op in self.ops if op.dst is not None}
        for op in list(self.ops):
            if op.kind != "conv" or op.res is None:
                continue
            d = producer.get(op.res.id)
            if d is None or d.kind != "conv" or d.ksize != (1, 1) or d.stride != 2 or d.pad != 0 or d.relu or \
                    d.site is not None or d.res is not None or getattr(d, "sc", None) is not None:
                continue
            readers = [o for o in self.ops if o.src is d.dst or o.res is d.dst]
            src_prod = producer.get(d.src.id)
            masked_src = (src_prod is not None and src_prod.site is not None and src_prod.site.kind == "mask"
                          and not allow_masked(src_prod))
            if readers != [op] or d.src.stoch != op.src.stoch or masked_src or not eligible(op) or not eligible(d) or \
                    d.src.H != 2 * op.dst.H or d.src.W != 2 * op.dst.W or op.stride != 1:
                continue
            op.sc = {"src": d.src, "weight": d.weight[:, :, 0, 0].contiguous(), "name": d.name}
            op.bias = op.bias + d.bias
            op.res = None
            op.name = op.name + "+" + d.name.split(".")[-1]
            self.ops.remove(d)

    def fuse_head_pools(self, eligible):
        """The convolution whose only reader is an exit head writes the head's global average pool instead of the map
        (`F.avg_pool2d(F.relu(out), k)` -> Linear, resnet18.py:309-314,:339-344): the map never reaches HBM and the head
        reads one pooled row per image."""
        for head in [o for o in self.ops if o.kind == "head"]:
            t = head.src
            conv = next((o for o in self.ops if o.kind == "conv" and o.dst is t), None)
            hw = t.H * t.W
            if conv is None or conv.site is not None or not conv.relu or not eligible(conv) or conv.dst.C % 256 != 0 or \
                    hw < 2 or hw > 32 or hw & (hw - 1) or getattr(conv, "sc", None) is not None:
                continue
            if sum(1 for o in self.ops if o.src is t or o.res is t) != 1:
                continue
            if conv.stride == 1 and conv.ksize == (3, 3) and hw <= 16 and os.environ.get("BNN_POOL_UNFUSE_SMALLMAPS") == "1":
                # experiment: 3x3 stride-1 on a 4x4 / 2x2 map could run with position-major tiles (31 % / 56 % of the taps
                # skipped, conv_tc.cu Params::pm_nb2) if it did not pool - an image's positions sit in different tiles.
                # Measured at C2 (layer4.1.conv2): the launch 0.466 -> 0.318 ms, the head 0.052 -> 0.063 ms, the STEP
                # unchanged (38.3 k img/s both ways): under the power cap the SM clock dropped 1 432 -> 1 387 MHz - the
                # extra 0.27 GB of HBM traffic costs the energy the skipped zero-MMAs saved.  Off.
                continue
            if conv.stride == 2 and any(o is not conv and o.kind == "conv" and o.src is conv.src and o.stride == 2
                                        for o in self.ops):
                continue                      # stays in its sibling group (the shared input is read once)
            pooled = self._new(t.C, 1, 1, t.stoch)
            conv.pool_from = (t.H, t.W)
            conv.dst = pooled
            head.src = pooled

    def fuse_sibling_convs(self, eligible):
        """Group stride-2 convolutions that read the same tensor (first conv of the next stage, its 1x1 shortcut,
        first conv of the exit branch) into one 'convg' op: the input is fetched from HBM once.  A 1x1 stride-2
        pad-0 conv is exactly the centre tap of a 3x3 stride-2 pad-1 conv."""
        groups = {}
        for op in self.ops:
            if op.kind != "conv" or op.res is not None or op.site is not None or op.stride != 2 or not eligible(op):
                continue
            if op.src is self.input or getattr(op, "sc", None) is not None:
                continue                      # the (channel-padded) network input and fused shortcuts stay single launches
            if not ((op.ksize == (3, 3) and op.pad == 1) or (op.ksize == (1, 1) and op.pad == 0)):
                continue
            groups.setdefault((op.src.id, op.dst.C, op.dst.H, op.dst.W), []).append(op)
        for members in groups.values():
            if not (2 <= len(members) <= 4) or not any(m.ksize == (3, 3) for m in members):
                continue
            ws = []
            for m in members:
                w = m.weight
                if m.ksize == (1, 1):
                    w3 = torch.zeros(w.shape[0], w.shape[1], 3, 3, dtype=w.dtype)
                    w3[:, :, 1, 1] = w[:, :, 0, 0]
                    w = w3
                ws.append(w)
            fused = Op("convg", members[0].src, None, None, torch.cat(ws), torch.cat([m.bias for m in members]),
                       (3, 3), 2, 1, name="+".join(m.name for m in members))
            fused.members = members
            fused.dsts = [m.dst for m in members]
            first = self.ops.index(members[0])
            self.ops[first] = fused
            for m in members[1:]:
                self.ops.remove(m)

    def macs(self):
        """(prefix MACs, per-sample suffix MACs) per image - the reference's cost model
        (results_analyzer.py:632-637) evaluated on this graph."""
        pre = suf = 0
        flat = []
        for op in self.ops:
            flat.extend(op.members if op.kind == "convg" else [op])
        for op in flat:
            if op.kind == "conv":
                oh, ow = getattr(op, "pool_from", None) or (op.dst.H, op.dst.W)
                m = oh * ow * op.dst.C * op.src.C * op.ksize[0] * op.ksize[1]
                if getattr(op, "sc", None) is not None:
                    m += oh * ow * op.dst.C * op.sc["src"].C
            elif op.kind == "head":
                m = op.weight.shape[0] * op.weight.shape[1]
                if op.site is not None or op.src.stoch:
                    suf += m
                else:
                    pre += m
                continue
            else:
                continue
            if op.dst.stoch:
                suf += m
            else:
                pre += m
        return pre, suf


# --------------------------------------------------------------------------------------------
# executor
# --------------------------------------------------------------------------------------------
@dataclass
class MCResult:
    mean_probs: torch.Tensor       # [E, B, C]   predictive mean (results_analyzer.py:248)
    mean_logits: torch.Tensor      # [E, B, C]   (results_analyzer.py:247)
    ens_probs: torch.Tensor        # [E, B, C]   cumulative exit ensembles (:260-269)
    ens_logits: torch.Tensor
    entropy: torch.Tensor          # [E, B]      entropy of the mean (metric_utils.py:3-6, per image)
    ens_entropy: torch.Tensor
    expected_entropy: torch.Tensor  # [E, B]     mean over samples of per-sample entropy
    all_logits: Optional[torch.Tensor] = None   # [S_local, E, B, C] when requested
    samples: int = 0
    # every statistic above in ONE contiguous device buffer - [4][E][B][C] (the four [E, B, C] tensors in the order of the
    # fields) followed by [3][E][B] (the entropies): fetch it with a single device->host copy
    flat: Optional[torch.Tensor] = None


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _pm_tap_fraction(n_img, H, W, OH, OW, stride, cout_g, ksize=3):
    """Executed / algorithmic taps of a 3x3 convolution under position-major tiling (conv_tc.cu, Params::pm_nb2): with one
    output position per row-tile, taps whose input pixel is zero padding are not issued.  Mirrors the host-side condition
    in conv_tc_impl; 1.0 when the pixel-major kernel runs."""
    pm_min = int(os.environ.get("BNN_TC_PM_MIN_IMAGES", "1024"))
    if not (ksize == 3 and 4 <= OH * OW <= 64 and cout_g % 256 == 0 and n_img >= pm_min and
            os.environ.get("BNN_TC_NO_PM") is None):
        return 1.0
    rows = sum(1 for oh in range(OH) for r in range(3) if 0 <= stride * oh + r - 1 < H)
    cols = sum(1 for ow in range(OW) for r in range(3) if 0 <= stride * ow + r - 1 < W)
    frac = rows * cols / (9.0 * OH * OW)
    return frac if frac <= 0.9 else 1.0        # conv_tc_impl: pm_pays


class Engine:
    """Executes a :class:`Graph` on the current CUDA device through the C ABI."""

    def __init__(self, graph, dtype="fp16", device=None, use_tc=True, fuse=True, mask_gather=None, sample_chunk=None,
                 fuse_graph=True):
        """fuse: sites into the producing convolution's epilogue; fuse_graph: the cross-op fusions on top of that
        (projection shortcuts as extra K, head pools, sibling groups) - the 8-bit plan (q8.Q8Plan) keeps one op per
        layer and switches them off."""
        if dtype not in DTYPES:
            raise ValueError("dtype must be one of %s" % sorted(DTYPES))
        if not torch.cuda.is_available():
            raise RuntimeError("bayesnn_fpga_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        with torch.cuda.device(self.device):
            _lib.require_device()
        self.graph = graph
        if fuse:
            graph.fuse_sites()
        self.dtype_name = dtype
        self.dcode, self.tdtype = DTYPES[dtype]
        # BNN_DISABLE_TC=1 routes the 16-bit path through the CUDA-core kernel too (debugging aid)
        self.use_tc = use_tc and dtype != "fp32" and os.environ.get("BNN_DISABLE_TC") != "1"
        # stem on the tensor cores: the network input (3 channels) is stored with its channels padded to one 64-channel
        # k-block (zeros) when every layer that reads it is a 3x3 / 1x1 convolution the tcgen05 kernel accepts otherwise
        self.in_pad = 0
        if self.use_tc and graph.input.C < 64 and os.environ.get("BNN_STEM_TC", "1") != "0":
            readers = [o for o in graph.ops if o.src is graph.input or o.res is graph.input]
            # (BNN_STEM_PAD=16: 16 channels, one MMA k-step per tap instead of four - measured SLOWER, 0.098 vs 0.079 ms
            # for the C2 stem: 32-byte TMA rows and 32-byte pixel stores cost more than the three skipped k-steps save)
            self.in_pad = 16 if (graph.input.C <= 16 and all(getattr(o, "stride", 1) == 1 for o in readers) and
                                 os.environ.get("BNN_STEM_PAD") == "16") else 64
            if not readers or not all(o.kind == "conv" and o.res is None and self._tc_eligible(o) for o in readers):
                self.in_pad = 0
        self.gather_mode = int(os.environ.get("BNN_MASK_GATHER", "1")) if mask_gather is None else int(mask_gather)
        if self.use_tc and fuse and fuse_graph and os.environ.get("BNN_SHORTCUT_FUSION", "1") != "0":
            # a Masksembles-masked block input may feed a fused shortcut only when it is materialised densely: a site
            # fused into its producer's epilogue (not the boundary site, which moves into per-mask weight sets, and not
            # the gathered layout of mode 2)
            graph.fuse_shortcuts(self._tc_eligible,
                                 allow_masked=lambda prod: self.gather_mode == 0 or
                                 (prod.kind == "conv" and self.gather_mode < 2))
        if self.use_tc and fuse and fuse_graph and os.environ.get("BNN_HEAD_POOL_FUSION", "1") != "0":
            graph.fuse_head_pools(self._tc_eligible)
        if self.use_tc and fuse and fuse_graph and os.environ.get("BNN_NO_SIBLING_FUSION") != "1":
            graph.fuse_sibling_convs(self._tc_eligible)
        self.sample_chunk = int(os.environ.get("BNN_SAMPLE_CHUNK", "0")) if sample_chunk is None else int(sample_chunk)
        self.launches = 0
        self._prof = None
        self._graphs = {}
        self._h2d_done = None
        self._graph_collectives = os.environ.get("BNN_GRAPH_COLLECTIVES", "1") != "0"
        self._bufs = {}
        self._prepare_weights()
        # Masksembles2D as smaller / re-weighted GEMMs.  0: dense 0/1 multiplies everywhere (the reference's form);
        # 1 (default): the site at the prefix boundary moves into per-mask weight sets of its consumers;
        # 2: additionally, sites fused into a producer's epilogue store the gathered (kept-channels-only) layout.
        self.gather = {}
        if self.use_tc and self.gather_mode > 0:
            self._plan_gather()
        self.boundary = {}
        # opt-in (BNN_BOUNDARY_FUSION=1): bit-identical, removes the S masked copies (1.07 GB -> 335 MB at C2) - but
        # measured SLOWER on B200 (C2 7.73 vs 7.02 ms/step): the in-shared-memory masking hop sits on the sibling launch's
        # critical path (1.66 vs 0.59 ms), see DESIGN.md
        if self.use_tc and fuse and fuse_graph and os.environ.get("BNN_BOUNDARY_FUSION", "0") == "1":
            self._plan_boundary()

    # ---- plan-time weight packing -----------------------------------------------------------
    def _tc_eligible(self, op):
        """Same rule as bnn_conv2d_tc (conv_tc.cu): decided at plan time, never a silent run-time fallback."""
        kh, kw = op.ksize
        pow2 = lambda v: v > 0 and (v & (v - 1)) == 0
        cin = self.in_pad if (op.src is self.graph.input and self.in_pad) else op.src.C
        oh, ow = getattr(op, "pool_from", None) or (op.dst.H, op.dst.W)
        return (self.use_tc and kh == kw and ((kh == 3 and op.pad == 1) or (kh == 1 and op.pad == 0))
                and op.stride in (1, 2) and op.dst.C % 64 == 0
                and (cin % 64 == 0 or (op.src is self.graph.input and cin == self.in_pad and cin % 16 == 0 and
                                       op.stride == 1 and getattr(op, "pool_from", None) is None))
                and (op.stride == 1 or (op.src.H % 2 == 0 and op.src.W % 2 == 0))
                and pow2(oh) and pow2(ow) and ow <= 128)

    def _prepare_weights(self):
        dev = self.device
        for op in self.graph.ops:
            if op.kind == "convg":
                op.d_w = op.weight.permute(0, 2, 3, 1).contiguous().to(dev, self.tdtype)
                op.d_b = op.bias.to(dev, torch.float32)
                op.relu_mask = sum(1 << i for i, m in enumerate(op.members) if m.relu)
                op.center_mask = sum(1 << i for i, m in enumerate(op.members) if m.ksize == (1, 1))
            elif op.kind == "conv":
                w = op.weight.permute(0, 2, 3, 1).contiguous()       # [Cout][KH][KW][Cin]
                op.use_tc = self._tc_eligible(op)
                # A 3x3 pad-1 convolution on a 2x2 map (the last VGG block at 32x32 inputs) is a DENSE map from the 4
                # input pixels to the 4 output pixels: every (output p, input q) pair is linked by exactly one tap.  In
                # NHWC memory the image is one row of 4*Cin values and the output one row of 4*Cout, so the layer runs
                # as ONE 1x1 GEMM with K = 4*Cin and 4*Cout outputs - 16 Cin x Cout blocks instead of the 36 the nine
                # taps would issue (20 of them multiplying zero padding): 2.25x fewer MMAs.  Element-wise dropout in the
                # epilogue sees the same linear element indices; channel-wise sites keep the tap form.
                op.fc22 = (op.use_tc and op.ksize == (3, 3) and op.pad == 1 and op.stride == 1 and op.src.H == 2 and
                           op.src.W == 2 and getattr(op, "sc", None) is None and getattr(op, "pool_from", None) is None and
                           (op.site is None or op.site.kind == "mc") and os.environ.get("BNN_SMALLMAP_FC", "1") != "0")
                if op.fc22:
                    co, ci = op.weight.shape[:2]
                    wf = torch.zeros((4 * co, 4 * ci), dtype=torch.float32)
                    for p_ in range(4):
                        for q_ in range(4):
                            kh_, kw_ = q_ // 2 - p_ // 2 + 1, q_ % 2 - p_ % 2 + 1
                            wf[p_ * co:(p_ + 1) * co, q_ * ci:(q_ + 1) * ci] = op.weight[:, :, kh_, kw_]
                    op.d_w = wf.contiguous().to(dev, self.tdtype)
                    op.d_b = op.bias.repeat(4).to(dev, torch.float32)
                    continue
                if op.use_tc and op.src is self.graph.input and self.in_pad:
                    w = torch.nn.functional.pad(w, (0, self.in_pad - w.shape[3]))   # zero weights on the padding channels
                if getattr(op, "sc", None) is not None:              # K-concatenated: 3x3 weights | shortcut weights
                    w = torch.cat([w.reshape(w.shape[0], -1), op.sc["weight"]], dim=1).contiguous()
                op.d_w = w.to(dev, self.tdtype if op.use_tc else torch.float32)
                op.d_b = op.bias.to(dev, torch.float32)
            elif op.kind == "affine":
                op.d_w = op.weight.to(dev, torch.float32)
                op.d_b = op.bias.to(dev, torch.float32)
            elif op.kind == "head":
                op.d_w = op.weight.t().contiguous().to(dev, torch.float32)      # [F][C] for the head kernel
                op.d_b = op.bias.to(dev, torch.float32)
                # wide heads on 16-bit features: the Linear runs on mma.sync with hi/lo-split operands
                C_, F_ = op.weight.shape
                op.head_mma = (self.dtype_name != "fp32" and C_ > 32 and F_ % 32 == 0 and
                               os.environ.get("BNN_HEAD_MMA", "1") != "0")
                # ... or, on 16-bit features with F % 64 == 0, the whole classifier as ONE tcgen05 GEMM over hi/lo-split
                # operands (bnn_exit_head_tc): [x_hi | x_lo] x [w_hi | w_lo | w_hi]^T with fp32 logits
                # (narrow heads, C <= 32, stay on the one-kernel FFMA form: three launches cost more than they save)
                op.head_tc = (self.dtype_name != "fp32" and self.use_tc and F_ % 64 == 0 and
                              C_ >= int(os.environ.get("BNN_HEAD_TC_MIN_C", "33")) and
                              os.environ.get("BNN_HEAD_TC", "1") != "0")
                # narrow heads: one warp per (sample, image) row + the sample-ordered soft-max kernel (bnn_exit_head_rows);
                # the block-per-image kernel is a 50-us latency chain per exit whatever the size
                # (measured at C2 / C5, ncu: 30 vs 34 us at S = 32, 87 vs 127 us at S = 128 for a pooled 512-feature head; an
                # un-pooled 4x4 map is SLOWER this way - 0.116 vs 0.063 ms - so those stay on the block-per-image kernel)
                op.head_rows = (not op.head_tc and C_ <= 32 and F_ % 8 == 0 and C_ * F_ * 4 <= 200 * 1024 and
                                op.src.H * op.src.W <= 4 and os.environ.get("BNN_HEAD_ROWS", "1") != "0")
                if op.head_rows:
                    op.d_w_cf = op.weight.contiguous().to(dev, torch.float32)       # [C][F], the nn.Linear layout
                if op.head_mma or op.head_tc:
                    w32 = op.weight.contiguous().to(dev, torch.float32)             # [C][F], the nn.Linear layout
                    op.d_w_hi = torch.empty((C_, F_), dtype=self.tdtype, device=dev)
                    op.d_w_lo = torch.empty((C_, F_), dtype=self.tdtype, device=dev)
                    with torch.cuda.device(dev):
                        _lib.check(self.lib.bnn_split16(_ptr(w32), _ptr(op.d_w_hi), _ptr(op.d_w_lo), C_ * F_, self.dcode,
                                                        ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
                        torch.cuda.current_stream(dev).synchronize()
                if op.head_tc:
                    # features exactly representable in 16 bits (already-pooled 16-bit rows, mask multipliers that are
                    # 0/1 or powers of two) need no x_lo block
                    p_ = op.site.p if op.site is not None and op.site.kind != "mask" else 0.0
                    keep_scale = 1.0 / (1.0 - p_) if p_ < 1.0 else 0.0
                    pow2 = keep_scale == 0.0 or abs(np.log2(keep_scale) - round(np.log2(keep_scale))) < 1e-12
                    op.head_with_lo = int(not (op.src.H * op.src.W == 1 and pow2))
                    op.c_pad = (C_ + 63) // 64 * 64 if C_ <= 128 else (C_ + 255) // 256 * 256
                    parts = [op.d_w_hi, op.d_w_lo] + ([op.d_w_hi] if op.head_with_lo else [])
                    w3 = torch.zeros((op.c_pad, len(parts) * F_), dtype=self.tdtype, device=dev)
                    w3[:C_] = torch.cat(parts, dim=1)
                    op.d_w3 = w3.contiguous()
                    op.d_b_pad = torch.zeros(op.c_pad, dtype=torch.float32, device=dev)
                    op.d_b_pad[:C_] = op.d_b
            if op.site is not None and op.site.kind == "mask":
                op.d_masks = op.site.module.masks.detach().to(dev, torch.float32).contiguous()

    # ---- prefix boundary without the S masked copies -------------------------------------------
    def _plan_boundary(self):
        """An element-wise MC-dropout site on the deterministic prefix whose output is read ONLY by an even number of
        128-channel stride-2 sibling convolutions (one grouped launch) and by fused 1x1 stride-2 projection shortcuts:
        the site launch writes keep BITS (+ the even/even plane the shortcuts read) instead of S masked copies, and the
        sibling launch ANDs the bits into its tiles in shared memory (north star items 1 and 3: the prefix is computed
        once AND never replicated; bnn_boundary_bits / bnn_conv2d_tc_grouped_masked / bnn_conv2d_tc_shortcut_plane)."""
        g = self.graph
        for op in g.ops:
            if op.kind != "site" or op.site.kind != "mc" or op.src.stoch or op.site.nchw_flat or op.site.p >= 1.0:
                continue
            t = op.dst
            if t.C % 64 != 0 or t.H % 2 or t.W % 2 or 128 % (t.W // 2) != 0:
                continue
            groups = [o for o in g.ops if o.kind == "convg" and o.src is t]
            scs = [o for o in g.ops if o.kind == "conv" and getattr(o, "sc", None) is not None and o.sc["src"] is t]
            others = [o for o in g.ops if (o.src is t or o.res is t) and o not in groups]
            if len(groups) != 1 or others or any(not o.use_tc or o.dst.H * 2 != t.H for o in scs):
                continue
            cg = groups[0]
            if len(cg.dsts) % 2 or any(d.C != 128 for d in cg.dsts) or cg.center_mask != 0:
                continue
            self.boundary[t.id] = {"site": op, "convg": cg, "shortcuts": scs}
            cg.boundary = op
            for o in scs:
                o.boundary = op

    # ---- Masksembles gathered layout ----------------------------------------------------------
    def _plan_gather(self):
        """Masksembles2D sites whose output is read by tensor-core convolutions only (north star (3); the reference
        multiplies by a 0/1 table and convolves the zeros, utils.py:165-169).

        * "weights": the site sits at the prefix boundary (its input has no sample dimension).  W (m . x) ==
          (W . m) x, so the consumers read the ONE deterministic tensor through per-mask weight sets whose dropped
          input channels are zero: the S masked copies are never written or read and the site launch disappears.
        * "compact" (gather_mode 2): the site is fused into its producer's epilogue, which stores only the channels
          the sample's mask keeps (kept count rounded up to 16); the consumers run a smaller GEMM over those channels
          with weights gathered per mask row.
        Whether a given batch size can use it is decided in :meth:`_gather_ids` (an MMA tile pair must not straddle
        two samples, one weight set per tile)."""
        g, dev = self.graph, self.device
        for op in g.ops:
            site = op.site
            if op.kind not in ("site", "conv") or site is None or site.kind != "mask" or op.dst is None:
                continue
            t = op.dst
            if op.kind == "site":
                mode = "weights" if not op.src.stoch else None
            else:
                mode = "compact" if (self.gather_mode >= 2 and op.use_tc and (t.C == 128 or t.C % 256 == 0)) else None
            if mode is None:
                continue
            readers = [o for o in g.ops if o.src is t or o.res is t]
            if any(getattr(o, "sc", None) is not None and o.sc["src"] is t for o in g.ops):
                continue                      # a fused shortcut reads the dense masked tensor
            if not readers or any(o.kind not in ("conv", "convg") or o.res is not None or o.site is not None
                                  or (o.kind == "conv" and not o.use_tc) for o in readers):
                continue
            masks = site.module.masks.detach().cpu().numpy() != 0                   # [n, C]
            n, C = masks.shape
            if C != t.C or C >= 32768:
                continue
            kept = [np.flatnonzero(masks[r]) for r in range(n)]
            info = {"site": site, "mode": mode, "kept": [len(k) for k in kept], "n": n, "readers": readers, "op": op}
            if mode == "compact":
                kc = (max(len(k) for k in kept) + 15) // 16 * 16
                if kc >= C:
                    continue                                                        # nothing to save
                pos = np.full((n, C), -1, np.int16)
                idx = np.full((n, kc), -1, np.int16)
                for r, k in enumerate(kept):
                    pos[r, k] = np.arange(len(k), dtype=np.int16)
                    idx[r, :len(k)] = k
                info.update(kc=kc, pos=torch.from_numpy(pos).to(dev), idx=torch.from_numpy(idx).to(dev))
            else:
                info.update(kc=C)
            for o in readers:
                w = o.weight                                                        # OIHW fp32 (grouped: concatenated)
                wg = torch.zeros((n, w.shape[0], w.shape[2], w.shape[3], info["kc"]), dtype=torch.float32)
                for r, k in enumerate(kept):
                    kk = torch.from_numpy(k)
                    if mode == "compact":
                        wg[r, :, :, :, :len(k)] = w[:, kk, :, :].permute(0, 2, 3, 1)
                    else:
                        wg[r][:, :, :, kk] = w[:, kk, :, :].permute(0, 2, 3, 1)     # dropped columns stay zero
                o.d_wg = wg.contiguous().to(dev, self.tdtype)
                o.gather = info
            self.gather[t.id] = info

    def _gather_ids(self, B):
        """{tensor id: "weights" | "compact"} for batch size B."""
        out = {}
        for tid, info in self.gather.items():
            t = self.graph.tensors[tid]
            ok = (B * t.H * t.W) % 32 == 0
            for o in info["readers"]:
                d = o.dsts[0] if o.kind == "convg" else o.dst
                ok = ok and (B * d.H * d.W) % 256 == 0
            if ok:
                out[tid] = info["mode"]
        return out

    def _compact_ids(self, B):
        return {tid for tid, m in self._gather_ids(B).items() if m == "compact"}

    # ---- buffers ----------------------------------------------------------------------------
    def _buffers(self, B, S_local, want_logits):
        key = (B, S_local, want_logits)
        if key in self._bufs:
            return self._bufs[key]
        g, dev = self.graph, self.device
        acts = {}
        gmode = self._gather_ids(B)
        compact = {tid for tid, m in gmode.items() if m == "compact"}
        chunk = S_local if self.sample_chunk <= 0 else min(S_local, self.sample_chunk)
        live = {g.input.id}
        for op in g.ops:
            live.update(t.id for t in (op.src, op.dst, op.res) + tuple(getattr(op, "dsts", ())) if t is not None)
            if getattr(op, "sc", None) is not None:
                live.add(op.sc["src"].id)
        bnd = {}
        for tid, info in self.boundary.items():
            t = g.tensors[tid]
            bnd[tid] = {
                "x_scaled": torch.empty((B, t.H, t.W, t.C), dtype=self.tdtype, device=dev),
                "bits": torch.empty((chunk * B * t.H * t.W * t.C // 8,), dtype=torch.uint8, device=dev),
                "plane": (torch.empty((chunk * B, t.H // 2, t.W // 2, t.C), dtype=self.tdtype, device=dev)
                          if info["shortcuts"] else None)}
        for t in g.tensors:
            if t.id not in live or gmode.get(t.id) == "weights" or t.id in self.boundary:
                continue                      # e.g. the un-masked output of a conv with a fused site
            n = (chunk if t.stoch else 1) * B
            if t is g.input and self.in_pad:
                acts[t.id] = torch.zeros((n, t.H, t.W, self.in_pad), dtype=self.tdtype, device=dev)
                continue
            if t.id in compact:
                # gathered layout; zero-filled once: the padding slots [kept, kc) are never written
                acts[t.id] = torch.zeros((n, t.H, t.W, self.gather[t.id]["kc"]), dtype=self.tdtype, device=dev)
                continue
            acts[t.id] = torch.empty((n, t.H, t.W, t.C), dtype=self.tdtype, device=dev)
        E, C = g.n_exits, g.n_classes
        head_ws = None
        tc_heads = [o for o in g.ops if o.kind == "head" and getattr(o, "head_tc", False)]
        row_heads = [o for o in g.ops if o.kind == "head" and getattr(o, "head_rows", False)]
        if tc_heads or row_heads:   # one workspace pair shared by all heads that use it (they run back to back on the stream)
            rows = chunk * B
            a_ws = (torch.empty(rows * max((2 if o.head_with_lo else 1) * o.src.C for o in tc_heads), dtype=self.tdtype,
                                device=dev) if tc_heads else None)
            head_ws = (a_ws, torch.empty(rows * max([o.c_pad for o in tc_heads] + [C for _ in row_heads]),
                                         dtype=torch.float32, device=dev))
        st = {
            "x": torch.empty(((chunk if g.input.stoch else 1) * B, g.input.C, g.input.H, g.input.W), dtype=torch.float32,
                             device=dev),
            "sums": torch.zeros((2 * E * B * C + E * B,), dtype=torch.float32, device=dev),
            "out": torch.empty((4 * E * B * C + 3 * E * B,), dtype=torch.float32, device=dev),
            "logits": (torch.empty((E, S_local, B, C), dtype=torch.float32, device=dev) if want_logits else None),
            "acts": acts,
            "head_ws": head_ws,
            "boundary": bnd,
            "compact": compact,
            "gmode": gmode,
            "chunk": chunk,
        }
        self._bufs[key] = st
        return st

    def release_buffers(self):
        self._graphs.clear()
        self._bufs.clear()

    # ---- launch helpers ---------------------------------------------------------------------
    def _drop_desc(self, op_site, d_masks, B, sample0, seed, mask_offset=None, compact=None):
        d = _lib.DropDesc()
        if op_site is None:
            d.kind = _lib.DROP_NONE
            d.batch = B
            return d
        d.kind = KINDS[op_site.kind]
        d.nchw_flat = int(op_site.nchw_flat)
        d.p = op_site.p
        d.seed = seed & 0xFFFFFFFFFFFFFFFF
        d.stream_id = op_site.stream
        d.sample0 = sample0
        d.batch = B
        if op_site.kind == "mask":
            d.masks = d_masks.data_ptr()
            d.n_masks = d_masks.shape[0]
            # kernel row = (cnt0 + sample0 + s) % n; we want (module.cnt + mask_offset + s) % n
            off = sample0 if mask_offset is None else mask_offset
            d.cnt0 = (int(op_site.module.cnt) + off - sample0) % d.n_masks
            if compact is not None:
                d.compact_pos = compact["pos"].data_ptr()
                d.compact_idx = compact["idx"].data_ptr()
                d.compact_c = compact["kc"]
        return d

    def _gathered_call(self, op, st, B, S_local, sample0, mask_offset, stream):
        """(call, flops, bytes) of a convolution (group) that reads a tensor in the gathered layout."""
        info, lib, acts = op.gather, self.lib, st["acts"]
        dsts = op.dsts if op.kind == "convg" else [op.dst]
        members = op.members if op.kind == "convg" else [op]
        d0 = dsts[0]
        n_img = S_local * B
        ys = (ctypes.c_void_p * len(dsts))(*[acts[d.id].data_ptr() for d in dsts])
        relu_mask = op.relu_mask if op.kind == "convg" else int(op.relu)
        center_mask = op.center_mask if op.kind == "convg" else 0
        ksize = 3 if op.kind == "convg" else op.ksize[0]
        off = sample0 if mask_offset is None else mask_offset
        cnt0 = (int(info["site"].module.cnt) + off) % info["n"]
        out_px = n_img * d0.H * d0.W
        # "weights" mode: read the site's deterministic INPUT (B images) - the masked copies do not exist
        src = info["op"].src if info["mode"] == "weights" else op.src
        has_samples = 0 if info["mode"] == "weights" else 1
        kept = sum(info["kept"]) / len(info["kept"])                 # algorithmic: the channels a mask keeps
        flops = sum(2 * out_px * m.dst.C * kept * m.ksize[0] * m.ksize[1] for m in members)
        es = 2
        nbytes = ((n_img if has_samples else B) * op.src.H * op.src.W * info["kc"] + out_px * d0.C * len(dsts)) * es \
            + op.d_wg.numel() * op.d_wg.element_size()
        call = lambda: lib.bnn_conv2d_tc_gathered(
            _ptr(acts[src.id]), _ptr(op.d_wg), _ptr(op.d_b), ys, len(dsts), relu_mask, center_mask, self.dcode, n_img,
            op.src.H, op.src.W, info["kc"], d0.C, ksize, op.stride, info["n"], cnt0, 0, B, has_samples, stream)
        return call, flops, nbytes

    def _batch(self, x, S_local):
        """Images per sample: the batch itself, or (input with a group dimension) one group of it."""
        if not self.graph.input.stoch:
            return x.shape[0]
        if S_local <= 0 or x.shape[0] % S_local != 0:
            raise ValueError('Batch size must be divisible by n, got batch {} and n {}'.format(x.shape[0], S_local))
        return x.shape[0] // S_local

    def sums_views(self, st, B):
        E, C = self.graph.n_exits, self.graph.n_classes
        n = E * B * C
        s = st["sums"]
        return s[:n].view(E, B, C), s[n:2 * n].view(E, B, C), s[2 * n:].view(E, B)

    def _launch(self, kernel, name, flops, nbytes, call, flops_exec=None):
        """One C-ABI launch; with profiling on, bracket it with CUDA events on the launching stream.
        flops = ALGORITHMIC work of the reference's layer (the hook-derived MAC count); flops_exec = what
        the kernel really multiplies when that is LESS (3x3 windows on 1x1 maps run their centre tap only)."""
        if self._prof is None:
            _lib.check(call())
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(call())
            e1.record()
            self._prof.append((kernel, name, flops, nbytes, e0, e1, flops if flops_exec is None else flops_exec))
        self.launches += 1

    def enqueue(self, x, S_local, sample0=0, seed=0x5EED, accumulate=False, want_logits=False, mask_offset=None):
        """Enqueue the whole pass for local samples [sample0, sample0 + S_local) on the current
        stream. Returns the buffer set (sums are in st['sums'])."""
        g, lib = self.graph, self.lib
        B = self._batch(x, S_local)
        if tuple(x.shape[1:]) != (g.input.C, g.input.H, g.input.W):
            raise ValueError("expected input [B, %d, %d, %d], got %s" % (g.input.C, g.input.H, g.input.W,
                                                                         tuple(x.shape)))
        if S_local < 0:
            raise ValueError("S_local must be >= 0")
        st = self._buffers(B, S_local, want_logits)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        acts = st["acts"]
        st["x"].copy_(x, non_blocking=True)
        if B == 0:
            return st
        es = 4 if self.dtype_name == "fp32" else 2
        n_img_in = x.shape[0]
        n_in = n_img_in * g.input.C * g.input.H * g.input.W
        self._launch("layout", "nchw_to_nhwc", 0, n_in * (4 + es), lambda: lib.bnn_nchw_to_nhwc_pitch(
            _ptr(st["x"]), _ptr(acts[g.input.id]), self.dcode, n_img_in, g.input.C, g.input.H, g.input.W,
            self.in_pad or g.input.C, stream))
        # deterministic prefix once; everything behind the first stochastic site in chunks of `chunk` samples so that
        # the activations handed from one layer to the next stay L2-resident (126 MB) instead of round-tripping HBM
        chunk = st["chunk"]
        self._run_ops(st, B, S_local, sample0, seed, accumulate, mask_offset, stream, "det", 0)
        for c in range(0, S_local, max(chunk, 1)):
            self._run_ops(st, B, min(chunk, S_local - c), sample0 + c, seed, accumulate or c > 0,
                          None if mask_offset is None else mask_offset + c, stream, "stoch", c)
        return st

    def _run_ops(self, st, B, S_local, sample0, seed, accumulate, mask_offset, stream, phase, s_off):
        """Enqueue the ops of one phase: "det" = ops on tensors without a sample dimension (and the heads that read
        them, which loop over all samples themselves), "stoch" = per-sample ops for local samples
        [s_off, s_off + S_local) whose global index starts at `sample0`."""
        g, lib, acts = self.graph, self.lib, st["acts"]
        es = 4 if self.dtype_name == "fp32" else 2
        sum_p, sum_l, sum_pl = self.sums_views(st, B)
        compact, gmode = st["compact"], st["gmode"]
        for op in g.ops:
            out_t = op.dsts[0] if op.kind == "convg" else (op.dst if op.dst is not None else op.src)
            is_stoch = True if op.kind == "site" else out_t.stoch
            if is_stoch != (phase == "stoch"):
                continue
            if op.kind in ("conv", "convg") and op.src.id in gmode:
                if S_local == 0:
                    continue
                call, flops, nbytes = self._gathered_call(op, st, B, S_local, sample0, mask_offset, stream)
                tag = " [gathered K]" if gmode[op.src.id] == "compact" else " [per-mask weights]"
                self._launch("conv_tc", op.name + tag, flops, nbytes, call)
            elif op.kind == "site" and gmode.get(op.dst.id) == "weights":
                continue                      # folded into the consumers' weight sets
            elif op.kind == "convg":
                d0 = op.dsts[0]
                n_img = (S_local if d0.stoch else 1) * B
                if n_img == 0:
                    continue
                ys = (ctypes.c_void_p * len(op.dsts))(*[acts[d.id].data_ptr() for d in op.dsts])
                out_px = n_img * d0.H * d0.W
                flops = sum(2 * out_px * m.dst.C * m.src.C * m.ksize[0] * m.ksize[1] for m in op.members)
                if getattr(op, "boundary", None) is not None:
                    bb = st["boundary"][op.src.id]
                    nbytes = (B * op.src.H * op.src.W * op.src.C + out_px * d0.C * len(op.dsts)) * es + bb["bits"].numel() \
                        + op.d_w.numel() * op.d_w.element_size()
                    self._launch("conv_tc", op.name + " [keep bits applied in shared memory]", flops, nbytes,
                                 lambda: lib.bnn_conv2d_tc_grouped_masked(
                                     _ptr(bb["x_scaled"]), _ptr(bb["bits"]), _ptr(op.d_w), _ptr(op.d_b), ys, len(op.dsts),
                                     op.relu_mask, self.dcode, n_img, B, op.src.H, op.src.W, op.src.C, d0.C, stream))
                    continue
                nbytes = (n_img * op.src.H * op.src.W * op.src.C + out_px * d0.C * len(op.dsts)) * es \
                    + op.d_w.numel() * op.d_w.element_size()
                frac = _pm_tap_fraction(n_img, op.src.H, op.src.W, d0.H, d0.W, 2, d0.C)
                self._launch("conv_tc", op.name + (" [position-major]" if frac < 1 else ""), flops, nbytes,
                             lambda: lib.bnn_conv2d_tc_grouped(
                    _ptr(acts[op.src.id]), _ptr(op.d_w), _ptr(op.d_b), ys, len(op.dsts), op.relu_mask, op.center_mask, self.dcode,
                    n_img, op.src.H, op.src.W, op.src.C, d0.C, 3, 2, stream), flops_exec=flops * frac)
            elif op.kind == "conv":
                n_img = (S_local if op.dst.stoch else 1) * B
                if n_img == 0:
                    continue
                src = acts[op.src.id]
                if op.dst.stoch and not op.src.stoch:
                    raise NotImplementedError("conv %s mixes deterministic input with stochastic residual" % op.name)
                res = acts[op.res.id] if op.res is not None else None
                if op.res is not None and op.dst.stoch and not op.res.stoch:
                    raise NotImplementedError("conv %s: deterministic residual into stochastic output" % op.name)
                dd = self._drop_desc(op.site, getattr(op, "d_masks", None), B, sample0, seed, mask_offset,
                                     self.gather[op.dst.id] if op.dst.id in compact else None)
                kh, kw = op.ksize
                pool_from = getattr(op, "pool_from", None)
                out_px = n_img * (pool_from[0] * pool_from[1] if pool_from else op.dst.H * op.dst.W)
                flops = 2 * out_px * op.dst.C * op.src.C * kh * kw
                nbytes = (n_img * op.src.H * op.src.W * op.src.C + out_px * op.dst.C * (2 if res is not None else 1)) * es \
                    + op.d_w.numel() * op.d_w.element_size()
                sc = getattr(op, "sc", None)
                if pool_from:
                    nbytes = (n_img * op.src.H * op.src.W * op.src.C + n_img * op.dst.C + (out_px * op.dst.C if res is not None else 0)) * es \
                        + op.d_w.numel() * op.d_w.element_size()
                cin = self.in_pad if (op.src is g.input and self.in_pad and op.use_tc) else op.src.C
                if sc is not None and getattr(op, "boundary", None) is not None:
                    flops += 2 * out_px * op.dst.C * sc["src"].C
                    nbytes += out_px * sc["src"].C * es
                    plane = st["boundary"][sc["src"].id]["plane"]
                    call = lambda: lib.bnn_conv2d_tc_shortcut_plane(
                        _ptr(src), _ptr(op.d_w), _ptr(op.d_b), _ptr(res), _ptr(acts[op.dst.id]), self.dcode, n_img,
                        op.src.H, op.src.W, op.src.C, op.dst.C, kh, op.stride, int(op.relu), ctypes.byref(dd),
                        _ptr(plane), sc["src"].C, stream)
                elif sc is not None:
                    flops += 2 * out_px * op.dst.C * sc["src"].C
                    nbytes += n_img * sc["src"].H * sc["src"].W * sc["src"].C * es
                    call = lambda: lib.bnn_conv2d_tc_shortcut(
                        _ptr(src), _ptr(op.d_w), _ptr(op.d_b), _ptr(res), _ptr(acts[op.dst.id]), self.dcode, n_img,
                        op.src.H, op.src.W, op.src.C, op.dst.C, kh, op.stride, int(op.relu), ctypes.byref(dd),
                        _ptr(acts[sc["src"].id]), sc["src"].H, sc["src"].W, sc["src"].C, stream)
                elif pool_from:
                    call = lambda: lib.bnn_conv2d_tc_pooled(
                        _ptr(src), _ptr(op.d_w), _ptr(op.d_b), _ptr(res), _ptr(acts[op.dst.id]), self.dcode, n_img,
                        op.src.H, op.src.W, op.src.C, op.dst.C, kh, op.stride, int(op.relu), stream)
                elif getattr(op, "fc22", False):
                    call = lambda: lib.bnn_conv2d_tc(
                        _ptr(src), _ptr(op.d_w), _ptr(op.d_b), _ptr(res), _ptr(acts[op.dst.id]), self.dcode, n_img,
                        1, 1, 4 * op.src.C, 4 * op.dst.C, 1, 1, int(op.relu), ctypes.byref(dd), stream)
                elif op.use_tc:
                    call = lambda: lib.bnn_conv2d_tc(
                        _ptr(src), _ptr(op.d_w), _ptr(op.d_b), _ptr(res), _ptr(acts[op.dst.id]), self.dcode, n_img,
                        op.src.H, op.src.W, cin, op.dst.C, kh, op.stride, int(op.relu), ctypes.byref(dd), stream)
                else:
                    call = lambda: lib.bnn_conv2d_simt(
                        _ptr(src), _ptr(op.d_w), _ptr(op.d_b), _ptr(res), _ptr(acts[op.dst.id]), self.dcode, n_img,
                        op.src.H, op.src.W, op.src.C, op.dst.C, kh, kw, op.stride, op.pad, int(op.relu),
                        ctypes.byref(dd), stream)
                # bnn_conv2d_tc runs only the centre tap of a 3x3 window on a 1x1 map (conv_tc.cu, bit-exact): the other
                # eight taps are zero padding and must not count as executed work
                tap_skip = (op.use_tc and kh == 3 and op.src.H == 1 and op.src.W == 1 and op.stride == 1 and
                            os.environ.get("BNN_TC_NO_TAP_SKIP") is None)
                fexec = flops / 9 if tap_skip else (flops * 16 / 36 if getattr(op, "fc22", False) else None)
                tag = " [2x2 map as one GEMM]" if getattr(op, "fc22", False) else ""
                if (fexec is None and op.use_tc and not pool_from and op.dst.id not in compact and
                        getattr(op, "boundary", None) is None):
                    frac = _pm_tap_fraction(n_img, op.src.H, op.src.W, op.dst.H, op.dst.W, op.stride, op.dst.C, kh)
                    if frac < 1:        # the fused 1x1 shortcut's k-blocks (sc) are all executed
                        conv_flops = 2 * out_px * op.dst.C * op.src.C * kh * kw
                        fexec, tag = flops - conv_flops * (1 - frac), " [position-major]"
                self._launch("conv_tc" if op.use_tc else "conv_simt", op.name + tag, flops, nbytes, call, flops_exec=fexec)
            elif op.kind == "site":
                if S_local == 0:
                    continue
                dd = self._drop_desc(op.site, getattr(op, "d_masks", None), B, sample0, seed, mask_offset,
                                     self.gather[op.dst.id] if op.dst.id in compact else None)
                per_image = op.src.H * op.src.W * op.src.C
                if op.dst.id in self.boundary:
                    bb = st["boundary"][op.dst.id]
                    nbytes = 2 * B * per_image * es + S_local * B * per_image // 8 + \
                        (S_local * B * per_image // 4 * es if bb["plane"] is not None else 0)
                    self._launch("dropout", op.name + " [keep bits + shortcut plane]", 0, nbytes, lambda: lib.bnn_boundary_bits(
                        _ptr(acts[op.src.id]), _ptr(bb["x_scaled"]), _ptr(bb["bits"]), _ptr(bb["plane"]), self.dcode, B,
                        op.src.H, op.src.W, op.src.C, S_local, ctypes.byref(dd), stream))
                    continue
                nbytes = B * per_image * es * ((S_local if op.src.stoch else 1) + S_local)
                if op.dst.id in compact:
                    nbytes = B * per_image * es * (S_local if op.src.stoch else 1) \
                        + B * S_local * op.src.H * op.src.W * self.gather[op.dst.id]["kc"] * es
                self._launch("dropout", op.name, 0, nbytes, lambda: lib.bnn_dropout(
                    _ptr(acts[op.src.id]), _ptr(acts[op.dst.id]), self.dcode, per_image, op.src.C, S_local,
                    int(op.src.stoch), ctypes.byref(dd), stream))
            elif op.kind == "affine":
                n_img = (S_local if op.dst.stoch else 1) * B
                if n_img == 0:
                    continue
                px = n_img * op.src.H * op.src.W
                self._launch("affine", op.name, 0, 2 * px * op.src.C * es, lambda: lib.bnn_channel_affine(
                    _ptr(acts[op.src.id]), _ptr(acts[op.dst.id]), _ptr(op.d_w), _ptr(op.d_b), self.dcode, px, op.src.C,
                    int(op.relu), stream))
            elif op.kind == "maxpool":
                n_img = (S_local if op.dst.stoch else 1) * B
                if n_img == 0:
                    continue
                nbytes = n_img * op.src.C * (op.src.H * op.src.W + op.dst.H * op.dst.W) * es
                self._launch("maxpool", op.name, 0, nbytes, lambda: lib.bnn_maxpool2d(
                    _ptr(acts[op.src.id]), _ptr(acts[op.dst.id]), self.dcode, n_img, op.src.H, op.src.W, op.src.C,
                    op.pool_k, stream))
            elif op.kind == "head":
                e = op.exit_index
                dd = self._drop_desc(op.site, getattr(op, "d_masks", None), B, sample0, seed, mask_offset)
                lo = st["logits"]
                # per-sample logits: one [S_local][B][C] slab per exit
                lo_e = lo[e][s_off:] if lo is not None else None
                hw = op.src.H * op.src.W
                nbytes = (S_local if op.src.stoch else 1) * B * hw * op.src.C * es + op.d_w.numel() * 4
                flops = 2 * S_local * B * op.src.C * g.n_classes
                if getattr(op, "head_tc", False) and S_local > 0:
                    a_ws, l_ws = st["head_ws"]
                    call = lambda: lib.bnn_exit_head_tc(
                        _ptr(acts[op.src.id]), self.dcode, int(op.src.stoch), B, S_local, hw, op.src.C, g.n_classes,
                        _ptr(op.d_w3), _ptr(op.d_b_pad), op.c_pad, op.head_with_lo, ctypes.byref(dd), _ptr(a_ws), _ptr(l_ws),
                        _ptr(sum_p[e]), _ptr(sum_l[e]), _ptr(sum_pl[e]), _ptr(lo_e), int(accumulate), stream)
                elif getattr(op, "head_rows", False) and S_local > 0:
                    l_ws = st["head_ws"][1]
                    call = lambda: lib.bnn_exit_head_rows(
                        _ptr(acts[op.src.id]), self.dcode, int(op.src.stoch), B, S_local, hw, op.src.C, g.n_classes,
                        _ptr(op.d_w_cf), _ptr(op.d_b), ctypes.byref(dd), _ptr(l_ws), _ptr(sum_p[e]), _ptr(sum_l[e]),
                        _ptr(sum_pl[e]), _ptr(lo_e), int(accumulate), stream)
                elif op.head_mma:
                    call = lambda: lib.bnn_exit_head_mma(
                        _ptr(acts[op.src.id]), self.dcode, int(op.src.stoch), B, S_local, hw, op.src.C, g.n_classes,
                        _ptr(op.d_w_hi), _ptr(op.d_w_lo), _ptr(op.d_b), ctypes.byref(dd), _ptr(sum_p[e]), _ptr(sum_l[e]),
                        _ptr(sum_pl[e]), _ptr(lo_e), int(accumulate), stream)
                else:
                    call = lambda: lib.bnn_exit_head(
                        _ptr(acts[op.src.id]), self.dcode, int(op.src.stoch), B, S_local, hw, op.src.C, g.n_classes,
                        _ptr(op.d_w), _ptr(op.d_b), ctypes.byref(dd), _ptr(sum_p[e]), _ptr(sum_l[e]), _ptr(sum_pl[e]),
                        _ptr(lo_e), int(accumulate), stream)
                self._launch("exit_head", op.name, flops, nbytes, call)
                if getattr(op, "head_tc", False) and S_local > 0:
                    self.launches += 2            # features + tcgen05 GEMM + soft-max / accumulate: three kernels
                elif getattr(op, "head_rows", False) and S_local > 0:
                    self.launches += 1            # logits rows + soft-max / accumulate: two kernels

    def profile_step(self, x, S_local, seed=0x5EED):
        """Device time of every launch of one step (CUDA events on the launching stream).
        -> list of {kernel, name, ms, flops (algorithmic), flops_exec (executed <= algorithmic), bytes (algorithmic)}
        in launch order."""
        x = x.to(self.device, torch.float32)
        with torch.cuda.device(self.device):
            self._prof = []
            try:
                st = self.enqueue(x, S_local, 0, seed)
                self.finalize(st, self._batch(x, S_local), max(S_local, 1))
                torch.cuda.synchronize(self.device)
                rec = self._prof
            finally:
                self._prof = None
        return [{"kernel": k, "name": n, "flops": f, "flops_exec": fx, "bytes": b, "ms": e0.elapsed_time(e1)}
                for k, n, f, b, e0, e1, fx in rec]

    def finalize(self, st, B, S_total):
        g = self.graph
        E, C = g.n_exits, g.n_classes
        n = E * B * C
        o = st["out"]
        views = [o[i * n:(i + 1) * n].view(E, B, C) for i in range(4)]
        ent = [o[4 * n + i * E * B: 4 * n + (i + 1) * E * B].view(E, B) for i in range(3)]
        sum_p, sum_l, sum_pl = self.sums_views(st, B)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if B > 0:
            _lib.check(self.lib.bnn_finalize(_ptr(sum_p), _ptr(sum_l), _ptr(sum_pl), E, B, C, S_total,
                                             _ptr(views[0]), _ptr(views[1]), _ptr(views[2]), _ptr(views[3]),
                                             _ptr(ent[0]), _ptr(ent[1]), _ptr(ent[2]), stream))
            self.launches += 2
        return views, ent

    def advance_masksembles(self, S):
        """Model the reference's stateful rotation (utils.py:168): every Masksembles module advances
        its `cnt` by one per forward pass."""
        seen = set()
        for s in self.graph.sites:
            if s.kind == "mask" and s.module is not None and id(s.module) not in seen:
                seen.add(id(s.module))
                s.module.cnt = (int(s.module.cnt) + S) % int(s.module.n)

    def _step(self, x, S, sample0, seed, want_logits, mask_offset, S_total, reduce_fn, gather_fn):
        """Everything one batch enqueues on the current stream: the launch sequence, the sample-sharding all-reduce of
        the sums (if any), the finaliser, the batch-sharding all-gather of the statistics (if any)."""
        st = self.enqueue(x, S, sample0, seed, False, want_logits, mask_offset)
        if reduce_fn is not None:
            reduce_fn(st["sums"])
        views, ent = self.finalize(st, self._batch(x, S), S_total)
        if gather_fn is not None:
            gather_fn(st["out"])
        return st, views, ent

    def _step_graphed(self, x, S, sample0, seed, want_logits, mask_offset, S_total, reduce_fn, gather_fn):
        """CUDA-graph replay of :meth:`_step`.  Kernel arguments (seed, sample range, Masksembles counters, buffer
        pointers) are baked into the graph, so graphs are cached per argument tuple: the first call with a tuple runs
        eagerly (this also initialises the NCCL communicator), the second captures, later ones replay (a benchmark or a
        sample-sharded serving loop repeats its tuple; an analysis loop that reseeds every batch simply stays eager).
        The NCCL collective is captured INSIDE the graph, right behind the kernels that produce its payload, so a step is
        one graph launch; if this torch / NCCL build refuses to capture it, the collectives are issued eagerly behind the
        replayed compute graph instead (decided once per engine)."""
        cnts = tuple(int(s.module.cnt) for s in self.graph.sites if s.kind == "mask")
        inside = self._graph_collectives and (reduce_fn is not None or gather_fn is not None)
        key = (self._batch(x, S), S, sample0, seed, want_logits, mask_offset, cnts, S_total, reduce_fn is not None,
               gather_fn is not None, inside)
        st = self._buffers(self._batch(x, S), S, want_logits)
        entry = self._graphs.get(key)
        if entry is None:
            self._graphs[key] = "seen"
            if len(self._graphs) > 64:
                self._graphs.pop(next(iter(self._graphs)))
            return self._step(x.to(self.device, non_blocking=True), S, sample0, seed, want_logits, mask_offset, S_total,
                              reduce_fn, gather_fn)
        st["x"].copy_(x, non_blocking=True)
        pinned_src = x.device.type == "cpu" and x.is_pinned()
        if pinned_src:
            # the caller may reuse its pinned buffer as soon as this call returns: the H2D copy (only) is waited for -
            # AFTER the graph has been launched behind it, so the device never sees the launch latency
            if self._h2d_done is None:
                self._h2d_done = torch.cuda.Event()
            self._h2d_done.record()
            if entry == "seen":
                self._h2d_done.synchronize()          # the capture below must not start with a copy in flight
        rf, gf = (reduce_fn, gather_fn) if inside else (None, None)
        if entry == "seen":
            n0 = self.launches
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    if rf is None and gf is None and (reduce_fn is not None):
                        # collectives stay outside: the graph ends in front of the reduction
                        out = (self.enqueue(st["x"], S, sample0, seed, False, want_logits, mask_offset), None, None)
                    else:
                        out = self._step(st["x"], S, sample0, seed, want_logits, mask_offset, S_total, rf, gf)
            except Exception as e:                      # noqa: BLE001 - capture of the NCCL call refused
                if not inside:
                    raise
                import warnings
                warnings.warn("CUDA-graph capture of the NCCL collective failed (%s): collectives run eagerly" % (e,))
                torch.cuda.synchronize(self.device)
                self._graph_collectives = False
                self._graphs.pop(key, None)
                self.launches = n0
                return self._step_graphed(x, S, sample0, seed, want_logits, mask_offset, S_total, reduce_fn, gather_fn)
            entry = self._graphs[key] = (g, self.launches - n0, out)
            self.launches = n0
        g, n, out = entry
        g.replay()
        if pinned_src:
            self._h2d_done.synchronize()
        self.launches += n
        st, views, ent = out
        if not inside:
            if reduce_fn is not None:                   # graph = launch sequence only; reduce, then finalise eagerly
                reduce_fn(st["sums"])
                views, ent = self.finalize(st, self._batch(x, S), S_total)
            if gather_fn is not None:
                gather_fn(st["out"])
        return st, views, ent

    def run(self, x, S, seed=0x5EED, sample0=0, S_total=None, want_logits=False, reduce_fn=None,
            mask_offset=None, use_graph=None, gather_fn=None):
        """S local samples starting at global sample index `sample0`; `reduce_fn(sums)` (optional)
        all-reduces the flat sums tensor across ranks before the finaliser, `gather_fn(out)` (optional) gathers the
        finished statistics (batch sharding).  Masksembles rows are (module.cnt + mask_offset + s) % n with
        mask_offset defaulting to sample0."""
        B = self._batch(x, S) if S > 0 else x.shape[0]
        if use_graph is None:
            use_graph = os.environ.get("BNN_CUDA_GRAPH", "1") != "0"
        graphed = use_graph and B > 0 and S > 0 and self._prof is None
        # a float32 HOST batch headed for a graph replay is copied straight into the graph's input buffer (one H2D, async
        # when the batch is pinned); everything else is moved / converted here
        if not (graphed and x.device.type == "cpu" and x.dtype == torch.float32 and x.is_contiguous()):
            x = x.to(self.device, torch.float32)
        S_total = S if S_total is None else S_total
        with torch.cuda.device(self.device):
            if graphed:
                st, views, ent = self._step_graphed(x, S, sample0, seed, want_logits, mask_offset, S_total, reduce_fn,
                                                    gather_fn)
            else:
                st, views, ent = self._step(x, S, sample0, seed, want_logits, mask_offset, S_total, reduce_fn, gather_fn)
        self.advance_masksembles(S_total)
        all_logits = None
        if want_logits:
            E, C = self.graph.n_exits, self.graph.n_classes
            all_logits = st["logits"].view(E, S, B, C).permute(1, 0, 2, 3)
        return MCResult(views[0], views[1], views[2], views[3], ent[0], ent[1], ent[2], all_logits, S_total, st["out"])
