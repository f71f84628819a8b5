"""CPU: host-side logic - graph lowering, BN folding, site fusion, sharding, the C-ABI library's
symbol table, and the no-CPU-fallback contract."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from bayesnn_fpga_b200 import _lib, engine, lenet, predict, resnet18, vgg19
from bayesnn_fpga_b200 import Dropouts, nn2bnn, utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "bnn_b200.h")).read()
    declared = set(re.findall(r"\b(bnn_[a-z0-9_]+)\s*\(", header))
    declared -= {"bnn_drop_desc"}
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.bnn_version() == 100


def test_no_cpu_fallback(lib):
    if torch.cuda.is_available():
        pytest.skip("CPU-only contract")
    # every compute entry point refuses to run without an sm_100 device
    assert lib.bnn_device_check() != 0
    buf = (ctypes.c_uint32 * 8)()
    assert lib.bnn_philox_words(ctypes.addressof(buf), 8, 1, 2, 3, None) != 0
    assert b"no CPU fallback" in lib.bnn_last_error()
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, out_dim=10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Dropouts.MCDropout(0.5)(torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        np.random.seed(0)
        utils.Masksembles1D(16, 4, 2.0).eval()(torch.zeros(4, 16))


def test_cost_model_matches_survey():
    # SURVEY.md section 8d, hook-derived MACs of the live reference
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", out_dim=10)
    assert m._bnn_graph().macs() == (152764416, 515919872)
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout=None, out_dim=10)
    pre, suf = m._bnn_graph().macs()
    assert pre + suf == 668684288 and suf == 4 * 5120
    v = vgg19.VGG19MCEarlyExit(dropout_exit=True, dropout=None, out_dim=100, n_exits=5).append_block_dropout((2, 3, 4))
    assert v._bnn_graph().macs() == (251854848, 174843904)


def test_site_numbering_and_fusion():
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", out_dim=10)
    g = m._bnn_graph()
    assert [s.name for s in g.sites] == ["layer1.1", "exit1_dropout", "layer2.1", "exit2_dropout", "layer3.1",
                                         "exit3_dropout", "exit_dropout"]
    assert [s.stream for s in g.sites] == list(range(7))
    n_before = len(g.ops)
    g.fuse_sites()
    fused = [op for op in g.ops if op.kind == "conv" and op.site is not None]
    # layer2 / layer3 sites ride in the epilogue of the block's last conv; the layer1 site is the
    # prefix -> suffix broadcast and stays a kernel of its own
    assert [op.site.name for op in fused] == ["layer2.1", "layer3.1"] and len(g.ops) == n_before - 2
    assert [op.site.name for op in g.ops if op.kind == "site"] == ["layer1.1"]
    # prefix = everything before the first site
    first = next(i for i, op in enumerate(g.ops) if op.kind == "site")
    assert all(not op.dst.stoch for op in g.ops[:first] if op.dst is not None)
    assert all(op.dst.stoch for op in g.ops[first:] if op.dst is not None)


def test_exit_only_dropout_keeps_all_convs_in_prefix():
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout=None, out_dim=10)
    g = m._bnn_graph()
    assert all(not op.dst.stoch for op in g.ops if op.kind == "conv")
    assert all(op.site is not None for op in g.ops if op.kind == "head")


def test_bn_fold_is_exact():
    torch.manual_seed(0)
    conv, bn = nn.Conv2d(5, 7, 3, padding=1, bias=True), nn.BatchNorm2d(7)
    bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5)
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.2)
    bn.eval()
    x = torch.randn(2, 5, 6, 6)
    w, b = engine.fold_conv_bn(conv, bn)
    assert torch.allclose(F.conv2d(x, w, b, 1, 1), bn(conv(x)), atol=1e-5)


def test_shard_samples_partitions():
    for S in (0, 1, 7, 32, 128):
        for G in (1, 2, 3, 4, 8):
            parts = [predict.shard_samples(S, G, r) for r in range(G)]
            assert sum(n for _, n in parts) == S
            assert parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(G - 1))
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
    with pytest.raises(ValueError):
        predict.shard_samples(8, 2, 2)


def test_converter_type_map_and_errors():
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.Conv2d(3, 4, 3)
            self.p = nn.MaxPool2d(2)
            self.f = nn.Sequential(nn.Flatten(), nn.Linear(4, 2), nn.ReLU())
            self.c3 = nn.Conv3d(1, 1, 1)
    bnn = nn2bnn.MCDropout(Net(), nSamples=5, p=0.25)
    assert type(bnn.model.c) is Dropouts.BayesianDropout2D
    assert type(bnn.model.p) is Dropouts.BayesianDropout
    assert type(bnn.model.f[1]) is Dropouts.BayesianDropout
    assert type(bnn.model.f[0]) is nn.Flatten and type(bnn.model.f[2]) is nn.ReLU
    assert type(bnn.model.c3) is Dropouts.BayesianDropout3D
    assert bnn.nSamples == 5 and bnn.p == 0.25 and bnn.model.c.p == 0.25
    assert "nSamples: 5" in bnn.extra_repr()
    with pytest.raises(ValueError, match="between 0 and 1"):
        Dropouts.BayesianDropout(nn.Linear(2, 2), p=1.5)          # Dropouts.py:15-17
    assert Dropouts.BayesianDropout(nn.Linear(2, 2), p=0.3).extra_repr() == "p=0.3, inplace=False"


def test_masksembles_module_contract():
    np.random.seed(0)
    m = utils.Masksembles2D(64, 4, 2.0)
    assert m.masks.shape == (4, 64) and not m.masks.requires_grad and m.cnt == 0
    assert "masks" in m.state_dict()
    assert set(np.unique(m.masks.numpy())) == {0.0, 1.0}
    assert [len(k) for k in utils.kept_channels(m.masks)] == [34] * 4          # 34/64 at scale 2
    m.train()
    with pytest.raises(ValueError, match="divisible by n"):
        m(torch.zeros(6, 64, 2, 2))                                          # utils.py:159-160
    assert m.extra_repr() == "scale=2.0, n=4"


def test_model_attributes_read_by_analyzers():
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", dropout_p=0.3, out_dim=10)
    assert (m.n_exits, m.out_dim, m.dropout, m.dropout_exit, m.dropout_p) == (4, 10, "block", True, 0.3)
    assert isinstance(m.layer1[1], Dropouts.MCDropout) and isinstance(m.exit1_dropout, Dropouts.MCDropout)
    with pytest.raises(NotImplementedError):
        np.random.seed(0)
        resnet18.ResNet18MCEarlyExit(dropout="layer", mask_type="mask")
    l = lenet.LeNetMCEarlyExit()
    g = l._bnn_graph()
    assert g.n_exits == 2 and g.n_classes == 10 and [s.name for s in g.sites] == ["bayes_2nd_exit", "bayes_1st_exit"]


# ---- generic lowering (the converter's output through the fused plan) ----------------
def _hook_macs(model, x):
    macs = []

    def hook(m, i, o):
        if isinstance(m, nn.Conv2d):
            macs.append(o[0].numel() * m.in_channels * m.kernel_size[0] * m.kernel_size[1])
        elif isinstance(m, nn.Linear):
            macs.append(m.in_features * m.out_features)
    hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, (nn.Conv2d, nn.Linear))]
    with torch.no_grad():
        model(x)
    for h in hs:
        h.remove()
    return sum(macs)


def test_lowering_residual_network_structure_and_macs():
    from bayesnn_fpga_b200 import lowering
    from tests.nets_generic import SmallResNet
    net = SmallResNet(lambda p: nn.Dropout(p)).eval()          # nn.Dropout is an eval-mode no-op: all prefix
    g, sites = lowering.lower_module(net, (3, 16, 16))
    assert not sites and g.n_exits == 2
    kinds = [(o.kind, o.name) for o in g.ops]
    # BN folded, ReLU / residual fused; the 1x1 shortcut is emitted before the conv that adds it; both global pools
    # disappear into exit heads
    assert kinds == [("conv", "stem"), ("conv", "b1_conv1"), ("conv", "b1_conv2"), ("head", "aux_fc"),
                     ("conv", "b2_conv1"), ("conv", "b2_down_0"), ("conv", "b2_conv2"), ("head", "fc")]
    by = {o.name: o for o in g.ops}
    assert by["b1_conv2"].res is by["stem"].dst and by["b2_conv2"].res is by["b2_down_0"].dst
    assert by["b2_conv2"].relu and not by["b2_down_0"].relu and by["aux_fc"].src is by["b1_conv2"].dst
    pre, suf = g.macs()
    assert suf == 0 and pre == _hook_macs(net, torch.zeros(1, 3, 16, 16))


def test_lowering_converted_network_prefix_suffix():
    from bayesnn_fpga_b200 import lowering
    from tests.nets_generic import plain_cnn
    bnn = nn2bnn.MCDropout(plain_cnn(), nSamples=4, p=0.25).reseed(7)
    g, sites = lowering.lower_module(bnn.model, (1, 28, 28))
    assert [s.stream for s, _ in sites] == list(range(6)) and [s.kind for s, _ in sites] == ["mc2d", "mc", "mc2d", "mc", "mc", "mc"]
    convs = [o for o in g.ops if o.kind == "conv"]
    assert all(o.relu for o in convs[:3]) and not convs[3].relu            # relu(drop(maxpool(drop(conv)))) commutes
    pre, suf = g.macs()
    assert pre == 28 * 28 * 8 * 25                                         # only the first conv is deterministic
    assert pre + suf == _hook_macs(plain_cnn(), torch.zeros(1, 1, 28, 28)) + 10 * 10   # + the identity output head
    g.fuse_sites()
    assert sum(o.kind == "site" for o in g.ops) == 3                       # boundary site + the two behind max-pools


def test_lowering_rejects_what_it_cannot_express():
    from bayesnn_fpga_b200 import lowering

    class Odd(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.Conv2d(3, 4, 3, padding=1)

        def forward(self, x):
            return torch.sigmoid(self.c(x)).mean((2, 3))
    with pytest.raises(NotImplementedError, match="sigmoid"):
        lowering.lower_module(Odd(), (3, 8, 8))
    with pytest.raises(NotImplementedError, match="kernel == stride"):
        lowering.lower_module(nn.Sequential(nn.Conv2d(3, 4, 3), nn.MaxPool2d(3, 2), nn.Flatten(), nn.Linear(4, 2)), (3, 9, 9))


def test_shortcut_fusion_pass_preserves_macs_and_removes_projection_convs():
    """Graph.fuse_shortcuts: every 1x1 stride-2 projection shortcut of the multi-exit ResNet becomes extra K of the
    convolution that adds it; MACs (results_analyzer.py:632-637 cost model) are unchanged."""
    from bayesnn_fpga_b200 import resnet18
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", dropout_p=0.5, out_dim=10)
    g = m._bnn_graph()
    g.fuse_sites()
    before = g.macs()
    n_ops = len(g.ops)
    g.fuse_shortcuts(lambda op: True)
    assert g.macs() == before
    fused = [o for o in g.ops if getattr(o, "sc", None) is not None]
    assert [o.name for o in fused] == ["layer2.0.0.conv2+downsample", "layer3.0.0.conv2+downsample",
                                       "layer4.0.conv2+downsample"]
    assert len(g.ops) == n_ops - 3 and all(o.res is None for o in fused)
    assert not any("downsample" == o.name.split(".")[-1] for o in g.ops)
    for o in fused:
        assert o.sc["src"].H == 2 * o.dst.H and o.sc["weight"].shape == (o.dst.C, o.sc["src"].C)
    # siblings: with the projections gone each group holds the exit branch's first conv + the next stage's first conv
    g.fuse_sibling_convs(lambda op: True)
    groups = [o for o in g.ops if o.kind == "convg"]
    assert [len(o.members) for o in groups] == [2, 2, 2]
    # a Masksembles-masked block input keeps its stand-alone projection (the per-mask weight path owns that tensor)
    np.random.seed(0)
    mm = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", out_dim=100, mask_type="mask", num_masks=4,
                                      mask_scale=2.0)
    gm = mm._bnn_graph()
    gm.fuse_sites()
    gm.fuse_shortcuts(lambda op: True)
    assert not any(getattr(o, "sc", None) is not None for o in gm.ops)


@pytest.mark.skipif(not os.path.isdir("/root/reference/Software_Artifact/software"),
                    reason="needs the reference checkout (build container only)")
def test_reference_model_objects_lower_to_the_same_graph_as_the_drop_in_classes():
    """lowering.lower_module applied to the UNMODIFIED reference model classes (their own MCDropout / Masksembles
    modules, F.relu / F.avg_pool2d / view forward) yields op for op the graph of this package's classes: same kinds,
    folded weights, ReLU flags, sites (kind, p, stream) and MACs - so `mc_predict(reference_model, x, S)` is the path
    the parity tests already cover."""
    import subprocess
    code = r'''
import sys
sys.path.insert(0, "%s"); sys.path.insert(0, "/root/reference/Software_Artifact/software")
import numpy as np, torch
from models.resnet18 import resnet18 as ref
from models.vgg19 import vgg19 as refv
from bayesnn_fpga_b200 import lowering, resnet18 as ours, vgg19 as oursv
def sig(g):
    return [(o.kind, o.relu, (o.site.kind, o.site.p, o.site.stream) if o.site else None,
             tuple(o.weight.shape) if o.weight is not None else None, o.stride, o.pad, o.pool_k) for o in g.ops]
def check(rm, om):
    om.load_state_dict(rm.state_dict())
    g1, _ = lowering.lower_module(rm.eval(), (3, 32, 32))
    g2 = om.eval()._bnn_graph()
    assert sig(g1) == sig(g2) and g1.macs() == g2.macs() and g1.out_order == list(range(g2.n_exits))
    for a, b in zip(g1.ops, g2.ops):
        if a.weight is not None:
            assert torch.equal(a.weight, b.weight) and torch.equal(a.bias, b.bias)
for kw in (dict(dropout_exit=True, dropout="block", dropout_p=0.5, out_dim=10),
           dict(dropout_exit=True, dropout=None, dropout_p=0.125, out_dim=100),
           dict(dropout_exit=True, dropout="block", out_dim=100, mask_type="mask", num_masks=4, mask_scale=2.0),
           dict(dropout_exit=True, dropout="layer", dropout_p=0.25, out_dim=10)):
    np.random.seed(0); rm = ref.ResNet18MCEarlyExit(**kw)
    np.random.seed(0); om = ours.ResNet18MCEarlyExit(**kw)
    check(rm, om)
np.random.seed(0)
check(ref.ResNet18MC(dropout_exit=True, dropout="block", dropout_p=0.375, out_dim=10),
      ours.ResNet18MC(dropout_exit=True, dropout="block", dropout_p=0.375, out_dim=10))
kw = dict(dropout_exit=True, dropout=None, dropout_p=0.25, out_dim=10, image_size=32, n_exits=5)
check(refv.VGG19MCEarlyExit(**kw), oursv.VGG19MCEarlyExit(**kw))
rm = refv.VGG19MCEarlyExit(**kw)
for i in (2, 3, 4):
    rm.blocks[i].append(refv.MCDropout(0.25))
check(rm, oursv.VGG19MCEarlyExit(**kw).append_block_dropout((2, 3, 4)))
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # a subprocess: the reference's top-level module names (`utils`, `models`) must not leak into this process
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_head_pool_fusion_pass():
    """Graph.fuse_head_pools: the convolutions whose only reader is an exit head write the pooled row; the one that
    shares its input with the next stage stays in the sibling group; MACs are unchanged."""
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", dropout_p=0.5, out_dim=10)
    g = m._bnn_graph()
    g.fuse_sites()
    before = g.macs()
    g.fuse_shortcuts(lambda op: True)
    g.fuse_head_pools(lambda op: True)
    pooled = [o.name for o in g.ops if getattr(o, "pool_from", None)]
    assert pooled == ["ex1conv3", "ex2conv2", "layer4.1.conv2"] and g.macs() == before
    for o in g.ops:
        if o.kind == "head" and o.name != "ex3linear":
            assert (o.src.H, o.src.W) == (1, 1)
    g.fuse_sibling_convs(lambda op: True)
    assert [len(o.members) for o in g.ops if o.kind == "convg"] == [2, 2, 2]


def test_keras_style_insertion_strategies_for_pytorch():
    """nn2bnn.convert_model mirrors the Keras converter's strategies (converter/keras/nn2bnn.py:9-72): position k =
    behind layer k; nothing in front of the first layer."""
    mk = lambda: nn.Sequential(nn.Conv2d(3, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(16, 32, 3, padding=1),
                               nn.ReLU(), nn.Flatten(), nn.Linear(32 * 4 * 4, 64), nn.ReLU(), nn.Linear(64, 10))
    layers = list(mk())
    assert nn2bnn._full_strategy(layers) == [-1, 2, 5, 7]
    assert nn2bnn._default_strategy(layers, 2) == [5, 7] and nn2bnn._default_strategy(layers, 9) == [-1, 2, 5, 7]
    assert nn2bnn._last_strategy(layers, 2) == [5, 4] and nn2bnn._last_strategy([nn.ReLU()], 1) == []
    sited = lambda m: [i for i, l in enumerate(m) if isinstance(l, nn.Sequential)]
    assert sited(nn2bnn.convert_model(mk(), "full")) == [3, 6, 8]          # never in front of the first layer
    assert sited(nn2bnn.convert_model(mk(), "default", num=1)) == [8]
    assert sited(nn2bnn.convert_model(mk(), "last", num=2)) == [5, 6]
    m = nn2bnn.convert_model(mk(), "default", num=2, p=0.3)
    assert isinstance(m[6][0], Dropouts.MCDropout) and m[6][0].p == 0.3 and isinstance(m[6][1], nn.Linear)
    np.random.seed(0)
    mm = nn2bnn.convert_model(mk(), "full", type="Masksembles", n=4, scale=2.0)
    assert isinstance(mm[3][0], utils.Masksembles2D) and mm[3][0].channels == 16
    assert isinstance(mm[6][0], utils.Masksembles1D) and mm[6][0].channels == 512
    with pytest.raises(ValueError):
        nn2bnn.convert_model(mk(), "last", num=2, type="Masksembles")      # in front of Flatten: width unknown
    with pytest.raises(ValueError):
        nn2bnn.convert_model(mk(), "nope")
    # the converted network lowers to a prefix / suffix plan: everything in front of the first site is deterministic
    from bayesnn_fpga_b200 import lowering
    g, sites = lowering.lower_module(nn2bnn.MCDropout(mk(), nSamples=4, p=0.25, strategy="default", num=2).model, (3, 8, 8))
    pre, suf = g.macs()
    assert len(sites) == 2 and suf == 512 * 64 + 64 * 10 and pre == 8 * 8 * 16 * 27 + 4 * 4 * 32 * 144


def test_get_network_factory(tmp_path):
    """model_loader.get_network (models/model_loader.py:8-23): factories by `call` / `resnet_type`, pickled models by
    `load_model`, AttributeError otherwise."""
    from bayesnn_fpga_b200 import model_loader
    hp = dict(call="ResNet18", load_model=None, resnet_type="mc_early_exit", dropout_exit=True, dropout="block",
              dropout_p=0.25, n_exits=4, out_dim=10)
    m = model_loader.get_network(hp)
    assert type(m).__name__ == "ResNet18MCEarlyExit" and m.dropout_p == 0.25 and m.out_dim == 10
    v = model_loader.get_network(dict(call="VGG19", load_model=None, resnet_type="mc_early_exit", dropout_exit=True,
                                      dropout=None, dropout_p=0.5, n_exits=5, out_dim=100, image_size=32))
    assert type(v).__name__ == "VGG19MCEarlyExit" and v.out_dim == 100
    # resnet18_loader.py:5-8: None builds the multi-exit class too (4 outputs), not ResNet18Base
    assert type(model_loader.get_network(dict(call="ResNet18", load_model=None, resnet_type=None, out_dim=10))).__name__ \
        == "ResNet18EarlyExit"
    # the reference's drivers always pass the FULL hyper-parameter dict (train/hyperparameters.py:86-109 + main.py's
    # mask keys): the plain / early-exit classes must be built with the six MC-only keys stripped
    full = dict(load_model=None, out_dim=100, dropout=None, dropout_exit=False, dropout_p=0.5, n_exits=4,
                mask_type="mc", num_masks=4, mask_scale=4.0)
    for call, rtype, want, n_out in (("ResNet18", "early_exit", "ResNet18EarlyExit", 4),
                                     ("ResNet18", None, "ResNet18EarlyExit", 4),
                                     ("VGG19", "early_exit", "VGG19EarlyExit", 5), ("VGG19", None, "VGG19", 1)):
        hp_full = dict(full, call=call, resnet_type=rtype, n_exits=n_out, image_size=32) if call == "VGG19" else \
            dict(full, call=call, resnet_type=rtype)
        net = model_loader.get_network(hp_full)
        assert type(net).__name__ == want and net.out_dim == 100
        assert net._bnn_graph().n_exits == n_out
    mc = model_loader.get_network(dict(full, call="ResNet18", resnet_type="mc", dropout_exit=True, n_exits=1))
    assert type(mc).__name__ == "ResNet18MC" and mc.dropout_exit
    with pytest.raises(AttributeError):
        model_loader.get_network(dict(call="LeNet", load_model=None, resnet_type=None))
    path = str(tmp_path / "m.pt")
    torch.save(m, path)
    again = model_loader.get_network(dict(call="ResNet18", load_model=path, resnet_type="mc_early_exit"))
    assert type(again).__name__ == "ResNet18MCEarlyExit"
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), again.state_dict().values()))


def test_drop_desc_struct_layout_matches_the_c_header(tmp_path):
    """ctypes `DropDesc` == `struct bnn_drop_desc` of include/bnn_b200.h as a C compiler lays it out (size and every
    field offset); the header itself must compile as plain C."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fields = [f[0] for f in _lib.DropDesc._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "bnn_b200.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(bnn_drop_desc));\n' +
                   "".join('  printf("%%zu\\n", offsetof(bnn_drop_desc, %s));\n' % f for f in fields) +
                   "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.DropDesc)
    assert out[1:] == [getattr(_lib.DropDesc, f).offset for f in fields]


def test_plans_live_outside_the_module_and_follow_the_parameters(tmp_path):
    """Plans (ctypes handles, CUDA graphs, device buffers) are kept in `_plans`, never in a module's __dict__: a model
    that has run still pickles (`torch.save(model)` is how the reference stores networks, model_loader.py:9-17) and
    deep-copies; and a plan is dropped as soon as the model it was built from changes."""
    import copy
    from bayesnn_fpga_b200 import _plans, resnet18
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", out_dim=10)
    plans = _plans.plans_for(m)
    plans["x"] = ctypes.CDLL(None).printf            # as un-picklable as an Engine (a ctypes function pointer)
    assert _plans.plans_for(m) is plans and "x" in _plans.plans_for(m)
    assert not any(k.startswith("_bnn") for k in m.__dict__)
    torch.save(m, str(tmp_path / "ran.pt"))
    m2 = copy.deepcopy(m)
    assert _plans.plans_for(m2) == {}                 # the copy plans for itself
    # every way the reference's drivers change a model invalidates the plan
    with torch.no_grad():
        m.layer1[0][0].conv1.weight.mul_(1.5)         # in-place update (an optimiser step)
    assert "x" not in _plans.plans_for(m)
    _plans.plans_for(m)["x"] = 1
    m.layer2.load_state_dict(m.layer2.state_dict())   # load_state_dict on a SUB-module
    assert "x" not in _plans.plans_for(m)
    _plans.plans_for(m)["x"] = 1
    m.bn1.running_mean.add_(0.1)                      # a buffer
    assert "x" not in _plans.plans_for(m)
    _plans.plans_for(m)["x"] = 1
    m.layer1.append(resnet18.MCDropout(0.1))          # structure: the reference idiom `blocks[i].append(MCDropout(p))`
    assert "x" not in _plans.plans_for(m)
    _plans.plans_for(m)["x"] = 1
    m.double()                                        # .to(dtype / device): new storage
    assert "x" not in _plans.plans_for(m)
    _plans.plans_for(m)["x"] = 1
    assert "x" in _plans.plans_for(m)                 # and an untouched model keeps its plans
    _plans.drop(m)
    assert _plans.plans_for(m) == {}


def test_converter_structure_matches_the_reference_converter():
    """`_convert_model` wraps exactly the leaves the reference's converter wraps, with the same wrapper class and p, in
    the same places (fixture frozen from the reference's own nn2bnn.py / Dropouts.py, make_golden_converter.py);
    `extra_repr` and the constructor's ValueError read the same."""
    from bayesnn_fpga_b200 import Dropouts, nn2bnn
    from tests.golden.make_golden_converter import structure
    from tests.nets_converter import NETS
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "converter.npz"))
    for tag, (make, in_shape, p, n_samples, seed) in NETS.items():
        w = nn2bnn.MCDropout(make(), nSamples=n_samples, p=p)
        assert structure(w.model) == [str(s) for s in z[tag + "/structure"]], tag
        assert w.extra_repr() == str(z[tag + "/repr"][0])
    for bad in (-0.1, 1.5):
        with pytest.raises(ValueError) as e:
            Dropouts.BayesianDropout(nn.Linear(2, 2), bad)
        assert str(e.value) == str(z["err/%g" % bad][0])
    assert Dropouts.BayesianDropout2D(nn.Conv2d(1, 1, 1), 0.25).extra_repr() == str(z["extra_repr"][0])


def test_q8_quantizers_follow_qkeras_definitions():
    """quantized_bits(8, ibit) / quantized_relu(8, ibit) (qkeras/quantizers.py; used by the reference's 8-bit models,
    t_qmodels_bayes_me.py:49-52): power-of-two steps, round half to even, saturating ranges - product == oracle."""
    from bayesnn_fpga_b200 import q8
    from oracle import q8 as oq8
    x = torch.tensor([-2.0, -1.0, -0.996, -0.00390625, -0.001953125, 0.0, 0.001953125, 0.005859375, 0.5, 0.99, 1.0, 3.0])
    q, step = q8.quantized_bits(x, 8, 0)
    assert step == 2.0 ** -7 and q.tolist() == [-128, -128, -127, 0, 0, 0, 0, 1, 64, 127, 127, 127]
    qo, so = oq8.quantized_bits_int(x.numpy(), 8, 0)
    assert so == step and qo.tolist() == q.tolist()
    q, step = q8.quantized_relu(x, 8, 0)
    assert step == 2.0 ** -8 and q.tolist() == [0, 0, 0, 0, 0, 0, 0, 2, 128, 253, 255, 255]      # 0.5 LSB -> even
    qo, so = oq8.quantized_relu_int(x.numpy(), 8, 0)
    assert so == step and qo.tolist() == q.tolist()
    q, step = q8.quantized_bits(x, 8, 2)                     # ibit = 2: range [-4, 4), step 2^-5
    assert step == 2.0 ** -5 and q.tolist()[0] == -64 and q.tolist()[-1] == 96


def test_bench_keeps_stdout_for_the_one_json_line():
    """bench.py's contract is ONE JSON line on stdout; whatever else writes to fd 1 (NCCL's version banner, stray prints)
    is sent to stderr once bench.claim_stdout() ran."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import bench, os; bench.claim_stdout(); os.write(1, b'NCCL version 2.x\\n'); print('chatter', flush=True); "
            "bench.emit({'ok': 1})")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, cwd=root, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stdout == b'{"ok": 1}\n'
    assert b"NCCL version" in r.stderr and b"chatter" in r.stderr


def test_position_major_tap_fraction_mirrors_the_kernel_rule(monkeypatch):
    """engine._pm_tap_fraction (EXECUTED / algorithmic multiply-adds under position-major tiling, what bench.py's
    roofline counts) follows conv_tc.cu's host rule: 3x3 kernels, 2x2 .. 8x8 output maps, 256-channel tiles, >= 1024
    images, and only where at least a tenth of the taps vanish."""
    from bayesnn_fpga_b200.engine import _pm_tap_fraction as frac
    monkeypatch.delenv("BNN_TC_NO_PM", raising=False)
    monkeypatch.delenv("BNN_TC_PM_MIN_IMAGES", raising=False)
    assert frac(8192, 4, 4, 4, 4, 1, 512) == pytest.approx(25 / 36)
    assert frac(8192, 8, 8, 8, 8, 1, 256) == pytest.approx(121 / 144)
    assert frac(8192, 2, 2, 2, 2, 1, 512) == pytest.approx(16 / 36)
    assert frac(8192, 8, 8, 4, 4, 2, 512) == pytest.approx(121 / 144)          # stride 2, 8x8 -> 4x4: 11 of 12 per axis
    assert frac(8192, 16, 16, 8, 8, 2, 256) == 1.0                              # saves 8 % only: pixel-major tiles
    assert frac(8192, 16, 16, 16, 16, 1, 256) == 1.0 and frac(8192, 4, 4, 4, 4, 1, 128) == 1.0
    assert frac(512, 4, 4, 4, 4, 1, 512) == 1.0                                 # below the batch threshold
    monkeypatch.setenv("BNN_TC_NO_PM", "1")
    assert frac(8192, 4, 4, 4, 4, 1, 512) == 1.0
