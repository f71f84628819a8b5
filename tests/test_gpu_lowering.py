"""GPU: arbitrary nn.Modules through the fused plan (lowering.lower_module) - SURVEY.md 8(f) rank 2.

Oracle: the same architecture on the CPU in float64 with the oracle's Philox masks injected at every stochastic
module (tests/nets_generic.py), i.e. the converter's eval semantics `mean over nSamples passes of the raw output`
(Hardware_Artifact/converter/pytorch/nn2bnn.py:26-27).  Tolerances: fp32 path 1e-5 * max(1, |out|), fp16 tensor-core
path 2e-3 * max(1, |out|) on raw outputs (logits).
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

from bayesnn_fpga_b200 import _plans, mc_predict, nn2bnn
from bayesnn_fpga_b200.Dropouts import BayesianDropout, BayesianDropout2D, MCDropout
from tests.gpu_util import report
from tests.nets_generic import InjectedDropout, SmallResNet, plain_cnn, randomize_bn, reference_mean

pytestmark = pytest.mark.gpu


def _converted_reference(bnn, seed):
    """CPU twin of a converted nn.Sequential: every wrapper becomes layer -> InjectedDropout with its stream."""
    layers, sites = [], []
    for m in bnn.model:
        if isinstance(m, (BayesianDropout, BayesianDropout2D)):
            d = InjectedDropout(m.p, m.bnn_stream, seed, "channel" if isinstance(m, BayesianDropout2D) else "element")
            layers += [copy.deepcopy(m.layer).cpu(), d]
            sites.append(d)
        else:
            layers.append(copy.deepcopy(m).cpu())
    return nn.Sequential(*layers), sites


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-5), ("fp16", 2e-3)])
def test_converted_cnn_fused_plan_vs_injected_reference(dtype, tol):
    torch.manual_seed(0)
    S, seed, B = 6, 77, 5
    bnn = nn2bnn.MCDropout(plain_cnn(), nSamples=S, p=0.25, dtype=dtype).reseed(seed).cuda().eval()
    ref, sites = _converted_reference(bnn, seed)
    x = torch.randn(B, 1, 28, 28)
    want = reference_mean(ref, x, sites, S)[0]
    got = bnn(x.cuda())
    assert bnn._plan((1, 28, 28), x.cuda().device) is not None            # the fused plan ran, not the fallback
    scale = max(1.0, want.abs().max().item())
    err = (got.double().cpu() - want).abs().max().item()
    report(test="converted_cnn_fused", dtype=dtype, err=err, scale=scale)
    assert err <= tol * scale
    # a second call draws the NEXT nSamples samples, like further stand-alone forward calls would
    want2 = reference_mean(ref, x, sites, S, sample0=S)[0]
    err2 = (bnn(x.cuda()).double().cpu() - want2).abs().max().item()
    assert err2 <= tol * scale and (want2 - want).abs().max().item() > 100 * tol


def test_converted_cnn_fused_equals_standalone_passes():
    """same masks, same arithmetic up to rounding: fused plan vs nSamples stand-alone passes (torch layers +
    bnn_dropout)."""
    torch.manual_seed(1)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        net = plain_cnn().cuda()
        a = nn2bnn.MCDropout(copy.deepcopy(net), nSamples=5, p=0.5).reseed(9).eval()
        b = nn2bnn.MCDropout(copy.deepcopy(net), nSamples=5, p=0.5, fused=False).reseed(9).eval()
        x = torch.randn(4, 1, 28, 28, device="cuda")
        ya, yb = a(x), b(x)
        assert (ya - yb).abs().max().item() <= 1e-5 * max(1.0, yb.abs().max().item())
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-5), ("fp16", 2e-3)])
def test_residual_network_with_sites_vs_injected_reference(dtype, tol):
    """BatchNorm folding, residual adds traced shortcut-last, global-pool heads, two outputs, prefix/suffix split."""
    torch.manual_seed(2)
    S, seed, B = 5, 0xBEEF, 6
    model = randomize_bn(SmallResNet(lambda p: MCDropout(p))).cuda().eval()
    streams = [m.bnn_stream for m in (model.d1, model.d2, model.d3)]
    ref_sites = []

    def drop(p):
        d = InjectedDropout(p, streams[len(ref_sites)], seed)
        ref_sites.append(d)
        return d
    ref = SmallResNet(drop)
    ref.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    x = torch.randn(B, 3, 16, 16)
    want = reference_mean(ref, x, ref_sites, S)
    r = mc_predict(model, x.cuda(), S, seed=seed, dtype=dtype)
    eng = _plans.plans_for(model)[("generic", (3, 16, 16), dtype, ())]
    pre, suf = eng.graph.macs()
    assert pre > 0 and suf > 0 and eng.graph.out_order == [0, 1]
    for e, w in enumerate(want):
        scale = max(1.0, w.abs().max().item())
        err = (r.mean_logits[e].double().cpu() - w).abs().max().item()
        report(test="generic_resnet", dtype=dtype, exit=e, err=err, scale=scale)
        assert err <= tol * scale
    # softmax statistics come with it
    assert torch.allclose(r.mean_probs.sum(-1), torch.ones_like(r.mean_probs.sum(-1)), atol=1e-4)


def test_unsupported_network_raises_unless_the_eager_loop_is_requested():
    """No silent library path: a network the lowering cannot express raises; `eager_fallback=True` opts in to the
    reference's literal loop (stand-alone passes through torch's own layers) with a warning."""
    mk = lambda: nn.Sequential(nn.Linear(6, 8), nn.Sigmoid(), nn.Linear(8, 3)).cuda()
    x = torch.randn(5, 6, device="cuda")
    with pytest.raises(NotImplementedError, match="eager_fallback=True"):
        nn2bnn.MCDropout(mk(), nSamples=3, p=0.5).reseed(4).eval()(x)
    bnn = nn2bnn.MCDropout(mk(), nSamples=3, p=0.5, eager_fallback=True).reseed(4).eval()
    with pytest.warns(UserWarning, match="stand-alone passes"):
        y = bnn(x)
    assert y.shape == (5, 3) and torch.isfinite(y).all()


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-5), ("fp16", 2e-3)])
def test_reference_style_model_object_runs_unmodified(dtype, tol):
    """A model written like the reference's own classes (its MCDropout = nn.Dropout subclass calling F.dropout,
    `out += residual`, F.relu on non-negative tensors, F.avg_pool2d + view heads, side-effect attribute) goes through
    mc_predict as is; sites are numbered in first-use order like the named models."""
    from tests.nets_generic import MCDropout as RefMCDropout, RefStyleNet
    torch.manual_seed(4)
    S, seed, B = 4, 0x5EED, 5
    model = randomize_bn(RefStyleNet(lambda p: RefMCDropout(p))).cuda().eval()
    ref_sites = []

    def drop(p):
        d = InjectedDropout(p, len(ref_sites), seed)
        ref_sites.append(d)
        return d
    ref = RefStyleNet(drop)
    # module construction order: layer1 dropout (first used), exit1_dropout, exit_dropout == first-use order
    ref.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    x = torch.randn(B, 3, 16, 16)
    want = reference_mean(ref, x, ref_sites, S)
    r = mc_predict(model, x.cuda(), S, seed=seed, dtype=dtype)
    eng = _plans.plans_for(model)[("generic", (3, 16, 16), dtype, ())]
    assert [s.stream for s in eng.graph.sites] == [0, 1, 2] and eng.graph.n_exits == 2
    assert sum(o.kind == "head" for o in eng.graph.ops) == 2 and not any(getattr(o, "is_gap", False) for o in eng.graph.ops)
    for e, w in enumerate(want):
        scale = max(1.0, w.abs().max().item())
        err = (r.mean_logits[e].double().cpu() - w).abs().max().item()
        report(test="reference_style_model", dtype=dtype, exit=e, err=err, scale=scale)
        assert err <= tol * scale


def test_keras_style_strategy_conversion_runs_fused():
    """nn2bnn.MCDropout(model, strategy="default", num=2): dropout in front of the last two Linear layers only - the
    convolutional prefix runs once per image; result == float64 reference with the same injected masks."""
    torch.manual_seed(7)
    S, seed, B = 6, 31, 4
    net = nn.Sequential(nn.Conv2d(3, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(16, 32, 3, padding=1), nn.ReLU(),
                        nn.Flatten(), nn.Linear(32 * 4 * 4, 64), nn.ReLU(), nn.Linear(64, 10))
    bnn = nn2bnn.MCDropout(copy.deepcopy(net), nSamples=S, p=0.25, strategy="default", num=2).reseed(seed).cuda().eval()
    layers, sites = [], []
    for m in bnn.model:
        if isinstance(m, nn.Sequential):
            d = InjectedDropout(m[0].p, m[0].bnn_stream, seed)
            sites.append(d)
            layers += [d, copy.deepcopy(m[1]).cpu()]
        else:
            layers.append(copy.deepcopy(m).cpu())
    assert len(sites) == 2
    x = torch.randn(B, 3, 8, 8)
    want = reference_mean(nn.Sequential(*layers), x, sites, S)[0]
    got = bnn(x.cuda())
    eng, _ = bnn._plan((3, 8, 8), x.cuda().device)
    pre, suf = eng.graph.macs()
    assert suf == 512 * 64 + 64 * 10 and pre == 8 * 8 * 16 * 27 + 4 * 4 * 32 * 144   # the convolutions are paid once, not S times
    assert (got.double().cpu() - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("tag", ["mlp", "cnn", "nested"])
def test_converter_matches_the_reference_converter_fixture(tag):
    """nn2bnn.MCDropout (fused plan AND stand-alone passes) vs outputs frozen from the reference's OWN converter modules
    (`Hardware_Artifact/converter/pytorch/{nn2bnn,Dropouts}.py` imported in the build container with the Philox masks
    injected, tests/golden/make_golden_converter.py): per-pass outputs and the mean over nSamples."""
    import os
    from oracle import seeded
    from tests.cases import GOLDEN
    from tests.nets_converter import NETS
    z = np.load(os.path.join(GOLDEN, "converter.npz"))
    make, in_shape, p, n_samples, seed = NETS[tag]
    x = torch.from_numpy(z[tag + "/x"]).cuda()
    want_mean, want_passes = z[tag + "/mean"], z[tag + "/passes"]
    scale = max(1.0, float(np.abs(want_passes).max()))

    def build(**kw):
        torch.manual_seed(0)
        net = make()
        net.load_state_dict(seeded.seeded_state_dict(net.state_dict(), seed=77))
        return nn2bnn.MCDropout(net.cuda(), nSamples=n_samples, p=p, **kw).reseed(seed).eval()
    errs = {}
    for dtype, tol in (("fp32", 1e-5), ("fp16", 2e-3)):
        got = build(dtype=dtype)(x)
        errs[dtype] = float(np.abs(got.double().cpu().numpy() - want_mean).max())
        assert errs[dtype] <= tol * scale, (tag, dtype, errs)
    # the reference's literal loop: nSamples stand-alone passes through the wrapped modules
    loop = build(fused=False)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # the leaves run in torch here
    try:
        passes = np.stack([loop.model(x).double().cpu().numpy() for _ in range(n_samples)])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    errs["standalone_passes"] = float(np.abs(passes - want_passes).max())
    assert errs["standalone_passes"] <= 1e-5 * scale
    report(test="converter_vs_reference_fixture", net=tag, scale=scale, **errs)


@pytest.mark.parametrize("kind,classes", [("resnet_mcd", 10), ("resnet_mask", 100), ("vgg_last3", 10)])
def test_unmodified_reference_model_object_is_a_drop_in(kind, classes):
    """The reference's OWN model object - built by its `models.model_loader.get_network` from the staged tree
    (oracle/_ref: its ResNet18MCEarlyExit / VGG19MCEarlyExit classes, its MCDropout / Masksembles layers) - handed to
    `mc_predict` as is: lowering.lower_module traces it to the same plan as this package's class of the same name, so the
    predictive means are IDENTICAL to the drop-in class holding the same state_dict (whose parity with the oracle the
    other GPU tests establish).  BASELINE configs 2, 3 and 4."""
    from oracle import ref_arm
    if not ref_arm.available():
        pytest.skip("oracle/_ref is not staged (python oracle/make_ref.py in the build container)")
    from bayesnn_fpga_b200 import model_loader
    from bayesnn_fpga_b200.Dropouts import MCDropout as OurMCDropout
    ref_model = ref_arm.build(kind, classes)
    assert type(ref_model).__module__.startswith("models.")          # really the reference's class
    base = dict(load_model=None, out_dim=classes, dropout_exit=True, dropout_p=0.5, mask_type="mc", num_masks=4, mask_scale=4.0)
    np.random.seed(0)                                                # Masksembles masks: same generator calls as ref_arm.build
    if kind == "resnet_mcd":
        ours = model_loader.get_network(dict(base, call="ResNet18", resnet_type="mc_early_exit", dropout="block", n_exits=4))
    elif kind == "resnet_mask":
        ours = model_loader.get_network(dict(base, call="ResNet18", resnet_type="mc_early_exit", dropout="block", n_exits=4,
                                             mask_type="mask", mask_scale=2.0))
    else:
        ours = model_loader.get_network(dict(base, call="VGG19", resnet_type="mc_early_exit", dropout=None, n_exits=5,
                                             image_size=32))
        for i in (2, 3, 4):
            ours.blocks[i].append(OurMCDropout(0.5))
    ours.load_state_dict(ref_model.state_dict())
    B, S = 24, 4
    x = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    a = mc_predict(ref_model.cuda().eval(), x.cuda(), S, seed=77, dtype="fp16")
    a_probs = [p.clone() for p in a.mean_probs]
    b = mc_predict(ours.cuda().eval(), x.cuda(), S, seed=77, dtype="fp16")
    assert len(a_probs) == len(b.mean_probs) == (5 if kind == "vgg_last3" else 4)
    for e, (pa, pb) in enumerate(zip(a_probs, b.mean_probs)):
        err = (pa - pb).abs().max().item()
        report(test="reference_object_drop_in", kind=kind, exit=e, err=err)
        assert torch.isfinite(pa).all() and abs(pa.sum(1) - 1).max().item() < 1e-3
        assert err == 0.0
