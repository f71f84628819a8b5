"""GPU: each C-ABI kernel against a plain fp32/fp64 restatement of the same op."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import philox, stats
from tests.gpu_util import TORCH_DT, drop_desc, report, stream

pytestmark = pytest.mark.gpu


def _conv_case(lib, fn, dt, N, H, W, Cin, Cout, k, stride, pad, relu, with_res, drop=None, seed=0):
    tdt, code = TORCH_DT[dt]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / np.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g)
    xq, wq = x.to(tdt).float(), (w.to(tdt).float() if fn == "tc" else w)
    want = F.conv2d(xq.double(), wq.double(), b.double(), stride, pad)
    res = None
    if with_res:
        res = torch.randn(want.shape, generator=g).to(tdt)
        want = want + res.double()
    if relu:
        want = want.relu()
    OH, OW = want.shape[2:]
    d_x = xq.permute(0, 2, 3, 1).contiguous().to(tdt).cuda()
    d_w = wq.permute(0, 2, 3, 1).contiguous().cuda()
    d_w = d_w.to(tdt) if fn == "tc" else d_w
    d_b = b.cuda()
    d_res = res.permute(0, 2, 3, 1).contiguous().cuda() if with_res else None
    d_y = torch.full((N, OH, OW, Cout), float("nan"), dtype=tdt, device="cuda")
    dd = drop if drop is not None else drop_desc(batch=N)
    rp = ctypes.c_void_p(d_res.data_ptr() if with_res else 0)
    if fn == "tc":
        rc = lib.bnn_conv2d_tc(d_x.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), rp, d_y.data_ptr(), code, N, H, W, Cin,
                               Cout, k, stride, int(relu), ctypes.byref(dd), stream())
    else:
        rc = lib.bnn_conv2d_simt(d_x.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), rp, d_y.data_ptr(), code, N, H, W,
                                 Cin, Cout, k, k, stride, pad, int(relu), ctypes.byref(dd), stream())
    assert rc == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    return d_y.cpu().float().permute(0, 3, 1, 2), want


SIMT_SHAPES = [
    # N, H, W, Cin, Cout, k, stride, pad, relu, res
    (3, 9, 7, 5, 11, 3, 1, 1, True, True),        # ragged everything
    (2, 32, 32, 3, 64, 3, 1, 1, False, False),    # conv1 of the ResNet
    (2, 14, 14, 20, 20, 5, 7, 0, True, False),    # LeNet second-exit conv ("same" at stride 7)
    (4, 28, 28, 1, 20, 5, 1, 2, True, False),     # LeNet conv2d_1
    (5, 2, 2, 20, 100, 2, 1, 0, True, False),     # an nn.Linear over a flattened 2x2 map
    (2, 16, 16, 64, 128, 1, 2, 0, False, False),  # 1x1 stride-2 shortcut
    (1, 8, 8, 128, 70, 3, 2, 1, True, True),      # stride 2, Cout not a multiple of the tile
    (130, 1, 1, 100, 10, 1, 1, 0, False, False),  # more than two tiles of 1x1 images
]


@pytest.mark.parametrize("shape", SIMT_SHAPES)
def test_conv_simt_fp32(lib, shape):
    got, want = _conv_case(lib, "simt", "fp32", *shape)
    err = (got.double() - want).abs().max().item()
    report(test="conv_simt_fp32", shape=list(shape), max_err=err)
    assert err <= 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
def test_conv_simt_16bit_storage(lib, dt):
    got, want = _conv_case(lib, "simt", dt, 2, 10, 10, 24, 40, 3, 1, 1, True, True)
    tol = 2e-3 if dt == "fp16" else 1.6e-2            # one rounding of the output to 16-bit storage
    assert (got.double() - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())


def test_conv_simt_fused_dropout(lib):
    N_img, B, S = 6, 2, 3
    dd = drop_desc(1, 0.5, 0x77, 4, 5, B)
    got, want = _conv_case(lib, "simt", "fp32", N_img, 6, 6, 8, 16, 3, 1, 1, True, False, drop=dd)
    for s in range(S):
        keep = torch.from_numpy(philox.keep_mask(0x77, 4, 5 + s, (B, 16, 6, 6), 0.5))
        ref = want[s * B:(s + 1) * B] * keep * 2.0
        assert (got[s * B:(s + 1) * B].double() - ref).abs().max().item() <= 2e-5


def test_conv_empty_and_bad_args(lib):
    y = torch.zeros(1, device="cuda")
    dd = drop_desc(batch=1)
    assert lib.bnn_conv2d_simt(y.data_ptr(), y.data_ptr(), y.data_ptr(), None, y.data_ptr(), 0, 0, 4, 4, 1, 1, 3, 3, 1,
                               1, 0, ctypes.byref(dd), stream()) == 0                 # empty batch: no-op
    assert lib.bnn_conv2d_simt(y.data_ptr(), y.data_ptr(), y.data_ptr(), None, y.data_ptr(), 0, 1, 2, 2, 1, 1, 5, 5, 1,
                               0, 0, ctypes.byref(dd), stream()) == -1                # empty output
    assert lib.bnn_conv2d_simt(y.data_ptr(), y.data_ptr(), y.data_ptr(), None, y.data_ptr(), 9, 1, 4, 4, 1, 1, 3, 3, 1,
                               1, 0, ctypes.byref(dd), stream()) == -1                # bad dtype


@pytest.mark.parametrize("dt", ["fp32", "fp16"])
@pytest.mark.parametrize("k", [2, 7])
@pytest.mark.parametrize("C", [5, 64])           # scalar path / 8-channel vector path
def test_maxpool(lib, dt, k, C):
    tdt, code = TORCH_DT[dt]
    x = torch.randn(3, C, 14, 14).to(tdt)
    d_x = x.permute(0, 2, 3, 1).contiguous().cuda()
    d_y = torch.empty(3, 14 // k, 14 // k, C, dtype=tdt, device="cuda")
    assert lib.bnn_maxpool2d(d_x.data_ptr(), d_y.data_ptr(), code, 3, 14, 14, C, k, stream()) == 0
    assert torch.equal(d_y.cpu().permute(0, 3, 1, 2), F.max_pool2d(x.float(), k, k).to(tdt))


def test_nchw_to_nhwc(lib):
    x = torch.randn(3, 5, 4, 6, device="cuda")
    for dt in ("fp32", "fp16", "bf16"):
        tdt, code = TORCH_DT[dt]
        y = torch.empty(3, 4, 6, 5, dtype=tdt, device="cuda")
        assert lib.bnn_nchw_to_nhwc(x.data_ptr(), y.data_ptr(), code, 3, 5, 4, 6, stream()) == 0
        assert torch.equal(y, x.permute(0, 2, 3, 1).to(tdt))


@pytest.mark.parametrize("C,F_,HW,has_samples,kind", [(10, 512, 16, 1, 1), (100, 512, 4, 0, 1), (7, 100, 1, 1, 0),
                                                      (100, 512, 1, 1, 3), (33, 48, 4, 1, 2)])
def test_exit_head_and_finalize(lib, C, F_, HW, has_samples, kind):
    """pool -> site -> linear -> softmax -> sums over samples, then the finaliser, vs float64 numpy."""
    B, S, s0, p, seed, sid = 5, 19, 3, 0.25, 0x99, 2          # S > HEAD_SCHUNK exercises the chunk loop
    g = torch.Generator().manual_seed(C)
    feat = torch.randn((S if has_samples else 1) * B, HW, F_, generator=g).abs()
    w = torch.randn(C, F_, generator=g) / np.sqrt(F_) * 3
    bias = torch.randn(C, generator=g)
    masks = (torch.rand(4, F_, generator=g) > 0.5).float()
    d_masks = masks.cuda()
    dd = drop_desc(kind, p, seed, sid, s0, B, d_masks if kind == 3 else None, cnt0=1)
    d = dict(feat=feat.cuda(), w=w.t().contiguous().cuda(), b=bias.cuda(), sp=torch.zeros(B, C).cuda(), sl=torch.zeros(B, C).cuda(),
             spl=torch.zeros(B).cuda(), lo=torch.zeros(S, B, C).cuda())
    rc = lib.bnn_exit_head(d["feat"].data_ptr(), 0, has_samples, B, S, HW, F_, C, d["w"].data_ptr(), d["b"].data_ptr(),
                           ctypes.byref(dd), d["sp"].data_ptr(), d["sl"].data_ptr(), d["spl"].data_ptr(),
                           d["lo"].data_ptr(), 0, stream())
    assert rc == 0, lib.bnn_last_error()
    fv = feat.double().view(-1, B, HW, F_).mean(2)                       # [S or 1, B, F]
    logits = np.zeros((S, B, C))
    for s in range(S):
        v = fv[s if has_samples else 0].numpy()
        if kind in (1, 2):
            v = v * philox.keep_mask(seed, sid, s0 + s, (B, F_), p) * (1 / (1 - p))
        elif kind == 3:
            v = v * masks[(1 + s0 + s) % 4].double().numpy()
        logits[s] = v @ w.double().numpy().T + bias.double().numpy()
    z = logits - logits.max(-1, keepdims=True)
    probs = np.exp(z) / np.exp(z).sum(-1, keepdims=True)
    assert np.abs(d["lo"].cpu().numpy() - logits).max() < 2e-5
    assert np.abs(d["sp"].cpu().numpy() - probs.sum(0)).max() < 2e-5
    assert np.abs(d["sl"].cpu().numpy() - logits.sum(0)).max() < 2e-4
    assert np.abs(d["spl"].cpu().numpy() - (probs * np.log(probs)).sum(-1).sum(0)).max() < 2e-4

    # accumulate=1 adds a second shard on top
    rc = lib.bnn_exit_head(d["feat"].data_ptr(), 0, has_samples, B, S, HW, F_, C, d["w"].data_ptr(), d["b"].data_ptr(),
                           ctypes.byref(dd), d["sp"].data_ptr(), d["sl"].data_ptr(), d["spl"].data_ptr(), None, 1, stream())
    assert rc == 0
    assert np.abs(d["sp"].cpu().numpy() - 2 * probs.sum(0)).max() < 4e-5

    # finaliser over E = 2 "exits" (the same sums twice, second one scaled) and 2*S samples
    E = 2
    sums_p = torch.stack([d["sp"], d["sp"]]).contiguous()
    sums_l = torch.stack([d["sl"], 0.5 * d["sl"]]).contiguous()
    sums_pl = torch.stack([d["spl"], d["spl"]]).contiguous()
    outs = [torch.zeros(E, B, C, device="cuda") for _ in range(4)] + [torch.zeros(E, B, device="cuda") for _ in range(3)]
    assert lib.bnn_finalize(sums_p.data_ptr(), sums_l.data_ptr(), sums_pl.data_ptr(), E, B, C, 2 * S,
                            *[o.data_ptr() for o in outs], stream()) == 0
    mp = np.stack([probs.mean(0), probs.mean(0)])
    ml = np.stack([logits.mean(0), 0.5 * logits.mean(0)])
    got = [o.cpu().numpy() for o in outs]
    assert np.abs(got[0] - mp).max() < 1e-5 and np.abs(got[1] - ml).max() < 1e-5
    assert np.abs(got[2][1] - mp.mean(0)).max() < 1e-5 and np.abs(got[3][1] - ml.mean(0)).max() < 1e-5
    assert np.abs(got[4][0] - stats.entropy_per_image(mp[0])).max() < 1e-5
    assert np.abs(got[6][0] - (-(probs * np.log(probs)).sum(-1).mean(0))).max() < 1e-5


def test_calibration_bins_and_ece(lib):
    from bayesnn_fpga_b200.results_analyzer import FullAnalysis

    class Dummy:
        n_exits, out_dim = 1, 10
    fa = FullAnalysis(Dummy(), None, run=False)
    z = np.load(__import__("tests.cases", fromlist=["GOLDEN"]).GOLDEN + "/ece_hist.npz")
    for k in range(3):
        p, lab = z["p%d" % k], z["label%d" % k]
        onehot = np.eye(p.shape[1])[lab]
        got = fa.ece_hist_binary(p, onehot)
        report(test="ece_hist_binary", case=k, diff=abs(got - float(z["ece%d" % k])))
        assert abs(got - float(z["ece%d" % k])) < 2e-7            # reference's own ece_hist_binary value (float32 sum)
        assert abs(fa.ece_width(p, lab) - stats.ece_width(p, lab)) < 1e-12
        # the statistic hls4ml_pred.py:90-91 really computes (probabilities re-soft-maxed by tfp) vs its restatement
        assert abs(fa.ece_tfp_as_called(p, lab) - stats.ece_tfp_as_called(p, lab)) < 1e-6
        assert abs(fa.ece_tfp_as_called(p, lab) - fa.ece_width(p, lab)) > 1e-3      # ... and it is a different number
    assert fa.ece_width(np.zeros((0, 10)), np.zeros(0, dtype=np.int64)) == 0.0        # empty dataset
    # NLL / MSE / accuracy of ece_eval_binary (results_analyzer.py:497-503) on the device vs the float64 restatement
    for k in range(3):
        p, lab = z["p%d" % k], z["label%d" % k]
        onehot = np.eye(p.shape[1])[lab]
        nll, mse, acc = stats.nll_mse_acc(p, onehot)
        ece, g_nll, g_mse, g_acc = fa.ece_eval_binary(p, onehot)
        assert abs(g_nll - nll) < 1e-12 * max(1.0, nll) and abs(g_mse - mse) < 1e-12 and g_acc == acc
        assert abs(ece - stats.ece_kde(p, onehot)) < 1e-8         # the ECE of ece_eval_binary is the KDE one (:503)


def test_statistics_branches_frozen_from_the_reference_source(lib):
    """ece_kde_binary with a separate integration set p_int / order 2 / the binary (C == 2) joint-calibration branch,
    ece_hist_binary's binary branch, and ece_eval_binary on probabilities only float64 can represent (NLL clip 1e-256):
    device (float64) vs values frozen from the reference's own source (tests/golden/analysis.npz)."""
    from bayesnn_fpga_b200.results_analyzer import FullAnalysis
    from tests.cases import GOLDEN
    z = np.load(GOLDEN + "/analysis.npz")

    class Dummy:
        n_exits, out_dim = 1, 10
    fa = FullAnalysis(Dummy(), None, run=False)
    p, p_int, lab = z["kde_p"], z["kde_p_int"], z["kde_lab"]
    onehot = np.eye(10)[lab]
    got = [fa.ece_kde_binary(p, onehot, p_int=p_int), fa.ece_kde_binary(p, onehot, p_int=p_int, order=2),
           fa.ece_kde_binary(p, onehot, order=2)]
    pb, pb_int, labb = z["bin_p"], z["bin_p_int"], z["bin_lab"]
    onehot_b = np.eye(2)[labb]
    got_b = [fa.ece_kde_binary(pb, onehot_b), fa.ece_kde_binary(pb, onehot_b, p_int=pb_int),
             fa.ece_kde_binary(pb, onehot_b, order=2)]
    d1, d2 = np.abs(np.asarray(got) - z["kde_vals"]).max(), np.abs(np.asarray(got_b) - z["bin_kde_vals"]).max()
    hist = [fa.ece_hist_binary(pb, onehot_b), fa.ece_hist_binary(pb, onehot_b, n_bins=10, order=2),
            fa.ece_hist_binary(p, onehot), fa.ece_hist_binary(p, onehot, n_bins=7)]
    d3 = np.abs(np.asarray(hist) - z["hist_vals"]).max()
    eb = np.asarray(fa.ece_eval_binary(pb, onehot_b))
    eu = np.asarray(fa.ece_eval_binary(z["under_p"], onehot))
    d4, d5 = np.abs(eb - z["bin_eval"]).max(), np.abs(eu - z["under_eval"]) / np.maximum(1.0, np.abs(z["under_eval"]))
    report(test="stats_branches", kde_p_int=float(d1), kde_binary=float(d2), hist=float(d3), eval_binary=float(d4),
           eval_underflow=float(d5.max()), nll_underflow=float(eu[1]))
    assert d1 < 1e-8 and d2 < 1e-8 and d3 < 2e-7 and d4 < 1e-8 and d5.max() < 1e-8
    assert eu[1] > 87.4 / 7                                        # mean NLL carries the 589.5-per-image penalties
    with pytest.raises(ValueError):
        fa.ece_kde_binary(p, onehot, p_int=pb_int)


TC_SHAPES = [
    # N, H, W, Cin, Cout, k, stride, pad, relu, res
    (8, 32, 32, 64, 64, 3, 1, 1, True, True),       # layer1 conv (BN = 64)
    (4, 32, 32, 64, 128, 3, 2, 1, True, False),     # layer2.0.conv1 / ex1conv1: stride 2 through parity planes
    (4, 32, 32, 64, 128, 1, 2, 0, False, False),    # 1x1 stride-2 shortcut
    (4, 16, 16, 128, 128, 3, 1, 1, True, True),     # layer2 conv + residual
    (16, 8, 8, 256, 256, 3, 1, 1, True, True),      # layer3 (BN = 256), two images per tile
    (40, 4, 4, 512, 512, 3, 1, 1, True, True),      # layer4: 8 images per tile, two N tiles
    (3, 8, 8, 128, 256, 3, 2, 1, True, False),      # M = 48 < one tile, box larger than the tensor
    (130, 2, 2, 512, 512, 3, 1, 1, True, False),    # VGG block 3 (2x2 maps)
    (300, 1, 1, 512, 512, 3, 1, 1, True, False),    # VGG block 4 (1x1 maps: only the centre tap is in bounds)
    (6, 4, 4, 256, 512, 3, 2, 1, True, False),      # VGG ex3 feature extractor
    (2, 64, 64, 64, 64, 3, 1, 1, False, False),     # wider image: 2 rows per tile
]


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_tc_matches_fp64_conv_of_the_rounded_operands(lib, shape, dt):
    got, want = _conv_case(lib, "tc", dt, *shape)
    err = (got.double() - want).abs().max().item()
    scale = max(1.0, want.abs().max().item())
    report(test="conv_tc", dtype=dt, shape=list(shape), max_err=err, scale=scale)
    # fp32 accumulation of exact 16-bit products + one rounding of the output to 16-bit storage
    assert err <= (1e-3 if dt == "fp16" else 8e-3) * scale


@pytest.mark.parametrize("cg2", ["0", "1"])
@pytest.mark.parametrize("shape", [s for s in TC_SHAPES if s[4] % 256 == 0])
def test_conv_tc_cta_pair_multicast(lib, shape, cg2, monkeypatch):
    """the CTA-pair kernels on small shapes incl. an odd number of row-tiles: cg2=0 weight halves TMA-multicast into
    both CTAs, 1-CTA MMAs; cg2=1 tcgen05.mma.cta_group::2 (M = 256 over the pair, each CTA keeps half the weights)."""
    monkeypatch.setenv("BNN_TC_MC_MIN_TILES", "1")
    monkeypatch.setenv("BNN_TC_CG2", cg2)
    got, want = _conv_case(lib, "tc", "fp16", *shape)
    assert (got.double() - want).abs().max().item() <= 1e-3 * max(1.0, want.abs().max().item())
    monkeypatch.setenv("BNN_TC_NOMC", "1")
    ref, _ = _conv_case(lib, "tc", "fp16", *shape)
    assert torch.equal(got, ref)                                  # bit-identical to the single-CTA kernel


@pytest.mark.parametrize("shape", [s for s in TC_SHAPES if s[4] % 256 != 0])
def test_conv_tc_cta_group2_narrow(lib, shape, monkeypatch):
    """cta_group::2 with 128 / 64 output channels (two row-tiles per CTA share each weight half)."""
    monkeypatch.setenv("BNN_TC_MC_MIN_TILES", "1")
    monkeypatch.setenv("BNN_TC_CG2_NARROW", "1")
    got, want = _conv_case(lib, "tc", "fp16", *shape)
    assert (got.double() - want).abs().max().item() <= 1e-3 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_conv_tc_fused_site(lib, kind):
    N_img, B, S, C, HW = 12, 4, 3, 128, 8
    masks = (torch.rand(4, C) > 0.5).float().cuda()
    dd = drop_desc(kind, 0.5, 0x77, 4, 5, B, masks if kind == 3 else None, cnt0=2)
    got, want = _conv_case(lib, "tc", "fp16", N_img, HW, HW, 64, C, 3, 1, 1, True, False, drop=dd)
    for s in range(S):
        if kind == 3:
            f = masks[(2 + 5 + s) % 4].cpu().view(1, C, 1, 1).double()
        else:
            keep = philox.keep_mask(0x77, 4, 5 + s, (B, C, HW, HW), 0.5, "channel" if kind == 2 else "element")
            f = torch.from_numpy(keep).double() * 2.0
        ref = want[s * B:(s + 1) * B] * f
        assert (got[s * B:(s + 1) * B].double() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def test_conv_tc_rejects_unsupported_geometry(lib):
    y = torch.zeros(64, device="cuda", dtype=torch.float16)
    dd = drop_desc(batch=1)
    args = lambda cin, cout, k, h: (y.data_ptr(), y.data_ptr(), y.data_ptr(), None, y.data_ptr(), 1, 1, h, h, cin,
                                    cout, k, 1, 0, ctypes.byref(dd), stream())
    assert lib.bnn_conv2d_tc(*args(3, 64, 3, 32)) == -4         # Cin % 64 != 0 -> use the CUDA-core kernel
    assert lib.bnn_conv2d_tc(*args(64, 64, 5, 32)) == -4        # 5x5
    assert lib.bnn_conv2d_tc(*args(64, 64, 3, 12)) == -4        # 12x12 output rows do not tile 128


@pytest.mark.parametrize("Cin,Cg,HW,G", [(64, 128, 16, 3), (128, 256, 8, 3), (256, 512, 8, 2)])
def test_conv_tc_grouped_equals_separate_convs(lib, Cin, Cg, HW, G):
    """sibling stride-2 convs on one input (3x3, 3x3, 1x1-in-the-centre-tap) as one launch == separate convs."""
    N = 6
    g = torch.Generator().manual_seed(Cg)
    x = torch.randn(N, Cin, HW, HW, generator=g).half()
    ws = [torch.randn(Cg, Cin, 3, 3, generator=g) / np.sqrt(9 * Cin) for _ in range(G - 1)]
    w1 = torch.randn(Cg, Cin, 1, 1, generator=g) / np.sqrt(Cin)
    w3 = torch.zeros(Cg, Cin, 3, 3)
    w3[:, :, 1, 1] = w1[:, :, 0, 0]
    wcat = torch.cat(ws + [w3]).half()
    bias = torch.randn(G * Cg, generator=g)
    relu_mask = (1 << (G - 1)) - 1                               # ReLU on the 3x3 outputs, not on the shortcut
    d_x = x.permute(0, 2, 3, 1).contiguous().cuda()
    d_w = wcat.permute(0, 2, 3, 1).contiguous().cuda()
    d_b = bias.cuda()
    outs = [torch.full((N, HW // 2, HW // 2, Cg), float("nan"), dtype=torch.float16, device="cuda") for _ in range(G)]
    ys = (ctypes.c_void_p * G)(*[o.data_ptr() for o in outs])
    rc = lib.bnn_conv2d_tc_grouped(d_x.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), ys, G, relu_mask, 1 << (G - 1), 1, N,
                                   HW, HW, Cin, Cg, 3, 2, stream())
    assert rc == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    xd = x.double()
    for i in range(G):
        if i < G - 1:
            want = F.conv2d(xd, ws[i].half().double(), bias[i * Cg:(i + 1) * Cg].double(), 2, 1).relu()
        else:
            want = F.conv2d(xd, w1.half().double(), bias[i * Cg:].double(), 2, 0)
        got = outs[i].cpu().permute(0, 3, 1, 2).double()
        assert (got - want).abs().max().item() <= 1e-3 * max(1.0, want.abs().max().item()), i


# ---- Masksembles gathered layout (bnn_drop_desc.compact_*, bnn_conv2d_tc_gathered) -------------------------
def _mask_tables(n, C, kept, seed):
    """n random 0/1 rows with `kept` ones each + the pos / idx tables of the gathered layout."""
    rng = np.random.RandomState(seed)
    kc = (kept + 15) // 16 * 16
    masks = np.zeros((n, C), np.float32)
    pos = np.full((n, C), -1, np.int16)
    idx = np.full((n, kc), -1, np.int16)
    for r in range(n):
        k = np.sort(rng.choice(C, kept, replace=False))
        masks[r, k] = 1
        pos[r, k] = np.arange(kept)
        idx[r, :kept] = k
    return masks, pos, idx, kc


def _compact_desc(masks, pos, idx, kc, B, sample0, cnt0):
    dm, dp, di = torch.from_numpy(masks).cuda(), torch.from_numpy(pos).cuda(), torch.from_numpy(idx).cuda()
    dd = drop_desc(3, 0.0, 0, 0, sample0, B, dm, cnt0=cnt0)
    dd.compact_pos, dd.compact_idx, dd.compact_c = dp.data_ptr(), di.data_ptr(), kc
    return dd, (dm, dp, di)


@pytest.mark.parametrize("x_has_samples", [0, 1])
def test_masksembles_compact_site(lib, x_has_samples):
    """stand-alone site in the gathered layout == the dense masked tensor with the dropped channels removed."""
    B, S, H, C, kept, n, s0, cnt0 = 3, 5, 4, 64, 34, 4, 2, 3
    masks, pos, idx, kc = _mask_tables(n, C, kept, 1)
    x = torch.randn((S if x_has_samples else 1) * B, H, H, C).half().cuda()
    y = torch.zeros(S * B, H, H, kc, dtype=torch.float16, device="cuda")
    dd, keep = _compact_desc(masks, pos, idx, kc, B, s0, cnt0)
    assert lib.bnn_dropout(x.data_ptr(), y.data_ptr(), 1, H * H * C, C, S, x_has_samples, ctypes.byref(dd),
                           stream()) == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    xs = x.view(-1, B, H, H, C).cpu()
    for s in range(S):
        r = (cnt0 + s0 + s) % n
        want = xs[s if x_has_samples else 0][..., idx[r, :kept].astype(np.int64)]
        got = y.view(S, B, H, H, kc)[s].cpu()
        assert torch.equal(got[..., :kept], want) and (got[..., kept:] == 0).all()


@pytest.mark.parametrize("Cout,M_img,pair", [(128, 16, "0"), (256, 16, "0"), (256, 28, "cg2"), (256, 28, "mc2")])
def test_conv_tc_fused_compact_epilogue(lib, Cout, M_img, pair, monkeypatch):
    """conv + residual + ReLU + Masksembles2D stored in the gathered layout == the dense fused epilogue with the
    dropped channels removed (bit-exact: same arithmetic, different addresses)."""
    if pair != "0":         # swapped / single-CTA / cta_group::2 / (multicast has no compact epilogue: single-CTA runs)
        monkeypatch.setenv("BNN_TC_MC_MIN_TILES", "1")
        monkeypatch.setenv("BNN_TC_CG2", "1" if pair == "cg2" else "0")
    B, H, Cin, n, s0, cnt0 = 4, 8, 64, 4, 1, 2
    S = M_img // B
    kept = {128: 68, 256: 136}[Cout]
    masks, pos, idx, kc = _mask_tables(n, Cout, kept, Cout)
    g = torch.Generator().manual_seed(Cout + M_img)
    x = torch.randn(S * B, H, H, Cin, generator=g).half().cuda()
    w = (torch.randn(Cout, 3, 3, Cin, generator=g) / 24).half().cuda()
    b = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(S * B, H, H, Cout, generator=g).half().cuda()
    dense = torch.empty(S * B, H, H, Cout, dtype=torch.float16, device="cuda")
    comp = torch.zeros(S * B, H, H, kc, dtype=torch.float16, device="cuda")
    dd_c, keep = _compact_desc(masks, pos, idx, kc, B, s0, cnt0)
    dd_d = drop_desc(3, 0.0, 0, 0, s0, B, keep[0], cnt0=cnt0)
    for dd, y in ((dd_d, dense), (dd_c, comp)):
        assert lib.bnn_conv2d_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), res.data_ptr(), y.data_ptr(), 1, S * B, H, H,
                                 Cin, Cout, 3, 1, 1, ctypes.byref(dd), stream()) == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    dv, cv = dense.view(S, B, H, H, Cout).cpu(), comp.view(S, B, H, H, kc).cpu()
    for s in range(S):
        r = (cnt0 + s0 + s) % n
        assert torch.equal(cv[s][..., :kept], dv[s][..., idx[r, :kept].astype(np.int64)]), s
        assert (cv[s][..., kept:] == 0).all()
    assert dv.abs().sum() > 0


@pytest.mark.parametrize("C,kept,cout_g,groups,stride,H,B", [
    (64, 34, 128, 3, 2, 16, 4),       # site 1 of the ResNet: layer2.0.conv1 + shortcut (centre tap) + ex1conv1
    (128, 68, 256, 3, 2, 8, 16),      # site 2
    (256, 136, 512, 3, 2, 8, 16),     # site 3: three channel blocks, the last one 16 channels wide
    (256, 136, 256, 1, 1, 8, 8),      # stride 1, single output
    (128, 68, 64, 1, 1, 16, 2),       # 64-channel tile kernel
    (256, 136, 512, 1, 2, 16, -4),    # B = 4 with the CTA-pair kernels forced: cta_group::2
    (256, 136, 256, 2, 2, 16, -8),
    (64, 34, 128, 3, 2, 16, -4),      # 128-channel groups paired into 256-column cta_group::2 tiles (+ a dead half-tile)
    (128, 68, 128, 2, 2, 16, -4),
])
def test_conv_tc_gathered_vs_dense(lib, C, kept, cout_g, groups, stride, H, B, monkeypatch):
    """gathered-K convolution on the compact tensor == the grouped convolution on the dense masked tensor (same
    products; the tensor cores sum them in a different order) and == float64 torch."""
    if B < 0:
        monkeypatch.setenv("BNN_TC_MC_MIN_TILES", "1")
        monkeypatch.setenv("BNN_TC_PAIR_GROUPS", "1")     # experiment flag: 128-channel groups in 256-column tiles
        B = -B
    n, S, s0, cnt0 = 4, 3, 1, 2
    masks, pos, idx, kc = _mask_tables(n, C, kept, C + cout_g)
    g = torch.Generator().manual_seed(C * 7 + cout_g)
    Cout = groups * cout_g
    xd = torch.randn(S, B, H, H, C, generator=g).half()
    for s in range(S):
        xd[s] *= torch.from_numpy(masks[(cnt0 + s0 + s) % n])
    w = (torch.randn(Cout, 3, 3, C, generator=g) / np.sqrt(9 * kept)).half()
    center_mask = 0b010 if groups >= 2 else 0
    if center_mask:
        w[cout_g:2 * cout_g, [0, 0, 0, 1, 1, 2, 2, 2], [0, 1, 2, 0, 2, 0, 1, 2]] = 0     # a 1x1 kernel in the centre tap
    b = torch.randn(Cout, generator=g)
    relu_mask = 0b101 if groups >= 2 else 1
    OH = H // stride
    xc = torch.zeros(S, B, H, H, kc, dtype=torch.float16)
    wg = torch.zeros(n, Cout, 3, 3, kc, dtype=torch.float16)
    for r in range(n):
        wg[r, ..., :kept] = w[..., idx[r, :kept].astype(np.int64)]
    for s in range(S):
        xc[s, ..., :kept] = xd[s][..., idx[(cnt0 + s0 + s) % n, :kept].astype(np.int64)]
    xc[..., kept:] = 7.0            # padding slots may hold anything finite: their weights are zero
    d = [t.cuda() for t in (xd, w, b, xc, wg)]
    outs = {}
    for name in ("dense", "gathered"):
        ys = [torch.full((S * B, OH, OH, cout_g), float("nan"), dtype=torch.float16, device="cuda") for _ in range(groups)]
        yp = (ctypes.c_void_p * groups)(*[y.data_ptr() for y in ys])
        if name == "dense":
            rc = lib.bnn_conv2d_tc_grouped(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), yp, groups, relu_mask,
                                           center_mask, 1, S * B, H, H, C, cout_g, 3, stride, stream())
        else:
            rc = lib.bnn_conv2d_tc_gathered(d[3].data_ptr(), d[4].data_ptr(), d[2].data_ptr(), yp, groups, relu_mask,
                                            center_mask, 1, S * B, H, H, kc, cout_g, 3, stride, n, cnt0, s0, B, 1, stream())
        assert rc == 0, lib.bnn_last_error()
        torch.cuda.synchronize()
        outs[name] = torch.cat([y.cpu().float() for y in ys], -1)
    want = F.conv2d(xd.view(S * B, H, H, C).permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), b.double(),
                    stride, 1).permute(0, 2, 3, 1)
    for gi in range(groups):
        if (relu_mask >> gi) & 1:
            want[..., gi * cout_g:(gi + 1) * cout_g].relu_()
    scale = max(1.0, want.abs().max().item())
    e_ref = (outs["gathered"].double() - want).abs().max().item()
    e_dense = (outs["gathered"] - outs["dense"]).abs().max().item()
    report(test="conv_tc_gathered", C=C, kept=kept, cout=Cout, err_vs_f64=e_ref, err_vs_dense=e_dense)
    assert e_ref <= 2e-3 * scale and e_dense <= 2e-3 * scale


def test_conv_tc_gathered_bad_args(lib):
    y = torch.zeros(16, device="cuda")
    yp = (ctypes.c_void_p * 1)(y.data_ptr())
    # batch * OH * OW not a multiple of 256: a tile pair would straddle two samples
    assert lib.bnn_conv2d_tc_gathered(y.data_ptr(), y.data_ptr(), y.data_ptr(), yp, 1, 0, 0, 1, 6, 8, 8, 48, 64, 3, 1, 4,
                                      0, 0, 3, 1, stream()) == -1
    assert b"multiple of 256" in lib.bnn_last_error()
    # Kc must be a multiple of 16
    assert lib.bnn_conv2d_tc_gathered(y.data_ptr(), y.data_ptr(), y.data_ptr(), yp, 1, 0, 0, 1, 8, 8, 8, 40, 64, 3, 1, 4,
                                      0, 0, 4, 1, stream()) == -4


def test_full_analysis_confidence_exiting_flops_and_kde_ece(lib):
    """FullAnalysis.confidence_exiting / flop_saver / flop_saver_ensembled / ece_eval_binary on the device vs the
    values frozen from the reference's own source (tests/golden/analysis.npz)."""
    from bayesnn_fpga_b200.results_analyzer import FullAnalysis
    from tests.cases import GOLDEN
    z = np.load(GOLDEN + "/analysis.npz")

    class Dummy:
        n_exits, out_dim = 1, 10
    for k, mt in enumerate(["resnet18", "vgg19"]):
        p_evals, lab = z["p%d" % k], z["lab%d" % k]
        E, N, C = p_evals.shape
        onehot = np.eye(C)[lab]
        fa = FullAnalysis(Dummy(), None, run=False)
        fa.model_type = mt
        fa.get_flops_per_module()
        assert fa.baseline_flops == stats.baseline_flops(mt) and fa.n_exits == E
        for r, row in enumerate(z["rows%d" % k]):
            thr, diff = float(row[0]), bool(row[1])
            idx, best, hist = fa._exit_scan(thr, p_evals, diff)
            assert (idx == z["exit%d_%d" % (k, r)]).all() and hist.sum() == N and hist[0] == 0
            accu, ece, nll = fa.confidence_exiting(thr, p_evals, onehot, diff=diff)
            report(test="confidence_exiting", model=mt, thr=thr, diff=diff, d_acc=abs(accu - row[2]), d_ece=abs(ece - row[3]),
                   d_nll=abs(nll - row[4]))
            assert abs(accu - row[2]) < 1e-12 and abs(ece - row[3]) < 1e-8 and abs(nll - row[4]) < 1e-12 * max(1.0, row[4])
            got = []
            for eo in (True, False):
                fa.exit_only = eo
                got += [fa.flop_saver(thr, p_evals, onehot, mc_passes=10, diff=diff),
                        fa.flop_saver_ensembled(thr, p_evals, onehot, mc_passes=10, diff=diff)]
            assert [float(v) for v in got] == list(row[5:9])
            for layer in range(E):
                assert fa.get_flops_standard_exit(layer, 10, True) == stats.flops_standard_exit(mt, layer, 10, True)
        for e in range(E):
            ece, nll, mse, accu = fa.ece_eval_binary(p_evals[e], onehot)
            w = z["per_exit%d" % k][e]
            report(test="kde_ece", model=mt, exit=e, ece=ece, want=float(w[0]))
            assert abs(ece - w[0]) < 1e-8 and abs(nll - w[1]) < 1e-12 * max(1.0, w[1]) and abs(mse - w[2]) < 1e-12 \
                and abs(accu - w[3]) < 1e-12
    # the KDE kernel alone against the float64 estimator, incl. flags and mirroring
    rng = np.random.RandomState(3)
    data = rng.beta(5, 2, 700)
    flags = (rng.rand(700) > 0.4).astype(np.int32)
    G, x0, dx, bw = 4096, -0.6, 2.2 / 4095, 0.013
    out = torch.zeros(G, dtype=torch.float64, device="cuda")
    d_data, d_flags = torch.from_numpy(data).cuda(), torch.from_numpy(flags).cuda()
    assert lib.bnn_kde_triweight(d_data.data_ptr(), d_flags.data_ptr(), 700, bw, float(flags.sum()), x0, dx, G, 0.0, 1.0,
                                 out.data_ptr(), stream()) == 0, lib.bnn_last_error()
    grid = x0 + dx * np.arange(G)
    sel = data[flags != 0].astype(np.float64).reshape(-1, 1)
    want = stats.kde_triweight_exact(stats.mirror_1d(sel, 0.0, 1.0), bw, grid) * 2
    want[(grid <= 0) | (grid >= 1)] = 0
    assert np.abs(out.cpu().numpy() - want).max() < 1e-9 * max(1.0, want.max())


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
@pytest.mark.parametrize("C,F_,HW,kind,S", [(100, 512, 4, 1, 19), (100, 512, 1, 3, 4), (37, 64, 16, 0, 33), (1000, 256, 1, 2, 16)])
def test_exit_head_mma_matches_ffma_head(lib, dt, C, F_, HW, kind, S):
    """bnn_exit_head_mma (mma.sync, hi/lo-split operands) == bnn_exit_head (fp32 FFMA) on the same 16-bit features:
    the split keeps the tensor-core logits within 1e-5 relative of the float32 form."""
    tdt, code = TORCH_DT[dt]
    B, s0, p, seed, sid = 5, 3, 0.25, 0x99, 2
    g = torch.Generator().manual_seed(C + S)
    feat = (torch.randn(S * B, HW, F_, generator=g).abs() * 3).to(tdt).cuda()
    w = (torch.randn(C, F_, generator=g) / np.sqrt(F_) * 3).cuda()
    bias = torch.randn(C, generator=g).cuda()
    masks = (torch.rand(4, F_, generator=g) > 0.5).float().cuda()
    dd = drop_desc(kind, p, seed, sid, s0, B, masks if kind == 3 else None, cnt0=1)
    w_hi, w_lo = torch.empty(C, F_, dtype=tdt, device="cuda"), torch.empty(C, F_, dtype=tdt, device="cuda")
    assert lib.bnn_split16(w.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), C * F_, code, stream()) == 0
    assert (w_hi.float() + w_lo.float() - w).abs().max().item() <= (2e-6 if dt == "fp16" else 2e-4) * w.abs().max().item()
    outs = {}
    for name in ("ffma", "mma"):
        o = dict(sp=torch.zeros(B, C).cuda(), sl=torch.zeros(B, C).cuda(), spl=torch.zeros(B).cuda(), lo=torch.zeros(S, B, C).cuda())
        if name == "ffma":
            wt = w.t().contiguous()
            rc = lib.bnn_exit_head(feat.data_ptr(), code, 1, B, S, HW, F_, C, wt.data_ptr(), bias.data_ptr(), ctypes.byref(dd),
                                   o["sp"].data_ptr(), o["sl"].data_ptr(), o["spl"].data_ptr(), o["lo"].data_ptr(), 0, stream())
        else:
            rc = lib.bnn_exit_head_mma(feat.data_ptr(), code, 1, B, S, HW, F_, C, w_hi.data_ptr(), w_lo.data_ptr(),
                                       bias.data_ptr(), ctypes.byref(dd), o["sp"].data_ptr(), o["sl"].data_ptr(),
                                       o["spl"].data_ptr(), o["lo"].data_ptr(), 0, stream())
        assert rc == 0, lib.bnn_last_error()
        torch.cuda.synchronize()
        outs[name] = o
    scale = max(1.0, outs["ffma"]["lo"].abs().max().item())
    e_lo = (outs["mma"]["lo"] - outs["ffma"]["lo"]).abs().max().item()
    e_p = (outs["mma"]["sp"] - outs["ffma"]["sp"]).abs().max().item()
    report(test="exit_head_mma", dtype=dt, C=C, F=F_, err_logits=e_lo, err_sum_p=e_p, scale=scale)
    tol = 1e-5 if dt == "fp16" else 3e-4          # bf16 hi+lo carries 16 mantissa bits
    assert e_lo <= tol * scale and e_p <= tol * S
    assert lib.bnn_exit_head_mma(feat.data_ptr(), code, 1, B, S, HW, 24, C, w_hi.data_ptr(), w_lo.data_ptr(), bias.data_ptr(),
                                 ctypes.byref(dd), o["sp"].data_ptr(), o["sl"].data_ptr(), o["spl"].data_ptr(), None, 0,
                                 stream()) == -1                     # F % 16 != 0


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
@pytest.mark.parametrize("C,F_,HW,kind,p,S,B", [(100, 512, 4, 1, 0.25, 19, 5), (100, 512, 1, 1, 0.5, 40, 37), (10, 512, 1, 3, 0.0, 4, 9),
                                              (10, 512, 16, 1, 0.5, 33, 6), (1000, 256, 1, 2, 0.125, 16, 3), (37, 64, 1, 0, 0.0, 3, 70)])
def test_exit_head_tc_matches_ffma_head(lib, dt, C, F_, HW, kind, p, S, B):
    """bnn_exit_head_tc - pool -> site -> hi/lo split, the classifier as ONE tcgen05 GEMM over [x_hi | x_lo] x
    [w_hi | w_lo | w_hi]^T with fp32 logits, soft-max / accumulation - == bnn_exit_head (fp32 FFMA) on the same 16-bit
    features, with and without the x_lo block (exact features: HW = 1 and a power-of-two keep scale)."""
    tdt, code = TORCH_DT[dt]
    s0, seed, sid = 3, 0x99, 2
    g = torch.Generator().manual_seed(C + S)
    feat = (torch.randn(S * B, HW, F_, generator=g).abs() * 3).to(tdt).cuda()
    w = (torch.randn(C, F_, generator=g) / np.sqrt(F_) * 3).cuda()
    bias = torch.randn(C, generator=g).cuda()
    masks = (torch.rand(4, F_, generator=g) > 0.5).float().cuda()
    dd = drop_desc(kind, p, seed, sid, s0, B, masks if kind == 3 else None, cnt0=1)
    w_hi, w_lo = torch.empty(C, F_, dtype=tdt, device="cuda"), torch.empty(C, F_, dtype=tdt, device="cuda")
    assert lib.bnn_split16(w.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), C * F_, code, stream()) == 0
    scale_keep = 1.0 / (1.0 - p) if kind in (1, 2) else 1.0
    with_lo = int(not (HW == 1 and abs(np.log2(scale_keep) - round(np.log2(scale_keep))) < 1e-12))
    c_pad = (C + 63) // 64 * 64 if C <= 128 else (C + 255) // 256 * 256
    parts = [w_hi, w_lo] + ([w_hi] if with_lo else [])
    w3 = torch.zeros(c_pad, len(parts) * F_, dtype=tdt, device="cuda")
    w3[:C] = torch.cat(parts, dim=1)
    b_pad = torch.zeros(c_pad, device="cuda")
    b_pad[:C] = bias
    a_ws = torch.empty(S * B * (2 if with_lo else 1) * F_, dtype=tdt, device="cuda")
    l_ws = torch.empty(S * B * c_pad, device="cuda")
    outs = {}
    for name in ("ffma", "tc"):
        o = dict(sp=torch.zeros(B, C).cuda(), sl=torch.zeros(B, C).cuda(), spl=torch.zeros(B).cuda(), lo=torch.zeros(S, B, C).cuda())
        if name == "ffma":
            wt = w.t().contiguous()
            rc = lib.bnn_exit_head(feat.data_ptr(), code, 1, B, S, HW, F_, C, wt.data_ptr(), bias.data_ptr(), ctypes.byref(dd),
                                   o["sp"].data_ptr(), o["sl"].data_ptr(), o["spl"].data_ptr(), o["lo"].data_ptr(), 0, stream())
        else:
            rc = lib.bnn_exit_head_tc(feat.data_ptr(), code, 1, B, S, HW, F_, C, w3.data_ptr(), b_pad.data_ptr(), c_pad, with_lo,
                                      ctypes.byref(dd), a_ws.data_ptr(), l_ws.data_ptr(), o["sp"].data_ptr(), o["sl"].data_ptr(),
                                      o["spl"].data_ptr(), o["lo"].data_ptr(), 0, stream())
        assert rc == 0, lib.bnn_last_error()
        torch.cuda.synchronize()
        outs[name] = o
    scale = max(1.0, outs["ffma"]["lo"].abs().max().item())
    e_lo = (outs["tc"]["lo"] - outs["ffma"]["lo"]).abs().max().item()
    e_p = (outs["tc"]["sp"] - outs["ffma"]["sp"]).abs().max().item()
    e_pl = (outs["tc"]["spl"] - outs["ffma"]["spl"]).abs().max().item()
    report(test="exit_head_tc", dtype=dt, C=C, F=F_, HW=HW, with_lo=with_lo, err_logits=e_lo, err_sum_p=e_p, scale=scale)
    tol = 1e-5 if dt == "fp16" else 3e-4          # bf16 hi+lo carries 16 mantissa bits
    assert e_lo <= tol * scale and e_p <= tol * S and e_pl <= 10 * tol * S
    assert lib.bnn_exit_head_tc(feat.data_ptr(), code, 1, B, S, HW, 24, C, w3.data_ptr(), b_pad.data_ptr(), c_pad, with_lo,
                                ctypes.byref(dd), a_ws.data_ptr(), l_ws.data_ptr(), o["sp"].data_ptr(), o["sl"].data_ptr(),
                                o["spl"].data_ptr(), None, 0, stream()) == -1          # F % 64 != 0


@pytest.mark.parametrize("Cin2,C,H,N,pair", [(64, 128, 16, 6, "0"), (128, 256, 8, 20, "0"), (256, 512, 4, 70, "0"),
                                              (128, 256, 8, 20, "cg2"), (128, 256, 8, 21, "mc2")])
@pytest.mark.parametrize("drop_kind", [0, 1])
def test_conv_tc_fused_shortcut(lib, Cin2, C, H, N, pair, drop_kind, monkeypatch):
    """conv3x3(h) + conv1x1_stride2(x2) + bias -> ReLU [-> dropout] as ONE GEMM with the shortcut as extra K
    (bnn_conv2d_tc_shortcut) vs float64 torch (BasicBlock.forward resnet18.py:41-46 with the BatchNorms folded)."""
    if pair != "0":
        monkeypatch.setenv("BNN_TC_MC_MIN_TILES", "1")
        monkeypatch.setenv("BNN_TC_CG2", "1" if pair == "cg2" else "0")
    g = torch.Generator().manual_seed(Cin2 + N)
    h = torch.randn(N, C, H, H, generator=g).half()
    x2 = torch.randn(N, Cin2, 2 * H, 2 * H, generator=g).half()
    w = (torch.randn(C, C, 3, 3, generator=g) / np.sqrt(9 * C)).half()
    wd = (torch.randn(C, Cin2, generator=g) / np.sqrt(Cin2)).half()
    b = torch.randn(C, generator=g)
    want = (F.conv2d(h.double(), w.double(), None, 1, 1) + F.conv2d(x2.double(), wd.double()[:, :, None, None], None, 2, 0)
            + b.double().view(1, -1, 1, 1)).relu()
    B, S = (N // 2, 2) if N % 2 == 0 else (N, 1)
    dd = drop_desc(drop_kind, 0.5, 0x77, 4, 5, B)
    if drop_kind == 1:
        for s_ in range(S):
            keep = torch.from_numpy(philox.keep_mask(0x77, 4, 5 + s_, (B, C, H, H), 0.5))
            want[s_ * B:(s_ + 1) * B] *= keep * 2.0
    d_h = h.permute(0, 2, 3, 1).contiguous().cuda()
    d_x2 = x2.permute(0, 2, 3, 1).contiguous().cuda()
    d_w = torch.cat([w.permute(0, 2, 3, 1).reshape(C, -1), wd], dim=1).contiguous().cuda()
    d_b = b.cuda()
    d_y = torch.full((N, H, H, C), float("nan"), dtype=torch.float16, device="cuda")
    rc = lib.bnn_conv2d_tc_shortcut(d_h.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), None, d_y.data_ptr(), 1, N, H, H, C, C, 3, 1,
                                    1, ctypes.byref(dd), d_x2.data_ptr(), 2 * H, 2 * H, Cin2, stream())
    assert rc == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    got = d_y.cpu().float().permute(0, 3, 1, 2)
    err = (got.double() - want).abs().max().item()
    report(test="conv_tc_shortcut", Cin2=Cin2, C=C, pair=pair, drop=drop_kind, err=err, scale=want.abs().max().item())
    assert err <= 2e-3 * max(1.0, want.abs().max().item())
    # geometry that is not a stride-2 projection is rejected
    assert lib.bnn_conv2d_tc_shortcut(d_h.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), None, d_y.data_ptr(), 1, N, H, H, C, C, 3,
                                      1, 1, ctypes.byref(dd), d_x2.data_ptr(), H, H, Cin2, stream()) == -1


@pytest.mark.parametrize("N,Cin,HW,G,center", [(5, 64, 32, 2, 0), (171, 64, 32, 2, 0), (7, 128, 16, 4, 0b1100), (3, 64, 16, 2, 0b10)])
def test_conv_tc_sibling_pair_cta_group2(lib, N, Cin, HW, G, center, monkeypatch):
    """Operand-swapped kernel as CTA pairs (tcgen05.mma.cta_group::2): two 128-channel sibling groups that read the same
    pixels run on one pair - one group per CTA, the 256-pixel tile split over the pair.  vs float64 torch per group,
    and bit-identical to the one-CTA-per-(tile, group) kernel (same accumulation order); more tiles than SM pairs;
    a pair of centre-tap (1x1) groups; a pair that DISAGREES on the centre-tap flag falls back to single CTAs.  The two
    Cin = 64 cases (one group pair, nine k-blocks) run with the weights RESIDENT in shared memory (Params::rw_kb)."""
    g = torch.Generator().manual_seed(N * G)
    x = torch.randn(N, Cin, HW, HW, generator=g).half()
    ws = []
    for i in range(G):
        w = torch.randn(128, Cin, 3, 3, generator=g) / np.sqrt(9 * Cin)
        if (center >> i) & 1:                                    # a 1x1 kernel held in the centre tap
            c = w[:, :, 1, 1].clone()
            w.zero_()
            w[:, :, 1, 1] = c
        ws.append(w.half())
    bias = torch.randn(G * 128, generator=g)
    relu_mask = 0b0101 & ((1 << G) - 1)
    d_x = x.permute(0, 2, 3, 1).contiguous().cuda()
    d_w = torch.cat(ws).permute(0, 2, 3, 1).contiguous().cuda()
    d_b = bias.cuda()

    def run():
        outs = [torch.full((N, HW // 2, HW // 2, 128), float("nan"), dtype=torch.float16, device="cuda") for _ in range(G)]
        ys = (ctypes.c_void_p * G)(*[o.data_ptr() for o in outs])
        rc = lib.bnn_conv2d_tc_grouped(d_x.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), ys, G, relu_mask, center, 1, N, HW, HW,
                                       Cin, 128, 3, 2, stream())
        assert rc == 0, lib.bnn_last_error()
        torch.cuda.synchronize()
        return [o.cpu() for o in outs]
    got = run()
    monkeypatch.setenv("BNN_TC_NO_SCG2", "1")
    ref = run()
    xd = x.double()
    for i in range(G):
        want = F.conv2d(xd, ws[i].double(), bias[i * 128:(i + 1) * 128].double(), 2, 1)
        if (relu_mask >> i) & 1:
            want = want.relu()
        err = (got[i].permute(0, 3, 1, 2).double() - want).abs().max().item()
        assert err <= 1e-3 * max(1.0, want.abs().max().item()), (i, err)
        assert torch.equal(got[i], ref[i]) and not torch.isnan(got[i]).any()


def test_conv_tc_tap_skip_on_1x1_maps_is_bit_exact(lib, monkeypatch):
    """3x3 pad-1 convolutions on 1x1 maps run the centre tap only; identical bits to the nine-tap form."""
    shape = (300, 1, 1, 512, 512, 3, 1, 1, True, True)
    got, want = _conv_case(lib, "tc", "fp16", *shape)
    monkeypatch.setenv("BNN_TC_NO_TAP_SKIP", "1")
    ref, _ = _conv_case(lib, "tc", "fp16", *shape)
    assert torch.equal(got, ref)
    assert (got.double() - want).abs().max().item() <= 1e-3 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("H,N,Cin,C,stride,with_res,pair", [(4, 40, 512, 512, 1, True, "0"), (8, 12, 256, 512, 2, False, "0"),
                                                           (2, 70, 256, 256, 1, False, "0"), (4, 41, 512, 512, 1, True, "cg2")])
def test_conv_tc_fused_average_pool(lib, H, N, Cin, C, stride, with_res, pair, monkeypatch):
    """bnn_conv2d_tc_pooled == mean over the OH x OW map of bnn_conv2d_tc's output (`F.avg_pool2d(F.relu(out), k)`,
    resnet18.py:309,:339), incl. a last tile that ends inside the warp."""
    if pair != "0":
        monkeypatch.setenv("BNN_TC_MC_MIN_TILES", "1")
    g = torch.Generator().manual_seed(H + N)
    x = torch.randn(N, H, H, Cin, generator=g).half().cuda()
    w = (torch.randn(C, 3, 3, Cin, generator=g) / np.sqrt(9 * Cin)).half().cuda()
    b = torch.randn(C, generator=g).cuda()
    OH = H // stride
    res = torch.randn(N, OH, OH, C, generator=g).half().cuda() if with_res else None
    full = torch.empty(N, OH, OH, C, dtype=torch.float16, device="cuda")
    pooled = torch.full((N, C), float("nan"), dtype=torch.float16, device="cuda")
    dd = drop_desc(batch=N)
    rp = ctypes.c_void_p(res.data_ptr() if with_res else 0)
    assert lib.bnn_conv2d_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), rp, full.data_ptr(), 1, N, H, H, Cin, C, 3, stride, 1,
                             ctypes.byref(dd), stream()) == 0, lib.bnn_last_error()
    assert lib.bnn_conv2d_tc_pooled(x.data_ptr(), w.data_ptr(), b.data_ptr(), rp, pooled.data_ptr(), 1, N, H, H, Cin, C, 3,
                                    stride, 1, stream()) == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    want = full.float().mean(dim=(1, 2))
    err = (pooled.float() - want).abs().max().item()
    # the fused form averages the fp32 accumulators, the reference form the fp16-rounded map: <= 1 fp16 ulp apart
    assert not torch.isnan(pooled).any() and err <= 2e-3 * max(1.0, want.abs().max().item())
    # maps that are not 2..32 pixels (power of two) are rejected
    assert lib.bnn_conv2d_tc_pooled(x.data_ptr(), w.data_ptr(), b.data_ptr(), rp, pooled.data_ptr(), 1, 1, 16, 16, Cin, C, 3,
                                    1, 1, stream()) == -1


@pytest.mark.parametrize("N,Cin,res,drop_kind", [(5, 128, True, 0), (200, 128, False, 0), (181, 128, True, 1), (150, 64, False, 3),
                                                 (2, 256, False, 0)])
def test_conv_tc_vertical_halo_form(lib, N, Cin, res, drop_kind, monkeypatch):
    """Operand-swapped kernel, vertical-halo form (3x3 stride-1, 16 x 16 maps, Cout = 128): one haloed 18-row tile per
    (kw, channel block) serves the three vertical taps through descriptor offsets.  vs float64 torch on the rounded
    operands, incl. more tiles than SMs (ring wrap-around, persistent loop), residual, fused element dropout /
    Masksembles epilogues - and vs the one-box-per-tap kernel (same values up to fp32 summation order)."""
    masks = (torch.rand(4, 128, generator=torch.Generator().manual_seed(1)) > 0.5).float().cuda()
    B = N
    dd = drop_desc(drop_kind, 0.5, 0x99, 3, 2, B, masks if drop_kind == 3 else None, cnt0=1) if drop_kind else None
    shape = (N, 16, 16, Cin, 128, 3, 1, 1, True, res)
    got, want = _conv_case(lib, "tc", "fp16", *shape, drop=dd)
    if drop_kind == 1:
        want = want * torch.from_numpy(philox.keep_mask(0x99, 3, 2, (B, 128, 16, 16), 0.5)).double() * 2.0
    elif drop_kind == 3:
        want = want * masks[(1 + 2) % 4].cpu().view(1, 128, 1, 1).double()
    err = (got.double() - want).abs().max().item()
    scale = max(1.0, want.abs().max().item())
    monkeypatch.setenv("BNN_TC_NO_VH", "1")
    ref, _ = _conv_case(lib, "tc", "fp16", *shape, drop=dd)
    d_old = (got.double() - ref.double()).abs().max().item()
    report(test="conv_tc_vertical_halo", N=N, Cin=Cin, res=res, drop=drop_kind, err=err, vs_per_tap_kernel=d_old, scale=scale)
    assert err <= 1e-3 * scale and d_old <= 2e-3 * scale and not torch.isnan(got).any()


@pytest.mark.parametrize("pair", ["single", "cg2"])
@pytest.mark.parametrize("shape,drop_kind", [((300, 4, 4, 512, 512, 3, 1, 1, True, True), 0), ((260, 8, 8, 256, 256, 3, 1, 1, True, False), 1),
                                             ((257, 8, 8, 256, 512, 3, 2, 1, True, False), 0), ((130, 2, 2, 512, 512, 3, 1, 1, True, False), 2),
                                             ((200, 4, 4, 256, 256, 3, 2, 1, False, False), 0),
                                             ((150, 2, 8, 256, 256, 3, 1, 1, True, True), 0), ((131, 8, 4, 256, 512, 3, 1, 1, True, False), 1)])
def test_conv_tc_position_major_tiles_skip_padding_taps(lib, shape, drop_kind, pair, monkeypatch):
    """Position-major tiling (one output position x 128 images per row-tile): taps that read only zero padding are not
    issued.  Skipped terms are exact zeros, so the result must be BIT-IDENTICAL to the pixel-major kernel - plain and
    strided convolutions, residual, fused element / channel dropout, ragged image blocks, single CTA and cta_group::2."""
    monkeypatch.setenv("BNN_TC_PM_MIN_IMAGES", "1")
    if pair == "cg2":
        monkeypatch.setenv("BNN_TC_MC_MIN_TILES", "1")
    else:
        monkeypatch.setenv("BNN_TC_NOMC", "1")
    N = shape[0]
    B = N // 2 if N % 2 == 0 else N
    dd = drop_desc(drop_kind, 0.5, 0x31, 4, 1, B) if drop_kind else None
    got, want = _conv_case(lib, "tc", "fp16", *shape, drop=dd)
    monkeypatch.setenv("BNN_TC_NO_PM", "1")
    ref, _ = _conv_case(lib, "tc", "fp16", *shape, drop=dd)
    assert torch.equal(got, ref) and not torch.isnan(got).any()
    if drop_kind == 0:
        assert (got.double() - want).abs().max().item() <= 1e-3 * max(1.0, want.abs().max().item())


def test_conv_tc_position_major_grouped_and_shortcut(lib, monkeypatch):
    """... and the grouped stride-2 siblings (two 512-channel groups, 8x8 -> 4x4) and the fused projection shortcut."""
    monkeypatch.setenv("BNN_TC_PM_MIN_IMAGES", "1")
    monkeypatch.setenv("BNN_TC_MC_MIN_TILES", "1")
    g = torch.Generator().manual_seed(4)
    N, Cin, HW, Cg, G = 140, 256, 8, 512, 2
    x = torch.randn(N, HW, HW, Cin, generator=g).half().cuda()
    w = (torch.randn(G * Cg, 3, 3, Cin, generator=g) / np.sqrt(9 * Cin)).half().cuda()
    b = torch.randn(G * Cg, generator=g).cuda()

    def grouped():
        outs = [torch.full((N, HW // 2, HW // 2, Cg), float("nan"), dtype=torch.half, device="cuda") for _ in range(G)]
        ys = (ctypes.c_void_p * G)(*[o.data_ptr() for o in outs])
        assert lib.bnn_conv2d_tc_grouped(x.data_ptr(), w.data_ptr(), b.data_ptr(), ys, G, 1, 0, 1, N, HW, HW, Cin, Cg, 3, 2,
                                         stream()) == 0, lib.bnn_last_error()
        torch.cuda.synchronize()
        return outs
    h = torch.randn(N, 4, 4, 512, generator=g).half().cuda()
    wsc = (torch.randn(512, 9 * 512 + Cin, generator=g) / 70).half().cuda()
    dd = drop_desc(batch=N)

    def shortcut():
        y = torch.full((N, 4, 4, 512), float("nan"), dtype=torch.half, device="cuda")
        assert lib.bnn_conv2d_tc_shortcut(h.data_ptr(), wsc.data_ptr(), b.data_ptr(), None, y.data_ptr(), 1, N, 4, 4, 512, 512, 3, 1, 1,
                                          ctypes.byref(dd), x.data_ptr(), HW, HW, Cin, stream()) == 0, lib.bnn_last_error()
        torch.cuda.synchronize()
        return y
    a, c = grouped(), shortcut()
    monkeypatch.setenv("BNN_TC_NO_PM", "1")
    a0, c0 = grouped(), shortcut()
    for u, v in zip(a, a0):
        assert torch.equal(u, v) and not torch.isnan(u).any()
    assert torch.equal(c, c0) and not torch.isnan(c).any()


@pytest.mark.parametrize("Cin,Cout,HW", [(16, 64, 32), (32, 128, 8), (48, 256, 8), (16, 128, 16)])
def test_conv_tc_narrow_input_one_kstep_per_tap(lib, Cin, Cout, HW):
    """Inputs narrower than one 64-channel k-block (the stem: 3 channels stored as 16): one k-block per tap of which only
    Cin / 16 MMA k-steps are issued; the TMA unit zero-fills the box columns past Cin.  vs float64 torch, all tile shapes
    (64 / 128-swapped / 256-channel tiles), ragged image count."""
    shape = (37, HW, HW, Cin, Cout, 3, 1, 1, True, False)
    got, want = _conv_case(lib, "tc", "fp16", *shape)
    err = (got.double() - want).abs().max().item()
    report(test="conv_tc_narrow_input", Cin=Cin, Cout=Cout, err=err)
    assert err <= 1e-3 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("dt,C,F_,HW,kind,p,S,B,has_samples", [
    ("fp16", 10, 512, 1, 1, 0.5, 32, 40, 1), ("fp16", 10, 512, 16, 2, 0.25, 5, 7, 1), ("bf16", 32, 256, 4, 3, 0.0, 6, 9, 1),
    ("fp32", 17, 264, 1, 1, 0.125, 9, 33, 1), ("fp16", 10, 512, 1, 0, 0.0, 3, 300, 1), ("fp16", 10, 64, 4, 1, 0.5, 4, 6, 0)])
def test_exit_head_rows_matches_block_per_image_head(lib, dt, C, F_, HW, kind, p, S, B, has_samples):
    """bnn_exit_head_rows (one warp per (sample, image) row, then the sample-ordered soft-max kernel) == bnn_exit_head on the
    same features: element / channel dropout, Masksembles rows, no site, pooled and un-pooled maps, features shared by the
    samples, fp32 / fp16 / bf16 storage, ragged row counts; the sums start from the accumulate flag."""
    tdt, code = TORCH_DT[dt]
    s0, seed, sid = 3, 0x99, 2
    g = torch.Generator().manual_seed(C + S + B)
    feat = (torch.randn((S if has_samples else 1) * B, HW, F_, generator=g).abs() * 3).to(tdt).cuda()
    w = (torch.randn(C, F_, generator=g) / np.sqrt(F_) * 3).cuda()
    bias = torch.randn(C, generator=g).cuda()
    masks = (torch.rand(4, F_, generator=g) > 0.5).float().cuda()
    dd = drop_desc(kind, p, seed, sid, s0, B, masks if kind == 3 else None, cnt0=1)
    l_ws = torch.empty(S * B * C, device="cuda")
    outs = {}
    for name in ("block", "rows"):
        o = dict(sp=torch.full((B, C), 0.5).cuda(), sl=torch.full((B, C), -1.0).cuda(), spl=torch.full((B,), 2.0).cuda(),
                 lo=torch.zeros(S, B, C).cuda())
        for acc in (0, 1):                          # second call accumulates on top of the first
            if name == "block":
                wt = w.t().contiguous()
                rc = lib.bnn_exit_head(feat.data_ptr(), code, has_samples, B, S, HW, F_, C, wt.data_ptr(), bias.data_ptr(),
                                       ctypes.byref(dd), o["sp"].data_ptr(), o["sl"].data_ptr(), o["spl"].data_ptr(),
                                       o["lo"].data_ptr(), acc, stream())
            else:
                rc = lib.bnn_exit_head_rows(feat.data_ptr(), code, has_samples, B, S, HW, F_, C, w.data_ptr(), bias.data_ptr(),
                                            ctypes.byref(dd), l_ws.data_ptr(), o["sp"].data_ptr(), o["sl"].data_ptr(),
                                            o["spl"].data_ptr(), o["lo"].data_ptr(), acc, stream())
            assert rc == 0, lib.bnn_last_error()
        torch.cuda.synchronize()
        outs[name] = o
    scale = max(1.0, outs["block"]["lo"].abs().max().item())
    e_lo = (outs["rows"]["lo"] - outs["block"]["lo"]).abs().max().item()
    e_p = (outs["rows"]["sp"] - outs["block"]["sp"]).abs().max().item()
    e_l = (outs["rows"]["sl"] - outs["block"]["sl"]).abs().max().item()
    e_pl = (outs["rows"]["spl"] - outs["block"]["spl"]).abs().max().item()
    report(test="exit_head_rows", dtype=dt, C=C, F=F_, HW=HW, kind=kind, err_logits=e_lo, err_sum_p=e_p, scale=scale)
    assert e_lo <= 2e-5 * scale and e_p <= 2e-5 * S and e_l <= 4e-5 * S * scale and e_pl <= 2e-4 * S
    assert lib.bnn_exit_head_rows(feat.data_ptr(), code, has_samples, B, S, HW, F_, 33, w.data_ptr(), bias.data_ptr(),
                                  ctypes.byref(dd), l_ws.data_ptr(), o["sp"].data_ptr(), o["sl"].data_ptr(), o["spl"].data_ptr(),
                                  None, 0, stream()) == -1          # wide heads go to bnn_exit_head_tc / bnn_exit_head_mma
