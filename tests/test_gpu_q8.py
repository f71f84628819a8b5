"""GPU: the 8-bit fixed-point path (reference: QKeras layers with quantized_bits(8, ibit) /
quantized_relu(8), t_qmodels_bayes_me.py:49-52).  Integer work: the bar is BIT-EXACT against the NumPy integer
restatement (oracle/q8.py) - int32 accumulation on tcgen05 kind::i8 is exact and the requantising epilogue is
restated operation for operation."""
import ctypes

import numpy as np
import pytest
import torch

from bayesnn_fpga_b200 import _lib, q8
from oracle import philox
from oracle import q8 as oq8
from tests.gpu_util import drop_desc, report, stream

pytestmark = pytest.mark.gpu

SHAPES = [
    # N, H, Cin, Cout, k, stride
    (6, 8, 256, 256, 3, 1),        # layer3-like, BN = 256
    (70, 4, 512, 512, 3, 1),       # layer4-like: 8 images per tile, two channel tiles, ragged last row-tile
    (5, 16, 128, 128, 3, 1),       # Cout = 128: two 128 x 128 MMAs per k-block
    (4, 8, 128, 64, 1, 1),         # 1x1, Cout = 64
    (3, 16, 128, 256, 3, 2),       # stride 2 through the parity planes
    (9, 8, 384, 256, 3, 1),        # three k-blocks per tap
    (300, 2, 512, 512, 3, 1),      # VGG block 4 (2x2 maps): most taps are zero padding
]


def _case(N, H, Cin, Cout, k, stride, seed):
    rng = np.random.RandomState(seed)
    x = rng.randint(0, 256, size=(N, Cin, H, H)).astype(np.uint8)
    x[rng.rand(*x.shape) < 0.5] = 0                                   # post-ReLU / post-dropout activations are sparse
    w = rng.randint(-128, 128, size=(Cout, Cin, k, k)).astype(np.int8)
    bias_q = (rng.randn(Cout) * 20).astype(np.float32)
    q_mult = np.float32(2.0 ** -(8 + int(np.log2(Cin * k * k)) // 2))    # keeps most outputs inside [0, 255]
    return x, w, bias_q, float(q_mult)


def _run(lib, x, w, bias_q, q_mult, k, stride, drop=None):
    N, Cin, H, _ = x.shape
    Cout = w.shape[0]
    pad = 1 if k == 3 else 0
    OH = (H + 2 * pad - k) // stride + 1
    d_x = torch.from_numpy(x).permute(0, 2, 3, 1).contiguous().cuda()
    d_w = torch.from_numpy(w).permute(0, 2, 3, 1).contiguous().cuda()
    d_b = torch.from_numpy(bias_q).cuda()
    d_y = torch.full((N, OH, OH, Cout), 77, dtype=torch.uint8, device="cuda")
    rc = lib.bnn_conv2d_tc_i8(d_x.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), d_y.data_ptr(), N, H, H, Cin, Cout, k, stride,
                              q_mult, ctypes.byref(drop) if drop is not None else None, stream())
    assert rc == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    return d_y.cpu().permute(0, 3, 1, 2).numpy()


@pytest.mark.parametrize("pair", ["single", "cg2"])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv_i8_bit_exact_vs_integer_oracle(lib, shape, pair, monkeypatch):
    N, H, Cin, Cout, k, stride = shape
    if pair == "cg2":
        if Cout % 256:
            pytest.skip("cta_group::2 kernel: 256-channel tiles")
        monkeypatch.setenv("BNN_TC_MC_MIN_TILES", "1")
    else:
        monkeypatch.setenv("BNN_TC_NOMC", "1")
    x, w, bias_q, q_mult = _case(*shape, seed=Cin + Cout + N)
    got = _run(lib, x, w, bias_q, q_mult, k, stride)
    want = oq8.qconv_relu(x, w, bias_q, q_mult, stride, 1 if k == 3 else 0)
    frac_interior = float(((want > 0) & (want < 255)).mean())
    report(test="conv_i8", shape=list(shape), pair=pair, mismatches=int((got != want).sum()), interior=frac_interior)
    assert frac_interior > 0.05                                        # the case really exercises the rounding
    assert np.array_equal(got, want)


@pytest.mark.parametrize("pair", ["single", "cg2"])
@pytest.mark.parametrize("shape", [(300, 4, 512, 512, 3, 1), (140, 8, 256, 256, 3, 1), (150, 8, 256, 512, 3, 2)])
def test_conv_i8_position_major_tiles_bit_exact(lib, shape, pair, monkeypatch):
    """position-major tiling (padding-only taps skipped, conv_tc.cu Params::pm_nb2) in the integer kernel: still the
    oracle's integers, at ragged image counts"""
    monkeypatch.setenv("BNN_TC_PM_MIN_IMAGES", "1")
    monkeypatch.setenv("BNN_TC_MC_MIN_TILES" if pair == "cg2" else "BNN_TC_NOMC", "1")
    N, H, Cin, Cout, k, stride = shape
    x, w, bias_q, q_mult = _case(*shape, seed=N)
    got = _run(lib, x, w, bias_q, q_mult, k, stride)
    assert np.array_equal(got, oq8.qconv_relu(x, w, bias_q, q_mult, stride, 1))


@pytest.mark.parametrize("kind", [1, 2])
def test_conv_i8_fused_dropout_bit_exact(lib, kind):
    """a stochastic site fused behind the ReLU: Philox element / channel masks, applied in float before requantisation"""
    B, S, H, Cin, Cout = 4, 3, 8, 256, 256
    x, w, bias_q, q_mult = _case(B * S, H, Cin, Cout, 3, 1, seed=5)
    dd = drop_desc(kind, 0.5, 0x51, 7, 2, B)
    got = _run(lib, x, w, bias_q, q_mult, 3, 1, drop=dd)
    keep = np.concatenate([philox.keep_mask(0x51, 7, 2 + s, (B, Cout, H, H), 0.5, "channel" if kind == 2 else "element")
                           for s in range(S)])
    want = oq8.qconv_relu(x, w, bias_q, q_mult, 1, 1, keep_scale=keep.astype(np.float32) * np.float32(2.0))
    assert np.array_equal(got, want)


def test_qkeras_style_layer_stack_bit_exact(lib):
    """Three QConv2d layers quantised like the reference's QKeras layers (quantized_bits(8, ibit) kernels and biases,
    quantized_relu(8) activations) run end to end on integers: every uint8 activation equals the NumPy restatement."""
    g = torch.Generator().manual_seed(3)
    x = torch.rand(5, 128, 16, 16, generator=g)                                   # inputs in [0, 1)
    specs = [(128, 256, 3, 2, 0.08), (256, 256, 3, 1, 0.05), (256, 512, 1, 1, 0.1)]
    xt = q8.QTensor.from_float_nchw(x)
    xq, x_step = oq8.quantized_relu_int(x.numpy())
    assert np.array_equal(xt.q.cpu().permute(0, 3, 1, 2).numpy(), xq.astype(np.uint8))
    for cin, cout, k, stride, std in specs:
        w = torch.randn(cout, cin, k, k, generator=g) * std
        b = torch.randn(cout, generator=g) * 0.1
        layer = q8.QConv2d(w, b, stride=stride, w_integer=0, out_integer=1)
        wq, w_step = oq8.quantized_bits_int(w.numpy(), 8, 0)
        bq, _ = oq8.quantized_bits_int(b.numpy(), 8, 0)
        assert np.array_equal(layer.w_q.numpy(), wq) and layer.w_step == w_step
        out_step = 2.0 ** (1 - 8)
        xt = layer(xt)
        xq = oq8.qconv_relu(xq, wq, (bq * w_step / out_step).astype(np.float32), np.float32(w_step * x_step / out_step),
                            stride, 1 if k == 3 else 0).astype(np.int32)
        x_step = out_step
        got = xt.q.cpu().permute(0, 3, 1, 2).numpy()
        assert np.array_equal(got, xq.astype(np.uint8)) and xt.step == x_step
        assert 0.02 < (got > 0).mean() < 0.98


def test_conv_i8_rejects_what_it_does_not_implement(lib):
    y = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    b = torch.zeros(512, device="cuda")
    args = lambda cin, cout, k: (y.data_ptr(), y.data_ptr(), b.data_ptr(), y.data_ptr(), 1, 8, 8, cin, cout, k, 1, 1.0, None,
                                 stream())
    assert lib.bnn_conv2d_tc_i8(*args(64, 256, 3)) == -4          # Cin must fill 128-byte k-blocks
    assert lib.bnn_conv2d_tc_i8(*args(128, 100, 3)) == -4
    assert lib.bnn_conv2d_tc_i8(*args(128, 256, 5)) == -4


def _np_maxpool(x, k):
    N, C, H, W = x.shape
    return x.reshape(N, C, H // k, k, W // k, k).max(axis=(3, 5))


def test_q8_suffix_plan_vgg_last3_bit_exact_and_close_to_fp16():
    """BASELINE config 4's network (multi-exit VGG-19, dropout on the last three blocks + exits) with its stochastic
    suffix in 8 bits (q8.Q8Plan): every uint8 activation tensor of the suffix equals the NumPy integer restatement
    bit for bit (same Philox masks, same prefix tensor), the mean probabilities agree with a float64 evaluation of the
    heads on those integers to 1e-5, and the 8-bit result stays close to the 16-bit engine's (reported)."""
    import bench
    from bayesnn_fpga_b200 import mc_predict
    model = bench.build_model("vgg_last3", 100).cuda()
    B, S, seed = 3, 2, 0x77
    g_ = torch.Generator().manual_seed(5)
    calib = torch.randn(8, 3, 32, 32, generator=g_)
    x = torch.randn(B, 3, 32, 32, generator=g_)
    plan = q8.Q8Plan(model, calib, S_calib=4, seed=1)
    r = plan.run(x, S, seed=seed)
    got_p = r.mean_probs.double().cpu().numpy()
    eng, g = plan.eng, plan.eng.graph
    st = eng._bufs[(B, S, False)]
    qb = plan._qbufs[(B, S)]
    nchw = lambda t: t.cpu().permute(0, 3, 1, 2).numpy()
    vals, n_checked, heads = {}, 0, {}
    for i, op in enumerate(g.ops):
        out_t = op.dst if op.dst is not None else op.src
        if not (op.kind == "site" or out_t.stoch):
            continue
        site = op.site
        keep = None
        if site is not None and op.kind != "head":
            shp = (B, op.dst.C, op.dst.H, op.dst.W)
            keep = np.concatenate([philox.keep_mask(seed, site.stream, s, shp, site.p) for s in range(S)]).astype(np.float32)
            keep = keep * np.float32(1.0 / (1.0 - site.p))
        if op.kind == "site":
            if op.src.stoch:
                xin, scale = vals[op.src.id].astype(np.float32), np.float32(plan.step[op.src.id] / plan.step[op.dst.id])
            else:
                xin = np.tile(nchw(st["acts"][op.src.id]).astype(np.float32), (S, 1, 1, 1))
                scale = np.float32(1.0 / plan.step[op.dst.id])
            y = np.clip(np.rint(((xin * scale).astype(np.float32) * keep).astype(np.float32)), 0, 255).astype(np.uint8)
        elif op.kind == "conv":
            L = plan.layers[i]
            y = oq8.qconv_relu(vals[op.src.id], L["w_q"].numpy(), L["d_bias_q"].cpu().numpy()[:op.dst.C], np.float32(L["q_mult"]),
                               op.stride, op.pad, keep_scale=keep)      # (integer sums: the 2x2-map GEMM form is bit-identical)
        elif op.kind == "maxpool":
            y = _np_maxpool(vals[op.src.id], op.pool_k)
        elif op.kind == "head":
            feat = vals[op.src.id].astype(np.float64).mean(axis=(2, 3)) * plan.step[op.src.id]           # [S*B, F]
            if site is not None:
                k = np.concatenate([philox.keep_mask(seed, site.stream, s, (B, op.src.C), site.p) for s in range(S)])
                feat = feat * k / (1.0 - site.p)
            logits = feat @ op.weight.double().numpy().T + op.bias.double().numpy()
            e = np.exp(logits - logits.max(1, keepdims=True))
            heads[op.exit_index] = (e / e.sum(1, keepdims=True)).reshape(S, B, -1).mean(0)
            continue
        vals[op.dst.id] = y
        assert np.array_equal(nchw(qb[op.dst.id]), y), (op.kind, op.name)
        assert 0.01 < (y > 0).mean() and y.max() > 64, (op.name, float((y > 0).mean()), int(y.max()))   # range is used
        n_checked += 1
    assert n_checked >= 12 and len(heads) >= 3
    for e, p in heads.items():
        assert np.abs(got_p[e] - p).max() <= 1e-5
    ref = mc_predict(model, x, S, seed=seed, dtype="fp16").mean_probs.double().cpu().numpy()
    err = float(np.abs(got_p - ref).max())
    report(test="q8_suffix_vgg_last3", tensors_bit_exact=n_checked, max_prob_diff_vs_fp16=err,
           argmax_same=float((got_p.argmax(-1) == ref.argmax(-1)).mean()))
    assert err < 0.1
    with pytest.raises(NotImplementedError):
        q8.Q8Plan(bench.build_model("resnet_mcd", 10).cuda(), calib)      # Cin = 64 consumers of the first site
