"""GPU: the mask contract is BIT-EXACT between the device Philox, the host restatement and therefore
the masks injected into the oracle; stand-alone stochastic layers (drop-ins for MCDropout,
BayesianDropout*, Masksembles1D/2D) against the oracle's masks."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import philox
from tests.gpu_util import TORCH_DT, drop_desc, stream

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("count", [0, 1, 3, 4, 5, 1003, 1 << 16])
def test_philox_words_bit_exact(lib, count):
    for seed, sid, sample in [(0, 0, 0), (0x5EED, 3, 17), (0xDEADBEEFCAFEF00D, 0xFFFFFFFF, 0xFFFFFFFE)]:
        out = torch.zeros(max(count, 1), dtype=torch.int32, device="cuda")
        assert lib.bnn_philox_words(out.data_ptr(), count, seed, sid, sample, stream()) == 0
        got = out.cpu().numpy().view(np.uint32)[:count]
        assert np.array_equal(got, philox.random_words(seed, sid, sample, count))


@pytest.mark.parametrize("p", [0.0, 0.125, 0.25, 0.375, 0.5, 0.9, 1.0])
def test_philox_keep_bit_exact(lib, p):
    n = 40001
    out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    assert lib.bnn_philox_keep(out.data_ptr(), n, p, 0x5EED, 2, 5, stream()) == 0
    want = philox.keep_mask_flat(0x5EED, 2, 5, n, p)
    assert np.array_equal(out.cpu().numpy().astype(bool), want)
    assert lib.bnn_philox_keep(out.data_ptr(), n, 1.5, 0, 0, 0, stream()) == -1      # ValueError class
    assert b"between 0 and 1" in lib.bnn_last_error()


@pytest.mark.parametrize("dt", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 8, 5, 7), (3, 64, 4, 4), (1, 3, 1, 1), (4, 10)])
def test_mcdropout_module_matches_injected_masks(dt, shape):
    from bayesnn_fpga_b200.Dropouts import MCDropout
    tdt, _ = TORCH_DT[dt]
    torch.manual_seed(1)
    x = torch.randn(*shape).to(tdt)
    m = MCDropout(0.25).reseed(seed=77, stream_id=9)
    for sample in range(3):                                       # successive calls = successive samples
        got = m(x.cuda())
        keep = torch.from_numpy(philox.keep_mask(77, 9, sample, shape, 0.25))
        want = (x.float() * keep * (1.0 / 0.75)).to(tdt)
        assert got.dtype == tdt and got.shape == x.shape
        assert torch.equal(got.cpu(), want)                        # one multiply, one rounding: exact


def test_dropout_p0_p1_and_empty():
    from bayesnn_fpga_b200.Dropouts import MCDropout
    x = torch.randn(2, 16, 3, 3, device="cuda")
    assert torch.equal(MCDropout(0.0)(x), x)
    assert torch.equal(MCDropout(1.0)(x), torch.zeros_like(x))      # F.dropout(p=1) -> zeros
    assert MCDropout(0.5)(torch.zeros(0, 16, 3, 3, device="cuda")).shape == (0, 16, 3, 3)


@pytest.mark.parametrize("dt", ["fp32", "fp16"])
@pytest.mark.parametrize("has_samples", [0, 1])
@pytest.mark.parametrize("p", [0.5, 0.25])
def test_prefix_broadcast_kernel(lib, dt, has_samples, p):
    """bnn_dropout: y[s] = x[(s)] * mask_s, one launch for all local samples, global sample index.  p = 0.5 takes the
    top-bit (PRMT) mask path of the 16-bit kernels, p = 0.25 the two-halfword compare; both with a broadcast source and with
    a source that already has a sample dimension."""
    tdt, code = TORCH_DT[dt]
    B, H, W, C, S, s0 = 3, 5, 6, 16, 4, 10
    torch.manual_seed(0)
    x = torch.randn((S if has_samples else 1) * B, H, W, C).to(tdt).cuda()
    y = torch.empty(S * B, H, W, C, dtype=tdt, device="cuda")
    d = drop_desc(1, p, 0xABC, 6, s0, B)
    assert lib.bnn_dropout(x.data_ptr(), y.data_ptr(), code, H * W * C, C, S, has_samples, ctypes.byref(d), stream()) == 0
    y = y.cpu().view(S, B, H, W, C)
    xs = x.cpu().view(-1, B, H, W, C)
    for s in range(S):
        keep = philox.keep_mask(0xABC, 6, s0 + s, (B, C, H, W), p)            # NCHW view of the NHWC contract
        want = (xs[s if has_samples else 0].float() * torch.from_numpy(keep).permute(0, 2, 3, 1) * (1.0 / (1.0 - p))).to(tdt)
        assert torch.equal(y[s], want)


def test_bayesian_dropout2d_is_channel_wise():
    from bayesnn_fpga_b200.Dropouts import BayesianDropout, BayesianDropout2D
    torch.manual_seed(0)
    conv = nn.Conv2d(3, 8, 3, padding=1).cuda()
    x = torch.randn(4, 3, 6, 6, device="cuda")
    m = BayesianDropout2D(conv, p=0.5).reseed(5, 1)
    got = m(x)
    keep = torch.from_numpy(philox.keep_mask(5, 1, 0, (4, 8, 6, 6), 0.5, mode="channel")).cuda()
    assert torch.equal(got, conv(x) * keep * 2.0)
    lin = nn.Linear(6, 5).cuda()
    m = BayesianDropout(lin, p=0.125).reseed(5, 2)
    z = torch.randn(7, 6, device="cuda")
    keep = torch.from_numpy(philox.keep_mask(5, 2, 0, (7, 5), 0.125)).cuda()
    assert torch.equal(m(z), lin(z) * keep * np.float32(1.0 / 0.875))


def test_masksembles_eval_rotation_and_train_groups():
    from bayesnn_fpga_b200.utils import Masksembles1D, Masksembles2D
    np.random.seed(0)
    m2 = Masksembles2D(64, 4, 2.0).cuda().eval()
    x = torch.randn(8, 64, 3, 3, device="cuda")
    for call in range(6):                                          # cnt rotates 0,1,2,3,0,1 (utils.py:168)
        got = m2(x)
        assert got.dtype == torch.float32
        assert torch.equal(got, x * m2.masks[call % 4].view(1, -1, 1, 1))
    assert m2.cnt == 2
    m2.train()
    got = m2(x)                                                    # group g of B/n images uses mask g
    want = torch.cat([x[2 * g:2 * g + 2] * m2.masks[g].view(1, -1, 1, 1) for g in range(4)])
    assert torch.equal(got, want)
    np.random.seed(1)
    m1 = Masksembles1D(512, 4, 2.0).cuda().eval()
    z = torch.randn(5, 512, device="cuda")
    assert torch.equal(m1(z), z * m1.masks[0]) and torch.equal(m1(z), z * m1.masks[1])


def test_converter_wrapper_eval_mean():
    """nn2bnn.MCDropout eval = mean over nSamples stochastic passes of the raw output (nn2bnn.py:26-27)."""
    from bayesnn_fpga_b200 import nn2bnn
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(6, 16), nn.ReLU(), nn.Linear(16, 3)).cuda()
    ref = [nn.Linear(6, 16).cuda(), nn.Linear(16, 3).cuda()]
    ref[0].load_state_dict(net[0].state_dict()); ref[1].load_state_dict(net[2].state_dict())
    bnn = nn2bnn.MCDropout(net, nSamples=4, p=0.5).reseed(123).eval()
    x = torch.randn(5, 6, device="cuda")
    got = bnn(x)
    acc = 0
    for s in range(4):
        k0 = torch.from_numpy(philox.keep_mask(123, 0, s, (5, 16), 0.5)).cuda()
        k1 = torch.from_numpy(philox.keep_mask(123, 1, s, (5, 3), 0.5)).cuda()
        acc = acc + ref[1](F.relu(ref[0](x) * k0 * 2.0)) * k1 * 2.0
    assert torch.allclose(got, acc / 4, atol=1e-6)
