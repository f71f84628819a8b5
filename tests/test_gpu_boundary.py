"""GPU: the prefix boundary without the S masked copies.  bnn_boundary_bits writes one scaled copy, the keep bits of
every sample and the even/even plane; the sibling-pair kernel ANDs the bits into its tiles in shared memory; the fused
projection shortcut reads the plane.  Everything must reproduce the masked-copies path BIT FOR BIT."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bayesnn_fpga_b200 import _lib, mc_predict
from oracle import philox
from tests.cases import build_seeded
from tests.gpu_util import drop_desc, stream

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,S,G,p", [(3, 4, 2, 0.5), (20, 9, 2, 0.25), (5, 3, 4, 0.5)])
def test_boundary_bits_and_masked_sibling_pair(lib, B, S, G, p):
    H = W = 32
    C = 64
    g = torch.Generator().manual_seed(B * S)
    x = torch.randn(B, H, W, C, generator=g).relu().half().cuda()
    dd = drop_desc(1, p, 0xB0B, 5, 3, B)
    copies = torch.empty(S * B, H, W, C, dtype=torch.half, device="cuda")
    assert lib.bnn_dropout(x.data_ptr(), copies.data_ptr(), 1, H * W * C, C, S, 0, ctypes.byref(dd), stream()) == 0
    xs = torch.empty_like(x)
    bits = torch.zeros(S * B * H * W * C // 8, dtype=torch.uint8, device="cuda")
    plane = torch.empty(S * B, H // 2, W // 2, C, dtype=torch.half, device="cuda")
    assert lib.bnn_boundary_bits(x.data_ptr(), xs.data_ptr(), bits.data_ptr(), plane.data_ptr(), 1, B, H, W, C, S,
                                 ctypes.byref(dd), stream()) == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    scale = np.float32(1.0 / (1.0 - p))
    assert torch.equal(xs, (x.float() * float(scale)).half())
    assert torch.equal(plane, copies[:, ::2, ::2, :].contiguous())
    keep = np.stack([philox.keep_mask(0xB0B, 5, 3 + s, (B, C, H, W), p) for s in range(S)])          # [S, B, C, H, W]
    want_bits = np.packbits(keep.transpose(0, 1, 3, 4, 2).reshape(-1, 8), axis=1, bitorder="little").reshape(-1)
    assert np.array_equal(bits.cpu().numpy(), want_bits)
    # ---- stride-2 sibling groups: masked kernel on (x_scaled, bits) == plain grouped kernel on the masked copies
    ws = torch.cat([torch.randn(128, C, 3, 3, generator=g) / np.sqrt(9 * C) for _ in range(G)]).half()
    d_w = ws.permute(0, 2, 3, 1).contiguous().cuda()
    d_b = torch.randn(G * 128, generator=g).cuda()

    def outs():
        o = [torch.full((S * B, H // 2, W // 2, 128), float("nan"), dtype=torch.half, device="cuda") for _ in range(G)]
        return o, (ctypes.c_void_p * G)(*[t.data_ptr() for t in o])
    ref, ys = outs()
    assert lib.bnn_conv2d_tc_grouped(copies.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), ys, G, 0b0101, 0, 1, S * B, H, W, C, 128,
                                     3, 2, stream()) == 0
    got, ys2 = outs()
    rc = lib.bnn_conv2d_tc_grouped_masked(xs.data_ptr(), bits.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), ys2, G, 0b0101, 1,
                                          S * B, B, H, W, C, 128, stream())
    assert rc == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    for a, b in zip(got, ref):
        assert not torch.isnan(a).any() and torch.equal(a, b)
    # an odd group count / a centre-tap group cannot take this form
    assert lib.bnn_conv2d_tc_grouped_masked(xs.data_ptr(), bits.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), ys2, 3, 0, 1, S * B, B,
                                            H, W, C, 128, stream()) == -1


def test_shortcut_reads_the_even_even_plane(lib):
    """bnn_conv2d_tc_shortcut_plane (vertical-halo kernel, dense plane operand) == bnn_conv2d_tc_shortcut on the copies."""
    N, C2, C, H = 7, 64, 128, 16
    g = torch.Generator().manual_seed(2)
    h = torch.randn(N, H, H, C, generator=g).half().cuda()
    x2 = torch.randn(N, 2 * H, 2 * H, C2, generator=g).half().cuda()
    w = torch.cat([(torch.randn(C, 3, 3, C, generator=g) / np.sqrt(9 * C)).reshape(C, -1),
                   torch.randn(C, C2, generator=g) / np.sqrt(C2)], dim=1).half().contiguous().cuda()
    b = torch.randn(C, generator=g).cuda()
    dd = drop_desc(batch=N)
    y0 = torch.empty(N, H, H, C, dtype=torch.half, device="cuda")
    y1 = torch.full_like(y0, float("nan"))
    assert lib.bnn_conv2d_tc_shortcut(h.data_ptr(), w.data_ptr(), b.data_ptr(), None, y0.data_ptr(), 1, N, H, H, C, C, 3, 1, 1,
                                      ctypes.byref(dd), x2.data_ptr(), 2 * H, 2 * H, C2, stream()) == 0
    plane = x2[:, ::2, ::2, :].contiguous()
    rc = lib.bnn_conv2d_tc_shortcut_plane(h.data_ptr(), w.data_ptr(), b.data_ptr(), None, y1.data_ptr(), 1, N, H, H, C, C, 3, 1, 1,
                                          ctypes.byref(dd), plane.data_ptr(), C2, stream())
    assert rc == 0, lib.bnn_last_error()
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)


def test_engine_boundary_fusion_is_bit_identical(monkeypatch):
    """The C2 network end to end: no S masked copies of the layer-1 tensor, no stand-alone broadcast launch - and the
    same per-sample logits, bit for bit, as the masked-copies plan."""
    model, sd, gold = build_seeded("resnet18_mcd_block")
    model.cuda()
    x = torch.from_numpy(gold["x"]).repeat(3, 1, 1, 1)            # 12 images
    S = 5
    monkeypatch.setenv("BNN_BOUNDARY_FUSION", "0")
    eng0 = model.bnn_engine("fp16", rebuild=True)
    assert not eng0.boundary
    r0 = eng0.run(x, S, seed=9, want_logits=True, use_graph=False)
    ref = r0.all_logits.clone()
    monkeypatch.setenv("BNN_BOUNDARY_FUSION", "1")
    eng1 = model.bnn_engine("fp16", rebuild=True)
    assert len(eng1.boundary) == 1
    r1 = eng1.run(x, S, seed=9, want_logits=True, use_graph=False)
    assert torch.equal(r1.all_logits, ref)
    names = [o["name"] for o in eng1.profile_step(x.cuda(), S)]
    assert any("keep bits applied in shared memory" in n for n in names) and any("keep bits + shortcut plane" in n for n in names)
    st = eng1._bufs[(12, S, False)]
    t = next(iter(eng1.boundary))
    assert t not in st["acts"]                                       # the S copies are never allocated
