"""GPU: BASELINE.json's full sizes.

* `test_full_size_rows_vs_oracle`: every BASELINE config (C2 S=32 B=256, C3 Masksembles S=4 B=256 C=100, C4 VGG-19
  S=64 B=512) runs at its REAL batch size and sample count, and rows from the head, the middle (a tile / cluster-pair
  boundary) and the tail of the batch are compared with the oracle evaluated on just those images: Philox element
  indices are per image within the batch, so the oracle injects the masks of batch position `batch_offset + i`
  (oracle.philox.keep_mask).  This exercises the tile-straddling, grid-stride and `cta_group::2` tail paths at
  M = 8 192 images against the reference's arithmetic, not only against properties.
* size-independent properties of the path: normalisation, determinism, shard invariance (global Philox sample index),
  exit-ensemble definition, Monte-Carlo consistency of the mask stream, and agreement of the tensor-core path with the
  exact fp32 path on the same masks."""
import numpy as np
import pytest
import torch

from bayesnn_fpga_b200 import mc_predict, resnet18, vgg19

pytestmark = pytest.mark.gpu

# (bench.py workload, oracle case of tests/cases.py, p, rows checked: (first image, count))
FULL = {
    "c2": ("resnet18_mcd_block", 0.5, [(0, 6), (125, 6), (250, 6)]),
    "c3": ("resnet18_mask_block", 0.0, [(0, 6), (125, 6), (250, 6)]),
    "c4": ("vgg19_mcd_last3", 0.5, [(0, 4), (253, 6), (508, 4)]),
}


@pytest.mark.parametrize("wl", sorted(FULL))
def test_full_size_rows_vs_oracle(wl):
    import bench
    from tests.cases import oracle_run
    from tests.gpu_util import report
    tag, p, rows = FULL[wl]
    _, kind, B, S, classes = bench.WORKLOADS[wl]
    model = bench.build_model(kind, classes)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.cuda()
    x = torch.randn(B, *bench.input_shape(kind), generator=torch.Generator().manual_seed(100))
    seed = 0x5EED
    want = []
    for first, n in rows:                      # the oracle sees ONLY these images, with the masks of their batch slots
        want.append(oracle_run(tag, sd, x[first:first + n], S, seed, p, batch_offset=first))
    cat = lambda k: np.concatenate([w[k] for w in want], axis=1)
    idx = np.concatenate([np.arange(f, f + n) for f, n in rows])
    scale = max(1.0, float(np.abs(cat("mean_logits")).max()))
    for dt, ptol, ltol in (("fp16", 1e-3, 1e-3), ("fp32", 1e-5, 1e-5)):
        for m in model.modules():              # Masksembles counters: every run starts from the oracle's cnt0 = 0
            if hasattr(m, "cnt"):
                m.cnt = 0
        r = mc_predict(model, x, S, seed=seed, dtype=dt)
        got_p = r.mean_probs.double().cpu().numpy()[:, idx]
        got_l = r.mean_logits.double().cpu().numpy()[:, idx]
        got_e = r.ens_probs.double().cpu().numpy()[:, idx]
        perr = float(np.abs(got_p - cat("mean_probs")).max())
        lerr = float(np.abs(got_l - cat("mean_logits")).max())
        eerr = float(np.abs(got_e - cat("ens_probs")).max())
        top2 = np.sort(cat("mean_probs"), axis=-1)[..., -2:]
        clear = (top2[..., 1] - top2[..., 0]) > 2 * ptol
        same = bool((got_p.argmax(-1)[clear] == cat("mean_probs").argmax(-1)[clear]).all())
        report(test="full_size_rows_vs_oracle", workload=wl, dtype=dt, B=B, S=S, rows=rows, prob_err=perr, logit_err=lerr,
               ens_prob_err=eerr, logit_scale=scale, argmax_same=same)
        assert perr <= ptol and eerr <= ptol and lerr <= ltol * scale and same
        model.bnn_engine(dt).release_buffers()


def _c2_model(classes=10):
    torch.manual_seed(0)
    m = resnet18.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", dropout_p=0.5, out_dim=classes)
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1, generator=g)
            mod.running_var.uniform_(0.5, 1.5, generator=g)
            mod.weight.data.uniform_(0.5, 1.5, generator=g)
            mod.bias.data.normal_(0, 0.1, generator=g)
    return m.cuda().eval()


def test_config2_full_size_properties():
    """multi-exit ResNet-18 MCD, S=32, batch 256 (BASELINE configs[1])."""
    B, S = 256, 32
    model = _c2_model()
    x = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    eng = model.bnn_engine("fp16")
    r = mc_predict(model, x, S, seed=7)
    p = r.mean_probs.clone()
    lg = r.mean_logits.clone()
    ens = r.ens_probs.clone()
    ent, eent = r.entropy.clone(), r.expected_entropy.clone()
    assert p.shape == (4, B, 10) and torch.isfinite(p).all() and torch.isfinite(lg).all()
    assert (p.sum(-1) - 1).abs().max().item() < 1e-5                     # a mean of softmaxes is a distribution
    assert (p >= 0).all()
    # cumulative exit ensembles are running means over exits (results_analyzer.py:260-269)
    for e in range(4):
        assert (ens[e] - p[:e + 1].mean(0)).abs().max().item() < 1e-6
    # entropy of the mean >= mean of the entropies (Jensen): mutual information is non-negative
    assert (ent - eent).min().item() > -1e-4
    assert (ent <= np.log(10) + 1e-4).all()
    # determinism: same seed -> bit-identical statistics (eager, graph capture and graph replay)
    for _ in range(2):
        again = mc_predict(model, x, S, seed=7)
        assert torch.equal(again.mean_probs, p) and torch.equal(again.mean_logits, lg)
    # a different seed really changes the masks
    assert not torch.equal(mc_predict(model, x, S, seed=8).mean_probs, p)
    # shard invariance: samples [0,16) + [16,32) summed == samples [0,32)
    st = eng.enqueue(x.cuda(), 16, sample0=0, seed=7)
    part = st["sums"].clone()
    st = eng.enqueue(x.cuda(), 16, sample0=16, seed=7)
    st["sums"] += part
    views, _ = eng.finalize(st, B, S)
    assert (views[0] - p).abs().max().item() < 2e-6 and (views[1] - lg).abs().max().item() < 2e-5
    # tensor-core fp16 path vs exact fp32 CUDA-core path on the same masks, first 32 images, 8 samples
    a = mc_predict(model, x[:32], 8, seed=7, dtype="fp16").mean_probs.clone()
    b = mc_predict(model, x[:32], 8, seed=7, dtype="fp32").mean_probs
    assert (a - b).abs().max().item() <= 1e-3


def test_mask_stream_statistics_at_full_size():
    """the keep rate of the fused-epilogue / broadcast masks is 1 - p and E[dropout(x)] = x: with S = 128 samples the
    mean over samples of the first stochastic tensor approaches the un-masked tensor like 1/sqrt(S)."""
    from bayesnn_fpga_b200.Dropouts import MCDropout
    x = torch.rand(64, 64, 32, 32, device="cuda") + 0.5
    m = MCDropout(0.5).reseed(123, 0)
    acc = torch.zeros_like(x)
    kept = 0.0
    S = 64
    for _ in range(S):
        y = m(x)
        acc += y
        kept += (y != 0).float().mean().item()
    assert abs(kept / S - 0.5) < 1e-3
    rel = ((acc / S - x).abs() / x).mean().item()
    assert rel < 1.5 / np.sqrt(S)                                   # ~0.8/sqrt(S) expected for p = 0.5


def test_config4_vgg_full_width_smoke():
    """multi-exit VGG-19, dropout on the last 3 blocks, S=64, batch 512 (BASELINE configs[3]) - finite, normalised,
    prefix reuse == per-sample passes for the deterministic exits."""
    torch.manual_seed(0)
    m = vgg19.VGG19MCEarlyExit(dropout_exit=True, dropout=None, dropout_p=0.5, out_dim=100, image_size=32,
                               n_exits=5).append_block_dropout((2, 3, 4)).cuda().eval()
    x = torch.randn(512, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    r = mc_predict(m, x, 64, seed=3)
    p = r.mean_probs.clone()
    assert p.shape == (5, 512, 100) and torch.isfinite(p).all()
    assert (p.sum(-1) - 1).abs().max().item() < 1e-5
    r2 = mc_predict(m, x[:16], 64, seed=3)                          # batch slicing does not change an image's result
    # (Philox element indices are per image *within the batch*, so only image 0..15 with identical b match)
    assert (r2.mean_probs - p[:, :16]).abs().max().item() < 1e-5
