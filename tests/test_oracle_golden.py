"""CPU: the oracle restatement against the fixtures frozen from the LIVE reference
(tests/golden/make_golden.py).  This is what pins the oracle (the reference ships no
golden vectors of its own)."""
import numpy as np
import pytest
import torch

from oracle import masksembles as o_masks
from oracle import nets, philox, seeded, stats
from tests.cases import CASES, GOLDEN, build_seeded, load_golden, oracle_run


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kats = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for ctr, key, want in kats:
        got = " ".join("%08x" % int(v) for v in philox.philox4x32_10(*ctr, *key))
        assert got == want


def test_philox_thresholds_and_edges():
    assert philox.threshold(0.0) == 0 and philox.threshold(0.5) == 2 ** 15 and philox.threshold(1.0) == 2 ** 16 - 1
    assert philox.keep_mask_flat(1, 2, 3, 1000, 0.0).all()          # p = 0 keeps everything
    assert not philox.keep_mask_flat(1, 2, 3, 1000, 1.0).any()      # p = 1 drops everything (F.dropout -> zeros)
    assert philox.keep_mask_flat(1, 2, 3, 0, 0.5).shape == (0,)     # empty
    m = philox.keep_mask_flat(1, 2, 3, 200003, 0.25)                # ragged length, statistics
    assert abs(m.mean() - 0.75) < 5e-3
    # NHWC contract: element e of the flat stream is (b, h, w, c)
    k = philox.keep_mask(9, 4, 7, (2, 8, 3, 5), 0.5)
    flat = philox.keep_mask_flat(9, 4, 7, 2 * 8 * 3 * 5, 0.5).reshape(2, 3, 5, 8)
    assert (k == flat.transpose(0, 3, 1, 2)).all()
    # channel-wise masks are constant over space
    c = philox.keep_mask(9, 4, 7, (2, 8, 3, 5), 0.5, mode="channel")
    assert (c == c[:, :, :1, :1]).all()


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_reproduces_live_reference(tag):
    model_sd_gold = build_seeded(tag)
    _, sd, gold = model_sd_gold
    x = torch.from_numpy(gold["x"])
    assert np.array_equal(x.numpy(), seeded.seeded_input(x.shape, seed=99).numpy())
    got = oracle_run(tag, sd, x, int(gold["S"]), int(gold["seed"]), float(gold["p"]))
    for k in ("mean_logits", "mean_probs", "ens_logits", "ens_probs"):
        assert np.abs(got[k] - gold[k]).max() <= 1e-6, k      # reference outputs, same torch ops
    assert np.abs(got["all_logits"] - gold["all_logits"]).max() <= 1e-6


@pytest.mark.parametrize("tag", sorted(CASES))
def test_state_dict_names_match_reference(tag):
    model, _, gold = build_seeded(tag)
    assert sorted(model.state_dict().keys()) == list(gold["state_keys"])


def test_masksembles_generator_bit_exact():
    from bayesnn_fpga_b200 import utils as product
    z = np.load(GOLDEN + "/masksembles_masks.npz")
    assert len(z.files) == 8
    for name in z.files:
        c, n, s, seed = name.split("_")
        c, n, s, seed = int(c[1:]), int(n[1:]), float(s[1:]), int(seed[4:])
        for gen in (o_masks.generation_wrapper, product.generation_wrapper):
            np.random.seed(seed)
            got = gen(c, n, s)
            assert got.shape == z[name].shape and (got == z[name]).all(), (name, gen.__module__)
        assert (z[name].sum(1) == z[name].sum(1)[0]).all()          # every mask keeps the same count


def test_masksembles_generator_errors():
    from bayesnn_fpga_b200 import utils as product
    for gen in (o_masks.generation_wrapper, product.generation_wrapper):
        with pytest.raises(ValueError):
            gen(9, 4, 2.0)          # c < 10 (utils.py:77-81)
        with pytest.raises(ValueError):
            gen(64, 4, 6.5)         # scale > 6 (utils.py:83-85)


def test_ece_hist_matches_reference_source():
    z = np.load(GOLDEN + "/ece_hist.npz")
    for k in range(3):
        p, lab = z["p%d" % k], z["label%d" % k]
        onehot = np.eye(p.shape[1])[lab]
        assert abs(stats.ece_hist(p, onehot) - float(z["ece%d" % k])) < 1e-7


def test_entropy_and_ece_width_definitions():
    p = np.array([[0.5, 0.5], [1.0, 0.0]])
    want = -(2 * 0.5 * np.log(0.5 + 1e-8) + 1.0 * np.log(1.0 + 1e-8) + 0.0) / 2
    assert abs(stats.entropy(p) - want) < 1e-12
    # perfectly calibrated two-bin toy: conf 0.75 with 3/4 correct
    p = np.array([[0.75, 0.25]] * 4)
    assert abs(stats.ece_width(p, np.array([0, 0, 0, 1]))) < 1e-12
    assert abs(stats.ece_width(p, np.array([0, 0, 1, 1])) - 0.25) < 1e-12


def test_analysis_statistics_oracle_matches_reference_fixture():
    """confidence exiting, FLOP accounting and KDE-ECE restatements vs values frozen from the reference's own source
    (tests/golden/make_golden_analysis.py; KDEpy's FFTKDE replaced by the exact estimator it approximates)."""
    z = np.load(GOLDEN + "/analysis.npz")
    for k, mt in enumerate(["resnet18", "vgg19"]):
        p_evals, lab = z["p%d" % k], z["lab%d" % k]
        E, N, C = p_evals.shape
        onehot = np.eye(C)[lab]
        rows = z["rows%d" % k]
        for r, row in enumerate(rows):
            thr, diff = float(row[0]), bool(row[1])
            best, idx = stats.confidence_exiting_preds(thr, p_evals, E, diff)
            assert (idx == z["exit%d_%d" % (k, r)]).all()
            nll, mse, acc = stats.nll_mse_acc(best, onehot)
            assert abs(acc - row[2]) < 1e-12 and abs(nll - row[4]) < 1e-12
            want_fl = row[5:9]
            got_fl = [stats.flop_saver(mt, thr, p_evals, 10, diff, eo, ens) for eo in (True, False) for ens in (False, True)]
            assert [float(v) for v in got_fl] == list(want_fl)
        assert idx.min() >= 1                                   # the reference never exits at exit 0 (:612)
        if k == 0:                                              # KDE-ECE: one exit is enough for the CPU suite
            assert abs(stats.ece_kde(p_evals[1], onehot) - z["per_exit%d" % k][1, 0]) < 1e-12
    # branches frozen by extra_cases(): separate integration set, order 2, binary joint calibration, float64-only inputs
    onehot = np.eye(10)[z["kde_lab"]]
    assert abs(stats.ece_kde(z["kde_p"], onehot, p_int=z["kde_p_int"], order=2) - z["kde_vals"][1]) < 1e-12
    onehot_b = np.eye(2)[z["bin_lab"]]
    assert abs(stats.ece_kde(z["bin_p"], onehot_b, p_int=z["bin_p_int"]) - z["bin_kde_vals"][1]) < 1e-12
    assert abs(stats.ece_hist(z["bin_p"], onehot_b) - z["hist_vals"][0]) < 1e-7
    assert abs(stats.nll_mse_acc(z["under_p"], onehot)[0] - z["under_eval"][1]) < 1e-9
    # the exact estimator integrates to one and reproduces a closed form at a single point
    grid = np.linspace(-1, 1, 4001)
    dens = stats.kde_triweight_exact(np.array([0.0]), 0.1, grid)
    assert abs(np.sum(dens) * (grid[1] - grid[0]) - 1.0) < 1e-6 and abs(dens[2000] - 35.0 / 32.0 / 0.3) < 1e-12


def test_layer_trackers_use_the_argmax_of_the_mean_logits():
    """_update_layer_tracker (results_analyzer.py:272-286) takes `output[output_id].max(1)`: the argmax of the mean
    LOGITS.  Fixture: a batch where that differs from the argmax of the mean probabilities, predictions frozen from
    the reference's own source; the oracle restatement and the product's host function must both reproduce them."""
    from bayesnn_fpga_b200.results_analyzer import layer_hits
    z = np.load(GOLDEN + "/analysis.npz")
    ml, mp, lab, want = z["trk_mean_logits"], z["trk_mean_probs"], z["trk_labels"], z["trk_pred"]
    assert (ml.argmax(-1) != mp.argmax(-1)).sum() >= 5
    pred_o, hit_o = stats.layer_tracker([torch.from_numpy(m) for m in ml], lab)
    assert (pred_o == want).all()
    pred, good, bad = layer_hits([torch.from_numpy(m) for m in ml], lab)
    assert (pred == want).all()
    for e in range(ml.shape[0]):
        assert good[e] == set(np.flatnonzero(want[e] == lab).tolist()) and bad[e] == set(np.flatnonzero(want[e] != lab).tolist())


def test_tfp_ece_as_called_restatement():
    """ece_tfp_as_called: probabilities re-soft-maxed, [lo, hi) bins.  For C classes the re-soft-maxed top probability
    lies in [1/C, e/(e + C - 1)], so with 10 classes only bins 1-2 can be hit and the value is ~|accuracy - 0.2|."""
    rng = np.random.RandomState(0)
    logits = 3 * rng.randn(500, 10)
    p = np.exp(logits - logits.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    lab = np.where(rng.rand(500) < 0.8, p.argmax(1), rng.randint(0, 10, 500))
    v = stats.ece_tfp_as_called(p, lab)
    acc = (p.argmax(1) == lab).mean()
    assert abs(v - abs(acc - 0.2)) < 0.06 and abs(v - stats.ece_width(p, lab)) > 0.3
    # two-bin closed form: all predictions [1, 0] -> softmax top = e / (e + 1) = 0.731 -> bin 7, all correct
    q = np.tile(np.array([[1.0, 0.0]]), (8, 1))
    assert abs(stats.ece_tfp_as_called(q, np.zeros(8, dtype=np.int64)) - (1 - np.e / (np.e + 1))) < 1e-6


def test_mask_contract_rows_of_a_larger_batch():
    """Element indices are per image WITHIN the batch: the masks of rows [o, o + b) of a B-image batch are a slice of
    the full mask - what lets the full-size GPU tests check rows of a B = 256 / 512 run against the oracle."""
    for shape, mode in (((6, 12, 3, 5), "element"), ((9, 7), "element"), ((9, 7, 4, 4), "channel")):
        full = philox.keep_mask(7, 2, 5, shape, 0.3, mode)
        for off, n in ((0, 2), (3, 2), (5, 1)):
            part = philox.keep_mask(7, 2, 5, (n,) + shape[1:], 0.3, mode, batch_offset=off)
            assert (full[off:off + n] == part).all()
    x = seeded.seeded_input((5, 1, 28, 28), seed=4)
    sd = seeded.seeded_state_dict(nets.LENET_SHAPES, seed=3)
    spec = nets.SiteSpec("mc", 0.2)
    with torch.no_grad():
        whole = nets.lenet_forward(sd, x, nets.InjectedSites(spec, 9, 1))
        tail = nets.lenet_forward(sd, x[3:], nets.InjectedSites(spec, 9, 1, batch_offset=3))
    for a, b in zip(whole, tail):
        assert torch.allclose(a[3:], b, atol=1e-6)


def test_masksembles_batched_formulation_oracle_matches_reference_fixture():
    """oracle.nets.GroupSites (batch split into n groups, group g x mask g) reproduces the outputs frozen from the
    reference's own Masksembles training branch inside its ResNet18MCEarlyExit, and the Keras batched inference
    formulation built on it (tests/golden/make_golden_masksembles_batched.py)."""
    z = np.load(GOLDEN + "/masksembles_batched.npz")
    model, sd, gold = build_seeded("resnet18_mask_block")
    tables = {k[:-len(".masks")]: v for k, v in sd.items() if k.endswith(".masks")}
    spec = nets.SiteSpec("mask", 0.0, tables)
    with torch.no_grad():
        got = nets.resnet18_forward(sd, torch.from_numpy(z["groups_x"]), nets.GroupSites(spec), "block", True)
        for a, w in zip(got, z["groups_logits"]):
            assert np.abs(a.numpy() - w).max() < 2e-6
        with pytest.raises(ValueError) as e:
            nets.resnet18_forward(sd, torch.from_numpy(z["groups_x"][:6]), nets.GroupSites(spec), "block", True)
        assert str(e.value) == str(z["groups_error"][0])
        xb = torch.from_numpy(z["batched_x"])
        preds = [torch.softmax(o, 1) for o in nets.resnet18_forward(sd, torch.cat([xb] * 4), nets.GroupSites(spec), "block", True)]
    per_exit = np.stack([p.reshape(4, -1, p.shape[-1]).mean(0).numpy() for p in preds])
    assert np.abs(per_exit - z["batched_exit_probs"]).max() < 2e-6
    assert np.abs(per_exit.mean(0) - z["batched_avg"]).max() < 2e-6
