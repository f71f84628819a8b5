"""Freeze golden vectors of the FullAnalysis statistics from the reference's OWN source (build container only).

    python tests/golden/make_golden_analysis.py     # needs /root/reference, writes tests/golden/analysis.npz

The methods are exec'd from Software_Artifact/software/train/results_analyzer.py inside a stub class (the module
itself imports KDEpy / matplotlib / sacred, which are not installed): _update_layer_tracker :272-286, mirror_1d
:339-349, ece_kde_binary :351-443, ece_hist_binary :446-495, ece_eval_binary :497-505, get_flops_per_module :568-580, confidence_exiting :606-630, get_flops_standard_exit
:632-637, flop_saver :639-672, flop_saver_ensembled :674-726, is_confident :728-735.

KDEpy is absent, so the name `FFTKDE` the reference source looks up is bound to a class that evaluates the SAME
estimator exactly (oracle.stats.kde_triweight_exact) instead of KDEpy's binned FFT approximation of it.  Everything
around the density estimate - clipping, top-label confidences, bandwidth rule, mirroring, the carry-forward
integration - is the reference's code, and the script asserts that the oracle restatement reproduces it.
"""
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/Software_Artifact/software/train/results_analyzer.py"

from oracle import philox, seeded, stats      # noqa: E402


class ExactKDE:
    def __init__(self, bw, kernel):
        assert kernel == "triweight"
        self.bw = bw

    def fit(self, data):
        self.data = np.asarray(data).reshape(-1)
        return self

    def evaluate(self, grid):
        return stats.kde_triweight_exact(self.data, self.bw, grid)


def load_stub():
    src = open(REF).read().split("\n")
    spans = [(271, 286), (338, 349), (350, 443), (445, 495), (496, 505), (567, 580), (605, 630), (631, 637), (638, 672), (673, 726), (727, 735)]
    body = "\n\n".join("\n".join(src[a:b]) for a, b in spans)
    code = "class Stub:\n" + textwrap.indent(textwrap.dedent(body), "    ")
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    ns = {"np": np, "torch": torch, "FFTKDE": ExactKDE}
    exec(code, ns)
    return ns["Stub"]


def synthetic_exits(E, N, C, seed):
    """[E, N, C] mean predictions whose confidence grows with the exit, and labels correlated with them."""
    ps = []
    for e in range(E):
        logits = (1.0 + 0.8 * e) * philox.normal(seed, e, 0, N * C).reshape(N, C)
        p = np.exp(logits - logits.max(1, keepdims=True))
        ps.append(p / p.sum(1, keepdims=True))
    p_evals = np.stack(ps)
    lab = seeded.seeded_labels(N, C, seed=seed)
    agree = philox.uniform01(seed + 1, 99, 0, N) < 0.7
    lab = np.where(agree, p_evals[-1].argmax(1), lab)
    return p_evals, lab


def main():
    Stub = load_stub()
    out = {}
    for k, (mt, N, C) in enumerate([("resnet18", 400, 10), ("vgg19", 300, 20)]):
        st = Stub()
        st.model_type = mt
        st.get_flops_per_module()
        E = st.n_exits
        assert stats.baseline_flops(mt) == st.baseline_flops
        p_evals, lab = synthetic_exits(E, N, C, seed=21 + k)
        onehot = np.eye(C)[lab]
        out["p%d" % k], out["lab%d" % k] = p_evals, lab
        for layer in range(E):
            for ens in (False, True):
                assert st.get_flops_standard_exit(layer, 10, ensemble=ens) == stats.flops_standard_exit(mt, layer, 10, ens)
        rows = []
        for thr, diff in [(0.5, False), (0.9, False), (0.25, True), (0.999, False)]:
            accu, ece, nll = st.confidence_exiting(thr, p_evals, onehot, diff=diff)
            best, idx = stats.confidence_exiting_preds(thr, p_evals, E, diff)
            nll_o, mse_o, acc_o = stats.nll_mse_acc(best, onehot)
            ece_o = stats.ece_kde(best, onehot)
            assert abs(accu - acc_o) < 1e-12 and abs(nll - nll_o) < 1e-12 and abs(ece - ece_o) < 1e-12, (thr, diff)
            fl = []
            for exit_only in (True, False):
                st.exit_only = exit_only
                a = st.flop_saver(thr, p_evals, onehot, mc_passes=10, diff=diff)
                b = st.flop_saver_ensembled(thr, p_evals, onehot, mc_passes=10, diff=diff)
                assert a == stats.flop_saver(mt, thr, p_evals, 10, diff, exit_only, False)
                assert b == stats.flop_saver(mt, thr, p_evals, 10, diff, exit_only, True)
                fl += [a, b]
            rows.append([thr, float(diff), accu, ece, nll] + [float(v) for v in fl])
            out["exit%d_%d" % (k, len(rows) - 1)] = idx
        out["rows%d" % k] = np.asarray(rows, dtype=np.float64)
        # per-exit KDE-ECE / NLL / MSE / accuracy (ece_eval_binary)
        per_exit = []
        for e in range(E):
            ece, nll, mse, accu = st.ece_eval_binary(p_evals[e], onehot)
            assert abs(ece - stats.ece_kde(p_evals[e], onehot)) < 1e-12
            per_exit.append([ece, nll, mse, accu])
        out["per_exit%d" % k] = np.asarray(per_exit)
        print("%s: confidence exiting, FLOP accounting and KDE-ECE - oracle == reference source" % mt)
    extra_cases(Stub, out)
    np.savez_compressed(os.path.join(HERE, "analysis.npz"), **out)


def extra_cases(Stub, out):
    """Branches the per-exit cases above do not reach: a separate integration set p_int, order 2, the binary (C == 2)
    joint-calibration branch of both ECEs, float64-only probabilities (NLL clip at 1e-256), and the layer trackers'
    argmax-of-mean-LOGITS rule on a batch where it disagrees with the argmax of the mean probabilities."""
    st = Stub()
    # ---- KDE-ECE with p_int / order / binary ---------------------------------------------------------------
    p_evals, lab = synthetic_exits(3, 500, 10, seed=40)
    onehot = np.eye(10)[lab]
    p, p_int = p_evals[1], p_evals[2][:320]
    vals = []
    for kw in (dict(p_int=p_int), dict(p_int=p_int, order=2), dict(order=2)):
        v = st.ece_kde_binary(p, onehot, **kw)
        assert abs(v - stats.ece_kde(p, onehot, **kw)) < 1e-12, kw
        vals.append(v)
    out["kde_p"], out["kde_p_int"], out["kde_lab"], out["kde_vals"] = p, p_int, lab, np.asarray(vals)
    pb_evals, labb = synthetic_exits(2, 600, 2, seed=41)
    onehot_b = np.eye(2)[labb]
    pb, pb_int = pb_evals[1], pb_evals[0][:450]
    vb = []
    for kw in (dict(), dict(p_int=pb_int), dict(order=2)):
        v = st.ece_kde_binary(pb, onehot_b, **kw)
        assert abs(v - stats.ece_kde(pb, onehot_b, **kw)) < 1e-12, ("binary", kw)
        vb.append(v)
    hb = [float(st.ece_hist_binary(pb, onehot_b)), float(st.ece_hist_binary(pb, onehot_b, n_bins=10, order=2)),
          float(st.ece_hist_binary(p, onehot)), float(st.ece_hist_binary(p, onehot, n_bins=7))]
    ho = [stats.ece_hist(pb, onehot_b), stats.ece_hist(pb, onehot_b, n_bins=10, order=2), stats.ece_hist(p, onehot),
          stats.ece_hist(p, onehot, n_bins=7)]
    assert np.abs(np.asarray(hb) - np.asarray(ho)).max() < 1e-7, (hb, ho)
    out["bin_p"], out["bin_p_int"], out["bin_lab"] = pb, pb_int, labb
    out["bin_kde_vals"], out["hist_vals"] = np.asarray(vb), np.asarray(hb)
    eb = st.ece_eval_binary(pb, onehot_b)
    assert abs(eb[1] - stats.nll_mse_acc(pb, onehot_b)[0]) < 1e-12
    out["bin_eval"] = np.asarray(eb, dtype=np.float64)
    # ---- probabilities only float64 can hold: the NLL clip is 1e-256, not FLT_MIN ----------------------------
    pu = p.copy()
    pu[np.arange(0, 500, 7), lab[::7]] = 1e-300           # the true class got (numerically) zero probability
    pu[np.arange(3, 500, 11), lab[3::11]] = 1e-60         # below float32's smallest normal, above the clip
    eu = st.ece_eval_binary(pu, onehot)
    nll_o, mse_o, acc_o = stats.nll_mse_acc(pu, onehot)
    assert abs(eu[1] - nll_o) < 1e-9 and abs(eu[2] - mse_o) < 1e-12 and eu[3] == acc_o
    assert abs(eu[0] - stats.ece_kde(pu, onehot)) < 1e-12
    out["under_p"], out["under_eval"] = pu, np.asarray(eu, dtype=np.float64)
    # ---- _update_layer_tracker: argmax of the mean logits ------------------------------------------------------
    E, B, C, S = 3, 64, 5, 4
    logits = 3.0 * philox.normal(77, 0, 0, S * E * B * C).reshape(S, E, B, C)
    logits[:, :, :, 0] += np.linspace(-2, 6, S)[:, None, None]     # a class with a wild logit in one pass
    probs = np.exp(logits - logits.max(-1, keepdims=True))
    probs /= probs.sum(-1, keepdims=True)
    mean_logits, mean_probs = logits.mean(0), probs.mean(0)
    assert (mean_logits.argmax(-1) != mean_probs.argmax(-1)).sum() >= 5        # the two rules really disagree here
    b_y = torch.from_numpy(seeded.seeded_labels(B, C, seed=5).astype(np.int64))
    st.device = torch.device("cpu")
    st.loader = type("L", (), {"batch_size": B})()
    trackers = [{e: set() for e in range(E)}, {e: set() for e in range(E)}, {e: {} for e in range(E)},
                {e: {} for e in range(E)}]
    output = [torch.from_numpy(mean_logits[e]) for e in range(E)]
    output_sm = [torch.from_numpy(mean_probs[e]) for e in range(E)]
    for e in range(E):
        trackers = list(st._update_layer_tracker(b_y, e, 0, *trackers, output, output_sm))
    pred_o, hit_o = stats.layer_tracker(output, b_y.numpy())
    for e in range(E):
        assert trackers[0][e] == set(np.flatnonzero(hit_o[e]).tolist())
        assert trackers[1][e] == set(np.flatnonzero(~hit_o[e]).tolist())
        assert all(int(trackers[2][e][i]) == pred_o[e, i] for i in range(B))
    out["trk_mean_logits"], out["trk_mean_probs"], out["trk_labels"] = mean_logits, mean_probs, b_y.numpy()
    out["trk_pred"] = pred_o
    print("extra cases (p_int, order 2, binary branches, float64-only probabilities, layer trackers) - "
          "oracle == reference source")


if __name__ == "__main__":
    main()
