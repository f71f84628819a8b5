"""Freeze golden vectors of the FullAnalysis statistics from the reference's OWN source (build container only).

    python tests/golden/make_golden_analysis.py     # needs /root/reference, writes tests/golden/analysis.npz

The methods are exec'd from Software_Artifact/software/train/results_analyzer.py inside a stub class (the module
itself imports KDEpy / matplotlib / sacred, which are not installed): mirror_1d :339-349, ece_kde_binary :351-443,
ece_eval_binary :497-505, get_flops_per_module :568-580, confidence_exiting :606-630, get_flops_standard_exit
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
    spans = [(338, 349), (350, 443), (496, 505), (567, 580), (605, 630), (631, 637), (638, 672), (673, 726), (727, 735)]
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
    np.savez_compressed(os.path.join(HERE, "analysis.npz"), **out)


if __name__ == "__main__":
    main()
