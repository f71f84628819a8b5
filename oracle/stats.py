"""Restatement of the S-pass statistics (test infrastructure, see oracle/__init__.py).

``mc_get_output`` follows ``FullAnalysis._get_output``
(Software_Artifact/software/train/results_analyzer.py:236-270): S sequential forward passes,
softmax per exit per pass, float64 averages of logits and of probabilities over the passes,
then the cumulative-exit ensembles.  ``entropy`` follows
Hardware_Artifact/bayes_hw/metric_utils.py:3-6, ``ece_hist`` follows
results_analyzer.py:446-495, ``ece_width`` restates the 10-equal-width-bin top-label ECE the
hardware scripts call (hls4ml_pred.py:90-91), ``nll_mse_acc`` follows results_analyzer.py:497-503.
"""
import numpy as np
import torch
import torch.nn.functional as F


def mc_get_output(forward_pass, passes):
    """forward_pass(i) -> list of E logits tensors [B, C].  Returns a dict with the
    reference's five outputs (as float64 numpy arrays) plus the per-pass buffers."""
    all_logits, all_probs = None, None
    for i in range(passes):                                   # :240-246
        outs = forward_pass(i)
        sm = np.asarray([F.softmax(o, dim=1).cpu().numpy() for o in outs])
        lg = np.asarray([o.cpu().numpy() for o in outs])
        if all_logits is None:
            all_logits = np.empty((passes,) + lg.shape)       # float64 like np.empty (:238-239)
            all_probs = np.empty((passes,) + sm.shape)
        all_logits[i] = lg
        all_probs[i] = sm
    mean_logits = np.average(all_logits, axis=0)              # :247
    mean_probs = np.average(all_probs, axis=0)                # :248
    E = mean_probs.shape[0]
    ens_probs = np.stack([np.mean(mean_probs[:i + 1], axis=0) for i in range(E)])    # :260-269
    ens_logits = np.stack([np.mean(mean_logits[:i + 1], axis=0) for i in range(E)])
    return {"mean_logits": mean_logits, "mean_probs": mean_probs,
            "ens_logits": ens_logits, "ens_probs": ens_probs,
            "all_logits": all_logits, "all_probs": all_probs}


def entropy(probs):
    """metric_utils.py:3-6: -sum(log(p + 1e-8) * p) / batch over a [B, C] array."""
    probs = np.asarray(probs)
    return -np.sum(np.log(probs + 1e-8) * probs) / probs.shape[0]


def entropy_per_image(probs):
    probs = np.asarray(probs)
    return -np.sum(np.log(probs + 1e-8) * probs, axis=-1)


def ece_hist(p, label_onehot, n_bins=15, order=1):
    """results_analyzer.py:446-495 for the multi-class (C != 2) branch: top-label confidence,
    bin edges at every (N // n_bins)-th sorted confidence, first edge 0, last edge 1."""
    p = np.clip(np.asarray(p, dtype=np.float64), 1e-256, 1 - 1e-256)
    N = p.shape[0]
    label_index = np.argmax(label_onehot, axis=1)
    pred = np.argmax(p, axis=1)
    # the reference stores confidences in a float32 torch tensor (torch.zeros(N,1), :455)
    conf = (p[np.arange(N), pred] / p.sum(axis=1)).astype(np.float32)
    hit = (pred == label_index).astype(np.float64)
    srt = np.sort(conf)
    per = int(N / n_bins)
    edges = np.zeros(n_bins + 1, dtype=np.float32)
    for i in range(n_bins):
        edges[i + 1] = srt[min((i + 1) * per, N - 1)]
    edges[0], edges[-1] = 0.0, 1.0
    conf_t = torch.from_numpy(conf)
    hit_t = torch.from_numpy(hit)
    total = torch.zeros(1)
    for lo, hi in zip(edges[:-1], edges[1:]):
        inside = conf_t.gt(float(lo)) * conf_t.le(float(hi))
        frac = inside.float().mean()
        if frac.item() > 0:
            acc = hit_t[inside].float().mean()
            avg = conf_t[inside].mean()
            total += torch.abs(avg - acc) ** order * frac
    return float(total)


def ece_width(p, label_index, n_bins=10):
    """Top-label ECE with equal-width bins (lo, hi] over [0, 1] (the statistic
    tfp.stats.expected_calibration_error computes; called at hls4ml_pred.py:90-91)."""
    p = np.asarray(p, dtype=np.float64)
    pred = np.argmax(p, axis=1)
    conf = p[np.arange(p.shape[0]), pred]
    hit = (pred == np.asarray(label_index)).astype(np.float64)
    b = np.clip(np.ceil(conf * n_bins).astype(np.int64) - 1, 0, n_bins - 1)
    out = 0.0
    for k in range(n_bins):
        sel = b == k
        if sel.any():
            out += abs(conf[sel].mean() - hit[sel].mean()) * sel.mean()
    return out


def nll_mse_acc(p, label_onehot):
    """results_analyzer.py:497-503 (everything but the KDE-ECE, which needs KDEpy)."""
    p = np.asarray(p, dtype=np.float64)
    mse = np.mean(np.sum((p - label_onehot) ** 2, 1))
    pc = np.clip(p, 1e-256, 1 - 1e-256)
    nll = -np.sum(label_onehot * np.log(pc)) / p.shape[0]
    acc = np.mean(np.argmax(pc, 1) == np.argmax(label_onehot, 1))
    return nll, mse, acc
