"""Restatement of the S-pass statistics (test infrastructure, see oracle/__init__.py).

``mc_get_output`` follows ``FullAnalysis._get_output``
(Software_Artifact/software/train/results_analyzer.py:236-270): S sequential forward passes,
softmax per exit per pass, float64 averages of logits and of probabilities over the passes,
then the cumulative-exit ensembles.  ``entropy`` follows
Hardware_Artifact/bayes_hw/metric_utils.py:3-6, ``ece_hist`` follows
results_analyzer.py:446-495, ``ece_width`` restates the 10-equal-width-bin top-label ECE the
hardware scripts mean to call (hls4ml_pred.py:90-91) and ``ece_tfp_as_called`` the number that call really produces,
``nll_mse_acc`` follows results_analyzer.py:497-503, ``layer_tracker`` :272-286.
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
    """results_analyzer.py:446-495: top-label confidence (C != 2; float32 like the reference's `torch.zeros(N,1)`
    buffer :455) or the class-1 probability with the label itself as the outcome (C == 2, :463-465; float64); bin
    edges at every (N // n_bins)-th sorted confidence, first edge 0, last edge 1."""
    p = np.clip(np.asarray(p, dtype=np.float64), 1e-256, 1 - 1e-256)
    N = p.shape[0]
    label_index = np.argmax(label_onehot, axis=1)
    if p.shape[1] != 2:
        pred = np.argmax(p, axis=1)
        conf = (p[np.arange(N), pred] / p.sum(axis=1)).astype(np.float32)
        hit = (pred == label_index).astype(np.float64)
    else:
        conf = (p / p.sum(axis=1)[:, None])[:, 1]
        hit = label_index.astype(np.int64)
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


def ece_tfp_as_called(y_prob, label_index, n_bins=10):
    """What hls4ml_pred.py:90-91,115-116 really evaluates:
    ``tfp.stats.expected_calibration_error(num_bins, logits=y_prob, labels_true, labels_predicted=argmax(y_prob))``
    with PROBABILITIES passed as `logits`.  tensorflow-probability (pinned by the reference: tensorflow_probability==0.16.0,
    Hardware_Artifact/requirements.txt:9) is not installed here, so this restates the published algorithm of
    tensorflow_probability/python/stats/calibration.py (`_compute_calibration_bin_statistics` +
    `expected_calibration_error`), float32 like TF:
        pred   = softmax(logits)                      <- the second soft-max of already-normalised probabilities
        prob_y = pred[i, labels_predicted[i]]
        bin    = tf.histogram_fixed_width_bins(prob_y, [0, 1], nbins) = clip(floor(nbins * prob_y), 0, nbins - 1)
        ece    = sum_b (n_b / N) * |correct_b / (n_b + tiny) - mean prob_y in b|   (empty bins contribute 0)
    parity unpinned (no tfp here); SURVEY.md A.3 lists the call as a reference defect - `ece_width` is what it meant."""
    y = np.asarray(y_prob, dtype=np.float32)
    pred_y = np.argmax(y, axis=1)
    z = np.exp(y - y.max(axis=1, keepdims=True))
    sm = (z / z.sum(axis=1, keepdims=True)).astype(np.float32)
    prob_y = sm[np.arange(y.shape[0]), pred_y]
    correct = (pred_y == np.asarray(label_index)).astype(np.float32)
    b = np.clip(np.floor(prob_y * np.float32(n_bins)).astype(np.int64), 0, n_bins - 1)
    tiny = np.finfo(np.float32).tiny
    out = 0.0
    n = float(y.shape[0])
    for k in range(n_bins):
        sel = b == k
        cnt = float(sel.sum())
        if cnt == 0:
            continue
        out += (cnt / n) * abs(float(correct[sel].sum()) / (cnt + tiny) - float(prob_y[sel].sum()) / cnt)
    return out


def layer_tracker(output, b_y):
    """_update_layer_tracker results_analyzer.py:272-286 without the per-instance dict bookkeeping: the prediction of
    an exit is ``output[output_id].max(1)[1]`` - the argmax of the MEAN LOGITS over the passes (for the ensemble
    trackers: of their cumulative mean over exits) - and an instance is correct when it equals the label.
    -> (pred [E, B] int64, correct [E, B] bool)."""
    pred = np.stack([np.asarray(o.max(1)[1]) if isinstance(o, torch.Tensor) else np.asarray(o).argmax(1)
                     for o in output])
    return pred, pred == np.asarray(b_y).reshape(1, -1)


def nll_mse_acc(p, label_onehot):
    """results_analyzer.py:497-503 (everything but the KDE-ECE, which needs KDEpy)."""
    p = np.asarray(p, dtype=np.float64)
    mse = np.mean(np.sum((p - label_onehot) ** 2, 1))
    pc = np.clip(p, 1e-256, 1 - 1e-256)
    nll = -np.sum(label_onehot * np.log(pc)) / p.shape[0]
    acc = np.mean(np.argmax(pc, 1) == np.argmax(label_onehot, 1))
    return nll, mse, acc


# ---- confidence-threshold early exiting and FLOP accounting (results_analyzer.py:568-580, :606-735) -----------
FLOP_TABLES = {     # get_flops_per_module :568-580
    "vgg19": dict(n_exits=5, flops_per_layer=[40173568, 56950784, 132448256, 132284416, 37789696],
                  flop_per_exit_convs=[14227456, 9467904, 4728832, 0, 0], flops_per_exit=[51200] * 5),
    "resnet18": dict(n_exits=4, flops_per_layer=[154402816, 135036928, 134627328, 134422528],
                     flop_per_exit_convs=[56909824, 37871616, 18915328, 0], flops_per_exit=[51200] * 4),
}


def baseline_flops(model_type):
    t = FLOP_TABLES[model_type]
    return sum(t["flops_per_layer"]) + t["flop_per_exit_convs"][-1] + t["flops_per_exit"][-1]      # :580


def is_confident(p_row, threshold, diff=False):
    """:728-735 - max p > threshold, or |top1 - top2| > threshold."""
    if diff:
        top2 = np.sort(p_row)[-2:]
        return abs(top2[1] - top2[0]) > threshold
    return np.max(p_row) > threshold


def exit_indices(threshold, p_evals, n_exits, diff=False, first_exit=1):
    """The exit every instance leaves at in confidence_exiting / flop_saver (:606-631): the scan starts at exit
    `first_exit` = 1 (`for layer in range(1, self.n_exits)`), the last exit takes whatever is left."""
    N = p_evals.shape[1]
    out = np.full(N, n_exits - 1, dtype=np.int64)
    for i in range(N):
        for layer in range(first_exit, n_exits - 1):
            if is_confident(p_evals[layer][i], threshold, diff):
                out[i] = layer
                break
    return out


def confidence_exiting_preds(threshold, p_evals, n_exits, diff=False):
    """best_preds of :606-629: every instance's prediction at the exit it leaves at."""
    idx = exit_indices(threshold, p_evals, n_exits, diff)
    return p_evals[idx, np.arange(p_evals.shape[1])], idx


def exit_flops(model_type, layer, mc_passes, exit_only, ensembled):
    """Cost charged to ONE instance leaving at `layer` by flop_saver (:639-672) / flop_saver_ensembled (:674-726)."""
    t = FLOP_TABLES[model_type]
    block = sum(t["flops_per_layer"][:layer + 1])
    if not ensembled:
        if exit_only:
            return block + t["flop_per_exit_convs"][layer] + mc_passes * t["flops_per_exit"][layer]
        return mc_passes * (block + t["flop_per_exit_convs"][layer] + t["flops_per_exit"][layer])
    f = block
    if exit_only:
        for prev in range(layer + 1):
            f += t["flop_per_exit_convs"][prev] + mc_passes * t["flops_per_exit"][prev]
        return f
    for prev in range(layer + 1):
        f += t["flop_per_exit_convs"][prev] + t["flops_per_exit"][prev]
    return f * mc_passes


def flop_saver(model_type, threshold, p_evals, mc_passes=10, diff=False, exit_only=True, ensembled=False):
    t = FLOP_TABLES[model_type]
    idx = exit_indices(threshold, p_evals, t["n_exits"], diff)
    return sum(exit_flops(model_type, int(l), mc_passes, exit_only, ensembled) for l in idx)


def flops_standard_exit(model_type, layer, mc_passes, ensemble=False):
    """get_flops_standard_exit :632-637."""
    t = FLOP_TABLES[model_type]
    if ensemble:
        return sum(t["flops_per_layer"][:layer + 1]) + sum(t["flop_per_exit_convs"][:layer + 1]) + \
            sum(t["flops_per_exit"][:layer + 1]) * mc_passes
    return sum(t["flops_per_layer"][:layer + 1]) + t["flop_per_exit_convs"][layer] + t["flops_per_exit"][layer] * mc_passes


# ---- KDE-ECE (results_analyzer.py:339-443) ----------------------------------------------------------------------
def mirror_1d(d, lo, hi):
    """:339-349 with both bounds: reflect the points below the midpoint about lo, the others about hi."""
    mid = (lo + hi) / 2
    return np.concatenate(((2 * lo - d[d < mid]).reshape(-1, 1), d, (2 * hi - d[d >= mid]).reshape(-1, 1)))


def kde_triweight_exact(data, bw, grid):
    """The estimator KDEpy's FFTKDE(bw, 'triweight').fit(data).evaluate(grid) approximates by linear binning + FFT:
    (1/n) sum_d K((x - d) / h) / h, K(u) = 35/32 (1 - u^2)^3 on |u| < 1, h = bw / sqrt(var K) = 3 bw
    (KDEpy scales every kernel so that `bw` is its standard deviation; triweight variance = 1/9)."""
    data = np.asarray(data, dtype=np.float64).reshape(-1)
    h = 3.0 * bw
    out = np.zeros(grid.shape[0])
    for lo in range(0, grid.shape[0], 1024):
        u = (grid[lo:lo + 1024, None] - data[None, :]) / h
        w = np.clip(1.0 - u * u, 0.0, None)
        out[lo:lo + 1024] = (w ** 3).sum(1)
    return out * (35.0 / 32.0) / (h * data.shape[0])


def ece_kde(p, label_onehot, order=1, kde=kde_triweight_exact, p_int=None):
    """ece_kde_binary :351-443, both branches (top-label for C != 2, joint calibration of the class-1 probability for
    C == 2) with an optional integration set p_int (default: p itself).  `kde(data, bw, grid)` stands in for KDEpy
    (not installed: parity of this statistic is pinned against the reference's own source with this exact estimator
    injected as `FFTKDE`, tests/golden/make_golden_analysis.py)."""
    p = np.asarray(p, dtype=np.float64)
    p_int = np.copy(p) if p_int is None else np.asarray(p_int, dtype=np.float64)
    p = np.clip(p, 1e-256, 1 - 1e-256)
    p_int = np.clip(p_int, 1e-256, 1 - 1e-256)
    x_int = np.linspace(-0.6, 1.6, num=2 ** 14)
    N = p.shape[0]
    label_index = np.argmax(label_onehot, axis=1)
    binary = p.shape[1] == 2
    if not binary:
        pred = np.argmax(p, axis=1)
        label_binary = (pred == label_index).astype(np.float64)
        # the reference keeps p_b in a float32 torch tensor (torch.zeros(N,1), :373)
        p_b = (p[np.arange(N), pred] / p.sum(1)).astype(np.float32)
    else:
        p_b = (p / p.sum(1)[:, None])[:, 1]                      # float64 (:382)
        label_binary = label_index
    dconf_1 = p_b[label_binary == 1].reshape(-1, 1)
    if np.std(dconf_1) != 0:
        kbw = np.std(dconf_1) * (N * 2) ** -0.2
    else:
        kbw = 0.0000000000000001 * (N * 2) ** -0.2
    pp1 = kde(mirror_1d(dconf_1, 0.0, 1.0), kbw, x_int)
    pp1[(x_int <= 0.0) | (x_int >= 1.0)] = 0
    pp1 = pp1 * 2
    p_int = p_int / p_int.sum(1)[:, None]
    N1 = p_int.shape[0]
    if not binary:
        pred_b_int = p_int[np.arange(N1), np.argmax(p_int, axis=1)].reshape(-1, 1)
    else:
        pred_b_int = p_int[:, 1].reshape(-1, 1)
    pp2 = kde(mirror_1d(pred_b_int, 0.0, 1.0), kbw, x_int)
    pp2[(x_int <= 0.0) | (x_int >= 1.0)] = 0
    pp2 = pp2 * 2
    perc = np.mean(label_binary) if not binary else np.mean(label_index)
    return kde_ece_integrate(x_int, pp1, pp2, perc, order)


def kde_ece_integrate(x_int, pp1, pp2, perc, order=1):
    """:426-443 - |conf - accuracy(conf)|^order weighted by the confidence density; where both densities vanish the
    previous integrand value is carried forward (the reference's sequential `integral[i] = integral[i-1]`)."""
    integral = np.zeros(x_int.shape)
    for i in range(x_int.shape[0]):
        conf = x_int[i]
        if max(pp1[i], pp2[i]) > 1e-6:
            with np.errstate(divide="ignore", invalid="ignore"):
                accu = np.min([perc * pp1[i] / pp2[i], 1.0])
            if not np.isnan(accu):
                integral[i] = np.abs(conf - accu) ** order * pp2[i]
        elif i > 1:
            integral[i] = integral[i - 1]
    ind = np.where((x_int >= 0.0) & (x_int <= 1.0))
    trapz = getattr(np, "trapezoid", None) or np.trapz
    return trapz(integral[ind], x_int[ind]) / trapz(pp2[ind], x_int[ind])
