"""``FullAnalysis`` adapter - the consumer-facing half of the drop-in for
``Software_Artifact/software/train/results_analyzer.py`` (``FullAnalysis`` :55-64,
``update_model_exits`` :66-71, ``sdn_get_detailed_results`` :113-177, ``_get_output`` :236-270,
``ece_hist_binary`` :446-495, ``ece_eval_binary`` :497-505).

``_get_output`` returns the reference's 5-tuple but computes it with ONE fused ``mc_predict`` call
(prefix once, S samples in the GEMM M dimension, sums on the device) and a single device->host copy
instead of 2*E*S synchronising copies.  Dataset-level calibration statistics are computed on the
device (``bnn_calibration_bins``); the equal-mass binning of ``ece_hist_binary`` sorts the N confidences
with torch.sort (plumbing) and bins on the host exactly like the reference.  The KDE-ECE
(:351-443) needs KDEpy and stays out of scope.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .predict import mc_predict


class FullAnalysis:
    def __init__(self, model, test_loader, gpu=0, mc_dropout=False, mc_passes=10, suffix="", seed=0x5EED,
                 dtype=None, run=True):
        self.model = model
        self.update_model_exits()
        self.loader = test_loader
        self.gpu = gpu
        self.mc_dropout = mc_dropout
        self.mc_passes = mc_passes
        self.seed = seed
        self.dtype = dtype
        self.filename_suffix = suffix
        self.outputs = list(range(self.model.n_exits))
        self.device = torch.device("cuda:%d" % gpu)
        self._batches = 0
        if run and test_loader is not None:
            self.sdn_get_detailed_results()

    def update_model_exits(self):
        """results_analyzer.py:66-71: single-exit metadata on a multi-exit class -> real exit count."""
        name = type(self.model).__name__
        if self.model.n_exits == 1 and name in ("VGG19EarlyExit", "VGG19MCEarlyExit"):
            self.model.n_exits = 5
        elif self.model.n_exits == 1 and name in ("ResNet18EarlyExit", "ResNet18MCEarlyExit"):
            self.model.n_exits = 4

    # ---- the hot path -----------------------------------------------------------------------
    def _predict(self, b_x):
        S = self.mc_passes if self.mc_dropout else 1
        # every batch draws fresh masks, like the reference whose RNG keeps advancing
        seed = (self.seed + 0x9E3779B97F4A7C15 * self._batches) & 0xFFFFFFFFFFFFFFFF
        self._batches += 1
        return mc_predict(self.model, b_x, S, seed=seed, dtype=self.dtype)

    def _get_output(self, b_x):
        """-> (output, output_sm, output_sm_np, ensemble_output, ensemble_output_sm) like :236-270:
        lists of E float64 tensors [B, C] (mean logits / mean probs / cumulative-exit means) and the
        [E, B, C] float64 array of mean probs."""
        r = self._predict(b_x)
        host = torch.stack([r.mean_logits, r.mean_probs, r.ens_logits, r.ens_probs]).double().cpu()
        E = host.shape[1]
        output = [host[0, e] for e in range(E)]
        output_sm = [host[1, e] for e in range(E)]
        output_sm_np = host[1].numpy()
        ensemble_output = [host[2, e] for e in range(E)]
        ensemble_output_sm = [host[3, e] for e in range(E)]
        return output, output_sm, output_sm_np, ensemble_output, ensemble_output_sm

    def sdn_get_detailed_results(self):
        """Batch loop of :113-177 without the per-sample Python bookkeeping: fills ``preds`` /
        ``ensemble_preds`` [E, N, C], ``labels`` (one-hot [N, C]) and per-exit correct/wrong id sets."""
        preds, ens, labels = [], [], []
        for b_x, b_y in self.loader:
            _, _, sm_np, _, ens_sm = self._get_output(b_x)
            preds.append(sm_np)
            ens.append(np.stack([t.numpy() for t in ens_sm]))
            labels.append(np.asarray(b_y).reshape(-1))
        self.preds = np.concatenate(preds, axis=1)
        self.ensemble_preds = np.concatenate(ens, axis=1)
        lab = np.concatenate(labels)
        self.label_index = lab
        self.labels = np.eye(self.model.out_dim)[lab]
        self.layer_correct, self.layer_wrong = {}, {}
        self.ensemble_layer_correct, self.ensemble_layer_wrong = {}, {}
        for e in self.outputs:
            for src, good, bad in ((self.preds, self.layer_correct, self.layer_wrong),
                                   (self.ensemble_preds, self.ensemble_layer_correct, self.ensemble_layer_wrong)):
                hit = src[e].argmax(1) == lab
                good[e] = set(np.flatnonzero(hit).tolist())
                bad[e] = set(np.flatnonzero(~hit).tolist())
        return self.preds

    # ---- calibration statistics ---------------------------------------------------------------
    def _device_bins(self, p, label_index, n_bins):
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.require_device()
            dp = torch.as_tensor(np.ascontiguousarray(p), dtype=torch.float32, device=self.device)
            dl = torch.as_tensor(np.ascontiguousarray(label_index), dtype=torch.int32, device=self.device)
            N, C = dp.shape
            conf = torch.empty(N, dtype=torch.float32, device=self.device)
            hit = torch.empty(N, dtype=torch.int32, device=self.device)
            bins = torch.zeros((n_bins, 3), dtype=torch.float32, device=self.device)
            if N == 0:
                return conf, hit, bins
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.bnn_calibration_bins(dp.data_ptr(), dl.data_ptr(), N, C, n_bins, conf.data_ptr(),
                                                hit.data_ptr(), bins.data_ptr(), stream))
        return conf, hit, bins

    def ece_width(self, p, label_index, n_bins=10):
        """Top-label ECE over equal-width bins (the statistic of hls4ml_pred.py:90-91), on the device."""
        _, _, bins = self._device_bins(p, label_index, n_bins)
        b = bins.double().cpu().numpy()
        n = b[:, 0].sum()
        if n == 0:
            return 0.0
        nz = b[:, 0] > 0
        return float(np.sum(np.abs(b[nz, 1] / b[nz, 0] - b[nz, 2] / b[nz, 0]) * b[nz, 0] / n))

    def ece_hist_binary(self, p, label, n_bins=15, order=1):
        """Equal-mass top-label ECE, same binning rule as results_analyzer.py:446-495 (multi-class branch)."""
        label_index = np.argmax(label, axis=1)
        conf_d, hit_d, _ = self._device_bins(np.clip(p, 1e-256, 1 - 1e-256), label_index, n_bins)
        srt = torch.sort(conf_d).values.cpu().numpy()
        conf, hit = conf_d.cpu().numpy(), hit_d.cpu().numpy().astype(np.float64)
        N = conf.shape[0]
        per = int(N / n_bins)
        edges = np.zeros(n_bins + 1, dtype=np.float32)
        for i in range(n_bins):
            edges[i + 1] = srt[min((i + 1) * per, N - 1)]
        edges[0], edges[-1] = 0.0, 1.0
        total = 0.0
        for lo, hi in zip(edges[:-1], edges[1:]):
            inside = (conf > lo) & (conf <= hi)
            if inside.any():
                total += abs(float(conf[inside].mean(dtype=np.float32)) - float(hit[inside].mean())) ** order \
                    * float(inside.mean())
        return total

    def dataset_metrics(self, p, label_index):
        """(nll, mse, accuracy) of results_analyzer.py:497-503 computed on the device (`bnn_dataset_metrics`)."""
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.require_device()
            dp = torch.as_tensor(np.ascontiguousarray(p), dtype=torch.float32, device=self.device)
            dl = torch.as_tensor(np.ascontiguousarray(label_index), dtype=torch.int32, device=self.device)
            N, C = dp.shape
            if N == 0:
                return 0.0, 0.0, 0.0
            ws = torch.empty(3 * ((N + 255) // 256), dtype=torch.float32, device=self.device)
            out = torch.empty(3, dtype=torch.float32, device=self.device)
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.bnn_dataset_metrics(dp.data_ptr(), dl.data_ptr(), N, C, ws.data_ptr(), out.data_ptr(), stream))
            nll, mse, acc = out.cpu().tolist()
        return nll, mse, acc

    def ece_eval_binary(self, p, label):
        """(ece, nll, mse, accuracy) like :497-505 with the histogram ECE in place of the KDE one (KDEpy is not
        available); everything is computed on the device."""
        nll, mse, accu = self.dataset_metrics(p, np.argmax(label, axis=1))
        return self.ece_hist_binary(p, label), nll, mse, accu

    def entropy(self, probs):
        """bayes_hw/metric_utils.py:3-6 on mean probabilities [N, C]."""
        probs = np.asarray(probs)
        return float(-np.sum(np.log(probs + 1e-8) * probs) / probs.shape[0])
