"""``FullAnalysis`` adapter - the consumer-facing half of the drop-in for
``Software_Artifact/software/train/results_analyzer.py`` (``FullAnalysis`` :55-64,
``update_model_exits`` :66-71, ``sdn_get_detailed_results`` :113-177, ``_get_output`` :236-270,
``ece_hist_binary`` :446-495, ``ece_eval_binary`` :497-505).

``_get_output`` returns the reference's 5-tuple but computes it with ONE fused ``mc_predict`` call
(prefix once, S samples in the GEMM M dimension, sums on the device) and a single device->host copy
instead of 2*E*S synchronising copies.  Dataset-level statistics are computed on the device:
``bnn_calibration_bins`` (histogram ECEs; the equal-mass binning of ``ece_hist_binary`` sorts the N confidences
with torch.sort - plumbing - and bins on the host exactly like the reference), ``bnn_dataset_metrics`` (NLL / MSE /
accuracy, :497-503), ``bnn_top_label`` + ``bnn_kde_triweight`` (the KDE-ECE of ``ece_kde_binary`` :351-443, with the
density estimate evaluated exactly instead of through KDEpy's FFT approximation, which is not installed) and
``bnn_confidence_exit`` (confidence-threshold early exiting :606-631; the FLOP accounting of ``flop_saver`` /
``flop_saver_ensembled`` :639-726 is the per-exit image histogram dotted with the reference's cost table).
``all_experiments`` (:288-337) computes the reference's per-exit lists without writing its log files.
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
        preds, ens, labels, top, ens_top = [], [], [], [], []
        for b_x, b_y in self.loader:
            output, _, sm_np, ens_output, ens_sm = self._get_output(b_x)
            preds.append(sm_np)
            ens.append(np.stack([t.numpy() for t in ens_sm]))
            labels.append(np.asarray(b_y).reshape(-1))
            # _update_layer_tracker (:273-276) decides correctness from `output[output_id].max(1)`: the argmax of the
            # MEAN LOGITS over the passes (for ensembles: of the mean over exits of those), not of the mean probabilities
            top.append(np.stack([t.max(1)[1].numpy() for t in output]))
            ens_top.append(np.stack([t.max(1)[1].numpy() for t in ens_output]))
        self.preds = np.concatenate(preds, axis=1)
        self.ensemble_preds = np.concatenate(ens, axis=1)
        lab = np.concatenate(labels)
        self.label_index = lab
        self.labels = np.eye(self.model.out_dim)[lab]
        self.layer_predictions = np.concatenate(top, axis=1)                   # [E, N] class index per exit
        self.ensemble_layer_predictions = np.concatenate(ens_top, axis=1)
        _, self.layer_correct, self.layer_wrong = layer_hits(self.layer_predictions, lab)
        _, self.ensemble_layer_correct, self.ensemble_layer_wrong = layer_hits(self.ensemble_layer_predictions, lab)
        return self.preds

    def get_validation_predictions(self, val_loader):
        """-> (preds [E, N, C], ensembled_preds [E, N, C], one-hot labels [N, C]) over `val_loader` like
        results_analyzer.py:179-215 (the trackers it also fills there are discarded by its only caller)."""
        preds, labels = [], []
        for b_x, b_y in val_loader:
            _, _, sm_np, _, _ = self._get_output(b_x)
            preds.append(sm_np)
            labels.append(np.asarray(b_y).reshape(-1))
        preds = np.concatenate(preds, axis=1)
        ens = np.stack([np.average(preds[:i], axis=0) for i in range(1, preds.shape[0] + 1)])       # :211-213
        return preds, ens, np.eye(self.model.out_dim)[np.concatenate(labels)]

    def save_validation(self, experiment_id, loader):
        """:217-222 - validation_predictions_<id>.npy holding preds, ensemble_preds, labels."""
        preds, ensemble_preds, labels = self.get_validation_predictions(loader)
        with open(f"validation_predictions_{experiment_id}.npy", "wb") as file:
            np.save(file, preds)
            np.save(file, ensemble_preds)
            np.save(file, labels)

    def average_results_accuracy(self):
        """:94-111 - mean over the exits of #correct and of the ECE, for single exits and cumulative ensembles."""
        n = len(self.layer_correct)
        acc = sum(len(self.layer_correct[l]) for l in range(n)) / n
        ens_acc = sum(len(self.ensemble_layer_correct[l]) for l in range(n)) / n
        ece = sum(self.ece_eval_binary(self.preds[l], self.labels)[0] for l in range(n)) / n
        ens_ece = sum(self.ece_eval_binary(self.ensemble_preds[l], self.labels)[0] for l in range(n)) / n
        return acc, ens_acc, ece, ens_ece

    def multipass_experiment(self, passes=range(1, 50)):
        """:73-92 - accuracy / ECE as a function of the number of MC passes; returns the four lists it prints."""
        out = ([], [], [], [])
        keep = self.mc_passes
        for s in passes:
            self.mc_passes = s
            self.sdn_get_detailed_results()
            for lst, v in zip(out, self.average_results_accuracy()):
                lst.append(v)
        self.mc_passes = keep if keep else 10
        for lst in out:
            print(",".join(str(v) for v in lst))
        return out

    # ---- calibration statistics ---------------------------------------------------------------
    def _f64(self, a):
        return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), dtype=torch.float64,
                               device=self.device)

    def _device_bins(self, p, label_index, n_bins, mode=_lib.CAL_TOP):
        """(conf float64 [N], hit int32 [N], bins float64 [n_bins, 3]) of `bnn_calibration_bins` - float64 end to end."""
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.require_device()
            dp = self._f64(p)
            dl = torch.as_tensor(np.ascontiguousarray(label_index), dtype=torch.int32, device=self.device)
            N, C = dp.shape
            conf = torch.empty(N, dtype=torch.float64, device=self.device)
            hit = torch.empty(N, dtype=torch.int32, device=self.device)
            bins = torch.zeros((n_bins, 3), dtype=torch.float64, device=self.device)
            if N == 0:
                return conf, hit, bins
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.bnn_calibration_bins(dp.data_ptr(), dl.data_ptr(), N, C, n_bins, mode, conf.data_ptr(),
                                                hit.data_ptr(), bins.data_ptr(), stream))
        return conf, hit, bins

    @staticmethod
    def _binned_ece(bins):
        b = bins.cpu().numpy()
        n = b[:, 0].sum()
        if n == 0:
            return 0.0
        nz = b[:, 0] > 0
        return float(np.sum(np.abs(b[nz, 1] / b[nz, 0] - b[nz, 2] / b[nz, 0]) * b[nz, 0] / n))

    def ece_width(self, p, label_index, n_bins=10):
        """Top-label ECE over equal-width bins - what hls4ml_pred.py:90-91 MEANS to compute - on the device."""
        return self._binned_ece(self._device_bins(p, label_index, n_bins)[2])

    def ece_tfp_as_called(self, y_prob, label_index, n_bins=10):
        """The number hls4ml_pred.py:90-91,115-116 REALLY prints: ``tfp.stats.expected_calibration_error(num_bins,
        logits=y_prob, labels_true, labels_predicted=argmax(y_prob))`` receives probabilities as `logits`, so tfp
        soft-maxes them a second time (float32) and bins that value into [lo, hi) bins.  Restated from
        tensorflow_probability/python/stats/calibration.py (tfp is not installed: unpinned); :meth:`ece_width` is the
        statistic the call was meant to compute."""
        return self._binned_ece(self._device_bins(np.asarray(y_prob, dtype=np.float32), label_index, n_bins,
                                                  _lib.CAL_TFP_RESOFTMAX)[2])

    def ece_hist_binary(self, p, label, n_bins=15, order=1):
        """Equal-mass ECE, same binning rule as results_analyzer.py:446-495: top-label confidence p[argmax] / sum(p)
        (multi-class) or the class-1 probability with the label itself as the outcome (C == 2, :463-465); confidences
        are float32 like the reference's `torch.zeros(N, 1)` buffer (:455) in the multi-class branch."""
        p = np.clip(np.asarray(p, dtype=np.float64), 1e-256, 1 - 1e-256)
        label_index = np.argmax(label, axis=1)
        binary = p.shape[1] == 2
        conf_d, hit_d = self._top_label(p, label_index, binary, round_f32=not binary)[:2]
        if not binary:
            conf_d = conf_d.float()
        srt = torch.sort(conf_d).values.cpu().numpy()
        conf, hit = conf_d.cpu().numpy(), hit_d.cpu().numpy().astype(np.float64)
        N = conf.shape[0]
        per = int(N / n_bins)
        edges = np.zeros(n_bins + 1, dtype=np.float32)                    # torch.zeros(len(bins)+1, 1) (:477)
        for i in range(n_bins):
            edges[i + 1] = srt[min((i + 1) * per, N - 1)]
        edges[0], edges[-1] = 0.0, 1.0
        total = 0.0
        for lo, hi in zip(edges[:-1], edges[1:]):
            inside = (conf > float(lo)) & (conf <= float(hi))
            if inside.any():
                avg = float(conf[inside].mean(dtype=conf.dtype))
                total += abs(avg - float(np.float32(hit[inside].mean()))) ** order * float(np.float32(inside.mean()))
        return total

    def dataset_metrics(self, p, label_index):
        """(nll, mse, accuracy) of results_analyzer.py:497-503 computed on the device in float64
        (`bnn_dataset_metrics`): the NLL clip is the reference's 1e-256, not a float32 stand-in."""
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.require_device()
            dp = self._f64(p)
            dl = torch.as_tensor(np.ascontiguousarray(label_index), dtype=torch.int32, device=self.device)
            N, C = dp.shape
            if N == 0:
                return 0.0, 0.0, 0.0
            ws = torch.empty(3 * ((N + 255) // 256), dtype=torch.float64, device=self.device)
            out = torch.empty(3, dtype=torch.float64, device=self.device)
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.bnn_dataset_metrics(dp.data_ptr(), dl.data_ptr(), N, C, ws.data_ptr(), out.data_ptr(), stream))
            nll, mse, acc = out.cpu().tolist()
        return nll, mse, acc

    def entropy(self, probs):
        """bayes_hw/metric_utils.py:3-6 on mean probabilities [N, C]."""
        probs = np.asarray(probs)
        return float(-np.sum(np.log(probs + 1e-8) * probs) / probs.shape[0])

    # ---- KDE-ECE ------------------------------------------------------------------------------------
    def _top_label(self, p, label_index, binary, round_f32):
        """`bnn_top_label` -> (conf float64 [N], flag int32 [N] | None, (n, sum, sum of squares) over the flagged)."""
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.require_device()
            dp = self._f64(p)
            N, C = dp.shape
            conf = torch.empty(N, dtype=torch.float64, device=self.device)
            st = torch.zeros(3, dtype=torch.float64, device=self.device)
            flag = dl = None
            if label_index is not None:
                dl = torch.as_tensor(np.ascontiguousarray(label_index), dtype=torch.int32, device=self.device)
                flag = torch.empty(N, dtype=torch.int32, device=self.device)
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.bnn_top_label(dp.data_ptr(), dl.data_ptr() if dl is not None else None, N, C, int(binary),
                                         int(round_f32), conf.data_ptr(), flag.data_ptr() if flag is not None else None,
                                         st.data_ptr(), stream))
            return conf, flag, st.cpu().tolist()

    def ece_kde_binary(self, p, label, p_int=None, order=1):
        """KDE-ECE like results_analyzer.py:351-443, both branches: top-label calibration (C != 2) and the joint
        calibration of the class-1 probability for binary problems (C == 2, :381-383, :417-419, :422-425), with an
        optional separate integration set `p_int` (:354-355, :409-419; default: `p` itself).  Confidences, bandwidth
        moments and both mirrored triweight density estimates are computed on the device in float64, the 2^14-point
        carry-forward integration (:426-443) on the host."""
        p = np.asarray(p, dtype=np.float64)
        p_int = p if p_int is None else np.asarray(p_int, dtype=np.float64)
        if p_int.shape[1] != p.shape[1]:
            raise ValueError("ece_kde_binary: p_int has %d classes, p has %d" % (p_int.shape[1], p.shape[1]))
        lib = _lib.load()
        binary = p.shape[1] == 2
        G = 2 ** 14
        x_int = np.linspace(-0.6, 1.6, num=G)
        N, N1 = p.shape[0], p_int.shape[0]
        label_index = np.argmax(label, axis=1)
        # the multi-class confidences of `p` live in a float32 tensor in the reference (:373); the binary branch and
        # the integration set stay float64 (:382, :414)
        conf, flag, (n1, s1, s2) = self._top_label(p, label_index, binary, round_f32=not binary)
        conf_int, _, _ = self._top_label(p_int, None, binary, round_f32=False)
        with torch.cuda.device(self.device):
            dens = torch.zeros((2, G), dtype=torch.float64, device=self.device)
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            var = max(s2 / n1 - (s1 / n1) ** 2, 0.0) if n1 > 0 else 0.0
            std = var ** 0.5
            kbw = (std if std > 1e-12 else 0.0000000000000001) * (N * 2) ** -0.2          # :389-393
            dx = float(x_int[1] - x_int[0])
            if n1 > 0:
                _lib.check(lib.bnn_kde_triweight(conf.data_ptr(), flag.data_ptr(), N, kbw, float(n1), float(x_int[0]), dx, G,
                                                 0.0, 1.0, dens[0].data_ptr(), stream))
            if N1 > 0:
                _lib.check(lib.bnn_kde_triweight(conf_int.data_ptr(), None, N1, kbw, float(N1), float(x_int[0]), dx, G,
                                                 0.0, 1.0, dens[1].data_ptr(), stream))
            pp = dens.cpu().numpy()
        perc = float(np.mean(label_index)) if binary else n1 / N                          # :422-425
        return _kde_ece_integrate(x_int, pp[0], pp[1], perc, order)

    def ece_eval_binary(self, p, label):
        """(ece, nll, mse, accuracy) like :497-505 - the ECE is the KDE-ECE, everything computed on the device."""
        nll, mse, accu = self.dataset_metrics(p, np.argmax(label, axis=1))
        return self.ece_kde_binary(p, label), nll, mse, accu

    # ---- per-exit experiments (:288-337) --------------------------------------------------------------
    def all_experiments(self, experiment_id=None):
        """The per-exit lists of results_analyzer.py:288-336 (correct / cumulative / destructive-overthinking /
        unique counts, ECE, NLL, MSE, accuracy for single exits and cumulative ensembles); the reference then
        writes them to log files (`saver`), which is left to the caller."""
        layers = sorted(self.layer_correct.keys())
        for prefix, preds, correct, wrong in (("", self.preds, self.layer_correct, self.layer_wrong),
                                              ("ensemble_", self.ensemble_preds, self.ensemble_layer_correct,
                                               self.ensemble_layer_wrong)):
            end_wrong = wrong[layers[-1]]
            cum = set()
            cols = {k: [] for k in ("cur_correct", "cum_correct", "overthinking", "unique", "ece", "nll", "mse", "accu")}
            for layer in layers:
                cur = correct[layer]
                cols["unique"].append(len(cur - cum))
                cum = cum | cur
                cols["cur_correct"].append(len(cur))
                cols["cum_correct"].append(len(cum))
                cols["overthinking"].append(len(cur & end_wrong))
                ece, nll, mse, accu = self.ece_eval_binary(preds[layer], self.labels)
                for k, v in (("ece", ece), ("nll", nll), ("mse", mse), ("accu", accu)):
                    cols[k].append(v)
            names = {"cur_correct": "cur_correct_saver", "cum_correct": "cum_correct_saver",
                     "overthinking": "destructive_overthinking", "unique": "unique_correct_saver" if not prefix else "unique_correct",
                     "ece": "ece_saver", "nll": "nll_saver", "mse": "mse_saver", "accu": "accu_saver"}
            for k, v in cols.items():
                setattr(self, prefix + names[k], v)
        return self

    # ---- confidence-threshold early exiting and FLOP accounting (:568-735) ----------------------------
    _FLOPS = {      # get_flops_per_module :568-580
        "vgg19": ([40173568, 56950784, 132448256, 132284416, 37789696], [14227456, 9467904, 4728832, 0, 0], [51200] * 5),
        "resnet18": ([154402816, 135036928, 134627328, 134422528], [56909824, 37871616, 18915328, 0], [51200] * 4),
    }

    def get_model_type(self):
        names = {c.__name__ for c in type(self.model).__mro__}
        if "VGG" in names:
            return "vgg19"
        if "ResNet" in names:
            return "resnet18"
        raise ValueError

    def get_dropout_type(self):
        """:582-596 -> (exit_only, dropout_rate, mc_passes)."""
        try:
            exit_only = True
            if self.model.dropout_exit and self.model.dropout is None:
                exit_only = True
            elif self.model.dropout is not None:
                exit_only = False
            return exit_only, self.model.dropout_p, 10
        except AttributeError:
            return True, 0, 1

    def get_flops_per_module(self):
        self.model_type = getattr(self, "model_type", None) or self.get_model_type()
        self.flops_per_layer, self.flop_per_exit_convs, self.flops_per_exit = [list(v) for v in self._FLOPS[self.model_type]]
        self.n_exits = len(self.flops_per_layer)
        self.baseline_flops = sum(self.flops_per_layer) + self.flop_per_exit_convs[-1] + self.flops_per_exit[-1]

    def is_confident(self, p_evals, threshold, layer, instance, diff=False):
        row = np.asarray(p_evals[layer][instance])
        if diff:
            top2 = np.sort(row)[-2:]
            return abs(top2[1] - top2[0]) > threshold
        return np.max(row).item() > threshold

    def _exit_scan(self, threshold, p_evals, diff):
        """bnn_confidence_exit over [E, N, C] -> (exit index per instance, chosen predictions, images per exit)."""
        lib = _lib.load()
        p_evals = np.asarray(p_evals)
        E, N, C = p_evals.shape
        with torch.cuda.device(self.device):
            _lib.require_device()
            dp = self._f64(p_evals)
            idx = torch.empty(N, dtype=torch.int32, device=self.device)
            best = torch.empty((N, C), dtype=torch.float64, device=self.device)
            hist = torch.zeros(E, dtype=torch.int32, device=self.device)
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.bnn_confidence_exit(dp.data_ptr(), E, N, C, min(1, E - 1), float(threshold), int(bool(diff)),
                                               idx.data_ptr(), best.data_ptr(), hist.data_ptr(), stream))
            return idx.cpu().numpy(), best.cpu().numpy(), hist.cpu().numpy()

    def confidence_exiting(self, threshold, p_evals, labels, diff=False):
        """-> (accuracy, ece, nll) of the predictions taken at the first confident exit (:606-630)."""
        _, best, _ = self._exit_scan(threshold, p_evals, diff)
        ece, nll, mse, accu = self.ece_eval_binary(best, labels)
        return accu, ece, nll

    def get_flops_standard_exit(self, layer, mc_passes, ensemble=False):
        if ensemble:
            return sum(self.flops_per_layer[:layer + 1]) + sum(self.flop_per_exit_convs[:layer + 1]) + \
                sum(self.flops_per_exit[:layer + 1]) * mc_passes
        return sum(self.flops_per_layer[:layer + 1]) + self.flop_per_exit_convs[layer] + self.flops_per_exit[layer] * mc_passes

    def _exit_cost(self, layer, mc_passes, ensembled):
        block = sum(self.flops_per_layer[:layer + 1])
        if not ensembled:
            if self.exit_only:
                return block + self.flop_per_exit_convs[layer] + mc_passes * self.flops_per_exit[layer]
            return mc_passes * (block + self.flop_per_exit_convs[layer] + self.flops_per_exit[layer])
        if self.exit_only:
            return block + sum(self.flop_per_exit_convs[:layer + 1]) + mc_passes * sum(self.flops_per_exit[:layer + 1])
        return mc_passes * (block + sum(self.flop_per_exit_convs[:layer + 1]) + sum(self.flops_per_exit[:layer + 1]))

    def flop_saver(self, threshold, p_evals, labels, mc_passes=10, diff=False):
        """Total FLOPs of confidence-threshold exiting over the dataset (:639-672)."""
        _, _, hist = self._exit_scan(threshold, p_evals, diff)
        return int(sum(int(n) * self._exit_cost(l, mc_passes, False) for l, n in enumerate(hist)))

    def flop_saver_ensembled(self, threshold, p_evals, labels, mc_passes=10, diff=False):
        """Same with cumulative-exit ensembles: all earlier exit branches are paid too (:674-726)."""
        _, _, hist = self._exit_scan(threshold, p_evals, diff)
        return int(sum(int(n) * self._exit_cost(l, mc_passes, True) for l, n in enumerate(hist)))

    def get_confidence_exiting_values(self, model_num=None, thresholds=(0.1, 0.15, 0.25, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95,
                                                                         0.99, 0.999)):
        """:543-566 on the predictions held by this object (the reference re-reads them from the .npy file `saver`
        wrote).  -> list of dicts with accuracy / ece / nll / flops relative to the baseline network."""
        self.model_type = self.get_model_type()
        self.exit_only, dropout_rate, mc_passes = self.get_dropout_type()
        self.get_flops_per_module()
        rows = []
        n = self.labels.shape[0]
        for thr in thresholds:
            for name, pe, fl in (("single", self.preds, self.flop_saver), ("ensemble", self.ensemble_preds,
                                                                            self.flop_saver_ensembled)):
                accu, ece, nll = self.confidence_exiting(thr, pe, self.labels)
                rows.append(dict(kind=name, threshold=thr, dropout_rate=dropout_rate, accuracy=accu, ece=ece, nll=nll,
                                 flops=fl(thr, pe, self.labels, mc_passes=mc_passes) / (self.baseline_flops * n)))
        return rows


def layer_hits(output, labels):
    """The layer trackers of results_analyzer.py:272-286 for one loader pass.  `output`: per-exit [N, C] mean LOGITS
    (tensors / arrays; the reference's `output[output_id].max(1)`), or an [E, N] array of already-taken argmaxes.
    -> (pred [E, N], {exit: set of correct instance ids}, {exit: set of wrong ids})."""
    if isinstance(output, np.ndarray) and output.ndim == 2:
        pred = output
    else:
        pred = np.stack([o.max(1)[1].cpu().numpy() if isinstance(o, torch.Tensor) else np.asarray(o).argmax(1)
                         for o in output])
    lab = np.asarray(labels).reshape(-1)
    good, bad = {}, {}
    for e in range(pred.shape[0]):
        hit = pred[e] == lab
        good[e] = set(np.flatnonzero(hit).tolist())
        bad[e] = set(np.flatnonzero(~hit).tolist())
    return pred, good, bad


def _kde_ece_integrate(x_int, pp1, pp2, perc, order=1):
    """results_analyzer.py:426-443, vectorised: integrand |conf - min(perc*pp1/pp2, 1)|^order * pp2 where a density
    exceeds 1e-6, else the previous integrand value carried forward (index > 1), then the ratio of trapezoid
    integrals over [0, 1]."""
    live = np.maximum(pp1, pp2) > 1e-6
    with np.errstate(divide="ignore", invalid="ignore"):
        accu = np.minimum(perc * pp1 / pp2, 1.0)
        val = np.abs(x_int - accu) ** order * pp2
    val = np.where(np.isnan(accu), 0.0, val)
    idx = np.where(live, np.arange(x_int.shape[0]), -1)
    idx[:2] = np.where(live[:2], idx[:2], np.arange(2))          # i <= 1 without density: stays 0 (own slot)
    val = np.where(live, val, 0.0)
    last = np.maximum.accumulate(idx)
    integral = val[last]
    ind = (x_int >= 0.0) & (x_int <= 1.0)
    trapz = getattr(np, "trapezoid", None) or np.trapz
    return float(trapz(integral[ind], x_int[ind]) / trapz(pp2[ind], x_int[ind]))
