// Dataset-level statistics of the reference's analyzers on the device, in FLOAT64 like the reference's NumPy code:
//   ece_hist_binary / ece_eval_binary / ece_kde_binary  (Software_Artifact/software/train/results_analyzer.py:351-505),
//   confidence_exiting / is_confident / flop_saver*      (:606-735),
//   the tfp ECE call of the hardware scripts              (Hardware_Artifact/bayes_hw/hls4ml_pred.py:90-91,115-116).
// The inputs are the [N, C] (or [E, N, C]) mean-probability arrays FullAnalysis keeps as float64 (:133-135); N <= 1e5,
// so every kernel here is latency-bound and double precision costs nothing measurable - but float32 copies would clip
// the NLL at -log(FLT_MIN) = 87.3 instead of the reference's -log(1e-256) = 589.5 and score a float32-underflowed
// probability differently.
#include "common.cuh"

namespace bnn {

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// first maximum wins, like np.argmax / torch.argmax on the reference's host arrays
__device__ __forceinline__ int row_argmax(const double* __restrict__ p, int C, double* best, double* sum) {
  int arg = 0;
  double bp = p[0], s = 0.0;
  for (int c = 0; c < C; ++c) {
    const double v = p[c];
    s += v;
    if (v > bp) {
      bp = v;
      arg = c;
    }
  }
  *best = bp;
  *sum = s;
  return arg;
}

__global__ void calibration_kernel(const double* __restrict__ probs, const int32_t* __restrict__ labels, int N, int C,
                                   int n_bins, int mode, double* __restrict__ conf, int32_t* __restrict__ correct,
                                   double* __restrict__ bin_stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double* p = probs + (size_t)i * C;
  double bp, sum;
  const int best = row_argmax(p, C, &bp, &sum);
  const int hit = best == labels[i];
  double cf = bp;
  int bin;
  if (mode == BNN_CAL_TOP_NORM) {
    cf = bp / sum;
  } else if (mode == BNN_CAL_TFP_RESOFTMAX) {
    // hls4ml_pred.py:90-91 passes PROBABILITIES as `logits=`: tfp soft-maxes them again (float32, like TF) and bins
    // the re-soft-maxed probability of the predicted label
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf((float)p[c] - (float)bp);
    cf = (double)(1.f / den);
  }
  if (mode == BNN_CAL_TFP_RESOFTMAX) {
    bin = (int)floorf((float)cf * (float)n_bins);          // tf.histogram_fixed_width_bins: [lo, hi) bins, clipped
  } else {
    bin = (int)ceil(cf * (double)n_bins) - 1;              // (lo, hi] bins
  }
  bin = max(0, min(n_bins - 1, bin));
  conf[i] = cf;
  correct[i] = hit;
  atomicAdd(&bin_stats[bin * 3 + 0], 1.0);
  atomicAdd(&bin_stats[bin * 3 + 1], cf);
  atomicAdd(&bin_stats[bin * 3 + 2], (double)hit);
}

// ---- confidence-threshold early exiting (results_analyzer.py:606-631, :728-735) ----------------------------------
// Image i leaves at the first exit e in [first_exit, E - 1) whose mean prediction is confident - max p > threshold, or
// (diff) |top1 - top2| > threshold - and at exit E - 1 otherwise.  One thread per image.
__global__ void confidence_exit_kernel(const double* __restrict__ probs, int E, int N, int C, int first_exit,
                                       double threshold, int diff, int32_t* __restrict__ exit_idx,
                                       double* __restrict__ best, int32_t* __restrict__ hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int chosen = E - 1;
  for (int e = first_exit; e < E - 1; ++e) {
    const double* p = probs + ((size_t)e * N + i) * C;
    double t1 = -INFINITY, t2 = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const double v = p[c];
      if (v > t1) {
        t2 = t1;
        t1 = v;
      } else if (v > t2) {
        t2 = v;
      }
    }
    const bool confident = diff ? (fabs(t1 - t2) > threshold) : (t1 > threshold);
    if (confident) {
      chosen = e;
      break;
    }
  }
  exit_idx[i] = chosen;
  atomicAdd(&hist[chosen], 1);
  const double* p = probs + ((size_t)chosen * N + i) * C;
  for (int c = 0; c < C; ++c) best[(size_t)i * C + c] = p[c];
}

// ---- KDE-ECE building blocks (results_analyzer.py:351-443) ---------------------------------------------------------
// Multi-class branch (:370-380): conf = p[argmax] / sum(p), flag = (argmax == label).
// Binary branch (C == 2, :381-383): conf = p[1] / sum(p), flag = label (the class index itself).
// Probabilities are clipped to [1e-256, 1 - 1e-256] first (:357-358).  round_f32: the reference stores the multi-class
// confidences of `p` in a float32 tensor (torch.zeros(N,1), :373); those of p_int stay float64 (:414).
// stats (only with labels): n, sum conf, sum conf^2 over the flagged images - the bandwidth rule (:389-393).
__global__ void top_label_kernel(const double* __restrict__ probs, const int32_t* __restrict__ labels, int N, int C,
                                 int binary, int round_f32, double* __restrict__ conf, int32_t* __restrict__ flag,
                                 double* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double n1 = 0.0, s1 = 0.0, s2 = 0.0;
  if (i < N) {
    const double* p = probs + (size_t)i * C;
    int best = 0;
    double bp = -1.0, sum = 0.0;
    for (int c = 0; c < C; ++c) {
      const double v = fmin(fmax(p[c], 1e-256), 1.0 - 1e-256);
      sum += v;
      if (v > bp) {
        bp = v;
        best = c;
      }
    }
    double cf = binary ? fmin(fmax(p[1], 1e-256), 1.0 - 1e-256) / sum : bp / sum;
    if (round_f32) cf = (double)(float)cf;
    conf[i] = cf;
    if (labels != nullptr) {
      const int f = binary ? (labels[i] == 1) : (best == labels[i]);
      flag[i] = f;
      if (f) {
        n1 = 1.0;
        s1 = cf;
        s2 = cf * cf;
      }
    }
  }
  n1 = warp_sum(n1);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0 && n1 > 0.0) {
    atomicAdd(&stats[0], n1);
    atomicAdd(&stats[1], s1);
    atomicAdd(&stats[2], s2);
  }
}

// Exact triweight kernel density estimate of the data mirrored about lo and hi (mirror_1d :339-349: points below the
// midpoint are reflected about lo, the others about hi), evaluated on the grid x0 + j*dx, set to zero outside
// (lo, hi) and doubled (:403-406) - i.e. (1/n) sum_d [K_h(x - d) + K_h(x - mirror(d))] inside the domain.
// K_h(u) = (35/32)(1 - (u/h)^2)^3 / h for |u| < h with h = 3*bw (KDEpy: bw is the kernel's standard deviation,
// triweight variance 1/9).  The reference evaluates the same estimate with KDEpy's FFT approximation.
// One block per 128 grid points; the data stream through shared memory.
__global__ void __launch_bounds__(128) kde_triweight_kernel(const double* __restrict__ data,
                                                            const int32_t* __restrict__ flags, int n, double h,
                                                            double inv_n, double x0, double dx, int G, double lo,
                                                            double hi, double* __restrict__ out) {
  __shared__ double sd[512];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const double x = x0 + (double)j * dx;
  const bool inside = j < G && x > lo && x < hi;
  const double inv_h = 1.0 / h, mid = 0.5 * (lo + hi);
  double acc = 0.0;
  for (int base = 0; base < n; base += 512) {
    __syncthreads();
    for (int t = threadIdx.x; t < 512; t += blockDim.x) {
      const int i = base + t;
      // excluded points are parked far outside every kernel support
      sd[t] = (i < n && (flags == nullptr || flags[i] != 0)) ? data[i] : 1e300;
    }
    __syncthreads();
    if (inside) {
      const int m = min(512, n - base);
      for (int t = 0; t < m; ++t) {
        const double d = sd[t];
        if (d > 1e299) continue;
        const double dm = d < mid ? 2.0 * lo - d : 2.0 * hi - d;
        double u = (x - d) * inv_h;
        if (fabs(u) < 1.0) {
          const double w = 1.0 - u * u;
          acc += w * w * w;
        }
        u = (x - dm) * inv_h;
        if (fabs(u) < 1.0) {
          const double w = 1.0 - u * u;
          acc += w * w * w;
        }
      }
    }
  }
  if (j < G) out[j] = inside ? acc * (35.0 / 32.0) * inv_h * inv_n : 0.0;
}

// per-image NLL / Brier (MSE) / top-1 hit (ece_eval_binary :497-503), block-reduced in a fixed order; partial[block][3]
__global__ void __launch_bounds__(256) dataset_metrics_kernel(const double* __restrict__ probs,
                                                              const int32_t* __restrict__ labels, int N, int C,
                                                              double* __restrict__ partial) {
  __shared__ double sh[3][256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double nll = 0.0, mse = 0.0, hit = 0.0;
  if (i < N) {
    const double* p = probs + (size_t)i * C;
    const int y = labels[i];
    int best = 0;
    double bp = -1.0;
    for (int c = 0; c < C; ++c) {
      const double v = p[c];
      const double t = c == y ? 1.0 : 0.0;
      mse += (v - t) * (v - t);                                     // :498 on the UNclipped probabilities
      const double vc = fmin(fmax(v, 1e-256), 1.0 - 1e-256);        // :500
      if (vc > bp) {
        bp = vc;
        best = c;
      }
    }
    nll = -log(fmin(fmax(p[y], 1e-256), 1.0 - 1e-256));             // :501
    hit = best == y ? 1.0 : 0.0;                                    // :502
  }
  sh[0][threadIdx.x] = nll;
  sh[1][threadIdx.x] = mse;
  sh[2][threadIdx.x] = hit;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 3) partial[blockIdx.x * 3 + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void dataset_metrics_reduce_kernel(const double* __restrict__ partial, int blocks, double inv_n,
                                              double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= 3) return;
  double a = 0.0;
  for (int b = 0; b < blocks; ++b) a += partial[b * 3 + k];
  out[k] = a * inv_n;
}

}  // namespace bnn

using namespace bnn;

extern "C" {

int bnn_calibration_bins(const double* probs, const int32_t* labels, int N, int C, int n_bins, int mode, double* conf,
                         int32_t* correct, double* bin_stats, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(probs && labels && conf && correct && bin_stats && N >= 0 && C > 0 && n_bins > 0,
              "bnn_calibration_bins: bad arguments");
  BNN_REQUIRE(mode == BNN_CAL_TOP || mode == BNN_CAL_TOP_NORM || mode == BNN_CAL_TFP_RESOFTMAX,
              "bnn_calibration_bins: unknown mode %d", mode);
  cudaStream_t st = (cudaStream_t)stream;
  BNN_CUDA_OK(cudaMemsetAsync(bin_stats, 0, sizeof(double) * 3 * n_bins, st));
  if (N == 0) return BNN_OK;
  calibration_kernel<<<(N + 127) / 128, 128, 0, st>>>(probs, labels, N, C, n_bins, mode, conf, correct, bin_stats);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_dataset_metrics(const double* probs, const int32_t* labels, int N, int C, double* workspace, double* out,
                        void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(probs && labels && workspace && out && N > 0 && C > 0, "bnn_dataset_metrics: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (N + 255) / 256;
  dataset_metrics_kernel<<<blocks, 256, 0, st>>>(probs, labels, N, C, workspace);
  BNN_LAUNCH_OK();
  dataset_metrics_reduce_kernel<<<1, 32, 0, st>>>(workspace, blocks, 1.0 / (double)N, out);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_confidence_exit(const double* probs, int E, int N, int C, int first_exit, double threshold, int diff,
                        int32_t* exit_idx, double* best_probs, int32_t* exit_hist, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(probs && exit_idx && best_probs && exit_hist, "bnn_confidence_exit: null pointer");
  BNN_REQUIRE(E > 0 && N >= 0 && C > 0 && first_exit >= 0 && first_exit < E, "bnn_confidence_exit: bad geometry");
  cudaStream_t st = (cudaStream_t)stream;
  BNN_CUDA_OK(cudaMemsetAsync(exit_hist, 0, sizeof(int32_t) * E, st));
  if (N == 0) return BNN_OK;
  confidence_exit_kernel<<<(N + 127) / 128, 128, 0, st>>>(probs, E, N, C, first_exit, threshold, diff, exit_idx,
                                                          best_probs, exit_hist);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_top_label(const double* probs, const int32_t* labels, int N, int C, int binary, int round_f32, double* conf,
                  int32_t* flag, double* stats, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(probs && conf && stats && N >= 0 && C > 0, "bnn_top_label: bad arguments");
  BNN_REQUIRE(labels == nullptr || flag != nullptr, "bnn_top_label: labels without a flag buffer");
  BNN_REQUIRE(!binary || C == 2, "bnn_top_label: the binary branch needs C == 2, got %d", C);
  cudaStream_t st = (cudaStream_t)stream;
  BNN_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 3, st));
  if (N == 0) return BNN_OK;
  top_label_kernel<<<(N + 127) / 128, 128, 0, st>>>(probs, labels, N, C, binary, round_f32, conf, flag, stats);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_kde_triweight(const double* data, const int32_t* flags, int n, double bw, double n_points, double x0,
                      double dx, int G, double lo, double hi, double* out, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(data && out && n >= 0 && G > 0 && bw > 0.0 && n_points > 0.0 && hi > lo,
              "bnn_kde_triweight: bad arguments");
  kde_triweight_kernel<<<(G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(data, flags, n, 3.0 * bw, 1.0 / n_points, x0,
                                                                          dx, G, lo, hi, out);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

}  // extern "C"
