// Exit-head and statistics kernels (sm_100a, CUDA cores; HBM/L2-bound, no tensor-core shape here:
// the FC is F x C with C = 10..100 per image).
//
//   exit_head_kernel : global average pool -> stochastic site -> Linear(F, C) -> softmax ->
//                      warp-shuffle accumulation over the local MC samples.  One CTA per image;
//                      all samples of that image are processed by the same CTA so the classifier
//                      weights are read once per image and the running sums never leave the SM.
//   finalize_kernel  : means, cumulative exit ensembles, entropies.
//   calibration_kernel: top-label confidence / correctness + equal-width bins.
#include "common.cuh"
#include "philox.cuh"

namespace bnn {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_WARPS = HEAD_THREADS / 32;
constexpr int HEAD_SCHUNK = 16;  // samples staged in shared memory at a time

template <typename T>
struct __align__(16) Vec8h {
  T v[8];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// logits[ns x C] = pooled^T * W^T + bias.  Thread = (class slot cl, K-split kq); it owns CPT classes
// (cl + j * nthr_c) x 16 samples in registers.  Per f: CPT coalesced weight loads (W^T is [F][C], served by L2 - the
// CTAs' shared memory leaves almost no L1) and four broadcast LDS.128 of the 16 pooled samples feed 16 * CPT FMAs;
// CPT = 4 for wide heads keeps the shared-memory pipe (the measured limiter at C = 100) off the critical path.
template <int CPT>
__device__ __forceinline__ void head_gemm(const float* __restrict__ pooled, float* __restrict__ part,
                                          float* __restrict__ logits, const float* __restrict__ wt,
                                          const float* __restrict__ bias, int F, int C, int ns, int tid) {
  constexpr int GROUP_MAX = HEAD_THREADS * CPT;          // classes handled per pass
  for (int cg0 = 0; cg0 < C; cg0 += GROUP_MAX) {
    const int cgroup = min(C - cg0, GROUP_MAX);
    int nthr_c = 1;
    while (nthr_c * CPT < cgroup) nthr_c <<= 1;          // class slots (power of two <= 256)
    const int ks = HEAD_THREADS / nthr_c;
    const int cl = tid % nthr_c, kq = tid / nthr_c;
    const int fk = (F + ks - 1) / ks;
    const int f_lo = kq * fk, f_hi = min(F, f_lo + fk);
    float accs[CPT][HEAD_SCHUNK];
#pragma unroll
    for (int j = 0; j < CPT; ++j)
#pragma unroll
      for (int i = 0; i < HEAD_SCHUNK; ++i) accs[j][i] = 0.f;
    bool cvalid[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) cvalid[j] = cg0 + cl + j * nthr_c < C;
    const float* wbase = wt + cg0 + cl;
    constexpr int U = CPT == 1 ? 8 : 4;                  // independent weight loads in flight: U * CPT
    int f = f_lo;
    for (; f + U <= f_hi; f += U) {
      float wv[U][CPT];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < CPT; ++j) wv[u][j] = cvalid[j] ? __ldg(wbase + (size_t)(f + u) * C + j * nthr_c) : 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4* pp = reinterpret_cast<const float4*>(pooled + (size_t)(f + u) * HEAD_SCHUNK);
#pragma unroll
        for (int i = 0; i < HEAD_SCHUNK / 4; ++i) {
          const float4 pv = pp[i];
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            accs[j][4 * i] = fmaf(wv[u][j], pv.x, accs[j][4 * i]);
            accs[j][4 * i + 1] = fmaf(wv[u][j], pv.y, accs[j][4 * i + 1]);
            accs[j][4 * i + 2] = fmaf(wv[u][j], pv.z, accs[j][4 * i + 2]);
            accs[j][4 * i + 3] = fmaf(wv[u][j], pv.w, accs[j][4 * i + 3]);
          }
        }
      }
    }
    for (; f < f_hi; ++f) {
      const float4* pp = reinterpret_cast<const float4*>(pooled + (size_t)f * HEAD_SCHUNK);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const float wv = cvalid[j] ? __ldg(wbase + (size_t)f * C + j * nthr_c) : 0.f;
#pragma unroll
        for (int i = 0; i < HEAD_SCHUNK / 4; ++i) {
          const float4 pv = pp[i];
          accs[j][4 * i] = fmaf(wv, pv.x, accs[j][4 * i]);
          accs[j][4 * i + 1] = fmaf(wv, pv.y, accs[j][4 * i + 1]);
          accs[j][4 * i + 2] = fmaf(wv, pv.z, accs[j][4 * i + 2]);
          accs[j][4 * i + 3] = fmaf(wv, pv.w, accs[j][4 * i + 3]);
        }
      }
    }
    // partial sums -> part[kq][sample][class slot j * nthr_c + cl]; then a fixed-order reduction over kq
    const int cw = nthr_c * CPT;
#pragma unroll
    for (int j = 0; j < CPT; ++j)
#pragma unroll
      for (int i = 0; i < HEAD_SCHUNK; ++i) part[(kq * HEAD_SCHUNK + i) * cw + j * nthr_c + cl] = accs[j][i];
    __syncthreads();
    for (int idx = tid; idx < ns * cw; idx += HEAD_THREADS) {
      const int sl = idx / cw, cc = idx - sl * cw;
      if (cg0 + cc < C) {
        float a = __ldg(bias + cg0 + cc);
        for (int k = 0; k < ks; ++k) a += part[(k * HEAD_SCHUNK + sl) * cw + cc];
        logits[sl * C + cg0 + cc] = a;
      }
    }
    __syncthreads();
  }
}

// dynamic shared memory layout (floats):
//   pooled[F][HEAD_SCHUNK] | part[HEAD_THREADS * HEAD_SCHUNK * (C > 32 ? 4 : 1)] | logits[HEAD_SCHUNK][C] | acc_p[C] | acc_l[C] |
//   red[HEAD_WARPS] | smax[HEAD_SCHUNK] | sinv[HEAD_SCHUNK]
template <typename T>
__global__ void __launch_bounds__(HEAD_THREADS) exit_head_kernel(
    const T* __restrict__ feat, int feat_has_samples, int B, int S_local, int HW, int F, int C,
    const float* __restrict__ wt, const float* __restrict__ bias, DropParams dp, float* __restrict__ sum_p,
    float* __restrict__ sum_logit, float* __restrict__ sum_plogp, float* __restrict__ logits_out, int accumulate) {
  extern __shared__ float sm[];
  float* pooled = sm;
  float* part = pooled + (size_t)HEAD_SCHUNK * F;
  float* logits = part + HEAD_THREADS * HEAD_SCHUNK * (C > 32 ? 4 : 1);
  float* acc_p = logits + (size_t)HEAD_SCHUNK * C;
  float* acc_l = acc_p + C;
  float* red = acc_l + C;
  float* smax = red + HEAD_WARPS;
  float* sinv = smax + HEAD_SCHUNK;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float inv_hw = 1.f / (float)HW;

  for (int c = tid; c < C; c += HEAD_THREADS) {
    acc_p[c] = 0.f;
    acc_l[c] = 0.f;
  }
  float plogp_acc = 0.f;  // meaningful on lane 0 of each warp

  for (int s0 = 0; s0 < S_local; s0 += HEAD_SCHUNK) {
    const int ns = min(HEAD_SCHUNK, S_local - s0);
    __syncthreads();
    // ---- phase 1: pooled + masked feature vectors for ns samples -----------------------------
    if (F % 8 == 0) {
      // vector path: a thread owns 8 consecutive features (16-byte loads for 16-bit storage, all HW loads of
      // one item in flight together) and one Philox block (= 8 elements) per (sample, feature octet)
      const int octs = F / 8;
      for (int idx = tid; idx < ns * octs; idx += HEAD_THREADS) {
        const int sl = idx / octs, f0 = (idx - sl * octs) * 8;
        const int s = s0 + sl;
        const T* src = feat + (((size_t)(feat_has_samples ? s : 0) * B + b) * HW) * F + f0;
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int p = 0; p < HW; ++p) {
          const Vec8h<T> v = *reinterpret_cast<const Vec8h<T>*>(src + (size_t)p * F);
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] += to_f32<T>(v.v[j]);
        }
        uint32_t k8 = 0xffu;
        float fac = inv_hw;
        if (dp.kind == BNN_DROP_ELEMENT || dp.kind == BNN_DROP_CHANNEL) {
          k8 = dp.scale == 0.f ? 0u
                               : philox_keep8(dp.seed, dp.stream_id, dp.sample0 + s, ((uint64_t)b * F + f0) >> 3, dp.thr);
          fac *= dp.scale;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = ((k8 >> j) & 1u) ? a[j] * fac : 0.f;
          if (dp.kind == BNN_DROP_MASKSEMBLES) v *= drop_factor(dp, (uint32_t)s, 0, 0, f0 + j);
          pooled[(f0 + j) * HEAD_SCHUNK + sl] = v;
        }
      }
    } else {
      for (int idx = tid; idx < ns * F; idx += HEAD_THREADS) {
        const int sl = idx / F, f = idx - sl * F;
        const int s = s0 + sl;
        const T* src = feat + (((size_t)(feat_has_samples ? s : 0) * B + b) * HW) * F + f;
        float a = 0.f;
        for (int p = 0; p < HW; ++p) a += to_f32<T>(src[(size_t)p * F]);
        a *= inv_hw;
        if (dp.kind != BNN_DROP_NONE)
          a *= drop_factor(dp, (uint32_t)s, (uint64_t)b * F + f, (uint64_t)b * F + f, f);
        pooled[f * HEAD_SCHUNK + sl] = a;
      }
    }
    __syncthreads();
    // ---- phase 2: logits[ns x C] = pooled[ns x F] * W^T[F x C] + bias as a register-tiled small GEMM ----
    if (C > 32)
      head_gemm<4>(pooled, part, logits, wt, bias, F, C, ns, tid);
    else
      head_gemm<1>(pooled, part, logits, wt, bias, F, C, ns, tid);
    // ---- phase 3a: softmax statistics per sample (one warp per sample) -----------------------
    for (int sl = warp; sl < ns; sl += HEAD_WARPS) {
      const float* lg = logits + sl * C;
      float mx = -INFINITY;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
      mx = warp_max(mx);
      float den = 0.f;
      for (int c = lane; c < C; c += 32) den += expf(lg[c] - mx);
      den = warp_sum(den);
      const float inv = 1.f / den, logden = logf(den);
      float pl = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float z = lg[c] - mx;
        pl += expf(z) * inv * (z - logden);
      }
      pl = warp_sum(pl);
      plogp_acc += pl;
      if (lane == 0) {
        smax[sl] = mx;
        sinv[sl] = inv;
      }
    }
    __syncthreads();
    // ---- phase 3b: one thread per class walks the samples in order (deterministic sums) ---------
    for (int c = tid; c < C; c += HEAD_THREADS) {
      float ap = 0.f, al = 0.f;
      for (int sl = 0; sl < ns; ++sl) {
        const float lg = logits[sl * C + c];
        ap += expf(lg - smax[sl]) * sinv[sl];
        al += lg;
        if (logits_out) logits_out[((size_t)(s0 + sl) * B + b) * C + c] = lg;
      }
      acc_p[c] += ap;
      acc_l[c] += al;
    }
  }
  __syncthreads();
  if (lane == 0) red[warp] = plogp_acc;
  __syncthreads();
  for (int c = tid; c < C; c += HEAD_THREADS) {
    const size_t o = (size_t)b * C + c;
    sum_p[o] = (accumulate ? sum_p[o] : 0.f) + acc_p[c];
    sum_logit[o] = (accumulate ? sum_logit[o] : 0.f) + acc_l[c];
  }
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < HEAD_WARPS; ++i) t += red[i];
    sum_plogp[b] = (accumulate ? sum_plogp[b] : 0.f) + t;
  }
}

// one thread per (image, class): means and cumulative exit ensembles
__global__ void finalize_means_kernel(const float* __restrict__ sum_p, const float* __restrict__ sum_logit, int E,
                                      int B, int C, float inv_s, float* __restrict__ mean_p,
                                      float* __restrict__ mean_logit, float* __restrict__ ens_p,
                                      float* __restrict__ ens_logit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over B*C
  if (i >= B * C) return;
  float run_p = 0.f, run_l = 0.f;
  for (int e = 0; e < E; ++e) {
    const size_t o = (size_t)e * B * C + i;
    const float mp = sum_p[o] * inv_s, ml = sum_logit[o] * inv_s;
    mean_p[o] = mp;
    mean_logit[o] = ml;
    run_p += mp;
    run_l += ml;
    ens_p[o] = run_p / (float)(e + 1);
    ens_logit[o] = run_l / (float)(e + 1);
  }
}

// one thread per (exit, image): entropy of the mean, -sum_c p log(p + 1e-8) (metric_utils.py:5)
__global__ void finalize_entropy_kernel(const float* __restrict__ mean_p, const float* __restrict__ ens_p,
                                        const float* __restrict__ sum_plogp, int EB, int C, float inv_s,
                                        float* __restrict__ entropy, float* __restrict__ ens_entropy,
                                        float* __restrict__ exp_entropy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over E*B
  if (i >= EB) return;
  const float* p = mean_p + (size_t)i * C;
  const float* q = ens_p + (size_t)i * C;
  float h = 0.f, he = 0.f;
  for (int c = 0; c < C; ++c) {
    h -= p[c] * logf(p[c] + 1e-8f);
    he -= q[c] * logf(q[c] + 1e-8f);
  }
  entropy[i] = h;
  ens_entropy[i] = he;
  exp_entropy[i] = -sum_plogp[i] * inv_s;
}

__global__ void calibration_kernel(const float* __restrict__ probs, const int32_t* __restrict__ labels, int N, int C,
                                   int n_bins, float* __restrict__ conf, int32_t* __restrict__ correct,
                                   float* __restrict__ bin_stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* p = probs + (size_t)i * C;
  int best = 0;
  float bp = p[0];
  for (int c = 1; c < C; ++c)
    if (p[c] > bp) {   // first maximum wins, like np.argmax
      bp = p[c];
      best = c;
    }
  const int hit = best == labels[i];
  conf[i] = bp;
  correct[i] = hit;
  int bin = (int)ceilf(bp * (float)n_bins) - 1;
  bin = max(0, min(n_bins - 1, bin));
  atomicAdd(&bin_stats[bin * 3 + 0], 1.f);
  atomicAdd(&bin_stats[bin * 3 + 1], bp);
  atomicAdd(&bin_stats[bin * 3 + 2], (float)hit);
}


// ---- confidence-threshold early exiting (results_analyzer.py:606-631, :728-735) ----------------------------------
// Image i leaves at the first exit e in [first_exit, E) whose mean prediction is confident - max p > threshold, or
// (diff) top1 - top2 > threshold - and at exit E-1 otherwise.  One thread per image.
__global__ void confidence_exit_kernel(const float* __restrict__ probs, int E, int N, int C, int first_exit,
                                       float threshold, int diff, int32_t* __restrict__ exit_idx,
                                       float* __restrict__ best, int32_t* __restrict__ hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int chosen = E - 1;
  for (int e = first_exit; e < E - 1; ++e) {
    const float* p = probs + ((size_t)e * N + i) * C;
    float t1 = -INFINITY, t2 = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const float v = p[c];
      if (v > t1) {
        t2 = t1;
        t1 = v;
      } else if (v > t2) {
        t2 = v;
      }
    }
    const bool confident = diff ? (fabsf(t1 - t2) > threshold) : (t1 > threshold);
    if (confident) {
      chosen = e;
      break;
    }
  }
  exit_idx[i] = chosen;
  atomicAdd(&hist[chosen], 1);
  const float* p = probs + ((size_t)chosen * N + i) * C;
  for (int c = 0; c < C; ++c) best[(size_t)i * C + c] = p[c];
}

// ---- KDE-ECE building blocks (results_analyzer.py:351-443) ---------------------------------------------------------
// Top-label confidence p[argmax] / sum(p), correctness, and the moments of the confidences of the CORRECT images
// (the bandwidth rule uses their standard deviation, :389-392).  stats: n_correct, sum conf, sum conf^2 (double).
__global__ void top_label_kernel(const float* __restrict__ probs, const int32_t* __restrict__ labels, int N, int C,
                                 float* __restrict__ conf, int32_t* __restrict__ correct, double* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double n1 = 0.0, s1 = 0.0, s2 = 0.0;
  if (i < N) {
    const float* p = probs + (size_t)i * C;
    int best = 0;
    float bp = p[0];
    double sum = 0.0;
    for (int c = 0; c < C; ++c) {
      const float v = fminf(fmaxf(p[c], 1e-38f), 1.f);    // the reference clips to [1e-256, 1 - 1e-256] in float64
      sum += (double)v;
      if (v > bp) {
        bp = v;
        best = c;
      }
    }
    const float cf = (float)((double)fminf(fmaxf(bp, 1e-38f), 1.f) / sum);
    const int hit = best == labels[i];
    conf[i] = cf;
    correct[i] = hit;
    if (hit) {
      n1 = 1.0;
      s1 = (double)cf;
      s2 = (double)cf * (double)cf;
    }
  }
  // warp reduce, one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    n1 += __shfl_down_sync(0xffffffffu, n1, o);
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
  }
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
__global__ void __launch_bounds__(128) kde_triweight_kernel(const float* __restrict__ data,
                                                            const int32_t* __restrict__ flags, int n, double h,
                                                            double inv_n, double x0, double dx, int G, double lo,
                                                            double hi, double* __restrict__ out) {
  __shared__ float sd[512];
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
      sd[t] = (i < n && (flags == nullptr || flags[i] != 0)) ? data[i] : 1e30f;
    }
    __syncthreads();
    if (inside) {
      const int m = min(512, n - base);
      for (int t = 0; t < m; ++t) {
        const double d = (double)sd[t];
        if (d > 1e29) continue;
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

// per-image NLL / Brier (MSE) / top-1 hit, block-reduced in a fixed order; partial[block][3]
__global__ void __launch_bounds__(256) dataset_metrics_kernel(const float* __restrict__ probs,
                                                              const int32_t* __restrict__ labels, int N, int C,
                                                              float* __restrict__ partial) {
  __shared__ float sh[3][256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float nll = 0.f, mse = 0.f, hit = 0.f;
  if (i < N) {
    const float* p = probs + (size_t)i * C;
    const int y = labels[i];
    int best = 0;
    float bp = p[0];
    for (int c = 0; c < C; ++c) {
      const float v = p[c];
      const float t = c == y ? 1.f : 0.f;
      mse += (v - t) * (v - t);
      if (v > bp) {
        bp = v;
        best = c;
      }
    }
    // results_analyzer.py:499-500: clip to [1e-256, 1 - 1e-256] in float64; in float32 the lower clip is the
    // smallest normal and the upper clip is 1
    nll = -logf(fmaxf(p[y], 1.17549435e-38f));
    hit = best == y ? 1.f : 0.f;
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

__global__ void dataset_metrics_reduce_kernel(const float* __restrict__ partial, int blocks, float inv_n,
                                              float* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= 3) return;
  double a = 0.0;
  for (int b = 0; b < blocks; ++b) a += (double)partial[b * 3 + k];
  out[k] = (float)(a * inv_n);
}

}  // namespace bnn

using namespace bnn;

extern "C" {

int bnn_exit_head(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                  const float* wt, const float* bias, const bnn_drop_desc* drop, float* sum_p, float* sum_logit,
                  float* sum_plogp, float* logits_out, int accumulate, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(feat && wt && bias && sum_p && sum_logit && sum_plogp, "bnn_exit_head: null pointer");
  BNN_REQUIRE(B >= 0 && S_local >= 0 && HW > 0 && F > 0 && C > 0, "bnn_exit_head: bad geometry");
  if (drop && drop->kind != BNN_DROP_NONE) {
    BNN_REQUIRE(drop->p >= 0.f && drop->p <= 1.f, "dropout probability has to be between 0 and 1, but got %g",
                drop->p);
    BNN_REQUIRE(drop->kind != BNN_DROP_MASKSEMBLES || (drop->masks && drop->n_masks > 0),
                "bnn_exit_head: Masksembles site without a mask table");
  }
  if (B == 0) return BNN_OK;
  const size_t smem = ((size_t)HEAD_SCHUNK * F + (size_t)HEAD_THREADS * HEAD_SCHUNK * (C > 32 ? 4 : 1) +
                       (size_t)HEAD_SCHUNK * C +
                       2 * (size_t)C + HEAD_WARPS + 2 * HEAD_SCHUNK) *
                      sizeof(float);
  BNN_REQUIRE(smem <= 200 * 1024, "bnn_exit_head: F=%d, C=%d need %zu bytes of shared memory", F, C, smem);
  DropParams dp = make_drop_params(drop, F);
  dp.batch = B;
  cudaStream_t st = (cudaStream_t)stream;
#define BNN_HEAD_LAUNCH(T)                                                                                         \
  do {                                                                                                             \
    BNN_CUDA_OK(cudaFuncSetAttribute(exit_head_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    exit_head_kernel<T><<<B, HEAD_THREADS, smem, st>>>((const T*)feat, feat_has_samples, B, S_local, HW, F, C, wt, \
                                                       bias, dp, sum_p, sum_logit, sum_plogp, logits_out,          \
                                                       accumulate);                                                \
  } while (0)
  switch (dtype) {
    case BNN_F32: BNN_HEAD_LAUNCH(float); break;
    case BNN_F16: BNN_HEAD_LAUNCH(__half); break;
    case BNN_BF16: BNN_HEAD_LAUNCH(__nv_bfloat16); break;
    default: set_error("unknown dtype code %d", dtype); return BNN_E_ARG;
  }
#undef BNN_HEAD_LAUNCH
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_finalize(const float* sum_p, const float* sum_logit, const float* sum_plogp, int E, int B, int C,
                 int S_total, float* mean_p, float* mean_logit, float* ens_p, float* ens_logit, float* entropy,
                 float* ens_entropy, float* exp_entropy, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(sum_p && sum_logit && sum_plogp && mean_p && mean_logit && ens_p && ens_logit && entropy &&
                  ens_entropy && exp_entropy,
              "bnn_finalize: null pointer");
  BNN_REQUIRE(E > 0 && B >= 0 && C > 0 && S_total > 0, "bnn_finalize: bad geometry");
  if (B == 0) return BNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int total = B * C;
  const float inv_s = 1.f / (float)S_total;
  finalize_means_kernel<<<(total + 127) / 128, 128, 0, st>>>(sum_p, sum_logit, E, B, C, inv_s, mean_p, mean_logit,
                                                             ens_p, ens_logit);
  BNN_LAUNCH_OK();
  finalize_entropy_kernel<<<(E * B + 127) / 128, 128, 0, st>>>(mean_p, ens_p, sum_plogp, E * B, C, inv_s, entropy,
                                                               ens_entropy, exp_entropy);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_calibration_bins(const float* probs, const int32_t* labels, int N, int C, int n_bins, float* conf,
                         int32_t* correct, float* bin_stats, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(probs && labels && conf && correct && bin_stats && N >= 0 && C > 0 && n_bins > 0,
              "bnn_calibration_bins: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  BNN_CUDA_OK(cudaMemsetAsync(bin_stats, 0, sizeof(float) * 3 * n_bins, st));
  if (N == 0) return BNN_OK;
  calibration_kernel<<<(N + 127) / 128, 128, 0, st>>>(probs, labels, N, C, n_bins, conf, correct, bin_stats);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_dataset_metrics(const float* probs, const int32_t* labels, int N, int C, float* workspace, float* out,
                        void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(probs && labels && workspace && out && N > 0 && C > 0, "bnn_dataset_metrics: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (N + 255) / 256;
  dataset_metrics_kernel<<<blocks, 256, 0, st>>>(probs, labels, N, C, workspace);
  BNN_LAUNCH_OK();
  dataset_metrics_reduce_kernel<<<1, 32, 0, st>>>(workspace, blocks, 1.f / (float)N, out);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_confidence_exit(const float* probs, int E, int N, int C, int first_exit, float threshold, int diff,
                        int32_t* exit_idx, float* best_probs, int32_t* exit_hist, void* stream) {
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

int bnn_top_label(const float* probs, const int32_t* labels, int N, int C, float* conf, int32_t* correct,
                  double* stats, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(probs && labels && conf && correct && stats && N >= 0 && C > 0, "bnn_top_label: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  BNN_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 3, st));
  if (N == 0) return BNN_OK;
  top_label_kernel<<<(N + 127) / 128, 128, 0, st>>>(probs, labels, N, C, conf, correct, stats);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_kde_triweight(const float* data, const int32_t* flags, int n, double bw, double n_points, double x0,
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
