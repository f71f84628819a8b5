// CUDA-core kernels of the MC-dropout / Masksembles inference path (sm_100a):
//   - error / device plumbing for the C ABI
//   - Philox test hooks
//   - NCHW -> NHWC input repack
//   - exact fp32 implicit-GEMM convolution with fused bias / residual / ReLU / dropout (FP32 parity
//     path + the layers the tensor-core kernel does not take: Cin % 64 != 0, 5x5, ...)
//   - stand-alone stochastic layer with prefix -> S-sample broadcast (HBM-bound)
//   - max-pool
// The exit-head / statistics kernels live in kernels_head.cu, the tcgen05 convolution in conv_tc.cu.
#include <stdarg.h>

#include "common.cuh"
#include "philox.cuh"

namespace bnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_dev_checked = -1000;  // cached result per process (device is fixed per process: one process per GPU)
static int g_sms = 0;

int check_device() {
  if (g_dev_checked != -1000) return g_dev_checked;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    cudaGetLastError();
    return BNN_E_CUDA;  // not cached: a device may appear later
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return BNN_E_CUDA;
  }
  g_sms = prop.multiProcessorCount;
  if (prop.major != 10) {
    set_error("device '%s' is sm_%d%d; this library is built for sm_100a only (no fallback)", prop.name,
              prop.major, prop.minor);
    g_dev_checked = BNN_E_ARCH;
    return g_dev_checked;
  }
  g_dev_checked = BNN_OK;
  return BNN_OK;
}

int sm_count() {
  if (check_device() != BNN_OK) return 0;
  return g_sms;
}

// ------------------------------------------------------------------------------------------
// Philox test hooks
// ------------------------------------------------------------------------------------------
__global__ void philox_words_kernel(uint32_t* out, int64_t count, uint64_t seed, uint32_t stream_id,
                                    uint32_t sample) {
  const int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk * 4 >= count) return;
  const uint4 r = philox_block(seed, stream_id, sample, (uint64_t)blk);
  const uint32_t v[4] = {r.x, r.y, r.z, r.w};
  for (int i = 0; i < 4; ++i)
    if (blk * 4 + i < count) out[blk * 4 + i] = v[i];
}

__global__ void philox_keep_kernel(uint8_t* out, int64_t count, uint32_t thr, int all_dropped, uint64_t seed,
                                   uint32_t stream_id, uint32_t sample) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  const uint32_t h = philox_half(seed, stream_id, sample, (uint64_t)e);
  out[e] = (!all_dropped && h >= thr) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC T
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int C, int HW, int64_t total,
                                    int pitch) {
  // one thread per output element; reads are strided by HW but the inputs are small (3 x 32 x 32 images).
  // pitch >= C: channels per pixel of y (the channels past C are left untouched: the caller zero-fills them once)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t t = i / C;
  const int hw = (int)(t % HW);
  const int64_t n = t / HW;
  y[t * pitch + c] = from_f32<T>(x[(n * C + c) * HW + hw]);
}

// ------------------------------------------------------------------------------------------
// fp32 implicit-GEMM convolution on CUDA cores
//   M = N*OH*OW output pixels, Ncols = Cout, K = KH*KW*Cin ((kh, kw, ci) order = NHWC patches)
//   64 x 64 output tile per CTA, 256 threads, 4 x 4 outputs per thread, K staged 16 at a time.
// ------------------------------------------------------------------------------------------
struct ConvGeom {
  int N, H, W, Cin, Cout, KH, KW, stride, pad, OH, OW, relu;
};

constexpr int SIMT_BM = 64, SIMT_BN = 64, SIMT_BK = 16;

template <typename T>
__global__ void __launch_bounds__(256) conv2d_simt_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias,
                                                          const T* __restrict__ res, T* __restrict__ y,
                                                          ConvGeom g, DropParams dp) {
  __shared__ float As[SIMT_BK][SIMT_BM + 4];
  __shared__ float Bs[SIMT_BK][SIMT_BN + 4];

  const int K = g.KH * g.KW * g.Cin;
  const int64_t M = (int64_t)g.N * g.OH * g.OW;
  const int64_t m0 = (int64_t)blockIdx.x * SIMT_BM;
  const int n0 = blockIdx.y * SIMT_BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // tx -> 4 couts, ty -> 4 pixels

  // loader mapping: each thread fetches 4 consecutive k for one row of A and one row of B
  const int lrow = tid / 4;          // 0..63
  const int lk = (tid % 4) * 4;      // 0,4,8,12
  const int64_t am = m0 + lrow;
  int a_n = 0, a_oh = 0, a_ow = 0;
  const bool a_valid = am < M;
  if (a_valid) {
    a_ow = (int)(am % g.OW);
    const int64_t t = am / g.OW;
    a_oh = (int)(t % g.OH);
    a_n = (int)(t / g.OH);
  }
  const int bn = n0 + lrow;
  const bool b_valid = bn < g.Cout;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SIMT_BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      float av = 0.f, bv = 0.f;
      if (k < K) {
        const int tap = k / g.Cin, ci = k - tap * g.Cin;
        const int kh = tap / g.KW, kw = tap - kh * g.KW;
        const int ih = a_oh * g.stride - g.pad + kh, iw = a_ow * g.stride - g.pad + kw;
        if (a_valid && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
          av = to_f32<T>(x[(((int64_t)a_n * g.H + ih) * g.W + iw) * g.Cin + ci]);
        if (b_valid) bv = __ldg(w + (int64_t)bn * K + k);
      }
      As[lk + j][lrow] = av;
      Bs[lk + j][lrow] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SIMT_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue: + bias (folded BN) [+ residual] [ReLU] [stochastic site] -> store
  const int64_t ohw = (int64_t)g.OH * g.OW;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int64_t img = m / ohw;                      // image index incl. the sample dimension
    const uint32_t s_local = (uint32_t)(img / dp.batch);
    const int64_t b = img % dp.batch;
    const int64_t pix = m - img * ohw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= g.Cout) continue;
      float v = acc[i][j] + __ldg(bias + c);
      if (res != nullptr) v += to_f32<T>(res[m * g.Cout + c]);
      if (g.relu) v = fmaxf(v, 0.f);
      if (dp.kind != BNN_DROP_NONE) {
        const uint64_t e_elem = (uint64_t)((b * ohw + pix) * g.Cout + c);
        const uint64_t e_chan = (uint64_t)(b * g.Cout + c);
        v *= drop_factor(dp, s_local, e_elem, e_chan, c);
      }
      y[m * g.Cout + c] = from_f32<T>(v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// stand-alone stochastic layer, with prefix -> S-sample broadcast
//   y[s][i] = x[(x_has_samples ? s : 0)][i] * factor(s, i),  i in [0, n_per) = B*H*W*C (one sample)
// HBM-bound: algorithmic bytes = (x_has_samples ? S : 1) * n_per * sizeof(T) read + S * n_per * sizeof(T)
// written.  Each thread owns 8 consecutive elements (one Philox block, 16-byte accesses for 16-bit T)
// and loops over the samples so the broadcast source is read from HBM once.
// ------------------------------------------------------------------------------------------
template <typename T>
struct __align__(8 * sizeof(T)) Vec8 {
  T v[8];
};

template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n_per,
                                                      int64_t per_image, int C, int S_local, int x_has_samples,
                                                      DropParams dp) {
  const int64_t v0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (v0 >= n_per) return;
  // vector path: the 8 elements exist, every sample slab stays 16-byte aligned, and (for the per-channel
  // kinds) the 8 elements sit inside one pixel's channel run
  const bool vec = (v0 + 8 <= n_per) && (n_per % 8 == 0) && (dp.kind == BNN_DROP_ELEMENT || C % 8 == 0) &&
                   dp.nchw_flat == 0;
  const int c0 = (int)(v0 % C);
  const int64_t b0 = v0 / per_image;

  Vec8<T> in;
  if (vec && !x_has_samples) in = *reinterpret_cast<const Vec8<T>*>(x + v0);   // broadcast source: read once

  if constexpr (sizeof(T) == 2) {
    // 16-bit storage, element-wise dropout on a tensor that already has a sample dimension (sites deeper in the suffix,
    // e.g. behind the max-pools of the VGG): the same mask arithmetic as the broadcast form below, input loaded and
    // scaled per sample
    if (vec && x_has_samples && dp.kind == BNN_DROP_ELEMENT && dp.scale != 0.f) {
      const uint32_t thr2s = dp.thr | (dp.thr << 16);
      const bool top_bit = dp.thr == 0x8000u;
      const uint64_t blk = (uint64_t)(v0 >> 3);
#pragma unroll 2
      for (int s = 0; s < S_local; ++s) {
        const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(x + (int64_t)s * n_per + v0);
        Vec8<T> sc;
#pragma unroll
        for (int j = 0; j < 8; ++j) sc.v[j] = from_f32<T>(to_f32<T>(v.v[j]) * dp.scale);
        const uint4 xs = *reinterpret_cast<const uint4*>(&sc);
        const uint4 r = philox_block(dp.seed, dp.stream_id, dp.sample0 + s, blk);
        auto mask2 = [&](uint32_t rv) -> uint32_t {
          uint32_t m;
          if (top_bit) {
            asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(m) : "r"(rv));
          } else {
            bool hi, lo;
            __vibmax_u16x2(rv, thr2s, &hi, &lo);
            m = (hi ? 0xffff0000u : 0u) | (lo ? 0x0000ffffu : 0u);
          }
          return m;
        };
        *reinterpret_cast<uint4*>(y + (int64_t)s * n_per + v0) =
            make_uint4(xs.x & mask2(r.x), xs.y & mask2(r.y), xs.z & mask2(r.z), xs.w & mask2(r.w));
      }
      return;
    }
    // 16-bit storage, element-wise dropout, prefix broadcast: scale once, then every sample is one Philox block,
    // four per-halfword compares (the 16-bit draws line up with the 16-bit elements) and four ANDs per 16 bytes
    if (vec && !x_has_samples && dp.kind == BNN_DROP_ELEMENT) {
      uint4 xs;
      {
        Vec8<T> sc;
#pragma unroll
        for (int j = 0; j < 8; ++j) sc.v[j] = from_f32<T>(to_f32<T>(in.v[j]) * dp.scale);
        xs = *reinterpret_cast<const uint4*>(&sc);
      }
      const uint32_t thr2 = dp.thr | (dp.thr << 16);
      const uint64_t blk = (uint64_t)(v0 >> 3);
      if (dp.thr == 0x8000u && dp.scale != 0.f) {
        // p = 0.5 (the reference's default): "draw >= 2^15" is the top bit of each 16-bit draw, and PRMT in its
        // sign-replicating mode (selector 0xbb99: bytes 1, 1, 3, 3 with the byte's msb copied to all eight bits) turns the
        // two top bits of a word into the two halfword masks in ONE instruction - 8 instead of ~24 integer instructions
        // per Philox block on the pipe that bounds this kernel.  Same bits as the general compare below.
#pragma unroll 2
        for (int s = 0; s < S_local; ++s) {
          const uint4 r = philox_block(dp.seed, dp.stream_id, dp.sample0 + s, blk);
          uint4 m;
          asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(m.x) : "r"(r.x));
          asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(m.y) : "r"(r.y));
          asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(m.z) : "r"(r.z));
          asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(m.w) : "r"(r.w));
          *reinterpret_cast<uint4*>(y + (int64_t)s * n_per + v0) = make_uint4(xs.x & m.x, xs.y & m.y, xs.z & m.z, xs.w & m.w);
        }
        return;
      }
#pragma unroll 2
      for (int s = 0; s < S_local; ++s) {
        const uint4 r = philox_block(dp.seed, dp.stream_id, dp.sample0 + s, blk);
        uint4 o;
        // the kernel is bound by the integer pipe (64 lanes / clock / SM: 60 integer instructions per Philox block x
        // 67 M blocks = 0.30 ms at C2), so the per-halfword compare is VIMNMX.U16x2 + two SELs, not the 7-instruction
        // software __vcmpgeu2
        auto keep2 = [&](uint32_t xv, uint32_t rv) -> uint32_t {
          bool hi, lo;
          __vibmax_u16x2(rv, thr2, &hi, &lo);                 // (r >= thr) per halfword
          return xv & ((hi ? 0xffff0000u : 0u) | (lo ? 0x0000ffffu : 0u));
        };
        o.x = keep2(xs.x, r.x);
        o.y = keep2(xs.y, r.y);
        o.z = keep2(xs.z, r.z);
        o.w = keep2(xs.w, r.w);
        if (dp.scale == 0.f) o = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(y + (int64_t)s * n_per + v0) = o;
      }
      return;
    }
  }

  for (int s = 0; s < S_local; ++s) {
    const T* xs = x + (x_has_samples ? (int64_t)s * n_per : 0);
    T* ys = y + (int64_t)s * n_per;
    if (vec) {
      if (x_has_samples) in = *reinterpret_cast<const Vec8<T>*>(xs + v0);
      float f[8];
      if (dp.kind == BNN_DROP_ELEMENT) {
        const uint32_t k8 = philox_keep8(dp.seed, dp.stream_id, dp.sample0 + s, (uint64_t)(v0 >> 3), dp.thr);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (dp.scale != 0.f && ((k8 >> j) & 1u)) ? dp.scale : 0.f;
      } else if (dp.kind == BNN_DROP_MASKSEMBLES) {
        const int row = (int)(((int64_t)dp.cnt0 + dp.sample0 + s) % dp.n_masks);
        const float4 m0 = __ldg(reinterpret_cast<const float4*>(dp.masks + (size_t)row * C + c0));
        const float4 m1 = __ldg(reinterpret_cast<const float4*>(dp.masks + (size_t)row * C + c0 + 4));
        f[0] = m0.x; f[1] = m0.y; f[2] = m0.z; f[3] = m0.w;
        f[4] = m1.x; f[5] = m1.y; f[6] = m1.z; f[7] = m1.w;
      } else {  // channel-wise: elements b*C + c0 .. +7 are one Philox block (C % 8 == 0)
        const uint64_t e = (uint64_t)(b0 * C + c0);
        const uint32_t k8 = philox_keep8(dp.seed, dp.stream_id, dp.sample0 + s, e >> 3, dp.thr);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (dp.scale != 0.f && ((k8 >> j) & 1u)) ? dp.scale : 0.f;
      }
      Vec8<T> out;
#pragma unroll
      for (int j = 0; j < 8; ++j) out.v[j] = from_f32<T>(to_f32<T>(in.v[j]) * f[j]);
      *reinterpret_cast<Vec8<T>*>(ys + v0) = out;
    } else {
      for (int j = 0; j < 8 && v0 + j < n_per; ++j) {
        const int64_t i = v0 + j;
        const int c = (int)(i % C);
        const int64_t b = i / per_image;
        uint64_t e = (uint64_t)i;
        if (dp.nchw_flat) {                       // NHWC buffer position -> NCHW-flattened element index
          const int64_t pix = (i - b * per_image) / C;
          e = (uint64_t)(b * per_image + (int64_t)c * (per_image / C) + pix);
        }
        const float f = drop_factor(dp, (uint32_t)s, e, (uint64_t)(b * C + c), c);
        ys[i] = from_f32<T>(to_f32<T>(xs[i]) * f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// max-pool k x k stride k, NHWC
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, int k, int OH,
                               int OW, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  int64_t t = i / C;
  const int ow = (int)(t % OW);
  t /= OW;
  const int oh = (int)(t % OH);
  const int64_t n = t / OH;
  float m = -INFINITY;
  for (int a = 0; a < k; ++a)
    for (int b = 0; b < k; ++b)
      m = fmaxf(m, to_f32<T>(x[((n * H + oh * k + a) * W + ow * k + b) * C + c]));
  y[i] = from_f32<T>(m);
}

// 8 channels (one 16-byte access for 16-bit storage) per thread; requires C % 8 == 0
template <typename T>
__global__ void __launch_bounds__(256) maxpool_vec8_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W,
                                                           int C8, int k, int OH, int OW, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % C8);
  int64_t t = i / C8;
  const int ow = (int)(t % OW);
  t /= OW;
  const int oh = (int)(t % OH);
  const int64_t n = t / OH;
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  for (int a = 0; a < k; ++a)
    for (int b = 0; b < k; ++b) {
      const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(x + (((n * H + oh * k + a) * W + ow * k + b) * C8 + c8) * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], to_f32<T>(v.v[j]));
    }
  Vec8<T> o;
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = from_f32<T>(m[j]);
  *reinterpret_cast<Vec8<T>*>(y + i * 8) = o;
}

// ------------------------------------------------------------------------------------------
// per-channel affine [+ ReLU], NHWC: y = relu?(x * scale[c] + bias[c]) - an eval-mode BatchNorm2d that cannot be folded
// into the convolution in front of it because a stochastic site sits in between (the converter's
// BayesianDropout2D(conv) -> BatchNorm2d -> ReLU chains, Hardware_Artifact/converter/pytorch/nn2bnn.py:32-45)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) channel_affine_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ bias, int C, int relu,
                                                             int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  float v = to_f32<T>(x[i]) * __ldg(scale + c) + __ldg(bias + c);
  if (relu) v = fmaxf(v, 0.f);
  y[i] = from_f32<T>(v);
}

// Masksembles "gathered" layout: y[s][b][px][j] = x[s or 0][b][px][idx[row(s)][j]] for the kept channels of sample
// s's mask row (utils.py:165-168 rotation), zeros in the padding slots.  One thread = 8 consecutive output slots of
// one pixel (one 16-byte store in 16-bit storage); the pixel's input channels come from one or two cache lines.
template <typename T>
__global__ void __launch_bounds__(256) masksembles_compact_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                  int64_t pixels, int C, int S_local,
                                                                  int x_has_samples, DropParams dp) {
  const int oct = dp.compact_c >> 3;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)S_local * pixels * oct) return;
  const int j8 = (int)(t % oct);
  const int64_t sp = t / oct;
  const int64_t px = sp % pixels;
  const int s = (int)(sp / pixels);
  const int row = (int)(((int64_t)dp.cnt0 + dp.sample0 + s) % dp.n_masks);
  const uint4 iw = __ldg(reinterpret_cast<const uint4*>(dp.compact_idx + (size_t)row * dp.compact_c) + j8);
  const uint32_t w[4] = {iw.x, iw.y, iw.z, iw.w};
  const T* xp = x + ((x_has_samples ? (int64_t)s * pixels : 0) + px) * C;
  Vec8<T> out;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int idx = (int)(int16_t)((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu));
    out.v[k] = idx >= 0 ? xp[idx] : from_f32<T>(0.f);
  }
  *reinterpret_cast<Vec8<T>*>(y + (sp * dp.compact_c + (int64_t)j8 * 8)) = out;
}

template <typename F>
static int dispatch_dtype(int dtype, F&& f) {
  switch (dtype) {
    case BNN_F32: return f((float*)nullptr);
    case BNN_F16: return f((__half*)nullptr);
    case BNN_BF16: return f((__nv_bfloat16*)nullptr);
    default: set_error("unknown dtype code %d", dtype); return BNN_E_ARG;
  }
}

// Prefix boundary of an element-wise MC-dropout site whose consumers are tensor-core convolutions that apply the keep
// mask THEMSELVES (conv_tc.cu, MASKED kernels): instead of S masked 16-bit copies of the boundary tensor (1.07 GB at C2)
// this writes ONE scaled copy x * 1/(1-p) (B images), the keep BITS of every sample (1 bit per element: 67 MB) and -
// for the fused 1x1 stride-2 projection shortcut, which reads only the even/even pixels - the masked even/even parity
// plane [S][B][H/2][W/2][C].  Same Philox blocks, same rounding as dropout_kernel: consumers reproduce its values bit
// for bit.  One thread = 8 consecutive channels of one pixel = one Philox block per sample.
template <typename T>
__global__ void __launch_bounds__(128) boundary_bits_kernel(const T* __restrict__ x, T* __restrict__ x_scaled,
                                                            uint8_t* __restrict__ bits, T* __restrict__ plane, int64_t n_pix,
                                                            int H, int W, int C, int S_local, DropParams dp) {
  // one thread = one pixel and 64 of its channels (eight Philox blocks per sample -> ONE 8-byte store of keep bits; a
  // warp writes 256 contiguous bytes)
  const int groups = C >> 6;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pix * groups) return;
  const int64_t pix = t / groups;                   // (b * H + h) * W + w
  const int c0 = (int)(t - pix * groups) << 6;
  const int64_t v0 = pix * C + c0;
  uint4 xs[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const Vec8<T> in = *reinterpret_cast<const Vec8<T>*>(x + v0 + 8 * o);
    Vec8<T> sc;
#pragma unroll
    for (int j = 0; j < 8; ++j) sc.v[j] = from_f32<T>(to_f32<T>(in.v[j]) * dp.scale);
    xs[o] = *reinterpret_cast<const uint4*>(&sc);
    *reinterpret_cast<uint4*>(x_scaled + v0 + 8 * o) = xs[o];
  }
  const int w = (int)(pix % W), h = (int)((pix / W) % H);
  const int64_t b = pix / ((int64_t)W * H);
  const bool ee = plane != nullptr && ((h | w) & 1) == 0;
  const int64_t plane_off = ((b * (H / 2) + h / 2) * (W / 2) + w / 2) * C + c0;
  const int64_t n_per = n_pix * C, plane_per = n_per / 4;
  const uint32_t thr2 = dp.thr | (dp.thr << 16);
  const uint64_t blk0 = (uint64_t)(v0 >> 3);
  for (int s = 0; s < S_local; ++s) {
    unsigned long long word = 0ull;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const uint4 r = philox_block(dp.seed, dp.stream_id, dp.sample0 + s, blk0 + o);
      // per-halfword compare masks (0xFFFF / 0) -> eight keep bits: take one bit of every halfword
      const uint32_t mx = __vcmpgeu2(r.x, thr2), my = __vcmpgeu2(r.y, thr2), mz = __vcmpgeu2(r.z, thr2), mw = __vcmpgeu2(r.w, thr2);
      const uint32_t k8 = (mx & 1u) | ((mx >> 15) & 2u) | ((my & 1u) << 2) | ((my >> 13) & 8u) | ((mz & 1u) << 4) |
                          ((mz >> 11) & 32u) | ((mw & 1u) << 6) | ((mw >> 9) & 128u);
      word |= (unsigned long long)(dp.scale == 0.f ? 0u : k8) << (8 * o);
      if (ee) {
        uint4 ov;
        ov.x = xs[o].x & mx;
        ov.y = xs[o].y & my;
        ov.z = xs[o].z & mz;
        ov.w = xs[o].w & mw;
        if (dp.scale == 0.f) ov = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(plane + (int64_t)s * plane_per + plane_off + 8 * o) = ov;
      }
    }
    *reinterpret_cast<unsigned long long*>(bits + (((int64_t)s * n_per + v0) >> 3)) = word;
  }
}

// same, plus unsigned 8-bit activations (max-pool and the quantising stochastic layer of the 8-bit path)
template <typename F>
static int dispatch_dtype_q(int dtype, F&& f) {
  if (dtype == BNN_I8) return f((uint8_t*)nullptr);
  return dispatch_dtype(dtype, f);
}

// Stochastic layer of the 8-bit path: y_q[s][i] = clip(rint(x[i or s][i] * in_to_out * factor_s(i)), 0, 255) with the same
// masks as dropout_kernel (factor = keep / (1 - p), a Masksembles row, ...).  x: 16-bit floats (the deterministic
// prefix -> first 8-bit tensor: in_to_out = 1 / step_y) or unsigned 8-bit (in_to_out = step_x / step_y).
template <typename TI>
__global__ void __launch_bounds__(256) dropout_q8_kernel(const TI* __restrict__ x, uint8_t* __restrict__ y, int64_t n_per,
                                                         int64_t per_image, int C, int S_local, int x_has_samples,
                                                         float in_to_out, DropParams dp) {
  const int64_t v0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (v0 >= n_per) return;
  const bool vec = (v0 + 8 <= n_per) && (n_per % 8 == 0) && (dp.kind == BNN_DROP_ELEMENT || C % 8 == 0) && dp.nchw_flat == 0;
  const int c0 = (int)(v0 % C);
  const int64_t b0 = v0 / per_image;
  for (int s = 0; s < S_local; ++s) {
    const TI* xs = x + (x_has_samples ? (int64_t)s * n_per : 0);
    uint8_t* ys = y + (int64_t)s * n_per;
    if (vec) {
      const Vec8<TI> in = *reinterpret_cast<const Vec8<TI>*>(xs + v0);
      float f[8];
      if (dp.kind == BNN_DROP_MASKSEMBLES) {
        const int row = (int)(((int64_t)dp.cnt0 + dp.sample0 + s) % dp.n_masks);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __ldg(dp.masks + (size_t)row * C + c0 + j);
      } else {
        const uint64_t e = dp.kind == BNN_DROP_ELEMENT ? (uint64_t)v0 : (uint64_t)(b0 * C + c0);
        const uint32_t k8 = dp.kind == BNN_DROP_NONE ? 0xffu : philox_keep8(dp.seed, dp.stream_id, dp.sample0 + s, e >> 3, dp.thr);
        const float sc = dp.kind == BNN_DROP_NONE ? 1.f : dp.scale;
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (sc != 0.f && ((k8 >> j) & 1u)) ? sc : 0.f;
      }
      Vec8<uint8_t> out;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        out.v[j] = from_f32<uint8_t>(__fmul_rn(__fmul_rn(to_f32<TI>(in.v[j]), in_to_out), f[j]));
      *reinterpret_cast<Vec8<uint8_t>*>(ys + v0) = out;
    } else {
      for (int j = 0; j < 8 && v0 + j < n_per; ++j) {
        const int64_t i = v0 + j;
        const int c = (int)(i % C);
        const int64_t b = i / per_image;
        uint64_t e = (uint64_t)i;
        if (dp.nchw_flat) {
          const int64_t pix = (i - b * per_image) / C;
          e = (uint64_t)(b * per_image + (int64_t)c * (per_image / C) + pix);
        }
        const float f = dp.kind == BNN_DROP_NONE ? 1.f : drop_factor(dp, (uint32_t)s, e, (uint64_t)(b * C + c), c);
        ys[i] = from_f32<uint8_t>(__fmul_rn(__fmul_rn(to_f32<TI>(xs[i]), in_to_out), f));
      }
    }
  }
}

}  // namespace bnn

using namespace bnn;

extern "C" {

int bnn_version(void) { return 100; }
const char* bnn_last_error(void) { return g_err; }
int bnn_device_check(void) { return check_device(); }
int bnn_sm_count(void) { return sm_count(); }

int bnn_philox_words(uint32_t* out, int64_t count, uint64_t seed, uint32_t stream_id, uint32_t sample,
                     void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(out != nullptr && count >= 0, "bnn_philox_words: bad arguments");
  if (count == 0) return BNN_OK;
  const int64_t blocks = ((count + 3) / 4 + 255) / 256;
  philox_words_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, count, seed, stream_id, sample);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_philox_keep(uint8_t* out, int64_t count, float p, uint64_t seed, uint32_t stream_id, uint32_t sample,
                    void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(out != nullptr && count >= 0, "bnn_philox_keep: bad arguments");
  BNN_REQUIRE(p >= 0.f && p <= 1.f, "dropout probability has to be between 0 and 1, but got %g", p);
  if (count == 0) return BNN_OK;
  const int64_t blocks = (count + 255) / 256;
  philox_keep_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, count, drop_threshold(p),
                                                                         p >= 1.f ? 1 : 0, seed, stream_id, sample);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_nchw_to_nhwc(const float* x, void* y, int dtype, int N, int C, int H, int W, void* stream) {
  return bnn_nchw_to_nhwc_pitch(x, y, dtype, N, C, H, W, C, stream);
}

int bnn_nchw_to_nhwc_pitch(const float* x, void* y, int dtype, int N, int C, int H, int W, int pitch, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(x && y && N >= 0 && C > 0 && H > 0 && W > 0 && pitch >= C, "bnn_nchw_to_nhwc: bad arguments");
  const int64_t total = (int64_t)N * C * H * W;
  if (total == 0) return BNN_OK;
  return dispatch_dtype(dtype, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    nchw_to_nhwc_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (T*)y, C, H * W,
                                                                                             total, pitch);
    BNN_LAUNCH_OK();
    return BNN_OK;
  });
}

int bnn_conv2d_simt(const void* x, const float* w, const float* bias, const void* res, void* y, int dtype, int N,
                    int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int relu,
                    const bnn_drop_desc* drop, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(x && w && bias && y, "bnn_conv2d_simt: null pointer");
  BNN_REQUIRE(N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0,
              "bnn_conv2d_simt: bad geometry");
  ConvGeom g{N, H, W, Cin, Cout, KH, KW, stride, pad, 0, 0, relu};
  g.OH = (H + 2 * pad - KH) / stride + 1;
  g.OW = (W + 2 * pad - KW) / stride + 1;
  BNN_REQUIRE(g.OH > 0 && g.OW > 0, "bnn_conv2d_simt: empty output (%d x %d)", g.OH, g.OW);
  if (drop && drop->kind != BNN_DROP_NONE) {
    BNN_REQUIRE(drop->batch > 0 && N % drop->batch == 0, "bnn_conv2d_simt: N=%d not a multiple of batch=%d", N,
                drop->batch);
    BNN_REQUIRE(drop->p >= 0.f && drop->p <= 1.f, "dropout probability has to be between 0 and 1, but got %g",
                drop->p);
    BNN_REQUIRE(drop->kind != BNN_DROP_MASKSEMBLES || (drop->masks && drop->n_masks > 0),
                "bnn_conv2d_simt: Masksembles site without a mask table");
  }
  const int64_t M = (int64_t)N * g.OH * g.OW;
  if (M == 0) return BNN_OK;
  const DropParams dp = make_drop_params(drop, Cout);
  dim3 grid((unsigned)((M + SIMT_BM - 1) / SIMT_BM), (unsigned)((Cout + SIMT_BN - 1) / SIMT_BN));
  return dispatch_dtype(dtype, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    conv2d_simt_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, w, bias, (const T*)res, (T*)y, g, dp);
    BNN_LAUNCH_OK();
    return BNN_OK;
  });
}

int bnn_dropout(const void* x, void* y, int dtype, int64_t per_image, int C, int S_local, int x_has_samples,
                const bnn_drop_desc* drop, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(x && y && drop, "bnn_dropout: null pointer");
  BNN_REQUIRE(per_image > 0 && C > 0 && per_image % C == 0 && S_local >= 0 && drop->batch >= 0,
              "bnn_dropout: bad geometry");
  BNN_REQUIRE(drop->kind == BNN_DROP_ELEMENT || drop->kind == BNN_DROP_CHANNEL || drop->kind == BNN_DROP_MASKSEMBLES,
              "bnn_dropout: unknown kind %d", drop->kind);
  BNN_REQUIRE(drop->p >= 0.f && drop->p <= 1.f, "dropout probability has to be between 0 and 1, but got %g",
              drop->p);
  BNN_REQUIRE(drop->kind != BNN_DROP_MASKSEMBLES || (drop->masks && drop->n_masks > 0),
              "bnn_dropout: Masksembles site without a mask table");
  const int64_t n_per = per_image * drop->batch;
  if (n_per == 0 || S_local == 0) return BNN_OK;
  const DropParams dp = make_drop_params(drop, C);
  if (dp.compact_pos != nullptr) {
    BNN_REQUIRE(dp.compact_idx != nullptr && dp.compact_c > 0 && dp.compact_c % 8 == 0 && C < 32768,
                "bnn_dropout: bad compact layout (compact_c=%d, C=%d)", dp.compact_c, C);
    const int64_t pixels = n_per / C;
    const int64_t nthr = (int64_t)S_local * pixels * (dp.compact_c / 8);
    return dispatch_dtype(dtype, [&](auto* tag) {
      using T = std::remove_pointer_t<decltype(tag)>;
      masksembles_compact_kernel<T><<<(unsigned)((nthr + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
          (const T*)x, (T*)y, pixels, C, S_local, x_has_samples, dp);
      BNN_LAUNCH_OK();
      return BNN_OK;
    });
  }
  const int64_t threads = (n_per + 7) / 8;
  return dispatch_dtype(dtype, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    dropout_kernel<T><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const T*)x, (T*)y, n_per, per_image, C, S_local, x_has_samples, dp);
    BNN_LAUNCH_OK();
    return BNN_OK;
  });
}

int bnn_boundary_bits(const void* x, void* x_scaled, void* bits, void* plane_ee, int dtype, int B, int H, int W, int C,
                      int S_local, const bnn_drop_desc* drop, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(x && x_scaled && bits && drop, "bnn_boundary_bits: null pointer");
  BNN_REQUIRE(dtype == BNN_F16 || dtype == BNN_BF16, "bnn_boundary_bits: 16-bit activations only");
  BNN_REQUIRE(drop->kind == BNN_DROP_ELEMENT && drop->p >= 0.f && drop->p <= 1.f,
              "bnn_boundary_bits: element-wise dropout with 0 <= p <= 1 (got kind %d, p %g)", drop->kind, drop->p);
  BNN_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && C % 64 == 0 && S_local >= 0 && (plane_ee == nullptr || (H % 2 == 0 && W % 2 == 0)),
              "bnn_boundary_bits: bad geometry (C = %d must be a multiple of 64)", C);
  const int64_t n_pix = (int64_t)B * H * W;
  if (n_pix == 0) return BNN_OK;
  const DropParams dp = make_drop_params(drop, C);
  const int64_t threads = n_pix * (C / 64);
  if (dtype == BNN_F16)
    boundary_bits_kernel<__half><<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const __half*)x, (__half*)x_scaled, (uint8_t*)bits, (__half*)plane_ee, n_pix, H, W, C, S_local, dp);
  else
    boundary_bits_kernel<__nv_bfloat16><<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)x_scaled, (uint8_t*)bits, (__nv_bfloat16*)plane_ee, n_pix, H, W, C, S_local, dp);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_dropout_q8(const void* x, void* y, int in_dtype, int64_t per_image, int C, int S_local, int x_has_samples,
                   float in_to_out, const bnn_drop_desc* drop, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(x && y && drop && per_image >= 0 && C > 0 && S_local >= 0 && drop->batch >= 0, "bnn_dropout_q8: bad arguments");
  BNN_REQUIRE(in_dtype == BNN_F16 || in_dtype == BNN_BF16 || in_dtype == BNN_I8, "bnn_dropout_q8: input must be 16-bit or 8-bit");
  if (drop->kind != BNN_DROP_NONE) {
    BNN_REQUIRE(drop->p >= 0.f && drop->p <= 1.f, "dropout probability has to be between 0 and 1, but got %g", drop->p);
    BNN_REQUIRE(drop->kind != BNN_DROP_MASKSEMBLES || (drop->masks && drop->n_masks > 0),
                "bnn_dropout_q8: Masksembles site without a mask table");
  }
  const int64_t n_per = per_image * drop->batch;
  if (n_per == 0 || S_local == 0) return BNN_OK;
  const DropParams dp = make_drop_params(drop, C);
  const int64_t threads = (n_per + 7) / 8;
  return dispatch_dtype_q(in_dtype, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    if constexpr (std::is_same<T, float>::value) {
      return (int)BNN_E_ARG;
    } else {
      dropout_q8_kernel<T><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
          (const T*)x, (uint8_t*)y, n_per, per_image, C, S_local, x_has_samples, in_to_out, dp);
      BNN_LAUNCH_OK();
      return (int)BNN_OK;
    }
  });
}

int bnn_channel_affine(const void* x, void* y, const float* scale, const float* bias, int dtype, int64_t pixels, int C,
                       int relu, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(x && y && scale && bias && pixels >= 0 && C > 0, "bnn_channel_affine: bad arguments");
  const int64_t total = pixels * C;
  if (total == 0) return BNN_OK;
  return dispatch_dtype(dtype, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    channel_affine_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const T*)x, (T*)y, scale, bias, C, relu, total);
    BNN_LAUNCH_OK();
    return BNN_OK;
  });
}

int bnn_maxpool2d(const void* x, void* y, int dtype, int N, int H, int W, int C, int k, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(x && y && N >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && H / k > 0 && W / k > 0,
              "bnn_maxpool2d: bad arguments");
  const int OH = H / k, OW = W / k;
  const int64_t total = (int64_t)N * OH * OW * C;
  if (total == 0) return BNN_OK;
  return dispatch_dtype_q(dtype, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    if (C % 8 == 0) {
      const int64_t tv = total / 8;
      maxpool_vec8_kernel<T><<<(unsigned)((tv + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, H, W,
                                                                                             C / 8, k, OH, OW, tv);
    } else {
      maxpool_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, H, W,
                                                                                            C, k, OH, OW, total);
    }
    BNN_LAUNCH_OK();
    return BNN_OK;
  });
}

}  // extern "C"
