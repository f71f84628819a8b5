// Exit-head and statistics kernels (sm_100a, CUDA cores; HBM/L2-bound, no tensor-core shape here:
// the FC is F x C with C = 10..100 per image).
//
//   exit_head_kernel : global average pool -> stochastic site -> Linear(F, C) -> softmax ->
//                      warp-shuffle accumulation over the local MC samples.  One CTA per image;
//                      all samples of that image are processed by the same CTA so the classifier
//                      weights are read once per image and the running sums never leave the SM.
//   finalize_kernel  : means, cumulative exit ensembles, entropies.
//   calibration_kernel: top-label confidence / correctness + equal-width bins.
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"

namespace bnn {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_WARPS = HEAD_THREADS / 32;
constexpr int HEAD_SCHUNK = 16;  // samples staged in shared memory at a time

template <typename T>
struct __align__(8 * sizeof(T)) Vec8h {
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


// ---- tensor-core variant of the head GEMM (wide heads: C > 32, 16-bit feature storage) ---------------------------
// logits[16 x C] = pooled[16 x F] * W[C x F]^T with mma.sync.m16n8k16 (legacy warp-level MMA: the GEMM is 16 rows per
// image, far below a tcgen05 tile).  Both operands are split into a 16-bit high part and a 16-bit remainder and three
// products are accumulated in fp32 (hi*hi + lo*hi + hi*lo), which keeps the result within ~2^-20 of the fp32 FFMA
// form - the head must not add error on top of the convolution path.  A fragments come from shared memory
// ([16][F + 8] 16-bit, the +8 padding makes the 32 lanes hit 32 banks), B fragments straight from the nn.Linear weight
// layout [C][F] in L2.  Warp w owns the 8-class tiles w, w + 8, ...
template <typename TM>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                                        uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// K permutation: inside every 32-wide K chunk lane t of a quad owns the PHYSICAL features k0 + 8t .. k0 + 8t + 7 and
// feeds them to two consecutive MMAs (slots 2t, 2t+1, 2t+8, 2t+9 <- physical 8t + {0,1,2,3}, then 8t + {4,5,6,7}).  A
// and B use the same assignment, so the products are unchanged, but every fragment becomes ONE 16-byte access: the B
// fragments of a class row are two full 32-byte sectors per quad instead of eight half-used ones (ncu on the C4 heads:
// the kernel sat at 72 % of the L1TEX pipe with 8 sectors per 4-byte LDG request, 30.5 M sectors per launch), and the A
// fragments are LDS.128 instead of four LDS.32.  HEAD_APAD keeps the 16-byte A accesses of a quarter-warp (rows g, g+1)
// on disjoint banks: (F + 32) * 2 bytes = 64 (mod 128).
constexpr int HEAD_APAD = 32;

template <typename TM, int MT>
__device__ __forceinline__ void head_gemm_mma(const TM* __restrict__ a_hi, const TM* __restrict__ a_lo,
                                              float* __restrict__ logits, const TM* __restrict__ w_hi,
                                              const TM* __restrict__ w_lo, const float* __restrict__ bias, int F, int C,
                                              int ns, int tid) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int pitch = F + HEAD_APAD;
  const int n_tiles = (C + 7) >> 3;
  for (int nt = warp; nt < n_tiles; nt += HEAD_WARPS) {
    const int n = nt * 8 + g;                               // the class whose weights this lane loads (B fragment)
    const bool nv = n < C;
    const TM* wh = w_hi + (size_t)(nv ? n : 0) * F + 8 * t;
    const TM* wl = w_lo + (size_t)(nv ? n : 0) * F + 8 * t;
    float acc[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][i] = 0.f;
#pragma unroll 2
    for (int k0 = 0; k0 < F; k0 += 32) {
      // one 16-byte weight fragment pair (hi + lo) feeds two MMA k-steps of all MT 16-sample row tiles
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      const uint4 bh = nv ? __ldg(reinterpret_cast<const uint4*>(wh + k0)) : z4;
      const uint4 bl = nv ? __ldg(reinterpret_cast<const uint4*>(wl + k0)) : z4;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m * 16 >= ns) break;                            // row tiles past the last sample of this pass
        const size_t r0 = (size_t)(m * 16 + g) * pitch + 8 * t + k0;
        const uint4 h0 = *reinterpret_cast<const uint4*>(a_hi + r0);
        const uint4 h1 = *reinterpret_cast<const uint4*>(a_hi + r0 + 8 * pitch);
        const uint4 l0 = *reinterpret_cast<const uint4*>(a_lo + r0);
        const uint4 l1 = *reinterpret_cast<const uint4*>(a_lo + r0 + 8 * pitch);
        {
          const uint32_t ah[4] = {h0.x, h1.x, h0.y, h1.y}, al[4] = {l0.x, l1.x, l0.y, l1.y};
          mma16816<TM>(acc[m], al, bh.x, bh.y);             // small terms first
          mma16816<TM>(acc[m], ah, bl.x, bl.y);
          mma16816<TM>(acc[m], ah, bh.x, bh.y);
        }
        {
          const uint32_t ah[4] = {h0.z, h1.z, h0.w, h1.w}, al[4] = {l0.z, l1.z, l0.w, l1.w};
          mma16816<TM>(acc[m], al, bh.z, bh.w);
          mma16816<TM>(acc[m], ah, bl.z, bl.w);
          mma16816<TM>(acc[m], ah, bh.z, bh.w);
        }
      }
    }
    // accumulator fragment: (row g, cols 2t, 2t+1), (row g + 8, cols 2t, 2t+1)
    const int c0 = nt * 8 + 2 * t;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int sl = m * 16 + g + (i >> 1) * 8, cc = c0 + (i & 1);
        if (sl < ns && cc < C) logits[sl * C + cc] = acc[m][i] + __ldg(bias + cc);
      }
  }
  __syncthreads();
}

template <typename TM>
__device__ __forceinline__ TM to_tm(float v);
template <>
__device__ __forceinline__ __half to_tm<__half>(float v) { return from_f32<__half>(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 to_tm<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ float to_tm<float>(float v) { return v; }

// fp32 weights [C][F] (the nn.Linear layout) -> 16-bit high part and 16-bit remainder
template <typename TM>
__global__ void split16_kernel(const float* __restrict__ w, TM* __restrict__ hi, TM* __restrict__ lo, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = w[i];
  const TM h = to_tm<TM>(v);
  hi[i] = h;
  lo[i] = to_tm<TM>(v - to_f32<TM>(h));
}

// Features of the tensor-core exit head: global average pool -> stochastic site -> hi / lo split.  One thread per
// (row = sample x image, feature octet): x_hi to a[row][f], and (when the features are not exactly representable in 16
// bits: a pooled mean, or a dropout scale that is not a power of two) the remainder x_lo to a[row][F + f].
template <typename T>
__global__ void __launch_bounds__(256) head_prep_kernel(const T* __restrict__ feat, int feat_has_samples, int B, int S_local,
                                                        int HW, int F, DropParams dp, T* __restrict__ a, int a_pitch,
                                                        int with_lo) {
  // 32-bit index arithmetic (the launcher checks S_local * B * F / 8 < 2^31): three 64-bit divisions per thread cost more
  // than the Philox block of this one-octet-per-thread kernel
  const uint32_t octs = (uint32_t)F / 8u;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (uint32_t)S_local * (uint32_t)B * octs) return;
  const uint32_t row = idx / octs;                       // s * B + b
  const int f0 = (int)(idx - row * octs) * 8;
  const int s = (int)(row / (uint32_t)B), b = (int)(row - (uint32_t)s * (uint32_t)B);
  const T* src = feat + (((size_t)(feat_has_samples ? s : 0) * B + b) * HW) * F + f0;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = 0; p < HW; ++p) {
    const Vec8h<T> v = *reinterpret_cast<const Vec8h<T>*>(src + (size_t)p * F);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += to_f32<T>(v.v[j]);
  }
  uint32_t k8 = 0xffu;
  float fac = 1.f / (float)HW;
  if (dp.kind == BNN_DROP_ELEMENT || dp.kind == BNN_DROP_CHANNEL) {
    k8 = dp.scale == 0.f ? 0u : philox_keep8(dp.seed, dp.stream_id, dp.sample0 + s, ((uint64_t)b * F + f0) >> 3, dp.thr);
    fac *= dp.scale;
  }
  Vec8h<T> vh, vl;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = ((k8 >> j) & 1u) ? acc[j] * fac : 0.f;
    if (dp.kind == BNN_DROP_MASKSEMBLES) v *= drop_factor(dp, (uint32_t)s, 0, 0, f0 + j);
    vh.v[j] = to_tm<T>(v);
    vl.v[j] = to_tm<T>(v - to_f32<T>(vh.v[j]));
  }
  *reinterpret_cast<Vec8h<T>*>(a + (size_t)row * a_pitch + f0) = vh;
  if (with_lo) *reinterpret_cast<Vec8h<T>*>(a + (size_t)row * a_pitch + F + f0) = vl;
}

// Soft-max + running sums over the samples for logits that already sit in global memory (the tensor-core head): one
// CTA of SM_WARPS warps per image, lanes over the classes (NPL per lane); warp w walks samples w, w + SM_WARPS, ... in
// order and the per-warp partial sums are combined in a fixed order (deterministic).  No barriers inside the sample
// loop - the block-per-image chunked form of exit_head_kernel spent ~7 us per 16-sample chunk in barriers and dependent
// loads (83 us per C4 head), a single warp per image 49 us in its own dependent shuffle / expf chain.
// warps per image: 16 for up to 128 classes (S = 32 ... 128 samples: two to eight rows per warp), 8 beyond (the per-warp
// partial sums live in static shared memory)
constexpr int SMW_NARROW = 16;
template <int NPL, int SM_WARPS>
__global__ void __launch_bounds__(SM_WARPS * 32) head_softmax_warp_kernel(
    const float* __restrict__ logits_in, int pitch, int B, int S_local, int C, float* __restrict__ sum_p,
    float* __restrict__ sum_logit, float* __restrict__ sum_plogp, float* __restrict__ logits_out, int accumulate) {
  __shared__ float part_p[SM_WARPS][NPL * 32], part_l[SM_WARPS][NPL * 32], part_pl[SM_WARPS];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ap[NPL], al[NPL], apl = 0.f;
#pragma unroll
  for (int j = 0; j < NPL; ++j) ap[j] = al[j] = 0.f;
  float v[NPL], nxt[NPL];
  auto load = [&](int s, float (&dst)[NPL]) {
    const float* row = logits_in + ((size_t)s * B + b) * pitch;
#pragma unroll
    for (int j = 0; j < NPL; ++j) dst[j] = (lane + 32 * j < C) ? __ldg(row + lane + 32 * j) : -INFINITY;
  };
  if (warp < S_local) load(warp, nxt);
  for (int s = warp; s < S_local; s += SM_WARPS) {
#pragma unroll
    for (int j = 0; j < NPL; ++j) v[j] = nxt[j];
    if (s + SM_WARPS < S_local) load(s + SM_WARPS, nxt);   // the next row is in flight while this one is reduced
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NPL; ++j) mx = fmaxf(mx, v[j]);
    mx = warp_max(mx);
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j)
      if (lane + 32 * j < C) den += expf(v[j] - mx);
    den = warp_sum(den);
    const float inv = 1.f / den, logden = logf(den);
    float pl = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int c = lane + 32 * j;
      if (c < C) {
        const float z = v[j] - mx, e = expf(z);
        pl += e * inv * (z - logden);
        ap[j] += e * inv;
        al[j] += v[j];
        if (logits_out) logits_out[((size_t)s * B + b) * C + c] = v[j];
      }
    }
    apl += warp_sum(pl);
  }
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    part_p[warp][lane + 32 * j] = ap[j];
    part_l[warp][lane + 32 * j] = al[j];
  }
  if (lane == 0) part_pl[warp] = apl;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += SM_WARPS * 32) {
    float tp = 0.f, tl = 0.f;
#pragma unroll
    for (int w = 0; w < SM_WARPS; ++w) {
      tp += part_p[w][c];
      tl += part_l[w][c];
    }
    const size_t o = (size_t)b * C + c;
    sum_p[o] = (accumulate ? sum_p[o] : 0.f) + tp;
    sum_logit[o] = (accumulate ? sum_logit[o] : 0.f) + tl;
  }
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < SM_WARPS; ++w) t += part_pl[w];
    sum_plogp[b] = (accumulate ? sum_plogp[b] : 0.f) + t;
  }
}

// Narrow heads (C <= 32) with all samples in parallel: one WARP per (sample, image) row.  The block-per-image kernel
// above walks the samples of an image in chunks with barriers between its phases and takes ~50 us per C2 exit whatever
// the bandwidth (8 MB of pooled features): a latency chain.  Here 8192 rows are 8192 independent warps: lane l owns the
// feature octets l, l + 32, ... (one 16-byte load per pixel, one Philox block per octet), multiplies them with the
// classifier weights held in shared memory ([C][F] fp32, the nn.Linear layout, split into two 4-float planes per octet
// so that the 16-byte reads of a warp are conflict-free) and the C partial sums are reduced with warp shuffles.  Logits
// go to a [rows][pitch] fp32 workspace; head_softmax_warp_kernel accumulates them per image in sample order.
template <typename T, int CMAX>
__global__ void __launch_bounds__(256) head_rows_kernel(const T* __restrict__ feat, int feat_has_samples, int B, int S_local,
                                                        int HW, int F, int C, const float* __restrict__ w_cf,
                                                        const float* __restrict__ bias, DropParams dp,
                                                        float* __restrict__ logits, int pitch, int rows_per_cta,
                                                        float feat_scale) {
  extern __shared__ float w_s[];                       // [C][2][F / 2]: plane h holds features 8 * oct + 4 * h + (0..3)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = F >> 1;
  for (int i = tid; i < C * F; i += 256) {
    const int c = i / F, f = i - c * F;
    w_s[c * F + ((f >> 2) & 1) * half + (f >> 3) * 4 + (f & 3)] = __ldg(w_cf + i);
  }
  __syncthreads();
  const int64_t rows = (int64_t)S_local * B;
  const int64_t r_end = min(rows, ((int64_t)blockIdx.x + 1) * rows_per_cta);
  const float inv_hw = feat_scale / (float)HW;
  for (int64_t row = (int64_t)blockIdx.x * rows_per_cta + warp; row < r_end; row += 8) {
    const int s = (int)((uint32_t)row / (uint32_t)B), b = (int)((uint32_t)row - (uint32_t)s * (uint32_t)B);   // rows < 2^31 (launcher)
    const T* src = feat + (((size_t)(feat_has_samples ? s : 0) * B + b) * HW) * F;
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
    for (int f0 = lane * 8; f0 < F; f0 += 256) {
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int p = 0; p < HW; ++p) {
        const Vec8h<T> v = *reinterpret_cast<const Vec8h<T>*>(src + (size_t)p * F + f0);
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
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[j] = ((k8 >> j) & 1u) ? a[j] * fac : 0.f;
        if (dp.kind == BNN_DROP_MASKSEMBLES) x[j] *= drop_factor(dp, (uint32_t)s, 0, 0, f0 + j);
      }
      const float* wp = w_s + (f0 >> 3) * 4;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          const float4 w0 = *reinterpret_cast<const float4*>(wp + c * F);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + c * F + half);
          float t = acc[c];
          t = fmaf(x[0], w0.x, t);
          t = fmaf(x[1], w0.y, t);
          t = fmaf(x[2], w0.z, t);
          t = fmaf(x[3], w0.w, t);
          t = fmaf(x[4], w1.x, t);
          t = fmaf(x[5], w1.y, t);
          t = fmaf(x[6], w1.z, t);
          t = fmaf(x[7], w1.w, t);
          acc[c] = t;
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        const float t = warp_sum(acc[c]);
        if (lane == c) mine = t;
      }
    if (lane < C) logits[(size_t)row * pitch + lane] = mine + __ldg(bias + lane);
  }
}

// dynamic shared memory layout (floats):
//   pooled[F][HEAD_SCHUNK] | part[HEAD_THREADS * HEAD_SCHUNK * (C > 32 ? 4 : 1)] | logits[HEAD_SCHUNK][C] | acc_p[C] | acc_l[C] |
//   red[HEAD_WARPS] | smax[HEAD_SCHUNK] | sinv[HEAD_SCHUNK]
// MMA: the head GEMM runs on mma.sync (head_gemm_mma) - shared memory then holds the pooled features as two 16-bit
// planes a_hi / a_lo [SCH][F + HEAD_APAD] instead of pooled[F][HEAD_SCHUNK] + part[]; w_hi / w_lo are [C][F].
// SCH = samples per pass (HEAD_SCHUNK; the MMA form can take 64 = four 16-row tiles per weight fragment, see the
// launcher for why that is not the default).
template <typename T, bool MMA, int SCH>
__global__ void __launch_bounds__(HEAD_THREADS) exit_head_kernel(
    const T* __restrict__ feat, int feat_has_samples, int B, int S_local, int HW, int F, int C,
    const float* __restrict__ wt, const float* __restrict__ bias, DropParams dp, float* __restrict__ sum_p,
    float* __restrict__ sum_logit, float* __restrict__ sum_plogp, float* __restrict__ logits_out, int accumulate,
    const T* __restrict__ w_hi, const T* __restrict__ w_lo, float feat_scale,
    const float* __restrict__ logits_in = nullptr, int logits_pitch = 0) {
  // logits_in != nullptr: the classifier already ran as a tensor-core GEMM (bnn_exit_head_tc) and left fp32 logits
  // [S_local][B][logits_pitch] in global memory - only the soft-max / accumulation phases run here
  extern __shared__ float sm[];
  float* pooled = sm;
  float* part = pooled + (size_t)SCH * F;
  float* logits = part + HEAD_THREADS * SCH * (C > 32 ? 4 : 1);
  T* a_hi = reinterpret_cast<T*>(sm);
  T* a_lo = a_hi + (size_t)SCH * (F + HEAD_APAD);
  if constexpr (MMA) logits = reinterpret_cast<float*>(a_lo + (size_t)SCH * (F + HEAD_APAD));
  if (logits_in != nullptr) logits = sm;                // no feature / partial-sum staging in this mode
  float* acc_p = logits + (size_t)SCH * C;
  float* acc_l = acc_p + C;
  float* red = acc_l + C;
  float* smax = red + HEAD_WARPS;
  float* sinv = smax + SCH;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float inv_hw = feat_scale / (float)HW;      // feat_scale: real value of one LSB of 8-bit features, else 1

  for (int c = tid; c < C; c += HEAD_THREADS) {
    acc_p[c] = 0.f;
    acc_l[c] = 0.f;
  }
  float plogp_acc = 0.f;  // meaningful on lane 0 of each warp

  for (int s0 = 0; s0 < S_local; s0 += SCH) {
    const int ns = min(SCH, S_local - s0);
    __syncthreads();
    if (logits_in != nullptr) {
      for (int idx = tid; idx < ns * C; idx += HEAD_THREADS) {
        const int sl = idx / C, c = idx - sl * C;
        logits[idx] = __ldg(logits_in + ((size_t)(s0 + sl) * B + b) * logits_pitch + c);
      }
      __syncthreads();
    } else {
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
        // all loads of a 16-pixel group are issued before the first add: a thread walks only ns*F/8/256 items, so the
        // memory latency has to be covered inside the item (4 loads in flight left the kernel at 1.9 TB/s)
        int p = 0;
        for (; p + 16 <= HW; p += 16) {
          Vec8h<T> v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = *reinterpret_cast<const Vec8h<T>*>(src + (size_t)(p + u) * F);
#pragma unroll
          for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += to_f32<T>(v[u].v[j]);
        }
#pragma unroll 4
        for (; p < HW; ++p) {
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
        Vec8h<T> vh, vl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = ((k8 >> j) & 1u) ? a[j] * fac : 0.f;
          if (dp.kind == BNN_DROP_MASKSEMBLES) v *= drop_factor(dp, (uint32_t)s, 0, 0, f0 + j);
          if constexpr (MMA) {
            vh.v[j] = to_tm<T>(v);
            vl.v[j] = to_tm<T>(v - to_f32<T>(vh.v[j]));
          } else {
            pooled[(f0 + j) * SCH + sl] = v;
          }
        }
        if constexpr (MMA) {
          *reinterpret_cast<Vec8h<T>*>(a_hi + (size_t)sl * (F + HEAD_APAD) + f0) = vh;
          *reinterpret_cast<Vec8h<T>*>(a_lo + (size_t)sl * (F + HEAD_APAD) + f0) = vl;
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
        pooled[f * SCH + sl] = a;
      }
    }
    __syncthreads();
    // ---- phase 2: logits[ns x C] = pooled[ns x F] * W^T[F x C] + bias as a register-tiled small GEMM ----
    if constexpr (MMA)
      head_gemm_mma<T, SCH / 16>(a_hi, a_lo, logits, w_hi, w_lo, bias, F, C, ns, tid);
    else if (C > 32)
      head_gemm<4>(pooled, part, logits, wt, bias, F, C, ns, tid);
    else
      head_gemm<1>(pooled, part, logits, wt, bias, F, C, ns, tid);
    }
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

}  // namespace bnn

using namespace bnn;

extern "C" {

int bnn_split16(const float* w, void* hi, void* lo, int64_t n, int dtype, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(w && hi && lo && n >= 0, "bnn_split16: bad arguments");
  BNN_REQUIRE(dtype == BNN_F16 || dtype == BNN_BF16, "bnn_split16: dtype must be float16 or bfloat16");
  if (n == 0) return BNN_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dtype == BNN_F16)
    split16_kernel<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__half*)hi, (__half*)lo, n);
  else
    split16_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

static int exit_head_run(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                         const float* wt, const float* bias, const bnn_drop_desc* drop, float* sum_p, float* sum_logit,
                         float* sum_plogp, float* logits_out, int accumulate, void* stream, const void* w_hi,
                         const void* w_lo, float feat_scale = 1.f, const float* logits_in = nullptr, int logits_pitch = 0);

int bnn_exit_head(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                  const float* wt, const float* bias, const bnn_drop_desc* drop, float* sum_p, float* sum_logit,
                  float* sum_plogp, float* logits_out, int accumulate, void* stream) {
  BNN_REQUIRE(wt, "bnn_exit_head: null pointer");
  return exit_head_run(feat, dtype, feat_has_samples, B, S_local, HW, F, C, wt, bias, drop, sum_p, sum_logit, sum_plogp,
                       logits_out, accumulate, stream, nullptr, nullptr);
}

int bnn_exit_head_mma(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                      const void* w_hi, const void* w_lo, const float* bias, const bnn_drop_desc* drop, float* sum_p,
                      float* sum_logit, float* sum_plogp, float* logits_out, int accumulate, void* stream) {
  BNN_REQUIRE(w_hi && w_lo, "bnn_exit_head_mma: null pointer");
  BNN_REQUIRE(dtype == BNN_F16 || dtype == BNN_BF16, "bnn_exit_head_mma: features must be float16 or bfloat16");
  BNN_REQUIRE(F % 32 == 0, "bnn_exit_head_mma: F=%d must be a multiple of 32", F);
  return exit_head_run(feat, dtype, feat_has_samples, B, S_local, HW, F, C, nullptr, bias, drop, sum_p, sum_logit,
                       sum_plogp, logits_out, accumulate, stream, w_hi, w_lo);
}

int bnn_exit_head_tc(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                     const void* w3, const float* bias_pad, int c_pad, int with_lo, const bnn_drop_desc* drop, void* a_ws,
                     float* logits_ws, float* sum_p, float* sum_logit, float* sum_plogp, float* logits_out, int accumulate,
                     void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(feat && w3 && bias_pad && a_ws && logits_ws && sum_p && sum_logit && sum_plogp, "bnn_exit_head_tc: null pointer");
  BNN_REQUIRE(dtype == BNN_F16 || dtype == BNN_BF16, "bnn_exit_head_tc: features must be float16 or bfloat16");
  BNN_REQUIRE(B >= 0 && S_local >= 0 && HW > 0 && F > 0 && F % 64 == 0 && C > 0 && c_pad >= C && c_pad % 64 == 0,
              "bnn_exit_head_tc: bad geometry (F=%d must be a multiple of 64, c_pad=%d a multiple of 64 >= C=%d)", F, c_pad, C);
  if (drop && drop->kind != BNN_DROP_NONE) {
    BNN_REQUIRE(drop->p >= 0.f && drop->p <= 1.f, "dropout probability has to be between 0 and 1, but got %g", drop->p);
    BNN_REQUIRE(drop->kind != BNN_DROP_MASKSEMBLES || (drop->masks && drop->n_masks > 0),
                "bnn_exit_head_tc: Masksembles site without a mask table");
  }
  if (B == 0 || S_local == 0) {
    if (B > 0 && !accumulate) {
      BNN_CUDA_OK(cudaMemsetAsync(sum_p, 0, sizeof(float) * (size_t)B * C, (cudaStream_t)stream));
      BNN_CUDA_OK(cudaMemsetAsync(sum_logit, 0, sizeof(float) * (size_t)B * C, (cudaStream_t)stream));
      BNN_CUDA_OK(cudaMemsetAsync(sum_plogp, 0, sizeof(float) * (size_t)B, (cudaStream_t)stream));
    }
    return BNN_OK;
  }
  DropParams dp = make_drop_params(drop, F);
  dp.batch = B;
  const int a_pitch = with_lo ? 2 * F : F;
  const int64_t items = (int64_t)S_local * B * (F / 8);
  BNN_REQUIRE(items < (int64_t)1 << 31, "bnn_exit_head_tc: S_local * B * F / 8 = %lld exceeds 32-bit indexing", (long long)items);
  const int M = S_local * B;
  // 1. features: pool -> site -> hi [| lo]     2. logits = [x_hi | x_lo] x [w_hi | w_lo | w_hi]^T on tcgen05 (fp32 out)
  if (dtype == BNN_F16)
    head_prep_kernel<__half><<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)feat, feat_has_samples, B, S_local, HW, F, dp, (__half*)a_ws, a_pitch, with_lo);
  else
    head_prep_kernel<__nv_bfloat16><<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)feat, feat_has_samples, B, S_local, HW, F, dp, (__nv_bfloat16*)a_ws, a_pitch, with_lo);
  BNN_LAUNCH_OK();
  const int head_c = F / 64;
  if (int rc = conv_tc_head_gemm(a_ws, w3, bias_pad, logits_ws, dtype, M, a_pitch, c_pad, (with_lo ? 3 : 2) * head_c, head_c,
                                 stream))
    return rc;
  // 3. soft-max, running sums over the samples (deterministic order), optional per-sample logits
  if (C <= 512) {
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 32)
      head_softmax_warp_kernel<1, SMW_NARROW><<<B, SMW_NARROW * 32, 0, st>>>(logits_ws, c_pad, B, S_local, C, sum_p, sum_logit, sum_plogp, logits_out, accumulate);
    else if (C <= 128)
      head_softmax_warp_kernel<4, SMW_NARROW><<<B, SMW_NARROW * 32, 0, st>>>(logits_ws, c_pad, B, S_local, C, sum_p, sum_logit, sum_plogp, logits_out, accumulate);
    else
      head_softmax_warp_kernel<16, 8><<<B, 8 * 32, 0, st>>>(logits_ws, c_pad, B, S_local, C, sum_p, sum_logit, sum_plogp, logits_out, accumulate);
    BNN_LAUNCH_OK();
    return BNN_OK;
  }
  return exit_head_run(nullptr, BNN_F32, 1, B, S_local, 1, F, C, nullptr, bias_pad, nullptr, sum_p, sum_logit, sum_plogp,
                       logits_out, accumulate, stream, nullptr, nullptr, 1.f, logits_ws, c_pad);
}

int bnn_exit_head_rows(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                       const float* w_cf, const float* bias, const bnn_drop_desc* drop, float* logits_ws, float* sum_p,
                       float* sum_logit, float* sum_plogp, float* logits_out, int accumulate, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(feat && w_cf && bias && logits_ws && sum_p && sum_logit && sum_plogp, "bnn_exit_head_rows: null pointer");
  BNN_REQUIRE(dtype == BNN_F16 || dtype == BNN_BF16 || dtype == BNN_F32, "bnn_exit_head_rows: float16 / bfloat16 / float32 features");
  BNN_REQUIRE(B >= 0 && S_local >= 0 && HW > 0 && F > 0 && F % 8 == 0 && C > 0 && C <= 32,
              "bnn_exit_head_rows: bad geometry (F=%d must be a multiple of 8, C=%d at most 32)", F, C);
  const size_t smem = (size_t)C * F * sizeof(float);
  BNN_REQUIRE(smem <= 200 * 1024, "bnn_exit_head_rows: F=%d x C=%d weights do not fit shared memory", F, C);
  if (drop && drop->kind != BNN_DROP_NONE) {
    BNN_REQUIRE(drop->p >= 0.f && drop->p <= 1.f, "dropout probability has to be between 0 and 1, but got %g", drop->p);
    BNN_REQUIRE(drop->kind != BNN_DROP_MASKSEMBLES || (drop->masks && drop->n_masks > 0),
                "bnn_exit_head_rows: Masksembles site without a mask table");
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0 || S_local == 0) {
    if (B > 0 && !accumulate) {
      BNN_CUDA_OK(cudaMemsetAsync(sum_p, 0, sizeof(float) * (size_t)B * C, st));
      BNN_CUDA_OK(cudaMemsetAsync(sum_logit, 0, sizeof(float) * (size_t)B * C, st));
      BNN_CUDA_OK(cudaMemsetAsync(sum_plogp, 0, sizeof(float) * (size_t)B, st));
    }
    return BNN_OK;
  }
  DropParams dp = make_drop_params(drop, F);
  dp.batch = B;
  const int64_t rows = (int64_t)S_local * B;
  BNN_REQUIRE(rows < (int64_t)1 << 31, "bnn_exit_head_rows: S_local * B = %lld exceeds 32-bit indexing", (long long)rows);
  // two CTAs per SM at most; every CTA fills its copy of the weights once and walks >= 8 rows per warp pass
  const int64_t max_grid = 2 * (int64_t)sm_count();
  int64_t grid = (rows + 7) / 8;
  if (grid > max_grid) grid = max_grid;
  const int rows_per_cta = (int)(((rows + grid - 1) / grid + 7) / 8 * 8);
  grid = (rows + rows_per_cta - 1) / rows_per_cta;
#define BNN_ROWS_LAUNCH(T, CM)                                                                                         \
  do {                                                                                                                 \
    BNN_CUDA_OK(cudaFuncSetAttribute(head_rows_kernel<T, CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    head_rows_kernel<T, CM><<<(unsigned)grid, 256, smem, st>>>((const T*)feat, feat_has_samples, B, S_local, HW, F, C,  \
                                                                w_cf, bias, dp, logits_ws, C, rows_per_cta, 1.f);       \
  } while (0)
#define BNN_ROWS_LAUNCH_T(T)           \
  do {                                 \
    if (C <= 16) BNN_ROWS_LAUNCH(T, 16); \
    else BNN_ROWS_LAUNCH(T, 32);       \
  } while (0)
  if (dtype == BNN_F16) BNN_ROWS_LAUNCH_T(__half);
  else if (dtype == BNN_BF16) BNN_ROWS_LAUNCH_T(__nv_bfloat16);
  else BNN_ROWS_LAUNCH_T(float);
#undef BNN_ROWS_LAUNCH_T
#undef BNN_ROWS_LAUNCH
  BNN_LAUNCH_OK();
  head_softmax_warp_kernel<1, SMW_NARROW><<<B, SMW_NARROW * 32, 0, st>>>(logits_ws, C, B, S_local, C, sum_p, sum_logit, sum_plogp, logits_out,
                                                           accumulate);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

int bnn_exit_head_q8(const void* feat_q, float feat_scale, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                     const float* wt, const float* bias, const bnn_drop_desc* drop, float* sum_p, float* sum_logit,
                     float* sum_plogp, float* logits_out, int accumulate, void* stream) {
  BNN_REQUIRE(wt && feat_scale > 0.f, "bnn_exit_head_q8: bad arguments");
  return exit_head_run(feat_q, BNN_I8, feat_has_samples, B, S_local, HW, F, C, wt, bias, drop, sum_p, sum_logit, sum_plogp,
                       logits_out, accumulate, stream, nullptr, nullptr, feat_scale);
}

static int exit_head_run(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                         const float* wt, const float* bias, const bnn_drop_desc* drop, float* sum_p, float* sum_logit,
                         float* sum_plogp, float* logits_out, int accumulate, void* stream, const void* w_hi,
                         const void* w_lo, float feat_scale, const float* logits_in, int logits_pitch) {
  if (int rc = check_device()) return rc;
  const bool mma = w_hi != nullptr;
  BNN_REQUIRE((feat || logits_in) && bias && sum_p && sum_logit && sum_plogp, "bnn_exit_head: null pointer");
  BNN_REQUIRE(B >= 0 && S_local >= 0 && HW > 0 && F > 0 && C > 0, "bnn_exit_head: bad geometry");
  if (drop && drop->kind != BNN_DROP_NONE) {
    BNN_REQUIRE(drop->p >= 0.f && drop->p <= 1.f, "dropout probability has to be between 0 and 1, but got %g",
                drop->p);
    BNN_REQUIRE(drop->kind != BNN_DROP_MASKSEMBLES || (drop->masks && drop->n_masks > 0),
                "bnn_exit_head: Masksembles site without a mask table");
  }
  if (B == 0) return BNN_OK;
  auto smem_for = [&](int sch) {
    const size_t tail = ((size_t)sch * C + 2 * (size_t)C + HEAD_WARPS + 2 * (size_t)sch) * sizeof(float);
    if (logits_in != nullptr) return tail;
    return mma ? 2 * (size_t)sch * (F + HEAD_APAD) * 2 + tail
               : ((size_t)sch * F + (size_t)HEAD_THREADS * sch * (C > 32 ? 4 : 1)) * sizeof(float) + tail;
  };
  // MMA form: 16 samples per pass.  64 per pass (four row tiles per weight fragment, 4x fewer weight re-reads from L2)
  // is available as an experiment (BNN_HEAD_SCH=64) but measured SLOWER at C4 (1.46 vs 1.22 ms for the five heads):
  // its 160 KB of shared memory leave one CTA per SM and nothing to hide the feature / weight load latency behind.
  // 32 samples per pass (two row tiles per weight fragment: half the weight re-reads) when the image has that many and
  // three CTAs still fit an SM; BNN_HEAD_SCH overrides (16 | 32 | 64).
  const char* sch_env = getenv("BNN_HEAD_SCH");
  int sch = HEAD_SCHUNK;
  if (mma) {
    const int want = sch_env ? atoi(sch_env) : (S_local >= 32 ? 32 : 16);
    if (want == 64 && S_local > HEAD_SCHUNK && smem_for(64) <= 160 * 1024) sch = 64;
    else if (want >= 32 && S_local > HEAD_SCHUNK && smem_for(32) <= 72 * 1024) sch = 32;
  }
  const size_t smem = smem_for(sch);
  BNN_REQUIRE(smem <= 200 * 1024, "bnn_exit_head: F=%d, C=%d need %zu bytes of shared memory", F, C, smem);
  DropParams dp = make_drop_params(drop, F);
  dp.batch = B;
  cudaStream_t st = (cudaStream_t)stream;
#define BNN_HEAD_LAUNCH(T, M, SCH)                                                                                  \
  do {                                                                                                              \
    BNN_CUDA_OK(cudaFuncSetAttribute(exit_head_kernel<T, M, SCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                     (int)smem));                                                                   \
    exit_head_kernel<T, M, SCH><<<B, HEAD_THREADS, smem, st>>>((const T*)feat, feat_has_samples, B, S_local, HW, F, \
                                                               C, wt, bias, dp, sum_p, sum_logit, sum_plogp,        \
                                                               logits_out, accumulate, (const T*)w_hi,              \
                                                               (const T*)w_lo, feat_scale, logits_in, logits_pitch);\
  } while (0)
#define BNN_HEAD_LAUNCH16(T)                                            \
  do {                                                                  \
    if (!mma) BNN_HEAD_LAUNCH(T, false, HEAD_SCHUNK);                   \
    else if (sch == 64) BNN_HEAD_LAUNCH(T, true, 64);                   \
    else if (sch == 32) BNN_HEAD_LAUNCH(T, true, 32);                   \
    else BNN_HEAD_LAUNCH(T, true, HEAD_SCHUNK);                         \
  } while (0)
  switch (dtype) {
    case BNN_F32: BNN_HEAD_LAUNCH(float, false, HEAD_SCHUNK); break;
    case BNN_F16: BNN_HEAD_LAUNCH16(__half); break;
    case BNN_BF16: BNN_HEAD_LAUNCH16(__nv_bfloat16); break;
    case BNN_I8: BNN_HEAD_LAUNCH(uint8_t, false, HEAD_SCHUNK); break;     // 8-bit features, fp32 FFMA classifier
    default: set_error("unknown dtype code %d", dtype); return BNN_E_ARG;
  }
#undef BNN_HEAD_LAUNCH16
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

}  // extern "C"
