// Counter-based mask stream (Philox-4x32-10, Salmon et al. SC'11) and the stochastic-site
// device functions shared by the stand-alone dropout kernel, the conv epilogues and the exit head.
//
// Contract (bit-exact mirror of oracle/philox.py, checked by tests/test_gpu_masks.py):
//   key = (seed lo, seed hi); counter = (e >> 3 lo, e >> 35, sample, stream): one block covers 8 elements;
//   element e uses the 16-bit half (e & 1) of word (e >> 1) & 3 (low half first)
//   keep <=> p < 1 && half >= min(rint(p * 2^16), 2^16 - 1)
// with e the per-sample element index (NHWC order for 4-D activations, b*F + f for 2-D, b*C + c
// for channel-wise dropout).  Replaces the torch global RNG behind F.dropout (resnet18.py:210).
#pragma once
#include "common.cuh"

namespace bnn {

__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = rint((double)p * 65536.0);
  if (t > 65535.0) t = 65535.0;
  if (t < 0.0) t = 0.0;
  return (uint32_t)t;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// the four words of block `blk` (raw generator output)
__device__ __forceinline__ uint4 philox_block(uint64_t seed, uint32_t stream, uint32_t sample, uint64_t blk) {
  return philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), sample, stream),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// 8 keep bits of the elements [8*blk, 8*blk + 7]: bit i <=> element 8*blk + i is kept
__device__ __forceinline__ uint32_t keep_bits8(const uint4 r, uint32_t thr) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  // VIMNMX.U16x2 compares both halves of a word at once and hands back the two (a >= b) predicates: 20 instead of 28
  // instructions per block (the fused-dropout epilogues are issue-bound)
  const uint32_t t2 = thr | (thr << 16);
  bool h0, l0, h1, l1, h2, l2, h3, l3;
  __vibmax_u16x2(r.x, t2, &h0, &l0);
  __vibmax_u16x2(r.y, t2, &h1, &l1);
  __vibmax_u16x2(r.z, t2, &h2, &l2);
  __vibmax_u16x2(r.w, t2, &h3, &l3);
  return (l0 ? 1u : 0u) | (h0 ? 2u : 0u) | (l1 ? 4u : 0u) | (h1 ? 8u : 0u) | (l2 ? 16u : 0u) | (h2 ? 32u : 0u) |
         (l3 ? 64u : 0u) | (h3 ? 128u : 0u);
#endif
  return ((r.x & 0xffffu) >= thr ? 1u : 0u) | ((r.x >> 16) >= thr ? 2u : 0u) | ((r.y & 0xffffu) >= thr ? 4u : 0u) |
         ((r.y >> 16) >= thr ? 8u : 0u) | ((r.z & 0xffffu) >= thr ? 16u : 0u) | ((r.z >> 16) >= thr ? 32u : 0u) |
         ((r.w & 0xffffu) >= thr ? 64u : 0u) | ((r.w >> 16) >= thr ? 128u : 0u);
}
__device__ __forceinline__ uint32_t philox_keep8(uint64_t seed, uint32_t stream, uint32_t sample, uint64_t blk,
                                                 uint32_t thr) {
  return keep_bits8(philox_block(seed, stream, sample, blk), thr);
}

// the 16-bit draw of one element (generic / scalar paths)
__device__ __forceinline__ uint32_t philox_half(uint64_t seed, uint32_t stream, uint32_t sample, uint64_t e) {
  const uint4 r = philox_block(seed, stream, sample, e >> 3);
  const uint32_t l = ((uint32_t)e >> 1) & 3u;
  const uint32_t w = l == 0 ? r.x : (l == 1 ? r.y : (l == 2 ? r.z : r.w));
  return (e & 1u) ? (w >> 16) : (w & 0xffffu);
}

// Device-side view of bnn_drop_desc with the derived constants.
struct DropParams {
  int kind;
  uint32_t thr;
  float scale;      // 1/(1-p), 0 when p >= 1
  uint64_t seed;
  uint32_t stream_id, sample0;
  int batch;
  const float* masks;
  int n_masks, cnt0;
  int channels;     // row length of `masks`
  // Masksembles gathered layout (optional): slot of channel c among the kept channels of mask row r (int16
  // [n_masks][channels], -1 = dropped), the inverse table (int16 [n_masks][compact_c], -1 = padding) and the
  // row length of the compact output
  const int16_t* compact_pos;
  const int16_t* compact_idx;
  int compact_c;
  int nchw_flat;    // element indices in NCHW-flattened order (site behind a Flatten of a spatial map)
};

inline DropParams make_drop_params(const bnn_drop_desc* d, int channels) {
  DropParams q{};
  if (d == nullptr || d->kind == BNN_DROP_NONE) {
    q.kind = BNN_DROP_NONE;
    q.batch = 1;
    return q;
  }
  q.kind = d->kind;
  q.thr = drop_threshold(d->p);
  q.scale = d->p >= 1.0f ? 0.0f : 1.0f / (1.0f - d->p);
  q.seed = d->seed;
  q.stream_id = d->stream_id;
  q.sample0 = d->sample0;
  q.batch = d->batch > 0 ? d->batch : 1;
  q.masks = d->masks;
  q.n_masks = d->n_masks;
  q.cnt0 = d->cnt0;
  q.channels = channels;
  q.nchw_flat = d->kind == BNN_DROP_ELEMENT ? d->nchw_flat : 0;
  if (d->kind == BNN_DROP_MASKSEMBLES && d->compact_pos != nullptr) {
    q.compact_pos = d->compact_pos;
    q.compact_idx = d->compact_idx;
    q.compact_c = d->compact_c;
  }
  return q;
}

// multiplier for one element (slow generic path; the vector paths amortise one Philox block over 8 elements)
__device__ __forceinline__ float drop_factor(const DropParams& q, uint32_t sample_local, uint64_t e_elem,
                                             uint64_t e_chan, int c) {
  if (q.kind == BNN_DROP_MASKSEMBLES) {
    const int row = (int)(((int64_t)q.cnt0 + q.sample0 + sample_local) % q.n_masks);
    return __ldg(q.masks + (size_t)row * q.channels + c);
  }
  const uint64_t e = q.kind == BNN_DROP_CHANNEL ? e_chan : e_elem;
  const uint32_t h = philox_half(q.seed, q.stream_id, q.sample0 + sample_local, e);
  return (q.scale != 0.f && h >= q.thr) ? q.scale : 0.f;
}

}  // namespace bnn
