// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bnn_conv2d_tc in include/bnn_b200.h).
//
// Replaces conv2d -> BatchNorm(eval) [-> += residual] [-> ReLU] [-> MCDropout | Masksembles2D] of the
// reference networks (resnet18.py:32-48,:306-308; vgg19.py:121-143) for the layers behind the first
// stochastic site, where the work is S x the prefix and really is a dense contraction.
//
// GEMM view: D[M, Cout] = A[M, K] * W[Cout, K]^T,  M = N*OH*OW (N counts samples x images),
//            K = taps * Cin ordered (kh, kw, ci) = NHWC patches.
//   A is never materialised: for k-block (tap, 64-channel block) the 128 x 64 operand tile is ONE TMA box of
//   the NHWC activation tensor shifted by the tap offset; out-of-bounds rows/columns are zero-filled by the
//   TMA unit, which is exactly the convolution's zero padding.
//     stride 1: 4-D map (C, W, H, N), box (64, tw, th, tn), coords (c0, kw-pad, oh0+kh-pad, n0)
//     stride 2: the same memory viewed as 5-D (2C, W/2, 2, H/2, N) (parity planes); a stride-2 tap is a
//               dense box of one (row parity, column parity) plane: box (64, tw, 1, th, tn)
//   Tiles are 128 consecutive output pixels (tw = OW, th rows, tn images) x BN output channels.
//
// CTA = 2 + EW warps (EW = 8 epilogue warps, 16 for the swapped dropout epilogue), persistent over tiles:
//   warp 0  : TMA producer (one elected lane) - A box + W box per k-block into a STAGES-deep smem ring
//   warp 1  : TMEM allocator + MMA issuer (one lane): tcgen05.mma kind::f16, 128 x BN x 16, fp32 accumulators
//             in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2.. : epilogue - tcgen05.ld 32 lanes x 32 columns -> + folded-BN bias -> + residual -> ReLU ->
//             Philox dropout / Masksembles mask -> 16-bit pack -> 64-byte row stores
//
// Variants of the one kernel template (picked per launch from the geometry by conv_tc_run):
//   PAIR = 2 (CG2)  cta_group::2 - one 256-row MMA over a CTA pair, each CTA holds half of the weight k-block
//   SWAP            Cout = 128: weights are the MMA's A operand (M = 128 channels), 256 pixels its N; transposed epilogue
//   VH              SWAP on 16 x 16 maps: one haloed pixel tile serves the three vertical taps (two rings, two producers)
//   SWAP + PAIR = 2 sibling pair: two 128-channel groups over the same pixels on one CTA pair; Params::rw_kb keeps the
//                   group's whole weight matrix resident in shared memory (K <= 9 k-blocks)
//   Params::pm_nb2  position-major row tiles on 2x2 .. 8x8 maps: taps that read only zero padding are not issued
//   T = int8_t      u8 x s8 -> int32 (kind::i8) with a requantising epilogue
//   fused shortcut (cblocks2), fused head pool (pool_hw), grouped outputs, Masksembles gathered K / per-mask weights,
//   head-GEMM mode (head_c), in-kernel keep-bit masking (MASKED, opt-in)
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "common.cuh"
#include "philox.cuh"

namespace bnn {
namespace tc {

#ifndef BNN_PHILOX_UNROLL
#define BNN_PHILOX_UNROLL 2
#endif
constexpr int PHILOX_UNROLL = BNN_PHILOX_UNROLL;   // independent Philox chains per thread in the 256-column epilogue
constexpr int BM = 128;          // rows (output pixels) per tile
constexpr int MAX_SMEM_OPTIN = 232448;   // 227 KB of dynamic shared memory per CTA
constexpr int BK = 64;           // K elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
// EW = epilogue warps.  The operand-swapped kernel's epilogue with a fused element-wise dropout (Philox bits exchanged
// by shuffles, residual, 2-byte stores) is latency-bound with two warps per scheduler - ncu: 47 % tensor-pipe activity
// on layer2.0.1.conv2 - and runs 16 epilogue warps (0.86 -> 0.74 ms at C2); the plain epilogues are FASTER with 8
// (measured: 16 warps cost 4-7 % on the layers without dropout), so EW is a template parameter chosen per launch.
__host__ __device__ constexpr int num_threads(int ew) { return 64 + ew * 32; }
constexpr int A_TILE_BYTES = BM * BK * 2;

__host__ __device__ constexpr int b_tile_bytes(int bn) { return bn * BK * 2; }
// as many pipeline stages as fit in ~200 KB of shared memory (at most 8)
__host__ __device__ constexpr int stages_for(int bn, int mt) {
  return (200 * 1024) / (mt * A_TILE_BYTES + b_tile_bytes(bn)) > 8 ? 8
                                                                  : (200 * 1024) / (mt * A_TILE_BYTES + b_tile_bytes(bn));
}
__host__ __device__ constexpr int stages_for_bytes(int stage_bytes) {
  return (200 * 1024) / stage_bytes > 8 ? 8 : (200 * 1024) / stage_bytes;
}
__host__ __device__ constexpr int smem_bytes_pair(int bn, int mt, int pair, bool swap = false) {
  // operand-swapped cta_group::2 ("sibling pair"): per CTA one 128-pixel tile + its own group's full weight tile
  const int stage = (swap && pair == 2) ? A_TILE_BYTES + b_tile_bytes(bn)
                                        : mt * A_TILE_BYTES + (pair == 2 ? b_tile_bytes(bn) / 2 : b_tile_bytes(bn));
  return stages_for_bytes(stage) * stage + 1024 + 256;
}
__host__ __device__ constexpr int smem_bytes(int bn, int mt) {
  return stages_for(bn, mt) * (mt * A_TILE_BYTES + b_tile_bytes(bn)) + 1024 /*align slack*/ + 256 /*barriers*/;
}

struct Params {
  int M, Cout, n_tiles_n, num_tiles;
  int taps, cblocks, Cin, stride, pad;
  int OH, OW, OHW;
  // grouped outputs: the GEMM's N = groups * cout_g concatenated output channels; group g is written to yg[g]
  // with row stride cout_g.  One group == an ordinary convolution (cout_g == Cout).
  int groups, cout_g;
  uint32_t relu_mask;   // bit g: ReLU on group g
  uint32_t center_mask; // bit g: group g is a 1x1 kernel stored in the centre tap of the 3x3 - only that tap runs
  // Masksembles gathered K (bnn_conv2d_tc_gathered): the input holds only the kept channels of each sample's mask
  // (Cin = kept count rounded up to 16), so the last channel block of a tap issues `last_ksteps` (< 4) MMA k-steps,
  // and the weights come in one pre-gathered set per mask row: rows [r * Cout, (r + 1) * Cout) for a tile of sample s,
  // r = (wsel_cnt0 + s) % wsel_n.  wsel_n == 0: one weight set.
  int last_ksteps;
  int wsel_n, wsel_cnt0, wsel_sample_px;
  // fused 1x1 stride-2 shortcut (bnn_conv2d_tc_shortcut): `cblocks2` extra k-blocks whose activation tile comes from a
  // SECOND tensor (tmap_a2: the block input x, centre tap of its stride-2 parity view) and whose weights are the
  // columns [taps * Cin, taps * Cin + 64 * cblocks2) of the K-concatenated weight matrix
  int cblocks2;
  int sc_dense;    // the shortcut operand is already the even/even parity plane [N][OH][OW][Cin2] (written by
                   // bnn_boundary_bits): a dense 4-D box instead of the centre tap of the 5-D parity view
  // fused global average pool (bnn_conv2d_tc_pooled): the output map of one image (pool_hw = OH * OW pixels, a power of
  // two <= 32, so one image = pool_hw adjacent lanes of an epilogue warp) is averaged with warp shuffles and only the
  // pooled row [N][Cout] is written - the exit head's input
  int pool_hw;
  int a_img_mod;   // > 0: the input has no sample dimension (deterministic prefix): output image n reads input n % a_img_mod
  // exit-head GEMM (bnn_exit_head_tc): a 1x1 "convolution" over [x_hi | x_lo] with weights [w_hi | w_lo | w_hi] - k-block
  // kb of the weights multiplies A block (kb < head_c ? kb : kb - head_c), so x_hi is used twice without being stored
  // twice; the epilogue writes fp32 logits
  // MASKED (sibling-pair kernel behind an element-wise MC-dropout site at the prefix boundary): x is the ONE scaled
  // deterministic tensor (a_img_mod images), mask_bits holds one keep bit per element of every sample
  // ([N][H][W][Cin / 8] bytes, written by bnn_boundary_bits); helper warps AND the bits into the tile in shared memory
  const uint8_t* mask_bits;
  int in_h, in_w;  // input map size (mask-bit addressing)
  // PM (position-major tiling, small maps): a row-tile is ONE output position (oh, ow) of 128 consecutive images, so a
  // tap is in bounds for all 128 rows or for none - the taps that would multiply only zero padding are not issued at all
  // (4x4 maps: 6.25 instead of 9 k-blocks per tile on average, 8x8: 7.6).  Row-tile order (pm_decode): tiles 2q and 2q+1
  // share a position (a CTA pair runs the same tap set); q walks the OHW positions of one 256-image block before moving
  // to the next block, so the tiles in flight consume whole images and the input stays in L2 between its 9 uses
  // (position-outermost order re-read the input from DRAM once per tap row).  pm_nb2 = image blocks rounded up to even.
  // RW (sibling-pair kernel, ONE group pair, K <= 9 k-blocks: ex1conv1 + layer2.0.0.conv1): each CTA's whole weight
  // matrix (rw_kb x 16 KB) is loaded once and stays in shared memory; only the 16 KB pixel tiles go through the ring
  // (rw_stages deep).  With the weights re-fetched per k-block a CTA ingests 32 KB per 512-cycle k-block = 64 B/clk, the
  // L2 -> SM ceiling, and the issuer waited for operands half of the time (profiles/r02_ncu_epilogue_bound.txt).
  int rw_kb, rw_stages;
  int l2pf;        // sibling-pair kernel, stride 2: L2-prefetch the pixel boxes of the tile `l2pf` rounds ahead (0 = off)
  int pm_nb2;      // > 0: position-major tiling
  int n_img;       // images (M / OHW)
  int head_c;      // > 0: channel blocks of x_hi
  int out_f32;     // epilogue stores float32 (non-swapped kernels only)
  float q_mult;    // 8-bit operands: output LSBs per accumulator LSB = w_scale * in_scale / out_scale (a power of two
                   // for the QKeras fixed-point formats, any float otherwise)
  int vh_a, vh_w;  // vertical-halo form: slots of the haloed-activation ring and of the weight ring
  // back-off (ns) of the waits with slack: producer <- free ring slot, epilogue <- accumulator full, MMA issuer <-
  // accumulator drained, MMA issuer <- operands landed (0 = tight loop).  BNN_TC_WAIT_NS="p,e,a,f"
  uint32_t wait_ns_prod, wait_ns_epi, wait_ns_acc, wait_ns_full;
  int exp_flags;   // MEASUREMENT ONLY (BNN_TC_EXP, results are garbage): bit 0 = do not load activation tiles, bit 1 = do
                   // not load weight tiles - isolates what operand delivery costs a launch; MASKED kernel: 4 = no
                   // masking, 8 = no proxy fence, 16 / 32 = CTA-scope instead of cluster-scope wait / arrive
  const float* bias;
  const void* res;
  void* yg[4];
  DropParams dp;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait with nanosleep back-off (experiment, BNN_TC_WAIT_NS): ncu counted 19 % of the issued instructions of the
// fused-dropout launch in try_wait loops, but backing off changes neither that launch nor the C2 step (try_wait already
// suspends the warp in hardware; profiles/r02_exp_wait_backoff.txt).  ns == 0 (default): plain try_wait loop.
__device__ __forceinline__ void mbar_wait_ns(uint64_t* bar, uint32_t parity, uint32_t ns) {
  if (mbar_try(bar, parity)) return;
  if (ns == 0) {
    mbar_wait(bar, parity);
    return;
  }
  do {
    __nanosleep(ns);
  } while (!mbar_try(bar, parity));
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// half of a weight tile, written to the same shared-memory offset of BOTH CTAs of the pair; each destination
// CTA's barrier (same offset) receives the bytes
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, "
      "%5}], [%2], %3;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// One elected lane of a fully converged warp (the MMA-issuer warp runs its loop with all 32 lanes so that the loop
// state - stage indices, shared-memory descriptors, TMEM addresses - is warp-uniform and lives in uniform registers;
// with the whole loop under `if (lane == 0)` every tcgen05.mma cost an ELECT + five R2UR.BROADCAST + predicate shuffling,
// ~140 cycles of dependent issue per 128-cycle MMA: ncu showed the issuing thread busy 80 % of the time).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 (fp16 / bf16 operands, fp32 accumulate) or - I8 - kind::i8
// (u8 x s8 operands, int32 accumulate)
template <bool I8 = false>
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  if constexpr (I8)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- cta_group::2 (CTA pair) variants ----
constexpr uint32_t kLeaderMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: rank 0 of the pair
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kLeaderMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kLeaderMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kLeaderMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// bring a box into L2 only (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(map), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
template <bool I8 = false>
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  if constexpr (I8)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// arrive on the barrier at this offset in the LEADER CTA of the pair (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kLeaderMask) : "memory");
}

// release at cluster scope: the arriving thread's shared-memory writes (of EITHER CTA of the pair) are visible to whoever
// acquires the leader's barrier at cluster scope
__device__ __forceinline__ void mbar_arrive_leader_release_cluster(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kLeaderMask) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP_C:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE_C;\n"
      "bra.uni WAIT_LOOP_C;\n"
      "WAIT_DONE_C:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// generic-proxy writes to shared memory (the mask warps' STS) -> visible to the async proxy (tensor-core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// same, arriving on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 bytes, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // bits  0-13 start address >> 4
  d |= (uint64_t)0 << 16;                           // bits 16-29 leading byte offset (unused: swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // bits 32-45 stride byte offset = 8 rows * 128 B
  d |= (uint64_t)1 << 46;                           // bits 46-47 descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                           // bits 61-63 layout: SWIZZLE_128B
  return d;
}

// instruction descriptor for kind::f16: fp32 accumulator, A/B type, both K-major, M = 128, N = BN
template <typename T>
__host__ __device__ constexpr uint32_t make_idesc_m(int m, int bn) {
  if (sizeof(T) == 1)   // kind::i8: D = S32 (2), A = activations, unsigned 8-bit (0), B = weights, signed 8-bit (1)
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
  const uint32_t ab = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  return (1u << 4) | (ab << 7) | (ab << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <typename T>
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  if (sizeof(T) == 1) return make_idesc_m<T>(BM, bn);
  const uint32_t ab = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;   // 0 = f16, 1 = bf16
  return (1u << 4)                 // D format f32
         | (ab << 7) | (ab << 10)  // A, B format
         | (0u << 15) | (0u << 16) // A, B K-major
         | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// SWAP (BN == 128, MT == 2): the roles of the operands are exchanged - the 128 x 64 weight tile is the MMA's
// A operand (M = 128 output channels) and the 256 x 64 pixel tile its B operand (N = 256), so that Cout = 128
// layers also issue 128 x 256 x 16 MMAs (a 128 x 128 MMA re-reads its operands from shared memory twice as often
// per FLOP and measured ~55% tensor-pipe utilisation).  The accumulator is then D^T: TMEM lanes = channels,
// columns = pixels, and the epilogue writes 2-byte elements that coalesce across the warp's 32 channels.
// MC2: the kernel runs as clusters of two CTAs that work on two adjacent row-tiles with the SAME weights; each CTA
// fetches half of every weight k-block and multicasts it into both shared memories (L2->SM weight traffic halves:
// 48 KB -> 32 KB per k-block at BN = 256).  A stage is reused only after the MMAs of BOTH CTAs released it.
// PAIR == 2 (CG2): the MMA itself spans the pair - tcgen05.mma.cta_group::2, M = 256 (128 rows per CTA), N = BN.
// Each CTA keeps only ITS half of the weight k-block in shared memory (the tensor cores of the two SMs exchange the
// halves), so a stage shrinks to 32 KB (6 stages) and weight reads from shared memory halve.  TMA loads of both
// CTAs signal the LEADER's full barrier; the leader's MMA thread commits (multicast) to the stage-empty and
// accumulator-full barriers of both CTAs; both epilogues arrive on the leader's accumulator-empty barrier.
// COMPACT (non-swapped epilogue only; the swapped one decides at run time): a fused Masksembles site may store
// only the kept channels of each sample's mask (DropParams::compact_pos), 2 bytes at a time.
// VH ("vertical halo", operand-swapped kernel, 3x3 stride-1 convolutions whose 256-pixel tile is ONE whole image of 16
// rows x 16 columns): the three vertical taps (kh = 0, 1, 2) of a (kw, channel block) read the SAME pixels shifted by
// one image row = 16 operand rows = 2048 bytes = two 128-byte-swizzle atoms.  The producer therefore loads ONE haloed
// tile (image rows -1..16, 18 x 16 pixels x 64 channels = 36 KB; rows -1 and 16 are zero-filled by the TMA unit) and
// the MMA issuer runs the three taps from it by offsetting the B-operand descriptor by kh * 2048 bytes.  Measured reason
// (profiles/r02_exp_operand_delivery.txt): this kernel is paced by the delivery of ACTIVATION tiles into the SM - 32 KB
// per 512-cycle k-block; with the activation loads switched off it runs at the MMA floor (1 430 vs 1 140 TFLOP/s), with
// the weight loads switched off nothing changes - so the activation bytes per k-block drop from 32 KB to 12 KB.
// Two rings instead of one: VH_A_SLOTS haloed activation tiles and VH_W_SLOTS weight tiles.
constexpr int VH_A_BYTES = 18 * 16 * BK * 2;      // 36 864
constexpr int VH_RING_BYTES = 220 * 1024;         // both rings together; the split is a launch parameter (vh_a / vh_w)
constexpr int VH_MAX_SLOTS = 12;
__host__ __device__ constexpr int smem_bytes_vh() { return VH_RING_BYTES + 1024 + 256; }

// position-major row-tile -> (output position, first image); see Params::pm_nb2
__device__ __forceinline__ void pm_decode(int m_tile, int ohw, int& pos, int& img0) {
  const int q = m_tile >> 1, bp = q / ohw;
  pos = q - bp * ohw;
  img0 = (2 * bp + (m_tile & 1)) * BM;
}

template <int BN, int MT, bool SWAP, int PAIR, bool COMPACT, int EW, typename T, bool VH = false, bool MASKED = false>
__global__ void __launch_bounds__(num_threads(EW) + (VH ? 32 : 0) + (MASKED ? 128 : 0), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_ah, const Params p) {
  constexpr bool MC2 = PAIR == 1;
  constexpr bool CG2 = PAIR == 2;
  constexpr bool PAIRED = PAIR != 0;
  // 8-bit operands (T = int8_t: unsigned 8-bit activations x signed 8-bit weights -> int32 accumulators, kind::i8): a
  // k-block is still one 128-byte swizzle row per pixel / weight row, i.e. 128 elements, and one MMA still consumes 32
  // bytes of K (32 elements), so only element COORDINATES, the instruction and the epilogue differ.
  constexpr bool I8 = sizeof(T) == 1;
  constexpr int BKE = I8 ? 2 * BK : BK;               // elements per k-block
  static_assert(!I8 || (!SWAP && !COMPACT && !VH && PAIR != 1), "8-bit operands: plain and cta_group::2 kernels only");
  static_assert(!VH || (SWAP && PAIR == 0 && BN == 128 && MT == 2), "VH is a variant of the operand-swapped kernel");
  static_assert(!(SWAP && PAIR == 1), "the multicast pairing of the operand-swapped kernel was retired");
  static_assert(!MASKED || (SWAP && PAIR == 2 && !VH), "in-kernel keep-bit masking: sibling-pair kernel only");
  // SCG2 (SWAP && CG2, "sibling pair"): the CTA pair works on the SAME 256 output pixels; CTA r owns output group
  // 2 * pair + r (its 128 weight rows are the pair-MMA's M rows r*128..) and loads only ITS half of the pixel tile (the
  // MMA's N = 256 columns are split over the pair), so a CTA ingests 16 KB of activations per k-block instead of 32 KB
  // - the operand whose delivery paces the operand-swapped kernel (profiles/r02_exp_operand_delivery.txt).
  constexpr bool SCG2 = SWAP && CG2;
  constexpr int A_STAGE = SCG2 ? A_TILE_BYTES : MT * A_TILE_BYTES;
  constexpr int B_TILE = (CG2 && !SWAP) ? b_tile_bytes(BN) / 2 : b_tile_bytes(BN);   // weights in THIS CTA's stage
  constexpr int STAGES = VH ? VH_MAX_SLOTS : stages_for_bytes(A_STAGE + B_TILE);   // VH: number of slot barriers
  constexpr int STAGE_BYTES = A_STAGE + B_TILE;
  constexpr int ACC_COLS = MT * BN;                 // one accumulator set: MT row-tiles of BN columns
  constexpr int TMEM_COLS = 2 * ACC_COLS;           // double-buffered (power of two, <= 512)
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  const int VH_A_SLOTS = p.vh_a, VH_W_SLOTS = p.vh_w;      // ring depths (vertical-halo form only)
  const bool RW = SCG2 && !MASKED && p.rw_kb > 0;          // resident weights (Params::rw_kb)
  const int n_stages = RW ? p.rw_stages : STAGES;          // ring depth in use (<= STAGES barriers)
  uint8_t* smem_b = smem + (VH ? VH_A_SLOTS * VH_A_BYTES : n_stages * A_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      smem + (VH ? VH_RING_BYTES : (RW ? n_stages * A_STAGE + p.rw_kb * B_TILE : STAGES * STAGE_BYTES)));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  // MASKED: afull = this CTA's pixel tile has landed (local); ready = both CTAs' mask warps are done with the stage (leader)
  uint64_t* afull_bar = tmem_empty + 2;
  uint64_t* ready_bar = afull_bar + (MASKED ? STAGES : 0);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ready_bar + (MASKED ? STAGES : 0));
  uint64_t* w_bar = reinterpret_cast<uint64_t*>(tmem_ptr) + 1;      // RW: the resident weights have landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.taps * p.cblocks;
  uint32_t cta_rank = 0;
  if constexpr (PAIRED) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  // scheduling unit: a CTA (tile = (row-tile, channel tile)) or a CTA pair (tile = (row-tile pair, channel tile))
  const int sched_id = PAIRED ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int sched_n = PAIRED ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], MC2 ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], CG2 ? 2 * EW : EW);   // one arrive per epilogue warp (of both CTAs)
    }
    mbar_init(w_bar, 1);
    if constexpr (MASKED)
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&afull_bar[i], 1);
        mbar_init(&ready_bar[i], 8);                  // four mask warps in each CTA of the pair
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CG2) tmem_alloc_2sm<TMEM_COLS>(tmem_ptr); else tmem_alloc<TMEM_COLS>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIRED) cluster_sync_all();   // the peer's barriers are initialised before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (VH && warp == 0) {
    // ===================== TMA producer, vertical-halo form =====================
    // barriers: full/empty[0 .. VH_A_SLOTS) belong to the activation ring, [VH_A_SLOTS .. STAGES) to the weight ring
    // Two producer threads, one per ring, each issuing in consumption order: warp 0 feeds the haloed activation tiles,
    // the extra warp behind the epilogue warps (index 2 + EW) feeds the weight tiles.  (One thread feeding both rings
    // in order lets the shallower ring throttle the other to a look-ahead of a single group - measured: 1 249 vs
    // 1 167 TFLOP/s for the per-tap kernel, against a 1 480 floor.)
    if (lane == 0) {
      int as = 0;
      uint32_t aph = 0;
      for (int tile = sched_id; tile < p.num_tiles; tile += sched_n) {
        for (int kw = 0; kw < 3; ++kw)
          for (int cb = 0; cb < p.cblocks; ++cb) {
            mbar_wait_ns(&empty_bar[as], aph ^ 1, p.wait_ns_prod);
            if (p.exp_flags & 1) {
              mbar_arrive(&full_bar[as]);                   // measurement mode: activations never fetched
            } else {
              mbar_expect_tx(&full_bar[as], VH_A_BYTES);
              tma_load_4d(smem_a + as * VH_A_BYTES, &tmap_ah, &full_bar[as], cb * BKE, kw - 1, -1, tile);
            }
            if (++as == VH_A_SLOTS) { as = 0; aph ^= 1; }
          }
        // fused 1x1 stride-2 shortcut: plain 2 x 128-pixel tiles of the block input's parity view
        for (int cb = 0; cb < p.cblocks2; ++cb) {
          mbar_wait_ns(&empty_bar[as], aph ^ 1, p.wait_ns_prod);
          mbar_expect_tx(&full_bar[as], A_STAGE);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint8_t* dst = smem_a + as * VH_A_BYTES + mt * A_TILE_BYTES;
            if (p.sc_dense)
              tma_load_4d(dst, &tmap_a2, &full_bar[as], cb * BKE, 0, mt * (BM / p.OW), tile);
            else
              tma_load_5d(dst, &tmap_a2, &full_bar[as], cb * BKE, 0, 0, mt * (BM / p.OW), tile);
          }
          if (++as == VH_A_SLOTS) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (VH && warp == 2 + EW) {
    // ===================== TMA producer of the weight ring (vertical-halo form) =====================
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      auto load_w = [&](int kcol) {
        mbar_wait_ns(&empty_bar[VH_A_SLOTS + ws], wph ^ 1, p.wait_ns_prod);
        if (p.exp_flags & 2) {
          mbar_arrive(&full_bar[VH_A_SLOTS + ws]);          // measurement mode: weights never fetched
        } else {
          mbar_expect_tx(&full_bar[VH_A_SLOTS + ws], B_TILE);
          tma_load_2d(smem_b + ws * B_TILE, &tmap_b, &full_bar[VH_A_SLOTS + ws], kcol, 0);
        }
        if (++ws == VH_W_SLOTS) { ws = 0; wph ^= 1; }
      };
      for (int tile = sched_id; tile < p.num_tiles; tile += sched_n) {
        for (int kw = 0; kw < 3; ++kw)
          for (int cb = 0; cb < p.cblocks; ++cb)
            for (int kh = 0; kh < 3; ++kh) load_w((kh * 3 + kw) * p.Cin + cb * BKE);
        for (int cb = 0; cb < p.cblocks2; ++cb) load_w(p.taps * p.Cin + cb * BKE);
      }
    }
  } else if (VH && warp == 1) {
    // ===================== MMA issuer, vertical-halo form =====================
    {
      constexpr uint32_t idesc = make_idesc<T>(MT * BM);
      const bool leader = elect_one();                      // the same lane issues every MMA and every commit
      int as = 0, ws = 0, acc = 0;
      uint32_t aph = 0, wph = 0, acc_phase = 0;
      for (int tile = sched_id; tile < p.num_tiles; tile += sched_n) {
        mbar_wait_ns(&tmem_empty[acc], acc_phase ^ 1, p.wait_ns_acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        uint32_t accum = 0;
        const int groups = 3 * p.cblocks + p.cblocks2;      // (kw, channel block) groups, then the shortcut k-blocks
        for (int g = 0; g < groups; ++g) {
          const int ntap = g < 3 * p.cblocks ? 3 : 1;
          mbar_wait_ns(&full_bar[as], aph, p.wait_ns_full);
          tc_fence_after();
          for (int kh = 0; kh < ntap; ++kh) {
            mbar_wait_ns(&full_bar[VH_A_SLOTS + ws], wph, p.wait_ns_full);
            tc_fence_after();
            const uint64_t w_desc = make_smem_desc(smem_u32(smem_b + ws * B_TILE));
            // tap kh reads image rows kh-1 .. kh+14 of the haloed tile: 16 pixels x 128 bytes further per kh
            const uint64_t p_desc = make_smem_desc(smem_u32(smem_a + as * VH_A_BYTES + kh * (16 * BK * 2)));
            if (leader) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_f16(d_tmem, w_desc + (uint64_t)(2 * k), p_desc + (uint64_t)(2 * k), idesc, (k | accum) != 0 ? 1u : 0u);
              umma_commit(&empty_bar[VH_A_SLOTS + ws]);
            }
            accum = 1;
            __syncwarp();
            if (++ws == VH_W_SLOTS) { ws = 0; wph ^= 1; }
          }
          if (leader) umma_commit(&empty_bar[as]);
          __syncwarp();
          if (++as == VH_A_SLOTS) { as = 0; aph ^= 1; }
        }
        if (leader) umma_commit(&tmem_full[acc]);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (MASKED && warp >= 2 + EW) {
    // ===================== mask warps (sibling pair behind the prefix-boundary MC-dropout site) =====================
    // One thread per pixel row of this CTA's 128-pixel tile: its 64 channels of the current channel block are eight
    // 16-byte chunks (128-byte swizzle: chunk j of row r sits at ((j ^ (r & 7)) << 4)); the pixel's keep bits for those
    // 64 channels are ONE 8-byte word of mask_bits.  AND, fence to the async proxy, hand the stage to the MMA issuer.
    const int r = (int)threadIdx.x - (2 + EW) * 32;       // 0 .. 127
    int stage = 0;
    uint32_t phase = 0;
    const int row_in_tile = r / p.OW, ow = r - row_in_tile * p.OW;
    const int cbytes = p.Cin >> 3;                          // mask bytes per input pixel
    for (int tile = sched_id; tile < p.num_tiles; tile += sched_n) {
      const int m_unit = tile / p.n_tiles_n;
      const int m0 = (m_unit * MT + (int)cta_rank) * BM;
      const bool live = m0 < p.M;
      const int img = live ? m0 / p.OHW : 0;               // sample x image index (NOT reduced modulo the batch)
      const int oh = (live ? (m0 - img * p.OHW) / p.OW : 0) + row_in_tile;
      auto bits_of = [&](int tap, int cb) -> uint64_t {
        const int kh = tap / 3, kw = tap - kh * 3;
        const int ih = 2 * oh + kh - 1, iw = 2 * ow + kw - 1;
        if (!live || ih < 0 || iw < 0 || ih >= p.in_h || iw >= p.in_w) return 0ull;     // zero padding: nothing to keep
        return __ldg(reinterpret_cast<const unsigned long long*>(
            p.mask_bits + (((size_t)img * p.in_h + ih) * p.in_w + iw) * cbytes + cb * 8));
      };
      // keep bits of the next PF k-blocks are in flight while a stage is processed (one dependent L2 load per stage
      // with a look-ahead of one paced the whole kernel: 2 600 cycles per k-block)
      constexpr int PF = 4;
      const int nkb = p.taps * p.cblocks;
      uint64_t ring[PF];
#pragma unroll
      for (int i = 0; i < PF; ++i) ring[i] = i < nkb ? bits_of(i / p.cblocks, i % p.cblocks) : 0ull;
      for (int kb = 0; kb < nkb; kb += PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
          if (kb + i >= nkb) break;
          const uint64_t cur = ring[i];
          const int nx = kb + i + PF;
          ring[i] = nx < nkb ? bits_of(nx / p.cblocks, nx % p.cblocks) : 0ull;
          mbar_wait(&afull_bar[stage], phase);
          uint8_t* row = smem_a + stage * A_STAGE + r * 128;
          if (!(p.exp_flags & 4))                     // (measurement switch: leave the tile unmasked)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4* q = reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4));
            uint4 v = *q;
            const uint32_t b8 = (uint32_t)(cur >> (8 * j)) & 0xffu;
            // bit i of b8 keeps 16-bit element i of the chunk: spread four bits into four 0x00 / 0xFF bytes with one
            // multiply, then duplicate each byte into a halfword mask with PRMT
            const uint32_t lo = (((b8 & 15u) * 0x00204081u) & 0x01010101u) * 0xffu;
            const uint32_t hi = (((b8 >> 4) * 0x00204081u) & 0x01010101u) * 0xffu;
            v.x &= __byte_perm(lo, 0u, 0x1100);
            v.y &= __byte_perm(lo, 0u, 0x3322);
            v.z &= __byte_perm(hi, 0u, 0x1100);
            v.w &= __byte_perm(hi, 0u, 0x3322);
            *q = v;
          }
          if (!(p.exp_flags & 8)) fence_proxy_async_smem();      // (measurement switch)
          __syncwarp();
          if (lane == 0) {
            if (p.exp_flags & 32) mbar_arrive_leader(&ready_bar[stage]); else mbar_arrive_leader_release_cluster(&ready_bar[stage]);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (SCG2 && !MASKED) {
        if (RW) {
          // this CTA's group (= its rank: one group pair) x all k-blocks, once; both CTAs' loads complete on the leader's
          // barrier
          if (cta_rank == 0) mbar_expect_tx(w_bar, 2 * p.rw_kb * B_TILE);
          for (int kb = 0; kb < p.rw_kb; ++kb)
            tma_load_2d_2sm(smem_b + kb * B_TILE, &tmap_b, w_bar, kb * BKE, (int)cta_rank * BN);
        }
      }
      for (int tile = sched_id; tile < p.num_tiles; tile += sched_n) {
        const int m_unit = tile / p.n_tiles_n, n_tile = tile - m_unit * p.n_tiles_n;
        const int m_tile = (PAIRED && !SCG2) ? 2 * m_unit + (int)cta_rank : m_unit;
        constexpr int NLOAD = SCG2 ? 1 : MT;          // 128-pixel tiles THIS CTA loads per k-block
        int img0[MT], oh0[MT];
        int pm_ow = 0;                                // PM: output column of this tile's position (0 otherwise)
#pragma unroll
        for (int mt = 0; mt < NLOAD; ++mt) {
          int m0 = SCG2 ? (m_unit * MT + (int)cta_rank) * BM : (m_tile * MT + mt) * BM;
          if (m0 >= p.M) m0 = 0;                      // row-tile past the end: load something valid, never stored
          img0[mt] = m0 / p.OHW;
          oh0[mt] = (m0 - img0[mt] * p.OHW) / p.OW;
          if (p.a_img_mod > 0) img0[mt] %= p.a_img_mod;
        }
        if (p.pm_nb2 > 0) {                           // (MT == 1) one position, 128 consecutive images
          int pos;
          pm_decode(m_tile, p.OHW, pos, img0[0]);     // image blocks past the end: the TMA unit zero-fills
          oh0[0] = pos / p.OW;
          pm_ow = pos - oh0[0] * p.OW;
        }
        if constexpr (SCG2 && !MASKED) {
          if (p.l2pf > 0 && p.stride == 2 && tile + p.l2pf * sched_n < p.num_tiles) {
            // the DRAM round trip of a later tile's pixels starts now: its TMA loads then hit L2
            const int t2 = tile + p.l2pf * sched_n, mu2 = t2 / p.n_tiles_n;
            int m2 = (mu2 * MT + (int)cta_rank) * BM;
            if (m2 >= p.M) m2 = 0;
            int i2 = m2 / p.OHW;
            const int o2 = (m2 - i2 * p.OHW) / p.OW;
            if (p.a_img_mod > 0) i2 %= p.a_img_mod;
            for (int tap = 0; tap < p.taps; ++tap) {
              const int rh = tap / 3 - p.pad, rw = tap - (tap / 3) * 3 - p.pad;
              const int hp = rh & 1, dh = (rh - hp) >> 1, wp = rw & 1, dw = (rw - wp) >> 1;
              for (int cb = 0; cb < p.cblocks; ++cb) tma_prefetch_5d(&tmap_a, wp * p.Cin + cb * BKE, dw, hp, o2 + dh, i2);
            }
          }
        }
        // SCG2: n_tile counts group PAIRS; this CTA's output group (and weight rows) is 2 * n_tile + rank
        const int grp_w = SCG2 ? 2 * n_tile + (int)cta_rank : n_tile;
        const bool center = (p.center_mask >> ((grp_w * BN) / p.cout_g)) & 1u;
        const int tap_lo = center ? 4 : 0, tap_hi = center ? 5 : p.taps;
        // weight set of this tile's MC sample (host guarantees that a tile / tile pair never straddles two samples)
        int wrow = grp_w * BN;
        if (p.wsel_n > 0) {
          int m0 = m_tile * MT * BM;
          if (m0 >= p.M) m0 = 0;
          wrow += ((p.wsel_cnt0 + m0 / p.wsel_sample_px) % p.wsel_n) * p.Cout;
        }
        for (int tap = tap_lo; tap < tap_hi; ++tap) {
          const int kh = p.taps == 1 ? p.pad : tap / 3, kw = p.taps == 1 ? p.pad : tap - (tap / 3) * 3;
          // stride 2: input row 2*oh + kh - pad -> (half-row index, row parity); same for columns
          const int rh = kh - p.pad, rw = kw - p.pad;                 // -1, 0, +1  (0 only for 1x1)
          const int hp = rh & 1, dh = (rh - hp) >> 1;                 // -1 -> (1, -1); 0 -> (0, 0); 1 -> (1, 0)
          const int wp = rw & 1, dw = (rw - wp) >> 1;
          if (p.pm_nb2 > 0) {                                         // the tap reads zero padding for EVERY row: skip
            const int ih = p.stride * oh0[0] + rh, iw = p.stride * pm_ow + rw;
            if (ih < 0 || ih >= p.in_h || iw < 0 || iw >= p.in_w) continue;
          }
          for (int cb = 0; cb < p.cblocks; ++cb) {
            const int a_cb = (p.head_c > 0 && cb >= p.head_c) ? cb - p.head_c : cb;     // exit-head GEMM: x_hi twice
            mbar_wait_ns(&empty_bar[stage], phase ^ 1, p.wait_ns_prod);
            if constexpr (CG2) {
              // both CTAs' loads complete on the leader's barrier: it expects the bytes of the pair
              if constexpr (MASKED) {
                // pixels: plain load that completes on THIS CTA's barrier (its mask warps pick the tile up there);
                // weights of both CTAs: on the leader's full barrier, as before
                if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * B_TILE);
                mbar_expect_tx(&afull_bar[stage], A_STAGE);
                tma_load_5d(smem_a + stage * A_STAGE, &tmap_a, &afull_bar[stage], wp * p.Cin + cb * BKE, dw, hp, oh0[0] + dh,
                            img0[0]);
              } else {
              if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], RW ? 2 * A_STAGE : 2 * STAGE_BYTES);
#pragma unroll
              for (int mt = 0; mt < NLOAD; ++mt) {
                uint8_t* a_dst = smem_a + stage * A_STAGE + mt * A_TILE_BYTES;
                if (p.stride == 1)
                  tma_load_4d_2sm(a_dst, &tmap_a, &full_bar[stage], a_cb * BKE, rw + pm_ow, oh0[mt] + rh, img0[mt]);
                else
                  tma_load_5d_2sm(a_dst, &tmap_a, &full_bar[stage], wp * p.Cin + cb * BKE, dw + pm_ow, hp, oh0[mt] + dh,
                                  img0[mt]);
              }
              }
              if (!RW)
                tma_load_2d_2sm(smem_b + stage * B_TILE, &tmap_b, &full_bar[stage], tap * p.Cin + cb * BKE,
                                SCG2 ? wrow : wrow + (int)cta_rank * (BN / 2));
            } else {
              if (p.exp_flags != 0 && !MC2) {
                // measurement mode: part of the operands is never fetched
                const uint32_t bytes = ((p.exp_flags & 1) ? 0 : A_STAGE) + ((p.exp_flags & 2) ? 0 : B_TILE);
                if (bytes == 0) mbar_arrive(&full_bar[stage]); else mbar_expect_tx(&full_bar[stage], bytes);
                if (!(p.exp_flags & 1)) {
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    uint8_t* a_dst = smem_a + stage * A_STAGE + mt * A_TILE_BYTES;
                    if (p.stride == 1)
                      tma_load_4d(a_dst, &tmap_a, &full_bar[stage], a_cb * BKE, rw + pm_ow, oh0[mt] + rh, img0[mt]);
                    else
                      tma_load_5d(a_dst, &tmap_a, &full_bar[stage], wp * p.Cin + cb * BKE, dw + pm_ow, hp, oh0[mt] + dh, img0[mt]);
                  }
                }
                if (!(p.exp_flags & 2))
                  tma_load_2d(smem_b + stage * B_TILE, &tmap_b, &full_bar[stage], tap * p.Cin + cb * BKE, wrow);
                if (++stage == STAGES) {
                  stage = 0;
                  phase ^= 1;
                }
                continue;
              }
              mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                uint8_t* a_dst = smem_a + stage * A_STAGE + mt * A_TILE_BYTES;
                if (p.stride == 1)
                  tma_load_4d(a_dst, &tmap_a, &full_bar[stage], a_cb * BKE, rw + pm_ow, oh0[mt] + rh, img0[mt]);
                else
                  tma_load_5d(a_dst, &tmap_a, &full_bar[stage], wp * p.Cin + cb * BKE, dw + pm_ow, hp, oh0[mt] + dh, img0[mt]);
              }
              if constexpr (MC2)
                tma_load_2d_mc(smem_b + stage * B_TILE + cta_rank * (B_TILE / 2), &tmap_b, &full_bar[stage],
                               tap * p.Cin + cb * BKE, wrow + (int)cta_rank * (BN / 2), (uint16_t)3);
              else
                tma_load_2d(smem_b + stage * B_TILE, &tmap_b, &full_bar[stage], tap * p.Cin + cb * BKE, wrow);
            }
            if (++stage == n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        // fused shortcut: the 1x1 stride-2 convolution of the block input rides in the same accumulator
        for (int cb = 0; cb < p.cblocks2; ++cb) {
          mbar_wait_ns(&empty_bar[stage], phase ^ 1, p.wait_ns_prod);
          if constexpr (CG2) {
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              if (p.sc_dense)
                tma_load_4d_2sm(smem_a + stage * A_STAGE + mt * A_TILE_BYTES, &tmap_a2, &full_bar[stage], cb * BKE, pm_ow, oh0[mt],
                                img0[mt]);
              else
                tma_load_5d_2sm(smem_a + stage * A_STAGE + mt * A_TILE_BYTES, &tmap_a2, &full_bar[stage], cb * BKE, pm_ow, 0,
                                oh0[mt], img0[mt]);
            }
            tma_load_2d_2sm(smem_b + stage * B_TILE, &tmap_b, &full_bar[stage], p.taps * p.Cin + cb * BKE,
                            wrow + (int)cta_rank * (BN / 2));
          } else {
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              if (p.sc_dense)
                tma_load_4d(smem_a + stage * A_STAGE + mt * A_TILE_BYTES, &tmap_a2, &full_bar[stage], cb * BKE, pm_ow, oh0[mt],
                            img0[mt]);
              else
                tma_load_5d(smem_a + stage * A_STAGE + mt * A_TILE_BYTES, &tmap_a2, &full_bar[stage], cb * BKE, pm_ow, 0, oh0[mt],
                            img0[mt]);
            }
            if constexpr (MC2)
              tma_load_2d_mc(smem_b + stage * B_TILE + cta_rank * (B_TILE / 2), &tmap_b, &full_bar[stage],
                             p.taps * p.Cin + cb * BKE, wrow + (int)cta_rank * (BN / 2), (uint16_t)3);
            else
              tma_load_2d(smem_b + stage * B_TILE, &tmap_b, &full_bar[stage], p.taps * p.Cin + cb * BKE, wrow);
          }
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (!(CG2 && cta_rank != 0)) {                      // CG2: only the leader CTA issues MMAs
      constexpr uint32_t idesc = SCG2 ? make_idesc_m<T>(2 * BM, MT * BM)
                                      : (CG2 ? make_idesc_m<T>(2 * BM, BN) : make_idesc<T>(SWAP ? MT * BM : BN));
      const bool leader = elect_one();                  // all 32 lanes run the loop (uniform state), one lane issues
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (RW) mbar_wait(w_bar, 0);
      for (int tile = sched_id; tile < p.num_tiles; tile += sched_n) {
        const int n_tile_mma = (tile % p.n_tiles_n) * (SCG2 ? 2 : 1);      // SCG2: first group of the pair
        const bool center_mma = (p.center_mask >> ((n_tile_mma * BN) / p.cout_g)) & 1u;
        int tile_kb = (center_mma ? p.cblocks : num_kb) + p.cblocks2;
        if (p.pm_nb2 > 0 && !center_mma && p.taps == 9) {
          // position-major: only the taps that are in bounds at this tile's position were loaded
          const int m_tile_mma = PAIRED ? 2 * (tile / p.n_tiles_n) : tile / p.n_tiles_n;
          int pos, n0_unused;
          pm_decode(m_tile_mma, p.OHW, pos, n0_unused);
          const int oh = pos / p.OW, ow = pos - oh * p.OW;
          int vh = 0, vw = 0;
          for (int k = 0; k < 3; ++k) {
            const int ih = p.stride * oh + k - p.pad, iw = p.stride * ow + k - p.pad;
            vh += (ih >= 0 && ih < p.in_h) ? 1 : 0;
            vw += (iw >= 0 && iw < p.in_w) ? 1 : 0;
          }
          tile_kb = vh * vw * p.cblocks + p.cblocks2;
        }
        mbar_wait_ns(&tmem_empty[acc], acc_phase ^ 1, p.wait_ns_acc);       // epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        int cb = 0;
        for (int kb = 0; kb < tile_kb; ++kb) {
          // gathered K: the last channel block of a tap holds fewer than 64 valid channels
          const int ksteps = (++cb == p.cblocks) ? p.last_ksteps : BK / UMMA_K;
          if (cb == p.cblocks) cb = 0;
          mbar_wait_ns(&full_bar[stage], phase, p.wait_ns_full);   // TMA bytes have landed
          if constexpr (MASKED) {                         // ... and both pixel halves are masked
            if (p.exp_flags & 16) mbar_wait(&ready_bar[stage], phase); else mbar_wait_acquire_cluster(&ready_bar[stage], phase);
          }
          tc_fence_after();
          const uint64_t w_desc = make_smem_desc(smem_u32(smem_b + (RW ? kb : stage) * B_TILE));
          if (leader) {
            if (SCG2) {
              // D^T[256 ch of the group pair, 256 px] += W[256 ch, 64] * P[256 px, 64]^T over the CTA pair: each CTA holds
              // its 128 weight rows (A) and its 128 pixels (half of B)
              const uint64_t p_desc = make_smem_desc(smem_u32(smem_a + stage * A_STAGE));
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                if (k < ksteps)
                  umma_f16_2sm(d_tmem, w_desc + (uint64_t)(2 * k), p_desc + (uint64_t)(2 * k), idesc,
                               (kb | k) != 0 ? 1u : 0u);
            } else if (CG2) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint64_t a_desc = make_smem_desc(smem_u32(smem_a + stage * A_STAGE + mt * A_TILE_BYTES));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                  if (k < ksteps)
                    umma_f16_2sm<I8>(d_tmem + (uint32_t)(mt * BN), a_desc + (uint64_t)(2 * k), w_desc + (uint64_t)(2 * k),
                                     idesc, (kb | k) != 0 ? 1u : 0u);
              }
            } else if (SWAP) {
              // D^T[128 ch, 256 px] += W[128 ch, 64] * P[256 px, 64]^T : the MT pixel tiles are contiguous in smem
              const uint64_t p_desc = make_smem_desc(smem_u32(smem_a + stage * A_STAGE));
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                if (k < ksteps)
                  umma_f16(d_tmem, w_desc + (uint64_t)(2 * k), p_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint64_t a_desc = make_smem_desc(smem_u32(smem_a + stage * A_STAGE + mt * A_TILE_BYTES));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  // advance 16 elements = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                  if (k < ksteps)
                    umma_f16<I8>(d_tmem + (uint32_t)(mt * BN), a_desc + (uint64_t)(2 * k), w_desc + (uint64_t)(2 * k), idesc,
                                 (kb | k) != 0 ? 1u : 0u);
                }
              }
            }
            if constexpr (CG2)
              umma_commit_2sm_mc(&empty_bar[stage], (uint16_t)3);   // both CTAs' producers may refill their halves
            else if constexpr (MC2)
              umma_commit_mc(&empty_bar[stage], (uint16_t)3);   // frees the slot in BOTH CTAs (the peer writes into it)
            else
              umma_commit(&empty_bar[stage]);               // frees the smem slot when these MMAs retire
          }
          __syncwarp();
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader) {
          if constexpr (CG2)
            umma_commit_2sm_mc(&tmem_full[acc], (uint16_t)3);   // accumulators complete in both CTAs -> both epilogues
          else
            umma_commit(&tmem_full[acc]);                   // accumulators complete -> epilogue
        }
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // warp w may touch TMEM lanes 32*(w%4)..+31; two warps share a lane group and split the columns.
    constexpr int NHFN = EW / 4;                          // warps per TMEM lane group (non-swapped layout): column shares
    constexpr int NCH = BN / (32 * NHFN);                 // 32-column chunks per warp
    static_assert(SWAP || (NCH >= 1 && NCH * 32 * NHFN == BN), "epilogue warps must tile the columns in 32-column chunks");
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;                       // which half of the columns
    const T* __restrict__ res = reinterpret_cast<const T*>(p.res);
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (SWAP) {
      // ---- transposed accumulator: this thread owns output channel c, TMEM columns are pixels ----
      // Host guarantees Cout == BN (one channel tile) so every row stride is the compile-time constant BN,
      // which keeps this fully unrolled loop small (the first version, with run-time strides and per-pixel
      // 64-bit address arithmetic, was 100 KB of code and instruction-cache bound).
      constexpr int PX = MT * BM;                         // pixels per tile (256)
      constexpr int NHF = EW / 4;             // warps per TMEM lane group: each takes PX / NHF pixels
      constexpr int PCH = PX / (32 * NHF);                // 32-pixel chunks per warp
      const uint16_t* __restrict__ res16 = reinterpret_cast<const uint16_t*>(p.res);
      const int64_t sample_px = (int64_t)p.dp.batch * p.OHW;          // pixels per MC sample
      const int c = q * 32 + lane;                                    // channel inside the group (cout_g == BN)
      const bool has_res = res16 != nullptr;
      for (int tile = sched_id; tile < p.num_tiles; tile += sched_n) {
        const int m_unit = tile / p.n_tiles_n;
        // output group: the channel tile itself, or (sibling pair) this CTA's member of the group pair
        const int n_tile = SCG2 ? 2 * (tile - m_unit * p.n_tiles_n) + (int)cta_rank : tile - m_unit * p.n_tiles_n;
        const int m_tile = (PAIRED && !SCG2) ? 2 * m_unit + (int)cta_rank : m_unit;
        uint16_t* __restrict__ y16 = reinterpret_cast<uint16_t*>(p.yg[n_tile]);
        const float bias_c = __ldg(p.bias + n_tile * BN + c);
        const bool relu = (p.relu_mask >> n_tile) & 1u;
        const int m_base = m_tile * PX + hf * (PX / NHF);
        if (has_res) {
          // pull this warp's residual rows (32 channels = 64 bytes per pixel) into L2 while the MMAs of this tile
          // are still running: lane l touches pixels l, l+32, l+64, l+96 of the warp's 128-pixel half
#pragma unroll
          for (int i = 0; i < PCH; ++i) {
            const int m = m_base + i * 32 + lane;
            if (m < p.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(res16 + (size_t)m * BN + q * 32));
          }
        }
        mbar_wait_ns(&tmem_full[acc], acc_phase, p.wait_ns_epi);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + hf * (PX / NHF));
#pragma unroll 1
        for (int ch = 0; ch < PCH; ++ch) {
          const int m0 = m_base + ch * 32;
          const int nvalid = p.M - m0;                    // pixels of this chunk that exist (may be <= 0)
          uint32_t v[32];
          tmem_ld_32x32(t_row + (uint32_t)(ch * 32), v);
          const size_t off0 = (size_t)m0 * BN + c;
          // stochastic site: one Philox block covers 8 consecutive channels of one pixel. Lane (c & 7) draws the
          // blocks of pixels j = (c & 7) + 8t (t < 4) for its channel octet and packs the 8 keep bits of each into
          // one word; 8 shuffles hand every lane the bits of its own channel
          uint32_t kw[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu,
                            0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
          float fac = 1.f;
          int cpos = 0;                                   // Masksembles gathered layout: slot of channel c, -1 = dropped
          if (p.dp.kind == BNN_DROP_ELEMENT) {
            // a 32-pixel chunk lies inside one MC sample whenever sample_px % 32 == 0 (host-checked)
            const uint32_t s_local = (uint32_t)((int64_t)m0 / sample_px);
            const uint64_t px0 = (uint64_t)((int64_t)m0 - (int64_t)s_local * sample_px) + (uint64_t)(lane & 7);
            uint32_t bits = 0;
#pragma unroll 1
            for (int t = 0; t < 4; ++t) {
              const uint64_t e = (px0 + 8 * t) * (uint64_t)BN + (uint64_t)(c & ~7);
              bits |= philox_keep8(p.dp.seed, p.dp.stream_id, p.dp.sample0 + s_local, e >> 3, p.dp.thr) << (8 * t);
            }
            if (p.dp.scale == 0.f) bits = 0;
            // kw[sl] bit 8t = keep bit of this lane's channel at pixel sl + 8t (pre-shifted by the lane's slot in its
            // octet, so that the per-element test below is an AND with a compile-time mask)
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) kw[sl] = __shfl_sync(0xffffffffu, bits, (lane & ~7) | sl) >> (lane & 7);
            fac = p.dp.scale;
          } else if (p.dp.kind == BNN_DROP_MASKSEMBLES) {
            // host guarantees sample_px % 32 == 0: a 32-pixel chunk never straddles two samples
            const int64_t s_local = (int64_t)m0 / sample_px;
            const int mrow = (int)(((int64_t)p.dp.cnt0 + p.dp.sample0 + s_local) % p.dp.n_masks);
            fac = __ldg(p.dp.masks + (size_t)mrow * BN + c);
            if (p.dp.compact_pos != nullptr) cpos = (int)__ldg(p.dp.compact_pos + (size_t)mrow * BN + c);
          }
          const uint16_t* rp = res16 + off0;
          uint16_t* yp = y16 + off0;
          // all 32 residual loads are issued together (each is a 64-byte coalesced warp access), then consumed
          uint32_t rr[32];
          if (has_res) {
            if (nvalid >= 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) rr[j] = (uint32_t)__ldg(rp + j * BN);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) rr[j] = (j < nvalid) ? (uint32_t)__ldg(rp + j * BN) : 0u;
            }
          }
          tmem_ld_wait();
          if (p.dp.kind == BNN_DROP_NONE && nvalid >= 32) {
            // fast path (no stochastic site, full chunk): ncu's source page showed the general path below issue-bound
            // on per-element predicates and keep-bit selects (stall_selected + stall_math + stall_not_selected > 40 %
            // of the samples on the sibling-group launch); here an element is FADD [+ residual] + FMNMX + F2FP + STG
            const float lo = relu ? 0.f : -INFINITY;
            if (has_res) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float f = fmaxf(__uint_as_float(v[j]) + bias_c + unpack2<T>(rr[j]).x, lo);
                yp[j * BN] = (uint16_t)(pack2<T>(f, 0.f) & 0xffffu);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float f = fmaxf(__uint_as_float(v[j]) + bias_c, lo);
                yp[j * BN] = (uint16_t)(pack2<T>(f, 0.f) & 0xffffu);
              }
            }
            continue;
          }
          auto value = [&](int j) -> uint16_t {
            float f = __uint_as_float(v[j]) + bias_c;
            if (has_res) f += unpack2<T>(rr[j]).x;
            if (relu) f = fmaxf(f, 0.f);
            f = (kw[j & 7] & (1u << (8 * (j >> 3)))) ? f * fac : 0.f;
            return (uint16_t)(pack2<T>(f, 0.f) & 0xffffu);
          };
          if (nvalid >= 32 && p.dp.compact_pos == nullptr) {
            // full chunk: no per-element predicates (ncu counted 42 issued instructions per element on the fused
            // element-dropout launch, a fifth of them BSSY / BSYNC / BRA / ISETP around `j < nvalid`)
#pragma unroll
            for (int j = 0; j < 32; ++j) yp[j * BN] = value(j);
          } else if (nvalid > 0 && p.dp.compact_pos == nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const uint16_t val = value(j);
              if (j < nvalid) yp[j * BN] = val;
            }
          } else if (nvalid > 0 && cpos >= 0) {
            // kept channels only, row stride = compact_c: the kept lanes of the warp write one contiguous run
            const int kc = p.dp.compact_c;
            uint16_t* yc = y16 + (size_t)m0 * kc + cpos;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) yc[j * kc] = value(j);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG2) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    } else
    for (int tile = sched_id; tile < p.num_tiles; tile += sched_n) {
      const int m_unit = tile / p.n_tiles_n, n_tile = tile - m_unit * p.n_tiles_n;
      const int m_tile = PAIRED ? 2 * m_unit + (int)cta_rank : m_unit;
      const int cbase = n_tile * BN + hf * (BN / NHFN);   // first (concatenated) output channel of this warp
      const int grp = cbase / p.cout_g;                   // output group and channel offset inside it
      const int cgrp = cbase - grp * p.cout_g;
      // 128-channel groups paired into 256-column tiles: an odd group count leaves the last half-tile without output
      const bool dead = grp >= p.groups;
      T* __restrict__ y = reinterpret_cast<T*>(p.yg[dead ? 0 : grp]);
      const bool relu = (p.relu_mask >> grp) & 1u;

      // residual rows are fetched BEFORE waiting for the accumulator: the DRAM latency hides behind the MMAs
      uint4 rpre[MT][NCH][4];
      if (res != nullptr) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          int m = (m_tile * MT + mt) * BM + q * 32 + lane;
          if (p.pm_nb2 > 0) {                         // position-major: row r of the tile is image n0 + r at one position
            int pos, n;
            pm_decode(m_tile, p.OHW, pos, n);
            n += q * 32 + lane;
            m = n < p.n_img ? n * p.OHW + pos : p.M;
          }
          if (m < p.M) {
            const uint4* src = reinterpret_cast<const uint4*>(res + (size_t)m * p.cout_g + cgrp);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
              for (int j = 0; j < 4; ++j) rpre[mt][ch][j] = __ldg(src + ch * 4 + j);
          }
        }
      }

      mbar_wait_ns(&tmem_full[acc], acc_phase, p.wait_ns_epi);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int m = (m_tile * MT + mt) * BM + q * 32 + lane;
        if (p.pm_nb2 > 0) {
          int pos, n;
          pm_decode(m_tile, p.OHW, pos, n);
          n += q * 32 + lane;
          m = n < p.n_img ? n * p.OHW + pos : p.M;
        }
        const bool valid = m < p.M && !dead;
        // stochastic-site coordinates of this output pixel
        int s_local = 0;
        uint64_t e_base = 0;
        if (p.dp.kind != BNN_DROP_NONE && valid) {
          const int img = m / p.OHW;
          s_local = img / p.dp.batch;
          const int b = img - s_local * p.dp.batch;
          const int pix = m - img * p.OHW;
          e_base =
              p.dp.kind == BNN_DROP_CHANNEL ? (uint64_t)b * p.Cout : ((uint64_t)b * p.OHW + pix) * (uint64_t)p.Cout;
        }
        const uint32_t t_row =
            tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + mt * BN + hf * (BN / NHFN));
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + (uint32_t)(ch * 32), v);
          const int c0 = cbase + ch * 32;
          float4 bia[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bia[j] = __ldg(reinterpret_cast<const float4*>(p.bias + (dead ? 0 : c0)) + j);
          tmem_ld_wait();
          if (valid || (p.pool_hw > 0 && !dead)) {      // pooling: all lanes of the warp take part in the shuffles
            const size_t off = (size_t)m * p.cout_g + (cgrp + ch * 32);
            float f[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if constexpr (I8) {
                // exact int32 accumulator -> float (round to nearest even), x per-layer multiplier, + bias in output
                // LSBs: two separately rounded operations (no FMA contraction) so that the NumPy restatement
                // (oracle/q8.py) reproduces every bit
                f[4 * j] = __fadd_rn(__fmul_rn((float)(int)v[4 * j], p.q_mult), bia[j].x);
                f[4 * j + 1] = __fadd_rn(__fmul_rn((float)(int)v[4 * j + 1], p.q_mult), bia[j].y);
                f[4 * j + 2] = __fadd_rn(__fmul_rn((float)(int)v[4 * j + 2], p.q_mult), bia[j].z);
                f[4 * j + 3] = __fadd_rn(__fmul_rn((float)(int)v[4 * j + 3], p.q_mult), bia[j].w);
              } else {
                f[4 * j] = __uint_as_float(v[4 * j]) + bia[j].x;
                f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bia[j].y;
                f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bia[j].z;
                f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bia[j].w;
              }
            }
            if (res != nullptr) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 r = rpre[mt][ch][j];
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 a = unpack2<T>(rw[t]);
                  f[8 * j + 2 * t] += a.x;
                  f[8 * j + 2 * t + 1] += a.y;
                }
              }
            }
            if (relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.dp.kind == BNN_DROP_ELEMENT || p.dp.kind == BNN_DROP_CHANNEL) {
              const uint64_t blk0 = (e_base + (uint64_t)c0) >> 3;      // c0 % 32 == 0 and Cout % 64 == 0
              // rolled loop (keeps the kernel inside the instruction cache): 4 Philox blocks -> 32 keep bits
              uint32_t bits = 0;
#pragma unroll PHILOX_UNROLL
              for (int j = 0; j < 4; ++j)
                bits |= philox_keep8(p.dp.seed, p.dp.stream_id, p.dp.sample0 + s_local, blk0 + j, p.dp.thr) << (8 * j);
              if (p.dp.scale == 0.f) bits = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] *= ((bits >> j) & 1u) ? p.dp.scale : 0.f;
            } else if (p.dp.kind == BNN_DROP_MASKSEMBLES) {
              const int mrow = (int)(((int64_t)p.dp.cnt0 + p.dp.sample0 + s_local) % p.dp.n_masks);
              if constexpr (COMPACT) {
                if (p.dp.compact_pos != nullptr) {
                  // gathered layout: channel c0 + j goes to slot pos[j] of this pixel's compact row (-1: dropped)
                  const uint4* pp = reinterpret_cast<const uint4*>(p.dp.compact_pos + (size_t)mrow * p.Cout + c0);
                  uint16_t* yc = reinterpret_cast<uint16_t*>(y) + (size_t)m * p.dp.compact_c;
#pragma unroll
                  for (int j4 = 0; j4 < 4; ++j4) {
                    const uint4 pw4 = __ldg(pp + j4);
                    const uint32_t pw[4] = {pw4.x, pw4.y, pw4.z, pw4.w};
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                      const int pos = (int)(int16_t)((t & 1) ? (pw[t >> 1] >> 16) : (pw[t >> 1] & 0xffffu));
                      if (pos >= 0) yc[pos] = (uint16_t)(pack2<T>(f[8 * j4 + t], 0.f) & 0xffffu);
                    }
                  }
                  continue;
                }
              }
              const float* mk = p.dp.masks + (size_t)mrow * p.Cout + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mk + j));
                f[j] *= m4.x;
                f[j + 1] *= m4.y;
                f[j + 2] *= m4.z;
                f[j + 3] *= m4.w;
              }
            }
            if (p.pool_hw > 0) {
              // M is a multiple of pool_hw, so an image is entirely valid or entirely past the end
#pragma unroll 1
              for (int o = p.pool_hw >> 1; o > 0; o >>= 1)
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] += __shfl_xor_sync(0xffffffffu, f[j], o);
              if (valid && (lane & (p.pool_hw - 1)) == 0) {
                const float inv = 1.f / (float)p.pool_hw;
                T* yp = y + (size_t)(m / p.pool_hw) * p.cout_g + (cgrp + ch * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 o;
                  o.x = pack2<T>(f[j] * inv, f[j + 1] * inv);
                  o.y = pack2<T>(f[j + 2] * inv, f[j + 3] * inv);
                  o.z = pack2<T>(f[j + 4] * inv, f[j + 5] * inv);
                  o.w = pack2<T>(f[j + 6] * inv, f[j + 7] * inv);
                  *reinterpret_cast<uint4*>(yp + j) = o;
                }
              }
              continue;
            }
            if (p.out_f32) {
              float* yf = reinterpret_cast<float*>(p.yg[dead ? 0 : grp]) + off;
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(yf + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              continue;
            }
            if constexpr (I8) {
              // requantise: round half to even, clip to the unsigned 8-bit grid of quantized_relu; 32 bytes per pixel
              uint32_t w8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                uint32_t word = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                  word |= (uint32_t)min(max(__float2int_rn(f[4 * j + t]), 0), 255) << (8 * t);
                w8[j] = word;
              }
              uint8_t* y8 = reinterpret_cast<uint8_t*>(y) + off;
              *reinterpret_cast<uint4*>(y8) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
              *reinterpret_cast<uint4*>(y8 + 16) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 o;
                o.x = pack2<T>(f[j], f[j + 1]);
                o.y = pack2<T>(f[j + 2], f[j + 3]);
                o.z = pack2<T>(f[j + 4], f[j + 5]);
                o.w = pack2<T>(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(y + off + j) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG2) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIRED) cluster_sync_all();   // nobody exits while its peer may still multicast into it / signal it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

static int encode_map(CUtensorMap* out, int dtype, int rank, const void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return BNN_E_CUDA;
  }
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const CUtensorMapDataType dt = dtype == BNN_I8    ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : dtype == BNN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                    : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return BNN_E_CUDA;
  }
  return BNN_OK;
}

template <int BN, int MT, bool SWAP, int PAIR, bool COMPACT, int EW, typename T, bool VH = false, bool MASKED = false>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, Params p, cudaStream_t st,
                  const CUtensorMap* tah = nullptr) {
  static bool configured = false;
  constexpr int smem_static = VH ? smem_bytes_vh() : smem_bytes_pair(BN, MT, PAIR, SWAP);
  // resident weights: (ring stages + weight k-blocks) x 16 KB
  const int smem = p.rw_kb > 0 ? (p.rw_stages + p.rw_kb) * A_TILE_BYTES + 1024 + 256 : smem_static;
  constexpr bool MC2 = PAIR != 0;
  auto kern = conv_tc_kernel<BN, MT, SWAP, PAIR, COMPACT, EW, T, VH, MASKED>;
  const CUtensorMap& th = tah ? *tah : ta;
  if (!configured) {
    BNN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (SWAP && PAIR == 2 && !MASKED) ? MAX_SMEM_OPTIN : smem_static));
    configured = true;
  }
  const int m_tiles = p.pm_nb2 > 0 ? p.OHW * p.pm_nb2 : (p.M + MT * BM - 1) / (MT * BM);
  if (MC2) {
    // pair-tiles: two adjacent row-tiles x one channel tile, or (sibling pair) one row-tile x two output groups
    p.num_tiles = (SWAP && PAIR == 2) ? m_tiles * p.n_tiles_n : ((m_tiles + 1) / 2) * p.n_tiles_n;
    int grid = 2 * p.num_tiles < (sm_count() & ~1) ? 2 * p.num_tiles : (sm_count() & ~1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(num_threads(EW) + (VH ? 32 : 0) + (MASKED ? 128 : 0));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BNN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, ta2, th, p));
  } else {
    p.num_tiles = m_tiles * p.n_tiles_n;
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    kern<<<grid, num_threads(EW) + (VH ? 32 : 0), smem, st>>>(ta, tb, ta2, th, p);
  }
  BNN_LAUNCH_OK();
  return BNN_OK;
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace tc
}  // namespace bnn

using namespace bnn;

// shared runner: `groups` outputs of `cout_g` channels each (groups == 1: an ordinary convolution)
struct GatherSel {
  int n_masks, cnt0, batch;   // weight set of image n: (cnt0 + n / batch) % n_masks; cnt0 already includes sample0
  int x_has_samples;          // 0: x holds `batch` images shared by all samples
};

struct HeadGemm {
  int kblocks;   // 64-element k-blocks of the weight matrix (2 or 3 times head_c)
  int head_c;    // k-blocks of x_hi (= F / 64); A block of weight block kb is kb < head_c ? kb : kb - head_c
};

struct Shortcut {
  const void* x2;   // [N][H2][W2][Cin2]: the input of the residual block, H2 = 2 * OH, W2 = 2 * OW
  int H2, W2, Cin2;
  int dense;        // x2 is already the even/even parity plane [N][OH][OW][Cin2]
};

struct MaskSel {
  const void* bits;   // keep bits of every sample of x ([N][H][W][Cin / 8] bytes, bnn_boundary_bits)
  int batch;          // x holds `batch` images shared by all samples
};

static int conv_tc_run(const char* who, const void* x, const void* w, const float* bias, const void* res,
                       void* const* y, int groups, uint32_t relu_mask, uint32_t center_mask, int dtype, int N, int H, int W, int Cin,
                       int cout_g, int ksize, int stride, const bnn_drop_desc* drop, void* stream,
                       const GatherSel* gsel = nullptr, const Shortcut* sc = nullptr, bool pool = false,
                       float q_mult = 1.f, const HeadGemm* hg = nullptr, const MaskSel* msk = nullptr) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(x && w && bias && y, "%s: null pointer", who);
  BNN_REQUIRE(groups >= 1 && groups <= 4, "%s: 1..4 output groups supported, got %d", who, groups);
  for (int g = 0; g < groups; ++g) BNN_REQUIRE(y[g] != nullptr, "%s: null output %d", who, g);
  BNN_REQUIRE(dtype == BNN_F16 || dtype == BNN_BF16 || dtype == BNN_I8, "%s: dtype must be float16, bfloat16 or int8", who);
  const bool i8 = dtype == BNN_I8;
  const int bke = i8 ? 2 * tc::BK : tc::BK;                  // elements of one 128-byte k-block
  BNN_REQUIRE(!i8 || (groups == 1 && gsel == nullptr && sc == nullptr && !pool && res == nullptr && (relu_mask & 1u)),
              "%s: the 8-bit kernel takes one dense output with ReLU, no residual / shortcut / pool fusion", who);
  BNN_REQUIRE(N >= 0 && H > 0 && W > 0, "%s: bad geometry", who);
  BNN_REQUIRE(groups == 1 || (res == nullptr && (drop == nullptr || drop->kind == BNN_DROP_NONE)),
              "%s: residual / stochastic epilogues need a single output group", who);
  const int Cout = groups * cout_g;
  const int pad = ksize == 3 ? 1 : 0;
  const int OH = (H + 2 * pad - ksize) / stride + 1, OW = (W + 2 * pad - ksize) / stride + 1;
  // narrow input (the stem: 3 channels stored as 16): ONE k-block per tap of which only Cin / 16 k-steps are issued
  // (Params::last_ksteps, as for gathered K); the TMA unit zero-fills the box columns past Cin
  const bool narrow_in = !i8 && gsel == nullptr && Cin > 0 && Cin < 64 && Cin % 16 == 0 && groups == 1 && sc == nullptr &&
                         !pool && hg == nullptr && msk == nullptr && stride == 1;
  const bool ok = (ksize == 1 || ksize == 3) && (stride == 1 || stride == 2) &&
                  (gsel ? (Cin % 16 == 0 && Cin > 0) : (Cin % bke == 0 || narrow_in)) && cout_g % 64 == 0 &&
                  (stride == 1 || (H % 2 == 0 && W % 2 == 0)) && tc::pow2(OW) && tc::pow2(OH) && OW <= 128;
  if (!ok) {
    set_error("%s: unsupported geometry k=%d s=%d Cin=%d Cout=%d %dx%d (use bnn_conv2d_simt)", who, ksize, stride, Cin,
              cout_g, H, W);
    return BNN_E_UNSUPPORTED;
  }
  if (drop && drop->kind != BNN_DROP_NONE) {
    BNN_REQUIRE(drop->batch > 0 && N % drop->batch == 0, "%s: N=%d not a multiple of batch=%d", who, N, drop->batch);
    BNN_REQUIRE(drop->p >= 0.f && drop->p <= 1.f, "dropout probability has to be between 0 and 1, but got %g", drop->p);
    BNN_REQUIRE(drop->kind != BNN_DROP_MASKSEMBLES || (drop->masks && drop->n_masks > 0),
                "%s: Masksembles site without a mask table", who);
  }
  if (gsel) {
    BNN_REQUIRE(gsel->n_masks > 0 && gsel->batch > 0 && N % gsel->batch == 0, "%s: bad mask rotation (n=%d batch=%d N=%d)",
                who, gsel->n_masks, gsel->batch, N);
    BNN_REQUIRE(((int64_t)gsel->batch * OH * OW) % (2 * tc::BM) == 0,
                "%s: batch * OH * OW = %lld must be a multiple of %d (one weight set per MMA tile pair)", who,
                (long long)gsel->batch * OH * OW, 2 * tc::BM);
  }
  if (sc) {
    BNN_REQUIRE(sc->x2 != nullptr && groups == 1 && gsel == nullptr, "%s: a fused shortcut needs one dense output", who);
    BNN_REQUIRE(sc->Cin2 > 0 && sc->Cin2 % 64 == 0 && sc->H2 == 2 * OH && sc->W2 == 2 * OW,
                "%s: shortcut input %dx%dx%d does not match a 1x1 stride-2 convolution onto %dx%d", who, sc->H2, sc->W2,
                sc->Cin2, OH, OW);
  }
  if (pool) {
    BNN_REQUIRE(groups == 1 && cout_g % 256 == 0 && (drop == nullptr || drop->kind == BNN_DROP_NONE) && gsel == nullptr,
                "%s: the fused average pool needs one dense output with Cout %% 256 == 0 and no stochastic epilogue", who);
    BNN_REQUIRE(OH * OW >= 2 && OH * OW <= 32 && tc::pow2(OH * OW), "%s: pooled map %dx%d must have 2..32 pixels (a power of two)",
                who, OH, OW);
  }
  const bool compact_out = drop && drop->kind == BNN_DROP_MASKSEMBLES && drop->compact_pos != nullptr;
  if (compact_out) {
    BNN_REQUIRE(drop->compact_c > 0 && drop->compact_c % 8 == 0 && drop->compact_c <= Cout,
                "%s: compact row length %d must be a positive multiple of 8 and <= Cout", who, drop->compact_c);
    BNN_REQUIRE(cout_g == 128 || cout_g % 256 == 0, "%s: compact Masksembles output needs Cout = 128 or a multiple of 256", who);
  }
  if (N == 0) return BNN_OK;
  BNN_REQUIRE((int64_t)N * OH * OW < (int64_t)1 << 31, "%s: M overflows int32", who);

  // tile geometry: 128 consecutive output pixels = tn images x th rows x OW columns
  int tw = OW;
  int th = OH < tc::BM / tw ? OH : tc::BM / tw;
  int tn = tc::BM / (tw * th);
  // ... or, position-major (small maps, 256-channel tiles): ONE output position of 128 consecutive images per row-tile,
  // so that taps which read only zero padding are skipped for the whole tile (conv_tc_kernel, Params::pm_nb2).
  // BNN_TC_NO_PM=1 keeps pixel-major tiles; BNN_TC_PM_MIN_IMAGES lowers the batch threshold (unit tests).
  const int pm_min = getenv("BNN_TC_PM_MIN_IMAGES") ? atoi(getenv("BNN_TC_PM_MIN_IMAGES")) : 1024;
  int pm_rows = 0, pm_cols = 0;      // in-bounds (position, tap) pairs per axis
  for (int k = 0; k < 3; ++k) {
    for (int o = 0; o < OH; ++o) pm_rows += (stride * o + k - 1 >= 0 && stride * o + k - 1 < H) ? 1 : 0;
    for (int o = 0; o < OW; ++o) pm_cols += (stride * o + k - 1 >= 0 && stride * o + k - 1 < W) ? 1 : 0;
  }
  // worth it when at least 10 % of the taps go away (8x8 stride 1: 16 %, 4x4: 31 %; 16x16 -> 8x8 stride 2 saves 8 % of the
  // MMAs but its 128-image gathers cost more than that: measured 0.515 vs 0.481 ms, profiles/README.md)
  const bool pm_pays = 10 * pm_rows * pm_cols <= 9 * (9 * OH * OW);
  const bool pm = ksize == 3 && pm_pays && !narrow_in && OH * OW <= 64 && OH * OW >= 4 && cout_g % 256 == 0 && gsel == nullptr && !pool && hg == nullptr &&
                  msk == nullptr && !(drop && drop->kind == BNN_DROP_MASKSEMBLES && drop->compact_pos != nullptr) &&
                  !(sc && sc->dense) && N >= pm_min && getenv("BNN_TC_NO_PM") == nullptr;
  if (pm) {
    tw = 1;
    th = 1;
    tn = tc::BM;
  }

  CUtensorMap ta, tb, ta2;
  const cuuint64_t eb = i8 ? 1 : 2;
  const int N_in = msk ? msk->batch : ((gsel && !gsel->x_has_samples) ? gsel->batch : N);     // images held by x
  if (msk) {
    BNN_REQUIRE(msk->bits != nullptr && msk->batch > 0 && N % msk->batch == 0 && gsel == nullptr && sc == nullptr && !pool &&
                    groups >= 2 && groups % 2 == 0 && cout_g == 128 && ksize == 3 && stride == 2 && center_mask == 0 &&
                    Cin % 64 == 0 && dtype != BNN_I8 && tc::BM % OW == 0,
                "%s: in-kernel keep-bit masking needs an even number of 3x3 stride-2 sibling groups of 128 channels", who);
  }
  if (stride == 1) {
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N_in};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * eb, (cuuint64_t)W * Cin * eb, (cuuint64_t)H * W * Cin * eb};
    const cuuint32_t box[4] = {(cuuint32_t)bke, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
    if (int rc = tc::encode_map(&ta, dtype, 4, x, dims, strides, box)) return rc;
  } else {
    const cuuint64_t dims[5] = {(cuuint64_t)2 * Cin, (cuuint64_t)W / 2, 2, (cuuint64_t)H / 2, (cuuint64_t)N_in};
    const cuuint64_t strides[4] = {(cuuint64_t)2 * Cin * eb, (cuuint64_t)W * Cin * eb, (cuuint64_t)2 * W * Cin * eb,
                                   (cuuint64_t)H * W * Cin * eb};
    const cuuint32_t box[5] = {(cuuint32_t)bke, (cuuint32_t)tw, 1, (cuuint32_t)th, (cuuint32_t)tn};
    if (int rc = tc::encode_map(&ta, dtype, 5, x, dims, strides, box)) return rc;
  }
  // a channel tile never straddles two output groups: BN divides cout_g ...
  int BN = cout_g % 256 == 0 ? 256 : (cout_g % 128 == 0 ? 128 : 64);
  // ... except, as an EXPERIMENT (BNN_TC_PAIR_GROUPS=1, off by default), sibling groups of exactly 128 channels
  // PAIRED into 256-column tiles of the cta_group::2 kernel (each epilogue half-tile is one group): the activation
  // tile is loaded once per two groups and a k-block moves 32 KB per 2.1 MMAC through L2 -> SM instead of the swapped
  // kernel's 48 KB.  Measured on the sibling launch behind the first site (K = 576): 1.42 ms vs 0.77 ms - with only 9
  // k-blocks per tile the row-per-thread epilogue of the 256-column kernel (16-byte stores into 32 different rows per
  // instruction) becomes the bottleneck, so the swapped kernel stays the default for 128-channel groups.  Both groups
  // of a pair must agree on the centre-tap flag; an odd group count leaves a half-tile of zero weights (TMA
  // zero-fills the rows past Cout) that the epilogue skips.
  if (groups >= 2 && cout_g == 128 && getenv("BNN_TC_PAIR_GROUPS") && atoi(getenv("BNN_TC_PAIR_GROUPS")) == 1) {
    bool ok_pairs = true;
    for (int g = 0; g + 1 < groups; g += 2)
      ok_pairs = ok_pairs && (((center_mask >> g) & 1u) == ((center_mask >> (g + 1)) & 1u) || ksize != 3);
    if (ok_pairs) BN = 256;
  }
  // CTA pairs with weight multicast: BN = 256 kernels with at least one pair of row-tiles per SM pair
  // (BNN_TC_MC_MIN_TILES overrides the threshold so that unit tests can drive the paired kernel with small shapes)
  const int64_t mc_min = getenv("BNN_TC_MC_MIN_TILES") ? atoll(getenv("BNN_TC_MC_MIN_TILES")) : 2 * (int64_t)sm_count();
  const bool mc2_any = getenv("BNN_TC_NOMC") == nullptr && ((int64_t)N * OH * OW + tc::BM - 1) / tc::BM >= mc_min;
  const bool cg2 = mc2_any && !(getenv("BNN_TC_CG2") && atoi(getenv("BNN_TC_CG2")) == 0);
  // (the compact-store epilogue exists for the single-CTA and the cta_group::2 kernels, not for the multicast one)
  const bool mc2 = BN == 256 && mc2_any && !(compact_out && !cg2);
  // transposed (operand-swapped) kernel: groups of exactly 128 channels; channel-wise dropout is not implemented
  // in its epilogue, and its stochastic epilogue assumes a 32-pixel chunk never straddles two MC samples
  const bool swap_ok =
      !i8 && hg == nullptr && BN == 128 && cout_g == 128 && !(drop && drop->kind == BNN_DROP_CHANNEL) &&
      !(drop && drop->kind != BNN_DROP_NONE && ((int64_t)drop->batch * OH * OW) % 32 != 0) &&
      getenv("BNN_TC_NOSWAP") == nullptr;
  const bool cg2_narrow_box = !i8 && mc2_any && BN < 256 && gsel == nullptr && getenv("BNN_TC_CG2_NARROW") &&
                              atoi(getenv("BNN_TC_CG2_NARROW")) == 1 &&
                              !(getenv("BNN_TC_CG2") && atoi(getenv("BNN_TC_CG2")) == 0);
  if (sc && sc->dense) {
    const int C2 = sc->Cin2;
    const cuuint64_t dims[4] = {(cuuint64_t)C2, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C2 * eb, (cuuint64_t)OW * C2 * eb, (cuuint64_t)OH * OW * C2 * eb};
    const cuuint32_t box[4] = {(cuuint32_t)tc::BK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
    if (int rc = tc::encode_map(&ta2, dtype, 4, sc->x2, dims, strides, box)) return rc;
  } else if (sc) {
    const int C2 = sc->Cin2, H2 = sc->H2, W2 = sc->W2;
    const cuuint64_t dims[5] = {(cuuint64_t)2 * C2, (cuuint64_t)W2 / 2, 2, (cuuint64_t)H2 / 2, (cuuint64_t)N};
    const cuuint64_t strides[4] = {(cuuint64_t)2 * C2 * eb, (cuuint64_t)W2 * C2 * eb, (cuuint64_t)2 * W2 * C2 * eb,
                                   (cuuint64_t)H2 * W2 * C2 * eb};
    const cuuint32_t box[5] = {(cuuint32_t)tc::BK, (cuuint32_t)tw, 1, (cuuint32_t)th, (cuuint32_t)tn};
    if (int rc = tc::encode_map(&ta2, dtype, 5, sc->x2, dims, strides, box)) return rc;
  } else {
    ta2 = ta;
  }
  {
    const int K = hg ? hg->kblocks * tc::BK : ksize * ksize * Cin + (sc ? sc->Cin2 : 0);
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout * (cuuint64_t)(gsel ? gsel->n_masks : 1)};
    const cuuint64_t strides[1] = {(cuuint64_t)K * eb};
    const cuuint32_t box[2] = {(cuuint32_t)bke, (cuuint32_t)((mc2 || cg2_narrow_box) ? BN / 2 : BN)};
    if (int rc = tc::encode_map(&tb, dtype, 2, w, dims, strides, box)) return rc;
  }

  tc::Params p{};
  p.M = N * OH * OW;
  p.Cout = Cout;
  p.n_tiles_n = (Cout + BN - 1) / BN;
  p.num_tiles = ((p.M + tc::BM - 1) / tc::BM) * p.n_tiles_n;
  p.taps = ksize * ksize;
  p.cblocks = (Cin + bke - 1) / bke;
  p.Cin = Cin;
  p.last_ksteps = (Cin - (p.cblocks - 1) * bke) / (i8 ? 2 * tc::UMMA_K : tc::UMMA_K);
  p.q_mult = q_mult;
  p.in_h = H;
  p.in_w = W;
  p.n_img = N;
  p.pm_nb2 = pm ? (((N + tc::BM - 1) / tc::BM + 1) & ~1) : 0;
  if (hg) {                                   // exit-head GEMM: K runs over the weight blocks, A blocks wrap (see Params)
    p.cblocks = hg->kblocks;
    p.last_ksteps = tc::BK / tc::UMMA_K;
    p.head_c = hg->head_c;
    p.out_f32 = 1;
  }
  if (gsel) {
    p.wsel_n = gsel->n_masks;
    p.wsel_cnt0 = ((gsel->cnt0 % gsel->n_masks) + gsel->n_masks) % gsel->n_masks;
    p.wsel_sample_px = gsel->batch * OH * OW;
    p.a_img_mod = gsel->x_has_samples ? 0 : gsel->batch;
  }
  p.exp_flags = getenv("BNN_TC_EXP") ? atoi(getenv("BNN_TC_EXP")) : 0;
  p.l2pf = getenv("BNN_TC_L2PF") ? atoi(getenv("BNN_TC_L2PF")) : 0;
  p.wait_ns_prod = p.wait_ns_epi = p.wait_ns_acc = p.wait_ns_full = 0;      // measured neutral (tools/exp_wait.py): off
  if (const char* e = getenv("BNN_TC_WAIT_NS")) {
    unsigned a = 0, b = 0, c = 0, d = 0;
    if (sscanf(e, "%u,%u,%u,%u", &a, &b, &c, &d) == 4) {
      p.wait_ns_prod = a;
      p.wait_ns_epi = b;
      p.wait_ns_acc = c;
      p.wait_ns_full = d;
    }
  }
  p.cblocks2 = sc ? sc->Cin2 / tc::BK : 0;
  p.sc_dense = sc ? sc->dense : 0;
  if (msk) {
    p.mask_bits = reinterpret_cast<const uint8_t*>(msk->bits);
    p.a_img_mod = msk->batch;
    p.in_h = H;
    p.in_w = W;
  }
  p.pool_hw = pool ? OH * OW : 0;
  p.stride = stride;
  p.pad = pad;
  p.OH = OH;
  p.OW = OW;
  p.OHW = OH * OW;
  p.groups = groups;
  p.cout_g = cout_g;
  p.relu_mask = relu_mask;
  p.center_mask = ksize == 3 ? center_mask : 0u;
  // 3x3 pad-1 convolution on 1x1 maps (the last VGG block at 32x32 inputs): eight of the nine taps only ever see zero
  // padding - run the centre tap alone (same result bit for bit, 9x fewer k-blocks)
  if (ksize == 3 && H == 1 && W == 1 && stride == 1 && getenv("BNN_TC_NO_TAP_SKIP") == nullptr)
    p.center_mask = groups >= 32 ? 0xffffffffu : ((1u << groups) - 1u);
  p.bias = bias;
  p.res = res;
  for (int g = 0; g < groups; ++g) p.yg[g] = y[g];
  p.dp = make_drop_params(drop, Cout);
  cudaStream_t st = (cudaStream_t)stream;

  // row-tiles per CTA tile: two 128-row accumulators share every weight k-block when TMEM allows (BN <= 128)
#define BNN_TC_DISPATCH_E(BN_, MT_, SWAP_, PAIR_, COMPACT_, EW_)                                          \
  case BN_:                                                                                              \
    return dtype == BNN_F16 ? tc::launch<BN_, MT_, SWAP_, PAIR_, COMPACT_, EW_, __half>(ta, tb, ta2, p, st)   \
                            : tc::launch<BN_, MT_, SWAP_, PAIR_, COMPACT_, EW_, __nv_bfloat16>(ta, tb, ta2, p, st);
#define BNN_TC_DISPATCH_C(BN_, MT_, SWAP_, PAIR_, COMPACT_) BNN_TC_DISPATCH_E(BN_, MT_, SWAP_, PAIR_, COMPACT_, 8)
#define BNN_TC_DISPATCH(BN_, MT_, SWAP_, PAIR_) BNN_TC_DISPATCH_C(BN_, MT_, SWAP_, PAIR_, false)
  if (i8) {
    // unsigned 8-bit activations x signed 8-bit weights -> int32 (tcgen05.mma kind::i8), requantising epilogue
    if (cg2 && mc2) return tc::launch<256, 1, false, 2, false, 8, int8_t>(ta, tb, ta2, p, st);
    switch (BN) {
      case 256: return tc::launch<256, 1, false, 0, false, 8, int8_t>(ta, tb, ta2, p, st);
      case 128: return tc::launch<128, 2, false, 0, false, 8, int8_t>(ta, tb, ta2, p, st);
      default: return tc::launch<64, 2, false, 0, false, 8, int8_t>(ta, tb, ta2, p, st);
    }
  }
  const bool cg2_narrow = cg2_narrow_box;   // experiment
  if (compact_out && BN == 256) {
    // compact Masksembles stores: dedicated instantiations so that the other kernels keep their code size
    if (cg2 && mc2) {
      switch (BN) { BNN_TC_DISPATCH_C(256, 1, false, 2, true) }
    }
    switch (BN) { BNN_TC_DISPATCH_C(256, 1, false, 0, true) }
  }
  if (compact_out && !(swap_ok && getenv("BNN_TC_NOSWAP") == nullptr)) {
    set_error("%s: compact Masksembles output is not available for this geometry", who);
    return BNN_E_UNSUPPORTED;
  }
  if (cg2 && cg2_narrow && BN < 256) {
    switch (BN) {
      BNN_TC_DISPATCH(128, 2, false, 2)
      BNN_TC_DISPATCH(64, 2, false, 2)
    }
  }
  if (swap_ok) {
    const bool wide_epi = drop && drop->kind == BNN_DROP_ELEMENT &&
                          !(getenv("BNN_TC_SWAP_EPI") && atoi(getenv("BNN_TC_SWAP_EPI")) == 8);
    // vertical-halo form: 3x3 stride-1 convolutions on 16 x 16 maps (one image per 256-pixel tile) - layer2 of the
    // ResNet, block 1 of the VGG.  BNN_TC_NO_VH=1 keeps the one-box-per-tap kernel (A/B measurements, unit tests).
    const bool vh = ksize == 3 && stride == 1 && OH == 16 && OW == 16 && groups == 1 && gsel == nullptr && Cin % 64 == 0 &&
                    !compact_out && getenv("BNN_TC_NO_VH") == nullptr;
    if (vh) {
      p.vh_a = 3;
      p.vh_w = 7;
      if (const char* e = getenv("BNN_TC_VH_SLOTS")) {      // "a,w" (tuning aid)
        int a = 0, w_ = 0;
        if (sscanf(e, "%d,%d", &a, &w_) == 2 && a >= 1 && w_ >= 1 && a + w_ <= tc::VH_MAX_SLOTS &&
            a * tc::VH_A_BYTES + w_ * tc::b_tile_bytes(128) <= tc::VH_RING_BYTES) {
          p.vh_a = a;
          p.vh_w = w_;
        }
      }
      CUtensorMap tah;
      const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N_in};
      const cuuint64_t strides[3] = {(cuuint64_t)Cin * eb, (cuuint64_t)W * Cin * eb, (cuuint64_t)H * W * Cin * eb};
      const cuuint32_t box[4] = {(cuuint32_t)tc::BK, 16u, 18u, 1u};
      if (int rc = tc::encode_map(&tah, dtype, 4, x, dims, strides, box)) return rc;
      if (wide_epi)
        return dtype == BNN_F16 ? tc::launch<128, 2, true, 0, false, 16, __half, true>(ta, tb, ta2, p, st, &tah)
                                : tc::launch<128, 2, true, 0, false, 16, __nv_bfloat16, true>(ta, tb, ta2, p, st, &tah);
      return dtype == BNN_F16 ? tc::launch<128, 2, true, 0, false, 8, __half, true>(ta, tb, ta2, p, st, &tah)
                              : tc::launch<128, 2, true, 0, false, 8, __nv_bfloat16, true>(ta, tb, ta2, p, st, &tah);
    }
    // sibling pair: an even number of 128-channel output groups reading the same pixels (the stride-2 siblings behind a
    // stochastic site) run as CTA pairs with tcgen05.mma.cta_group::2 - one group per CTA, the pixel tile split over the
    // pair.  Both groups of a pair must agree on the centre-tap flag.  BNN_TC_NO_SCG2=1: one CTA per (tile, group).
    bool scg2 = groups >= 2 && groups % 2 == 0 && gsel == nullptr && getenv("BNN_TC_NO_SCG2") == nullptr;
    for (int g = 0; scg2 && g + 1 < groups; g += 2)
      scg2 = ((p.center_mask >> g) & 1u) == ((p.center_mask >> (g + 1)) & 1u);
    if (msk) {
      BNN_REQUIRE(scg2, "%s: in-kernel keep-bit masking needs the sibling-pair kernel", who);
      p.n_tiles_n = groups / 2;
      return dtype == BNN_F16 ? tc::launch<128, 2, true, 2, false, 8, __half, false, true>(ta, tb, ta2, p, st)
                              : tc::launch<128, 2, true, 2, false, 8, __nv_bfloat16, false, true>(ta, tb, ta2, p, st);
    }
    if (scg2) {
      p.n_tiles_n = groups / 2;
      // one group pair whose weights fit beside a >= 4-deep pixel ring: keep them resident (Params::rw_kb)
      const int kb_all = p.taps * p.cblocks;
      const int room = (tc::MAX_SMEM_OPTIN - 1024 - 256) / tc::A_TILE_BYTES - kb_all;
      if (groups == 2 && p.center_mask == 0 && sc == nullptr && BN == 128 && room >= 4 && p.last_ksteps == tc::BK / tc::UMMA_K &&
          getenv("BNN_TC_NO_RW") == nullptr) {
        p.rw_kb = kb_all;
        p.rw_stages = room < 6 ? room : 6;
      }
      switch (BN) { BNN_TC_DISPATCH(128, 2, true, 2) }
    }
    if (wide_epi) {
      switch (BN) { BNN_TC_DISPATCH_E(128, 2, true, 0, false, 16) }
    }
    switch (BN) { BNN_TC_DISPATCH(128, 2, true, 0) }
  }
  if (cg2 && mc2) {
    // (sixteen epilogue warps for the residual + stochastic-site form of this kernel: measured 3 % SLOWER than eight,
    // tools/exp_epi256.py - the instantiation was dropped again)
    switch (BN) { BNN_TC_DISPATCH(256, 1, false, 2) }
  }
  if (mc2) {
    switch (BN) { BNN_TC_DISPATCH(256, 1, false, 1) }
  }
  switch (BN) {
    BNN_TC_DISPATCH(256, 1, false, 0)
    BNN_TC_DISPATCH(128, 2, false, 0)
    BNN_TC_DISPATCH(64, 2, false, 0)
  }
#undef BNN_TC_DISPATCH
#undef BNN_TC_DISPATCH_C
#undef BNN_TC_DISPATCH_E
  return BNN_E_UNSUPPORTED;
}

extern "C" int bnn_conv2d_tc(const void* x, const void* w, const float* bias, const void* res, void* y, int dtype,
                             int N, int H, int W, int Cin, int Cout, int ksize, int stride, int relu,
                             const bnn_drop_desc* drop, void* stream) {
  void* ys[1] = {y};
  return conv_tc_run("bnn_conv2d_tc", x, w, bias, res, ys, 1, relu ? 1u : 0u, 0u, dtype, N, H, W, Cin, Cout, ksize, stride,
                     drop, stream);
}

extern "C" int bnn_conv2d_tc_grouped(const void* x, const void* w, const float* bias, void* const* y, int n_groups,
                                     uint32_t relu_mask, uint32_t center_mask, int dtype, int N, int H, int W, int Cin,
                                     int cout_per_group, int ksize, int stride, void* stream) {
  return conv_tc_run("bnn_conv2d_tc_grouped", x, w, bias, nullptr, y, n_groups, relu_mask, center_mask, dtype, N, H, W, Cin,
                     cout_per_group, ksize, stride, nullptr, stream);
}

extern "C" int bnn_conv2d_tc_gathered(const void* x, const void* w, const float* bias, void* const* y, int n_groups,
                                      uint32_t relu_mask, uint32_t center_mask, int dtype, int N, int H, int W, int Kc,
                                      int cout_per_group, int ksize, int stride, int n_masks, int cnt0,
                                      uint32_t sample0, int batch, int x_has_samples, void* stream) {
  BNN_REQUIRE(n_masks > 0, "bnn_conv2d_tc_gathered: n_masks must be positive");
  GatherSel g{n_masks, (int)(((int64_t)cnt0 + (int64_t)sample0) % n_masks), batch, x_has_samples};
  return conv_tc_run("bnn_conv2d_tc_gathered", x, w, bias, nullptr, y, n_groups, relu_mask, center_mask, dtype, N, H, W, Kc,
                     cout_per_group, ksize, stride, nullptr, stream, &g);
}

extern "C" int bnn_conv2d_tc_shortcut(const void* x, const void* w, const float* bias, const void* res, void* y, int dtype,
                                      int N, int H, int W, int Cin, int Cout, int ksize, int stride, int relu,
                                      const bnn_drop_desc* drop, const void* x2, int H2, int W2, int Cin2, void* stream) {
  void* ys[1] = {y};
  Shortcut sc{x2, H2, W2, Cin2, 0};
  return conv_tc_run("bnn_conv2d_tc_shortcut", x, w, bias, res, ys, 1, relu ? 1u : 0u, 0u, dtype, N, H, W, Cin, Cout, ksize,
                     stride, drop, stream, nullptr, &sc);
}

extern "C" int bnn_conv2d_tc_pooled(const void* x, const void* w, const float* bias, const void* res, void* y_pooled,
                                    int dtype, int N, int H, int W, int Cin, int Cout, int ksize, int stride, int relu,
                                    void* stream) {
  void* ys[1] = {y_pooled};
  return conv_tc_run("bnn_conv2d_tc_pooled", x, w, bias, res, ys, 1, relu ? 1u : 0u, 0u, dtype, N, H, W, Cin, Cout, ksize,
                     stride, nullptr, stream, nullptr, nullptr, true);
}

extern "C" int bnn_conv2d_tc_grouped_masked(const void* x_scaled, const void* mask_bits, const void* w, const float* bias,
                                            void* const* y, int n_groups, uint32_t relu_mask, int dtype, int N, int batch, int H,
                                            int W, int Cin, int cout_per_group, void* stream) {
  MaskSel msk{mask_bits, batch};
  return conv_tc_run("bnn_conv2d_tc_grouped_masked", x_scaled, w, bias, nullptr, y, n_groups, relu_mask, 0u, dtype, N, H, W, Cin,
                     cout_per_group, 3, 2, nullptr, stream, nullptr, nullptr, false, 1.f, nullptr, &msk);
}

extern "C" int bnn_conv2d_tc_shortcut_plane(const void* x, const void* w, const float* bias, const void* res, void* y, int dtype,
                                            int N, int H, int W, int Cin, int Cout, int ksize, int stride, int relu,
                                            const bnn_drop_desc* drop, const void* x2_plane, int Cin2, void* stream) {
  void* ys[1] = {y};
  Shortcut sc{x2_plane, 2 * H, 2 * W, Cin2, 1};
  return conv_tc_run("bnn_conv2d_tc_shortcut_plane", x, w, bias, res, ys, 1, relu ? 1u : 0u, 0u, dtype, N, H, W, Cin, Cout, ksize,
                     stride, drop, stream, nullptr, &sc);
}

extern "C" int bnn_conv2d_tc_i8(const void* x, const void* w, const float* bias_q, void* y, int N, int H, int W, int Cin,
                                int Cout, int ksize, int stride, float q_mult, const bnn_drop_desc* drop, void* stream) {
  void* ys[1] = {y};
  return conv_tc_run("bnn_conv2d_tc_i8", x, w, bias_q, nullptr, ys, 1, 1u, 0u, BNN_I8, N, H, W, Cin, Cout, ksize, stride, drop,
                     stream, nullptr, nullptr, false, q_mult);
}

// logits[M][Cpad] (fp32) = [x_hi | x_lo][M][Fa] * W3[Cpad][kblocks * 64]^T + bias - the classifier of an exit head as a
// tcgen05 GEMM with hi/lo-split 16-bit operands (see kernels_head.cu: bnn_exit_head_tc)
int bnn::conv_tc_head_gemm(const void* a, const void* w3, const float* bias_pad, float* logits, int dtype, int M, int Fa,
                           int Cpad, int kblocks, int head_c, void* stream) {
  void* ys[1] = {logits};
  HeadGemm hg{kblocks, head_c};
  return conv_tc_run("bnn_exit_head_tc", a, w3, bias_pad, nullptr, ys, 1, 0u, 0u, dtype, M, 1, 1, Fa, Cpad, 1, 1, nullptr, stream,
                     nullptr, nullptr, false, 1.f, &hg);
}
