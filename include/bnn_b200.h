/*
 * bnn_b200.h - C ABI of the B200-native multi-exit MC-dropout / Masksembles inference path.
 *
 * The reference (os-hxfan/BayesNN_FPGA) has no plugin / operator / FFI layer of its own: its hot
 * path is plain PyTorch (`FullAnalysis._get_output`, Software_Artifact/software/train/
 * results_analyzer.py:236-270, calling `model(b_x)` S times).  The drop-in boundary is therefore
 * the Python object model (same class names / kwargs / return structures), and
 * THIS header is what those Python classes bind underneath with ctypes.  Every entry point names
 * the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *     buffers are borrowed (the caller, i.e. torch, owns every allocation)
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *   - return value: 0 on success, negative BNN_E_* otherwise; bnn_last_error() gives the message
 *   - activations are NHWC ("channels last"), one image = H*W*C contiguous elements; a tensor that
 *     carries Monte-Carlo samples is [S_local][B][H][W][C] (sample-major)
 *   - dtype codes: 0 = float32, 1 = float16, 2 = bfloat16 (storage of activations; accumulation is
 *     always float32)
 *   - there is NO CPU fallback: every compute entry point returns BNN_E_ARCH unless the current
 *     device is sm_100
 */
#ifndef BNN_B200_H
#define BNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNN_OK 0
#define BNN_E_ARG (-1)
#define BNN_E_ARCH (-2)
#define BNN_E_CUDA (-3)
#define BNN_E_UNSUPPORTED (-4)

#define BNN_F32 0
#define BNN_F16 1
#define BNN_BF16 2
#define BNN_I8 3   /* unsigned 8-bit activations / signed 8-bit weights (bnn_conv2d_tc_i8 only) */

/* stochastic-layer kinds (which reference module a site replaces) */
#define BNN_DROP_NONE 0
#define BNN_DROP_ELEMENT 1 /* MCDropout / F.dropout: resnet18.py:207-210, vgg19.py:384-387, Dropouts.py:32-34 */
#define BNN_DROP_CHANNEL 2 /* F.dropout2d / dropout3d: converter Dropouts.py:43-45, :54-56 */
#define BNN_DROP_MASKSEMBLES 3 /* Masksembles1D/2D eval branch: utils.py:165-169, :227-231 */

int bnn_version(void);
const char* bnn_last_error(void);
/* 0 if the current CUDA device is an sm_100 part, BNN_E_ARCH / BNN_E_CUDA otherwise. */
int bnn_device_check(void);
int bnn_sm_count(void);

/* ---- mask contract (test hooks; the same device functions are inlined in the fused kernels) ----
 * Philox-4x32-10 words / keep flags for element indices [0, count) of (seed, stream, sample).
 * Replaces torch's global dropout RNG inside F.dropout (resnet18.py:210) with a counter-based
 * stream so that masks are reproducible, shard-invariant and injectable into the oracle. */
int bnn_philox_words(uint32_t* out, int64_t count, uint64_t seed, uint32_t stream_id, uint32_t sample,
                     void* stream);
int bnn_philox_keep(uint8_t* out, int64_t count, float p, uint64_t seed, uint32_t stream_id, uint32_t sample,
                    void* stream);

/* ---- layout ----
 * x[N][C][H][W] float32 (what `model(b_x)` receives, results_analyzer.py:144-146) -> y[N][H][W][C] in
 * `dtype`. */
int bnn_nchw_to_nhwc(const float* x, void* y, int dtype, int N, int C, int H, int W, void* stream);
/* Same with `pitch` >= C channels per pixel in y (channels [C, pitch) are not written: zero-fill the buffer once).
 * Lets the 3-channel stem convolution (resnet18.py:303, vgg19.py make_layers) run on the tensor cores with its input
 * channels padded to one 64-channel k-block. */
int bnn_nchw_to_nhwc_pitch(const float* x, void* y, int dtype, int N, int C, int H, int W, int pitch, void* stream);

/* ---- convolution + folded BatchNorm + residual + ReLU (+ fused dropout) ----
 * Replaces conv2d -> BatchNorm2d(eval) [-> += residual] [-> ReLU] [-> MCDropout / Masksembles2D] chains:
 * BasicBlock.forward resnet18.py:32-48, exit branches :306-308, VGG make_layers vgg19.py:121-143, LeNet
 * t_qmodels_bayes_me.py:49-52.  `w` is [Cout][KH][KW][Cin] with the BN scale folded in, `bias` the
 * folded shift (float32 [Cout]); `res` (nullable) has the output's shape.  N counts images including the
 * sample dimension.
 *
 * Fused stochastic epilogue (drop_kind != BNN_DROP_NONE), applied after the ReLU exactly like the
 * reference's Sequential(layerN, MCDropout) (resnet18.py:278): the output is [S_local][B] images,
 * image n belongs to sample sample0 + n / B and to batch element n % B.
 *
 * bnn_conv2d_simt : exact fp32 CUDA-core implicit GEMM, any geometry, w is float32 (FP32 parity path,
 *                   and the layers whose Cin is not a multiple of 64).
 * bnn_conv2d_tc   : tcgen05 / TMEM / TMA implicit GEMM; x, w, y, res in float16 or bfloat16; 3x3 pad 1
 *                   or 1x1 pad 0, stride 1 or 2, Cin % 64 == 0, Cout % 64 == 0.
 */
typedef struct bnn_drop_desc {
  int kind;             /* BNN_DROP_* */
  float p;              /* drop probability (ELEMENT / CHANNEL) */
  uint64_t seed;        /* Philox key */
  uint32_t stream_id;   /* dropout site number */
  uint32_t sample0;     /* GLOBAL index of the first local sample */
  int batch;            /* B: images per sample */
  const float* masks;   /* MASKSEMBLES: float32 [n_masks][C] on the device */
  int n_masks;
  int cnt0;             /* MASKSEMBLES: value of the module's rotating `cnt` at sample 0 */
  /* MASKSEMBLES, optional "gathered" output layout (all NULL / 0: dense output, dropped channels stored as zeros).
   * With compact_pos set the output has compact_c (= kept channels rounded up to 16) channels per pixel and holds
   * only the channels the sample's mask keeps, in increasing channel order; dropped channels are never written
   * and the consuming convolutions (bnn_conv2d_tc_gathered) never read them.  The padding slots
   * [kept, compact_c) are not written either: the caller zero-fills the buffer once.
   *   compact_pos  int16 [n_masks][C]          slot of channel c under mask row r, -1 if dropped
   *   compact_idx  int16 [n_masks][compact_c]  channel held by slot j under mask row r, -1 for padding */
  const int16_t* compact_pos;
  const int16_t* compact_idx;
  int compact_c;
  /* ELEMENT, bnn_dropout only: the site follows a Flatten of a C x H x W map (`x.view(B, -1)` then dropout): element
   * indices of the mask stream follow the reference's NCHW-flattened order b*C*HW + c*HW + pixel instead of the
   * buffer's NHWC order, so the fused run draws the masks a stand-alone call on the flattened tensor would. */
  int nchw_flat;
} bnn_drop_desc;

int bnn_conv2d_simt(const void* x, const float* w, const float* bias, const void* res, void* y, int dtype, int N,
                    int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int relu,
                    const bnn_drop_desc* drop, void* stream);

int bnn_conv2d_tc(const void* x, const void* w, const float* bias, const void* res, void* y, int dtype, int N,
                  int H, int W, int Cin, int Cout, int ksize, int stride, int relu, const bnn_drop_desc* drop,
                  void* stream);

/* Sibling convolutions that read the SAME input with the same geometry (in the multi-exit ResNet: the first conv of
 * stage N+1, its 1x1 shortcut embedded in the centre tap of a 3x3 kernel, and the first conv of exit branch N -
 * resnet18.py:306,:316 and BasicBlock :35,:43) as ONE launch: the activation tile is fetched from HBM once.
 * `w` is [n_groups * cout_per_group][k][k][Cin], `bias` [n_groups * cout_per_group], `y` a HOST array of n_groups
 * device pointers, each output [N][OH][OW][cout_per_group]; bit g of relu_mask enables the ReLU of output g; bit g
 * of center_mask says output g is a 1x1 kernel held in the centre tap (only that tap is multiplied). */
int bnn_conv2d_tc_grouped(const void* x, const void* w, const float* bias, void* const* y, int n_groups,
                          uint32_t relu_mask, uint32_t center_mask, int dtype, int N, int H, int W, int Cin,
                          int cout_per_group, int ksize, int stride, void* stream);

/* A residual block's second convolution with its 1x1 stride-2 projection shortcut fused in (BasicBlock.forward,
 * resnet18.py:41-46: `out = bn2(conv2(out)); out += downsample(x); relu`): the shortcut's input x2 [N][H2][W2][Cin2]
 * (H2 = 2*OH, W2 = 2*OW) contributes Cin2/64 extra k-blocks to the SAME accumulator, so the shortcut tensor is never
 * written or re-read as a residual.  w is [Cout][k*k*Cin + Cin2] (the 3x3 weights followed by the folded shortcut
 * weights), bias the sum of both folded shifts.  Everything else as bnn_conv2d_tc. */
int bnn_conv2d_tc_shortcut(const void* x, const void* w, const float* bias, const void* res, void* y, int dtype, int N,
                           int H, int W, int Cin, int Cout, int ksize, int stride, int relu, const bnn_drop_desc* drop,
                           const void* x2, int H2, int W2, int Cin2, void* stream);

/* The last convolution in front of an exit head with the head's global average pool fused into its epilogue
 * (`F.avg_pool2d(F.relu(out), 4)`, resnet18.py:309,:339): the OH x OW map of every image (2..32 pixels, a power of two)
 * is averaged with warp shuffles and only y_pooled [N][Cout] is written; the map itself never reaches HBM and the head
 * (bnn_exit_head with HW = 1) reads 1/(OH*OW) of the bytes.  Cout % 256 == 0, no stochastic epilogue. */
int bnn_conv2d_tc_pooled(const void* x, const void* w, const float* bias, const void* res, void* y_pooled, int dtype,
                         int N, int H, int W, int Cin, int Cout, int ksize, int stride, int relu, void* stream);

/* Masksembles "gathered" convolution (north star (3): structured channel masks become smaller GEMMs and dropped
 * channels are never read from HBM).  Same as bnn_conv2d_tc_grouped, but
 *   x  is [N][H][W][Kc]: the compact output of a Masksembles2D site (bnn_drop_desc.compact_pos), Kc % 16 == 0;
 *   w  is [n_masks][n_groups * cout_per_group][k][k][Kc]: one weight set per mask row, input channels gathered
 *      with the same kept-channel list (zeros in the padding slots);
 *   image n belongs to local sample n / batch and uses weight set (cnt0 + sample0 + n / batch) % n_masks - the
 *      rotation of Masksembles2D.forward (utils.py:165-168).
 * x_has_samples == 0: x holds only `batch` images - the deterministic prefix in front of the first Masksembles2D
 *   site - and every sample reads them through its own weight set whose dropped input channels are zero
 *   (W (m . x) == (W . m) x): the S masked copies of the prefix tensor are never materialised (Kc == C then).
 * batch * OH * OW must be a multiple of 256 (an MMA tile pair never straddles two samples). */
int bnn_conv2d_tc_gathered(const void* x, const void* w, const float* bias, void* const* y, int n_groups,
                           uint32_t relu_mask, uint32_t center_mask, int dtype, int N, int H, int W, int Kc,
                           int cout_per_group, int ksize, int stride, int n_masks, int cnt0, uint32_t sample0,
                           int batch, int x_has_samples, void* stream);

/* ---- 8-bit fixed-point convolution on the tensor cores (tcgen05.mma kind::i8) ----
 * The B200 analogue of the reference's 8-bit QKeras layers (bayes_hw/models/t_qmodels_bayes_me.py:49-52, :59-63:
 * QConv2D with kernel/bias quantizer quantized_bits(tbit = 8, ibit, alpha = 1) followed by QActivation
 * quantized_relu(8)): activations are UNSIGNED 8-bit integers x_q (real value x_q * s_x), weights SIGNED 8-bit
 * integers w_q (real w_q * s_w), the accumulation is exact in int32, and the epilogue requantises
 *     y_q = clip(rint(float(acc) * q_mult + bias_q[c]) [* dropout], 0, 255)        rint = round half to even
 * with q_mult = s_w * s_x / s_y and bias_q = bias / s_y (both float32; powers of two for QKeras formats).  The ReLU is
 * the clip at 0.  x: NHWC uint8 [N][H][W][Cin], Cin % 128 == 0; w: int8 [Cout][k][k][Cin]; y: NHWC uint8; Cout % 64 == 0;
 * k in {1, 3}, stride in {1, 2} (same geometry rules as bnn_conv2d_tc).  `drop` (may be NULL): a stochastic site fused
 * behind the ReLU, applied in float before the requantisation (Philox element / channel masks, Masksembles rows). */
int bnn_conv2d_tc_i8(const void* x, const void* w, const float* bias_q, void* y, int N, int H, int W, int Cin, int Cout,
                     int ksize, int stride, float q_mult, const bnn_drop_desc* drop, void* stream);

/* ---- prefix boundary without the S masked copies (element-wise MC dropout in front of tensor-core convolutions) ----
 * bnn_boundary_bits replaces bnn_dropout at a site whose input is the deterministic prefix (computed once per image):
 *   x_scaled [B][H][W][C]            = x * 1/(1-p), rounded to the 16-bit storage type (ONE copy)
 *   bits     [S_local][B][H][W][C/8] = the keep bits of every sample (bit i of byte j <=> channel 8j + i kept)
 *   plane_ee [S_local][B][H/2][W/2][C] (may be NULL) = the masked even/even pixels, all a fused 1x1 stride-2 projection
 *                                      shortcut reads (bnn_conv2d_tc_shortcut_plane)
 * - 67 + 268 MB instead of 1.07 GB at BASELINE config 2.  Same Philox blocks and rounding as bnn_dropout, so
 * bnn_conv2d_tc_grouped_masked (the stride-2 sibling groups behind the site: x_scaled is read through TMA and the keep
 * bits are ANDed into the tile in shared memory by helper warps before the MMAs; N = S_local * batch output images) and
 * bnn_conv2d_tc_shortcut_plane reproduce bnn_conv2d_tc_grouped / bnn_conv2d_tc_shortcut on the masked copies bit for bit. */
int bnn_boundary_bits(const void* x, void* x_scaled, void* bits, void* plane_ee, int dtype, int B, int H, int W, int C,
                      int S_local, const bnn_drop_desc* drop, void* stream);
int bnn_conv2d_tc_grouped_masked(const void* x_scaled, const void* mask_bits, const void* w, const float* bias, void* const* y,
                                 int n_groups, uint32_t relu_mask, int dtype, int N, int batch, int H, int W, int Cin,
                                 int cout_per_group, void* stream);
int bnn_conv2d_tc_shortcut_plane(const void* x, const void* w, const float* bias, const void* res, void* y, int dtype, int N,
                                 int H, int W, int Cin, int Cout, int ksize, int stride, int relu, const bnn_drop_desc* drop,
                                 const void* x2_plane, int Cin2, void* stream);

/* ---- the rest of the 8-bit path (unsigned 8-bit NHWC activations; real value = q * step) ----
 * bnn_dropout_q8: the stochastic layer that produces / masks 8-bit tensors:
 *     y_q[s][i] = clip(rint(x[..][i] * in_to_out * factor_s(i)), 0, 255)
 *   with the masks of bnn_dropout.  x is the 16-bit deterministic prefix tensor (in_dtype BNN_F16 / BNN_BF16, in_to_out =
 *   1 / step_y: the prefix is computed once per image and enters the 8-bit suffix here) or an 8-bit tensor (BNN_I8,
 *   in_to_out = step_x / step_y).
 * bnn_maxpool2d accepts dtype BNN_I8 (max commutes with the monotone quantiser).
 * bnn_exit_head_q8: bnn_exit_head on 8-bit features (real value feat_q * feat_scale); the classifier runs in fp32. */
int bnn_dropout_q8(const void* x, void* y, int in_dtype, int64_t per_image, int C, int S_local, int x_has_samples,
                   float in_to_out, const bnn_drop_desc* drop, void* stream);
int bnn_exit_head_q8(const void* feat_q, float feat_scale, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                     const float* wt, const float* bias, const bnn_drop_desc* drop, float* sum_p, float* sum_logit,
                     float* sum_plogp, float* logits_out, int accumulate, void* stream);

/* ---- stand-alone stochastic layer (prefix -> suffix broadcast) ----
 * y[s][b][...] = drop_s(x[b][...]) for s in [0, S_local) when x_has_samples == 0 (the deterministic
 * prefix is computed once per image and broadcast across the S samples), or drop_s(x[s][b][...])
 * otherwise.  `per_image` = H*W*C elements, C = channel count (innermost). Replaces MCDropout.forward
 * (resnet18.py:207-210), BayesianDropout*.forward's dropout half (Dropouts.py:32-56) and the
 * Masksembles eval branch (utils.py:165-169, :227-231). */
int bnn_dropout(const void* x, void* y, int dtype, int64_t per_image, int C, int S_local, int x_has_samples,
                const bnn_drop_desc* drop, void* stream);

/* ---- per-channel affine [+ ReLU] on an NHWC tensor of `pixels` pixels: y = relu?(x * scale[c] + bias[c]) ----
 * An eval-mode BatchNorm2d (scale = gamma / sqrt(var + eps), bias = beta - mean * scale) that cannot be folded into
 * the convolution in front of it because a stochastic layer sits in between: the reference's converter wraps every
 * Conv2d leaf (nn2bnn.py:32-45), which turns conv -> BN -> ReLU into BayesianDropout2D(conv) -> BN -> ReLU. */
int bnn_channel_affine(const void* x, void* y, const float* scale, const float* bias, int dtype, int64_t pixels, int C,
                       int relu, void* stream);

/* ---- max-pool k x k, stride k (vgg19.py:128, LeNet t_qmodels_bayes_me.py:54,:104) ---- */
int bnn_maxpool2d(const void* x, void* y, int dtype, int N, int H, int W, int C, int k, void* stream);

/* ---- exit head: global average pool -> stochastic layer -> Linear -> softmax -> accumulate over samples ----
 * Replaces, per exit, `F.avg_pool2d(F.relu(out), k)` -> view -> exitN_dropout -> exNlinear
 * (resnet18.py:309-314, vgg19.py:297-322) plus `softmax(out, dim=1)` and the per-pass buffers of
 * _get_output (results_analyzer.py:242-248): instead of S x E device->host copies the kernel keeps
 * running sums over the local samples.
 *   feat        [S_local or 1][B][HW][F]   (feat_has_samples says which)
 *   wt, bias    float32 [F][C] (the nn.Linear weight TRANSPOSED, so that a warp reads consecutive classes), [C]
 *   sum_p, sum_logit   float32 [B][C]   (+= if accumulate != 0, else overwritten)
 *   sum_plogp          float32 [B]      sum_s sum_c p log p
 *   logits_out         nullable float32 [S_local][B][C]: per-sample logits (for parity tests)
 */
int bnn_exit_head(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                  const float* wt, const float* bias, const bnn_drop_desc* drop, float* sum_p, float* sum_logit,
                  float* sum_plogp, float* logits_out, int accumulate, void* stream);

/* Same exit head with the Linear on the warp-level tensor cores (mma.sync m16n8k16; 16 samples x C classes per image is
 * far below a tcgen05 tile) for wide heads (C = 100: the FFMA form is 14 % of the C4 step).  w_hi / w_lo: the
 * nn.Linear weight [C][F] split by bnn_split16 into a 16-bit high part and a 16-bit remainder, in the feature dtype
 * (float16 / bfloat16); the pooled features are split the same way and three products are accumulated in float32, so
 * the logits stay within ~2^-20 relative of the float32 form.  F % 16 == 0. */
int bnn_exit_head_mma(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                      const void* w_hi, const void* w_lo, const float* bias, const bnn_drop_desc* drop, float* sum_p,
                      float* sum_logit, float* sum_plogp, float* logits_out, int accumulate, void* stream);
/* hi[i] = round16(w[i]), lo[i] = round16(w[i] - hi[i]) in `dtype` (1 = float16, 2 = bfloat16). */
int bnn_split16(const float* w, void* hi, void* lo, int64_t n, int dtype, void* stream);

/* bnn_exit_head with the classifier on the 5th-generation tensor cores (wide heads: C = 100 ... 1000):
 *   1. features: average pool -> stochastic site -> 16-bit high part x_hi [and remainder x_lo when with_lo != 0: pooled
 *      means, dropout scales that are not powers of two] into a_ws [S_local * B][F or 2F]
 *   2. logits_ws [S_local * B][c_pad] (fp32) = [x_hi | x_lo] x w3^T + bias_pad as ONE tcgen05 GEMM; w3 is the nn.Linear
 *      weight split into hi / lo 16-bit parts and laid out [c_pad][w_hi | w_lo (| w_hi)] (F columns each, rows >= C zero):
 *      x_hi * w_hi + x_hi * w_lo (+ x_lo * w_hi) keeps the logits within ~2^-20 of the fp32 classifier
 *   3. soft-max, running sums over the samples, per-sample logits - as bnn_exit_head.
 * a_ws / logits_ws are caller-owned workspaces. */
int bnn_exit_head_tc(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                     const void* w3, const float* bias_pad, int c_pad, int with_lo, const bnn_drop_desc* drop, void* a_ws,
                     float* logits_ws, float* sum_p, float* sum_logit, float* sum_plogp, float* logits_out, int accumulate,
                     void* stream);

/* ---- statistics finaliser ----
 * From the (all-reduced) sums over S_total samples: predictive mean, mean logits, cumulative exit ensembles
 * (results_analyzer.py:247-248, :260-269), entropy of the mean with the reference's 1e-8 epsilon
 * (bayes_hw/metric_utils.py:3-6) and the expected per-sample entropy.
 *   sums layout: sum_p [E][B][C], sum_logit [E][B][C], sum_plogp [E][B]
 *   outputs: mean_p, mean_logit, ens_p, ens_logit [E][B][C]; entropy, ens_entropy, exp_entropy [E][B]
 */
int bnn_finalize(const float* sum_p, const float* sum_logit, const float* sum_plogp, int E, int B, int C,
                 int S_total, float* mean_p, float* mean_logit, float* ens_p, float* ens_logit, float* entropy,
                 float* ens_entropy, float* exp_entropy, void* stream);

/* ---- dataset-level statistics on the device (csrc/kernels_stats.cu) ----
 * All of them take the FLOAT64 [N][C] mean-probability arrays FullAnalysis keeps on the host
 * (results_analyzer.py:133-135) and compute in float64 like the reference's NumPy code; labels are int32 [N].
 *
 * bnn_calibration_bins: per-image top-label confidence / correctness and equal-width bin statistics.
 *   mode BNN_CAL_TOP            conf = p[argmax]           bins (lo, hi]    (the corrected 10-bin ECE)
 *        BNN_CAL_TOP_NORM       conf = p[argmax] / sum(p)  bins (lo, hi]    (confidences of ece_hist_binary :446-460)
 *        BNN_CAL_TFP_RESOFTMAX  conf = softmax_f32(p)[argmax p], bins [lo, hi) = floor(n * conf): the statistic
 *                               hls4ml_pred.py:90-91,115-116 REALLY computes - it passes probabilities as `logits=`
 *                               and tfp.stats.expected_calibration_error soft-maxes them again
 * bin_stats: float64 [n_bins][3] = (count, sum confidence, sum correct), zeroed by the call. */
enum { BNN_CAL_TOP = 0, BNN_CAL_TOP_NORM = 1, BNN_CAL_TFP_RESOFTMAX = 2 };
int bnn_calibration_bins(const double* probs, const int32_t* labels, int N, int C, int n_bins, int mode, double* conf,
                         int32_t* correct, double* bin_stats, void* stream);

/* Dataset-level scores of ece_eval_binary (results_analyzer.py:497-503):
 * out[0] = NLL = -mean log clip(p[label], 1e-256, 1 - 1e-256), out[1] = MSE = mean sum_c (p_c - onehot_c)^2,
 * out[2] = top-1 accuracy.  workspace: float64 [3 * ceil(N / 256)] (fixed-order two-stage reduction: deterministic). */
int bnn_dataset_metrics(const double* probs, const int32_t* labels, int N, int C, double* workspace, double* out,
                        void* stream);

/* ---- confidence-threshold early exiting (results_analyzer.py:606-631 confidence_exiting, :728-735 is_confident) ----
 * probs [E][N][C] float64 (mean predictions of every exit).  Image i leaves at the first exit e in
 * [first_exit, E-1) with max_c p > threshold (diff == 0) or |top1 - top2| > threshold (diff != 0), else at E-1; the
 * reference starts its scan at exit 1 (`for layer in range(1, self.n_exits)`), pass first_exit = 1 to mirror it.
 * exit_idx int32 [N]; best_probs float64 [N][C] = the chosen exit's prediction; exit_hist int32 [E] = images per
 * exit (zeroed by the call) - the FLOP accounting of flop_saver / flop_saver_ensembled (:639-726) is a dot product
 * of this histogram with the per-exit cost table. */
int bnn_confidence_exit(const double* probs, int E, int N, int C, int first_exit, double threshold, int diff,
                        int32_t* exit_idx, double* best_probs, int32_t* exit_hist, void* stream);

/* ---- KDE-ECE building blocks (ece_kde_binary, results_analyzer.py:351-443) ----
 * bnn_top_label: probabilities clipped to [1e-256, 1 - 1e-256] (:357-358), then
 *     binary == 0 (top-label branch :370-380, :412-416): conf[i] = p[i][argmax] / sum_c p[i][c], flag[i] = argmax == label
 *     binary != 0 (C == 2, joint-calibration branch :381-383, :417-419): conf[i] = p[i][1] / sum_c p[i][c], flag[i] = label
 *   round_f32 != 0 rounds conf to float32 (the reference stores the top-label confidences of `p` in a float32 tensor,
 *   :373; those of the integration set p_int stay float64, :414).  labels may be NULL (integration set: no flags, no
 *   stats).  stats (double [3], zeroed by the call) = (n, sum conf, sum conf^2) over the flagged images for the
 *   bandwidth rule std(conf | flag) * (2N)^-0.2 (:389-392).
 * bnn_kde_triweight: density of `data` (only entries with flags[i] != 0 when flags is given) mirrored about lo / hi
 *   (mirror_1d :339-349), triweight kernel of standard deviation `bw`, on the grid x0 + j*dx, j < G; zero outside
 *   (lo, hi) and doubled (:403-406); n_points = number of un-mirrored points used.  The estimate is evaluated
 *   EXACTLY; the reference uses KDEpy's FFTKDE (linear binning + FFT convolution) for the same estimator. */
int bnn_top_label(const double* probs, const int32_t* labels, int N, int C, int binary, int round_f32, double* conf,
                  int32_t* flag, double* stats, void* stream);
int bnn_kde_triweight(const double* data, const int32_t* flags, int n, double bw, double n_points, double x0,
                      double dx, int G, double lo, double hi, double* out, void* stream);

/* Narrow exit heads (C <= 32) with all samples in parallel (same contract as bnn_exit_head; replaces the same reference
 * code: resnet18.py:309-314 + results_analyzer.py:242-248): one warp per (sample, image) row computes the logits
 * (pool -> site -> Linear in fp32; w_cf = the nn.Linear weight [C][F] float32) into logits_ws [S_local * B][C] float32,
 * then one CTA per image runs the soft-max and adds the samples in order (deterministic).  F % 8 == 0. */
int bnn_exit_head_rows(const void* feat, int dtype, int feat_has_samples, int B, int S_local, int HW, int F, int C,
                       const float* w_cf, const float* bias, const bnn_drop_desc* drop, float* logits_ws, float* sum_p,
                       float* sum_logit, float* sum_plogp, float* logits_out, int accumulate, void* stream);

/* ---- the path's one collective as a kernel over NVLink peer memory (sample sharding) ----
 * The reference is single-device; sharding its S passes over the GPUs of a box leaves one exchange: the sum of the
 * per-exit statistics sums (what `np.average(all_output_probs, axis=0)` / `all_outputs` run over, results_analyzer.py:
 * 247-248) before the finaliser.  bnn_peer_allreduce sums `flat` (n float32, device, 16-byte aligned) over `world`
 * ranks IN PLACE on `stream`: peer_base[r] (HOST array of `world` device pointers) is rank r's workspace as mapped into
 * THIS process (symmetric memory: same layout everywhere) = a slot of >= n floats followed, at flags_offset_bytes, by
 * 16 uint32 flags (zeroed once before the first call on any rank); `state` = 32 bytes of local device memory, zeroed
 * once.  Ranks add the slots in rank order, so every rank gets the same bits.  All ranks must call it the same number
 * of times.  A peer that does not answer within 2 s sets an error flag instead of hanging the GPU:
 * bnn_peer_allreduce_status (synchronises `stream`) returns BNN_E_CUDA then. */
int bnn_peer_allreduce(float* flat, int64_t n, void* const* peer_base, int world, int rank, int64_t flags_offset_bytes,
                       void* state, void* stream);
int bnn_peer_allreduce_status(const void* state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BNN_B200_H */
