"""8-bit fixed-point restatement (test infrastructure, see oracle/__init__.py).

The reference's hardware models are QKeras networks (Hardware_Artifact/bayes_hw/models/t_qmodels_bayes_me.py:49-52,
:59-71, :99-117): every QConv2D / QDense carries ``kernel_quantizer = bias_quantizer = quantized_bits(tbit, ibit,
alpha=1)`` and is followed by ``QActivation(quantized_relu(tbit))``; the scripts run ``--quant_tbit 8``.  QKeras is not
installed here (and is unpinned in Hardware_Artifact/requirements.txt:8), so the two quantizers are restated from their
published definitions (qkeras/quantizers.py) - parity of this path is UNPINNED against the reference:

  quantized_bits(bits, integer, keep_negative=True, alpha=1):   m = 2^(bits-1), m_i = 2^integer
      q(x) = clip(round(x / m_i * m), -m, m - 1) * m_i / m          -> signed `bits`-bit integers x step 2^(integer-bits+1)
  quantized_relu(bits, integer=0):                               m = 2^bits, m_i = 2^integer
      q(x) = m_i * clip(round(x * m / m_i) / m, 0, 1 - 1/m)         -> unsigned `bits`-bit integers x step 2^(integer-bits)
  round = tf.round = round half to even.

Everything downstream is integer arithmetic: a layer is  y_q = clip(rint(acc * q_mult + bias_q), 0, 2^bits - 1)  with
acc = sum w_q * x_q exact in int32 - which is what the tcgen05 kind::i8 kernel (bnn_conv2d_tc_i8) computes bit for bit.
"""
import numpy as np


def quantized_bits_int(x, bits=8, integer=0):
    """-> (signed integers in [-2^(bits-1), 2^(bits-1) - 1], real value of one LSB)."""
    m = 2.0 ** (bits - 1)
    step = 2.0 ** (integer - bits + 1)
    q = np.clip(np.rint(np.asarray(x, dtype=np.float64) / step), -m, m - 1)
    return q.astype(np.int32), step


def quantized_relu_int(x, bits=8, integer=0):
    """-> (unsigned integers in [0, 2^bits - 1], real value of one LSB)."""
    step = 2.0 ** (integer - bits)
    q = np.clip(np.rint(np.asarray(x, dtype=np.float64) / step), 0, 2.0 ** bits - 1)
    return q.astype(np.int32), step


def conv2d_int(x_q, w_q, stride, pad):
    """Exact integer convolution. x_q [N, C, H, W] (any int), w_q [O, C, k, k] -> int64 [N, O, OH, OW]."""
    x = np.asarray(x_q, dtype=np.int64)
    w = np.asarray(w_q, dtype=np.int64)
    N, C, H, W = x.shape
    O, _, k, _ = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = np.zeros((N, O, OH, OW), dtype=np.int64)
    for a in range(k):
        for b in range(k):
            patch = xp[:, :, a:a + stride * OH:stride, b:b + stride * OW:stride]          # [N, C, OH, OW]
            out += np.einsum("nchw,oc->nohw", patch, w[:, :, a, b], optimize=True)
    return out


def requantize(acc, q_mult, bias_q, keep_scale=None, bits=8):
    """The epilogue of bnn_conv2d_tc_i8 in float32, operation for operation:
    f = fl(fl(float32(acc)) * q_mult) + bias_q[c];  f = max(f, 0);  [f = fl(f * keep_scale)];  clip(rint(f), 0, 2^bits-1)."""
    a = np.asarray(acc).astype(np.int32).astype(np.float32)                 # int32 -> float32, round to nearest even
    f = (a * np.float32(q_mult)).astype(np.float32)
    f = (f + np.asarray(bias_q, dtype=np.float32).reshape(1, -1, 1, 1)).astype(np.float32)
    f = np.maximum(f, np.float32(0))
    if keep_scale is not None:
        f = (f * np.asarray(keep_scale, dtype=np.float32)).astype(np.float32)
    return np.clip(np.rint(f), 0, 2 ** bits - 1).astype(np.uint8)


def qconv_relu(x_q, w_q, bias_q, q_mult, stride, pad, keep_scale=None):
    """QConv2D + quantized_relu(8) on integers: uint8 [N, C, H, W] -> uint8 [N, O, OH, OW]."""
    return requantize(conv2d_int(x_q, w_q, stride, pad), q_mult, bias_q, keep_scale)
