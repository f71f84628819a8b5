"""Host restatement of the mask contract (test infrastructure, see oracle/__init__.py).

The reference draws dropout masks from torch's global RNG (``F.dropout(x, p, True)``,
``resnet18.py:207-210``); that stream is not reproducible across devices, so parity is
checked with INJECTED masks.  This file is the single source of truth for those masks:
Philox-4x32-10 (Salmon et al., SC'11), counter-based.

Contract (mirrored bit-exactly by ``bayesnn_fpga_b200/csrc/philox.cuh``):

  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (e >> 3  [low 32 bits], e >> 35, sample, stream)      one Philox block covers 8 elements
  half     = 16-bit half (e & 1) of word (e >> 1) & 3 of philox4x32_10(counter, key)   (low half first)
  keep     = (p < 1) and (half >= thr(p)),   thr(p) = min(rint(p * 2**16), 2**16 - 1)

(16-bit decisions halve the generator cost inside the GEMM epilogues; the drop probability is realised to
within 2**-17, finer than any dropout rate is ever specified.)

``e`` is the linear element index of the masked tensor for ONE sample:
  * element-wise dropout on a 4-D activation: NHWC order, e = ((b*H + h)*W + w)*C + c
  * element-wise dropout on a 2-D activation [B, F]: e = b*F + f
  * channel-wise dropout (``F.dropout2d``, converter Dropouts.py:43-45): e = b*C + c
``sample`` is the GLOBAL Monte-Carlo sample index (results do not depend on how samples
are sharded over GPUs); ``stream`` identifies the dropout site in the network.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox-4x32-10. All inputs broadcastable uint32 arrays / ints.
    Returns four uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0          # 32x32 -> 64, no overflow in uint64
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def random_words(seed, stream, sample, count, first_block=0):
    """``count`` uint32 words of one (stream, sample), starting at Philox block `first_block` (4 words per block)."""
    nblk = (count + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64) + np.uint64(first_block)
    r = philox4x32_10(blk & _MASK32, blk >> np.uint64(32), sample, stream,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(r, axis=1).reshape(-1)[:count]


def random_halves(seed, stream, sample, count, first=0):
    """``count`` 16-bit draws for element indices first..first+count-1: two per 32-bit word, low half first (one
    Philox block = 8 elements)."""
    first = int(first)
    blk0, skip = first // 8, first % 8
    w = random_words(seed, stream, sample, (skip + count + 1) // 2, first_block=blk0)
    h = np.stack([w & np.uint32(0xFFFF), w >> np.uint32(16)], axis=1).reshape(-1)
    return h[skip:skip + count].astype(np.uint32)


def threshold(p):
    """16-bit drop threshold: keep iff half >= threshold(p)."""
    return int(min(np.rint(float(p) * 65536.0), 65535.0))


def keep_mask_flat(seed, stream, sample, count, p, first=0):
    if p >= 1.0:
        return np.zeros(count, dtype=bool)
    return random_halves(seed, stream, sample, count, first) >= np.uint32(threshold(p))


def keep_mask(seed, stream, sample, shape, p, mode="element", batch_offset=0):
    """Boolean keep mask in TORCH (NCHW / [B,F]) layout for one MC sample.

    mode "element": the NHWC-ordered contract above, returned permuted to NCHW.
    mode "channel": one draw per (b, c), broadcast over the spatial dims (dropout2d).
    batch_offset: the tensor holds images batch_offset.. of a larger batch (element indices are per image WITHIN the
    batch, so rows of a full-size run can be checked without materialising the whole batch on the host).
    """
    shape = tuple(int(v) for v in shape)
    if mode == "channel":
        b, c = shape[0], shape[1]
        m = keep_mask_flat(seed, stream, sample, b * c, p, batch_offset * c).reshape(b, c)
        return np.broadcast_to(m.reshape(b, c, *([1] * (len(shape) - 2))), shape).copy()
    if len(shape) == 4:
        b, c, h, w = shape
        m = keep_mask_flat(seed, stream, sample, b * h * w * c, p, batch_offset * h * w * c).reshape(b, h, w, c)
        return np.ascontiguousarray(m.transpose(0, 3, 1, 2))
    count = int(np.prod(shape))
    return keep_mask_flat(seed, stream, sample, count, p, batch_offset * (count // max(shape[0], 1))).reshape(shape)


def uniform01(seed, stream, sample, count):
    """float64 uniforms in (0,1) - used only to build seeded synthetic weights."""
    w = random_words(seed, stream, sample, count).astype(np.float64)
    return (w + 0.5) / 4294967296.0


def normal(seed, stream, sample, count):
    """float64 standard normals by Box-Muller over two Philox streams."""
    u1 = uniform01(seed, stream, 2 * sample, count)
    u2 = uniform01(seed, stream, 2 * sample + 1, count)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
