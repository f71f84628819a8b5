"""8-bit fixed-point inference on the tensor cores - the B200 analogue of the reference's QKeras models
(``Hardware_Artifact/bayes_hw/models/t_qmodels_bayes_me.py:49-52``: ``QConv2D(kernel_quantizer=quantized_bits(tbit,
ibit, alpha=1), bias_quantizer=...)`` + ``QActivation(quantized_relu(tbit))``; scripts run ``--quant_tbit 8``).

``quantized_bits`` / ``quantized_relu`` restate QKeras' quantizers (power-of-two steps, round half to even);
:class:`QConv2d` holds an integer layer (signed 8-bit weights, bias and multiplier in output LSBs) and runs it through
``bnn_conv2d_tc_i8`` (``tcgen05.mma kind::i8``: u8 x s8 -> exact int32, requantising epilogue with an optional fused
stochastic site).  Activations are NHWC ``torch.uint8`` tensors on the device with a real value per LSB
(:class:`QTensor`).  There is no CPU path.
"""
import ctypes
from dataclasses import dataclass

import torch

from . import _lib


def quantized_bits(x, bits=8, integer=0):
    """QKeras ``quantized_bits(bits, integer, alpha=1)`` -> (int8-range integer tensor, real value of one LSB)."""
    step = 2.0 ** (integer - bits + 1)
    m = 2.0 ** (bits - 1)
    return torch.clamp(torch.round(x.double() / step), -m, m - 1).to(torch.int32), step


def quantized_relu(x, bits=8, integer=0):
    """QKeras ``quantized_relu(bits, integer)`` -> (unsigned integer tensor, real value of one LSB)."""
    step = 2.0 ** (integer - bits)
    return torch.clamp(torch.round(x.double() / step), 0, 2.0 ** bits - 1).to(torch.int32), step


def integer_bits_for(max_abs, signed):
    """Smallest `integer` (QKeras' ibit) whose range covers max_abs."""
    import math
    if max_abs <= 0:
        return 0
    return max(int(math.ceil(math.log2(max_abs + 1e-30))), -8) + (0 if signed else 0)


@dataclass
class QTensor:
    """NHWC uint8 activations on the device; real value = q * step."""
    q: torch.Tensor
    step: float

    def dequantize(self):
        return self.q.float() * self.step

    @staticmethod
    def from_float_nchw(x, integer=0, bits=8):
        q, step = quantized_relu(x, bits, integer)
        return QTensor(q.to(torch.uint8).permute(0, 2, 3, 1).contiguous().cuda(), step)


class QConv2d:
    """QConv2D (3x3 pad 1 | 1x1 pad 0, stride 1 | 2) + quantized_relu(8) as one integer layer.

    weight [O, C, k, k] and bias [O] are REAL-valued tensors; they are quantised with ``quantized_bits(8, w_integer)``
    like the reference's kernel / bias quantizers.  ``out_integer`` is the integer-bit count of the output's
    ``quantized_relu`` (0 in the reference: activations in [0, 1))."""

    def __init__(self, weight, bias, stride=1, w_integer=0, out_integer=0):
        self.k, self.stride = int(weight.shape[2]), int(stride)
        if self.k not in (1, 3) or weight.shape[2] != weight.shape[3]:
            raise ValueError("QConv2d: 1x1 or 3x3 kernels")
        self.w_q, self.w_step = quantized_bits(weight, 8, w_integer)
        self.b_q, _ = quantized_bits(bias, 8, w_integer)                     # bias_quantizer == kernel_quantizer
        self.bias_real = self.b_q.double() * self.w_step
        self.out_step = 2.0 ** (out_integer - 8)
        self.d_w = self.w_q.to(torch.int8).permute(0, 2, 3, 1).contiguous().cuda()        # [O][k][k][C]
        self.d_bias_q = (self.bias_real / self.out_step).float().cuda()      # bias in output LSBs

    def q_mult(self, in_step):
        return float(self.w_step * in_step / self.out_step)

    def __call__(self, x: QTensor, drop=None):
        lib = _lib.load()
        N, H, W, C = x.q.shape
        O = self.d_w.shape[0]
        pad = 1 if self.k == 3 else 0
        OH, OW = (H + 2 * pad - self.k) // self.stride + 1, (W + 2 * pad - self.k) // self.stride + 1
        y = torch.empty((N, OH, OW, O), dtype=torch.uint8, device=x.q.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.q.device).cuda_stream)
        _lib.check(lib.bnn_conv2d_tc_i8(x.q.data_ptr(), self.d_w.data_ptr(), self.d_bias_q.data_ptr(), y.data_ptr(), N, H, W,
                                        C, O, self.k, self.stride, self.q_mult(x.step),
                                        ctypes.byref(drop) if drop is not None else None, stream))
        return QTensor(y, self.out_step)


# --------------------------------------------------------------------------------------------------------------------
# 8-bit SUFFIX of a multi-exit network: the deterministic prefix stays on the 16-bit path (it runs once per image),
# everything behind the first stochastic site - S times the work - runs on unsigned 8-bit activations and signed 8-bit
# weights with power-of-two steps, like the reference's QKeras models (quantized_bits(8, ibit) / quantized_relu(8, ibit)).
# --------------------------------------------------------------------------------------------------------------------
def _ibits(max_abs):
    import math
    return int(math.ceil(math.log2(max(float(max_abs), 2.0 ** -20))))


class Q8Plan:
    """8-bit execution plan of `model`'s stochastic suffix.

    Calibration (once, on `calib_x`): the 16-bit engine runs S_calib samples; every stochastic tensor gets the
    power-of-two step whose unsigned 8-bit range covers its observed maximum (QKeras' `integer` bits), every suffix
    convolution gets `quantized_bits(8, ibit)` weights / biases with ibit from the BN-folded tensors' ranges.
    Supported suffix vocabulary: 3x3 / 1x1 convolutions with ReLU (Cin % 128 == 0, Cout % 64 == 0, no residual),
    max-pools, stochastic sites (stand-alone or fused behind a convolution), exit heads - e.g. the multi-exit VGG-19
    with dropout on its last blocks (BASELINE config 4).  Anything else raises NotImplementedError."""

    def __init__(self, model, calib_x, S_calib=4, seed=0x5EED, margin=1.0):
        from . import engine as _engine
        self.model = model
        dev = next(model.parameters()).device
        self.eng = _engine.Engine(model._bnn_graph(), dtype="fp16", device=dev, fuse_graph=False)
        g = self.eng.graph
        self.lib = _lib.load()
        # ---- calibration pass on the 16-bit path
        r = self.eng.run(calib_x, S_calib, seed=seed, use_graph=False)
        torch.cuda.synchronize(dev)
        st = self.eng._bufs[(calib_x.shape[0], S_calib, False)]
        self.step, self.amax = {}, {}
        for op in g.ops:
            t = op.dst
            if t is None or not t.stoch:
                continue
            amax = float(st["acts"][t.id].float().abs().max()) * margin
            self.amax[t.id] = amax
            self.step[t.id] = 2.0 ** (_ibits(amax) - 8)
        del r
        self.eng.release_buffers()
        # ---- integer layers
        self.layers = {}
        for op in g.ops:
            if op.kind == "maxpool" and op.dst.stoch:
                self.step[op.dst.id] = self.step[op.src.id]                 # max commutes with the quantiser
        for i, op in enumerate(g.ops):
            if not (op.dst is not None and op.dst.stoch):
                continue
            if op.kind == "conv":
                kh, kw = op.ksize
                ok = (op.src.stoch and op.res is None and op.relu and kh == kw and kh in (1, 3) and op.pad == (kh == 3)
                      and op.src.C % 128 == 0 and op.dst.C % 64 == 0 and getattr(op, "use_tc", False))
                if not ok:
                    raise NotImplementedError("8-bit plan: convolution %s (Cin=%d, Cout=%d, k=%d, relu=%s, residual=%s) is "
                                              "outside the 8-bit suffix vocabulary" % (op.name, op.src.C, op.dst.C, kh,
                                                                                       op.relu, op.res is not None))
                wi, bi = _ibits(op.weight.abs().max()), _ibits(max(float(op.bias.abs().max()), 2.0 ** -20))
                w_q, w_step = quantized_bits(op.weight, 8, wi)
                b_q, b_step = quantized_bits(op.bias, 8, bi)
                out_step = self.step[op.dst.id]
                bias_q = (b_q.double() * b_step / out_step).float()
                # 3x3 convolution on a 2x2 map == one dense GEMM over the image's 4*Cin values (see Engine: fc22)
                fc22 = (kh == 3 and op.stride == 1 and op.src.H == 2 and op.src.W == 2 and
                        (op.site is None or op.site.kind == "mc") and (4 * op.src.C) % 128 == 0)
                if fc22:
                    co, ci = w_q.shape[:2]
                    wf = torch.zeros((4 * co, 4 * ci), dtype=torch.int8)
                    for p_ in range(4):
                        for q_ in range(4):
                            kh_, kw_ = q_ // 2 - p_ // 2 + 1, q_ % 2 - p_ % 2 + 1
                            wf[p_ * co:(p_ + 1) * co, q_ * ci:(q_ + 1) * ci] = w_q[:, :, kh_, kw_].to(torch.int8)
                    d_w, bias_q = wf.contiguous().to(dev), bias_q.repeat(4)
                else:
                    d_w = w_q.to(torch.int8).permute(0, 2, 3, 1).contiguous().to(dev)
                self.layers[i] = dict(
                    w_q=w_q, w_step=w_step, b_q=b_q, b_step=b_step, fc22=fc22, d_w=d_w, d_bias_q=bias_q.to(dev),
                    q_mult=float(w_step * self.step[op.src.id] / out_step))
            elif op.kind not in ("site", "maxpool"):
                raise NotImplementedError("8-bit plan: op %s (%s) in the stochastic suffix" % (op.name, op.kind))
        self._qbufs = {}

    def _bufs(self, B, S):
        key = (B, S)
        if key not in self._qbufs:
            g = self.eng.graph
            self._qbufs[key] = {t.id: torch.empty((S * B, t.H, t.W, t.C), dtype=torch.uint8, device=self.eng.device)
                                for t in g.tensors if t.stoch and t.id in self.step}
        return self._qbufs[key]

    def run(self, x, S, seed=0x5EED, sample0=0, keep_activations=False):
        """-> MCResult like Engine.run (mean probabilities, logits, ensembles, entropies)."""
        from .engine import MCResult, _ptr
        eng, g, lib = self.eng, self.eng.graph, self.lib
        x = x.to(eng.device, torch.float32)
        B = x.shape[0]
        with torch.cuda.device(eng.device):
            st = eng._buffers(B, S, False)
            stream = ctypes.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
            acts, q = st["acts"], self._bufs(B, S)
            st["x"].copy_(x, non_blocking=True)
            _lib.check(lib.bnn_nchw_to_nhwc_pitch(_ptr(st["x"]), _ptr(acts[g.input.id]), eng.dcode, B, g.input.C, g.input.H,
                                                  g.input.W, eng.in_pad or g.input.C, stream))
            eng._run_ops(st, B, S, sample0, seed, False, None, stream, "det", 0)        # 16-bit prefix (+ its heads)
            sum_p, sum_l, sum_pl = eng.sums_views(st, B)
            for i, op in enumerate(g.ops):
                out_t = op.dst if op.dst is not None else op.src
                if not (True if op.kind == "site" else out_t.stoch):
                    continue
                if op.kind == "site":
                    dd = eng._drop_desc(op.site, getattr(op, "d_masks", None), B, sample0, seed, None)
                    per_image = op.src.H * op.src.W * op.src.C
                    if op.src.stoch:
                        src, code, scale = q[op.src.id], _lib.I8, self.step[op.src.id] / self.step[op.dst.id]
                    else:
                        src, code, scale = acts[op.src.id], eng.dcode, 1.0 / self.step[op.dst.id]
                    _lib.check(lib.bnn_dropout_q8(_ptr(src), _ptr(q[op.dst.id]), code, per_image, op.src.C, S,
                                                  int(op.src.stoch), scale, ctypes.byref(dd), stream))
                elif op.kind == "conv":
                    L = self.layers[i]
                    dd = eng._drop_desc(op.site, getattr(op, "d_masks", None), B, sample0, seed, None)
                    geo = (1, 1, 4 * op.src.C, 4 * op.dst.C, 1, 1) if L["fc22"] else \
                        (op.src.H, op.src.W, op.src.C, op.dst.C, op.ksize[0], op.stride)
                    _lib.check(lib.bnn_conv2d_tc_i8(_ptr(q[op.src.id]), _ptr(L["d_w"]), _ptr(L["d_bias_q"]), _ptr(q[op.dst.id]),
                                                    S * B, *geo, L["q_mult"], ctypes.byref(dd), stream))
                elif op.kind == "maxpool":
                    _lib.check(lib.bnn_maxpool2d(_ptr(q[op.src.id]), _ptr(q[op.dst.id]), _lib.I8, S * B, op.src.H, op.src.W,
                                                 op.src.C, op.pool_k, stream))
                elif op.kind == "head":
                    e = op.exit_index
                    dd = eng._drop_desc(op.site, getattr(op, "d_masks", None), B, sample0, seed, None)
                    _lib.check(lib.bnn_exit_head_q8(_ptr(q[op.src.id]), self.step[op.src.id], 1, B, S, op.src.H * op.src.W,
                                                    op.src.C, g.n_classes, _ptr(op.d_w), _ptr(op.d_b), ctypes.byref(dd),
                                                    _ptr(sum_p[e]), _ptr(sum_l[e]), _ptr(sum_pl[e]), None, 0, stream))
            views, ent = eng.finalize(st, B, S)
        eng.advance_masksembles(S)
        return MCResult(views[0], views[1], views[2], views[3], ent[0], ent[1], ent[2], None, S)
