"""Masksembles layers and mask generator - drop-in for the reference's
``Software_Artifact/software/utils.py`` (generator :18-110, ``Masksembles2D`` :115-174,
``Masksembles1D`` :177-236) with the forward pass on the B200 C-ABI library.

The generator must be BIT-EXACT with the reference: it consumes the global NumPy RNG through the
same sequence of ``np.random.choice(range(width), ones, replace=False)`` calls, so a model built
after ``np.random.seed(k)`` gets the masks the reference would get.
"""
import ctypes

import numpy as np
import torch
from torch import nn

from . import _lib


def dict_drop(dic, *keys):
    """utils.py:7-12."""
    return {k: v for k, v in dic.items() if k not in keys}


def generate_masks_(m: int, n: int, s: float) -> np.ndarray:
    """n binary rows with m ones among int(m*s) positions; unused positions dropped (utils.py:18-41)."""
    width = int(m * s)
    rows = np.zeros((n, width))
    for i in range(n):
        rows[i, np.random.choice(range(width), m, replace=False)] = 1
    return rows[:, rows.any(axis=0)]


def generate_masks(m: int, n: int, s: float) -> np.ndarray:
    """Redraw until the width equals the expected m*s*(1-(1-1/s)^n) (utils.py:44-63)."""
    target = int(m * s * (1 - (1 - 1 / s) ** n))
    masks = generate_masks_(m, n, s)
    while masks.shape[1] != target:
        masks = generate_masks_(m, n, s)
    return masks


def generation_wrapper(c: int, n: int, scale: float) -> np.ndarray:
    """[n, c] masks for a layer with c channels (utils.py:66-110); same errors as the reference."""
    if c < 10:
        raise ValueError("Masksembles approach couldn't be used in such setups where "
                         f"number of channels is less then 10. Current value is (channels={c}). "
                         "Please increase number of features in your layer or remove this "
                         "particular instance of Masksembles from your architecture.")
    if scale > 6.:
        raise ValueError("Masksembles approach couldn't be used in such setups where "
                         f"scale parameter is larger then 6. Current value is (scale={scale}).")
    active = int(int(c) / (scale * (1 - (1 - 1 / scale) ** n)))
    masks = generate_masks(active, n, scale)
    s = None
    for s in np.linspace(max(0.8 * scale, 1.0), 1.5 * scale, 300):
        if masks.shape[-1] >= c:
            break
        masks = generate_masks(active, n, s)
    upper = s
    if masks.shape[-1] != c:
        for s in np.linspace(max(0.8 * scale, 1.0), upper, 1000):
            if masks.shape[-1] >= c:
                break
            masks = generate_masks(active, n, s)
    if masks.shape[-1] != c:
        raise ValueError("generation_wrapper function failed to generate masks with "
                         "requested number of features. Please try to change scale parameter")
    return masks


def kept_channels(masks):
    """Per-mask sorted indices of the kept channels (what a gathered GEMM would index with)."""
    m = masks.detach().cpu().numpy() if hasattr(masks, "detach") else np.asarray(masks)
    return [np.flatnonzero(row > 0).astype(np.int32) for row in m]


def _site_forward(x, kind, p, masks, mask_rows, seed, stream_id, sample0, channel_dim_last=False):
    """Run one stand-alone stochastic layer on a CUDA tensor through ``bnn_dropout``.

    x is NCHW / [B, F] like in the reference; the kernel works channels-last, so 4-D inputs are
    viewed through ``channels_last`` strides (a layout change of the caller's tensor, not arithmetic).
    ``mask_rows`` (Masksembles training branch) gives the mask row of every image group.
    """
    if not x.is_cuda:
        raise RuntimeError("bayesnn_fpga_b200 layers run on a CUDA (sm_100) device only; there is no CPU "
                           "fallback - move the input with .cuda()")
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.require_device()
        dt = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}.get(x.dtype)
        if dt is None:
            raise TypeError("unsupported dtype %s" % x.dtype)
        if x.dim() > 2:
            perm = [0] + list(range(2, x.dim())) + [1]
            xin = x.permute(*perm).contiguous()
        else:
            xin = x.contiguous()
        B = xin.shape[0]
        C = xin.shape[-1]
        per_image = xin[0].numel() if B > 0 else C
        out = torch.empty_like(xin)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        d = _lib.DropDesc()
        d.kind, d.p, d.seed, d.stream_id = kind, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id)
        if masks is not None:
            masks = masks.contiguous()
            d.masks, d.n_masks = masks.data_ptr(), masks.shape[0]
        if B > 0:
            if mask_rows is None:
                d.sample0, d.batch, d.cnt0 = int(sample0), B, 0
                _lib.check(lib.bnn_dropout(xin.data_ptr(), out.data_ptr(), dt, per_image, C, 1, 1,
                                           ctypes.byref(d), stream))
            else:
                # training branch: image group g uses mask row g (utils.py:158-164) - one launch per group
                n = masks.shape[0]
                grp = B // n
                for g in range(n):
                    d.sample0, d.batch, d.cnt0 = 0, grp, g
                    off = g * grp * per_image * xin.element_size()
                    _lib.check(lib.bnn_dropout(xin.data_ptr() + off, out.data_ptr() + off, dt, per_image, C, 1, 1,
                                               ctypes.byref(d), stream))
        if x.dim() > 2:
            inv = [0, x.dim() - 1] + list(range(1, x.dim() - 1))
            out = out.permute(*inv)
    return out


class _MasksemblesBase(nn.Module):
    def __init__(self, channels: int, n: int, scale: float):
        super().__init__()
        self.channels = channels
        self.n = n
        self.scale = scale
        self.cnt = 0
        # numpy's column selection inside the generator may return a non-C-contiguous array; the
        # kernels index the table as row-major [n][C]
        masks = torch.from_numpy(np.ascontiguousarray(generation_wrapper(channels, n, scale))).float()
        self.masks = torch.nn.Parameter(masks, requires_grad=False)

    def forward(self, inputs):
        batch = inputs.shape[0]
        masks = self.masks.to(inputs.device)
        if self.training:
            if batch % self.n != 0:
                raise ValueError('Batch size must be divisible by n, got batch {} and n {}'.format(batch, self.n))
            return _site_forward(inputs, _lib.DROP_MASKSEMBLES, 0.0, masks, True, 0, 0, 0)
        d_cnt = self.cnt
        self.cnt = (self.cnt + 1) % self.n
        out = _site_forward(inputs, _lib.DROP_MASKSEMBLES, 0.0, masks, None, 0, 0, d_cnt)
        return out

    def extra_repr(self):
        return 'scale={}, n={}'.format(self.scale, self.n)


class Masksembles2D(_MasksemblesBase):
    """Drop-in for utils.py:115-174. Input/Output (N, C, H, W); eval: ``x * masks[cnt]`` then
    ``cnt = (cnt + 1) % n`` (no rescale); train: batch split in n groups, group i times mask i."""

    def forward(self, inputs):
        return super().forward(inputs).float()


class Masksembles1D(_MasksemblesBase):
    """Drop-in for utils.py:177-236. Input/Output (N, C)."""
