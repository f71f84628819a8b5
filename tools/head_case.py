"""One exit head at BASELINE config 4 / config 2 shapes through bnn_exit_head_tc (three kernels) and bnn_exit_head_mma /
bnn_exit_head - for `ncu --metrics gpu__time_duration.sum` launch lists."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesnn_fpga_b200 import _lib
from tests.gpu_util import drop_desc, stream
lib = _lib.load()
for (B, S, C, F_, HW) in ((512, 64, 100, 512, 1), (256, 32, 10, 512, 1), (256, 32, 10, 512, 16)):
    feat = (torch.randn(S * B, HW, F_).abs()).half().cuda()
    w = (torch.randn(C, F_) / np.sqrt(F_)).cuda()
    bias = torch.randn(C).cuda()
    dd = drop_desc(1, 0.5, 1, 2, 0, B)
    w_hi, w_lo = torch.empty(C, F_, dtype=torch.half, device="cuda"), torch.empty(C, F_, dtype=torch.half, device="cuda")
    lib.bnn_split16(w.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), C * F_, 1, stream())
    with_lo = int(HW != 1)
    c_pad = (C + 63) // 64 * 64
    parts = [w_hi, w_lo] + ([w_hi] if with_lo else [])
    w3 = torch.zeros(c_pad, len(parts) * F_, dtype=torch.half, device="cuda"); w3[:C] = torch.cat(parts, 1)
    b_pad = torch.zeros(c_pad, device="cuda"); b_pad[:C] = bias
    a_ws = torch.empty(S * B * (2 if with_lo else 1) * F_, dtype=torch.half, device="cuda")
    l_ws = torch.empty(S * B * c_pad, device="cuda")
    sp, sl, spl = torch.zeros(B, C).cuda(), torch.zeros(B, C).cuda(), torch.zeros(B).cuda()
    wt = w.t().contiguous()
    for it in range(2):
        assert lib.bnn_exit_head_tc(feat.data_ptr(), 1, 1, B, S, HW, F_, C, w3.data_ptr(), b_pad.data_ptr(), c_pad, with_lo,
                                    ctypes.byref(dd), a_ws.data_ptr(), l_ws.data_ptr(), sp.data_ptr(), sl.data_ptr(),
                                    spl.data_ptr(), None, 0, stream()) == 0
        if C > 32:
            lib.bnn_exit_head_mma(feat.data_ptr(), 1, 1, B, S, HW, F_, C, w_hi.data_ptr(), w_lo.data_ptr(), bias.data_ptr(),
                                  ctypes.byref(dd), sp.data_ptr(), sl.data_ptr(), spl.data_ptr(), None, 0, stream())
        else:
            lib.bnn_exit_head(feat.data_ptr(), 1, 1, B, S, HW, F_, C, wt.data_ptr(), bias.data_ptr(), ctypes.byref(dd),
                              sp.data_ptr(), sl.data_ptr(), spl.data_ptr(), None, 0, stream())
    torch.cuda.synchronize()
