"""One call each of bnn_exit_head (block per image) and bnn_exit_head_rows at the C2 / C5 head shapes, for
`ncu --metrics gpu__time_duration.sum` (per-kernel durations)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesnn_fpga_b200 import _lib
from tests.gpu_util import drop_desc, stream
lib = _lib.load()
C, F_, HW, B = 10, 512, 1, 256
for S in (32, 128):
    feat = (torch.randn(S * B, HW, F_).abs() * 3).half().cuda()
    w = (torch.randn(C, F_) / np.sqrt(F_) * 3).cuda()
    wt = w.t().contiguous()
    bias = torch.randn(C).cuda()
    dd = drop_desc(1, 0.5, 0x99, 2, 0, B)
    l_ws = torch.empty(S * B * C, device="cuda")
    sp, sl, spl = torch.zeros(B, C).cuda(), torch.zeros(B, C).cuda(), torch.zeros(B).cuda()
    for _ in range(2):
        assert lib.bnn_exit_head(feat.data_ptr(), 1, 1, B, S, HW, F_, C, wt.data_ptr(), bias.data_ptr(), ctypes.byref(dd),
                                 sp.data_ptr(), sl.data_ptr(), spl.data_ptr(), None, 0, stream()) == 0
        assert lib.bnn_exit_head_rows(feat.data_ptr(), 1, 1, B, S, HW, F_, C, w.data_ptr(), bias.data_ptr(), ctypes.byref(dd),
                                      l_ws.data_ptr(), sp.data_ptr(), sl.data_ptr(), spl.data_ptr(), None, 0, stream()) == 0
    torch.cuda.synchronize()
