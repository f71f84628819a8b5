"""Two launches for `ncu -k regex:conv_tc`: layer3.0.1.conv2 of the C2 step (8x8 maps, 256 -> 256 channels, cta_group::2,
position-major tiles) with residual + fused element dropout, then the same convolution without site and residual."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesnn_fpga_b200 import _lib
from tests.gpu_util import drop_desc, stream
lib = _lib.load()
N, H, C = 8192, 8, 256
x = torch.randn(N, H, H, C, device="cuda", dtype=torch.float16)
w = (torch.randn(C, 3, 3, C, device="cuda") / (C * 9) ** 0.5).half()
b = torch.randn(C, device="cuda")
y = torch.empty(N, H, H, C, device="cuda", dtype=torch.float16)
r = torch.randn_like(y)
dd = drop_desc(1, 0.25, 0x77, 3, 0, 256)
assert lib.bnn_conv2d_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr(), y.data_ptr(), 1, N, H, H, C, C, 3, 1, 1,
                         ctypes.byref(dd), stream()) == 0
d0 = drop_desc(batch=N)
assert lib.bnn_conv2d_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, y.data_ptr(), 1, N, H, H, C, C, 3, 1, 1,
                         ctypes.byref(d0), stream()) == 0
torch.cuda.synchronize()
