"""Exactly two launches of the layer2 3x3 conv (8192 x 16x16x128 -> 128): per-tap operand-swapped kernel, then the
vertical-halo form (for `ncu -k regex:conv_tc`)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesnn_fpga_b200 import _lib
from tests.gpu_util import drop_desc, stream
lib = _lib.load()
N, H, C = 8192, 16, 128
x = torch.randn(N, H, H, C, device="cuda", dtype=torch.float16)
w = (torch.randn(C, 3, 3, C, device="cuda") / (C * 9) ** 0.5).half()
b = torch.randn(C, device="cuda")
y = torch.empty(N, H, H, C, device="cuda", dtype=torch.float16)
dd = drop_desc(batch=N)
for env in ({"BNN_TC_NO_VH": "1"}, {}):
    os.environ.update(env)
    assert lib.bnn_conv2d_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, y.data_ptr(), 1, N, H, H, C, C, 3, 1, 1,
                             ctypes.byref(dd), stream()) == 0
    for k in env:
        os.environ.pop(k)
torch.cuda.synchronize()
