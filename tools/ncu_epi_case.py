"""Two launches for `ncu -k regex:conv_tc`: (0) layer2.0.1.conv2 of the C2 step - vertical-halo kernel with residual and
fused element dropout (16 epilogue warps); (1) sibling group 1 (ex1conv1 + layer2.0.0.conv1: two 128-channel stride-2
groups over 32x32x64, sibling-pair kernel)."""
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
r = torch.randn_like(y)
dd = drop_desc(1, 0.25, 0x77, 3, 0, 256)
assert lib.bnn_conv2d_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr(), y.data_ptr(), 1, N, H, H, C, C, 3, 1, 1,
                         ctypes.byref(dd), stream()) == 0
x2 = torch.randn(N, 32, 32, 64, device="cuda", dtype=torch.float16)
w2 = (torch.randn(256, 3, 3, 64, device="cuda") / 24).half()
b2 = torch.randn(256, device="cuda")
outs = [torch.empty(N, 16, 16, 128, device="cuda", dtype=torch.float16) for _ in range(2)]
ys = (ctypes.c_void_p * 2)(*[o.data_ptr() for o in outs])
assert lib.bnn_conv2d_tc_grouped(x2.data_ptr(), w2.data_ptr(), b2.data_ptr(), ys, 2, 3, 0, 1, N, 32, 32, 64, 128, 3, 2, stream()) == 0
torch.cuda.synchronize()
