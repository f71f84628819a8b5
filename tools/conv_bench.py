import sys, os, ctypes; sys.path.insert(0, "/root/repo")
import torch
from bayesnn_fpga_b200 import _lib
from tests.gpu_util import drop_desc, stream
_lib.build(); lib = _lib.load()
def bench(N,H,W,Cin,Cout,k,stride,res=False,relu=True,iters=10):
    pad = 1 if k==3 else 0
    OH=(H+2*pad-k)//stride+1
    x=torch.randn(N,H,W,Cin,device="cuda",dtype=torch.float16)
    w=(torch.randn(Cout,k,k,Cin,device="cuda")/ (Cin*k*k)**0.5).half()
    b=torch.randn(Cout,device="cuda")
    y=torch.empty(N,OH,OH,Cout,device="cuda",dtype=torch.float16)
    r=torch.randn_like(y) if res else None
    dd=drop_desc(batch=N)
    call=lambda: lib.bnn_conv2d_tc(x.data_ptr(),w.data_ptr(),b.data_ptr(),ctypes.c_void_p(r.data_ptr() if res else 0),y.data_ptr(),1,N,H,W,Cin,Cout,k,stride,int(relu),ctypes.byref(dd),stream())
    for _ in range(3): assert call()==0
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): call()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/iters
    fl=2*N*OH*OH*Cout*Cin*k*k
    return ms, fl/ms/1e9
shapes={"l2conv":(8192,16,16,128,128,3,1),"l2s2":(8192,32,32,64,128,3,2),"ds":(8192,32,32,64,128,1,2),"l3conv":(8192,8,8,256,256,3,1)}
for name,sh in shapes.items():
    for env in sys.argv[1:] or ["",]:
        for kv in env.split(","):
            if kv: k_,v_=kv.split("="); os.environ[k_]=v_
        ms,tf=bench(*sh)
        print("%-8s %-40s %.3f ms %7.1f TFLOP/s"%(name,env,ms,tf),flush=True)
        for kv in env.split(","):
            if kv: os.environ.pop(kv.split("=")[0])
