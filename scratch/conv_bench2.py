import sys, os; sys.path.insert(0, "/root/repo")
from scratch.conv_bench import bench
for name, sh, kw in [("l2conv+res",(8192,16,16,128,128,3,1),dict(res=True)),("l2conv",(8192,16,16,128,128,3,1),{}),("l3conv+res",(8192,8,8,256,256,3,1),dict(res=True)),("l4conv+res",(8192,4,4,512,512,3,1),dict(res=True))]:
    for env in ["BNN_TC_NOSWAP=1",""]:
        if env: os.environ["BNN_TC_NOSWAP"]="1"
        ms,tf=bench(*sh,**kw)
        print("%-12s %-18s %.3f ms %7.1f TFLOP/s"%(name,env,ms,tf),flush=True)
        os.environ.pop("BNN_TC_NOSWAP",None)
