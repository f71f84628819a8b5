import sys, os; sys.path.insert(0, "/root/repo")
from tools.conv_bench import bench
for name, sh, kw in [("l2conv+res",(8192,16,16,128,128,3,1),dict(res=True)),("l2conv",(8192,16,16,128,128,3,1),{}),("l2s2",(8192,32,32,64,128,3,2),{}),("l1conv",(256,32,32,64,64,3,1),dict(res=True)),("l3conv+res",(8192,8,8,256,256,3,1),dict(res=True)),("l4conv+res",(8192,4,4,512,512,3,1),dict(res=True))]:
    for env in sys.argv[1:] or [""]:
        for kv in env.split(","):
            if kv: k_,v_=kv.split("="); os.environ[k_]=v_
        ms,tf=bench(*sh,**kw)
        print("%-12s %-28s %.3f ms %7.1f TFLOP/s"%(name,env,ms,tf),flush=True)
        for kv in env.split(","):
            if kv: os.environ.pop(kv.split("=")[0])
