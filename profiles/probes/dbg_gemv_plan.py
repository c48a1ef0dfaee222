import os, sys, torch
sys.path.insert(0, os.getcwd())
os.environ["MXQ_GEMV_VERBOSE"]="1"
from mxq_b200 import ops
dev=torch.device("cuda:0")
for oc, ic, B in ((4096,11008,4),(4096,11008,2),(4096,11008,1),(11008,4096,4)):
    p={k: torch.zeros(s, dtype=d, device=dev) for k,(s,d) in ops.packed_shapes(oc,ic).items()}
    x=torch.randn(B,ic,device=dev).half()
    try:
        ops.gemv(x,p); torch.cuda.synchronize(); print("ok",oc,ic,B)
    except Exception as e:
        print("FAIL",oc,ic,B,e)
