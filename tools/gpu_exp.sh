#!/bin/bash
python - <<'PY'
import sys, torch, time
sys.path.insert(0, '.')
from ultragrid_b200 import api
W,H=7680,4320
dev=torch.device('cuda',0)
src=[torch.randint(0,256,(W*H*2,),dtype=torch.uint8,device=dev) for _ in range(4)]
out=torch.empty(W*H,dtype=torch.uint8,device=dev)
for t in (6,1):
    for i in range(3): api.uyvy_to_dxt(src[i%4],W,H,dxt_type=t,out=out)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): api.uyvy_to_dxt(src[i%4],W,H,dxt_type=t,out=out)
    e1.record(); torch.cuda.synchronize()
    print("dxt type",t,"us/frame",e0.elapsed_time(e1)/20*1e3)
PY
