import os, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/cuda-fem_b200')
import torch, femx
ctx = femx.Context(0)
mesh = ctx.rectangle_mesh(0,1,0,1,4096,4096)
for spec in ("1","0","1"):
    os.environ["FEMX_SPEC"]=spec
    ts=[]
    for i in range(5):
        torch.cuda.synchronize(); t0=time.perf_counter()
        p=femx.Pattern(ctx, mesh); torch.cuda.synchronize()
        ts.append(round(1e3*(time.perf_counter()-t0),2)); p.close()
    print("FEMX_SPEC",spec,ts)
