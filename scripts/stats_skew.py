import sys, torch
sys.path.insert(0, '/root/repo')
from vq_seg_b200 import ops
dev = torch.device('cuda:0')
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    return sorted(ts)[n//2]
n,d,k=32768,512,1024
x=torch.randn(1,n,d,device=dev)
for nm,idx in [("uniform", torch.randint(0,k,(1,n),device=dev)), ("one code", torch.zeros(1,n,dtype=torch.long,device=dev)),
               ("two codes", (torch.arange(n,device=dev)%2).view(1,n)), ("sorted", (torch.arange(n,device=dev)*k//n).view(1,n))]:
    print(nm, "det", t(lambda: ops.code_stats(x,idx,k,True)), "us; atomic", t(lambda: ops.code_stats(x,idx,k,False)), "us")
