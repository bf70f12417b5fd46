import torch
dev='cuda:0'
n=1<<30
a=torch.empty(n,dtype=torch.float32,device=dev); b=torch.empty(n,dtype=torch.float32,device=dev)
def t(fn,it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it*1e-3
tw=t(lambda:a.zero_()); print('write only (zero_ 4 GiB):', 4*n/tw/1e9,'GB/s')
tr=t(lambda:a.sum()); print('read only (sum 4 GiB):', 4*n/tr/1e9,'GB/s')
tc=t(lambda:b.copy_(a)); print('copy (r+w 8 GiB):', 8*n/tc/1e9,'GB/s')
