import torch
x=torch.empty(1<<30,dtype=torch.bfloat16,device='cuda'); y=torch.empty_like(x)
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms=t(lambda: x.fill_(1.0)); print("fill 2GiB write-only: %.3f ms %.0f GB/s"%(ms, 2*(1<<30)/ms/1e6))
ms=t(lambda: y.copy_(x)); print("copy 2GiB+2GiB: %.3f ms %.0f GB/s"%(ms, 4*(1<<30)/ms/1e6))
ms=t(lambda: x.sum()); print("read-only 2GiB: %.3f ms %.0f GB/s"%(ms, 2*(1<<30)/ms/1e6))
z=torch.empty(1<<28,dtype=torch.bfloat16,device='cuda')
ms=t(lambda: torch.add(z,z,out=x[:1<<28])); print("r1 w1 (0.5+0.5 GiB)", 2*(1<<29)/ms/1e6)
