import torch
x=[torch.randn(1024,257,1024,device='cuda').to(torch.bfloat16) for _ in range(4)]
def t(f,n=12):
    for i in range(3): f(x[i%4])
    torch.cuda.synchronize()
    ev=[(torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i,(a,b) in enumerate(ev):
        a.record(); f(x[i%4]); b.record()
    torch.cuda.synchronize()
    ms=sorted(a.elapsed_time(b) for a,b in ev)
    return ms[len(ms)//2]
nb=x[0].numel()*2
for name,f in [("sum",lambda a: a.sum()),("max",lambda a: a.amax()),("view int32 sum",lambda a: a.view(torch.int32).sum()),("sum dim-1",lambda a: a.sum(-1))]:
    ms=t(f); print(name, "%.1f us  %.2f TB/s"%(ms*1e3, nb/ms/1e9))
y=torch.empty_like(x[0])
ms=t(lambda a: y.copy_(a)); print("copy %.1f us %.2f TB/s (r+w)"%(ms*1e3, 2*nb/ms/1e9))
