import torch
x = torch.empty(2*1024**3, dtype=torch.uint8, device="cuda")
y = torch.empty(2*1024**3, dtype=torch.uint8, device="cuda")
def t(f, n=10):
    f(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
ms = t(lambda: x.zero_()); print("memset   %.0f GB/s" % (x.numel()/ms/1e6))
ms = t(lambda: y.copy_(x)); print("copy     %.0f GB/s (r+w)" % (2*x.numel()/ms/1e6))
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print("read sum %.0f GB/s" % (x.numel()/ms/1e6))
# 20% read / 80% write mix: out[4n] = f(in[n])
a = torch.empty(512*1024**2//2, dtype=torch.bfloat16, device="cuda"); 
o = torch.empty(4*a.numel(), dtype=torch.bfloat16, device="cuda")
ms = t(lambda: torch.mul(a.view(-1,1), 2, out=o.view(-1,4)[:, :1]) if False else o.view(-1,4).copy_(a.view(-1,1).expand(-1,4)))
print("1r:4w    %.0f GB/s" % ((a.numel()*2*5)/ms/1e6))
