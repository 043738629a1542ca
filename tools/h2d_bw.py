import torch, time
x = torch.empty(400*1044484, dtype=torch.int32).pin_memory()
d = torch.empty_like(x, device='cuda')
for chunk in (None, 48<<20, 256<<20):
    torch.cuda.synchronize(); t0=time.perf_counter()
    if chunk is None:
        d.copy_(x, non_blocking=True)
    else:
        n = chunk//4
        for i in range(0, x.numel(), n):
            d[i:i+n].copy_(x[i:i+n], non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(chunk, f"{x.numel()*4/dt/1e9:.1f} GB/s", f"{dt*1e3:.1f} ms")
y = torch.empty(400*1044484, dtype=torch.int32)
torch.cuda.synchronize(); t0=time.perf_counter(); d.copy_(y); torch.cuda.synchronize(); print('pageable', f"{y.numel()*4/(time.perf_counter()-t0)/1e9:.1f} GB/s")
