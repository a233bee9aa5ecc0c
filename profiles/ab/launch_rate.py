import torch, time
x = torch.zeros(1024, device='cuda')
torch.cuda.synchronize()
for n in (1000, 5000):
    t0=time.perf_counter()
    for _ in range(n): x.add_(1.0)
    t1=time.perf_counter()
    torch.cuda.synchronize()
    t2=time.perf_counter()
    print("n=%d host enqueue %.2f us/launch, incl. drain %.2f us/launch"%(n,(t1-t0)/n*1e6,(t2-t0)/n*1e6))
