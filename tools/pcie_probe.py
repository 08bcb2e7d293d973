import torch, time
n = 2_700_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for _ in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('H2D pinned GB/s', n/dt/1e9, 'ms', dt*1e3)
