import torch, time
n=584*1024*1024
h=torch.empty(n,dtype=torch.uint8,pin_memory=True); d=torch.empty(n,dtype=torch.uint8,device='cuda')
for name,fn in (('H2D',lambda: d.copy_(h,non_blocking=True)),('D2H',lambda: h.copy_(d,non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print(name, '%.1f GB/s  %.1f ms'%(n/dt/1e9, dt*1e3))
# both directions at once on two streams
s1,s2=torch.cuda.Stream(),torch.cuda.Stream(); h2=torch.empty(n,dtype=torch.uint8,pin_memory=True); d2=torch.empty(n,dtype=torch.uint8,device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print('both', '%.1f GB/s each  %.1f ms'%(n/dt/1e9, dt*1e3))
