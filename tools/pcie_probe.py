import torch, time
n = 210*1024*1024
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(120*1024*1024, dtype=torch.uint8).pin_memory(); d2 = torch.empty(120*1024*1024, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
a = t(lambda: d.copy_(h, non_blocking=True)); print("H2D 210MB alone: %.2f ms %.1f GB/s" % (a*1e3, n/a/1e9))
b = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H 120MB alone: %.2f ms %.1f GB/s" % (b*1e3, h2.numel()/b/1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print("both concurrently: %.2f ms -> H2D %.1f GB/s + D2H %.1f GB/s" % (c*1e3, n/c/1e9, h2.numel()/c/1e9))
