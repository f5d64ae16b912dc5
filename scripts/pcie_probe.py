import torch, time
x = torch.empty(85*1024*1024, dtype=torch.uint8).pin_memory(); d = torch.empty_like(x, device="cuda")
y = torch.empty(57*1024*1024, dtype=torch.uint8).pin_memory(); e = torch.empty(57*1024*1024, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("H2D 85MB ms", t(lambda: d.copy_(x, non_blocking=True)))
print("D2H 57MB ms", t(lambda: y.copy_(e, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(e, non_blocking=True)
print("both concurrently ms", t(both))
