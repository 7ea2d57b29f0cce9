import torch, time
n = 160*1024*1024//4
h = torch.empty(n, dtype=torch.float32).pin_memory(); h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, device="cuda"); d2 = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, k=5):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/k
a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print(f"H2D {160/1024/a:.1f} GiB/s  D2H {160/1024/b:.1f} GiB/s  both: {c*1e3:.2f} ms -> {160/1024/c:.1f} GiB/s each")
