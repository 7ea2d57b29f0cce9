"""PCIe probe: pinned host <-> device copy rates, one direction and both at once, for one 160 MB copy and for
the same bytes cut into chunks (what the host pipeline of msda_host_fwd_bwd_f32 issues)."""
import time
import torch
n = 160 * 1024 * 1024 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory(); h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, device="cuda"); d2 = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, k=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k):
        fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k


for mb in (160, 16, 4, 1):
    c = mb * 1024 * 1024 // 4
    def up():
        with torch.cuda.stream(s1):
            for o in range(0, n, c):
                d[o:o + c].copy_(h[o:o + c], non_blocking=True)
    def down():
        with torch.cuda.stream(s2):
            for o in range(0, n, c):
                h2[o:o + c].copy_(d2[o:o + c], non_blocking=True)
    def both():
        up(); down()
    a, b, e = t(up), t(down), t(both)
    print(f"copies of {mb:4d} MiB: H2D {160/1024/a:5.1f} GiB/s  D2H {160/1024/b:5.1f} GiB/s  both at once {e*1e3:5.2f} ms -> {160/1024/e:5.1f} GiB/s each")
