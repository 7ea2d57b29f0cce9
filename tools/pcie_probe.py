#!/usr/bin/env python
"""Host<->device copy rates of this box with pinned buffers (the bound of bench.py's e2e line).
    python tools/pcie_probe.py [--mb 159]"""
import argparse
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=159.3)
a = ap.parse_args()
n = int(a.mb * 1e6)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
s3, s4 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s1.synchronize(); s2.synchronize(); s3.synchronize(); s4.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d(chunks=1):
    with torch.cuda.stream(s1):
        s1.wait_stream(torch.cuda.current_stream())
        for c in range(chunks):
            lo, hi = c * n // chunks, (c + 1) * n // chunks
            d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)


def d2h(chunks=1):
    with torch.cuda.stream(s2):
        s2.wait_stream(torch.cuda.current_stream())
        for c in range(chunks):
            lo, hi = c * n // chunks, (c + 1) * n // chunks
            h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s2)


def both2(chunks):
    """two streams per direction, chunks alternating between them"""
    cur = torch.cuda.current_stream()
    for st in (s1, s2, s3, s4):
        st.wait_stream(cur)
    for c in range(chunks):
        lo, hi = c * n // chunks, (c + 1) * n // chunks
        with torch.cuda.stream(s1 if c & 1 else s3):
            d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
        with torch.cuda.stream(s2 if c & 1 else s4):
            h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
    for st in (s1, s2, s3, s4):
        cur.wait_stream(st)


def both(chunks=1):
    h2d(chunks)
    d2h(chunks)


for name, fn in (("H2D alone", h2d), ("D2H alone", d2h), ("both directions at once", both),
                 ("both, 2 chunks each", lambda: both(2)), ("both, 4 chunks each", lambda: both(4)),
                 ("both, 8 chunks each", lambda: both(8)), ("both, 16 chunks each", lambda: both(16)), ("both, 64 chunks each", lambda: both(64)),
                 ("H2D alone, 16 chunks", lambda: h2d(16)), ("D2H alone, 16 chunks", lambda: d2h(16)),
                 ("both, 16 chunks, 2 streams per direction", lambda: both2(16)),
                 ("both, 64 chunks, 2 streams per direction", lambda: both2(64))):
    ms = timed(fn)
    print(f"{name:28s} {ms:7.3f} ms  {n / ms / 1e6:6.1f} GB/s per direction")
