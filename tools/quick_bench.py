#!/usr/bin/env python
"""Quick device-resident timing of fwd / bwd for a few workloads (CUDA events, rotating inputs),
plus a sanity check against the reference's recompiled CUDA kernels when oracle/_ref is present.

    [MSDA_B200_LIB=path/to/variant.so] python tools/quick_bench.py [--iters 50] [--workloads a,b]
"""
import argparse
import ctypes
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uvhand_b200 import cabi  # noqa: E402
from uvhand_b200.workloads import WORKLOADS, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--workloads", default="c2_detr_encoder:f32:encoder,c2_detr_encoder:f32:uniform,c2_detr_encoder:bf16:encoder,c3_arctic_encoder:f32:encoder,c4_dino_decoder:f32:twostage")
ap.add_argument("--check", action="store_true")
ap.add_argument("--opts", default="", help="library options, e.g. tile_l0=0x34,tile_mode=1")
a = ap.parse_args()
print("lib:", cabi.LIB_PATH, "opts:", a.opts)
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    cabi.set_option(k, int(v, 0))
st = torch.cuda.current_stream().cuda_stream

ref = None
ref_so = os.path.join(ROOT, "oracle", "_ref", "libmsda_refcuda.so")
if a.check and os.path.exists(ref_so):
    ref = ctypes.CDLL(ref_so)
    vp = ctypes.c_void_p
    ref.msda_ref_fwd_f32.argtypes = [vp] * 5 + [ctypes.c_int] * 8 + [vp, vp]
    ref.msda_ref_bwd_f32.argtypes = [vp] * 6 + [ctypes.c_int] * 8 + [vp] * 4

for spec in a.workloads.split(","):
    name, kind, dist = spec.split(":")
    wl = WORKLOADS[name]
    vdt = torch.bfloat16 if kind == "bf16" else torch.float32
    sets = [make_inputs(wl, seed=s, device="cuda", value_dtype=vdt, loc_dist=dist) for s in range(3)]
    out = torch.empty(wl.N, wl.Lq, wl.M * wl.D, device="cuda")
    gv = torch.zeros_like(sets[0]["value"])
    gl = torch.empty_like(sets[0]["loc"])
    ga = torch.empty_like(sets[0]["aw"])
    dims = wl.dims()
    scratch = torch.zeros(gv.shape, device="cuda") if kind == "bf16" else None

    def fwd(x):
        cabi.forward(kind, x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(), x["loc"].data_ptr(),
                     x["aw"].data_ptr(), dims, out.data_ptr(), st)

    def bwd(x):
        cabi.backward(kind, x["grad_out"].data_ptr(), x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(),
                      x["loc"].data_ptr(), x["aw"].data_ptr(), dims, gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), st,
                      bf16_scratch_ptr=scratch.data_ptr() if scratch is not None else None)

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.iters)]
    for i in range(5):
        (scratch if scratch is not None else gv).zero_(); fwd(sets[i % 3]); bwd(sets[i % 3])
    torch.cuda.synchronize()
    for i in range(a.iters):
        x = sets[i % 3]
        (scratch if scratch is not None else gv).zero_()
        ev[i][0].record(); fwd(x); ev[i][1].record(); bwd(x); ev[i][2].record()
    torch.cuda.synchronize()
    f = [e[0].elapsed_time(e[1]) * 1e3 for e in ev]
    b = [e[1].elapsed_time(e[2]) * 1e3 for e in ev]
    msg = f"{spec:45s} fwd {statistics.median(f):8.1f} us (min {min(f):7.1f})   bwd {statistics.median(b):8.1f} us (min {min(b):7.1f})"
    if ref is not None and kind == "f32":
        x = sets[0]
        N, S, M, D, L, Lq, P = dims
        gv.zero_(); fwd(x); bwd(x)
        o2 = torch.empty_like(out); gv2 = torch.empty_like(gv); gl2 = torch.empty_like(gl); ga2 = torch.empty_like(ga)
        ref.msda_ref_fwd_f32(x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(), x["loc"].data_ptr(),
                             x["aw"].data_ptr(), N, S, M, D, L, Lq, P, 64, o2.data_ptr(), st)
        ref.msda_ref_bwd_f32(x["grad_out"].data_ptr(), x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(),
                             x["loc"].data_ptr(), x["aw"].data_ptr(), N, S, M, D, L, Lq, P, 64, gv2.data_ptr(),
                             gl2.data_ptr(), ga2.data_ptr(), st)
        torch.cuda.synchronize()
        rel = lambda p, q: float((p - q).abs().max() / q.abs().max())
        msg += f"   err vs ref-cuda: out {rel(out, o2):.1e} gv {rel(gv, gv2):.1e} gl {rel(gl, gl2):.1e} ga {rel(ga, ga2):.1e}"
    print(msg, flush=True)
    del sets
    torch.cuda.empty_cache()
