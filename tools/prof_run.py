#!/usr/bin/env python
"""Minimal driver for ncu captures: a few fwd+bwd steps of one workload through the C ABI.

    python tools/prof_run.py --workload c2_detr_encoder [--loc-dist uniform] [--kind f32|bf16] [--steps 3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from uvhand_b200 import cabi  # noqa: E402
from uvhand_b200.workloads import WORKLOADS, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2_detr_encoder")
ap.add_argument("--loc-dist", default=None)
ap.add_argument("--kind", default="f32")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--fwd-path", type=int, default=0)
ap.add_argument("--bwd-path", type=int, default=0)
ap.add_argument("--opts", default="", help="comma-separated key=value library options")
a = ap.parse_args()

wl = WORKLOADS[a.workload]
vdt = torch.bfloat16 if a.kind == "bf16" else torch.float32
sets = [make_inputs(wl, seed=s, device="cuda", value_dtype=vdt, loc_dist=a.loc_dist) for s in range(2)]
out = torch.empty(wl.N, wl.Lq, wl.M * wl.D, device="cuda")
gv = torch.zeros_like(sets[0]["value"])
gl = torch.empty_like(sets[0]["loc"])
ga = torch.empty_like(sets[0]["aw"])
scratch = torch.zeros(gv.shape, device="cuda") if a.kind == "bf16" else None
cabi.set_option("fwd_path", a.fwd_path)
cabi.set_option("bwd_path", a.bwd_path)
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    cabi.set_option(k, int(v))
st = torch.cuda.current_stream().cuda_stream
torch.cuda.synchronize()
for i in range(a.steps):
    x = sets[i % 2]
    (scratch if scratch is not None else gv).zero_()
    cabi.forward(a.kind, x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(), x["loc"].data_ptr(),
                 x["aw"].data_ptr(), wl.dims(), out.data_ptr(), st)
    cabi.backward(a.kind, x["grad_out"].data_ptr(), x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(),
                  x["loc"].data_ptr(), x["aw"].data_ptr(), wl.dims(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), st,
                  bf16_scratch_ptr=scratch.data_ptr() if scratch is not None else None)
torch.cuda.synchronize()
print("ok", cabi.last_path(False), cabi.last_path(True), float(out.abs().mean()), float(gv.float().abs().mean()))
