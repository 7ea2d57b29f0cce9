#!/usr/bin/env python
"""Kernel times of the fused-prologue entry points (msda_fused_fwd_f32 / msda_fused_bwd_f32) through the C ABI.
    python tools/fused_bench.py [--workloads c2_detr_encoder,c3_arctic_encoder] [--iters 30]"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from uvhand_b200 import cabi  # noqa: E402
from uvhand_b200.workloads import WORKLOADS, encoder_reference_points, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="c2_detr_encoder,c3_arctic_encoder")
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--no-gref", action="store_true", help="reference points do not need a gradient (the encoder)")
a = ap.parse_args()
st = torch.cuda.current_stream().cuda_stream
for name in a.workloads.split(","):
    wl = WORKLOADS[name]
    N, S, M, D, L, Lq, P = wl.dims()
    sets = []
    for seed in range(3):
        x = make_inputs(wl, seed=seed, device="cuda")
        x["ref"] = (encoder_reference_points(wl.levels, "cuda")[None, :, None, :].expand(N, Lq, L, 2).contiguous()
                    if Lq == S else (torch.rand(N, Lq, 1, 2, device="cuda") * 2 - 1).expand(N, Lq, L, 2).contiguous())   # "twostage"
        x["off"] = torch.randn(N, Lq, M, L, P, 2, device="cuda") * 2.0
        x["logits"] = torch.randn(N, Lq, M, L * P, device="cuda")
        sets.append(x)
    out = torch.empty(N, Lq, M * D, device="cuda")
    gv = torch.zeros(N, S, M, D, device="cuda")
    goff, glog = torch.empty_like(sets[0]["off"]), torch.empty_like(sets[0]["logits"])
    gref = torch.zeros(N, Lq, L, 2, device="cuda")
    tf, tb = [], []
    for it in range(a.iters + 5):
        x = sets[it % 3]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        gv.zero_(); gref.zero_()
        e[0].record()
        cabi.fused_forward(x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(), x["ref"].data_ptr(),
                           x["off"].data_ptr(), x["logits"].data_ptr(), wl.dims(), out.data_ptr(), st)
        e[1].record()
        cabi.fused_backward(x["grad_out"].data_ptr(), x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(),
                            x["ref"].data_ptr(), x["off"].data_ptr(), x["logits"].data_ptr(), wl.dims(), gv.data_ptr(),
                            goff.data_ptr(), glog.data_ptr(), None if a.no_gref else gref.data_ptr(), st)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 5:
            tf.append(e[0].elapsed_time(e[1]) * 1e3)
            tb.append(e[1].elapsed_time(e[2]) * 1e3)
    print(f"{name:28s} fused fwd {statistics.median(tf):7.1f} us   fused bwd {statistics.median(tb):7.1f} us   "
          f"({cabi.last_path(False)}, {cabi.last_path(True)})", flush=True)
