#!/usr/bin/env python
"""Host-buffer call (msda_host_fwd_bwd_f32) timing under the plan's chunk-count knob (MSDA_B200_HOST_CHUNKS).

    python tools/e2e_probe.py [--workload c2_detr_encoder] [--chunks 2,4,8] [--steps 10]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from uvhand_b200 import cabi  # noqa: E402
from uvhand_b200.workloads import WORKLOADS, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2_detr_encoder")
ap.add_argument("--chunks", default="")
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
wl = WORKLOADS[a.workload]
t = make_inputs(wl, seed=0, device="cpu")
host = {k: np.ascontiguousarray(t[k].numpy()) for k in ("value", "loc", "aw", "grad_out")}
shapes, lsi = t["shapes"].numpy(), t["lsi"].numpy()


def outputs():
    o = {"out": np.empty((wl.N, wl.Lq, wl.M * wl.D), np.float32), "gv": np.empty(host["value"].shape, np.float32),
         "gl": np.empty(host["loc"].shape, np.float32), "ga": np.empty(host["aw"].shape, np.float32)}
    for v in o.values():
        v.fill(np.nan)
        cabi.host_register(v)
    return o


for v in host.values():
    cabi.host_register(v)
for _once in (0,):
    for chunks in ([None] + [int(c) for c in a.chunks.split(",") if c]):
        if chunks is None:
            os.environ.pop("MSDA_B200_HOST_CHUNKS", None)
        else:
            os.environ["MSDA_B200_HOST_CHUNKS"] = str(chunks)
        o = outputs()
        plan = cabi.HostPlan(*wl.dims(), device=0)

        def step():
            plan.forward_backward(host["value"], shapes, lsi, host["loc"], host["aw"], host["grad_out"], o["out"], o["gv"],
                                  o["gl"], o["ga"])
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(a.steps):
            step()
        ms = (time.perf_counter() - w0) / a.steps * 1e3
        plan.close()
        nan_left = {k: int(np.isnan(v).sum()) for k, v in o.items()}
        for v in o.values():
            cabi.host_unregister(v)
        print(f"chunks={'default' if chunks is None else chunks:>7}  {ms:7.3f} ms per call   nan left: {nan_left}", flush=True)
