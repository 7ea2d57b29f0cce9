"""e2e probe: msda_host_fwd_bwd_f32 at C2 with (a) numpy buffers pinned by msda_host_register and (b) torch
pinned allocations; prints ms/step and checks the results against the device-pointer path."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from uvhand_b200 import cabi
from uvhand_b200.workloads import WORKLOADS, make_inputs
wl = WORKLOADS["c2_detr_encoder"]
s = make_inputs(wl, seed=0, device="cuda")
import os
for mode in os.environ.get("E2E_MODES", "register").split(","):
    if mode == "register":
        x = {k: np.ascontiguousarray(v.cpu().numpy()) for k, v in s.items()}
        outs = {"out": np.empty((wl.N, wl.Lq, wl.M*wl.D), np.float32), "gv": np.empty(x["value"].shape, np.float32), "gl": np.empty(x["loc"].shape, np.float32), "ga": np.empty(x["aw"].shape, np.float32)}
        for a in [x["value"], x["loc"], x["aw"], x["grad_out"]] + list(outs.values()): cabi.host_register(a)
    else:
        keep = {k: v.cpu().pin_memory() for k, v in s.items()}
        x = {k: v.numpy() for k, v in keep.items()}
        ko = {"out": torch.empty((wl.N, wl.Lq, wl.M*wl.D)).pin_memory(), "gv": torch.empty(s["value"].shape).pin_memory(), "gl": torch.empty(s["loc"].shape).pin_memory(), "ga": torch.empty(s["aw"].shape).pin_memory()}
        outs = {k: v.numpy() for k, v in ko.items()}
    plan = cabi.HostPlan(*wl.dims())
    def step(): plan.forward_backward(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], outs["out"], outs["gv"], outs["gl"], outs["ga"])
    for _ in range(3): step()
    t=time.perf_counter()
    for _ in range(20): step()
    dt=(time.perf_counter()-t)/20
    print(mode, "e2e ms/step", round(dt*1e3, 3), "items/s", wl.items/dt)
    plan.close()
st = torch.cuda.current_stream().cuda_stream
out = torch.empty(wl.N, wl.Lq, wl.M*wl.D, device="cuda"); gv=torch.zeros_like(s["value"]); gl=torch.empty_like(s["loc"]); ga=torch.empty_like(s["aw"])
cabi.forward("f32", s["value"].data_ptr(), s["shapes"].data_ptr(), s["lsi"].data_ptr(), s["loc"].data_ptr(), s["aw"].data_ptr(), wl.dims(), out.data_ptr(), st)
cabi.backward("f32", s["grad_out"].data_ptr(), s["value"].data_ptr(), s["shapes"].data_ptr(), s["lsi"].data_ptr(), s["loc"].data_ptr(), s["aw"].data_ptr(), wl.dims(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), st)
torch.cuda.synchronize()
rel=lambda a,b: float(np.abs(a-b).max()/np.abs(b).max())
print("host vs device:", rel(outs["out"], out.cpu().numpy()), rel(outs["gv"], gv.cpu().numpy()), rel(outs["gl"], gl.cpu().numpy()), rel(outs["ga"], ga.cpu().numpy()))
