import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from uvhand_b200 import cabi
from uvhand_b200.workloads import WORKLOADS, make_inputs
wl = WORKLOADS["c2_detr_encoder"]
s = make_inputs(wl, seed=0, device="cuda")
x = {k: np.ascontiguousarray(v.cpu().numpy()) for k, v in s.items()}
outs = {"out": np.empty((wl.N, wl.Lq, wl.M*wl.D), np.float32), "gv": np.empty(x["value"].shape, np.float32), "gl": np.empty(x["loc"].shape, np.float32), "ga": np.empty(x["aw"].shape, np.float32)}
for a in [x["value"], x["loc"], x["aw"], x["grad_out"]] + list(outs.values()): cabi.host_register(a)
plan = cabi.HostPlan(*wl.dims())
def step(): plan.forward_backward(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], outs["out"], outs["gv"], outs["gl"], outs["ga"])
for _ in range(3): step()
t=time.perf_counter()
for _ in range(20): step()
dt=(time.perf_counter()-t)/20
print("e2e ms/step", dt*1e3, "items/s", wl.items/dt, "h2d,d2h", plan.traffic())
# check vs device path
st = torch.cuda.current_stream().cuda_stream
out = torch.empty(wl.N, wl.Lq, wl.M*wl.D, device="cuda"); gv=torch.zeros_like(s["value"]); gl=torch.empty_like(s["loc"]); ga=torch.empty_like(s["aw"])
cabi.forward("f32", s["value"].data_ptr(), s["shapes"].data_ptr(), s["lsi"].data_ptr(), s["loc"].data_ptr(), s["aw"].data_ptr(), wl.dims(), out.data_ptr(), st)
cabi.backward("f32", s["grad_out"].data_ptr(), s["value"].data_ptr(), s["shapes"].data_ptr(), s["lsi"].data_ptr(), s["loc"].data_ptr(), s["aw"].data_ptr(), wl.dims(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), st)
torch.cuda.synchronize()
rel=lambda a,b: float(np.abs(a-b).max()/np.abs(b).max())
print("host vs device:", rel(outs["out"], out.cpu().numpy()), rel(outs["gv"], gv.cpu().numpy()), rel(outs["gl"], gl.cpu().numpy()), rel(outs["ga"], ga.cpu().numpy()))
