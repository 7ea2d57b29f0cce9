#!/usr/bin/env python
"""One data-parallel training step of a bare MSDeformAttn encoder layer under DistributedDataParallel
(SURVEY.md section 8e): shows that the operator itself exchanges nothing -- every rank runs it on its own
clip -- and that NCCL appears only in DDP's gradient all-reduce of the layer's parameters.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/ddp_step.py [--steps 20] [--workload c3_arctic_encoder] [--no-fuse]

Prints (rank 0) one JSON line: ms per step with DDP, ms per step of the same layer without DDP (no all-reduce),
and the aggregate queries*heads/s of the operator inside the step.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from uvhand_b200 import sharding  # noqa: E402
from uvhand_b200.modules import MSDeformAttn  # noqa: E402
from uvhand_b200.workloads import WORKLOADS, encoder_reference_points, level_tensors  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--workload", default="c3_arctic_encoder")
ap.add_argument("--no-fuse", action="store_true", help="materialise locations / weights with torch ops (the reference module)")
a = ap.parse_args()

world, rank, local = sharding.env_world()
torch.cuda.set_device(local)
d = sharding.init_process_group("nccl", torch.device("cuda", local))
wl = WORKLOADS[a.workload]
torch.manual_seed(0)                                         # same parameters on every rank
layer = MSDeformAttn(wl.M * wl.D, wl.L, wl.M, wl.P, fuse_prologue=not a.no_fuse).cuda()
with torch.no_grad():
    layer.sampling_offsets.weight.normal_(0, 0.01)
ddp = torch.nn.parallel.DistributedDataParallel(layer, device_ids=[local]) if d is not None else layer
opt = torch.optim.SGD(layer.parameters(), lr=1e-3)
shapes, lsi = level_tensors(wl.levels, "cuda")
g = torch.Generator(device="cuda").manual_seed(100 + rank)   # each rank its own clip (weak scaling)
src = torch.randn(wl.N, wl.S, wl.M * wl.D, device="cuda", generator=g)
ref = encoder_reference_points(wl.levels, "cuda")[None, :, None, :].expand(wl.N, wl.S, wl.L, 2).contiguous()


def run(model, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        out = model(src, ref, src, shapes, lsi, None)
        out.square().mean().backward()
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


run(ddp, a.warmup)
if d is not None:
    d.barrier()
ms_ddp = run(ddp, a.steps)
ms_local = run(layer, a.steps)
ms_ddp, ms_local = sharding.max_over_ranks([ms_ddp, ms_local], d, "cuda")
if rank == 0:
    n_param = sum(p.numel() for p in layer.parameters())
    print(json.dumps({"workload": a.workload, "n_gpus": world, "fused_prologue": not a.no_fuse,
                      "ms_per_step_ddp": ms_ddp, "ms_per_step_no_allreduce": ms_local,
                      "allreduce_share_of_step": max(0.0, (ms_ddp - ms_local) / ms_ddp) if world > 1 else 0.0,
                      "allreduce_bytes_per_step": 4 * n_param,
                      "layer_items_per_s": sharding.aggregate_throughput(wl.items, world, ms_ddp * 1e-3),
                      "collectives_inside_the_operator": 0}))
if d is not None:
    d.destroy_process_group()
