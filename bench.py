#!/usr/bin/env python
"""Benchmark of the MSDeformAttn hot path (BASELINE.json metric: fwd+bwd queries*heads/s and HBM
GB/s as a fraction of peak, at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--extended]

A "step" is one forward + one backward of the operator over one batch of synthetic input
(workload c2_detr_encoder by default = BASELINE.json configs[1]: N=2, M=8, D=32, L=4, P=4,
levels of an 800x1333 image, Lq = S = 22223), including the zero-fill of grad_value (timed as part
of the backward: it is part of the op).  The default line also carries, under
"workloads", forward / backward times and roofline fractions of the other BASELINE configurations
(C3 ARCTIC encoder + decoder, C4 DINO decoder, C5 high-res bf16 in the accurate accumulation mode)
and, under "prologue", the module-level comparison fused / unfused prologue and fp32 / bf16 value.

 value    whole-job queries*heads/s with the inputs resident in HBM (device pointers through the
          C ABI).  Each step uses a different one of several input sets so that nothing a step
          reads is left in L2 from its previous use ("inputs larger than L2").
 e2e      the same metric through the host-buffer C-ABI call msda_host_fwd_bwd_f32: pinned HOST
          buffers in, HOST buffers out, all host<->device copies inside the timed region.
 roofline dominant kernel (backward): algorithmic bytes / its mean CUDA-event duration in the timed
          region, against the measured HBM copy peak (MEASURED_PEAKS.json).
 cpu_baseline  the reference's CPU path ms_deform_attn_core_pytorch -- the UNMODIFIED reference file staged in
          oracle/_ref/ by __graft_entry__.build() (kind "reference"), else the port in oracle/grid_sample_port.py
          (kind "port") -- timed on this box's host cores on a bounded sample of the same workload (rank 0, N=1 only).

Multi-GPU (torchrun, one rank per GPU): the path shards over the batch with no exchange step, so
every rank runs its own batch (weak scaling) and the only collective is the timing reduction.
--impl reference: rank 0 alone times that CPU path with all host threads on the full configuration (every image
of the batch), --steps / --warmup honoured (clamped only by a wall-clock budget); other ranks exit 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "msda_fwd_bwd_queries_x_heads_per_s"
UNIT = "queries*heads/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_detr_encoder")
    ap.add_argument("--loc-dist", default=None, help="override the workload's location distribution")
    ap.add_argument("--extended", action="store_true", help="also time the other workloads / kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--main-only", action="store_true", help="skip the other-workloads / prologue blocks (launch lists)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.lower().startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_cpu_function():
    """(callable(value, levels, loc, aw) -> out, kind).  The unmodified reference file
    models/ops/functions/ms_deform_attn_func.py is staged by __graft_entry__.build() under oracle/_ref/ (git-ignored,
    travels to the GPU box); it imports the CUDA extension at module top, which is satisfied with an EMPTY stub module
     -- none of this repository's kernels is on that path.  Falls back to the port."""
    import importlib.util
    import types
    path = os.path.join(ROOT, "oracle", "_ref", "reference_func", "ms_deform_attn_func.py")
    if os.path.exists(path):
        try:
            import torch
            saved = sys.modules.get("MultiScaleDeformableAttention")
            sys.modules["MultiScaleDeformableAttention"] = types.ModuleType("MultiScaleDeformableAttention")
            try:
                spec = importlib.util.spec_from_file_location("_uvhand_reference_func", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
            finally:
                if saved is not None:
                    sys.modules["MultiScaleDeformableAttention"] = saved
                else:
                    sys.modules.pop("MultiScaleDeformableAttention", None)
            core = mod.ms_deform_attn_core_pytorch

            def run(value, levels, loc, aw):
                return core(value, torch.as_tensor(levels, dtype=torch.long), loc, aw)
            return run, "reference"
        except Exception as e:      # noqa: BLE001
            print(f"bench.py: staged reference function unusable ({e!r}); using the port", file=sys.stderr)
    from oracle.grid_sample_port import msda_grid_sample
    return msda_grid_sample, "port"


def cpu_baseline_run(wl_name: str, loc_dist, threads: int, steps: int, warmup: int, n_images=None, budget_s=150.0):
    """Times the reference's CPU path (fwd + autograd bwd, fp32) on `n_images` images of the workload (all queries;
    None = the whole batch).  Returns items/s, seconds per step, description, kind, steps actually timed."""
    import torch
    from uvhand_b200.workloads import WORKLOADS, Workload, make_inputs
    fn, kind = reference_cpu_function()
    wl = WORKLOADS[wl_name]
    n = wl.N if n_images is None else n_images
    sample = Workload(wl.name + f"_n{n}", n, wl.levels, wl.Lq, M=wl.M, D=wl.D, P=wl.P, loc_dist=wl.loc_dist)
    torch.set_num_threads(threads)
    x = make_inputs(sample, seed=0, device="cpu", loc_dist=loc_dist)
    levels = [tuple(l) for l in sample.levels]
    times = []
    t_start = time.perf_counter()
    it = 0
    while it < warmup + steps:
        v = x["value"].clone().requires_grad_(True)
        lo = x["loc"].clone().requires_grad_(True)
        a = x["aw"].clone().requires_grad_(True)
        t0 = time.perf_counter()
        out = fn(v, levels, lo, a)
        out.backward(x["grad_out"])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        it += 1
        if times and time.perf_counter() - t_start > budget_s:      # wall-clock budget: stop early, say so
            break
    t = statistics.mean(times)
    desc = (f"{sample.name}: {n} of {wl.N} images of {wl.name} (all {wl.Lq} queries each), fp32 fwd + autograd bwd, "
            f"{len(times)} timed runs after {warmup} warm-up, {threads} threads")
    return sample.items / t, t, desc, kind, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    threads = os.cpu_count() or 1
    val, t, desc, kind, steps = cpu_baseline_run(args.workload, args.loc_dist, threads, max(1, args.steps),
                                                 max(0, args.warmup))
    from uvhand_b200.workloads import WORKLOADS
    wl = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "N": wl.N, "M": wl.M, "D": wl.D, "L": wl.L, "P": wl.P,
                   "Lq": wl.Lq, "S": wl.S, "loc_dist": args.loc_dist or wl.loc_dist,
                   "note": "reference CPU path ms_deform_attn_core_pytorch (F.grid_sample), whole batch per step, host cores only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc,
                         "torch": torch.__version__, "cpu": cpu_model()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if steps < args.steps:
        line["config"]["note"] += f"; stopped after {steps} of {args.steps} steps (wall-clock budget)"
    print(json.dumps(line))
    return 0


def time_device_steps(torch, cabi, wl, sets, bufs, kind, steps, warmup, per_kernel=True, bf16_direct=False):
    """K steps of memset(grad_value) + fwd + bwd on the current stream.  Returns total ms and the
    per-kernel mean durations (CUDA events on the launching stream)."""
    dims = wl.dims()
    st = torch.cuda.current_stream().cuda_stream

    def step(i, ev=None):
        # order: forward | zero-fill of grad_value (or of the fp32 accumulation buffer) | backward.  The zero-fill
        # is part of the op (the reference does it inside ms_deform_attn_cuda_backward, .cu:121): it is timed
        # with the backward (events 1 -> 3), event 2 separates it for the record.
        x = sets[i % len(sets)]
        scratch = None
        if ev:
            ev[0].record()
        cabi.forward(kind, x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(), x["loc"].data_ptr(),
                     x["aw"].data_ptr(), dims, bufs["out"].data_ptr(), st)
        if ev:
            ev[1].record()
        if kind == "bf16" and not bf16_direct:
            if "scratch" not in bufs:
                bufs["scratch"] = torch.zeros(bufs["gv"].shape, device="cuda", dtype=torch.float32)
            scratch = bufs["scratch"]
            scratch.zero_()          # fp32 accumulation buffer; grad_value itself is fully overwritten
        else:
            bufs["gv"].zero_()
        if ev:
            ev[2].record()
        cabi.backward(kind, x["grad_out"].data_ptr(), x["value"].data_ptr(), x["shapes"].data_ptr(),
                      x["lsi"].data_ptr(), x["loc"].data_ptr(), x["aw"].data_ptr(), dims, bufs["gv"].data_ptr(),
                      bufs["gl"].data_ptr(), bufs["ga"].data_ptr(), st,
                      bf16_scratch_ptr=scratch.data_ptr() if scratch is not None else None)
        if ev:
            ev[3].record()

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)] if per_kernel else None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    return step, evs, t0, t1


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from uvhand_b200 import sharding
    world, rank, local_rank = sharding.env_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the operator has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dist = sharding.init_process_group("nccl", torch.device("cuda", local_rank))
    n_gpus = world if world > 1 else 1

    from uvhand_b200 import cabi
    from uvhand_b200.workloads import WORKLOADS, make_inputs
    wl = WORKLOADS[args.workload]
    kind = "bf16" if args.workload.startswith("c5") else "f32"
    vdt = torch.bfloat16 if kind == "bf16" else torch.float32
    vbytes = 2 if kind == "bf16" else 4
    loc_dist = args.loc_dist or wl.loc_dist
    n_sets = 4
    sets = [make_inputs(wl, seed=1000 * rank + s, device="cuda", value_dtype=vdt, loc_dist=loc_dist) for s in range(n_sets)]
    bufs = {"out": torch.empty(wl.N, wl.Lq, wl.M * wl.D, device="cuda"),
            "gv": torch.zeros_like(sets[0]["value"]), "gl": torch.empty_like(sets[0]["loc"]),
            "ga": torch.empty_like(sets[0]["aw"])}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    step, evs, t0, t1 = time_device_steps(torch, cabi, wl, sets, bufs, kind, args.steps, args.warmup)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    n_launch0 = cabi.launch_count()
    t0.record()
    for i in range(args.steps):
        step(i, evs[i])
    t1.record()
    barrier()
    launches = cabi.launch_count() - n_launch0
    fwd_name, bwd_name = cabi.last_path(False), cabi.last_path(True)     # the kernels the timed steps ran
    clocks = sampler.stop() if sampler else None
    total_ms = t0.elapsed_time(t1)
    fwd_all = [e[0].elapsed_time(e[1]) for e in evs]
    bwd_all = [e[1].elapsed_time(e[3]) for e in evs]          # zero-fill + backward kernel(s)
    memset_all = [e[1].elapsed_time(e[2]) for e in evs]
    fwd_ms, bwd_ms, memset_ms = statistics.mean(fwd_all), statistics.mean(bwd_all), statistics.mean(memset_all)
    spread = {"fwd_ms_median": statistics.median(fwd_all), "fwd_ms_min": min(fwd_all),
              "bwd_ms_median": statistics.median(bwd_all), "bwd_ms_min": min(bwd_all)}   # this rank's
    total_ms, fwd_ms, bwd_ms, memset_ms = sharding.max_over_ranks([total_ms, fwd_ms, bwd_ms, memset_ms], dist, "cuda")
    ms_per_step = total_ms / args.steps
    value = sharding.aggregate_throughput(wl.items, n_gpus, ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        import numpy as np
        x = {k: v.cpu().numpy() for k, v in sets[0].items()}
        if kind == "bf16":
            x["value"] = sets[0]["value"].float().cpu().numpy()   # host API is fp32
        host = {k: np.ascontiguousarray(x[k]) for k in ("value", "loc", "aw", "grad_out")}
        outs = {"out": np.empty((wl.N, wl.Lq, wl.M * wl.D), np.float32), "gv": np.empty(host["value"].shape, np.float32),
                "gl": np.empty(host["loc"].shape, np.float32), "ga": np.empty(host["aw"].shape, np.float32)}
        for a in list(host.values()) + list(outs.values()):
            cabi.host_register(a)
        plan = cabi.HostPlan(*wl.dims(), device=local_rank)
        h2d, d2h = plan.traffic()
        e2e_steps = max(3, min(args.steps, 20))

        def e2e_step():
            plan.forward_backward(host["value"], x["shapes"], x["lsi"], host["loc"], host["aw"], host["grad_out"],
                                  outs["out"], outs["gv"], outs["gl"], outs["ga"])
        for _ in range(3):
            e2e_step()
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - w0) / e2e_steps
        e2e_s = sharding.max_over_ranks([e2e_s], dist, "cuda")[0]
        e2e = {"value": sharding.aggregate_throughput(wl.items, n_gpus, e2e_s), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "per_gpu_value": wl.items / e2e_s,
               "pcie_gbs_per_gpu_each_way": [h2d / e2e_s / 1e9, d2h / e2e_s / 1e9],
               "api": "msda_host_fwd_bwd_f32 (C ABI, pinned host buffers in and out)",
               "note": "host<->device copies dominate (PCIe); with N GPUs on one host the per-GPU value falls as the "
                       "ranks share the host's pinned-memory / PCIe root bandwidth -- not NCCL, not a kernel"}
        plan.close()
        # A stream of steps, `depth` in flight (msda_host_pipe_*): what a training loop that prefetches its next batch
        # does.  Every step's inputs still go up and every step's results still come down inside the timed region (each
        # slot has its own pinned result buffers); what goes away is the per-call pipeline fill and drain, because the
        # uploads of step i + 1 overlap the kernels and downloads of step i.
        try:
            depth = 3
            slot_outs = [outs] + [{k: np.empty_like(v) for k, v in outs.items()} for _ in range(depth - 1)]
            for o in slot_outs[1:]:
                for v in o.values():
                    cabi.host_register(v)
            pipe = cabi.HostPipe(*wl.dims(), device=local_rank, depth=depth)

            def pipe_steps(n):
                for i in range(n):
                    o = slot_outs[i % depth]
                    pipe.submit(i % depth, host["value"], x["shapes"], x["lsi"], host["loc"], host["aw"], host["grad_out"],
                                o["out"], o["gv"], o["gl"], o["ga"])
                pipe.wait()
            pipe_steps(depth + 1)
            barrier()
            w0 = time.perf_counter()
            pipe_steps(e2e_steps)
            torch.cuda.synchronize()
            ps = sharding.max_over_ranks([(time.perf_counter() - w0) / e2e_steps], dist, "cuda")[0]
            same = all(bool(np.array_equal(slot_outs[s]["out"], outs["out"])) for s in range(1, depth))
            e2e["blocking_call"] = {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"], "api": e2e["api"],
                                    "note": "one synchronous msda_host_fwd_bwd_f32 call per step (round-1 definition of e2e)"}
            e2e.update({"value": sharding.aggregate_throughput(wl.items, n_gpus, ps), "ms_per_step": ps * 1e3,
                        "per_gpu_value": wl.items / ps, "pcie_gbs_per_gpu_each_way": [h2d / ps / 1e9, d2h / ps / 1e9],
                        "steps_in_flight": depth, "slots_agree": same,
                        "api": "msda_host_pipe_submit_f32 / msda_host_pipe_wait (C ABI, pinned host buffers in and out, "
                               f"{depth} steps in flight; every step uploads its inputs and downloads its results)"})
            pipe.close()
            for o in slot_outs[1:]:
                for v in o.values():
                    cabi.host_unregister(v)
        except Exception as exc:      # noqa: BLE001
            e2e["pipelined_error"] = repr(exc)
        # the same call with bf16 storage for the two value-shaped tensors (value up, grad_value down): informational,
        # the contract's e2e above is the fp32 reference-facing call
        try:
            v16 = np.ascontiguousarray(sets[0]["value"].to(torch.bfloat16).view(torch.int16).cpu().numpy().view(np.uint16))
            gv16 = np.empty(v16.shape, np.uint16)
            cabi.host_register(v16)
            cabi.host_register(gv16)
            plan16 = cabi.HostPlan(*wl.dims(), device=local_rank, value_bf16=True)
            h2d16, d2h16 = plan16.traffic()

            def e2e16_step():
                plan16.forward_backward(v16, x["shapes"], x["lsi"], host["loc"], host["aw"], host["grad_out"],
                                        outs["out"], gv16, outs["gl"], outs["ga"])
            for _ in range(3):
                e2e16_step()
            barrier()
            w0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e16_step()
            torch.cuda.synchronize()
            e16 = sharding.max_over_ranks([(time.perf_counter() - w0) / e2e_steps], dist, "cuda")[0]
            e2e["bf16_value_storage"] = {"value": sharding.aggregate_throughput(wl.items, n_gpus, e16), "ms_per_step": e16 * 1e3,
                                         "h2d_bytes_per_step": h2d16, "d2h_bytes_per_step": d2h16,
                                         "api": "msda_host_fwd_bwd_bf16 (value / grad_value travel as bf16; 1e-2 tolerance mode)"}
            plan16.close()
            # ... and as a stream of steps (three in flight), like e2e.value
            depth16 = 3
            gv16s = [gv16] + [np.empty_like(gv16) for _ in range(depth16 - 1)]
            o16 = [outs] + [{k: np.empty_like(v) for k, v in outs.items() if k != "gv"} for _ in range(depth16 - 1)]
            for g in gv16s[1:]:
                cabi.host_register(g)
            for o in o16[1:]:
                for v in o.values():
                    cabi.host_register(v)
            pipe16 = cabi.HostPipe(*wl.dims(), device=local_rank, depth=depth16, value_bf16=True)

            def pipe16_steps(n):
                for i in range(n):
                    o = o16[i % depth16]
                    pipe16.submit(i % depth16, v16, x["shapes"], x["lsi"], host["loc"], host["aw"], host["grad_out"],
                                  o["out"], gv16s[i % depth16], o["gl"], o["ga"])
                pipe16.wait()
            pipe16_steps(depth16 + 1)
            barrier()
            w0 = time.perf_counter()
            pipe16_steps(e2e_steps)
            torch.cuda.synchronize()
            p16 = sharding.max_over_ranks([(time.perf_counter() - w0) / e2e_steps], dist, "cuda")[0]
            e2e["bf16_value_storage"]["steps_in_flight"] = {"value": sharding.aggregate_throughput(wl.items, n_gpus, p16),
                                                            "ms_per_step": p16 * 1e3, "depth": depth16,
                                                            "api": "msda_host_pipe_submit_bf16 / msda_host_pipe_wait"}
            pipe16.close()
            for g in gv16s[1:]:
                cabi.host_unregister(g)
            for o in o16[1:]:
                for v in o.values():
                    cabi.host_unregister(v)
            cabi.host_unregister(v16)
            cabi.host_unregister(gv16)
        except Exception as exc:      # noqa: BLE001
            e2e["bf16_value_storage"] = {"error": repr(exc)}
        for a in list(host.values()) + list(outs.values()):
            cabi.host_unregister(a)

    # ---- extras: other kernels / workloads (rank 0, not part of the contract keys) ------------
    extras, other, prologue = {}, {}, {}
    if rank == 0 and n_gpus == 1 and not args.main_only:
        del sets, bufs
        torch.cuda.empty_cache()
        other = other_workloads(torch, cabi)
        prologue = prologue_block(torch)
    if args.extended and rank == 0:
        extras = extended(torch, cabi, args)

    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    bytes_bwd = wl.bytes_bwd(vbytes)
    bytes_fwd = wl.bytes_fwd(vbytes)
    achieved = bytes_bwd / (bwd_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get("bwd_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16-value/f32" if kind == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "N_per_gpu": wl.N, "M": wl.M, "D": wl.D, "L": wl.L, "P": wl.P,
                   "Lq": wl.Lq, "S": wl.S, "levels": [list(l) for l in wl.levels], "loc_dist": loc_dist,
                   "parallelism": f"batch-sharded x{n_gpus}, no collective in the op",
                   "l2": f"inputs larger than L2: {n_sets} rotating input sets, {(bytes_fwd + bytes_bwd) / 1e6:.0f} MB touched per step vs 126 MB L2",
                   "step": "memset(grad_value) + msda_fwd + msda_bwd through the C ABI on device pointers"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "kernels": {"fwd": fwd_name, "bwd": bwd_name, "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "bwd_memset_ms": memset_ms,
                    "bwd_note": "bwd_ms = zero-fill of grad_value + backward kernel (CUDA events on the launching stream)",
                    "fwd_gbs": bytes_fwd / (fwd_ms * 1e-3) / 1e9, "fwd_frac_of_peak": bytes_fwd / (fwd_ms * 1e-3) / 1e9 / peak,
                    "fwd_bwd_gbs": (bytes_fwd + bytes_bwd) / (ms_per_step * 1e-3) / 1e9,
                    "fwd_bwd_frac_of_peak": (bytes_fwd + bytes_bwd) / (ms_per_step * 1e-3) / 1e9 / peak,
                    "fwd_bwd_frac_of_8TBs": (bytes_fwd + bytes_bwd) / (ms_per_step * 1e-3) / 1e9 / 8000.0, **spread},
        "roofline": {"bound": "hbm", "kernel": bwd_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": bytes_bwd,
                     "peak_source": peak_src,
                     "duration_note": "achieved = algorithmic bytes / (zero-fill + backward kernel) mean duration",
                     "traffic_source": "ncu --set full capture of the same command, profiles/roofline_traffic.json"},
    }
    if other:
        line["workloads"] = other
    if prologue:
        line["prologue"] = prologue
    if n_gpus == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, t, desc, ckind, _ = cpu_baseline_run(args.workload, args.loc_dist, threads, 3, 1, n_images=1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": ckind, "sample": desc,
                                "ms_per_sample_step": t * 1e3, "cpu": cpu_model()}
        # SURVEY 8d also asks for the single-thread figure (one timed run after one warm-up)
        v1, t1_, _, _, _ = cpu_baseline_run(args.workload, args.loc_dist, 1, 1, 1, n_images=1)
        line["cpu_baseline"]["single_thread_value"] = v1
        line["cpu_baseline"]["single_thread_ms_per_sample_step"] = t1_ * 1e3
    else:
        line["cpu_baseline"] = None
    if extras:
        line["extras"] = extras
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def _time_workload(torch, cabi, name, kind, loc_dist=None, steps=20, warmup=3, n_sets=3, bf16_direct=False):
    from uvhand_b200.workloads import WORKLOADS, make_inputs
    wl = WORKLOADS[name]
    vdt = torch.bfloat16 if kind == "bf16" else torch.float32
    sets = [make_inputs(wl, seed=s, device="cuda", value_dtype=vdt, loc_dist=loc_dist) for s in range(n_sets)]
    bufs = {"out": torch.empty(wl.N, wl.Lq, wl.M * wl.D, device="cuda"), "gv": torch.zeros_like(sets[0]["value"]),
            "gl": torch.empty_like(sets[0]["loc"]), "ga": torch.empty_like(sets[0]["aw"])}
    step, evs, t0, t1 = time_device_steps(torch, cabi, wl, sets, bufs, kind, steps, warmup, bf16_direct=bf16_direct)
    t0.record()
    for i in range(steps):
        step(i, evs[i])
    t1.record()
    torch.cuda.synchronize()
    peak, _ = peaks()
    vb = 2 if kind == "bf16" else 4
    ms = t0.elapsed_time(t1) / steps
    f = statistics.median(e[0].elapsed_time(e[1]) for e in evs)
    b = statistics.median(e[1].elapsed_time(e[3]) for e in evs)
    res = {"ms_per_step": ms, "fwd_ms": f, "bwd_ms": b, "items_per_s": wl.items / (ms * 1e-3),
           "fwd_frac_of_peak": wl.bytes_fwd(vb) / (f * 1e-3) / 1e9 / peak,
           "fwd_bwd_frac_of_peak": (wl.bytes_fwd(vb) + wl.bytes_bwd(vb)) / ((f + b) * 1e-3) / 1e9 / peak,
           "fwd_bwd_gbs": (wl.bytes_fwd(vb) + wl.bytes_bwd(vb)) / (ms * 1e-3) / 1e9,
           "kernels": [cabi.last_path(False), cabi.last_path(True)], "dtype": "bf16-value/f32" if kind == "bf16" else "f32",
           "shape": {"N": wl.N, "Lq": wl.Lq, "S": wl.S}, "loc_dist": loc_dist or wl.loc_dist}
    del sets, bufs
    torch.cuda.empty_cache()
    return res


def other_workloads(torch, cabi):
    """The other BASELINE configurations, device-resident, medians of 20 steps (rotating input sets; bwd_ms includes
    the zero-fill).  Fractions are of the measured HBM copy peak (MEASURED_PEAKS.json)."""
    res = {}
    for name, kind in (("c3_arctic_encoder", "f32"), ("c3_arctic_decoder", "f32"), ("c4_dino_decoder", "f32"),
                       ("c4_dino_decoder_train", "f32"), ("c2_detr_encoder", "bf16"), ("c5_hires", "bf16"),
                       ("c5_hires_contended", "bf16")):
        steps = 6 if name.startswith("c5") else 20
        res[f"{name}:{kind}"] = _time_workload(torch, cabi, name, kind, steps=steps, n_sets=2 if name.startswith("c5") else 3)
    return res


def prologue_block(torch):
    """SURVEY 8f ranks 1 and 2 at module level: autograd forward + backward from the raw Linear outputs
    (softmax + location arithmetic + operator), unfused vs fused prologue; and the value path of the module
    (value_proj output -> padding mask -> operator input) in fp32 vs packed once as bf16."""
    from uvhand_b200.functions import MSDeformAttnFunction, MSDeformAttnFusedFunction
    from uvhand_b200.workloads import WORKLOADS, encoder_reference_points, make_inputs
    res = {}
    steps, warmup = 20, 4
    for name in ("c2_detr_encoder", "c3_arctic_encoder"):
        wl = WORKLOADS[name]
        N, S, M, D, L, Lq, P = wl.dims()
        x = make_inputs(wl, seed=0, device="cuda")
        ref = encoder_reference_points(wl.levels, "cuda")[None, :, None, :].expand(N, Lq, L, 2).contiguous()
        leaves = {"value": x["value"], "off": torch.randn(N, Lq, M, L, P, 2, device="cuda") * 2.0,
                  "logits": torch.randn(N, Lq, M, L * P, device="cuda")}
        wh = torch.stack([x["shapes"][:, 1], x["shapes"][:, 0]], -1)

        def unfused(v, off, lg):
            w = torch.softmax(lg, -1).view(N, Lq, M, L, P)
            loc = ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
            return MSDeformAttnFunction.apply(v, x["shapes"], x["lsi"], loc, w, 64)

        def fused(v, off, lg):
            return MSDeformAttnFusedFunction.apply(v, x["shapes"], x["lsi"], ref, off, lg)

        for tag, fn in (("unfused", unfused), ("fused", fused)):
            ts = []
            for it in range(warmup + steps):
                lv = {k: t.detach().requires_grad_(True) for k, t in leaves.items()}
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn(lv["value"], lv["off"], lv["logits"])
                out.backward(x["grad_out"])
                e1.record()
                torch.cuda.synchronize()
                if it >= warmup:
                    ts.append(e0.elapsed_time(e1))
            res[f"{tag}:{name}"] = {"ms_fwd_bwd_from_raw_linear_outputs": statistics.median(ts)}
        del x, leaves
        torch.cuda.empty_cache()
    try:
        res.update(value_path_block(torch))
    except Exception as e:      # noqa: BLE001
        res["value_path_error"] = repr(e)
    return res


def value_path_block(torch):
    """f2: value_proj output [N, S, 256] fp32 -> operator input.  'fp32' = masked_fill + fp32 kernels (the reference
    module); 'bf16' = one fused cast+mask pass to bf16 + bf16-value kernels (MSDeformAttn(value_dtype=bfloat16))."""
    from uvhand_b200.modules import MSDeformAttn
    from uvhand_b200.workloads import WORKLOADS, encoder_reference_points, level_tensors
    res = {}
    wl = WORKLOADS["c2_detr_encoder"]
    N, S, M, D, L, Lq, P = wl.dims()
    shapes, lsi = level_tensors(wl.levels, "cuda")
    refp = encoder_reference_points(wl.levels, "cuda")[None, :, None, :].expand(N, Lq, L, 2).contiguous()
    pad = torch.zeros(N, S, dtype=torch.bool, device="cuda")
    pad[:, -S // 10:] = True
    torch.manual_seed(0)
    base = MSDeformAttn(256, L, M, P, fuse_prologue=False).cuda()
    with torch.no_grad():
        base.sampling_offsets.weight.normal_(0, 0.01)
    for tag, kw in (("module_fp32_unfused", dict(fuse_prologue=False)), ("module_fp32_fused", dict(fuse_prologue=True)),
                    ("module_bf16_value", dict(fuse_prologue=False, value_dtype=torch.bfloat16))):
        mod = MSDeformAttn(256, L, M, P, **kw).cuda()
        mod.load_state_dict(base.state_dict())
        ts = []
        for it in range(12):
            q = torch.randn(N, Lq, 256, device="cuda", requires_grad=True)
            f = torch.randn(N, S, 256, device="cuda", requires_grad=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = mod(q, refp, f, shapes, lsi, pad)
            out.sum().backward()
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        res[f"{tag}:c2_detr_encoder"] = {"ms_module_fwd_bwd": statistics.median(ts)}
        del mod
    return res


def extended(torch, cabi, args):
    """Other workloads, location distributions and kernel families, device-resident, rank 0."""
    import ctypes
    from uvhand_b200.workloads import WORKLOADS, make_inputs
    res = {}
    steps, warmup = 30, 5

    def run(name, kind, loc_dist=None, fwd_path=0, bwd_path=0, tag=None, bf16_direct=False):
        wl = WORKLOADS[name]
        vdt = torch.bfloat16 if kind == "bf16" else torch.float32
        sets = [make_inputs(wl, seed=s, device="cuda", value_dtype=vdt, loc_dist=loc_dist) for s in range(3)]
        bufs = {"out": torch.empty(wl.N, wl.Lq, wl.M * wl.D, device="cuda"), "gv": torch.zeros_like(sets[0]["value"]),
                "gl": torch.empty_like(sets[0]["loc"]), "ga": torch.empty_like(sets[0]["aw"])}
        cabi.set_option("fwd_path", fwd_path)
        cabi.set_option("bwd_path", bwd_path)
        try:
            step, evs, t0, t1 = time_device_steps(torch, cabi, wl, sets, bufs, kind, steps, warmup,
                                                  bf16_direct=bf16_direct)
            t0.record()
            for i in range(steps):
                step(i, evs[i])
            t1.record()
            torch.cuda.synchronize()
        finally:
            cabi.set_option("fwd_path", 0)
            cabi.set_option("bwd_path", 0)
        ms = t0.elapsed_time(t1) / steps
        vb = 2 if kind == "bf16" else 4
        res[tag or f"{name}:{kind}:{loc_dist or wl.loc_dist}"] = {
            "ms_per_step": ms, "fwd_ms": statistics.mean(e[0].elapsed_time(e[1]) for e in evs),
            "bwd_ms": statistics.mean(e[1].elapsed_time(e[3]) for e in evs),
            "items_per_s": wl.items / (ms * 1e-3),
            "fwd_bwd_gbs": (wl.bytes_fwd(vb) + wl.bytes_bwd(vb)) / (ms * 1e-3) / 1e9,
            "kernels": [cabi.last_path(False), cabi.last_path(True)]}
        del sets, bufs
        torch.cuda.empty_cache()

    run("c2_detr_encoder", "f32", "uniform")
    run("c2_detr_encoder", "f32", None, 1, 1, tag="c2_detr_encoder:f32:encoder:anyshape-kernels")
    run("c2_detr_encoder", "bf16")
    for name in ("c3_arctic_encoder", "c3_arctic_decoder", "c4_dino_decoder", "c4_dino_decoder_train"):
        run(name, "f32")
    run("c5_hires", "bf16")
    run("c5_hires", "bf16", bf16_direct=True, tag="c5_hires:bf16:encoder:direct-bf16x2-red")
    run("c5_hires_contended", "bf16")
    run("c5_hires_contended", "bf16", bf16_direct=True, tag="c5_hires_contended:bf16:contended:direct-bf16x2-red")
    run("c5_hires", "f32")

    # the reference's own kernels recompiled for sm_100a (informational; oracle/_ref is test infra)
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libmsda_refcuda.so")
    if os.path.exists(ref_so):
        lib = ctypes.CDLL(ref_so)
        vp = ctypes.c_void_p
        ints = [ctypes.c_int] * 8
        lib.msda_ref_fwd_f32.argtypes = [vp] * 5 + ints + [vp, vp]
        lib.msda_ref_bwd_f32.argtypes = [vp] * 6 + ints + [vp] * 4
        for name, ld in (("c2_detr_encoder", None), ("c2_detr_encoder", "uniform"), ("c3_arctic_encoder", None)):
            wl = WORKLOADS[name]
            N, S, M, D, L, Lq, P = wl.dims()
            sets = [make_inputs(wl, seed=s, device="cuda", loc_dist=ld) for s in range(3)]
            out = torch.empty(N, Lq, M * D, device="cuda")
            gv, gl, ga = torch.empty_like(sets[0]["value"]), torch.empty_like(sets[0]["loc"]), torch.empty_like(sets[0]["aw"])
            st = torch.cuda.current_stream().cuda_stream
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]

            def rstep(i, e=None):
                x = sets[i % 3]
                if e:
                    e[0].record()
                lib.msda_ref_fwd_f32(x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(), x["loc"].data_ptr(),
                                     x["aw"].data_ptr(), N, S, M, D, L, Lq, P, 64, out.data_ptr(), st)
                if e:
                    e[1].record()
                lib.msda_ref_bwd_f32(x["grad_out"].data_ptr(), x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(),
                                     x["loc"].data_ptr(), x["aw"].data_ptr(), N, S, M, D, L, Lq, P, 64, gv.data_ptr(),
                                     gl.data_ptr(), ga.data_ptr(), st)
                if e:
                    e[2].record()
            for i in range(warmup):
                rstep(i)
            torch.cuda.synchronize()
            for i in range(steps):
                rstep(i, ev[i])
            torch.cuda.synchronize()
            f = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
            b = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
            res[f"reference_cuda_sm100a_recompile:{name}:{ld or wl.loc_dist}"] = {
                "fwd_ms": f, "bwd_ms": b, "ms_per_step": f + b, "items_per_s": wl.items / ((f + b) * 1e-3)}
            del sets
            torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    sys.exit(main())
