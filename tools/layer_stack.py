#!/usr/bin/env python
"""SURVEY.md 8f rank 3 -- per-layer launch overhead of the operator inside a model step.

A model step calls the operator once per transformer layer: 6 encoder + 6 decoder calls
(/root/reference/models/arctic_transformer.py:325-331 encoder loop, :412-419 decoder loop), forward in
layer order and backward in reverse.  This tool runs that call stack through the C ABI for

    c3  H2OTR/ARCTIC   encoder N=32, Lq=S=1045       decoder N=32, Lq=300     (one 32-frame window)
    c4  DINO 4-scale   encoder N=2,  Lq=S=22223      decoder N=2,  Lq=1098    (training, DN queries)

(a) eagerly (12 forward + 12 x (zero-fill + backward) host calls per step) and (b) as ONE CUDA graph replay, checks
every layer's output and gradients of the REPLAYED step against the CPU oracle on `--check-frames` samples per layer
(not against the eager run), and prints one JSON line with microseconds per step and per layer call.

    python tools/layer_stack.py [--models c3,c4] [--reps 20] [--check-frames 1]
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STACKS = {"c3": ("c3_arctic_encoder", "c3_arctic_decoder"), "c4": ("c2_detr_encoder", "c4_dino_decoder_train")}
N_ENC = N_DEC = 6


def build_stack(model, device="cuda"):
    """12 layers, each with its own seeded inputs and its own outputs (nothing is shared, as in a real step)."""
    import torch
    from uvhand_b200.workloads import WORKLOADS, make_inputs
    layers = []
    for k in range(N_ENC + N_DEC):
        wl = WORKLOADS[STACKS[model][0 if k < N_ENC else 1]]
        x = make_inputs(wl, seed=100 + k, device=device)
        layers.append({"wl": wl, "x": x, "out": torch.empty(wl.N, wl.Lq, wl.M * wl.D, device=device),
                       "gv": torch.zeros_like(x["value"]), "gl": torch.empty_like(x["loc"]), "ga": torch.empty_like(x["aw"])})
    return layers


def run_stack(layers):
    """One step on torch's current stream: forward through the layers, backward in reverse."""
    import torch
    from uvhand_b200 import cabi
    st = torch.cuda.current_stream().cuda_stream
    for ly in layers:
        x = ly["x"]
        cabi.forward("f32", x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(), x["loc"].data_ptr(),
                     x["aw"].data_ptr(), ly["wl"].dims(), ly["out"].data_ptr(), st)
    for ly in reversed(layers):
        x = ly["x"]
        ly["gv"].zero_()
        cabi.backward("f32", x["grad_out"].data_ptr(), x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(),
                      x["loc"].data_ptr(), x["aw"].data_ptr(), ly["wl"].dims(), ly["gv"].data_ptr(), ly["gl"].data_ptr(),
                      ly["ga"].data_ptr(), st)


def capture(layers):
    import torch
    run_stack(layers)                      # warm-up outside the capture (one-time attribute calls, allocator)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        run_stack(layers)
    return graph


def poison(layers):
    for ly in layers:
        for k in ("out", "gl", "ga", "gv"):
            ly[k].fill_(float("nan"))


def check_against_oracle(layers, frames=1, fwd_rtol=1e-5, grad_rtol=1e-4):
    """Every layer's results (as they stand on the device) against the C oracle, on the first `frames` samples of
    each layer.  Returns the worst relative errors; raises AssertionError outside tolerance."""
    import numpy as np
    from oracle import c_oracle
    worst = {"out": 0.0, "grad_value": 0.0, "grad_loc": 0.0, "grad_aw": 0.0}
    for k, ly in enumerate(layers):
        x = ly["x"]
        sl = slice(0, min(frames, ly["wl"].N))
        a = {n: x[n][sl].cpu().numpy() for n in ("value", "loc", "aw", "grad_out")}
        shapes, lsi = x["shapes"].cpu().numpy(), x["lsi"].cpu().numpy()
        ref_out = c_oracle.forward(a["value"], shapes, lsi, a["loc"], a["aw"], "f32x")
        ref_gv, ref_gl, ref_ga = c_oracle.backward(a["grad_out"], a["value"], shapes, lsi, a["loc"], a["aw"], "f32x")
        for name, got, ref, tol in (("out", ly["out"][sl], ref_out, fwd_rtol), ("grad_value", ly["gv"][sl], ref_gv, grad_rtol),
                                    ("grad_loc", ly["gl"][sl], ref_gl, grad_rtol), ("grad_aw", ly["ga"][sl], ref_ga, grad_rtol)):
            g = got.cpu().numpy().astype(np.float64).reshape(ref.shape)
            err = float(np.abs(g - ref).max() / max(np.abs(ref).max(), 1e-30))
            assert np.isfinite(g).all() and err <= tol, f"layer {k} {name}: relative error {err:.2e} > {tol}"
            worst[name] = max(worst[name], err)
    return worst


def time_stack(fn, reps, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts), min(ts)


def main():
    import torch
    from uvhand_b200 import cabi
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="c3,c4")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--check-frames", type=int, default=1)
    a = ap.parse_args()
    res = {}
    for model in a.models.split(","):
        layers = build_stack(model)
        eager_us, eager_min = time_stack(lambda: run_stack(layers), a.reps)
        graph = capture(layers)
        poison(layers)
        n0 = cabi.launch_count()
        graph.replay()
        torch.cuda.synchronize()
        assert cabi.launch_count() == n0, "a graph replay must not go through the host launch path"
        worst = check_against_oracle(layers, a.check_frames)
        graph_us, graph_min = time_stack(graph.replay, a.reps)
        calls = 2 * len(layers)
        res[model] = {"workloads": list(STACKS[model]), "layers": len(layers), "operator_calls_per_step": calls,
                      "eager_us_per_step": eager_us, "graph_us_per_step": graph_us,
                      "eager_us_per_call": eager_us / calls, "graph_us_per_call": graph_us / calls,
                      "eager_min_us": eager_min, "graph_min_us": graph_min,
                      "host_overhead_removed_us_per_call": (eager_us - graph_us) / calls,
                      "replay_vs_oracle_worst_rel_err": worst, "checked_frames_per_layer": a.check_frames,
                      "kernels_last": [cabi.last_path(False), cabi.last_path(True)]}
        del layers, graph
        torch.cuda.empty_cache()
    print(json.dumps({"tool": "layer_stack", "device": torch.cuda.get_device_name(0), "stacks": res}))


if __name__ == "__main__":
    main()
