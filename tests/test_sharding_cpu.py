"""Multi-rank host logic on CPU: world_size-2 gloo processes (no GPU).  The operator shards over the
batch with no exchange step, so what there is to test is the partition, that the
union of the per-rank results IS the unsharded result (the oracle stands in for the kernels here:
tests may call it), the max-over-ranks timing reduction, and the reference arm's rank behaviour."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from uvhand_b200 import sharding  # noqa: E402
from uvhand_b200.workloads import Workload, make_inputs  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n,world", [(0, 1), (1, 1), (2, 2), (32, 8), (7, 4), (3, 8), (2048, 8)])
def test_shard_range_partitions_the_batch(n, world):
    spans = [sharding.shard_range(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
        assert b0 == a1 and a0 <= b0
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_shard_range_rejects_bad_requests():
    for bad in [(4, 0, 0), (4, 2, 2), (4, 2, -1), (-1, 2, 0)]:
        with pytest.raises(ValueError):
            sharding.shard_range(*bad)


WL = Workload("shard", 4, ((6, 5), (3, 3)), 11, M=2, D=4, P=2, loc_dist="uniform")


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    d = sharding.init_process_group("gloo")
    assert d is not None and d.get_world_size() == world
    from oracle import c_oracle
    full = make_inputs(WL, seed=11, device="cpu")
    mine = sharding.shard_batch(full, world, rank)
    x = {k: v.numpy() for k, v in mine.items()}
    out = c_oracle.forward(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], "f64")
    gv, gl, ga = c_oracle.backward(x["grad_out"], x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], "f64")
    # no collective on the data path: results are gathered here only so that the TEST can compare
    outs = [None] * world
    d.all_gather_object(outs, (out, gv, gl, ga))
    t = sharding.max_over_ranks([1.0 + rank, 5.0 - rank], d)
    d.barrier()
    if rank == 0:
        q.put((outs, t))
    d.destroy_process_group()


def test_two_rank_gloo_shards_reassemble_to_the_unsharded_result():
    from oracle import c_oracle
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs, t = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert t == [2.0, 5.0]                      # slowest rank wins, element-wise
    full = {k: v.numpy() for k, v in make_inputs(WL, seed=11, device="cpu").items()}
    ref_out = c_oracle.forward(full["value"], full["shapes"], full["lsi"], full["loc"], full["aw"], "f64")
    ref_g = c_oracle.backward(full["grad_out"], full["value"], full["shapes"], full["lsi"], full["loc"], full["aw"], "f64")
    got = [np.concatenate([o[i] for o in outs], 0) for i in range(4)]
    for g, r in zip(got, (ref_out,) + tuple(ref_g)):
        assert g.shape == r.shape and np.array_equal(g, r)     # bit-identical: nothing crosses samples


def test_aggregate_throughput_is_all_units_over_slowest_rank():
    assert sharding.aggregate_throughput(100, 8, 2.0) == 400.0


def test_reference_arm_under_two_ranks_only_rank0_prints():
    """bench.py --impl reference under torchrun: rank 0 alone runs and prints; others exit 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", LOCAL_RANK="1", RANK="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env.update(RANK="0", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1", "--workload", "c3_arctic_decoder"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    # "reference" when __graft_entry__.build() staged the unmodified reference function (oracle/_ref/), else the port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["steps"] == 1 and line["warmup"] == 1
