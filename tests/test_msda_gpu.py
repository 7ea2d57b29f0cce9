"""GPU parity tests (run with -m gpu on the B200 box).  Everything here calls the product through
the C ABI (ctypes on libmsda_b200.so) or through the reference-facing extension / autograd
Function, and scores it against the CPU oracle (oracle/) or the committed golden vectors.

Tolerances (BASELINE.json north_star): integer work bit-exact; fp32 forward rtol 1e-5; fp32
gradients rtol 1e-4 (atomic ordering); bf16 rtol 1e-2; with atol = rtol * max|reference|.
"""
import glob
import os

import ctypes

import numpy as np
import pytest
import torch

from oracle import c_oracle
from uvhand_b200 import cabi
from uvhand_b200.workloads import WORKLOADS, Workload, make_inputs

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]

FWD_RTOL, GRAD_RTOL, BF16_RTOL = 1e-5, 1e-4, 1e-2


def close(got, want, rtol):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    atol = rtol * max(float(np.abs(want).max()), 1e-30)
    bad = np.abs(got - want) > atol + rtol * np.abs(want)
    if bad.any():
        i = np.unravel_index(np.argmax(np.abs(got - want)), got.shape)
        raise AssertionError(f"{bad.sum()} / {bad.size} outside rtol={rtol}: worst at {i}: got {got[i]!r} want {want[i]!r}")


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


class Paths:
    """Context manager forcing a kernel family.  fwd/bwd: 0 auto, 1 any-shape kernels;
    split: 0 auto (by problem size), 1 never (4 items per warp), 2 always (1 item per warp);
    tiles: the on-chip-combining backward (L = P = 4): 0 by shape, 1 never, 2 whenever the shape is eligible;
    wide: the 8-channels-per-lane forward (L = P = 4): 0 / 2 whenever eligible and not split, 1 never;
    slab: the shared-memory slab forward (fp32, L = P = 4, S <= 1600): 0 by shape, 1 never, 2 whenever eligible."""

    def __init__(self, fwd=0, bwd=0, split=0, tiles=1, wide=1, slab=1):
        self.f, self.b, self.s, self.t, self.w, self.sl = fwd, bwd, split, tiles, wide, slab

    def __enter__(self):
        cabi.set_option("fwd_path", self.f)
        cabi.set_option("bwd_path", self.b)
        cabi.set_option("split_mode", self.s)
        cabi.set_option("bwd_tiles", self.t)
        cabi.set_option("fwd_wide", self.w)
        cabi.set_option("fwd_slab", self.sl)

    def __exit__(self, *a):
        cabi.set_option("fwd_path", 0)
        cabi.set_option("bwd_path", 0)
        cabi.set_option("split_mode", 0)
        cabi.set_option("bwd_tiles", 0)
        cabi.set_option("fwd_wide", 0)
        cabi.set_option("fwd_slab", 0)


# the kernel families every parity case runs through (D = 32 shapes; "tiles" applies to L = P = 4 only)
FAMILIES = {"d32": (0, 0, 1, 1, 1), "split": (0, 0, 2, 1, 1), "tiles": (0, 0, 1, 2, 2), "slab": (0, 0, 1, 1, 1, 2),
            "anyshape": (1, 1, 0, 1, 1)}
FAMILY_IDS = list(FAMILIES)


def run_cabi(kind, value, shapes, lsi, loc, aw, grad_out=None, bf16_direct=False):
    """numpy in -> numpy out through msda_fwd_* / msda_bwd_* with device buffers."""
    tdt = {"f32": torch.float32, "f64": torch.float64, "bf16": torch.float32}[kind]
    vdt = torch.bfloat16 if kind == "bf16" else tdt
    v = dev(value, vdt)
    sh, ls = dev(np.asarray(shapes, np.int64)), dev(np.asarray(lsi, np.int64))
    lo, a = dev(loc, tdt), dev(aw, tdt)
    N, S, M, D = v.shape
    _, Lq, _, L, P, _ = lo.shape
    dims = (N, S, M, D, L, Lq, P)
    out = torch.full((N, Lq, M * D), float("nan"), device="cuda", dtype=tdt)
    st = torch.cuda.current_stream().cuda_stream
    cabi.forward(kind, v.data_ptr(), sh.data_ptr(), ls.data_ptr(), lo.data_ptr(), a.data_ptr(), dims,
                 out.data_ptr(), st)
    res = {"out": out.cpu().numpy(), "fwd_path": cabi.last_path(False)}
    if grad_out is not None:
        g = dev(grad_out, tdt).reshape(N, Lq, M * D).contiguous()
        gv = torch.zeros_like(v)
        gl = torch.full(lo.shape, float("nan"), device="cuda", dtype=tdt)
        ga = torch.full(a.shape, float("nan"), device="cuda", dtype=tdt)
        scratch = None
        if kind == "bf16" and not bf16_direct:
            scratch = torch.zeros(v.shape, device="cuda", dtype=torch.float32)
            gv.fill_(float("nan"))          # accurate mode fully overwrites grad_value
        cabi.backward(kind, g.data_ptr(), v.data_ptr(), sh.data_ptr(), ls.data_ptr(), lo.data_ptr(),
                      a.data_ptr(), dims, gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), st,
                      bf16_scratch_ptr=scratch.data_ptr() if scratch is not None else None)
        res.update(grad_value=gv.float().cpu().numpy() if kind == "bf16" else gv.cpu().numpy(),
                   grad_loc=gl.cpu().numpy(), grad_aw=ga.cpu().numpy(), bwd_path=cabi.last_path(True))
    torch.cuda.synchronize()
    return res


def oracle(value, shapes, lsi, loc, aw, grad_out, variant="f32x"):
    out = c_oracle.forward(value, shapes, lsi, loc, aw, variant)
    gv, gl, ga = c_oracle.backward(grad_out, value, shapes, lsi, loc, aw, variant)
    return {"out": out, "grad_value": gv, "grad_loc": gl, "grad_aw": ga}


def check_all(res, ref, fwd_rtol=FWD_RTOL, grad_rtol=GRAD_RTOL):
    close(res["out"], ref["out"], fwd_rtol)
    close(res["grad_value"], ref["grad_value"], grad_rtol)
    close(res["grad_loc"], ref["grad_loc"], grad_rtol)
    close(res["grad_aw"], ref["grad_aw"], grad_rtol)


def np_inputs(wl, seed, **kw):
    t = make_inputs(wl, seed=seed, device="cpu", **kw)
    return {k: v.numpy() for k, v in t.items()}


# ------------------------------------------------------------------------------------------------
# golden vectors written by the unmodified reference (fp64 grid_sample path + autograd)
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("forced", FAMILY_IDS)
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_golden_f32(path, forced):
    g = np.load(path)
    with Paths(*FAMILIES[forced]):
        res = run_cabi("f32", g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], g["grad_out"])
    check_all(res, g)
    if g["value"].shape[3] == 32:
        l4p4 = g["loc"].shape[3] == 4 and g["loc"].shape[4] == 4
        want = {"d32": ("fwd_d32_f32", "bwd_d32_f32"), "split": ("fwd_d32_split_f32", "bwd_d32_split_f32"),
                "tiles": ("fwd_d32_wide_f32" if l4p4 else "fwd_d32_f32", "bwd_d32_tiles_f32" if l4p4 else "bwd_d32_f32"),
                "slab": ("fwd_d32_slab_f32" if l4p4 and g["value"].shape[1] <= 1600 else "fwd_d32_f32", "bwd_d32_f32"),
                "anyshape": ("fwd_generic_f32", "bwd_generic_f32")}[forced]
        assert (res["fwd_path"], res["bwd_path"]) == want


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_golden_f64(path):
    g = np.load(path)
    res = run_cabi("f64", g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], g["grad_out"])
    check_all(res, g, 1e-11, 1e-10)


def test_reference_test_anchor():
    # models/ops/test.py seed-3 forward, 4 d.p.
    g = np.load(GOLDEN[IDS.index("testpy_n1")])
    res = run_cabi("f32", g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"])
    want = np.array([[0.0019, 0.0046, 0.0047, 0.0044], [0.0038, 0.0025, 0.0018, 0.0036]])
    assert np.allclose(np.round(res["out"][0], 4), want, atol=1e-4)


# ------------------------------------------------------------------------------------------------
# integer work: bit-exact against the oracle
# ------------------------------------------------------------------------------------------------

def _index_gpu(shapes, loc):
    sh, lo = dev(np.asarray(shapes, np.int64)), dev(loc, torch.float32)
    N, Lq, M, L, P, _ = lo.shape
    h = torch.full((N, Lq, M, L, P), -7, device="cuda", dtype=torch.int32)
    w = torch.full_like(h, -7)
    k = torch.full((N, Lq, M, L, P), 255, device="cuda", dtype=torch.uint8)
    cabi.sample_index(sh.data_ptr(), lo.data_ptr(), N, Lq, M, L, P, h.data_ptr(), w.data_ptr(), k.data_ptr(),
                      torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return h.cpu().numpy(), w.cpu().numpy(), k.cpu().numpy()


@pytest.mark.parametrize("levels", [WORKLOADS["c2_detr_encoder"].levels, WORKLOADS["c3_arctic_encoder"].levels,
                                    WORKLOADS["c5_hires"].levels, ((6, 4), (3, 2)), ((1, 1), (1, 7), (9, 1))],
                         ids=["c2", "c3", "c5", "c1", "degenerate"])
def test_sample_index_bit_exact(levels):
    rng = np.random.default_rng(42)
    shapes = np.array(levels, dtype=np.int64)
    L = len(levels)
    parts = [rng.random((1, 4000, 2, L, 4, 2)).astype(np.float32),                    # U[0,1)
             (rng.random((1, 4000, 2, L, 4, 2)) * 3 - 1).astype(np.float32)]          # U[-1,2)
    # exact pixel centres, exact pixel edges, half a pixel outside, for every level
    edge = np.zeros((1, 2048, 2, L, 4, 2), dtype=np.float32)
    for l, (H, W) in enumerate(levels):
        k = np.arange(2048 * 2 * 4, dtype=np.float64).reshape(2048, 2, 4)
        cx = ((k % (W + 2)) - 1 + 0.5) / W
        cy = (((k // 3) % (H + 2)) - 1 + 0.5) / H
        ex = (k % (W + 1)) / W
        ey = ((k // 5) % (H + 1)) / H
        pick = (k.astype(np.int64) % 2 == 0)
        edge[0, :, :, l, :, 0] = np.where(pick, cx, ex)
        edge[0, :, :, l, :, 1] = np.where(pick, cy, ey)
    parts.append(edge)
    special = rng.random((1, 64, 2, L, 4, 2)).astype(np.float32)
    special[0, 0::4, :, :, 0, 0] = np.nan
    special[0, 1::4, :, :, 1, 1] = np.inf
    special[0, 2::4, :, :, 2, 0] = -np.inf
    special[0, 3::4, :, :, 3, 1] = 1e30
    parts.append(special)
    for loc in parts:
        h, w, k = _index_gpu(shapes, loc)
        h0, w0, k0 = c_oracle.sample_index(shapes, loc, "f32")
        assert np.array_equal(k, k0)
        assert np.array_equal(h, h0)
        assert np.array_equal(w, w0)


@pytest.mark.parametrize("forced", FAMILY_IDS)
def test_grad_value_touch_pattern_exact(forced):
    """Which value rows receive gradient is pure index/mask work: the set must match exactly."""
    wl = Workload("t", 2, ((9, 13), (5, 7), (3, 4), (2, 2)), 60, loc_dist="twostage")
    x = np_inputs(wl, 7, pad_frac=0.0)
    x["grad_out"] = np.ones_like(x["grad_out"])
    with Paths(*FAMILIES[forced]):
        res = run_cabi("f32", x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    ref = oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    assert np.array_equal(res["grad_value"] != 0, ref["grad_value"] != 0)
    # rejected samples give exactly zero location / weight gradients
    _, _, k = c_oracle.sample_index(x["shapes"], x["loc"], "f32")
    assert (k == 0).any()
    assert np.all(res["grad_aw"][k == 0] == 0) and np.all(res["grad_loc"][k == 0] == 0)


# ------------------------------------------------------------------------------------------------
# seeded workloads at sizes the oracle finishes in seconds
# ------------------------------------------------------------------------------------------------

MID_LEVELS = ((40, 57), (20, 29), (10, 15), (5, 8))
MID_S = sum(h * w for h, w in MID_LEVELS)


@pytest.mark.parametrize("forced", FAMILY_IDS)
@pytest.mark.parametrize("dist,Lq", [("encoder", MID_S), ("uniform", 777), ("twostage", 300), ("contended", 512)])
def test_mid_size_f32(dist, Lq, forced):
    wl = Workload("mid", 2, MID_LEVELS, Lq, loc_dist=dist)
    x = np_inputs(wl, 3)
    with Paths(*FAMILIES[forced]):
        res = run_cabi("f32", x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    check_all(res, oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"]))


@pytest.mark.parametrize("split", [1, 2, 3], ids=["d32", "split", "wide+tiles"])
@pytest.mark.parametrize("direct", [False, True], ids=["f32acc", "direct-bf16x2"])
@pytest.mark.parametrize("dist,Lq", [("encoder", MID_S), ("twostage", 300)])
def test_mid_size_bf16(dist, Lq, direct, split):
    wl = Workload("mid", 2, MID_LEVELS, Lq, loc_dist=dist)
    x = np_inputs(wl, 4)
    v_bf16 = torch.from_numpy(x["value"]).to(torch.bfloat16)
    v_round = v_bf16.float().numpy()                      # what the kernel actually reads
    with Paths(0, 0, min(split, 2), 1, 1) if split < 3 else Paths(0, 0, 1, 2, 2):
        res = run_cabi("bf16", v_round, x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], bf16_direct=direct)
    ref = oracle(v_round, x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    assert res["fwd_path"] == {1: "fwd_d32_bf16", 2: "fwd_d32_split_bf16", 3: "fwd_d32_wide_bf16"}[split]
    tl = "tiles_" if split == 3 else ""
    assert res["bwd_path"] == (f"bwd_d32_{tl}bf16" if direct else f"bwd_d32_{tl}bf16_f32acc")
    close(res["out"], ref["out"], FWD_RTOL)               # same rounded inputs -> fp32-level agreement
    close(res["grad_loc"], ref["grad_loc"], GRAD_RTOL)
    close(res["grad_aw"], ref["grad_aw"], GRAD_RTOL)
    gv, gv_ref = res["grad_value"].astype(np.float64), ref["grad_value"]
    l2 = np.linalg.norm(gv - gv_ref) / np.linalg.norm(gv_ref)
    worst = np.abs(gv - gv_ref).max() / np.abs(gv_ref).max()
    print(f"bf16 grad_value {dist} direct={direct}: rel-L2 {l2:.3e}, worst/max {worst:.3e}")
    if not direct:
        # default bf16 mode: fp32 accumulation, ONE rounding to bf16 at the end -> the stated 1e-2
        close(res["grad_value"], ref["grad_value"], BF16_RTOL)
    else:
        # direct red.global.add.noftz.bf16x2: every add rounds the running sum to 8 bits, heavily hit
        # rows drift; documented looser bound (include/msda_b200.h), not the north_star tolerance.
        assert l2 < 5e-2 and worst < 0.25, (l2, worst)
    # and against the unrounded value: the stated bf16 tolerance
    ref_full = oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    close(res["out"], ref_full["out"], BF16_RTOL)


SHAPES = [
    # (N, M, D, L, P, Lq, levels)                 what it exercises
    (1, 1, 1, 1, 1, 1, ((1, 1),)),                # smallest possible
    (3, 3, 32, 4, 4, 5, ((7, 5), (4, 3), (2, 2), (1, 1))),     # M % 4 != 0: a warp spans queries
    (2, 8, 32, 3, 2, 9, ((6, 5), (3, 3), (2, 1))),             # L*P = 6: runtime-LP specialised kernel
    (1, 2, 32, 8, 4, 6, tuple((3, 2) for _ in range(8))),      # L*P = 32: record-buffer limit
    (1, 2, 32, 5, 7, 6, tuple((3, 2) for _ in range(5))),      # L*P = 35: falls back to any-shape
    (1, 2, 32, 9, 1, 6, tuple((2, 2) for _ in range(9))),      # L = 9 > 8: any-shape
    (2, 2, 30, 2, 2, 3, ((6, 4), (3, 2))),        # test.py:85 channel counts
    (2, 2, 33, 2, 2, 3, ((6, 4), (3, 2))),
    (1, 2, 64, 2, 2, 3, ((6, 4), (3, 2))),
    (1, 2, 71, 2, 2, 3, ((6, 4), (3, 2))),
    (1, 2, 1025, 2, 2, 2, ((6, 4), (3, 2))),
    (1, 1, 2048, 2, 2, 2, ((6, 4), (3, 2))),
    (1, 1, 3096, 2, 2, 2, ((6, 4), (3, 2))),
    (5, 4, 32, 4, 4, 1, ((3, 3), (2, 2), (1, 2), (1, 1))),     # one query
]


@pytest.mark.parametrize("shape", SHAPES, ids=[f"N{s[0]}M{s[1]}D{s[2]}L{s[3]}P{s[4]}Lq{s[5]}" for s in SHAPES])
@pytest.mark.parametrize("kind", ["f32", "f32-split", "f32-tiles", "f64"])
def test_shape_sweep(shape, kind):
    N, M, D, L, P, Lq, levels = shape
    wl = Workload("s", N, levels, Lq, M=M, D=D, P=P, loc_dist="twostage")
    x = np_inputs(wl, 9, pad_frac=0.0, offset_px=1.0)
    x["loc"] = (x["loc"] * 0.4 + 0.5).astype(np.float32)   # mostly inside, some out
    split = 2 if kind == "f32-split" else 1                # 4-items-per-warp vs 1-item-per-warp kernels
    tiles = 2 if kind == "f32-tiles" else 1                # on-chip-combining backward (only L = P = 4 takes it)
    kind = kind.split("-")[0]
    with Paths(0, 0, split, tiles, tiles):
        res = run_cabi(kind, x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    variant = "f32x" if kind == "f32" else "f64"
    ref = oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], variant)
    if kind == "f32":
        check_all(res, ref)
    else:
        check_all(res, ref, 1e-12, 1e-11)


@pytest.mark.parametrize("seed", range(int(os.environ.get("MSDA_FUZZ_SEEDS", "12"))))
def test_random_geometry_fuzz(seed):
    """Seeded random problems (ragged level sizes incl. 1xW / Hx1 / 1x1, any M, L, P with D = 32, locations
    from well outside to well inside, a sprinkle of NaN / inf / exact pixel centres / exact borders) through
    all three D=32-capable kernel families, against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    L = int(rng.integers(1, 9))
    P = int(rng.integers(1, 32 // L + 1))
    M = int(rng.integers(1, 10))
    N = int(rng.integers(1, 4))
    Lq = int(rng.integers(1, 70))
    levels = tuple((int(rng.integers(1, 12)), int(rng.integers(1, 12))) for _ in range(L))
    wl = Workload("fz", N, levels, Lq, M=M, D=32, P=P, loc_dist="uniform")
    x = np_inputs(wl, 50 + seed, pad_frac=0.0)
    loc = (rng.random(x["loc"].shape) * 1.6 - 0.3).astype(np.float32)
    wh = np.array([[w, h] for h, w in levels], np.float32)[None, None, None, :, None, :]
    flat = loc.reshape(-1, 2)
    k = flat.shape[0]
    pick = rng.choice(k, size=max(1, k // 6), replace=False)
    kinds = rng.integers(0, 6, size=pick.size)
    lvl_of = np.broadcast_to(np.arange(L)[None, None, None, :, None], loc.shape[:-1]).reshape(-1)
    for i, kind in zip(pick, kinds):
        w_, h_ = wh.reshape(-1, 2)[lvl_of[i]]
        if kind == 0:
            flat[i] = [np.nan, 0.5]
        elif kind == 1:
            flat[i] = [0.5, np.inf if rng.random() < 0.5 else -np.inf]
        elif kind == 2:   # exact pixel centres
            flat[i] = [(rng.integers(0, int(w_)) + 0.5) / w_, (rng.integers(0, int(h_)) + 0.5) / h_]
        elif kind == 3:   # exact pixel edges, including the outer borders 0 and 1
            flat[i] = [rng.integers(0, int(w_) + 1) / w_, rng.integers(0, int(h_) + 1) / h_]
        elif kind == 4:   # half a pixel outside: the last accepted / first rejected coordinates
            flat[i] = [-0.5 / w_, 1.0 + 0.5 / h_]
        else:
            flat[i] = [1.0 + 0.4999 / w_, -0.4999 / h_]
    x["loc"] = loc
    ref = oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    h0, w0, k0 = c_oracle.sample_index(x["shapes"], x["loc"], "f32")
    for fam in FAMILY_IDS:
        with Paths(*FAMILIES[fam]):
            res = run_cabi("f32", x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
        for key in ("out", "grad_value", "grad_loc", "grad_aw"):
            assert np.isfinite(res[key]).all(), (fam, key)
        check_all(res, ref)
        assert np.all(res["grad_loc"][k0 == 0] == 0) and np.all(res["grad_aw"][k0 == 0] == 0), fam


def test_special_locations_contribute_zero():
    wl = Workload("s", 1, ((4, 5), (2, 3)), 8, M=8, D=32, P=4, loc_dist="uniform")
    x = np_inputs(wl, 1)
    loc = x["loc"]
    loc[0, 0, :, 0, 0] = [np.nan, 0.5]
    loc[0, 1, :, 1, 1] = [0.5, np.inf]
    loc[0, 2, :, 0, 2] = [-np.inf, np.nan]
    loc[0, 3, :, 1, 3] = [7.0, 0.5]
    for forced in FAMILY_IDS:
        with Paths(*FAMILIES[forced]):
            res = run_cabi("f32", x["value"], x["shapes"], x["lsi"], loc, x["aw"], x["grad_out"])
        ref = oracle(x["value"], x["shapes"], x["lsi"], loc, x["aw"], x["grad_out"])
        for k in ("out", "grad_value", "grad_loc", "grad_aw"):
            assert np.isfinite(res[k]).all(), k
        check_all(res, ref)
        assert np.all(res["grad_loc"][0, 0, :, 0, 0] == 0) and np.all(res["grad_aw"][0, 1, :, 1, 1] == 0)


def test_invalid_corners_are_never_read():
    """Zero padding means 'do not load', not 'load and multiply by zero': an inf in a row that is
    only ever an out-of-range neighbour must not reach the output (cuh:55-78)."""
    levels = ((3, 3),)
    wl = Workload("s", 1, levels, 4, M=8, D=32, P=4, loc_dist="uniform")
    x = np_inputs(wl, 2, pad_frac=0.0)
    # all samples in the top-left pixel's upper-left quadrant: h_low = w_low = -1, only corner v4 = (0,0) valid
    x["loc"][...] = np.float32(0.05)
    x["value"][0, 1:] = np.inf                      # every row except pixel (0,0)
    for forced in FAMILY_IDS:
        with Paths(*FAMILIES[forced]):
            res = run_cabi("f32", x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
        assert np.isfinite(res["out"]).all()
        assert np.isfinite(res["grad_loc"]).all() and np.isfinite(res["grad_aw"]).all()
        assert np.all(res["grad_value"][0, 1:] == 0)


# ------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties + oracle on a slice
# ------------------------------------------------------------------------------------------------

def _full_size(wl, kind="f32", seed=0):
    t = make_inputs(wl, seed=seed, device="cuda", value_dtype=torch.bfloat16 if kind == "bf16" else torch.float32)
    dims = wl.dims()
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty(wl.N, wl.Lq, wl.M * wl.D, device="cuda")
    cabi.forward(kind, t["value"].data_ptr(), t["shapes"].data_ptr(), t["lsi"].data_ptr(), t["loc"].data_ptr(),
                 t["aw"].data_ptr(), dims, out.data_ptr(), st)
    gv = torch.zeros_like(t["value"])
    gl = torch.empty_like(t["loc"])
    ga = torch.empty_like(t["aw"])
    scratch = torch.zeros(t["value"].shape, device="cuda") if kind == "bf16" else None
    cabi.backward(kind, t["grad_out"].data_ptr(), t["value"].data_ptr(), t["shapes"].data_ptr(), t["lsi"].data_ptr(),
                  t["loc"].data_ptr(), t["aw"].data_ptr(), dims, gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), st,
                  bf16_scratch_ptr=scratch.data_ptr() if scratch is not None else None)
    torch.cuda.synchronize()
    return t, out, gv, gl, ga


@pytest.mark.parametrize("name", ["c2_detr_encoder", "c3_arctic_encoder", "c3_arctic_decoder", "c4_dino_decoder",
                                  "c4_dino_decoder_train"])
def test_full_size_properties_f32(name):
    wl = WORKLOADS[name]
    t, out, gv, gl, ga = _full_size(wl)
    # the 900/1098-query decoders are small enough for the one-item-per-warp forward; their queries have no
    # relation to the (large) pyramid, so the backward reduces straight to L2.  Encoder self-attention and the
    # small ARCTIC pyramid take the on-chip-combining backward.
    # The small ARCTIC pyramid is gathered from a shared-memory slab in the forward.
    want = ("fwd_d32_split_f32" if name.startswith("c4") else "fwd_d32_slab_f32" if name.startswith("c3") else "fwd_d32_f32",
            "bwd_d32_f32" if name.startswith("c4") else "bwd_d32_tiles_f32")
    assert (cabi.last_path(False), cabi.last_path(True)) == want
    g = t["grad_out"].view_as(out).double()
    # forward is linear in value and in the weights: <out, g> = <value, grad_value> = <w, grad_w>
    lhs = float((out.double() * g).sum())
    rhs_v = float((t["value"].double() * gv.double()).sum())
    rhs_w = float((t["aw"].double() * ga.double()).sum())
    scale = float((out.double() * g).abs().sum())
    assert abs(lhs - rhs_v) < 1e-5 * scale
    assert abs(lhs - rhs_w) < 1e-5 * scale
    # oracle on a slice of queries (full value): forward, grad_loc, grad_w for those queries
    gen = torch.Generator().manual_seed(1)
    idx = torch.randperm(wl.Lq, generator=gen)[:96].sort().values
    cpu = {k: v.cpu() for k, v in t.items()}
    sub_loc = cpu["loc"][:, idx].contiguous().numpy()
    sub_aw = cpu["aw"][:, idx].contiguous().numpy()
    sub_g = cpu["grad_out"][:, idx].contiguous().numpy()
    ref = oracle(cpu["value"].numpy(), cpu["shapes"].numpy(), cpu["lsi"].numpy(), sub_loc, sub_aw, sub_g)
    close(out.cpu()[:, idx].numpy(), ref["out"], FWD_RTOL)
    close(gl.cpu()[:, idx].numpy(), ref["grad_loc"], GRAD_RTOL)
    close(ga.cpu()[:, idx].numpy(), ref["grad_aw"], GRAD_RTOL)
    # index decisions for the whole tensor, bit-exact
    h, w, k = _index_gpu(cpu["shapes"].numpy(), cpu["loc"].numpy()[:1, :4096])
    h0, w0, k0 = c_oracle.sample_index(cpu["shapes"].numpy(), cpu["loc"].numpy()[:1, :4096], "f32")
    assert np.array_equal(h, h0) and np.array_equal(w, w0) and np.array_equal(k, k0)


def test_full_size_c2_grad_value_vs_oracle():
    """C2 is small enough (5.7 M samples) for the OpenMP oracle to do the whole backward."""
    wl = WORKLOADS["c2_detr_encoder"]
    t, out, gv, gl, ga = _full_size(wl, seed=2)
    cpu = {k: v.cpu().numpy() for k, v in t.items()}
    ref = oracle(cpu["value"], cpu["shapes"], cpu["lsi"], cpu["loc"], cpu["aw"], cpu["grad_out"])
    close(out.cpu().numpy(), ref["out"], FWD_RTOL)
    close(gv.cpu().numpy(), ref["grad_value"], GRAD_RTOL)
    close(gl.cpu().numpy(), ref["grad_loc"], GRAD_RTOL)
    close(ga.cpu().numpy(), ref["grad_aw"], GRAD_RTOL)


def test_arctic_eval_batch_of_2048_frames_equals_the_reference_chunk_loop():
    """ARCTIC evaluation hands the operator N = 64 clips x 32 frames = 2048 samples (README.md:173-174); the
    reference walks them in 32 chunks of im2col_step = 64 with pointer offsets (ms_deform_attn_cuda.cu:56-75,
    131-148).  One launch over the whole batch (9 GB of operands, 64-bit addressing) must give the very same
    bits as that chunk loop run through the same kernels."""
    from uvhand_b200.workloads import ARCTIC_LEVELS
    wl = Workload("arctic_eval", 2048, ARCTIC_LEVELS, sum(h * w for h, w in ARCTIC_LEVELS))
    t = make_inputs(wl, seed=21, device="cuda")
    N, S, M, D, L, Lq, P = wl.dims()
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty(N, Lq, M * D, device="cuda")
    gl, ga = torch.empty_like(t["loc"]), torch.empty_like(t["aw"])
    gv = torch.zeros_like(t["value"])
    cabi.forward("f32", t["value"].data_ptr(), t["shapes"].data_ptr(), t["lsi"].data_ptr(), t["loc"].data_ptr(),
                 t["aw"].data_ptr(), wl.dims(), out.data_ptr(), st)
    cabi.backward("f32", t["grad_out"].data_ptr(), t["value"].data_ptr(), t["shapes"].data_ptr(), t["lsi"].data_ptr(),
                  t["loc"].data_ptr(), t["aw"].data_ptr(), wl.dims(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), st)
    out_c, gl_c, ga_c = torch.empty_like(out), torch.empty_like(gl), torch.empty_like(ga)
    gv_c = torch.zeros_like(gv)
    step = 64
    assert cabi.lib().msda_b200_check_im2col_step(N, step) == 0
    with Paths(0, 0, 1, 0, 0):    # chunks of 64 are small enough to pick the split forward; keep the kernels identical
        for b0 in range(0, N, step):
            sl = slice(b0, b0 + step)
            dims = (step, S, M, D, L, Lq, P)
            cabi.forward("f32", t["value"][sl].data_ptr(), t["shapes"].data_ptr(), t["lsi"].data_ptr(),
                         t["loc"][sl].data_ptr(), t["aw"][sl].data_ptr(), dims, out_c[sl].data_ptr(), st)
            cabi.backward("f32", t["grad_out"][sl].data_ptr(), t["value"][sl].data_ptr(), t["shapes"].data_ptr(),
                          t["lsi"].data_ptr(), t["loc"][sl].data_ptr(), t["aw"][sl].data_ptr(), dims,
                          gv_c[sl].data_ptr(), gl_c[sl].data_ptr(), ga_c[sl].data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(out, out_c) and torch.equal(gl, gl_c) and torch.equal(ga, ga_c)
    # grad_value: the same additions per row, in an order the atomics decide
    assert float((gv - gv_c).abs().max()) <= GRAD_RTOL * float(gv_c.abs().max())
    # and the last frame against the oracle (addresses beyond 2^31 elements are exercised by its rows)
    cpu = {k: v[-1:].cpu().numpy() if k in ("value", "loc", "aw", "grad_out") else v.cpu().numpy() for k, v in t.items()}
    ref = oracle(cpu["value"], cpu["shapes"], cpu["lsi"], cpu["loc"], cpu["aw"], cpu["grad_out"])
    close(out[-1:].cpu().numpy(), ref["out"], FWD_RTOL)
    close(gv[-1:].cpu().numpy(), ref["grad_value"], GRAD_RTOL)
    close(gl[-1:].cpu().numpy(), ref["grad_loc"], GRAD_RTOL)


def test_full_size_c5_bf16_properties():
    wl = WORKLOADS["c5_hires"]
    t, out, gv, gl, ga = _full_size(wl, kind="bf16")
    assert cabi.last_path(False) == "fwd_d32_wide_bf16" and cabi.last_path(True) == "bwd_d32_tiles_bf16_f32acc"
    g = t["grad_out"].view_as(out).double()
    lhs = float((out.double() * g).sum())
    rhs_w = float((t["aw"].double() * ga.double()).sum())
    scale = float((out.double() * g).abs().sum())
    assert abs(lhs - rhs_w) < 1e-5 * scale
    rhs_v = float((t["value"].double() * gv.double()).sum())        # grad_value rounded once to bf16
    assert abs(lhs - rhs_v) < 1e-3 * scale
    gen = torch.Generator().manual_seed(1)
    idx = torch.randperm(wl.Lq, generator=gen)[:64].sort().values
    cpu = {k: v.cpu() for k, v in t.items()}
    v1 = cpu["value"][:1].float().numpy()
    ref_out = c_oracle.forward(v1, cpu["shapes"].numpy(), cpu["lsi"].numpy(),
                               cpu["loc"][:1, idx].contiguous().numpy(), cpu["aw"][:1, idx].contiguous().numpy(), "f32x")
    close(out.cpu()[:1, idx].numpy(), ref_out, FWD_RTOL)


# ------------------------------------------------------------------------------------------------
# the reference-facing layers: extension module, autograd Function, nn.Module
# ------------------------------------------------------------------------------------------------

def test_function_matches_oracle_and_is_chunk_invariant():
    from uvhand_b200.functions import MSDeformAttnFunction
    wl = Workload("f", 6, ((9, 13), (5, 7), (3, 4), (2, 2)), 50, loc_dist="twostage")
    x = make_inputs(wl, seed=21, device="cpu")
    ref = oracle(*[x[k].numpy() for k in ("value", "shapes", "lsi", "loc", "aw", "grad_out")])
    outs = []
    for step in (64, 6, 3, 2, 1):
        v = x["value"].cuda().requires_grad_(True)
        lo = x["loc"].cuda().requires_grad_(True)
        a = x["aw"].cuda().requires_grad_(True)
        out = MSDeformAttnFunction.apply(v, x["shapes"].cuda(), x["lsi"].cuda(), lo, a, step)
        assert out.shape == (wl.N, wl.Lq, wl.M * wl.D) and out.is_contiguous()
        out.backward(x["grad_out"].cuda())
        close(out.detach().cpu().numpy(), ref["out"], FWD_RTOL)
        close(v.grad.cpu().numpy(), ref["grad_value"], GRAD_RTOL)
        close(lo.grad.cpu().numpy(), ref["grad_loc"], GRAD_RTOL)
        close(a.grad.cpu().numpy(), ref["grad_aw"], GRAD_RTOL)
        outs.append(out.detach())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])       # forward is deterministic and independent of im2col_step
    # ms_deform_attn_cuda.cu:52: batch 6, im2col_step 4 -> error
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        MSDeformAttnFunction.apply(x["value"].cuda(), x["shapes"].cuda(), x["lsi"].cuda(), x["loc"].cuda(),
                                   x["aw"].cuda(), 4)


def test_function_dtype_policy():
    from uvhand_b200.functions import MSDeformAttnFunction
    from uvhand_b200.functions import ms_deform_attn_func as F
    wl = Workload("f", 2, ((6, 4), (3, 2)), 5, M=8, D=32, P=4, loc_dist="uniform")
    x = make_inputs(wl, seed=22, device="cuda")
    args = (x["shapes"], x["lsi"], x["loc"], x["aw"], 64)
    base = MSDeformAttnFunction.apply(x["value"], *args)
    # models/ops casts every value to float32 (func.py:26): half / bf16 values behave that way by default
    v16 = x["value"].half()
    out16 = MSDeformAttnFunction.apply(v16, *args)
    assert out16.dtype == torch.float32
    assert torch.allclose(out16, MSDeformAttnFunction.apply(v16.float(), *args), rtol=0, atol=0)
    vb = x["value"].bfloat16().requires_grad_(True)
    assert not F.get_bf16_value_mode()
    out_b = MSDeformAttnFunction.apply(vb, *args)
    assert cabi.last_path(False) in ("fwd_d32_f32", "fwd_d32_split_f32")
    F.set_bf16_value_mode(True)
    try:
        out_n = MSDeformAttnFunction.apply(vb, *args)
        assert cabi.last_path(False) in ("fwd_d32_bf16", "fwd_d32_split_bf16")
        assert torch.allclose(out_n, out_b, rtol=1e-5, atol=1e-5)
        out_n.backward(x["grad_out"])
        assert vb.grad.dtype == torch.bfloat16 and cabi.last_path(True) == "bwd_d32_bf16_f32acc"
    finally:
        F.set_bf16_value_mode(False)
    assert torch.allclose(out_b, base, rtol=2e-2, atol=2e-2)
    # float64 all the way (what models/dino/ops permits and test.py's double checks need)
    out64 = MSDeformAttnFunction.apply(x["value"].double(), x["shapes"], x["lsi"], x["loc"].double(),
                                       x["aw"].double(), 64)
    assert out64.dtype == torch.float64 and torch.allclose(out64.float(), base, rtol=1e-5, atol=1e-6)


def test_extension_preconditions_raise_like_the_reference():
    import uvhand_b200
    M = uvhand_b200.load_extension()
    wl = Workload("f", 2, ((6, 4), (3, 2)), 5, M=2, D=4, P=2, loc_dist="uniform")
    x = make_inputs(wl, seed=23, device="cuda")
    v, sh, ls, lo, a = x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"]
    M.ms_deform_attn_forward(v, sh, ls, lo, a, 64)
    with pytest.raises(RuntimeError, match="value tensor has to be contiguous"):
        M.ms_deform_attn_forward(v.transpose(1, 2).contiguous().transpose(1, 2), sh, ls, lo, a, 64)
    with pytest.raises(RuntimeError, match="sampling_loc tensor has to be contiguous"):
        M.ms_deform_attn_forward(v, sh, ls, lo.flip(-1).transpose(1, 2).contiguous().transpose(1, 2), a, 64)
    with pytest.raises(RuntimeError, match="spatial_shapes must be a CUDA tensor"):
        M.ms_deform_attn_forward(v, sh.cpu(), ls, lo, a, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        M.ms_deform_attn_forward(v.cpu(), sh, ls, lo, a, 64)
    with pytest.raises(RuntimeError, match="dtype"):
        M.ms_deform_attn_forward(v, sh, ls, lo.double(), a, 64)
    with pytest.raises(RuntimeError, match="int64"):
        M.ms_deform_attn_forward(v, sh.int(), ls, lo, a, 64)
    with pytest.raises(RuntimeError, match="not implemented for"):
        M.ms_deform_attn_forward(v.half(), sh, ls, lo.half(), a.half(), 64)
    with pytest.raises(RuntimeError, match="grad_output tensor has to be contiguous"):
        g = torch.randn(2, 8, 5, device="cuda").transpose(1, 2)
        M.ms_deform_attn_backward(v, sh, ls, lo, a, g, 64)
    res = M.ms_deform_attn_backward(v, sh, ls, lo, a, torch.randn(2, 5, 8, device="cuda"), 64)
    assert isinstance(res, list) and [tuple(r.shape) for r in res] == [tuple(v.shape), tuple(lo.shape), tuple(a.shape)]


@pytest.mark.parametrize("channels", [30, 32, 64, 71])
def test_gradcheck_like_reference_test(channels):
    """models/ops/test.py:63-78: torch.autograd.gradcheck in float64 at the test's shape."""
    from uvhand_b200.functions import MSDeformAttnFunction
    torch.manual_seed(3)
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = 30
    value = (torch.rand(N, S, M, channels) * 0.01).cuda().double().requires_grad_(True)
    loc = torch.rand(N, Lq, M, L, P, 2).cuda().double().requires_grad_(True)
    aw = torch.rand(N, Lq, M, L, P).cuda() + 1e-5
    aw = (aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().requires_grad_(True)
    assert torch.autograd.gradcheck(MSDeformAttnFunction.apply, (value, shapes, lsi, loc, aw, 2))


@pytest.mark.parametrize("ref_dim", [2, 42, 4])
def test_module_against_grid_sample_port(ref_dim):
    """MSDeformAttn end to end (projections + op) against the same module with the op replaced by the
    port of the reference's CPU path."""
    from oracle.grid_sample_port import msda_grid_sample
    from uvhand_b200.modules import MSDeformAttn
    torch.manual_seed(5)
    levels = ((12, 17), (6, 9), (3, 5), (2, 3))
    shapes = torch.as_tensor(levels, dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    N, Lq = 2, 37
    mod = MSDeformAttn(256, 4, 8, 4).cuda()
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    query = torch.randn(N, Lq, 256, device="cuda", requires_grad=True)
    feat = torch.randn(N, S, 256, device="cuda", requires_grad=True)
    refp = torch.rand(N, Lq, 4, ref_dim, device="cuda")
    if ref_dim == 4:
        refp[..., 2:] *= 0.3
    pad = torch.zeros(N, S, dtype=torch.bool, device="cuda")
    pad[:, -5:] = True
    out = mod(query, refp, feat, shapes, lsi, pad)
    out.sum().backward()
    gq, gf = query.grad.clone(), feat.grad.clone()
    query.grad = feat.grad = None

    # same module arithmetic, op swapped for the grid_sample port
    value = mod.value_proj(feat).masked_fill(pad[..., None], 0.0).view(N, S, 8, 32)
    off = mod.sampling_offsets(query).view(N, Lq, 8, 4, 4, 2)
    w = torch.softmax(mod.attention_weights(query).view(N, Lq, 8, 16), -1).view(N, Lq, 8, 4, 4)
    loc = MSDeformAttn._sampling_locations(refp, off, shapes)
    out2 = mod.output_proj(msda_grid_sample(value, levels, loc, w))
    out2.sum().backward()
    assert torch.allclose(out, out2, rtol=1e-4, atol=1e-5)
    assert torch.allclose(gq, query.grad, rtol=1e-3, atol=1e-5)
    assert torch.allclose(gf, feat.grad, rtol=1e-3, atol=1e-5)


# ------------------------------------------------------------------------------------------------
# fused prologue: softmax + reference/offset arithmetic inside the kernels
# ------------------------------------------------------------------------------------------------

def _fused_inputs(N, M, L, P, Lq, levels, seed, ref_spread=1.0):
    """Raw module-side operands: reference points, pixel offsets, logits (+ value, grad_out)."""
    rng = np.random.default_rng(seed)
    S = sum(h * w for h, w in levels)
    shapes = np.asarray(levels, np.int64).reshape(-1, 2)
    lsi = np.concatenate([[0], np.cumsum(shapes.prod(1))[:-1]]).astype(np.int64)
    value = rng.standard_normal((N, S, M, 32)).astype(np.float32)
    ref = (0.5 + ref_spread * (rng.random((N, Lq, L, 2)) - 0.5)).astype(np.float32)
    off = (rng.standard_normal((N, Lq, M, L, P, 2)) * 2.0).astype(np.float32)
    logits = rng.standard_normal((N, Lq, M, L * P)).astype(np.float32)
    go = rng.standard_normal((N, Lq, M * 32)).astype(np.float32)
    return dict(value=value, shapes=shapes, lsi=lsi, ref=ref, off=off, logits=logits, grad_out=go)


def _fused_reference(x):
    """Unfused restatement: torch's fp32 location arithmetic in numpy fp32 (division then addition,
    each correctly rounded), softmax in fp64, then the C oracle and the chain rule by hand."""
    shapes = x["shapes"]
    wh = np.stack([shapes[:, 1], shapes[:, 0]], -1).astype(np.float32)            # (W, H) per level
    loc = (x["ref"][:, :, None, :, None, :] + x["off"] / wh[None, None, None, :, None, :]).astype(np.float32)
    lg = x["logits"].astype(np.float64)
    e = np.exp(lg - lg.max(-1, keepdims=True))
    a64 = e / e.sum(-1, keepdims=True)
    N, Lq, M, L, P, _ = x["off"].shape
    a = a64.astype(np.float32).reshape(N, Lq, M, L, P)
    out = c_oracle.forward(x["value"], shapes, x["lsi"], loc, a, "f32x")
    gv, gl, ga = c_oracle.backward(x["grad_out"], x["value"], shapes, x["lsi"], loc, a, "f32x")
    g_off = gl / wh[None, None, None, :, None, :].astype(np.float64)
    ga_f = ga.reshape(N, Lq, M, L * P)
    g_logits = a64 * (ga_f - (a64 * ga_f).sum(-1, keepdims=True))
    g_ref = gl.sum(axis=(2, 4))
    return loc, a, dict(out=out, grad_value=gv, grad_off=g_off, grad_logits=g_logits, grad_ref=g_ref)


def _run_fused(x, want_ref=True):
    t = {k: dev(v) for k, v in x.items()}
    N, S, M, D = t["value"].shape
    _, Lq, _, L, P, _ = t["off"].shape
    dims = (N, S, M, D, L, Lq, P)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.full((N, Lq, M * D), float("nan"), device="cuda")
    cabi.fused_forward(t["value"].data_ptr(), t["shapes"].data_ptr(), t["lsi"].data_ptr(), t["ref"].data_ptr(),
                       t["off"].data_ptr(), t["logits"].data_ptr(), dims, out.data_ptr(), st)
    fwd_path = cabi.last_path(False)
    gv = torch.zeros_like(t["value"])
    goff = torch.full_like(t["off"], float("nan"))
    glog = torch.full_like(t["logits"], float("nan"))
    gref = torch.zeros_like(t["ref"]) if want_ref else None
    cabi.fused_backward(t["grad_out"].data_ptr(), t["value"].data_ptr(), t["shapes"].data_ptr(), t["lsi"].data_ptr(),
                        t["ref"].data_ptr(), t["off"].data_ptr(), t["logits"].data_ptr(), dims, gv.data_ptr(),
                        goff.data_ptr(), glog.data_ptr(), gref.data_ptr() if want_ref else None, st)
    torch.cuda.synchronize()
    return dict(out=out.cpu().numpy(), grad_value=gv.cpu().numpy(), grad_off=goff.cpu().numpy(),
                grad_logits=glog.cpu().numpy(), grad_ref=gref.cpu().numpy() if want_ref else None,
                fwd_path=fwd_path, bwd_path=cabi.last_path(True)), t


FUSED_CASES = [
    # (N, M, L, P, Lq, levels, split)                                    what it exercises
    (2, 8, 4, 4, 611, ((20, 27), (10, 14), (5, 7), (3, 4)), 1),      # shipped geometry: shuffle softmax, pipelined phase A
    (2, 8, 4, 4, 77, ((20, 27), (10, 14), (5, 7), (3, 4)), 2),       # one-item-per-warp forward
    (1, 3, 3, 2, 50, ((9, 7), (5, 4), (2, 2)), 1),                   # L*P = 6, M % 4 != 0: looped softmax
    (1, 2, 8, 4, 21, tuple((4, 3) for _ in range(8)), 1),            # L*P = 32
    (1, 2, 8, 4, 21, tuple((4, 3) for _ in range(8)), 2),
    (2, 8, 2, 4, 33, ((6, 5), (3, 3)), 1),                           # L*P = 8 (power of two, runtime path)
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=[f"N{c[0]}M{c[1]}L{c[2]}P{c[3]}Lq{c[4]}s{c[6]}" for c in FUSED_CASES])
def test_fused_prologue_matches_unfused_restatement(case):
    N, M, L, P, Lq, levels, split = case
    x = _fused_inputs(N, M, L, P, Lq, levels, seed=17, ref_spread=1.3)     # some references outside [0, 1]
    loc, a, ref = _fused_reference(x)
    with Paths(0, 0, split):
        res, t = _run_fused(x)
        assert "fused" in res["fwd_path"] and ("split" in res["fwd_path"]) == (split == 2)
        assert res["bwd_path"] == "bwd_d32_fused_f32"
        # same locations bit for bit => the unfused kernels fed torch's own loc / softmax agree to rounding
        unf = run_cabi("f32", x["value"], x["shapes"], x["lsi"], loc, a, x["grad_out"])
    close(res["out"], ref["out"], FWD_RTOL)
    close(res["out"], unf["out"], 2e-6)
    close(res["grad_value"], ref["grad_value"], GRAD_RTOL)
    close(res["grad_off"], ref["grad_off"], GRAD_RTOL)
    close(res["grad_logits"], ref["grad_logits"], GRAD_RTOL)
    close(res["grad_ref"], ref["grad_ref"], GRAD_RTOL)
    # indices: which value rows receive gradient is integer work and must match the unfused path exactly
    assert np.array_equal(res["grad_value"] != 0, unf["grad_value"] != 0)
    res2, _ = _run_fused(x, want_ref=False)
    assert np.array_equal(res2["grad_off"], res["grad_off"]) and res2["grad_ref"] is None


@pytest.mark.parametrize("seed", range(int(os.environ.get("MSDA_FUZZ_SEEDS", "8"))))
def test_fused_prologue_fuzz(seed):
    """Random head / level / point counts and ragged level sizes through the fused kernels (both forward
    decompositions), against the unfused restatement."""
    rng = np.random.default_rng(7000 + seed)
    L = int(rng.integers(1, 9))
    P = int(rng.integers(1, 32 // L + 1))
    M = int(rng.integers(1, 10))
    N = int(rng.integers(1, 3))
    Lq = int(rng.integers(1, 90))
    levels = tuple((int(rng.integers(1, 14)), int(rng.integers(1, 14))) for _ in range(L))
    x = _fused_inputs(N, M, L, P, Lq, levels, seed=90 + seed, ref_spread=1.5)
    loc, a, ref = _fused_reference(x)
    for split in (1, 2):
        with Paths(0, 0, split):
            res, _ = _run_fused(x)
        close(res["out"], ref["out"], FWD_RTOL)
        close(res["grad_value"], ref["grad_value"], GRAD_RTOL)
        close(res["grad_off"], ref["grad_off"], GRAD_RTOL)
        close(res["grad_logits"], ref["grad_logits"], GRAD_RTOL)
        close(res["grad_ref"], ref["grad_ref"], GRAD_RTOL)


def test_fused_locations_are_torch_bits():
    """The fused kernels must land on exactly the pixels torch's `ref + off / wh` lands on, including
    at exact pixel centres where a 1-ulp difference flips the floor."""
    levels = ((100, 167), (50, 84), (25, 42), (13, 21))
    wl = Workload("f", 1, levels, sum(h * w for h, w in levels))
    shapes, lsi = (t.cuda() for t in __import__("uvhand_b200.workloads", fromlist=["level_tensors"]).level_tensors(levels))
    from uvhand_b200.workloads import encoder_reference_points
    ref = encoder_reference_points(levels, "cuda")[None, :, None, :].expand(1, wl.Lq, 4, 2).contiguous()
    # whole-pixel offsets (the module's initial ring, ms_deform_attn.py:62-70): samples sit ON pixel centres
    off = torch.randint(-3, 4, (1, wl.Lq, 8, 4, 4, 2), device="cuda").float()
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
    loc = (ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]).contiguous()
    value = torch.randn(1, wl.S, 8, 32, device="cuda")
    logits = torch.zeros(1, wl.Lq, 8, 16, device="cuda")
    aw = torch.full((1, wl.Lq, 8, 4, 4), 1.0 / 16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    out_f = torch.empty(1, wl.Lq, 256, device="cuda")
    out_u = torch.empty_like(out_f)
    cabi.fused_forward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), ref.data_ptr(), off.data_ptr(),
                       logits.data_ptr(), wl.dims(), out_f.data_ptr(), st)
    cabi.forward("f32", value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), aw.data_ptr(), wl.dims(),
                 out_u.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(out_f, out_u)          # uniform weights are exact (1/16): any difference is an index difference


@pytest.mark.parametrize("ref_dim", [2, 42])
def test_module_fused_prologue_equals_unfused(ref_dim):
    from uvhand_b200.modules import MSDeformAttn
    torch.manual_seed(11)
    levels = ((12, 17), (6, 9), (3, 5), (2, 3))
    shapes = torch.as_tensor(levels, dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    N, Lq = 2, 41
    plain = MSDeformAttn(256, 4, 8, 4, fuse_prologue=False).cuda()
    with torch.no_grad():
        plain.sampling_offsets.weight.normal_(0, 0.02)
        plain.attention_weights.weight.normal_(0, 0.1)
    assert MSDeformAttn(256, 4, 8, 4).fuse_prologue          # the default (round 2)
    fused = MSDeformAttn(256, 4, 8, 4, fuse_prologue=True).cuda()
    fused.load_state_dict(plain.state_dict())
    pad = torch.zeros(N, S, dtype=torch.bool, device="cuda")
    pad[:, -7:] = True
    base = dict(query=torch.randn(N, Lq, 256, device="cuda"), feat=torch.randn(N, S, 256, device="cuda"),
                refp=torch.rand(N, Lq, 4, ref_dim, device="cuda"))
    outs = []
    for mod in (plain, fused):
        leaf = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        n0 = cabi.launch_count()
        out = mod(leaf["query"], leaf["refp"], leaf["feat"], shapes, lsi, pad)
        out.square().sum().backward()
        torch.cuda.synchronize()
        assert cabi.launch_count() - n0 == 2
        assert ("fused" in cabi.last_path(False)) == mod.fuse_prologue
        outs.append((out.detach(), {k: v.grad for k, v in leaf.items()}, {k: p.grad for k, p in mod.named_parameters()}))
    (o0, g0, p0), (o1, g1, p1) = outs
    assert torch.allclose(o0, o1, rtol=1e-5, atol=1e-6)
    for k in g0:
        assert torch.allclose(g0[k], g1[k], rtol=1e-4, atol=1e-4 * float(g0[k].abs().max())), k
    for k in p0:
        assert torch.allclose(p0[k], p1[k], rtol=1e-4, atol=1e-4 * float(p0[k].abs().max())), k


def test_tensors_on_a_device_that_is_not_current():
    """The reference takes no device guard: with tensors on cuda:1 while cuda:0 is current it
    would launch on the wrong GPU.  The host layer here guards on value's device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from uvhand_b200.functions import MSDeformAttnFunction
    wl = Workload("dev1", 2, ((9, 13), (5, 7), (3, 4), (2, 2)), 40, loc_dist="twostage")
    cpu = make_inputs(wl, seed=13, device="cpu")
    with torch.cuda.device(0):
        x = {k: v.to("cuda:1") for k, v in cpu.items()}
        v = x["value"].requires_grad_(True)
        out = MSDeformAttnFunction.apply(v, x["shapes"], x["lsi"], x["loc"], x["aw"], 64)
        out.backward(x["grad_out"])
        torch.cuda.synchronize(1)
    assert out.device == torch.device("cuda:1") and v.grad.device == torch.device("cuda:1")
    n = {k: t.numpy() for k, t in cpu.items()}
    ref = oracle(n["value"], n["shapes"], n["lsi"], n["loc"], n["aw"], n["grad_out"])
    close(out.detach().cpu().numpy(), ref["out"], FWD_RTOL)
    close(v.grad.cpu().numpy(), ref["grad_value"], GRAD_RTOL)


# ------------------------------------------------------------------------------------------------
# launch hygiene: the entry points never synchronise, so a stack of calls can be
# captured in a CUDA graph and replayed
# ------------------------------------------------------------------------------------------------

def test_calls_are_capturable_in_a_cuda_graph():
    wl = WORKLOADS["c4_dino_decoder"]
    x = make_inputs(wl, seed=3, device="cuda")
    dims = wl.dims()
    out = torch.zeros(wl.N, wl.Lq, wl.M * wl.D, device="cuda")
    gv, gl, ga = torch.zeros_like(x["value"]), torch.zeros_like(x["loc"]), torch.zeros_like(x["aw"])

    def step(stream):
        gv.zero_()
        cabi.forward("f32", x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(), x["loc"].data_ptr(),
                     x["aw"].data_ptr(), dims, out.data_ptr(), stream)
        cabi.backward("f32", x["grad_out"].data_ptr(), x["value"].data_ptr(), x["shapes"].data_ptr(), x["lsi"].data_ptr(),
                      x["loc"].data_ptr(), x["aw"].data_ptr(), dims, gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), stream)

    step(torch.cuda.current_stream().cuda_stream)          # eager reference (also warms the library up)
    torch.cuda.synchronize()
    want = [t.clone() for t in (out, gl, ga)]
    want_gv = gv.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(6):                                 # e.g. the six decoder layers of one model step
            step(torch.cuda.current_stream().cuda_stream)
    for t in (out, gl, ga, gv):
        t.fill_(float("nan"))
    n0 = cabi.launch_count()
    graph.replay()
    torch.cuda.synchronize()
    assert cabi.launch_count() == n0                       # replay launches nothing through the host path
    for got, ref in zip((out, gl, ga), want):
        assert torch.equal(got, ref)
    close(gv.cpu().numpy(), want_gv.cpu().numpy(), GRAD_RTOL)   # atomics: order may differ between runs


# ------------------------------------------------------------------------------------------------
# host-buffer C ABI (what a non-torch caller binds)
# ------------------------------------------------------------------------------------------------

def test_host_plan_round_trip():
    wl = Workload("h", 2, ((20, 27), (10, 14), (5, 7), (3, 4)), 333, loc_dist="twostage")
    x = np_inputs(wl, 31)
    plan = cabi.HostPlan(*wl.dims())
    out = np.full((wl.N, wl.Lq, wl.M * wl.D), np.nan, np.float32)
    gv = np.full(x["value"].shape, np.nan, np.float32)
    gl = np.full(x["loc"].shape, np.nan, np.float32)
    ga = np.full(x["aw"].shape, np.nan, np.float32)
    plan.forward_backward(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], out, gv, gl, ga)
    ref = oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    check_all({"out": out, "grad_value": gv, "grad_loc": gl, "grad_aw": ga}, ref)
    out2 = np.empty_like(out)
    plan.forward(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], out2)
    assert np.array_equal(out, out2)
    # second call must not accumulate into stale grad_value
    plan.forward_backward(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], out, gv, gl, ga)
    close(gv, ref["grad_value"], GRAD_RTOL)
    h2d, d2h = plan.traffic()
    assert h2d >= x["value"].nbytes + x["loc"].nbytes + x["aw"].nbytes + x["grad_out"].nbytes
    assert d2h == out.nbytes + gv.nbytes + gl.nbytes + ga.nbytes
    plan.close()


# ------------------------------------------------------------------------------------------------
# the reference's own CUDA kernels, recompiled (oracle/_ref), as a second oracle
# ------------------------------------------------------------------------------------------------

def _ref_cuda():
    import ctypes
    p = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "libmsda_refcuda.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/libmsda_refcuda.so not built (needs /root/reference at build time)")
    return ctypes.CDLL(p)


@pytest.mark.parametrize("dist,Lq", [("encoder", MID_S), ("twostage", 300)])
def test_against_reference_cuda_kernels(dist, Lq):
    import ctypes
    lib = _ref_cuda()
    wl = Workload("mid", 2, MID_LEVELS, Lq, loc_dist=dist)
    t = make_inputs(wl, seed=17, device="cuda")
    N, S, M, D, L, Lq, P = wl.dims()
    vp = ctypes.c_void_p
    ints = [ctypes.c_int] * 8
    lib.msda_ref_fwd_f32.argtypes = [vp] * 5 + ints + [vp, vp]
    lib.msda_ref_bwd_f32.argtypes = [vp] * 6 + ints + [vp] * 4
    out_r = torch.empty(N, Lq, M * D, device="cuda")
    gv_r, gl_r, ga_r = torch.empty_like(t["value"]), torch.empty_like(t["loc"]), torch.empty_like(t["aw"])
    st = torch.cuda.current_stream().cuda_stream
    assert lib.msda_ref_fwd_f32(t["value"].data_ptr(), t["shapes"].data_ptr(), t["lsi"].data_ptr(), t["loc"].data_ptr(),
                                t["aw"].data_ptr(), N, S, M, D, L, Lq, P, 64, out_r.data_ptr(), st) == 0
    assert lib.msda_ref_bwd_f32(t["grad_out"].data_ptr(), t["value"].data_ptr(), t["shapes"].data_ptr(),
                                t["lsi"].data_ptr(), t["loc"].data_ptr(), t["aw"].data_ptr(), N, S, M, D, L, Lq, P, 64,
                                gv_r.data_ptr(), gl_r.data_ptr(), ga_r.data_ptr(), st) == 0
    torch.cuda.synchronize()
    res = run_cabi("f32", *[t[k].cpu().numpy() for k in ("value", "shapes", "lsi", "loc", "aw", "grad_out")])
    close(res["out"], out_r.cpu().numpy(), FWD_RTOL)
    close(res["grad_value"], gv_r.cpu().numpy(), GRAD_RTOL)
    close(res["grad_loc"], gl_r.cpu().numpy(), GRAD_RTOL)
    close(res["grad_aw"], ga_r.cpu().numpy(), GRAD_RTOL)
    # identical index decisions => identical touch pattern of grad_value
    assert np.array_equal(res["grad_value"] != 0, gv_r.cpu().numpy() != 0)


# ------------------------------------------------------------------------------------------------
# round 2: full-size gradients against the oracle for every BASELINE configuration
# ------------------------------------------------------------------------------------------------

def _oracle_one_sample(t, b, value_from_bf16=False):
    """Oracle forward + backward for sample b of a full-size input set (the operator has no cross-sample term)."""
    cpu = {k: (v[b:b + 1] if k in ("value", "loc", "aw", "grad_out") else v).cpu() for k, v in t.items()}
    value = cpu["value"].float().numpy() if value_from_bf16 else cpu["value"].numpy()
    return oracle(value, cpu["shapes"].numpy(), cpu["lsi"].numpy(), cpu["loc"].numpy(), cpu["aw"].numpy(),
                  cpu["grad_out"].numpy())


@pytest.mark.parametrize("name", ["c5_hires", "c5_hires_contended"])
def test_full_size_c5_bf16_gradients_vs_oracle(name):
    """BASELINE configs[4] (bf16 value, grad_value contention) at full size: all three gradients and the output of
    one of the 8 images against the OpenMP oracle run on the values the kernel actually read."""
    wl = WORKLOADS[name]
    t, out, gv, gl, ga = _full_size(wl, kind="bf16", seed=3)
    assert cabi.last_path(True) == "bwd_d32_tiles_bf16_f32acc"
    b = 5
    ref = _oracle_one_sample(t, b, value_from_bf16=True)
    close(out[b:b + 1].cpu().numpy(), ref["out"], FWD_RTOL)
    close(gl[b:b + 1].cpu().numpy(), ref["grad_loc"], GRAD_RTOL)
    close(ga[b:b + 1].cpu().numpy(), ref["grad_aw"], GRAD_RTOL)
    close(gv[b:b + 1].float().cpu().numpy(), ref["grad_value"], BF16_RTOL)     # rounded once to bf16
    # which rows receive gradient is index work: exact
    assert np.array_equal(gv[b:b + 1].float().cpu().numpy() != 0, ref["grad_value"] != 0)


@pytest.mark.parametrize("tiles", [1, 0], ids=["direct", "auto"])
def test_full_size_c5_f32_contended_grad_value(tiles):
    """The contended distribution in fp32 through both backward families (tolerance of the fp32 path)."""
    wl = WORKLOADS["c5_hires_contended"]
    with Paths(0, 0, 1, tiles, 1):
        t, out, gv, gl, ga = _full_size(wl, kind="f32", seed=4)
        assert cabi.last_path(True) == ("bwd_d32_f32" if tiles == 1 else "bwd_d32_tiles_f32")
    ref = _oracle_one_sample(t, 2)
    close(gv[2:3].cpu().numpy(), ref["grad_value"], GRAD_RTOL)
    close(gl[2:3].cpu().numpy(), ref["grad_loc"], GRAD_RTOL)


@pytest.mark.parametrize("name,b", [("c3_arctic_encoder", 17), ("c3_arctic_decoder", 30), ("c4_dino_decoder", 1),
                                    ("c4_dino_decoder_train", 0)])
def test_full_size_c3_c4_grad_value_vs_oracle(name, b):
    wl = WORKLOADS[name]
    t, out, gv, gl, ga = _full_size(wl, seed=5)
    ref = _oracle_one_sample(t, b)
    close(out[b:b + 1].cpu().numpy(), ref["out"], FWD_RTOL)
    close(gv[b:b + 1].cpu().numpy(), ref["grad_value"], GRAD_RTOL)
    close(gl[b:b + 1].cpu().numpy(), ref["grad_loc"], GRAD_RTOL)
    close(ga[b:b + 1].cpu().numpy(), ref["grad_aw"], GRAD_RTOL)


def test_tile_backward_when_its_windows_miss():
    """Encoder-shaped call (Lq == S picks the tile backward) whose locations have nothing to do with the query
    grid: almost every row pair misses the unit's windows and is served without combining (the kernel's leftover
    path).  Slower than the direct kernel by construction, but it must still be right."""
    wl = Workload("miss", 1, ((60, 70), (30, 35), (24, 25), (22, 23)), 60 * 70 + 30 * 35 + 24 * 25 + 22 * 23,
                  loc_dist="uniform")
    x = np_inputs(wl, 31)
    with Paths(0, 0, 1, 0, 1):
        res = run_cabi("f32", x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    assert res["bwd_path"] == "bwd_d32_tiles_f32"
    check_all(res, oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"]))


@pytest.mark.parametrize("levels,lsi,S", [
    # tokens nobody's level owns at both ends and between the levels; levels stored out of order
    (((20, 27), (10, 14), (5, 7), (3, 4)), (40, 620, 800, 580), 877),
    # the same with a pyramid too large to be windowed whole (8x8 pixel tiles with moving windows)
    (((40, 57), (20, 29), (10, 15), (5, 8)), (2300, 1500, 100, 0), 4700),
    # two levels that overlap in `value` (legal for the operator: the levels are just views)
    (((20, 27), (10, 14), (5, 7), (3, 4)), (0, 300, 300, 0), 540),
])
def test_tile_backward_with_irregular_level_layout(levels, lsi, S):
    """Lq == S selects the pixel-block tile backward, which ASSUMES query i sits at token i's pixel only to place its
    windows: whatever level_start_index says (gaps, other orders, overlapping levels, tokens beyond the pyramid),
    every query in [0, Lq) must be processed exactly once (the operator itself never checks the layout: cuh:258-297)."""
    N, M, D, L, P = 2, 8, 32, 4, 4
    rng = np.random.default_rng(S)
    x = {
        "value": rng.standard_normal((N, S, M, D), dtype=np.float32),
        "shapes": np.asarray(levels, np.int64), "lsi": np.asarray(lsi, np.int64),
        "loc": rng.random((N, S, M, L, P, 2), dtype=np.float32) * 1.2 - 0.1,
        "aw": rng.random((N, S, M, L, P), dtype=np.float32),
        "grad_out": rng.standard_normal((N, S, M * D), dtype=np.float32),
    }
    with Paths(0, 0, 1, 2, 1, 2):       # tile backward and, where the pyramid fits, the slab forward (it stages ALL S tokens)
        res = run_cabi("f32", x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    assert res["bwd_path"] == "bwd_d32_tiles_f32"
    assert res["fwd_path"] == ("fwd_d32_slab_f32" if S <= 1600 else "fwd_d32_f32")
    check_all(res, oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"]))


def test_slab_forward_edges():
    """The slab forward's share arithmetic: fewer passes than SMs, a last pass with a few queries, one query, and a
    pyramid at the row limit of the slab."""
    for N, Lq, levels, seed in ((1, 1, ((5, 6), (3, 3), (2, 2), (1, 1)), 1), (3, 97, ((9, 11), (5, 6), (3, 3), (2, 2)), 2),
                                (2, 700, ((36, 36), (15, 16), (8, 8), (0, 0)), 3), (1, 130, ((30, 40), (15, 20), (8, 10), (4, 5)), 4)):
        wl = Workload("slab_edge", N, levels, Lq, loc_dist="uniform")
        x = np_inputs(wl, seed)
        with Paths(0, 0, 1, 1, 1, 2):
            res = run_cabi("f32", x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
        assert res["fwd_path"] == "fwd_d32_slab_f32", (levels, res["fwd_path"])
        check_all(res, oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"]))


def test_backward_with_no_value_tokens_writes_zero_gradients():
    """S == 0: msda_bwd_* promise grad_sampling_loc / grad_attn_weight 'fully written' (include/msda_b200.h)."""
    N, M, D, L, Lq, P = 2, 8, 32, 4, 5, 4
    sh = torch.zeros(L, 2, dtype=torch.long, device="cuda")
    ls = torch.zeros(L, dtype=torch.long, device="cuda")
    lo = torch.rand(N, Lq, M, L, P, 2, device="cuda")
    a = torch.rand(N, Lq, M, L, P, device="cuda")
    g = torch.randn(N, Lq, M * D, device="cuda")
    v = torch.zeros(1, device="cuda")
    gl = torch.full_like(lo, float("nan"))
    ga = torch.full_like(a, float("nan"))
    cabi.backward("f32", g.data_ptr(), v.data_ptr(), sh.data_ptr(), ls.data_ptr(), lo.data_ptr(), a.data_ptr(),
                  (N, 0, M, D, L, Lq, P), v.data_ptr(), gl.data_ptr(), ga.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert float(gl.abs().max()) == 0 and float(ga.abs().max()) == 0


def test_stale_cuda_error_is_not_blamed_on_the_operator():
    """A non-sticky error left by an earlier caller (here: an impossible allocation) must not fail the next
    operator call -- and must be gone afterwards."""
    import ctypes
    g = np.load(GOLDEN[IDS.index("d32_m8_l4_p4")])
    v, lo, a = dev(g["value"]), dev(g["loc"]), dev(g["aw"])
    sh, ls = dev(g["shapes"].astype(np.int64)), dev(g["lsi"].astype(np.int64))
    N, S, M, D = v.shape
    _, Lq, _, L, P, _ = lo.shape
    out = torch.empty(N, Lq, M * D, device="cuda")
    torch.cuda.synchronize()
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    ptr = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(ptr), ctypes.c_size_t(1 << 62)) != 0      # leaves cudaErrorMemoryAllocation behind
    try:
        cabi.forward("f32", v.data_ptr(), sh.data_ptr(), ls.data_ptr(), lo.data_ptr(), a.data_ptr(),
                     (N, S, M, D, L, Lq, P), out.data_ptr(), torch.cuda.current_stream().cuda_stream)   # raises on rc != 0
    finally:
        # the error was planted in the process's shared CUDA runtime (the one torch uses); the library links its own
        # runtime instance and neither sees nor may clear it -- the test has to, or the next torch call trips over it
        rt.cudaGetLastError()
    torch.cuda.synchronize()
    close(out.cpu().numpy(), g["out"], FWD_RTOL)


def test_module_half_precision_branch():
    """models/dino/ops/modules/ms_deform_attn.py:124-131: a half-precision module upcasts value / locations / weights
    to fp32 around the operator and casts the result back."""
    from uvhand_b200.modules import MSDeformAttn
    torch.manual_seed(9)
    levels = ((12, 17), (6, 9), (3, 5), (2, 3))
    shapes = torch.as_tensor(levels, dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    N, Lq = 2, 23
    mod32 = MSDeformAttn(256, 4, 8, 4).cuda()
    with torch.no_grad():
        mod32.sampling_offsets.weight.normal_(0, 0.02)
        mod32.attention_weights.weight.normal_(0, 0.1)
    mod16 = MSDeformAttn(256, 4, 8, 4).cuda()
    mod16.load_state_dict(mod32.state_dict())
    mod16 = mod16.half()
    query = torch.randn(N, Lq, 256, device="cuda")
    feat = torch.randn(N, S, 256, device="cuda")
    refp = torch.rand(N, Lq, 4, 2, device="cuda")
    q16 = query.half().requires_grad_(True)
    f16 = feat.half().requires_grad_(True)
    out16 = mod16(q16, refp.half(), f16, shapes, lsi, None)
    assert out16.dtype == torch.float16
    assert cabi.last_path(False) in ("fwd_d32_f32", "fwd_d32_split_f32")          # the operator itself ran in fp32
    out16.float().sum().backward()
    assert q16.grad.dtype == torch.float16 and f16.grad.dtype == torch.float16
    out32 = mod32(query, refp, feat, shapes, lsi, None)
    assert torch.allclose(out16.float(), out32, rtol=2e-2, atol=2e-2)
    assert torch.isfinite(q16.grad.float()).all() and torch.isfinite(f16.grad.float()).all()


def test_value_pack_bf16_equals_torch_mask_then_cast():
    """one pass for padding mask + bf16 cast, and its backward, bit for bit what the torch
    sequence masked_fill -> to(bfloat16) (and its autograd) gives."""
    from uvhand_b200.functions import ValueToBF16Function
    torch.manual_seed(3)
    N, S, C = 3, 517, 256
    v = torch.randn(N, S, C, device="cuda")
    mask = torch.rand(N, S, device="cuda") < 0.2
    for m in (mask, None):
        a = v.clone().requires_grad_(True)
        b = v.clone().requires_grad_(True)
        out = ValueToBF16Function.apply(a, m)
        want = (b.masked_fill(m[..., None], 0.0) if m is not None else b).to(torch.bfloat16)
        assert out.dtype == torch.bfloat16 and torch.equal(out, want)
        g = torch.randn(N, S, C, device="cuda").to(torch.bfloat16)
        out.backward(g)
        want.backward(g)
        assert torch.equal(a.grad, b.grad)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_module_bf16_value_path(ref_dim):
    """MSDeformAttn(value_dtype=bfloat16): value written once as bf16 (mask folded in), bf16-value kernels, fp32
    everything else -- against the fp32 module with the operator replaced by the port of the reference CPU path,
    at the bf16 tolerance of BASELINE.json (1e-2)."""
    from oracle.grid_sample_port import msda_grid_sample
    from uvhand_b200.modules import MSDeformAttn
    torch.manual_seed(6)
    levels = ((12, 17), (6, 9), (3, 5), (2, 3))
    shapes = torch.as_tensor(levels, dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    N, Lq = 2, 37
    mod = MSDeformAttn(256, 4, 8, 4, value_dtype=torch.bfloat16).cuda()
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    query = torch.randn(N, Lq, 256, device="cuda", requires_grad=True)
    feat = torch.randn(N, S, 256, device="cuda", requires_grad=True)
    refp = torch.rand(N, Lq, 4, ref_dim, device="cuda")
    if ref_dim == 4:
        refp[..., 2:] *= 0.3
    pad = torch.zeros(N, S, dtype=torch.bool, device="cuda")
    pad[:, -5:] = True
    n0 = cabi.launch_count()
    out = mod(query, refp, feat, shapes, lsi, pad)
    assert out.dtype == torch.float32 and "bf16" in cabi.last_path(False)
    out.sum().backward()
    torch.cuda.synchronize()
    assert "bf16" in cabi.last_path(True)
    assert cabi.launch_count() - n0 == 5       # pack, forward, backward + final rounding, unpack
    gq, gf = query.grad.clone(), feat.grad.clone()
    query.grad = feat.grad = None
    value = mod.value_proj(feat).masked_fill(pad[..., None], 0.0).view(N, S, 8, 32)
    off = mod.sampling_offsets(query).view(N, Lq, 8, 4, 4, 2)
    w = torch.softmax(mod.attention_weights(query).view(N, Lq, 8, 16), -1).view(N, Lq, 8, 4, 4)
    loc = MSDeformAttn._sampling_locations(refp, off, shapes)
    out2 = mod.output_proj(msda_grid_sample(value, levels, loc, w))
    out2.sum().backward()
    tol = BF16_RTOL
    assert torch.allclose(out, out2, rtol=tol, atol=tol * float(out2.abs().max()))
    assert torch.allclose(gq, query.grad, rtol=tol, atol=tol * float(query.grad.abs().max()))
    assert torch.allclose(gf, feat.grad, rtol=tol, atol=tol * float(feat.grad.abs().max()))
    assert float(gf[:, -5:].abs().max()) == 0          # padded tokens receive no gradient


@pytest.mark.parametrize("case", [
    (2, 8, 735, ((20, 27), (10, 14), (5, 7), (3, 4)), 0.4),        # small pyramid windowed whole, linear tiles, refs mostly inside
    (1, 8, 40 * 57 + 20 * 29 + 10 * 15 + 5 * 8, MID_LEVELS, 1.3),  # Lq == S: 8x8 pixel tiles, windows, leftovers, refs outside
    (2, 3, 200, ((9, 13), (5, 7), (3, 4), (2, 2)), 1.0),           # 3 heads
], ids=["whole", "grid", "m3"])
def test_fused_prologue_through_the_tile_backward(case):
    """msda_fused_bwd_f32 dispatched to the on-chip-combining backward (softmax / location prologue in its geometry
    phase, their backward in its epilogue), against the unfused restatement."""
    N, M, Lq, levels, spread = case
    x = _fused_inputs(N, M, 4, 4, Lq, levels, seed=23, ref_spread=spread)
    loc, a, ref = _fused_reference(x)
    with Paths(0, 0, 1, 2, 1):
        res, _ = _run_fused(x)
        assert res["bwd_path"] == "bwd_d32_tiles_fused_f32"
        res2, _ = _run_fused(x, want_ref=False)
    close(res["out"], ref["out"], FWD_RTOL)
    close(res["grad_value"], ref["grad_value"], GRAD_RTOL)
    close(res["grad_off"], ref["grad_off"], GRAD_RTOL)
    close(res["grad_logits"], ref["grad_logits"], GRAD_RTOL)
    close(res["grad_ref"], ref["grad_ref"], GRAD_RTOL)
    assert np.array_equal(res2["grad_off"], res["grad_off"]) and res2["grad_ref"] is None


@pytest.mark.parametrize("model", ["c3", "c4"])
def test_layer_stack_graph_replay_against_the_oracle(model):
    """the 6 encoder + 6 decoder operator calls of one model step (forward in order, backward in
    reverse) captured as ONE CUDA graph; the replayed step's outputs and gradients of every layer are scored against
    the ORACLE (tools/layer_stack.py), and a replay goes through no host launch path."""
    sys_path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools")
    import importlib.util
    spec = importlib.util.spec_from_file_location("layer_stack", os.path.join(sys_path, "layer_stack.py"))
    ls = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ls)
    layers = ls.build_stack(model)
    graph = ls.capture(layers)
    ls.poison(layers)
    n0 = cabi.launch_count()
    graph.replay()
    torch.cuda.synchronize()
    assert cabi.launch_count() == n0
    worst = ls.check_against_oracle(layers, frames=1, fwd_rtol=FWD_RTOL, grad_rtol=GRAD_RTOL)
    assert worst["out"] <= FWD_RTOL and worst["grad_value"] <= GRAD_RTOL


@pytest.mark.parametrize("chunks", [None, "2"])
def test_host_pipe_steps_in_flight(chunks, monkeypatch):
    """msda_host_pipe_*: five steps with different inputs through a pipe of depth 2 (a slot is reused while the other
    is in flight; pageable AND page-locked buffers), every step against the oracle; the level tables change on the way
    (the plan caches them)."""
    if chunks:
        monkeypatch.setenv("MSDA_B200_PIPE_CHUNKS", chunks)
    levels_a, levels_b = ((20, 27), (10, 14), (5, 7), (3, 4)), ((27, 20), (14, 10), (7, 5), (4, 3))    # same S
    S = sum(h * w for h, w in levels_a)
    pipe = cabi.HostPipe(2, S, 8, 32, 4, 77, 4, depth=2)
    with pytest.raises(ValueError):
        pipe.submit(2, *([None] * 10))
    steps = []
    for i in range(5):
        wl = Workload("pipe", 2, levels_b if i == 3 else levels_a, 77, loc_dist="twostage")
        x = np_inputs(wl, 40 + i)
        o = {"out": np.full((2, 77, 256), np.nan, np.float32), "gv": np.full(x["value"].shape, np.nan, np.float32),
             "gl": np.full(x["loc"].shape, np.nan, np.float32), "ga": np.full(x["aw"].shape, np.nan, np.float32)}
        if i % 2:
            for a in list(o.values()) + [x[k] for k in ("value", "loc", "aw", "grad_out")]:
                cabi.host_register(a)
        steps.append((x, o))
    for i, (x, o) in enumerate(steps):
        pipe.submit(i % 2, x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], o["out"], o["gv"], o["gl"], o["ga"])
    pipe.wait()
    h2d, d2h = pipe.traffic()
    assert h2d > 0 and d2h > 0
    pipe.close()
    for i, (x, o) in enumerate(steps):
        if i % 2:
            for a in list(o.values()) + [x[k] for k in ("value", "loc", "aw", "grad_out")]:
                cabi.host_unregister(a)
        ref = oracle(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
        check_all({"out": o["out"], "grad_value": o["gv"], "grad_loc": o["gl"], "grad_aw": o["ga"]}, ref)
    assert cabi.lib().msda_host_pipe_create(ctypes.byref(ctypes.c_void_p()), 0, 0, 1, 1, 1, 1, 1, 1, 1) == -2   # depth 0
    assert cabi.lib().msda_host_pipe_create(ctypes.byref(ctypes.c_void_p()), 0, 9, 1, 1, 1, 1, 1, 1, 1) == -2


def test_host_pipe_bf16_storage():
    """msda_host_pipe_*_bf16: three steps in flight with value / grad_value as bf16 bit patterns, against the oracle on the
    rounded value; an fp32 submit on a bf16 pipe is rejected."""
    wl = Workload("p16", 2, ((20, 27), (10, 14), (5, 7), (3, 4)), 150, loc_dist="twostage")
    pipe = cabi.HostPipe(*wl.dims(), depth=3, value_bf16=True)
    steps = []
    for i in range(3):
        x = np_inputs(wl, 50 + i)
        v16 = torch.from_numpy(x["value"]).to(torch.bfloat16)
        o = {"out": np.full((wl.N, wl.Lq, wl.M * wl.D), np.nan, np.float32), "gv": np.zeros(x["value"].shape, np.uint16),
             "gl": np.full(x["loc"].shape, np.nan, np.float32), "ga": np.full(x["aw"].shape, np.nan, np.float32)}
        steps.append((x, v16.view(torch.int16).numpy().view(np.uint16).copy(), v16.float().numpy(), o))
    for i, (x, bits, _, o) in enumerate(steps):
        pipe.submit(i, bits, x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], o["out"], o["gv"], o["gl"], o["ga"])
    pipe.wait()
    x, bits, _, o = steps[0]
    rc = cabi.lib().msda_host_pipe_submit_f32(pipe._h, 0, *[a.ctypes.data_as(ctypes.c_void_p) for a in (
        x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], o["out"], x["value"].copy(), o["gl"], o["ga"])])
    assert rc == -4                                        # MSDA_ERR_UNSUPPORTED
    pipe.close()
    for x, _, v_round, o in steps:
        ref = oracle(v_round, x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
        close(o["out"], ref["out"], FWD_RTOL)
        close(o["gl"], ref["grad_loc"], GRAD_RTOL)
        close(o["ga"], ref["grad_aw"], GRAD_RTOL)
        close(torch.from_numpy(o["gv"].view(np.int16)).view(torch.bfloat16).float().numpy(), ref["grad_value"], BF16_RTOL)


def test_host_plan_bf16_storage_round_trip():
    """msda_host_fwd_bwd_bf16: value up / grad_value down as bf16 (fp32 accumulation on the device, one rounding),
    everything else fp32, against the oracle run on the bf16-rounded value."""
    wl = Workload("h16", 2, ((20, 27), (10, 14), (5, 7), (3, 4)), 333, loc_dist="twostage")
    x = np_inputs(wl, 33)
    v16 = torch.from_numpy(x["value"]).to(torch.bfloat16)
    v_bits = v16.view(torch.int16).numpy().view(np.uint16).copy()
    v_round = v16.float().numpy()
    plan = cabi.HostPlan(*wl.dims(), value_bf16=True)
    out = np.full((wl.N, wl.Lq, wl.M * wl.D), np.nan, np.float32)
    gv = np.zeros(x["value"].shape, np.uint16)
    gl = np.full(x["loc"].shape, np.nan, np.float32)
    ga = np.full(x["aw"].shape, np.nan, np.float32)
    for _ in range(2):                                   # a plan is reusable
        plan.forward_backward(v_bits, x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], out, gv, gl, ga)
    h2d, d2h = plan.traffic()
    assert h2d < 4 * (x["value"].size + x["loc"].size + x["aw"].size + x["grad_out"].size)
    with pytest.raises(ValueError):
        plan.forward_backward(x["value"], x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"], out, gv, gl, ga)
    plan.close()
    ref = oracle(v_round, x["shapes"], x["lsi"], x["loc"], x["aw"], x["grad_out"])
    close(out, ref["out"], FWD_RTOL)
    close(gl, ref["grad_loc"], GRAD_RTOL)
    close(ga, ref["grad_aw"], GRAD_RTOL)
    gv_f = torch.from_numpy(gv.view(np.int16)).view(torch.bfloat16).float().numpy()
    close(gv_f, ref["grad_value"], BF16_RTOL)
