"""CPU tests: pin the oracle (C restatement, numpy restatement, grid_sample port) to the golden
vectors written by the unmodified reference, and to each other."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, msda_numpy
from oracle.grid_sample_port import msda_grid_sample

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def _rel(a, b):
    scale = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() / scale


@pytest.fixture(scope="module", autouse=True)
def _built():
    c_oracle.build()


def test_golden_present():
    assert len(GOLDEN) >= 7


def test_reference_anchor_value():
    # first forward of models/ops/test.py with seed 3, N=1.
    g = np.load(GOLDEN[IDS.index("testpy_n1")])
    want = np.array([[0.0019, 0.0046, 0.0047, 0.0044], [0.0038, 0.0025, 0.0018, 0.0036]])
    assert np.allclose(np.round(g["out"][0], 4), want, atol=1e-4)


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_c_oracle_f64_matches_reference(path):
    g = np.load(path)
    out = c_oracle.forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], "f64")
    gv, gl, ga = c_oracle.backward(g["grad_out"], g["value"], g["shapes"], g["lsi"], g["loc"],
                                   g["aw"], "f64")
    # Both sides are float64; the only difference is grid_sample's un-normalisation
    # ((g+1)*W-1)/2 versus loc*W-0.5, a few ulps of the pixel coordinate.
    assert _rel(out, g["out"]) < 1e-12
    assert _rel(gv, g["grad_value"]) < 1e-12
    assert _rel(gl, g["grad_loc"]) < 1e-11
    assert _rel(ga, g["grad_aw"]) < 1e-12


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_c_oracle_f32x_close_to_reference(path):
    """fp32 pixel coordinates, unrounded sums: differs from the fp64 reference only by the
    rounding of loc*size-0.5 to float32 (relative 6e-8 of a coordinate of size <= W)."""
    g = np.load(path)
    out = c_oracle.forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], "f32x")
    gv, gl, ga = c_oracle.backward(g["grad_out"], g["value"], g["shapes"], g["lsi"], g["loc"],
                                   g["aw"], "f32x")
    assert _rel(out, g["out"]) < 2e-6
    assert _rel(gv, g["grad_value"]) < 2e-6
    assert _rel(ga, g["grad_aw"]) < 2e-6
    assert _rel(gl, g["grad_loc"]) < 2e-5


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_c_oracle_f32_strict(path):
    g = np.load(path)
    out = c_oracle.forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], "f32")
    assert out.dtype == np.float32
    assert _rel(out.astype(np.float64), g["out"]) < 1e-5


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_numpy_restatement_matches_c(path):
    g = np.load(path)
    h, w, k = c_oracle.sample_index(g["shapes"], g["loc"], "f32")
    h2, w2, k2, _, _ = msda_numpy.sample_index(g["shapes"], g["loc"])
    assert np.array_equal(h, h2) and np.array_equal(w, w2) and np.array_equal(k, k2)
    out_c = c_oracle.forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], "f32x")
    out_n = msda_numpy.forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"])
    assert _rel(out_n, out_c) < 1e-13
    gc = c_oracle.backward(g["grad_out"], g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], "f32x")
    gn = msda_numpy.backward(g["grad_out"], g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"])
    for a, b in zip(gn, gc):
        assert _rel(a, b) < 1e-12


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_grid_sample_port_matches_reference(path):
    g = np.load(path)
    v = torch.from_numpy(g["value"]).double().requires_grad_(True)
    lo = torch.from_numpy(g["loc"]).double().requires_grad_(True)
    a = torch.from_numpy(g["aw"]).double().requires_grad_(True)
    out = msda_grid_sample(v, g["shapes"].tolist(), lo, a)
    out.backward(torch.from_numpy(g["grad_out"]).double())
    assert _rel(out.detach().numpy(), g["out"]) < 1e-14
    assert _rel(v.grad.numpy(), g["grad_value"]) < 1e-13
    assert _rel(lo.grad.numpy(), g["grad_loc"]) < 1e-13
    assert _rel(a.grad.numpy(), g["grad_aw"]) < 1e-13
    out32 = msda_grid_sample(torch.from_numpy(g["value"]), g["shapes"].tolist(),
                             torch.from_numpy(g["loc"]), torch.from_numpy(g["aw"]))
    assert _rel(out32.numpy(), g["out_ref_fp32"]) < 1e-6


def test_special_locations_rejected():
    """NaN / +-inf / far-outside locations contribute exactly zero (cuh:288: every comparison
    is false) -- the grid_sample path returns NaN there, so only the restatements can say so."""
    shapes = np.array([[4, 5], [2, 3]], dtype=np.int64)
    lsi = np.array([0, 20], dtype=np.int64)
    rng = np.random.default_rng(0)
    value = rng.standard_normal((1, 26, 2, 4)).astype(np.float32)
    loc = rng.random((1, 3, 2, 2, 2, 2)).astype(np.float32)
    loc[0, 0, 0, 0, 0] = [np.nan, 0.5]
    loc[0, 0, 0, 0, 1] = [0.5, np.inf]
    loc[0, 0, 1, 1, 0] = [-np.inf, 0.5]
    loc[0, 1, 0, 0, 0] = [7.0, 0.5]
    loc[0, 1, 0, 1, 1] = [0.5, -3.0]
    aw = np.full((1, 3, 2, 2, 2), 0.25, dtype=np.float32)
    h, w, k = c_oracle.sample_index(shapes, loc, "f32")
    for idx in [(0, 0, 0, 0, 0), (0, 0, 0, 0, 1), (0, 0, 1, 1, 0), (0, 1, 0, 0, 0), (0, 1, 0, 1, 1)]:
        assert k[idx] == 0 and h[idx] == 0 and w[idx] == 0
    out = c_oracle.forward(value, shapes, lsi, loc, aw, "f32x")
    assert np.isfinite(out).all()
    g = rng.standard_normal((1, 3, 8)).astype(np.float32)
    gv, gl, ga = c_oracle.backward(g, value, shapes, lsi, loc, aw, "f32x")
    assert np.isfinite(gv).all() and np.isfinite(gl).all() and np.isfinite(ga).all()
    assert gl[0, 0, 0, 0, 0].tolist() == [0.0, 0.0] and ga[0, 0, 0, 0, 0] == 0.0
    out_n = msda_numpy.forward(value, shapes, lsi, loc, aw)
    assert _rel(out_n, out) < 1e-13


def test_border_masks_exact():
    """Hand-checked corner masks at the borders of a 3x4 level (cuh:55-78)."""
    shapes = np.array([[3, 4]], dtype=np.int64)
    H, W = 3, 4

    def one(x_pix, y_pix):
        loc = np.array([(x_pix + 0.5) / W, (y_pix + 0.5) / H], dtype=np.float32).reshape(1, 1, 1, 1, 1, 2)
        h, w, k = c_oracle.sample_index(shapes, loc, "f32")
        return int(h.ravel()[0]), int(w.ravel()[0]), int(k.ravel()[0])

    assert one(1.25, 1.25) == (1, 1, 16 | 15)
    assert one(-0.5, 1.25) == (1, -1, 16 | 2 | 8)           # left of column 0: only the x+1 corners
    assert one(3.25, 0.25) == (0, 3, 16 | 1 | 4)            # right edge: only the x corners
    assert one(1.25, -0.75) == (-1, 1, 16 | 4 | 8)          # above row 0
    assert one(1.25, 2.5) == (2, 1, 16 | 1 | 2)             # below last row
    assert one(-0.25, -0.25) == (-1, -1, 16 | 8)
    assert one(-1.0, 1.0)[2] == 0                           # w_im == -1 is rejected (strict >)
    assert one(4.0, 1.0)[2] == 0                            # w_im == W is rejected (strict <)


def test_fused_vs_unfused_floor_is_the_fused_one():
    """at pixel centres the two schedules disagree on floor for ~1% of
    points; the oracle must follow the compiled reference (one rounding)."""
    W = 167
    k = np.arange(W, dtype=np.float64)
    locx = ((k + 0.5) / W).astype(np.float32)
    fused = np.floor((locx.astype(np.float64) * W - 0.5).astype(np.float32)).astype(np.int32)
    unfused = np.floor((locx * np.float32(W)).astype(np.float32) - np.float32(0.5)).astype(np.int32)
    assert (fused != unfused).any()
    loc = np.stack([locx, np.full_like(locx, 0.5)], -1).reshape(1, W, 1, 1, 1, 2)
    shapes = np.array([[1, W]], dtype=np.int64)
    _, w_low, mask = c_oracle.sample_index(shapes, loc, "f32")
    assert np.array_equal(w_low.ravel(), fused)
    assert (mask.ravel() & 16).all()
