"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/_build/libmsda_oracle.so.

The library is the plain-C restatement of the reference CUDA arithmetic
(oracle/msda_oracle.c).  Inputs and outputs are numpy arrays; nothing here touches a GPU.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmsda_oracle.so")
_lib = None

_VARIANTS = {
    # name -> (storage dtype, carried dtype)
    "f32": (np.float32, np.float32),
    "f32x": (np.float32, np.float64),
    "f64": (np.float64, np.float64),
}


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc + OpenMP).  Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("msda_oracle.c", "msda_oracle_body.inc")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, aw, variant):
    st, acc = _VARIANTS[variant]
    value = np.ascontiguousarray(value, dtype=st)
    loc = np.ascontiguousarray(loc, dtype=st)
    aw = np.ascontiguousarray(aw, dtype=st)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, M2, L, P, two = loc.shape
    assert M2 == M and two == 2 and shapes.shape == (L, 2) and lsi.shape == (L,)
    assert aw.shape == (N, Lq, M, L, P)
    return value, shapes, lsi, loc, aw, (N, S, M, D, L, Lq, P), st, acc


def forward(value, shapes, lsi, loc, aw, variant: str = "f32x") -> np.ndarray:
    """out[N, Lq, M*D] in the variant's carried dtype."""
    value, shapes, lsi, loc, aw, dims, st, acc = _prep(value, shapes, lsi, loc, aw, variant)
    N, S, M, D, L, Lq, P = dims
    out = np.empty((N, Lq, M * D), dtype=acc)
    fn = getattr(lib(), f"msda_oracle_fwd_{variant}")
    rc = fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(aw),
            *[ctypes.c_int(x) for x in dims], _ptr(out))
    if rc != 0:
        raise RuntimeError(f"msda_oracle_fwd_{variant} failed rc={rc}")
    return out


def backward(grad_out, value, shapes, lsi, loc, aw, variant: str = "f32x"):
    """(grad_value[N,S,M,D], grad_loc[N,Lq,M,L,P,2], grad_aw[N,Lq,M,L,P]) in the carried dtype."""
    value, shapes, lsi, loc, aw, dims, st, acc = _prep(value, shapes, lsi, loc, aw, variant)
    N, S, M, D, L, Lq, P = dims
    grad_out = np.ascontiguousarray(grad_out, dtype=st).reshape(N, Lq, M * D)
    gv = np.empty((N, S, M, D), dtype=acc)
    gl = np.empty((N, Lq, M, L, P, 2), dtype=acc)
    ga = np.empty((N, Lq, M, L, P), dtype=acc)
    fn = getattr(lib(), f"msda_oracle_bwd_{variant}")
    rc = fn(_ptr(grad_out), _ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(aw),
            *[ctypes.c_int(x) for x in dims], _ptr(gv), _ptr(gl), _ptr(ga))
    if rc != 0:
        raise RuntimeError(f"msda_oracle_bwd_{variant} failed rc={rc}")
    return gv, gl, ga


def sample_index(shapes, loc, variant: str = "f32"):
    """(h_low int32, w_low int32, mask uint8), each [N,Lq,M,L,P].

    mask bit0..3 = corner v1..v4 inside the level, bit4 = sample passed the coarse test;
    rejected samples (outside, NaN, +-inf) report h_low = w_low = mask = 0."""
    st, _ = _VARIANTS[variant]
    loc = np.ascontiguousarray(loc, dtype=st)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    N, Lq, M, L, P, two = loc.shape
    assert two == 2 and shapes.shape == (L, 2)
    h = np.empty((N, Lq, M, L, P), dtype=np.int32)
    w = np.empty((N, Lq, M, L, P), dtype=np.int32)
    k = np.empty((N, Lq, M, L, P), dtype=np.uint8)
    fn = getattr(lib(), f"msda_oracle_index_{variant}")
    fn(_ptr(shapes), _ptr(loc), *[ctypes.c_int(x) for x in (N, Lq, M, L, P)],
       _ptr(h), _ptr(w), _ptr(k))
    return h, w, k


def set_threads(n: int) -> None:
    """Best effort: OpenMP thread count for subsequent calls."""
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(ctypes.c_int(int(n)))
    except OSError:
        os.environ["OMP_NUM_THREADS"] = str(int(n))
