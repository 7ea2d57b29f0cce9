"""TEST INFRASTRUCTURE ONLY -- vectorised numpy restatement of the reference CUDA arithmetic.

Written independently of oracle/msda_oracle.c so that the two can check each other: the
integer / predicate part (pixel floor, corner masks) is compared bit-for-bit, the sums with
a tolerance.  Follows /root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:
  cuh:285-288  pixel coordinate (ONE rounding -- the compiled FFMA)
               and the strict coarse test that also rejects NaN / +-inf
  cuh:38-78    floor, corner validity, zero padding
  cuh:80-83    bilinear weights         cuh:114-158  backward formulas
"""
from __future__ import annotations

import numpy as np

_CORNERS = ((0, 0, 1), (0, 1, 2), (1, 0, 4), (1, 1, 8))  # (dy, dx, mask bit) = v1..v4


def pixel_coords(loc: np.ndarray, shapes: np.ndarray):
    """h_im, w_im with the single-rounding schedule, same dtype as loc.

    float32: the product loc*size and the -0.5 are exact in float64 (24+11 bits), so one
    final rounding to float32 equals fmaf.  float64 input: plain (two-rounding) arithmetic;
    numpy has no fma -- documented limitation, only the f32 variants are index-exact here."""
    L = shapes.shape[0]
    H = shapes[:, 0].reshape(1, 1, 1, L, 1).astype(np.float64)
    W = shapes[:, 1].reshape(1, 1, 1, L, 1).astype(np.float64)
    x = loc[..., 0].astype(np.float64)
    y = loc[..., 1].astype(np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        h_im = (y * H - 0.5).astype(loc.dtype)
        w_im = (x * W - 0.5).astype(loc.dtype)
    return h_im, w_im


def sample_index(shapes: np.ndarray, loc: np.ndarray):
    """(h_low int32, w_low int32, mask uint8, lh, lw) -- lh/lw in loc.dtype (exact)."""
    shapes = np.asarray(shapes, dtype=np.int64)
    L = shapes.shape[0]
    Hi = shapes[:, 0].reshape(1, 1, 1, L, 1)
    Wi = shapes[:, 1].reshape(1, 1, 1, L, 1)
    h_im, w_im = pixel_coords(loc, shapes)
    with np.errstate(invalid="ignore"):
        inside = (h_im > -1) & (w_im > -1) & (h_im < Hi.astype(loc.dtype)) & (w_im < Wi.astype(loc.dtype))
    h_safe = np.where(inside, h_im, 0).astype(loc.dtype)
    w_safe = np.where(inside, w_im, 0).astype(loc.dtype)
    hf = np.floor(h_safe)
    wf = np.floor(w_safe)
    h_low = hf.astype(np.int32)
    w_low = wf.astype(np.int32)
    lh = (h_safe - hf).astype(loc.dtype)
    lw = (w_safe - wf).astype(loc.dtype)
    mask = np.zeros(h_low.shape, dtype=np.uint8)
    mask |= (((h_low >= 0) & (w_low >= 0)) * 1).astype(np.uint8)
    mask |= (((h_low >= 0) & (w_low + 1 <= Wi - 1)) * 2).astype(np.uint8)
    mask |= (((h_low + 1 <= Hi - 1) & (w_low >= 0)) * 4).astype(np.uint8)
    mask |= (((h_low + 1 <= Hi - 1) & (w_low + 1 <= Wi - 1)) * 8).astype(np.uint8)
    mask |= 16
    mask = np.where(inside, mask, 0).astype(np.uint8)
    h_low = np.where(inside, h_low, 0).astype(np.int32)
    w_low = np.where(inside, w_low, 0).astype(np.int32)
    return h_low, w_low, mask, lh, lw


def _corner_terms(value, shapes, lsi, loc, carried):
    """Yield per (l, p, corner): flat row index into value[N*S], validity, bilinear weight,
    and the two derivative coefficients (d/dh, d/dw of that corner's weight)."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    h_low, w_low, mask, lh, lw = sample_index(shapes, loc)
    lh = lh.astype(carried)
    lw = lw.astype(carried)
    hh = 1 - lh
    hw = 1 - lw
    b_idx = np.arange(N).reshape(N, 1, 1)
    for l in range(L):
        Hl, Wl = int(shapes[l, 0]), int(shapes[l, 1])
        for p in range(P):
            hl, wl_ = h_low[..., l, p], w_low[..., l, p]
            k = mask[..., l, p]
            a_lh, a_lw, a_hh, a_hw = lh[..., l, p], lw[..., l, p], hh[..., l, p], hw[..., l, p]
            weights = (a_hh * a_hw, a_hh * a_lw, a_lh * a_hw, a_lh * a_lw)
            dh = (-a_hw, -a_lw, a_hw, a_lw)     # cuh:123-151, grad_h_weight signs
            dw = (-a_hh, a_hh, -a_lh, a_lh)     # grad_w_weight signs
            for ci, (dy, dx, bit) in enumerate(_CORNERS):
                ok = (k & bit) != 0
                y = np.clip(hl + dy, 0, Hl - 1)
                x = np.clip(wl_ + dx, 0, Wl - 1)
                row = b_idx * S + int(lsi[l]) + y * Wl + x     # [N,Lq,M]
                yield l, p, ci, row, ok, weights[ci], dh[ci], dw[ci], Hl, Wl


def forward(value, shapes, lsi, loc, aw, carried=np.float64) -> np.ndarray:
    value = np.asarray(value)
    shapes = np.asarray(shapes, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    vflat = value.reshape(N * S, M, D).astype(carried)
    m_idx = np.arange(M).reshape(1, 1, M)
    out = np.zeros((N, Lq, M, D), dtype=carried)
    awc = np.asarray(aw).astype(carried)
    for l, p, ci, row, ok, wgt, _, _, _, _ in _corner_terms(value, shapes, lsi, loc, carried):
        v = vflat[row, m_idx]                                   # [N,Lq,M,D]
        v = np.where(ok[..., None], v, 0)
        out += (wgt * awc[..., l, p])[..., None] * v
    return out.reshape(N, Lq, M * D)


def backward(grad_out, value, shapes, lsi, loc, aw, carried=np.float64):
    value = np.asarray(value)
    shapes = np.asarray(shapes, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    g = np.asarray(grad_out).reshape(N, Lq, M, D).astype(carried)
    awc = np.asarray(aw).astype(carried)
    vflat = value.reshape(N * S, M, D).astype(carried)
    m_idx = np.broadcast_to(np.arange(M).reshape(1, 1, M), (N, Lq, M))
    gv = np.zeros((N * S, M, D), dtype=carried)
    gl = np.zeros((N, Lq, M, L, P, 2), dtype=carried)
    ga = np.zeros((N, Lq, M, L, P), dtype=carried)
    for l, p, ci, row, ok, wgt, dh, dw, Hl, Wl in _corner_terms(value, shapes, lsi, loc, carried):
        v = np.where(ok[..., None], vflat[row, m_idx], 0)       # [N,Lq,M,D]
        a = awc[..., l, p]
        gdotv = (g * v).sum(-1)                                 # sum_c g_c v_c
        ga[..., l, p] += wgt * gdotv
        gl[..., l, p, 0] += Wl * dw * a * gdotv
        gl[..., l, p, 1] += Hl * dh * a * gdotv
        contrib = np.where(ok[..., None], (wgt * a)[..., None] * g, 0)
        np.add.at(gv, (row[ok], m_idx[ok]), contrib[ok])
    return gv.reshape(N, S, M, D), gl, ga
